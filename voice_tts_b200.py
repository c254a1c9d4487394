"""Importable alias: `import voice_tts_b200` loads the package directory
`voice-tts_b200/` (whose name, fixed by the repo layout, is not a Python
identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("voice-tts_b200")
sys.modules[__name__] = _pkg
