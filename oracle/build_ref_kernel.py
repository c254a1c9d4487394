"""ORACLE support - builds the REFERENCE's own fused activation kernel for sm_100a.

The reference ships one native component on this path: `anti_alias_activation_cuda` (pybind shim + one CUDA kernel,
indextts/s2mel/modules/bigvgan/alias_free_activation/cuda/{anti_alias_activation.cpp, anti_alias_activation_cuda.cu}),
JIT-built by its `load.py:17-65` for sm_70/sm_80 only - as shipped it has no image for B200.  This recipe compiles
the two source files WHERE THEY LIE under /root/reference (nothing is copied into the repo) with the reference's own flags
(`load.py:36-56`: -O3 --use_fast_math, the half-operator undefs) and `-gencode arch=compute_100a,code=sm_100a` instead of its
arch list, into `oracle/_ref/` (git-ignored, travels to the GPU box).  The result is
  * a GPU-side comparator: the kernel a user of the reference would run on B200 after fixing its arch list, timed beside
    ours by `bench.py --act-sweep` (series "reference_kernel");
  * a second pin of the activation semantics: `tests/test_gpu_ref_kernel.py` checks that it equals our kernel in the
    interior of every row and differs at the 3 samples next to each end, exactly as SURVEY 2.3 derived.
It is NOT a numerical oracle (fast-math sin, edge deviation) and is never on a product path.

    python oracle/build_ref_kernel.py          (needs /root/reference; about 5 minutes: torch headers)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BVG_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "indextts/s2mel/modules/bigvgan/alias_free_activation/cuda")
OUT = os.path.join(ROOT, "oracle", "_ref")
NAME = "anti_alias_activation_cuda"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(SRC):
        return None
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"       # only consulted for torch's default flags; the real arch is below
    from torch.utils import cpp_extension
    os.makedirs(OUT, exist_ok=True)
    cpp_extension.load(
        name=NAME,
        sources=[os.path.join(SRC, "anti_alias_activation.cpp"), os.path.join(SRC, "anti_alias_activation_cuda.cu")],
        build_directory=OUT,
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a", "--use_fast_math",
                           "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                           "--expt-relaxed-constexpr", "--expt-extended-lambda"],
        is_python_module=False, verbose=verbose)
    return so_path() if os.path.exists(so_path()) else None


def load():
    """imports the built module (GPU box: the prebuilt .so only) - returns None when it is not there."""
    if not os.path.exists(so_path()):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location(NAME, so_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print(p or "reference sources not found - nothing built")
