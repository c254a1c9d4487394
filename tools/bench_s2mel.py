"""The s2mel-tail extra of bench.py on its own:  python tools/bench_s2mel.py [T] [B]  -> one JSON line"""
import json, os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1863
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
print(json.dumps(bench.s2mel_tail_bench(T, B)))
