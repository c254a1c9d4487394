"""Does a CTA of another kernel sharing the SM corrupt the persistent tcgen05 conv kernel?
A filler kernel (FMA spin or memory streaming, 256 threads, no shared memory) runs on a second stream while a wide
Conv1d (+residual) runs on the first; the conv result is compared with the result of an undisturbed run."""
import importlib, os, sys, warnings, ctypes
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ops = importlib.import_module("voice-tts_b200.ops"); _lib = importlib.import_module("voice-tts_b200._lib")
lib = _lib.load()
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
bf = lambda t: t.to(torch.bfloat16).float()
cases = [(4, 384, 3444, 7, 1), (4, 96, 13776, 7, 3), (8, 24, 55104, 11, 1), (2, 768, 3444, 3, 1)]
scratch = torch.zeros(64 << 20, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for (B, C, T, k, d) in cases:
    x = bf(torch.randn(B, C, T, generator=g)).to(dev); w = bf(torch.randn(C, C, k, generator=g) / (C * k) ** .5).to(dev)
    b = torch.randn(C, generator=g).to(dev); res = torch.randn(B, C, T, generator=g).to(dev); e = torch.empty(0, device=dev)
    ref = ops.conv1d_res(x, w, b, res, e, 1.0, False, d, "bf16", 0); torch.cuda.synchronize()
    for mode, name in ((None, "undisturbed"), (0, "FMA filler"), (1, "memory filler")):
        for threads in (128, 256):
            nd = 0
            for trial in range(6):
                if mode is not None:
                    with torch.cuda.stream(s2):
                        rc = lib.bvg_debug_spin(148 * 8, threads, 20000 if mode == 0 else 2000, mode, scratch.data_ptr(), scratch.numel(), s2.cuda_stream)
                        assert rc == 0
                with torch.cuda.stream(s1):
                    y = ops.conv1d_res(x, w, b, res, e, 1.0, False, d, "bf16", 0)
                torch.cuda.synchronize()
                nd += int((y != ref).sum())
            print("conv %s  %-16s filler threads %3d: differing elements %d" % ((B, C, T, k, d), name, threads, nd), flush=True)
            if mode is None: break
