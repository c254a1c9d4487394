"""Which kernel family is sensitive to co-resident blocks (conv_own_sm = 0)?  Soak of one shape over option sets.
  python tools/soak3.py N "fuse_act=0,fuse_res=0" "..."     (streams = 3 and conv_own_sm = 0 are always set)"""
import importlib, os, sys, warnings, contextlib, io
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("voice-tts_b200"); synth = importlib.import_module("voice-tts_b200.synth"); cfg = importlib.import_module("voice-tts_b200.config")
N = int(sys.argv[1])
B = int(os.environ.get("SOAK_B", "4")); T0 = int(os.environ.get("SOAK_T0", "172"))
h = cfg.default_hparams(); sd = synth.make_state_dict(h, 1234)
def make(opts):
    m = pkg.BigVGAN(h, precision="bf16")
    with contextlib.redirect_stdout(io.StringIO()): m.remove_weight_norm()
    m.load_state_dict(sd); m = m.to("cuda:0").eval()
    for k, v in opts.items(): m.set_option(k, v)
    return m
mel = synth.make_mel(B, 80, T0).to("cuda:0")
for spec in sys.argv[2:]:
    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in spec.split(",") if "=" in kv}
    ser = dict(opts); ser.update(streams=1, conv_own_sm=1)
    base = make(ser)
    with torch.no_grad(): ref = base(mel).clone()
    del base
    par = dict(opts); par.setdefault("streams", 3); par.setdefault("conv_own_sm", 0)
    m = make(par)
    nbad = 0; info = []
    with torch.no_grad():
        for i in range(N):
            y = m(mel)
            if not torch.equal(y, ref):
                nbad += 1
                if len(info) < 4:
                    d = (y != ref); idx = d.nonzero()
                    info.append("iter %d: %d samples, utt %s, first %d last %d, max|err| %.2e" % (i, int(d.sum()), d.view(B, -1).sum(1).tolist(), idx[0][2], idx[-1][2], float((y - ref).abs().max())))
    print("%-50s %d of %d forwards differ %s" % (spec, nbad, N, " | ".join(info)), flush=True)
    del m; torch.cuda.empty_cache()
