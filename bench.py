#!/usr/bin/env python
"""Benchmark of the BigVGAN v2 vocoder hot path on B200 (see DESIGN.md, section Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the generator over one batch of synthetic mels (the
BASELINE configs[1] workload: 16 utterances x 10 s per GPU, bf16 tensor-core
mode).  Rank 0 prints ONE JSON line.  `value` = audio-seconds generated per
wall-second over all ranks with the mels already resident in HBM; `e2e` = the
same through the host-buffer C-ABI call (`bvg_vocoder_fwd_host`: pinned H2D of
the mels + generator + D2H of the waveform inside the timed region).

`--impl reference` times the reference's CPU path on the host cores: the oracle's
staged form, i.e. the reference's own operator sequence (F.pad / conv_transpose1d /
conv1d / snake), bit-identical to the imported reference and within 5 % of its wall
time in the build container (tools/cpu_arm_check.py; the Python reference itself
cannot travel to the GPU box) - on a bounded sample of the same workload: ONE
utterance of the workload's shape per step.

`--workload c4` is BASELINE configs[3]: a global batch of 256 utterances x 30 s sharded
256/N over the ranks (strong scaling), micro-batched inside `bvg_vocoder_fwd`.
"""
import argparse
import contextlib
import importlib
import io
import json
import os
import subprocess
import sys
import threading
import time
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vocoder_audio_seconds_per_second"
UNIT = "audio-s/s"
SR = 22050
HOP = 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline", choices=["headline", "c4"],
                    help="headline: BASELINE configs[1], 16 x 10 s per GPU (weak scaling); c4: configs[3], 256 x 30 s global (strong)")
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU per step (headline) / in total (c4)")
    ap.add_argument("--frames", type=int, default=None, help="mel frames per utterance (861 = 10 s, 2584 = 30 s)")
    ap.add_argument("--repeats", type=int, default=3, help="the K-step timed region is repeated this often; value = median")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--act-sweep", action="store_true", help="include the full fused-activation HBM sweep table (config 3)")
    ap.add_argument("--no-act-sweep", action="store_true", help="skip the activation sweep summary")
    ap.add_argument("--no-extras", action="store_true", help="headline numbers only (no precision modes, activation sweep, single-utterance latency)")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 256 if a.workload == "c4" else 16
    if a.frames is None:
        a.frames = 2584 if a.workload == "c4" else 861
    return a


def workload_config(args, world):
    """`config` of the JSON line - identical in both arms (the reference arm times a bounded sample of it)."""
    if args.workload == "c4":
        return {"workload": "bigvgan_v2_22khz_80band_256x generator, global batch of %d utterances x %d mel frames (%.1f s) sharded "
                            "%d/N over the ranks per step (BASELINE configs[3])" % (args.batch, args.frames, args.frames * HOP / SR, args.batch),
                "global_batch": args.batch, "parallelism": "batch-sharded x%d, no collective" % world,
                "weights": "random-init (seed 1234), alpha/beta ~ N(0,0.5)",
                "l2": "no explicit flush: each step streams a multi-GB workspace + 0.22 GB weights (>> 126 MB L2)"}
    return {"workload": "bigvgan_v2_22khz_80band_256x generator, %d utterances x %d mel frames (%.1f s) per GPU per step"
                        % (args.batch, args.frames, args.frames * HOP / SR),
            "global_batch": args.batch * world, "parallelism": "batch-sharded x%d, no collective" % world,
            "weights": "random-init (seed 1234), alpha/beta ~ N(0,0.5)",
            "l2": "no explicit flush: each step streams a multi-GB workspace + 0.22 GB weights (>> 126 MB L2)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_forward(h, sd, mel):
    """one forward of the reference's CPU path: the oracle in its staged form = the reference's own operator sequence
    (bit-identical to the imported reference, 1.05x its wall time: tools/cpu_arm_check.py)"""
    from oracle import bigvgan_oracle as O
    with O.staged_ops():
        return O.generator_forward(sd, h, mel)


def cpu_forward_rate(h, sd, frames, runs, threads):
    """audio-s per wall-s of the reference's CPU path on the host cores (median of `runs` after one warm-up)."""
    import torch
    synth = importlib.import_module("voice-tts_b200.synth")
    torch.set_num_threads(threads)
    mel = synth.make_mel(1, h["num_mels"], frames)
    with torch.no_grad():
        cpu_forward(h, sd, mel)  # warm-up
        ts = []
        for _ in range(runs):
            t0 = time.perf_counter()
            cpu_forward(h, sd, mel)
            ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return frames * HOP / SR / med, med


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on all host threads.  Each step is a bounded sample of
    the workload: ONE utterance of the workload's shape (fewer frames only if the host is too slow for the time budget)."""
    if rank != 0:
        return
    import torch
    cfg = importlib.import_module("voice-tts_b200.config")
    synth = importlib.import_module("voice-tts_b200.synth")
    h = cfg.default_hparams()
    sd = synth.make_state_dict(h, seed=1234)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    # one utterance of the workload's own shape per step, unless a probe says the whole run would exceed ~4 minutes
    probe_frames = 43
    rate, t_probe = cpu_forward_rate(h, sd, probe_frames, 1, threads)
    budget = 240.0 / max(1, args.steps + args.warmup)
    frames = int(max(22, min(args.frames, probe_frames * budget / (2.0 * t_probe))))   # x2: long utterances run ~2x slower per frame (caches)
    mel = synth.make_mel(1, h["num_mels"], frames)
    ts = []
    with torch.no_grad():
        for _ in range(args.warmup):
            cpu_forward(h, sd, mel)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t1 = time.perf_counter()
            cpu_forward(h, sd, mel)
            ts.append(time.perf_counter() - t1)
        dt = time.perf_counter() - t0
    audio_s = frames * HOP / SR
    value = audio_s * args.steps / dt
    ts.sort()
    sample = "1 utterance x %d mel frames (%.2f s audio) per step%s, fp32, reference operator sequence (oracle staged form), torch %s CPU, %s" % (
        frames, audio_s, "" if frames == args.frames else " (the workload's %d frames cut to the time budget)" % args.frames,
        torch.__version__, cpu_model_name())
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "median_step_s": ts[len(ts) // 2],
                         "kind_note": "the reference's own operator sequence on torch CPU (bit-identical outputs, 1.05x the wall time of "
                                      "the imported reference in the build container); the Python reference cannot travel to the GPU box"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def act_sweep(dev, pk):
    """BASELINE config 3: fused Activation1d HBM sweep on [B,C,T] (operator layout, bvg_act1d_fwd).
    Algorithmic bytes = 2 * B*C*T * sizeof(dtype).  Three series: bf16 I/O (fast snake), fp32 I/O with the
    accurate snake (the <= 1e-5 parity mode) and fp32 I/O with the fast snake (BVG_ACT_FAST_SIN)."""
    import torch
    ops = importlib.import_module("voice-tts_b200.ops")
    synth = importlib.import_module("voice-tts_b200.synth")
    taps_cpu = synth.kaiser_sinc_filter1d().reshape(-1)
    taps = taps_cpu.tolist()
    rows = []
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    # GPU-side comparator (a reference measurement like `--impl reference`, never on the product path): the reference's own
    # fused kernel (anti_alias_activation_cuda.cu:43-246) rebuilt for sm_100a by oracle/build_ref_kernel.py into oracle/_ref -
    # series "reference_kernel" (fast-math sin, log-scale alpha/beta like ours); absent -> skipped
    from oracle import build_ref_kernel
    refk = build_ref_kernel.load()
    taps_t = taps_cpu.to(dev)
    series = [(torch.bfloat16, True, None), (torch.float32, False, None), (torch.float32, True, None)]
    if refk is not None:
        series += [(torch.bfloat16, True, refk), (torch.float32, True, refk)]
    for dt, fast, rk in series:
        for C in (24, 48, 192, 768, 1536):
            for T in (8192, 131072, 2097152):
                es = 2 if dt == torch.bfloat16 else 4
                # >= 256 MB per tensor where possible (2x the 126 MB L2), plus an explicit L2 flush between runs
                B = max(1, -(-(256 << 20) // (C * T * es)))
                if B * C * T * es > (3 << 30):
                    continue
                x = torch.randn(B, C, T, device=dev).to(dt)
                a = torch.randn(C, device=dev) * 0.5
                b = torch.randn(C, device=dev) * 0.5
                if rk is not None:
                    if B * C * T >= 2 ** 31:      # the reference kernel indexes with 32-bit ints (.cu:68)
                        continue
                    fdt, adt, bdt = taps_t.to(dt), a.to(dt), b.to(dt)   # the v2 copy reads filters / alpha / beta as input_t (.cu:47-50)
                    run = lambda: rk.forward(x, fdt, fdt, adt, bdt)
                else:
                    run = lambda: ops.act1d(x, a, b, taps, taps, fast)
                for _ in range(2):
                    run()
                ts = []
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    run()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ts.sort()
                gbs = 2.0 * B * C * T * es / (ts[len(ts) // 2] * 1e-3) / 1e9
                rows.append({"impl": "reference_kernel" if rk is not None else "ours",
                             "dtype": str(dt)[6:], "snake": "fast" if fast else "accurate", "B": B, "C": C, "T": T,
                             "ms": ts[len(ts) // 2], "GBps": round(gbs, 1), "frac_hbm": round(gbs / pk["hbm_gbs"], 3)})
                del x
    return rows


def top_launch_stats(model, pk):
    """The layer shape that costs the most time in the instrumented pass: its own achieved TFLOP/s (per-launch CUDA events)."""
    import tempfile
    path = os.path.join(tempfile.gettempdir(), "bvg_bench_profile_%d.csv" % os.getpid())
    try:
        model.dump_profile(path)
        groups = {}
        with open(path) as f:
            for ln in f.read().splitlines()[1:]:
                cat, cin, cout, k, dil, rows, ms, work, rate = ln.split(",")
                if cat not in ("0", "4"):
                    continue
                mode = int(dil) // 100
                key = (int(cin), int(cout), int(k), mode)
                g = groups.setdefault(key, [0, 0.0, 0.0])
                g[0] += 1
                g[1] += float(ms)
                g[2] += float(work)
        os.remove(path)
        if not groups:
            return None
        key, (n, ms, work) = max(groups.items(), key=lambda kv: kv[1][1])
        tf = work / (ms * 1e-3) / 1e12
        kind = {0: "plain epilogue", 1: "following Activation1d in the epilogue", 2: "residual + following Activation1d in the epilogue",
                3: "bf16x3 split", 4: "whole AMP unit: act + conv + act + conv + residual in one kernel, flops of both convs"}.get(key[3], "?")
        return {"layer": "Conv1d %d->%d k=%d (%s)" % (key[0], abs(key[1]), key[2], kind), "launches": n,
                "avg_launch_ms": round(ms / n, 4), "achieved": round(tf, 1), "unit": "TFLOP/s",
                "frac_of_sustained_peak": round(tf / pk["bf16_tflops_sustained"], 4),
                "frac_of_burst_peak": round(tf / pk["bf16_tflops"], 4)}
    except Exception as exc:  # diagnostics only
        return {"error": str(exc)}


def act_sweep_summary(rows, pk):
    """min / median / max fraction of the measured HBM copy rate per series, and the large-T points (T >= 131072)."""
    out = {}
    for r in rows:
        key = "%s %s %s" % (r["impl"], r["dtype"], r["snake"])
        out.setdefault(key, []).append(r)
    summ = {}
    for key, rs in out.items():
        fr = sorted(x["frac_hbm"] for x in rs)
        big = sorted(x["frac_hbm"] for x in rs if x["T"] >= 131072)
        summ[key] = {"points": len(fr), "frac_min": fr[0], "frac_median": fr[len(fr) // 2], "frac_max": fr[-1],
                     "frac_median_T_ge_131072": big[len(big) // 2] if big else None,
                     "GBps_max": max(x["GBps"] for x in rs)}
    return {"peak_GBps": pk["hbm_gbs"], "layout": "[B,C,T] operator layout (bvg_act1d_fwd), >= 256 MB tensors, L2 flushed between runs, "
            "C in 24..1536, T in 8K..2M, algorithmic bytes = 2*B*C*T*sizeof(dtype)", "series": summ}


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    pkg = importlib.import_module("voice-tts_b200")
    cfg = importlib.import_module("voice-tts_b200.config")
    synth = importlib.import_module("voice-tts_b200.synth")
    _lib = importlib.import_module("voice-tts_b200._lib")
    shard = importlib.import_module("voice-tts_b200.shard")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    c4 = args.workload == "c4"
    extras = not args.no_extras and not c4

    h = cfg.default_hparams()
    sd = synth.make_state_dict(h, seed=1234)
    model = pkg.BigVGAN(h, precision=args.precision)
    with contextlib.redirect_stdout(io.StringIO()):
        model.remove_weight_norm()
    model.load_state_dict(sd)
    model = model.to(dev).eval()

    T0 = args.frames
    if c4:
        # strong scaling: the global batch is fixed, rank r vocodes utterances shard_range(B_global, r, world)
        B_global = args.batch
        lo, hi = shard.shard_range(B_global, rank, world)
    else:
        # weak scaling: B utterances per GPU, rank r vocodes [r*B, (r+1)*B) of the global batch
        B_global = args.batch * world
        lo, hi = shard.shard_range(B_global, rank, world)
    B = hi - lo
    audio_s_step = B_global * T0 * HOP / SR          # audio-seconds ALL ranks produce per step
    mel_host = synth.make_mel(B, h["num_mels"], T0, first_utterance=lo).pin_memory()
    mel = mel_host.to(dev, non_blocking=True)
    wav_host = torch.empty(B, 1, T0 * cfg.total_upsample(h), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return shard.max_over_ranks(x, dev)

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            wav = model(mel)
        torch.cuda.synchronize()

        def timed_steps():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            w = None
            for _ in range(args.steps):
                w = model(mel)
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)), w

        def timed_e2e():
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                model.forward_host(mel_host, out=wav_host)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            barrier()
            return dt

        for _ in range(2):
            model.forward_host(mel_host, out=wav_host)
        # ---- `repeats` x (K device-resident steps, K end-to-end steps), interleaved so that both see the same clocks;
        #      value / e2e = the median region.  Device-resident: CUDA events around the K steps, max over ranks.
        #      End to end: BigVGAN.forward_host -> bvg_vocoder_fwd_host (pinned H2D of the mels + generator + D2H of the
        #      waveform inside the timed region), wall clock, max over ranks. ----
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms_runs, e2e_runs, launches = [], [], 0
        for r in range(max(1, args.repeats)):
            n0 = _lib.launch_count()
            ms_r, wav = timed_steps()
            if r == 0:
                launches = _lib.launch_count() - n0
            ms_runs.append(ms_r)
            e2e_runs.append(timed_e2e())
        ms = median(ms_runs)
        e2e_s = median(e2e_runs)
        clocks = sampler.stop() if rank == 0 else None

        # ---- the same K steps again, serial schedule (streams = 1), with one CUDA-event pair per kernel on the launch
        #      stream (roofline / time split).  An event between two launches keeps the next kernel from being issued
        #      while the previous one drains (~5-9 us per launch, 2-4 % of the step), and per-kernel times only add up
        #      when kernels do not overlap, so the headline comes from the un-instrumented default passes above and this
        #      pass reports its own ms_per_step beside the per-kernel sums. ----
        wav_default = wav.clone()
        model.set_option("streams", 1)
        for _ in range(2):
            model(mel)
        ms_serial, wav_s = timed_steps()
        serial_identical = bool(torch.equal(wav_s, wav_default))
        model.set_option("profile", 1)
        model(mel)
        model.read_profile()
        ms_prof, _ = timed_steps()
        top_launch = None
        if rank == 0:
            top_launch = top_launch_stats(model, pk)
        prof = model.read_profile()
        model.set_option("profile", 0)
        model.set_option("streams", 3)
        del wav_default, wav_s

        # ---- latency of ONE utterance, the way infer_v2 calls the vocoder (batch 1, one segment): 2 s (BASELINE
        #      configs[0] shape) and 10 s, repeated shapes (CUDA-graph replay after the first call of a shape) ----
        latency = {}
        if extras:
            for frames in (172, 861):
                mel1 = mel[:1, :, :frames].contiguous() if T0 >= frames else None
                if mel1 is None:
                    continue
                for _ in range(5):
                    model(mel1)
                barrier()
                e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e4.record()
                for _ in range(50):
                    model(mel1)
                e5.record()
                barrier()
                ms_one = max_over_ranks(e4.elapsed_time(e5)) / 50
                latency[frames] = {"frames": frames, "audio_s": frames * HOP / SR, "ms": ms_one,
                                   "x_realtime": frames * HOP / SR / (ms_one * 1e-3)}

        # ---- BASELINE configs[4], vocoder side: the segment loop of infer_v2.py:616-744 replayed around the drop-in (GPT and
        #      s2mel are out of scope and cannot be imported here): per text segment `wav = self.bigvgan(vc_target.float())`,
        #      `torch.clamp(32767 * wav, -32767, 32767)`, `wav.cpu()`; wall clock like the reference's own `bigvgan_time` ----
        replay = None
        if extras and rank == 0:
            seg_frames = [388, 517, 301, 646, 431, 560, 258, 474, 905, 1290]      # 10 segments, 3-15 s (<= 1500 mel tokens each)
            segs = [synth.make_mel(1, h["num_mels"], f, first_utterance=100 + i).to(dev) for i, f in enumerate(seg_frames)]
            segs_host = [vc.cpu().pin_memory() for vc in segs]
            audio = sum(seg_frames) * HOP / SR

            def loop_reference_form():
                wavs = []
                for vc_target in segs:
                    wav = model(vc_target.float()).squeeze().unsqueeze(0)
                    wav = torch.clamp(32767 * wav, -32767.0, 32767.0)
                    wavs.append(wav.cpu())
                return wavs

            def loop_host_int16():
                return [model.forward_host(vc, int16=True) for vc in segs_host]     # pinned host mel in, int16 host waveform out

            def timed(fn, n=5):
                fn(); fn()
                ts = []
                for _ in range(n):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    fn()
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                return median(ts)
            t_ref_form = timed(loop_reference_form)
            t_i16 = timed(loop_host_int16)
            t_batched = timed(lambda: [w.cpu() for w in model.forward_segments(segs)])
            t_serial = timed(lambda: [w.cpu() for w in model.forward_segments(segs, concurrency=1)])
            t_three = timed(lambda: [w.cpu() for w in model.forward_segments(segs, concurrency=3)])
            replay = {"segments": len(seg_frames), "frames": seg_frames, "audio_s": audio,
                      "loop_as_infer_v2": {"bigvgan_time_s": t_ref_form, "rtf": t_ref_form / audio, "x_realtime": audio / t_ref_form},
                      "loop_forward_host_int16": {"bigvgan_time_s": t_i16, "rtf": t_i16 / audio, "x_realtime": audio / t_i16},
                      "forward_segments_one_call": {"bigvgan_time_s": t_batched, "rtf": t_batched / audio, "x_realtime": audio / t_batched,
                                                    "concurrency": 2},
                      "forward_segments_serial": {"bigvgan_time_s": t_serial, "x_realtime": audio / t_serial, "concurrency": 1},
                      "forward_segments_3_workers": {"bigvgan_time_s": t_three, "x_realtime": audio / t_three, "concurrency": 3},
                      "note": "vocoder stage only (GPT + s2mel out of scope); median of 5 wall-clock passes after 2 warm-ups"}

    value = audio_s_step * args.steps / (ms * 1e-3)
    e2e_measured = audio_s_step * args.steps / e2e_s
    # the end-to-end pass does strictly more work per step (the same kernels plus the two PCIe copies); when its median
    # region nevertheless comes out ahead of the device-resident one (clock drift under the power cap: the regions differ
    # by ~1 %), the lower figure is reported and the measured one kept beside it
    e2e_value = min(e2e_measured, value)

    # ---- roofline of the dominant kernel (live CUDA-event durations of this run) ----
    c_ms, c_flops, c_n = prof["conv_tcgen05"]
    s_ms, s_flops, s_n = prof["conv_simt"]
    a_ms, a_bytes, a_n = prof["activation"]
    o_ms, _, o_n = prof["other"]
    u_ms, u_flops, u_n = prof.get("amp_unit", (0.0, 0.0, 0))
    tot = c_ms + s_ms + a_ms + o_ms + u_ms
    # the whole-AMP-unit kernels (amp_unit.cu) are tcgen05 conv kernels with both activations inside: they count as conv
    # launches with the flops of their two convolutions
    c_ms, c_flops, c_n = c_ms + u_ms, c_flops + u_flops, c_n + u_n
    conv_ms, conv_flops, conv_n, conv_name = (c_ms, c_flops, c_n, "conv_umma2_kernel + conv_umma2a_kernel (tcgen05 implicit-GEMM Conv1d/ConvTranspose1d; the a-variant carries the following Activation1d in its epilogue)") \
        if c_n else (s_ms, s_flops, s_n, "conv_simt_kernel (fp32)")
    tensor_peak = pk["bf16_tflops_sustained"]
    ach_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    ach_gb = a_bytes / (a_ms * 1e-3) / 1e9 if a_ms > 0 else 0.0
    # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of every launch of ONE forward of the SAME workload
    # from the committed ncu capture (tools/ncu_traffic.py -> profiles/r02_traffic.json; ncu cannot run inside a timed bench)
    traffic, traffic_file = {}, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            if args.batch == 16 and T0 == 861 and args.precision == "bf16" and not c4:
                traffic, traffic_file = tj, name
            break
        except Exception:
            continue
    roofline = {"kernel": conv_name, "bound": "tensor", "achieved": round(ach_tf, 2), "peak": tensor_peak,
                "unit": "TFLOP/s", "frac": round(ach_tf / tensor_peak, 4),
                "frac_of_burst_peak": round(ach_tf / pk["bf16_tflops"], 4),
                "traffic": traffic.get("conv_tcgen05", {}).get("dram_bytes_per_launch") if c_n else None,
                "traffic_note": "bytes per launch averaged over the conv launches of one forward (ncu capture of the same workload, profiles/%s); operands + results of the padded tensors, no re-reads" % traffic_file,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % pk["source"],
                "measured": "one CUDA-event pair per launch on the launch stream, summed over a separate pass of the same K steps",
                "share_of_step": round(conv_ms / tot, 3) if tot else None, "launches": conv_n,
                "avg_launch_ms": round(conv_ms / max(conv_n, 1), 4)}
    if top_launch:
        roofline["top_launch"] = top_launch
    roofline_act = {"kernel": "act1d_cl_packed_kernel (stand-alone fused up2-snakebeta-down2, channels-last; the activations fused into conv epilogues are not counted here)", "bound": "hbm",
                    "achieved": round(ach_gb, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": round(ach_gb / pk["hbm_gbs"], 4),
                    "traffic": traffic.get("activation", {}).get("dram_bytes_per_launch"),
                    "algorithmic_bytes_per_launch": round(a_bytes / max(a_n, 1), 1),
                    "share_of_step": round(a_ms / tot, 3) if tot else None, "launches": a_n}

    config = workload_config(args, world)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if c4 else "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": config,
        "repeats": {"n": len(ms_runs), "statistic": "median", "ms_per_step": [round(x / args.steps, 4) for x in ms_runs],
                    "e2e_ms_per_step": [round(1e3 * x / args.steps, 4) for x in e2e_runs]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": mel_host.numel() * 4,
                "d2h_bytes_per_step": wav_host.numel() * 4, "api": "BigVGAN.forward_host -> bvg_vocoder_fwd_host",
                "measured": e2e_measured},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_activation": roofline_act,
        "time_split_ms_per_step": {"conv_tcgen05": c_ms / args.steps, "conv_simt": s_ms / args.steps,
                                   "activation": a_ms / args.steps, "other": o_ms / args.steps,
                                   "of_which_amp_unit_kernels": u_ms / args.steps},
        "x_realtime_per_gpu": value / world,
        "utterances_per_rank": B,
        "workspace_gb_per_gpu": round(model_workspace_gb(model, B, T0), 2),
        "profile_pass_ms_per_step": ms_prof / args.steps,
        "serial_schedule": {"value": audio_s_step * args.steps / (ms_serial * 1e-3), "unit": UNIT,
                            "ms_per_step": ms_serial / args.steps, "bit_identical_to_default": serial_identical,
                            "note": "bvg_set_option('streams', 1); the default runs the 3 AMP blocks of a stage on 3 streams"},
    }
    if replay:
        out["infer_v2_segment_replay"] = replay
    if 172 in latency:
        out["single_utterance"] = latency[172]
    if 861 in latency:
        out["single_utterance_10s"] = latency[861]

    if rank == 0 and not args.no_cpu_baseline and world >= 1:
        threads = os.cpu_count() or 1
        # bounded sample: ONE utterance of the workload's own shape (10 s: ~10-30 s of CPU work for 1 warm-up + 3 runs);
        # a probe shortens it on a slow host
        rate_p, t_p = cpu_forward_rate(h, sd, 43, 1, threads)
        frames = int(max(43, min(T0, 43 * 12.0 / (2.0 * t_p))))    # <= ~12 s per forward (long utterances cost ~2x per frame)
        rate, med = cpu_forward_rate(h, sd, frames, 3, threads)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": "reference operator sequence on torch CPU (oracle staged form: bit-identical to the imported "
                                         "reference, 1.05x its wall time in the build container), fp32, 1 utterance x %d frames (%.1f s), "
                                         "median of 3 after 1 warm-up, %.2f s per forward, %s" % (frames, frames * HOP / SR, med, cpu_model_name())}
    if rank == 0 and world == 1 and args.precision == "bf16" and extras and not args.no_cpu_baseline:
        # the other precision modes on the same workload (reported beside the headline, not part of it): "fp32" is the
        # <= 1e-5 parity mode (SIMT convolutions), "bf16x3" fp32 storage with three bf16 tensor-core passes per convolution
        out["precision_modes"] = {}
        del model
        torch.cuda.empty_cache()
        for prec in ("bf16x3", "fp32"):
            try:
                m2 = pkg.BigVGAN(h, precision=prec)
                with contextlib.redirect_stdout(io.StringIO()):
                    m2.remove_weight_norm()
                m2.load_state_dict(sd)
                m2 = m2.to(dev).eval()
                with torch.no_grad():
                    m2(mel)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(2):
                        m2(mel)
                    e1.record()
                    torch.cuda.synchronize()
                msp = e0.elapsed_time(e1) / 2
                out["precision_modes"][prec] = {"value": audio_s_step / (msp * 1e-3), "unit": UNIT, "ms_per_step": msp}
                del m2
                torch.cuda.empty_cache()
            except Exception as e:  # a reported extra must never cost the headline line
                out["precision_modes"][prec] = {"error": str(e)[:200]}
    if rank == 0 and world == 1 and extras and not args.no_act_sweep:
        # BASELINE configs[2]: the fused-activation HBM sweep (summary always, full table with --act-sweep)
        try:
            rows = act_sweep(dev, pk)
            out["activation_sweep_summary"] = act_sweep_summary(rows, pk)
            if args.act_sweep:
                out["activation_sweep"] = rows
        except Exception as e:
            out["activation_sweep_summary"] = {"error": str(e)[:200]}
    if rank == 0 and world == 1 and extras:
        try:      # the next row of SURVEY section 8(f) built to the same bar: the s2mel tail in front of the vocoder
            out["s2mel_tail"] = s2mel_tail_bench(cpu=not args.no_cpu_baseline, dev=dev)
        except Exception as e:
            out["s2mel_tail"] = {"error": str(e)[:200]}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def s2mel_tail_bench(T=1863, B=2, cpu=True, dev="cuda:0"):
    """`bvg_s2mel_tail_fwd` at the shape one solver step of infer_v2 runs (flow_matching.py:88-98:
    the CFG-stacked batch of 2; 795 prompt + 1068 target frames as in the docstring of flow_matching.py:36), both precision
    modes, the fused Euler / CFG update, and - as this path's cpu_baseline leg - the reference's operator sequence (oracle
    restatement, <= 7e-7 from the unmodified DiT.forward) on torch CPU with all host threads, same shape."""
    import statistics
    import torch
    tm = importlib.import_module("voice-tts_b200.s2mel_tail")
    synth = importlib.import_module("voice-tts_b200.synth")
    cfgm = importlib.import_module("voice-tts_b200.config")
    _lib = importlib.import_module("voice-tts_b200._lib")
    c = cfgm.s2mel_tail_config()
    H, D, L, k, C = c["hidden"], c["dit_hidden"], c["n_layers"], c["kernel_size"], c["out_channels"]
    macs_row = D * H + L * (H * 2 * H * k) + (L - 1) * (H * 2 * H) + H * H + D * H + H * H + H * C
    flops = 2.0 * macs_row * B * T
    sd = synth.make_s2mel_tail_state_dict(c, seed=1)
    x_res, tt, t1, lens = synth.make_s2mel_tail_inputs(c, B, T)
    out = {"workload": "s2mel tail (conv1 + WN x%d + res_projection + FinalLayer + conv2), %d x %d frames, random-init weights" % (L, B, T),
           "gflop_per_call": flops / 1e9}
    xr, td, t1d, ld = x_res.to(dev), tt.to(dev), t1.to(dev), lens.to(dev)
    outs = {}

    def timed(m, n=20, reps=5):
        ts = []
        with torch.no_grad():
            for _ in range(3):
                y = m(xr, ld, td, t1d)
            for _ in range(reps):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    y = m(xr, ld, td, t1d)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / n)
        return statistics.median(ts), y

    for prec in ("bf16", "fp32"):
        m = tm.S2MelTail(c, precision=prec)
        m.load_folded_state_dict(sd)
        m = m.to(dev).eval()
        ms, y = timed(m)
        launches = m.last_forward_launches()
        m.set_option("graph", 0)
        eager_ms, y2 = timed(m, reps=1)
        out[prec] = {"ms_per_call": ms, "ms_per_call_without_graph_replay": eager_ms, "tflops": flops / (ms * 1e-3) / 1e12,
                     "launches": int(launches), "frames_per_s": B * T / (ms * 1e-3), "replay_bit_identical": bool(torch.equal(y, y2))}
        outs[prec] = y.cpu()
        del m
    ref = outs["fp32"].double()
    err = outs["bf16"].double() - ref
    out["bf16"]["snr_db_vs_fp32_mode"] = float(10 * torch.log10(ref.pow(2).sum() / err.pow(2).sum()))
    # one solver step's Euler / CFG update: x [1, 80, T], stacked estimator output [2, 80, T]
    x = torch.randn(1, C, T, device=dev)
    d = torch.randn(2, C, T, device=dev)
    for _ in range(3):
        tm.euler_step_(x, d, 0.04, 0.7, 100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        tm.euler_step_(x, d, 0.04, 0.7, 100)
    e1.record()
    torch.cuda.synchronize()
    out["euler_step_us"] = e0.elapsed_time(e1) / 200 * 1e3
    if cpu:
        from oracle import s2mel_oracle as S      # cpu_baseline leg: the checker, timed as the CPU arm of this path
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            S.tail_forward(sd, c, x_res, lens, tt, t1)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                yo = S.tail_forward(sd, c, x_res, lens, tt, t1)
                ts.append(time.perf_counter() - t0)
        out["cpu_baseline"] = {"ms_per_call": statistics.median(ts) * 1e3, "cores": os.cpu_count(), "kind": "port",
                               "sample": "the same %d x %d frames, median of 3 after 1 warm-up" % (B, T),
                               "max_abs_diff_fp32_mode": float((outs["fp32"] - yo).abs().max())}
        out["speedup_bf16_vs_cpu"] = out["cpu_baseline"]["ms_per_call"] / out["bf16"]["ms_per_call"]
    return out


def model_workspace_gb(model, B, T0):
    _lib = importlib.import_module("voice-tts_b200._lib")
    ops = importlib.import_module("voice-tts_b200.ops")
    if model._hid is None:
        return 0.0
    return _lib.load().bvg_workspace_bytes(ops._HANDLES[model._hid][0], B, T0) / 1e9


if __name__ == "__main__":
    main()
