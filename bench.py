#!/usr/bin/env python
"""bench.py - TRU-Net training-step throughput (BASELINE.json: "train 4-s clips/sec") plus the inference real-time factors.

    python bench.py --gpus N --steps K --warmup W             # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU (oracle port)
    python bench.py --impl torch-gpu  --steps K ...           # the reference algorithm on stock PyTorch on the same GPU
                                                              # (cuDNN / cuBLAS / cuFFT): the bar the hand kernels must beat

A "step" = one pass of the hot path over one batch of synthetic 16 kHz clips:
front end -> TRU-Net -> mask + iSTFT -> L1 + multi-resolution STFT loss -> backward ->
(N > 1: one NCCL all-reduce of the flat gradient bucket) -> grad norm + LR schedule + AdamW.
Workload = BASELINE.json configs[1]: tiny.json, 32 distinct clean/noisy 4-s pairs per GPU (weak scaling: global batch 32*N;
for N > 1 the strong-scaling point of configs[2], global batch 256, is timed as well).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_4s_clips_per_sec"
UNIT = "clips/s"
CLIP_SAMPLES = 64000           # 4 s at 16 kHz
LONG_SAMPLES = 160000          # 10 s at 16 kHz (configs[4])
WORKLOAD = ("tiny.json training step (BASELINE.json configs[1]): %d clean/noisy 4-s 16 kHz pairs per GPU, front end + TRU-Net + "
            "mask/iSTFT + L1/MRSTFT loss, fwd+bwd, flat-bucket NCCL all-reduce (N>1), grad norm + LR schedule + AdamW")
STFT_CFG = dict(fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240], win_lengths=[240, 600, 1200],
                sc_lambda=0.5, mag_lambda=0.5)          # config/tiny.json:30-37
# block-boundary bytes of the model per frame (forward), front end / back end bytes per clip
MODEL_BYTES_PER_FRAME = 801328
STEP_BYTES_B32 = 3 * MODEL_BYTES_PER_FRAME * 32 * 501      # 38.5 GB: "3 x forward" convention for fwd+bwd


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "torch-gpu"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=150.0, help="time budget of the CPU arm (steps are cut to fit)")
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"],
                    help="flat: optim.FlatAdamW (grad norm + AdamW in one C call); torch: stock fused AdamW (A/B aid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the streaming / offline inference measurements")
    ap.add_argument("--no-stock-gpu", action="store_true", help="skip the stock-PyTorch-on-this-GPU baseline")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the global-batch-256 strong-scaling point")
    ap.add_argument("--streams", type=int, default=4096, help="concurrent streams of the streaming measurement (configs[3])")
    ap.add_argument("--offline-clips", type=int, default=1250, help="10-s clips per GPU of the offline measurement (configs[4])")
    ap.add_argument("--offline-batch", type=int, default=37,
                    help="clips per offline batch: 37 x 16 TGRU sequences = 148 CTAs of 4, one full wave of the recurrence kernel")
    return ap.parse_args()


def config_of(args, world):
    """The workload description; identical for every --impl so the driver can pair the arms."""
    B = args.batch
    return {"workload": WORKLOAD % B, "clips_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
            "clip_samples": CLIP_SAMPLES, "distinct_clips_per_gpu": B,
            "l2": "no explicit flush: one step streams >10 GB of activations through the 126 MB L2"}


# ------------------------------------------------------------------ synthetic workload
def synthetic_batch(b, n, first=0):
    """Deterministic (clean, noisy) clips of SURVEY section 8d: seeded Gaussian speech-band noise under a 3 Hz envelope plus three
    tones, and white noise at -30 dB.  Kept here so that the native arm does not touch oracle/ (the tests use the oracle's own
    copy of the same recipe)."""
    import math
    import torch
    cleans, noisies = [], []
    t = torch.arange(n, dtype=torch.float64)
    env = 0.5 * (1.0 + torch.sin(2 * math.pi * 3 * t / n))
    tones = sum(0.05 * torch.sin(2 * math.pi * f0 * t / 16000) for f0 in (220.0, 440.0, 1760.0))
    for i in range(first, first + b):
        g = torch.Generator().manual_seed(1000 + i)
        clean = (0.1 * torch.randn(n, generator=g, dtype=torch.float32).double() * env + tones).float()
        noise = 0.03 * torch.randn(n, generator=g, dtype=torch.float32).double()
        cleans.append(clean)
        noisies.append((clean.double() + noise).float())
    return torch.stack(cleans), torch.stack(noisies)


# ------------------------------------------------------------------ baseline legs (the ONLY code here that touches oracle/)
def oracle_training_step(device, batch, allow_tf32=False):
    """The reference algorithm (oracle restatement, SURVEY section 8c: the reference itself does not run) as a training step
    on `device` with stock PyTorch: train.py:118-140.  Returns (step, stage_step): callables running one step; stage_step
    returns per-stage times."""
    import torch
    from oracle import tru_oracle as O
    torch.manual_seed(0)
    net = O.randomize_bn(O.TRUNet()).to(device).train()
    opt = torch.optim.AdamW(net.parameters(), lr=4e-4)
    clean, noisy = synthetic_batch(batch, CLIP_SAMPLES)
    clean, noisy = clean.to(device), noisy.to(device)
    if device.type == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = allow_tf32
        torch.backends.cudnn.allow_tf32 = allow_tf32

    def step():
        opt.zero_grad(set_to_none=True)
        loss, _, _ = O.loss_fn(net, clean, noisy)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1e9)      # train.py:138
        opt.step()
        return loss

    def stage_step():
        """One step with CUDA events between the stages (front end / network forward / mask+iSTFT+loss / backward / optimizer)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        feats = O.frontend(noisy)
        ev[1].record()
        out = net(feats)
        ev[2].record()
        den = O.backend(out)
        l1 = torch.abs(torch.nn.functional.l1_loss(den, clean))
        sc, mg = O.mrstft_loss(den, clean)
        loss = l1 + sc + mg
        ev[3].record()
        loss.backward()
        ev[4].record()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1e9)
        opt.step()
        ev[5].record()
        torch.cuda.synchronize()
        names = ["frontend", "net_fwd", "mask_istft_loss_fwd", "backward_all", "gradnorm_adamw"]
        return {n: round(ev[i].elapsed_time(ev[i + 1]), 3) for i, n in enumerate(names)}
    return step, stage_step


def cpu_training_clips_per_sec(batch, steps, warmup, threads=None, budget_s=150.0):
    """Oracle training step on the host cores.  Returns (clips/s, threads, timed steps)."""
    import torch
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, _ = oracle_training_step(torch.device("cpu"), batch)
    t_begin = time.perf_counter()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        # stop early when the next step would not fit the budget (at least one timed step)
        if times and time.perf_counter() - t_begin + dt > budget_s:
            break
    return batch * len(times) / sum(times), torch.get_num_threads(), len(times)


def cpu_forward_latency_ms(threads):
    """configs[0]: features + network + mask + iSTFT of ONE 4-s clip on the CPU (oracle, eval mode)."""
    import torch
    from oracle import tru_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = O.randomize_bn(O.TRUNet()).eval()
    _, noisy = synthetic_batch(1, CLIP_SAMPLES)
    ts = []
    with torch.no_grad():
        for it in range(4):
            t0 = time.perf_counter()
            O.backend(net(O.frontend(noisy)))
            ts.append(1000.0 * (time.perf_counter() - t0))
    return round(statistics.median(ts[1:]), 2)


def run_reference(args):
    """--impl reference: rank 0 alone times the oracle on every host core, at the native arm's batch size and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = 1 if args.warmup >= 1 else 0
    v, cores, steps = cpu_training_clips_per_sec(args.batch, max(1, args.steps), warm, budget_s=args.cpu_seconds)
    sample = ("%d clips/step x %d timed steps (+%d warm-up) of the same tiny.json training step, oracle port of the reference on torch "
              "CPU, %d threads; steps cut from %d to fit %.0f s" % (args.batch, steps, warm, cores, args.steps, args.cpu_seconds))
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(1000.0 * args.batch / v, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, args.gpus),
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def stock_gpu_numbers(dev, batch, steps=3, warmup=2):
    """The oracle with .cuda(): stock PyTorch (cuDNN convs / GRUs, cuBLAS, cuFFT, ATen elementwise + a Python PCEN loop) on this
    GPU, same batch, fp32 with TF32 off and on.  The bar of SURVEY section 8(d) / BASELINE.md."""
    import torch
    out = {}
    for tf32 in (False, True):
        step, stage_step = oracle_training_step(dev, batch, allow_tf32=tf32)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        key = "tf32" if tf32 else "fp32"
        out[key] = {"clips_per_sec": round(batch / (ms / 1000.0), 2), "ms_per_step": round(ms, 3), "stages_ms": stage_step()}
        del step, stage_step
        torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = True
    out["what"] = ("oracle restatement of the reference with .cuda(): torch %s, cuDNN convs and GRUs, cuFFT, Python PCEN loop; "
                   "%d clips/step, %d timed steps" % (torch.__version__, batch, steps))
    return out


def run_torch_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    s = stock_gpu_numbers(dev, args.batch, steps=max(1, min(args.steps, 10)), warmup=max(1, min(args.warmup, 3)))
    v = s["fp32"]["clips_per_sec"]
    line = {"impl": "torch-gpu", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": max(1, min(args.steps, 10)),
            "warmup": max(1, min(args.warmup, 3)), "ms_per_step": s["fp32"]["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(args, 1),
            "stock_gpu": s}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ helpers
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index
        self.nvml = None

    # NVML from a thread (first sample within a millisecond; an `nvidia-smi -lms` child needs most of a 0.5-s timed region to start
    # and sometimes delivered no sample at all); the nvidia-smi loop is the fallback when the binding is missing
    def _nvml_start(self):
        import threading
        import pynvml as N
        N.nvmlInit()
        h = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            h = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = N.nvmlDeviceGetHandleByIndex(self.idx)
        self.nvml = {"N": N, "h": h, "sm": [], "reasons": 0, "stop": False}
        self.nvml["max"] = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))

        def loop():
            st = self.nvml
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            while not st["stop"]:
                try:
                    st["sm"].append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                    st["reasons"] |= int(get_reasons(h))
                except Exception:
                    pass
                time.sleep(0.02)
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def _nvml_stop(self):
        st = self.nvml
        st["stop"] = True
        self.thread.join(timeout=2)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        out = {"sm_mhz": None, "sm_max_mhz": st["max"], "reasons": sorted(k for k, b in bits.items() if st["reasons"] & b)}
        if st["sm"]:
            out.update(sm_mhz=statistics.median(st["sm"]), samples=len(st["sm"]), source="nvml, 20 ms period")
        return out

    def start(self):
        try:
            self._nvml_start()
            return
        except Exception:
            self.nvml = None
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.nvml is not None:
            return self._nvml_stop()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel family from the committed ncu pass (profiles/*_traffic.json)."""
    import glob
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            d = json.load(open(f))
            if kernel in d:
                return d[kernel]["dram_bytes_per_launch"], os.path.basename(f)
        except Exception:
            pass
    return None, None


def device_ms(fn, n, warm, sync):
    """Average milliseconds of fn() over n calls, CUDA events on the current stream, after `warm` untimed calls."""
    import torch
    for _ in range(warm):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / n


def measure_inference(args, dev, state_dict, world, rank, dist):
    """BASELINE.json metric, second half: inference real-time factor (seconds of audio per second).
    configs[3]  S concurrent streams, one frame per stream and step (front-end step -> TRU-Net step with carried TGRU state ->
                mask + iSTFT step); e2e: the step's frames come from pinned host memory and its audio goes back, every step.
    configs[4]  offline batch denoising of 10-s clips, `--offline-clips` per GPU (1,250 = 10,000 / 8), host buffers in and out
                (util.denoise_host_batches); clips are independent: replicas, no collective; aggregate = sum over ranks / max time.
    configs[0]  ONE 4-s clip, host buffer in, host buffer out, synchronous: latency."""
    import torch
    from tinyrecurrentunet_b200 import _lib as L, network, util
    peak, _ = measured_peaks()
    net = network.TRUNet().to(dev)
    net.load_state_dict(state_dict)
    net.eval()
    sync = torch.cuda.synchronize
    out = {}
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    if rank == 0:
        # ---- configs[3] ------------------------------------------------------------------------------------------
        S = args.streams
        frames_h = (0.1 * torch.randn(S, 512, generator=g)).pin_memory()
        audio_h = torch.empty(S, 128).pin_memory()
        frames_d = frames_h.to(dev)
        sd = util.StreamingDenoiser(net, S, device=dev)
        ms = device_ms(lambda: sd.step(frames_d), 30, 5, sync)

        def e2e_stream():
            audio_h.copy_(sd.step(frames_h.to(dev, non_blocking=True)), non_blocking=True)
        ms_e = device_ms(e2e_stream, 30, 3, sync)

        def pipelined_ms(sdx, n=60):
            """n hops through util.stream_host_frames (upload of hop t+1 and download of hop t-1 overlap hop t), wall clock
            from the first upload to the last block on the host, every block touched by the consumer."""
            for _ in util.stream_host_frames(sdx, (frames_h for _ in range(6)), dev):
                pass
            sync()
            t0 = time.perf_counter()
            chk = 0.0
            for blk in util.stream_host_frames(sdx, (frames_h for _ in range(n)), dev):
                chk += float(blk[0, 0])
            sync()
            return 1000.0 * (time.perf_counter() - t0) / n
        ms_p = pipelined_ms(sd)
        L.profile_enable(True)
        for _ in range(5):
            sd.step(frames_d)
        prof = L.profile_report()
        L.profile_enable(False)
        fam = {}
        for k, v in prof.items():
            a = fam.setdefault(k.split(":")[0], [0.0, 0])
            a[0] += v["ms"] / 5
            a[1] += v["launches"] // 5
        alg = MODEL_BYTES_PER_FRAME * S + 2 * (S * 16 * 128 * 4) + 2 * (S * 257 * 4) + S * 512 * 4 + S * 128 * 4 + 2 * S * 384 * 4
        out["stream"] = {"streams": S, "ms_per_step": round(ms, 4), "audio_ms_per_step": 8.0,
                         "rtf": round(S * 0.008 / (ms / 1000.0), 1),
                         "e2e": {"ms_per_step": round(ms_e, 4), "rtf": round(S * 0.008 / (ms_e / 1000.0), 1),
                                 "h2d_bytes_per_step": S * 512 * 4, "d2h_bytes_per_step": S * 128 * 4,
                                 "pipelined_ms_per_step": round(ms_p, 4), "pipelined_rtf": round(S * 0.008 / (ms_p / 1000.0), 1),
                                 "pipelined_note": "util.stream_host_frames: same copies every hop, on side streams, one hop of "
                                                   "pipeline latency; wall clock over 60 hops"},
                         "roofline": {"bound": "hbm", "algorithmic_bytes_per_step": alg, "achieved": round(alg / ms / 1e6, 1),
                                      "peak": peak, "unit": "GB/s", "frac": round(alg / ms / 1e6 / peak, 4)},
                         "kernels_ms": {k: [round(v[0], 4), v[1]] for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])},
                         "workload": "configs[3]: %d concurrent streams, 1 frame (hop 128 @16 kHz) per stream and step, "
                                     "PCEN / TGRU / overlap-add state carried" % S}
        del sd
        try:                               # the same step replayed as one CUDA graph (util.StreamingDenoiser(cuda_graph=True))
            sdg = util.StreamingDenoiser(net, S, device=dev, cuda_graph=True)
            ms_g = device_ms(lambda: sdg.step(frames_d), 30, 6, sync)

            def e2e_stream_graph():
                audio_h.copy_(sdg.step(frames_h.to(dev, non_blocking=True)), non_blocking=True)
            ms_ge = device_ms(e2e_stream_graph, 30, 3, sync)
            ms_gp = pipelined_ms(sdg)
            out["stream"]["cuda_graph"] = {"ms_per_step": round(ms_g, 4), "rtf": round(S * 0.008 / (ms_g / 1000.0), 1),
                                           "e2e_ms_per_step": round(ms_ge, 4), "e2e_rtf": round(S * 0.008 / (ms_ge / 1000.0), 1),
                                           "e2e_pipelined_ms_per_step": round(ms_gp, 4),
                                           "e2e_pipelined_rtf": round(S * 0.008 / (ms_gp / 1000.0), 1),
                                           "frac": round(alg / ms_g / 1e6 / peak, 4)}
            del sdg
        except Exception as exc:           # reported, never hidden: the launch-by-launch numbers above stand on their own
            out["stream"]["cuda_graph"] = {"error": repr(exc)[:300]}
        # ---- configs[0] ------------------------------------------------------------------------------------------
        one_h = synthetic_batch(1, CLIP_SAMPLES)[1].pin_memory()
        res_h = torch.empty(1, CLIP_SAMPLES).pin_memory()
        lat = []
        with torch.no_grad():
            for it in range(25):
                sync()
                t0 = time.perf_counter()
                res_h.copy_(util.denoise(net, one_h.to(dev, non_blocking=True))[0], non_blocking=True)
                sync()
                lat.append(1000.0 * (time.perf_counter() - t0))
        latg = []
        try:
            gd = util.GraphedDenoise(net, 1, CLIP_SAMPLES, device=dev)
            for it in range(25):
                sync()
                t0 = time.perf_counter()
                res_h.copy_(gd(one_h.to(dev, non_blocking=True))[0], non_blocking=True)
                sync()
                latg.append(1000.0 * (time.perf_counter() - t0))
            del gd
        except Exception as exc:
            latg = repr(exc)[:300]
        sd1 = util.StreamingDenoiser(net, 1, device=dev)
        fr1 = torch.zeros(1, 512, device=dev)
        lat1 = []
        for it in range(60):
            sync()
            t0 = time.perf_counter()
            sd1.step(fr1)
            sync()
            lat1.append(1000.0 * (time.perf_counter() - t0))
        lat1g = []
        try:
            sd1g = util.StreamingDenoiser(net, 1, device=dev, cuda_graph=True)
            for it in range(60):
                sync()
                t0 = time.perf_counter()
                sd1g.step(fr1)
                sync()
                lat1g.append(1000.0 * (time.perf_counter() - t0))
            del sd1g
        except Exception as exc:
            lat1g = repr(exc)[:300]
        out["latency"] = {"workload": "configs[0]: one 4-s clip, batch 1: features + network + mask + iSTFT, host buffer in and out, "
                                      "synchronous (wall clock, median of 20)",
                          "gpu_ms": round(statistics.median(lat[5:]), 3), "gpu_rtf": round(4000.0 / statistics.median(lat[5:]), 1),
                          "gpu_cuda_graph_ms": round(statistics.median(latg[5:]), 3) if isinstance(latg, list) else latg,
                          "gpu_single_frame_step_ms": round(statistics.median(lat1[10:]), 3),
                          "gpu_single_frame_step_cuda_graph_ms": (round(statistics.median(lat1g[10:]), 3)
                                                                  if isinstance(lat1g, list) else lat1g),
                          "single_frame_note": "rt.py:20-27: one stream, one frame per call, state carried; hop = 8 ms of audio; "
                                               "cuda_graph = the same step replayed as one captured graph"}
        if not args.no_cpu_baseline and world == 1:      # (at N > 1 the other ranks keep the host cores busy)
            cores = os.cpu_count() or 1
            out["latency"]["cpu_ms"] = cpu_forward_latency_ms(cores)
            out["latency"]["cpu_threads"] = cores
            out["latency"]["cpu_1thread_ms"] = cpu_forward_latency_ms(1)
    # ---- configs[4] (every rank) ---------------------------------------------------------------------------------
    Bo = max(1, args.offline_batch)
    nclips = max(Bo, args.offline_clips)
    pool = [(0.1 * torch.randn(Bo, LONG_SAMPLES, generator=g)).pin_memory() for _ in range(2)]   # 2 x Bo distinct clips, cycled

    def host_batches(n):                                                # n clips: full batches, then the ragged rest
        for i in range((n + Bo - 1) // Bo):
            yield pool[i & 1][:min(Bo, n - i * Bo)]
    for _ in util.denoise_host_batches(net, host_batches(2 * Bo + nclips % Bo), dev):      # warm-up (allocations, pinned staging of both batch sizes)
        pass
    if world > 1:
        dist.barrier()
    sync()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    checksum = 0.0
    for host_audio in util.denoise_host_batches(net, host_batches(nclips), dev):
        checksum += float(host_audio[0, 1000])                         # the consumer touches every result
    e1.record()
    sync()
    ms_off = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_off, op=dist.ReduceOp.MAX)
    ms_off = ms_off.item()
    with torch.no_grad():
        dev_ms = device_ms(lambda: util.denoise(net, pool[0].to(dev))[0], 3, 1, sync)     # device-side time of one batch (incl. its H2D)
    if rank == 0:
        frames = 1 + LONG_SAMPLES // 128
        alg = Bo * (MODEL_BYTES_PER_FRAME * frames + 4 * LONG_SAMPLES + 16 * 257 * frames + 8 * 257 * 4 * frames + 4 * LONG_SAMPLES)
        out["offline"] = {"clips_per_gpu": nclips, "gpus": world, "clip_seconds": 10.0, "batch": Bo,
                          "seconds": round(ms_off / 1000.0, 4),
                          "rtf": round(world * nclips * 10.0 / (ms_off / 1000.0), 1),
                          "rtf_per_gpu": round(nclips * 10.0 / (ms_off / 1000.0), 1),
                          "h2d_bytes_per_batch": Bo * LONG_SAMPLES * 4, "d2h_bytes_per_batch": Bo * LONG_SAMPLES * 4,
                          "ms_per_batch_device": round(dev_ms, 3),
                          "roofline": {"bound": "hbm", "algorithmic_bytes_per_batch": alg, "achieved": round(alg / dev_ms / 1e6, 1),
                                       "peak": peak, "unit": "GB/s", "frac": round(alg / dev_ms / 1e6 / peak, 4)},
                          "workload": "configs[4]: %d x 10-s clips per GPU on %d GPU(s) in batches of %d, pinned host buffers in and "
                                      "out every batch (util.denoise_host_batches), %d distinct clips cycled; no collective"
                                      % (nclips, world, Bo, 2 * Bo)}
    return out


# ------------------------------------------------------------------ native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun exactly like the driver does
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; the contract is ONE JSON line on
        # stdout, so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", init_method="env://", world_size=world, rank=rank, device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from tinyrecurrentunet_b200 import _lib as L, network, optim, stft_loss, util
    from tinyrecurrentunet_b200 import distributed as tdist

    if os.environ.get("TRU_DBG_FLAGS"):                    # bottleneck hunting: disable parts of the GEMM kernel (results are garbage)
        L.lib.tru_debug_set_flags(int(os.environ["TRU_DBG_FLAGS"]))
    B = args.batch
    torch.manual_seed(0)
    net = network.TRUNet(3, 64, 3, 128, [5, 3], [2, 1], 192).to(dev).train()
    mr = stft_loss.MultiResolutionSTFTLoss(**STFT_CFG).to(dev)
    if world > 1:
        tdist.apply_gradient_allreduce(net)
    if args.optimizer == "flat":                          # train.py:68 + :138-140 as one C call (two launches)
        opt = optim.FlatAdamW(net.parameters(), lr=4e-4, max_grad_norm=1e9)
    else:                                                 # A/B aid: stock torch (fused multi-tensor AdamW, no norm)
        opt = torch.optim.AdamW(net.parameters(), lr=4e-4, fused=True)
    sched = util.LinearWarmupCosineDecay(opt, lr_max=4e-4, n_iter=25_000_000, iteration=0, divider=25,
                                         warmup_proportion=0.05)          # train.py:98-104 (tiny.json: 25M iterations)

    # synthetic clips: B distinct clips per rank
    clean_h, noisy_h = synthetic_batch(B, CLIP_SAMPLES, first=rank * B)
    clean_h, noisy_h = clean_h.pin_memory(), noisy_h.pin_memory()
    clean_d, noisy_d = clean_h.to(dev), noisy_h.to(dev)

    def step(clean, noisy):
        opt.zero_grad(set_to_none=True)
        loss, _ = util.loss_fn(net, (clean, noisy), ell_p=1, ell_p_lambda=1, stft_lambda=1, mrstftloss=mr)
        if world > 1:
            tdist.attach_loss(net, loss)     # train.py:133's logging mean rides in the gradient bucket (net.reduced_loss)
        loss.backward()                      # N>1: the all-reduce fires from the engine callback
        sched.step()                         # train.py:139 (host float, no sync)
        opt.step()                           # train.py:138 + :140
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(nsteps, fn):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nsteps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    dp_check = None
    for it in range(max(args.warmup, 3)):
        step(clean_d, noisy_d)
        if it == 0 and world > 1:
            # after the all-reduce every rank must hold the same (mean) gradients, bit for bit, and the same mean loss
            flat = net._tru_flat_grad
            hi, lo = flat.clone(), flat.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            same = bool(torch.equal(hi, lo)) and bool(torch.isfinite(flat).all())
            if not same:
                raise RuntimeError("data-parallel check failed: gradients differ across ranks after the all-reduce")
            dp_check = {"gradients_identical_across_ranks": True, "mean_loss_in_bucket_tail": float(net.reduced_loss.item())}
    sync_all()

    # ---- (1) device-resident throughput -------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = L.lib.tru_launch_count()
    ms = timed(args.steps, lambda: step(clean_d, noisy_d))
    launches = L.lib.tru_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else {}
    value = world * B * args.steps / (ms / 1000.0)

    # ---- (2) end to end: pinned host buffers in, loss read back, every step ----------
    e2e = None
    if not args.no_e2e:
        def host_batches(n):
            for _ in range(n):
                yield clean_h, noisy_h

        def e2e_epoch():
            # every step: this step's inputs travel pinned host -> device (one batch ahead, on a side stream) and the loss
            # comes back to the host
            for clean, noisy in util.CudaPrefetcher(host_batches(args.steps), dev):
                step(clean, noisy).item()
        for clean, noisy in util.CudaPrefetcher(host_batches(1), dev):
            step(clean, noisy).item()
        ms_e = timed(1, e2e_epoch)
        e2e = {"value": round(world * B * args.steps / (ms_e / 1000.0), 2), "unit": UNIT,
               "h2d_bytes_per_step": 2 * B * CLIP_SAMPLES * 4 * world, "d2h_bytes_per_step": 4 * world}

    # ---- (3) per-kernel CUDA-event profile over a second timed region ---------------
    L.profile_enable(True)
    ms_p = timed(args.steps, lambda: step(clean_d, noisy_d))
    prof_detail = L.profile_report()
    L.profile_enable(False)
    prof = {}
    for k, v in prof_detail.items():               # "igemm_tc:M=..,K=.." -> "igemm_tc"
        a = prof.setdefault(k.split(":")[0], dict(launches=0, ms=0.0, bytes=0.0, flops=0.0))
        for f in a:
            a[f] += v[f]
    if os.environ.get("TRU_BENCH_DETAIL") and rank == 0:
        for k, v in sorted(prof_detail.items(), key=lambda kv: -kv[1]["ms"])[:80]:
            print("# %-60s n=%5.1f %8.3f ms/step %7.1f GB/s %7.1f TF" % (
                k, v["launches"] / args.steps, v["ms"] / args.steps, v["bytes"] / max(v["ms"], 1e-9) / 1e6,
                v["flops"] / max(v["ms"], 1e-9) / 1e9), file=sys.stderr)
    peak, peak_src = measured_peaks()
    total_kernel_ms = sum(v["ms"] for v in prof.values()) or 1.0
    top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    tname, t = top
    ach = t["bytes"] / (t["ms"] / 1000.0) / 1e9
    roofline = {"kernel": tname, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": round(t["bytes"] / t["launches"], 0),
                "launches_per_step": t["launches"] / args.steps,
                "avg_launch_ms": round(t["ms"] / t["launches"], 4),
                "share_of_kernel_time": round(t["ms"] / total_kernel_ms, 4),
                "tflops_fp32": round(t["flops"] / (t["ms"] / 1000.0) / 1e12, 2),
                "ms_per_step_profiled": round(ms_p / args.steps, 3),
                # the whole step against SURVEY section 8(d)'s algorithmic bytes (3 x forward block-boundary traffic)
                "whole_step": {"algorithmic_bytes": STEP_BYTES_B32 * B // 32, "achieved": round(STEP_BYTES_B32 * B / 32 / (ms / args.steps) / 1e6, 1),
                               "frac": round(STEP_BYTES_B32 * B / 32 / (ms / args.steps) / 1e6 / peak, 4),
                               "bytes_moved_by_kernels": round(sum(v["bytes"] for v in prof.values()) / args.steps, 0)}}
    traffic, tsrc = measured_traffic(tname)
    if traffic is not None:
        roofline["traffic"] = traffic
        roofline["traffic_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per launch (profiles/%s)" % tsrc
    kernels = {k: {"ms_per_step": round(v["ms"] / args.steps, 4), "launches_per_step": v["launches"] / args.steps,
                   "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                   "TFLOPs": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2)}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- (4) strong-scaling point of configs[2]: global batch 256 over the N ranks ---------
    strong = None
    if world > 1 and not args.no_strong and 256 % world == 0:
        per = 256 // world
        try:
            torch.cuda.empty_cache()
            reps = (per + B - 1) // B
            cs, ns = clean_d.repeat(reps, 1)[:per].contiguous(), noisy_d.repeat(reps, 1)[:per].contiguous()
            step(cs, ns)
            ms_s = timed(3, lambda: step(cs, ns))
            strong = {"global_batch": 256, "clips_per_gpu": per, "steps": 3, "ms_per_step": round(ms_s / 3, 3),
                      "value": round(256 * 3 / (ms_s / 1000.0), 2), "unit": UNIT,
                      "note": "configs[2] as BASELINE.json states it (config/tiny.json:24 x 256): the %d distinct clips of this rank "
                              "tiled to %d" % (B, per)}
            del cs, ns
        except Exception as e:          # e.g. out of memory at N = 2 on a smaller part: report, do not lose the main line
            strong = {"global_batch": 256, "clips_per_gpu": per, "error": str(e)[:200]}
        torch.cuda.empty_cache()

    # ---- (5) baselines on this box (rank 0, N = 1 only): host CPU, stock PyTorch on this GPU ---------------
    cpu = None
    stock = None
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    opt = None
    net = None
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_stock_gpu:
        stock = stock_gpu_numbers(dev, B)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, n = cpu_training_clips_per_sec(B, 1, 1, budget_s=60.0)
        v1, _, _ = cpu_training_clips_per_sec(2, 1, 0, threads=1, budget_s=30.0)
        torch.set_num_threads(os.cpu_count() or 1)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d clips/step x %d timed step (+1 warm-up) of the same training step (oracle port, torch CPU, %d threads)" % (B, n, cores),
               "value_1thread": round(v1, 4), "sample_1thread": "2 clips x 1 step, 1 thread"}

    # ---- (6) inference real-time factors ----------------------------------------------------------------------
    inference = None
    if not args.no_inference:
        inference = measure_inference(args, dev, state, world, rank, dist)

    if rank == 0:
        cfg = config_of(args, world)
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "optimizer": "FlatAdamW (tru_flat_adamw_step)" if args.optimizer == "flat" else "torch.optim.AdamW(fused)",
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu, "stock_gpu": stock, "strong_scaling": strong, "dp_check": dp_check,
                "inference": inference, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-gpu":
        run_torch_gpu(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
