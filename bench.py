#!/usr/bin/env python
"""bench.py - TRU-Net training-step throughput (BASELINE.json: "train 4-s clips/sec").

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host CPU

A "step" = one pass of the hot path over one batch of synthetic 16 kHz clips:
front end -> TRU-Net -> mask + iSTFT -> L1 + multi-resolution STFT loss -> backward ->
(N > 1: one NCCL all-reduce of the flat gradient bucket) -> AdamW step.
Workload = BASELINE.json configs[1]: tiny.json, 32 clean/noisy 4-s pairs per GPU
(weak scaling: global batch 32*N).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_4s_clips_per_sec"
UNIT = "clips/s"
CLIP_SAMPLES = 64000           # 4 s at 16 kHz
WORKLOAD = ("tiny.json training step (BASELINE.json configs[1]): %d clean/noisy 4-s 16 kHz pairs per GPU, front end + TRU-Net + "
            "mask/iSTFT + L1/MRSTFT loss, fwd+bwd, flat-bucket NCCL all-reduce (N>1), grad norm + LR schedule + AdamW")
STFT_CFG = dict(fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240], win_lengths=[240, 600, 1200],
                sc_lambda=0.5, mag_lambda=0.5)          # config/tiny.json:30-37


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--cpu-batch", type=int, default=2, help="clips per CPU step (bounded sample)")
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"],
                    help="flat: optim.FlatAdamW (grad norm + AdamW in one C call); torch: stock fused AdamW (A/B aid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the streaming / offline inference RTF measurements")
    ap.add_argument("--streams", type=int, default=4096, help="concurrent streams of the streaming measurement (configs[3])")
    return ap.parse_args()


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


# ------------------------------------------------------------------ CPU arm (oracle)
def cpu_training_clips_per_sec(batch, steps, warmup):
    """The reference algorithm (oracle restatement, SURVEY section 8c: the reference itself does
    not run) on the host cores with every thread torch can use.  Returns (clips/s, cores)."""
    import torch
    from oracle import tru_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = O.randomize_bn(O.TRUNet()).train()
    opt = torch.optim.AdamW(net.parameters(), lr=4e-4)
    clean, noisy = O.synthetic_batch(batch, n=CLIP_SAMPLES)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _, _ = O.loss_fn(net, clean, noisy)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1e9)      # train.py:138
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    v, cores = cpu_training_clips_per_sec(args.cpu_batch, steps, warm)
    sample = "%d clips/step x %d steps of the tiny.json training step (fwd+bwd+AdamW), torch CPU, %d threads" % (
        args.cpu_batch, steps, cores)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(1000.0 * args.cpu_batch / v, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % args.batch, "clips_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
                       "parallelism": "dp%d" % args.gpus, "device": "host CPU (rank 0 only)",
                       "sample_clips_per_step": args.cpu_batch},
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
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


def measure_inference(args, dev, state_dict):
    """BASELINE.json metric, second half: inference real-time factor.  configs[3]: S concurrent streams, one frame per
    stream and step (front end step -> TRU-Net step with carried TGRU state -> mask + iSTFT step);  configs[4]: offline
    batch denoising of 10-s clips (front end -> TRU-Net -> mask + iSTFT).  RTF = seconds of audio per second."""
    import torch
    from tinyrecurrentunet_b200 import network, util
    net = network.TRUNet().to(dev)
    net.load_state_dict(state_dict)
    net.eval()

    def timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {}
    S = args.streams
    g = torch.Generator(device="cpu").manual_seed(7)
    frames = (0.1 * torch.randn(S, 512, generator=g)).to(dev)
    sd = util.StreamingDenoiser(net, S, device=dev)
    ms = timed(lambda: sd.step(frames), 30)
    out["stream"] = {"streams": S, "ms_per_step": round(ms, 4), "audio_ms_per_step": 8.0,
                     "rtf": round(S * 0.008 / (ms / 1000.0), 1),
                     "workload": "configs[3]: %d concurrent streams, 1 frame (hop 128 @16 kHz) per stream and step, "
                                 "PCEN / TGRU / overlap-add state carried" % S}
    del sd
    Bo, No = 32, 160000
    audio = (0.1 * torch.randn(Bo, No, generator=g)).to(dev)
    with torch.no_grad():
        ms = timed(lambda: util.denoise(net, audio)[0], 3, warm=1)
    out["offline"] = {"clips": Bo, "clip_seconds": 10.0, "ms_per_batch": round(ms, 3),
                      "rtf": round(Bo * 10.0 / (ms / 1000.0), 1),
                      "workload": "configs[4] shard: %d x 10-s clips per call on one GPU (clips are independent: replicas, "
                                  "no collective)" % Bo}
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

    if os.environ.get("TRU_LOADER_WARPS"):                 # tuning aid (8 or 16 loader warps in the GEMM kernels)
        L.lib.tru_debug_set_loader_warps(int(os.environ["TRU_LOADER_WARPS"]))
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

    # synthetic clips: a pool of distinct clips, rank-dependent, tiled to the batch
    pool = min(B, 8)
    clean_h, noisy_h = synthetic_batch(pool, CLIP_SAMPLES, first=rank * pool)
    reps = (B + pool - 1) // pool
    clean_h = clean_h.repeat(reps, 1)[:B].contiguous().pin_memory()
    noisy_h = noisy_h.repeat(reps, 1)[:B].contiguous().pin_memory()
    clean_d = clean_h.to(dev)
    noisy_d = noisy_h.to(dev)

    def step(clean, noisy):
        opt.zero_grad(set_to_none=True)
        loss, _ = util.loss_fn(net, (clean, noisy), ell_p=1, ell_p_lambda=1, stft_lambda=1, mrstftloss=mr)
        loss.backward()                  # N>1: the all-reduce fires from the engine callback
        sched.step()                     # train.py:139 (host float, no sync)
        opt.step()                       # train.py:138 + :140
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

    for _ in range(max(args.warmup, 3)):
        step(clean_d, noisy_d)
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
        pre = util.CudaPrefetcher(dev)      # H2D of the next batch on a side stream while the current step runs

        def e2e_step():
            bufs = pre.next(clean_h, noisy_h)          # every step: pinned host buffers -> device (this step's inputs)
            loss = step(bufs[0], bufs[1])
            pre.release(bufs)
            return loss.item()                         # every step: loss read back
        e2e_step()
        ms_e = timed(args.steps, e2e_step)
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
        for k, v in sorted(prof_detail.items(), key=lambda kv: -kv[1]["ms"])[:60]:
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
                "ms_per_step_profiled": round(ms_p / args.steps, 3)}
    traffic, tsrc = measured_traffic(tname)
    if traffic is not None:
        roofline["traffic"] = traffic
        roofline["traffic_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per launch (profiles/%s)" % tsrc
    kernels = {k: {"ms_per_step": round(v["ms"] / args.steps, 4), "launches_per_step": v["launches"] / args.steps,
                   "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                   "TFLOPs": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2)}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- (4) CPU baseline on this box's host cores (rank 0, N = 1 only) ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores = cpu_training_clips_per_sec(args.cpu_batch, 3, 1)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d clips/step x 3 steps of the same training step (oracle, torch CPU)" % args.cpu_batch}

    # ---- (5) inference real-time factors (rank 0, N = 1 only) ----------------------------
    inference = None
    if rank == 0 and world == 1 and not args.no_inference:
        state = {k: v.detach().clone() for k, v in net.state_dict().items()}
        opt = None
        net = None
        torch.cuda.empty_cache()
        inference = measure_inference(args, dev, state)

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD % B,
                           "clips_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                           "optimizer": "FlatAdamW (tru_flat_adamw_step)" if args.optimizer == "flat" else "torch.optim.AdamW(fused)",
                           "l2": "no explicit flush: one step streams >10 GB of activations through the 126 MB L2"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu, "inference": inference, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
