"""Per-launch profile of the inference paths (CUDA events around every launch, library profiler):
    python tools/prof_infer.py [--streams 4096] [--offline-batch 25]
prints one line per kernel shape: launches per step, ms per launch, GB/s (algorithmic bytes of the launch), TFLOP/s."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tinyrecurrentunet_b200 import _lib as L, network, util  # noqa: E402


WARM = 3


def report(title, fn, reps):
    for _ in range(WARM):
        fn()
    torch.cuda.synchronize()
    L.profile_enable(True)
    for _ in range(reps):
        fn()
    prof = L.profile_report()
    L.profile_enable(False)
    tot = sum(v["ms"] for v in prof.values()) / reps
    print("== %s: %.4f ms of kernels per step" % (title, tot))
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        n = max(1, v["launches"])
        ms = v["ms"] / n
        print("  %-58s x%-3d %8.4f ms  %7.0f GB/s %6.1f TF/s" % (k, v["launches"] // reps, ms, v["bytes"] / n / ms / 1e6 if ms else 0,
                                                                   v["flops"] / n / ms / 1e9 if ms else 0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--offline-batch", type=int, default=37)
    ap.add_argument("--fusion", type=int, default=1)
    ap.add_argument("--ncu", action="store_true", help="one warm-up and one measured pass per path (for an ncu launch list)")
    a = ap.parse_args()
    global WARM
    if a.ncu:
        WARM = 1
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = network.TRUNet().to(dev).eval()
    L.lib.tru_debug_set_eval_fusion(a.fusion)
    frames = 0.1 * torch.randn(a.streams, 512, device=dev)
    sd = util.StreamingDenoiser(net, a.streams, device=dev)
    report("streaming step, %d streams" % a.streams, lambda: sd.step(frames), 1 if a.ncu else 5)
    del sd
    if a.offline_batch:
        clips = 0.1 * torch.randn(a.offline_batch, 160000, device=dev)
        with torch.no_grad():
            report("offline batch, %d x 10-s clips" % a.offline_batch, lambda: util.denoise(net, clips), 1 if a.ncu else 3)


if __name__ == "__main__":
    main()
