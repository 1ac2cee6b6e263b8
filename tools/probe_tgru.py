"""TGRU recurrence kernels: ms per launch and us per time step for 1 / 2 / 4 sequences per CTA (gru.cu tgru_seqs_per_cta),
each at 144 CTAs: B = 9, 18, 36 clips of 4 s (T' = 501), forward (eval) and forward + BPTT (train)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tinyrecurrentunet_b200 import _lib as L, network  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = network.TRUNet().to(dev)
T = 501
BS = [int(v) for v in os.environ.get("PROBE_B", "1,9,18,36").split(",")]
MODES = os.environ.get("PROBE_MODE", "eval,train").split(",")
for B in BS:
    x = torch.randn(B, T, 4, 257, device=dev)
    for mode in MODES:
        net.train(mode == "train")

        def run():
            if mode == "eval":
                with torch.no_grad():
                    net(x)
            else:
                net.zero_grad(set_to_none=True)
                net(x).square().mean().backward()
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        L.profile_enable(True)
        for _ in range(3):
            run()
        prof = L.profile_report()
        L.profile_enable(False)
        for k in ("tgru_fwd", "tgru_bwd"):
            v = [p for n, p in prof.items() if n.split(":")[0] == k]
            if v:
                ms = sum(p["ms"] for p in v) / sum(p["launches"] for p in v)
                print("B=%2d (%3d sequences) %-5s %-8s %.4f ms  %.3f us per step" % (B, 16 * B, mode, k, ms, 1000.0 * ms / T))
