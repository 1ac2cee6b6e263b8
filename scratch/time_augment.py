"""Time the batch-assembly kernels at the bench batch (32 x 4 s)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyrecurrentunet_b200 import _lib as L, dataset
aug = dataset.DataAugment()
B, N = 32, 64000
clean = torch.randn(B, N + 6000, device="cuda") * 0.1
noise = torch.randn(B, N, device="cuda") * 0.5
params = [aug.sample_params() for _ in range(B)]
for _ in range(3):
    dataset.assemble_batch(clean, noise, N, aug, params)
torch.cuda.synchronize()
L.profile_enable(True)
for _ in range(20):
    dataset.assemble_batch(clean, noise, N, aug, params)
rep = L.profile_report()
for k, v in rep.items():
    print(k, v["launches"], "launches", round(1000 * v["ms"] / v["launches"], 1), "us each", round(v["bytes"] / v["launches"] / (v["ms"] / v["launches"]) / 1e6, 1), "GB/s")
