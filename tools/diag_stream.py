#!/usr/bin/env python
"""Diagnostic: where do frontend_step and frontend differ, and which one is closer to the CPU oracle?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyrecurrentunet_b200 import ops
from oracle import tru_oracle as O
S, T = 56, 10
g = torch.Generator().manual_seed(99)
audio = 0.1 * torch.randn(4096, 128 * (T - 1), generator=g)[:S].contiguous()
audio[::7] *= 0.01
x = audio.cuda()
off = ops.frontend(x).cpu()
ref = O.frontend(audio)
ref64 = O.frontend(audio.double()).float() if False else None
xp = torch.nn.functional.pad(x.unsqueeze(1), (256, 256), mode="reflect").squeeze(1)
pcen = torch.zeros(S, 257, device="cuda")
mag = O.stft_rect(audio).abs().transpose(1, 2)      # (S, T, F)
for t in range(T):
    f = ops.frontend_step(xp[:, 128 * t:128 * t + 512].contiguous(), pcen).cpu()
    for name, a in (("step", f), ("offline", off[:, t])):
        d = (a - ref[:, t]).abs()
        q = d[::7]; l = torch.cat([d[i::7] for i in range(1, 7)])
        print("t=%d %-8s vs oracle: quiet streams %s   loud streams %s" % (
            t, name, ["%.1e" % q[:, c].max().item() for c in range(4)], ["%.1e" % l[:, c].max().item() for c in range(4)]))
    d = (f - off[:, t]).abs()[:, 1]
    i = d.argmax().item(); s_, k = divmod(i, 257)
    print("   worst PCEN step-vs-offline: stream %d bin %d: step %.5f offline %.5f oracle %.5f |X| %.3e (max |X| of frame %.3e)" % (
        s_, k, f[s_, 1, k], off[s_, t, 1, k], ref[s_, t, 1, k], mag[s_, t, k], mag[s_, t].max()))
