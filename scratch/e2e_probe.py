"""Where does the gap between device-resident and end-to-end throughput come from?  Variants of bench.py's e2e loop."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyrecurrentunet_b200 import network, optim, stft_loss, util

dev = torch.device("cuda")
B, N = 32, 64000
torch.manual_seed(0)
net = network.TRUNet().to(dev).train()
mr = stft_loss.MultiResolutionSTFTLoss(fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240], win_lengths=[240, 600, 1200]).to(dev)
opt = optim.FlatAdamW(net.parameters(), lr=4e-4, max_grad_norm=1e9)
sched = util.LinearWarmupCosineDecay(opt, lr_max=4e-4, n_iter=25_000_000, iteration=0, divider=25, warmup_proportion=0.05)
clean_h = torch.randn(8, N) * 0.1
noisy_h = clean_h + 0.03 * torch.randn(8, N)
clean_h = clean_h.repeat(4, 1).contiguous().pin_memory()
noisy_h = noisy_h.repeat(4, 1).contiguous().pin_memory()
clean_d, noisy_d = clean_h.to(dev), noisy_h.to(dev)


def step(c, n, zero_first=True):
    if zero_first:
        opt.zero_grad(set_to_none=True)
    loss, _ = util.loss_fn(net, (c, n), mrstftloss=mr)
    loss.backward()
    sched.step()
    opt.step()
    if not zero_first:
        opt.zero_grad(set_to_none=True)
    return loss


def timed(fn, n=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


pre = util.CudaPrefetcher("cuda")


def v_resident():
    step(clean_d, noisy_d)


def v_item_only():
    return step(clean_d, noisy_d).item()


def v_copy_only():
    bufs = pre.next(clean_h, noisy_h)
    step(bufs[0], bufs[1])
    pre.release(bufs)


def v_e2e():
    bufs = pre.next(clean_h, noisy_h)
    loss = step(bufs[0], bufs[1])
    pre.release(bufs)
    return loss.item()


state = {}


def v_e2e_reordered():
    bufs = state.get("bufs") or pre.next(clean_h, noisy_h)
    loss = step(bufs[0], bufs[1], zero_first=False)
    pre.release(bufs)
    state["bufs"] = pre.next(clean_h, noisy_h)
    return loss.item()


pinned_loss = torch.zeros(2, pin_memory=True)
ev = [torch.cuda.Event(), torch.cuda.Event()]
cnt = {"i": 0}


def v_e2e_lagged_read():
    """loss copied to pinned memory asynchronously, read one step later (still one D2H per step)."""
    i = cnt["i"]
    bufs = pre.next(clean_h, noisy_h)
    loss = step(bufs[0], bufs[1])
    pre.release(bufs)
    pinned_loss[i & 1].copy_(loss.detach(), non_blocking=True)
    ev[i & 1].record()
    if i > 0:
        ev[(i - 1) & 1].synchronize()
        _ = float(pinned_loss[(i - 1) & 1])
    cnt["i"] = i + 1


for name, fn in (("resident", v_resident), ("item only", v_item_only), ("copy only", v_copy_only), ("e2e (bench)", v_e2e),
                 ("e2e reordered", v_e2e_reordered), ("e2e lagged read", v_e2e_lagged_read), ("resident", v_resident)):
    g, w = timed(fn)
    print("%-18s gpu %.3f ms/step  wall %.3f ms/step  -> %.1f clips/s" % (name, g, w, B / g * 1e3))
