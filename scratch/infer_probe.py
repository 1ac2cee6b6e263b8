import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyrecurrentunet_b200 import _lib as L, network, ops, util
dev = torch.device("cuda")
torch.manual_seed(0)
net = network.TRUNet().to(dev).eval()
def timeit(fn, n, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def prof(fn, n, tag):
    L.profile_enable(True); 
    for _ in range(n): fn()
    rep = L.profile_report(); L.profile_enable(False)
    print("## profile", tag)
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:25]:
        print("#  %-62s n=%5.1f %8.3f ms %7.1f GB/s %6.1f TF" % (k, v["launches"]/n, v["ms"]/n, v["bytes"]/max(v["ms"],1e-9)/1e6, v["flops"]/max(v["ms"],1e-9)/1e9))
# --- streaming
S = int(os.environ.get("S", "4096"))
frames = torch.randn(S, 512, device=dev) * 0.1
pst = torch.zeros(S, 257, device=dev); h = torch.zeros(S * 16, 128, device=dev)
def sstep():
    global h
    f = ops.frontend_step(frames, pst)
    out, h = net.step(f, h)
    return out
ms = timeit(sstep, 20)
print("stream S=%d: %.3f ms/step -> RTF %.1f (x realtime, all streams)" % (S, ms, S * 0.008 / (ms / 1e3)))
prof(sstep, 5, "stream")
# --- offline 10 s clips
for B in (16, 64):
    audio = torch.randn(B, 160000, device=dev) * 0.1
    with torch.no_grad():
        fn = lambda: util.denoise(net, audio)[0]
        ms = timeit(fn, 5)
        print("offline B=%d x 10 s: %.3f ms -> RTF %.1f" % (B, ms, B * 10.0 / (ms / 1e3)))
with torch.no_grad():
    prof(fn, 3, "offline B=64")
print("mem GB", torch.cuda.max_memory_allocated() / 1e9)
