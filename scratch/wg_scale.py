import ctypes as C, torch, sys
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_wgrad_stream
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*8 + [C.c_int]*4 + [C.c_void_p]
torch.manual_seed(0)
def run(M, Lq, Cc, N, bn):
    a = torch.randn(M, Cc, device='cuda'); dy = torch.randn(M, N, device='cuda'); z = torch.randn(M, N, device='cuda') if bn else None
    q = [torch.rand(N, device='cuda') + 0.5 for _ in range(3)] if bn else [None]*3
    dw = torch.zeros(N, Cc, device='cuda'); db = torch.zeros(N, device='cuda')
    args = (a.data_ptr(), dy.data_ptr(), z.data_ptr() if bn else None, *[t.data_ptr() if bn else None for t in q], dw.data_ptr(), db.data_ptr(), M, Lq, Cc, N, None)
    L.check(fn(*args), "wgs"); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn(*args)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byt = 4.0 * M * (Cc + N * (2 if bn else 1))
    print("M=%8d Lq=%5d C=%3d N=%3d bn=%d: %.3f ms  %5.0f GB/s  (%.2f us per unit per CTA)" % (M, Lq, Cc, N, bn, ms, byt / ms / 1e6, ms * 1e3 / (M / 16 / 148)), flush=True)
for M in (256512, 2052096):
    for (Cc, N) in ((128, 128), (128, 192), (128, 256), (128, 384), (64, 384), (64, 192)):
        run(M, 16, Cc, N, False)
