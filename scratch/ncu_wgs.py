import ctypes as C, torch, sys
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_wgrad_stream
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*8 + [C.c_int]*4 + [C.c_void_p]
M, Lq, Cc, N = int(sys.argv[1]) if len(sys.argv) > 1 else 513024, 64, 128, 128
torch.manual_seed(0)
a = torch.randn(M, Cc, device='cuda'); dy = torch.randn(M, N, device='cuda'); z = torch.randn(M, N, device='cuda')
q0 = torch.rand(N, device='cuda') + 0.5; q1 = torch.randn(N, device='cuda') * 0.1; q2 = torch.randn(N, device='cuda') * 0.1
dw = torch.zeros(N, Cc, device='cuda'); db = torch.zeros(N, device='cuda')
def run():
    L.check(fn(a.data_ptr(), dy.data_ptr(), z.data_ptr(), q0.data_ptr(), q1.data_ptr(), q2.data_ptr(), dw.data_ptr(), db.data_ptr(), M, Lq, Cc, N, None), "wgs")
run(); torch.cuda.synchronize()
dz = (q0 * dy + q1 * z + q2).double()
ref = dz.t() @ a.double()
print("dW rel err", ((dw.double() - ref).abs().max() / ref.abs().max()).item(), "db rel err", ((db.double() - dz.sum(0)).abs().max() / dz.sum(0).abs().max()).item())
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("M=%d C=%d N=%d: %.3f ms  %.0f GB/s" % (M, Cc, N, ms, 4.0 * M * (Cc + 2 * N) / ms / 1e6))
