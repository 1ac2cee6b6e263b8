import ctypes as C, torch, sys
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_wgrad
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*4 + [C.c_int]*4 + [C.c_void_p]
M,Cc,N = 2052096,128,128
a = torch.randn(M, Cc, device='cuda'); z = torch.randn(M, N, device='cuda'); dw = torch.zeros(N,Cc,device='cuda'); db=torch.zeros(N,device='cuda')
for _ in range(3): fn(a.data_ptr(), z.data_ptr(), dw.data_ptr(), db.data_ptr(), M, Cc, N, 1, None)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn(a.data_ptr(), z.data_ptr(), dw.data_ptr(), db.data_ptr(), M, Cc, N, 1, None)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/5
print("wgrad tc %.3f ms  %.1f TF %.0f GB/s" % (ms, 2*M*Cc*N/ms/1e9, 4*M*(Cc+N)/ms/1e6))
