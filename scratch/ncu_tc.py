import ctypes as C, torch, sys
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_pw
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*7 + [C.c_int]*4 + [C.c_void_p]
M,K,N = 2052096,128,128
x = torch.randn(M, K, device='cuda'); w = torch.randn(N,K,device='cuda'); b = torch.randn(N, device='cuda'); out = torch.empty(M,N,device='cuda')
for _ in range(3): fn(x.data_ptr(), None, None, w.data_ptr(), b.data_ptr(), out.data_ptr(), None, M,K,N,1,None)
torch.cuda.synchronize()
print("done")
