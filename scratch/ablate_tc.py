import ctypes as C, torch, sys, os
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_pw
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*7 + [C.c_int]*4 + [C.c_void_p]
torch.manual_seed(0)
def t(M,K,N,affine,stats,flags,lw=0):
    L.lib.tru_debug_set_flags(flags); L.lib.tru_debug_set_loader_warps(lw)
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N,K,device='cuda'); b = torch.randn(N, device='cuda'); out = torch.empty(M,N,device='cuda')
    p0 = (torch.rand(K, device='cuda') + 0.5) if affine else None
    p2 = torch.randn(K, device='cuda') if affine else None
    st = torch.zeros(2*N, device='cuda', dtype=torch.float64) if stats else None
    a = (x.data_ptr(), p0.data_ptr() if affine else None, p2.data_ptr() if affine else None, w.data_ptr(), b.data_ptr(), out.data_ptr(), st.data_ptr() if stats else None, M,K,N,1,None)
    for _ in range(2): L.check(fn(*a))
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn(*a)
    e1.record(); torch.cuda.synchronize()
    L.lib.tru_debug_set_flags(0)
    return e0.elapsed_time(e1)/5
names = {0:"full",1:"noMMA",2:"noLDG",4:"noSTS",8:"noEPI",3:"noMMA+noLDG",9:"noMMA+noEPI",10:"noLDG+noEPI",6:"noLDG+noSTS", 14:"onlyMMA(noLDG,STS,EPI)",15:"nothing", 7:"onlyEPI", 13:"onlyLDG", 11:"onlySTS"}
for (M,K,N) in [(2052096,128,128),(2052096,192,64),(2052096,64,128),(2052096,64,64)]:
    for affine in (True,):
        print("shape", (M,K,N), "affine+stats" if affine else "plain", "ideal HBM ms %.3f" % (4*(M*K+M*N)/6.5237e9))
        for f in (0,1,2,4,8,3,9,10,6,14,7,13,11,15):
            print("   %-26s %.3f ms" % (names[f], t(M,K,N,affine,affine,f)), flush=True)
