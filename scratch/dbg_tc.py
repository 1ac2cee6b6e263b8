import ctypes as C, torch, sys, time
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_pw
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*7 + [C.c_int]*4 + [C.c_void_p]
torch.manual_seed(0)
def run(M, K, N, affine, stats, use_tc):
    x = torch.randn(M, K, device='cuda')
    w = torch.randn(N, K, device='cuda') / K**0.5
    b = torch.randn(N, device='cuda')
    p0 = (torch.rand(K, device='cuda') + 0.5) if affine else None
    p2 = torch.randn(K, device='cuda') if affine else None
    out = torch.full((M, N), float('nan'), device='cuda')
    st = torch.zeros(2*N, device='cuda', dtype=torch.float64) if stats else None
    rc = fn(x.data_ptr(), p0.data_ptr() if affine else None, p2.data_ptr() if affine else None, w.data_ptr(), b.data_ptr(),
            out.data_ptr(), st.data_ptr() if stats else None, M, K, N, use_tc, None)
    L.check(rc, "debug_pw")
    torch.cuda.synchronize()
    a = x.double()
    if affine: a = torch.relu(a * p0.double() + p2.double())
    ref = a @ w.double().t() + b.double()
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    serr = 0.0
    if stats:
        s1 = ref.sum(0); s2 = (ref*ref).sum(0)
        serr = max(((st[:N]-s1).abs().max()/s1.abs().max()).item(), ((st[N:]-s2).abs().max()/s2.abs().max()).item())
    return err, serr
for shape in [(1000,128,8),(3000,8,64),(5000,40,8),(128,32,32),(128,64,64),(256,64,128),(1000,128,128),(128*300+17,128,128),(5000,192,64),(4097,64,384),(3000,128,192),(777,384,128),(2048,320,64)]:
    for affine, stats in ((False, False), (True, True)):
        try:
            e_tc = run(*shape, affine, stats, 1)
        except Exception as ex:
            e_tc = str(ex)[:80]
        e_si = run(*shape, affine, stats, 0)
        print(shape, "affine/stats" if affine else "plain", " TC:", e_tc, " SIMT:", e_si, flush=True)
# timing
M,K,N = 2052096,128,128
x = torch.randn(M, K, device='cuda'); w = torch.randn(N,K,device='cuda'); b = torch.randn(N, device='cuda'); out = torch.empty(M,N,device='cuda')
for use_tc in (1,0):
    for _ in range(2): fn(x.data_ptr(), None, None, w.data_ptr(), b.data_ptr(), out.data_ptr(), None, M,K,N,use_tc,None)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn(x.data_ptr(), None, None, w.data_ptr(), b.data_ptr(), out.data_ptr(), None, M,K,N,use_tc,None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print("use_tc", use_tc, "M,K,N", (M,K,N), "%.3f ms  %.1f TFLOP/s  %.0f GB/s" % (ms, 2*M*K*N/ms/1e9, 4*(M*K+M*N)/ms/1e6))
