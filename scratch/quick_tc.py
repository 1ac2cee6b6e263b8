import ctypes as C, torch, sys
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_pw
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*7 + [C.c_int]*4 + [C.c_void_p]
torch.manual_seed(0)
def run(M, K, N, affine, stats):
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') / K**0.5; b = torch.randn(N, device="cuda")
    p0 = (torch.rand(K, device='cuda') + 0.5) if affine else None
    p2 = torch.randn(K, device='cuda') if affine else None
    out = torch.full((M, N), float('nan'), device='cuda')
    st = torch.zeros(2*N, device='cuda', dtype=torch.float64) if stats else None
    L.check(fn(x.data_ptr(), p0.data_ptr() if affine else None, p2.data_ptr() if affine else None, w.data_ptr(), b.data_ptr(),
            out.data_ptr(), st.data_ptr() if stats else None, M, K, N, 1, None), "debug_pw")
    try:
        torch.cuda.synchronize()
    finally:
        buf = (C.c_uint * 1028)()
        if L.lib.tru_debug_read_mbar(buf, 1028):
            n = buf[0]
            print("mbar timeouts:", n)
            import collections
            cnt = collections.Counter()
            for i in range(min(n, 255)):
                bb, t, tag, par = buf[4*i+4:4*i+8]
                cnt[(t // 32, tag, par)] += 1
            for k, v in sorted(cnt.items()): print("   warp %d tag %d parity %d : %d" % (k[0], k[1], k[2], v))
    a = x.double()
    if affine: a = torch.relu(a * p0.double() + p2.double())
    ref = a @ w.double().t() + b.double()
    return ((out.double() - ref).abs().max() / ref.abs().max()).item()
for shape in [(128*300+17,128,128),(128*700+5,192,64),(128*1000,64,128),(200000,320,64),(100000,384,128),(150001,128,192)]:
    for affine, stats in ((False, False), (True, True)):
        print(shape, affine, flush=True)
        print("   err", run(*shape, affine, stats), flush=True)
