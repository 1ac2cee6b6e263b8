import ctypes as C, torch, sys, os
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_pw_bwd
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*14 + [C.c_int]*3 + [C.c_void_p]
torch.manual_seed(0)
def mk(M, K, N, extra=False):
    d = dict(dy=torch.randn(M, K, device='cuda'), z=torch.randn(M, K, device='cuda'), q0=torch.rand(K, device='cuda') + 0.5,
             q1=torch.randn(K, device='cuda') * 0.1, q2=torch.randn(K, device='cuda') * 0.1, w=torch.randn(K, N, device='cuda') / K**0.5,
             dx=torch.empty(M, N, device='cuda'), zm=torch.randn(M, N, device='cuda'), mp0=torch.rand(N, device='cuda') + 0.5,
             mp2=torch.randn(N, device='cuda') * 0.3, bmean=torch.randn(N, device='cuda') * 0.1, binv=torch.rand(N, device='cuda') + 0.5,
             bst=torch.zeros(2 * N, device='cuda', dtype=torch.float64), ex=torch.randn(M, N, device='cuda') if extra else None)
    return d
def call(d, M, K, N):
    L.check(fn(d['dy'].data_ptr(), d['z'].data_ptr(), d['q0'].data_ptr(), d['q1'].data_ptr(), d['q2'].data_ptr(), d['w'].data_ptr(),
               d['dx'].data_ptr(), d['zm'].data_ptr(), d['mp0'].data_ptr(), d['mp2'].data_ptr(), d['bmean'].data_ptr(), d['binv'].data_ptr(),
               d['bst'].data_ptr(), d['ex'].data_ptr() if d['ex'] is not None else None, M, K, N, None), "pw_bwd")
def check(M, K, N, extra):
    d = mk(M, K, N, extra); call(d, M, K, N); torch.cuda.synchronize()
    dz = d['dy'].double() * d['q0'].double() + d['z'].double() * d['q1'].double() + d['q2'].double()
    ref = dz @ d['w'].double()
    if extra: ref = ref + d['ex'].double()
    ref = ref * ((d['zm'].double() * d['mp0'].double() + d['mp2'].double()) > 0)
    err = ((d['dx'].double() - ref).abs().max() / ref.abs().max()).item()
    s1 = ref.sum(0); s2 = (ref * (d['zm'].double() - d['bmean'].double())).sum(0) * d['binv'].double()
    e2 = max(((d['bst'][:N] - s1).abs().max() / s1.abs().max()).item(), ((d['bst'][N:] - s2).abs().max() / s2.abs().max()).item())
    return err, e2
def t(M, K, N, flags=0, extra=False, n=5):
    d = mk(M, K, N, extra); L.lib.tru_debug_set_flags(flags)
    for _ in range(2): call(d, M, K, N)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): call(d, M, K, N)
    e1.record(); torch.cuda.synchronize(); L.lib.tru_debug_set_flags(0)
    return e0.elapsed_time(e1) / n
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":
        d = mk(2052096, 128, 128); call(d, 2052096, 128, 128); call(d, 2052096, 128, 128); torch.cuda.synchronize(); sys.exit(0)
    for (M, K, N, ex) in [(5000, 128, 128, False), (40001, 128, 128, True), (30000, 64, 64, False), (30000, 320, 64, True)]:
        print((M, K, N, ex), "err %.2e  bstats err %.2e" % check(M, K, N, ex), flush=True)
    for (M, K, N) in [(2052096, 128, 128), (2052096, 128, 64), (2052096, 64, 64)]:
        print("shape", (M, K, N), "ideal HBM ms %.3f" % (4 * (2 * M * K + 2 * M * N) / 6.5237e9))
        for f, nm in ((0, "full"), (2, "noLDG"), (8, "noEPI loads/stores"), (16, "no epilogue body")):
            print("   %-26s %.3f ms   with extra: %.3f ms" % (nm, t(M, K, N, f), t(M, K, N, f, True)), flush=True)
