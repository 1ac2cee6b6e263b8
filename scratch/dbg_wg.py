import ctypes as C, torch, sys
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_wgrad
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*4 + [C.c_int]*4 + [C.c_void_p]
torch.manual_seed(0)
def run(M, Cc, N, use_tc, mode="rand"):
    if mode == "ones":
        a = torch.ones(M, Cc, device='cuda'); z = torch.ones(M, N, device='cuda')
    elif mode == "idx":
        a = torch.arange(Cc, device='cuda').float().repeat(M,1); z = torch.ones(M, N, device='cuda')
    else:
        a = torch.randn(M, Cc, device='cuda'); z = torch.randn(M, N, device='cuda')
    dw = torch.zeros(N, Cc, device='cuda'); db = torch.zeros(N, device='cuda')
    L.check(fn(a.data_ptr(), z.data_ptr(), dw.data_ptr(), db.data_ptr(), M, Cc, N, use_tc, None))
    torch.cuda.synchronize()
    ref = z.double().t() @ a.double()
    return dw, ref, db, z.double().sum(0)
dw, ref, db, dbr = run(64, 128, 128, 1, "ones")
print("ones: dw[0,:8]", dw[0,:8].tolist(), "expected", ref[0,0].item(), "nonzero frac", (dw!=0).float().mean().item())
dw, ref, db, dbr = run(64, 128, 128, 1, "idx")
print("idx: dw[0,:8]", dw[0,:8].tolist(), " dw[1,:4]", dw[1,:4].tolist(), "expected", ref[0,:8].tolist())
for shape in [(5000,128,8),(5000,8,8),(32,128,128),(64,128,128),(1000,128,128),(5000,64,128),(5000,128,64),(4097,64,64),(3000,128,384),(3000,64,192),(100000,128,128)]:
    for tc in (1,0):
        dw, ref, db, dbr = run(*shape, tc)
        print(shape, "tc" if tc else "simt", "dW rel err", ((dw.double()-ref).abs().max()/ref.abs().max()).item(), "db rel err", ((db.double()-dbr).abs().max()/dbr.abs().max()).item(), flush=True)
