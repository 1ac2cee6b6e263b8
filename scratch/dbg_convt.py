import ctypes as C, torch, sys
sys.path.insert(0, '.')
from tinyrecurrentunet_b200 import _lib as L
fn = L.lib.tru_debug_convt_bwd_data
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]*3 + [C.c_int]*7 + [C.c_void_p]
torch.manual_seed(0)
for (BT, Ls, Cin, Cout, k, s) in [(12,64,64,64,5,2),(12,128,64,64,3,1),(12,128,8,8,5,2),(12,64,64,64,3,2),(12,64,64,64,5,1),(12,64,32,32,5,2),(12,64,64,32,5,2),(12,64,32,64,5,2),(3,64,64,64,5,2),(12,64,64,64,4,2)]:
    p = s//2
    Lout = (Ls-1)*s - 2*p + k
    x = torch.randn(BT, Cin, Ls, device='cuda', requires_grad=True)
    w = torch.randn(Cin, Cout, k, device='cuda')
    y = torch.nn.functional.conv_transpose1d(x, w, stride=s, padding=p)
    dy = torch.randn_like(y)
    y.backward(dy)
    ref = x.grad.transpose(1,2).contiguous()
    dycl = dy.transpose(1,2).contiguous()
    dx = torch.empty(BT, Ls, Cin, device='cuda')
    L.check(fn(dycl.data_ptr(), w.data_ptr(), dx.data_ptr(), BT, Ls, Lout, Cin, Cout, k, s, None))
    torch.cuda.synchronize()
    print((BT,Ls,Cin,Cout,k,s), "rel err", ((dx-ref).abs().max()/ref.abs().max()).item())
