import ctypes as C, torch, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import test_gpu_network as tg
from oracle import tru_oracle as O
ref, net = tg.make_pair(2)
B, T = 2, 6
x = tg.feats_like(B, T, 5)
net.train(); net._debug_keep_ws = True
w = torch.randn(B, T, 8, 257)
y = net(x.cuda()); (y * w.cuda()).sum().backward()
small = tg.gpu_buffer(net, "small", 0, (23, 7, 128))
DEC_K=[3,5,3,5,3,5]; DEC_S=[2,2,1,2,1,2]; DEC_LP=[16,32,64,64,128,128]; DEC_LT=[31,65,66,129,130,257]
sd = net.state_dict()
for d in (4,3,2,1,0):
    BT=B*T; Co=64
    dYt = tg.gpu_buffer(net, "dZDt", d, (BT, DEC_LT[d], Co)); Zt = tg.gpu_buffer(net, "ZDt", d, (BT, DEC_LT[d], Co))
    bn_t = 11+2*d; bn_p = 10+2*d
    q0,q1,q2 = (small[bn_t, j, :Co] for j in (4,5,6))
    dZt = q0*dYt + q1*Zt + q2
    cls = "FirstTrCNN" if d==0 else "TrCNN"
    W = sd["decoder.%d.%s.3.weight"%(d,cls)].cpu()
    s=DEC_S[d]; k=DEC_K[d]; p=s//2
    # data grad of conv_transpose1d = conv1d(dZ, W) with stride s, padding p
    dA = torch.nn.functional.conv1d(dZt.transpose(1,2), W, stride=s, padding=p)   # (BT, Cin, L)
    dA = dA.transpose(1,2)[:, :DEC_LP[d]]
    Zp = tg.gpu_buffer(net, "ZDp", d, (BT, DEC_LP[d], Co))
    p0,p2 = small[bn_p,0,:Co], small[bn_p,1,:Co]
    exp = dA * ((Zp*p0+p2) > 0)
    mine = tg.gpu_buffer(net, "dZDp", d, (BT, DEC_LP[d], Co))
    print("dec", d, "k", k, "s", s, "bwd-data self-consistency rel err:", tg.rel(mine, exp), " unmasked-vs-mine nonzero frac", (mine!=0).float().mean().item(), (exp!=0).float().mean().item())

# ---- oracle side-by-side for decoder 3 ----
ref.train()
cap = {}
seq = ref.decoder[3].TrCNN
def _pre(m, inp):
    inp[0].register_hook(lambda g: cap.__setitem__("dA", g.detach()))
    return None
def _post(m, inp, out):
    out.register_hook(lambda g: cap.__setitem__("dZt", g.detach()))
    return None
seq[3].register_forward_pre_hook(_pre)
seq[3].register_forward_hook(_post)
yr = ref(x); (yr * w).sum().backward()
d = 3; BT = B*T; Co = 64
dYt = tg.gpu_buffer(net, "dZDt", d, (BT, DEC_LT[d], Co)); Zt = tg.gpu_buffer(net, "ZDt", d, (BT, DEC_LT[d], Co))
q0,q1,q2 = (small[11+2*d, j, :Co] for j in (4,5,6))
dZt = q0*dYt + q1*Zt + q2
print("dZ(ZDt3) mine vs oracle grad of convT output:", tg.rel(dZt, cap["dZt"].transpose(1,2)))
W = sd["decoder.3.TrCNN.3.weight"].cpu()
dA = torch.nn.functional.conv1d(dZt.transpose(1,2), W, stride=2, padding=1).transpose(1,2)
print("dA shapes", dA.shape, cap["dA"].shape, " dA vs oracle grad of convT input:", tg.rel(dA[:, :64], cap["dA"].transpose(1,2)))
mine = tg.gpu_buffer(net, "dZDp", d, (BT, 64, Co))
Zp = tg.gpu_buffer(net, "ZDp", d, (BT, 64, Co))
p0,p2 = small[10+2*d,0,:Co], small[10+2*d,1,:Co]
print("dY mine vs mask*oracle dA:", tg.rel(mine, cap["dA"].transpose(1,2) * ((Zp*p0+p2) > 0)))
def _pre2(m, inp):
    cap["act3"] = inp[0].detach().clone()
    return None
h = seq[3].register_forward_pre_hook(_pre2)
ref.zero_grad()
yr = ref(x); (yr * w).sum().backward()
omask = cap["act3"].transpose(1,2) > 0
mmask = (Zp*p0+p2) > 0
print("mask mismatches:", (omask != mmask).sum().item(), "of", omask.numel())
print("sum my dY vs oracle dbeta:", tg.rel(mine.sum((0,1)), seq[1].bias.grad), " sum(omask*dA) vs dbeta", tg.rel((cap["dA"].transpose(1,2)*omask).sum((0,1)), seq[1].bias.grad))
g = dict(net.named_parameters())
for k in ("decoder.3.TrCNN.1.bias","decoder.3.TrCNN.1.weight","decoder.3.TrCNN.0.weight","decoder.2.TrCNN.3.weight"):
    print(k, tg.rel(g[k].grad, dict(ref.named_parameters())[k].grad))
