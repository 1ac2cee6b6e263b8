"""GPU parity of the individual GEMM-shaped kernels through their C-ABI test entry points, and a
full-size property test of the whole training step.

Tolerances: <= 1e-4 relative on forward values (north_star), <= 1e-3 on gradients.  The tensor-core
kernels run 3xTF32 (fp32-accurate); a weight gradient additionally accumulates ~14,000 rows per CTA
in one TMEM accumulator, whose truncating adds bias the sum by up to ~1e-4 at the full batch size
(measured 1.1e-4 at M = 2,052,096), which is why that case is checked against 3e-4."""
import ctypes as C

import pytest
import torch

from oracle import tru_oracle as O

pytestmark = pytest.mark.gpu


def _lib():
    from tinyrecurrentunet_b200 import _lib as L
    return L


def _pw(M, K, N, affine, stats, use_tc=1):
    L = _lib()
    fn = L.lib.tru_debug_pw
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 7 + [C.c_int] * 4 + [C.c_void_p]
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    p0 = (torch.rand(K, device="cuda") + 0.5) if affine else None
    p2 = torch.randn(K, device="cuda") if affine else None
    out = torch.full((M, N), float("nan"), device="cuda")
    st = torch.zeros(2 * N, device="cuda", dtype=torch.float64) if stats else None
    L.check(fn(x.data_ptr(), p0.data_ptr() if affine else None, p2.data_ptr() if affine else None, w.data_ptr(),
               b.data_ptr(), out.data_ptr(), st.data_ptr() if stats else None, M, K, N, use_tc, None), "debug_pw")
    torch.cuda.synchronize()
    a = x.double()
    if affine:
        a = torch.relu(a * p0.double() + p2.double())
    ref = a @ w.double().t() + b.double()
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    serr = 0.0
    if stats:
        s1, s2 = ref.sum(0), (ref * ref).sum(0)
        serr = max(((st[:N] - s1).abs().max() / s1.abs().max()).item(), ((st[N:] - s2).abs().max() / s2.abs().max()).item())
    return err, serr


# M = 64 and M = 128 accumulators, ragged last tile, channel counts that need zero padding, K split into passes
@pytest.mark.parametrize("M,K,N", [(1000, 128, 8), (3000, 8, 64), (5000, 40, 8), (128, 32, 32), (128 * 300 + 17, 128, 128),
                                   (5000, 192, 64), (4097, 64, 384), (3000, 128, 192), (777, 384, 128), (2048, 320, 64)])
def test_tensor_core_pointwise_conv(M, K, N):
    torch.manual_seed(M + K + N)
    for affine, stats in ((False, False), (True, True)):
        err, serr = _pw(M, K, N, affine, stats)
        assert err <= 1e-5, (affine, err)                 # 10x inside the 1e-4 budget of a single layer
        assert serr <= 1e-5, (affine, serr)


# The transposed convs of the decoder (network.py:67-73): k3 s1 (dec2, dec4), k5 s2 (dec1, dec3), k3 s2 (dec0); 64 -> 64 channels.
# Forward and data gradient run as TAP-SHARED launches (one staged tile, one row-shifted descriptor per tap); frame counts that
# put tile boundaries at every position of a frame, and ragged last tiles.
@pytest.mark.parametrize("BT,L,k,s", [(37, 64, 3, 1), (9, 128, 3, 1), (41, 32, 5, 2), (23, 64, 5, 2), (50, 16, 3, 2), (3, 5, 5, 2),
                                      (252, 64, 5, 2), (252, 32, 5, 2), (1002, 64, 5, 2), (1002, 128, 3, 1)])
def test_transposed_conv_forward_and_data_gradient(BT, L, k, s):
    L_ = _lib()
    torch.manual_seed(BT * 7 + L + k + s)
    Cin = Cout = 64
    pad = s // 2
    Lout = (L - 1) * s - 2 * pad + k
    x = torch.randn(BT, L, Cin, device="cuda")
    w = torch.randn(Cin, Cout, k, device="cuda") / (Cin * k) ** 0.5
    b = torch.randn(Cout, device="cuda")
    ref = torch.nn.functional.conv_transpose1d(x.double().transpose(1, 2), w.double(), b.double(), stride=s, padding=pad)
    ref = ref.transpose(1, 2).contiguous()                                   # (BT, Lout, Cout)
    assert ref.shape == (BT, Lout, Cout)
    fwd = L_.lib.tru_debug_convt_fwd
    fwd.restype = C.c_int
    fwd.argtypes = [C.c_void_p] * 5 + [C.c_int] * 7 + [C.c_void_p]
    out = torch.full((BT, Lout, Cout), float("nan"), device="cuda")
    st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    L_.check(fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), st.data_ptr(), BT, L, Lout, Cin, Cout, k, s, None), "convt_fwd")
    torch.cuda.synchronize()
    assert ((out.double() - ref).abs().max() / ref.abs().max()).item() <= 1e-5
    s1, s2 = ref.sum((0, 1)), (ref * ref).sum((0, 1))
    assert ((st[:Cout] - s1).abs().max() / s1.abs().max()).item() <= 1e-5
    assert ((st[Cout:] - s2).abs().max() / s2.abs().max()).item() <= 1e-5
    # data gradient: dx = conv1d(dz, w) (the adjoint), both launch forms, without and with the BatchNorm-backward affine on load
    dy = torch.randn(BT, Lout, Cout, device="cuda")
    z = torch.randn(BT, Lout, Cout, device="cuda")
    q0 = torch.rand(Cout, device="cuda") + 0.5
    q1 = torch.randn(Cout, device="cuda") * 0.1
    q2 = torch.randn(Cout, device="cuda") * 0.1
    bwd = L_.lib.tru_debug_convt_bwd_data
    bwd.restype = C.c_int
    bwd.argtypes = [C.c_void_p] * 13 + [C.c_int] * 8 + [C.c_void_p]
    zm = torch.randn(BT, L, Cin, device="cuda")                     # pre-BN activation of the layer receiving the gradient
    mp0, mp2 = torch.rand(Cin, device="cuda") + 0.5, torch.randn(Cin, device="cuda") * 0.3
    bmean, binv = torch.randn(Cin, device="cuda") * 0.1, torch.rand(Cin, device="cuda") + 0.5
    ptr = lambda t, on=True: t.data_ptr() if on else None
    for bn in (False, True):
        dz = (q0 * dy + q1 * z + q2).double() if bn else dy.double()
        dref = torch.nn.functional.conv1d(dz.transpose(1, 2), w.double(), stride=s, padding=pad).transpose(1, 2)
        assert dref.shape == (BT, L, Cin)
        for shared in (0, 1):
            for masked in (False, True):
                want = dref * (zm.double() * mp0.double() + mp2.double() > 0) if masked else dref
                dx = torch.full((BT, L, Cin), float("nan"), device="cuda")
                bst = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
                L_.check(bwd(dy.data_ptr(), ptr(z, bn), ptr(q0, bn), ptr(q1, bn), ptr(q2, bn), w.data_ptr(), dx.data_ptr(),
                             ptr(zm, masked), ptr(mp0, masked), ptr(mp2, masked), ptr(bmean, masked), ptr(binv, masked),
                             ptr(bst, masked), BT, L, Lout, Cin, Cout, k, s, shared, None), "convt_bwd_data")
                torch.cuda.synchronize()
                err = ((dx.double() - want).abs().max() / want.abs().max()).item()
                assert err <= 1e-5, (bn, shared, masked, err)
                if masked:
                    s1 = want.sum((0, 1))
                    s2 = (want * (zm.double() - bmean.double())).sum((0, 1)) * binv.double()
                    e1 = ((bst[:Cin] - s1).abs().max() / s1.abs().max()).item()
                    e2 = ((bst[Cin:] - s2).abs().max() / s2.abs().max()).item()
                    assert e1 <= 1e-4 and e2 <= 1e-4, (bn, shared, e1, e2)


def _wgrad_stream(M, Lq, Cc, N, bn):
    L = _lib()
    fn = L.lib.tru_debug_wgrad_stream
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 8 + [C.c_int] * 4 + [C.c_void_p]
    a = torch.randn(M, Cc, device="cuda")
    dy = torch.randn(M, N, device="cuda")
    z = torch.randn(M, N, device="cuda")
    q0 = torch.rand(N, device="cuda") + 0.5
    q1 = torch.randn(N, device="cuda") * 0.1
    q2 = torch.randn(N, device="cuda") * 0.1
    dw = torch.zeros(N, Cc, device="cuda")
    db = torch.zeros(N, device="cuda")
    nul = None
    L.check(fn(a.data_ptr(), dy.data_ptr(), z.data_ptr() if bn else nul, q0.data_ptr() if bn else nul,
               q1.data_ptr() if bn else nul, q2.data_ptr() if bn else nul, dw.data_ptr(), db.data_ptr(), M, Lq, Cc, N, None),
            "debug_wgrad_stream")
    torch.cuda.synchronize()
    dz = (q0 * dy + q1 * z + q2).double() if bn else dy.double()
    ref = dz.t() @ a.double()
    e_w = ((dw.double() - ref).abs().max() / ref.abs().max()).item()
    e_b = ((db.double() - dz.sum(0)).abs().max() / dz.sum(0).abs().max()).item()
    return e_w, e_b


@pytest.mark.parametrize("M,Lq,Cc,N,tol", [(16 * 37, 16, 128, 128, 1e-5), (64 * 501, 64, 64, 128, 1e-5), (128 * 96, 128, 128, 64, 1e-5),
                                           (32 * 50, 32, 128, 8, 1e-5), (128 * 16032, 128, 128, 128, 3e-4)])
def test_streaming_weight_gradient(M, Lq, Cc, N, tol):
    torch.manual_seed(M + Cc)
    for bn in (False, True):
        e_w, e_b = _wgrad_stream(M, Lq, Cc, N, bn)
        assert e_w <= tol, (bn, e_w)
        assert e_b <= 1e-4, (bn, e_b)


def test_full_size_step_equals_small_batch_replicated():
    """BASELINE.json configs[1] size (32 clips x 4 s) through a size-independent property: a batch made of
    8 copies of 4 distinct clips has exactly the BatchNorm statistics, loss and (mean-reduced) gradients
    of the 4-clip batch, which the CPU oracle can evaluate."""
    from tinyrecurrentunet_b200 import network, stft_loss, util
    torch.manual_seed(0)
    ref = O.randomize_bn(O.TRUNet()).train()
    net = network.TRUNet(3, 64, 3, 128, [5, 3], [2, 1], 192)
    net.load_state_dict(ref.state_dict())
    net = net.cuda().train()
    clean, noisy = O.synthetic_batch(4, n=64000, first=11)
    loss_ref, d_ref, _ = O.loss_fn(ref, clean, noisy)
    loss_ref.backward()
    mr = stft_loss.MultiResolutionSTFTLoss(fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240],
                                           win_lengths=[240, 600, 1200], sc_lambda=0.5, mag_lambda=0.5).cuda()
    loss, d = util.loss_fn(net, (clean.repeat(8, 1).cuda(), noisy.repeat(8, 1).cuda()), ell_p=1, ell_p_lambda=1,
                           stft_lambda=1, mrstftloss=mr)
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()) <= 1e-4
    for k in ("l1", "stft_sc", "stft_mag"):
        assert abs(d[k].item() - d_ref[k].item()) / abs(d_ref[k].item()) <= 1e-4, k
    gref = dict(ref.named_parameters())
    num = den = 0.0
    for k, p in net.named_parameters():
        assert torch.isfinite(p.grad).all(), k
        num += (p.grad.cpu().double() - gref[k].grad.double()).pow(2).sum().item()
        den += gref[k].grad.double().pow(2).sum().item()
    # end to end a few ReLU masks differ between the two front ends (see test_loss_fn_end_to_end_matches_oracle)
    assert (num / den) ** 0.5 <= 2e-2, (num / den) ** 0.5
