"""GPU parity of the optimizer step: optim.FlatAdamW (one C call: tru_flat_adamw_step) against
what train.py:138-140 runs -- nn.utils.clip_grad_norm_ + torch.optim.AdamW.step -- on the same parameters and gradients.
The checker here is torch itself (the reference's third-party dependency for this step), on the same device.

Tolerance: the kernel follows torch's multi-tensor path op for op, so parameters and moments agree to a few fp32 ulps;
the tests allow 2e-6 relative to the tensor's largest magnitude after 25 updates (bit-comparable, not bit-identical:
torch's own single-tensor / foreach / fused paths differ from each other by the same amount)."""
import copy
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-6


def close(a, b, tol=TOL):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item() <= tol


def make_nets(seed=0):
    from tinyrecurrentunet_b200 import network
    torch.manual_seed(seed)
    net = network.TRUNet().cuda()
    twin = copy.deepcopy(net)
    return net, twin


def set_grads(net, twin, gen, scale=1.0):
    for p, q in zip(net.parameters(), twin.parameters()):
        g = torch.randn(p.shape, device="cuda", generator=gen) * scale
        p.grad = g.clone()
        q.grad = g.clone()


def assert_same_state(net, twin, opt, ref, tol=TOL):
    for (name, p), q in zip(net.named_parameters(), twin.parameters()):
        assert close(p, q, tol), name
        assert close(opt.state[p]["exp_avg"], ref.state[q]["exp_avg"], tol), name
        assert close(opt.state[p]["exp_avg_sq"], ref.state[q]["exp_avg_sq"], tol), name
        assert float(opt.state[p]["step"]) == float(ref.state[q]["step"])


@pytest.mark.parametrize("max_norm", [1e9, 3.0])
def test_flat_adamw_matches_clip_grad_norm_plus_torch_adamw(max_norm):
    """25 iterations of train.py:138-140 with the learning-rate schedule stepping both optimizers; max_norm = 1e9 is the
    reference's setting (norm only), 3.0 makes every step clip (random grads of 381,472 elements have norm ~ 600)."""
    from tinyrecurrentunet_b200 import optim, util
    net, twin = make_nets()
    opt = optim.FlatAdamW(net.parameters(), lr=4e-4, max_grad_norm=max_norm)
    ref = torch.optim.AdamW(twin.parameters(), lr=4e-4)
    sch = util.LinearWarmupCosineDecay(opt, lr_max=4e-4, n_iter=40, iteration=0, divider=25, warmup_proportion=0.3)
    sch_ref = util.LinearWarmupCosineDecay(ref, lr_max=4e-4, n_iter=40, iteration=0, divider=25, warmup_proportion=0.3)
    gen = torch.Generator(device="cuda").manual_seed(7)
    for it in range(25):
        set_grads(net, twin, gen)
        want_norm = torch.nn.utils.clip_grad_norm_(twin.parameters(), max_norm)
        sch.step(); sch_ref.step()
        opt.step(); ref.step()
        assert abs(opt.grad_norm.item() - want_norm.item()) <= 2e-6 * want_norm.item(), it
    assert_same_state(net, twin, opt, ref)
    # the parameters are views of ONE flat buffer, in PARAM_ORDER, and the module still runs on them
    base = opt._flat["p"].data_ptr()
    assert all(p.data_ptr() == base + 4 * o for p, o in zip(net.parameters(), opt._flat["offs"]))
    net.eval()
    x = torch.randn(1, 6, 4, 257, device="cuda")
    twin.eval()
    assert close(net(x), twin(x), 1e-5)


def test_flat_adamw_on_the_gradients_of_a_real_backward():
    """Gradients delivered by TRUNet's backward are views of one flat buffer in the optimizer's layout: the step must use
    them in place (no gather copy) and agree with torch AdamW fed the same gradients."""
    from oracle import tru_oracle as O
    from tinyrecurrentunet_b200 import optim, stft_loss, util
    net, twin = make_nets(1)
    opt = optim.FlatAdamW(net.parameters(), lr=4e-4, max_grad_norm=1e9)
    ref = torch.optim.AdamW(twin.parameters(), lr=4e-4)
    mr = stft_loss.MultiResolutionSTFTLoss(fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240],
                                           win_lengths=[240, 600, 1200]).cuda()
    for it in range(3):
        clean, noisy = O.synthetic_batch(2, n=128 * 24, first=10 * it)
        opt.zero_grad(set_to_none=True)
        loss, _ = util.loss_fn(net, (clean.cuda(), noisy.cuda()), mrstftloss=mr)
        loss.backward()
        for p, q in zip(net.parameters(), twin.parameters()):
            q.grad = p.grad.clone()
        want_norm = torch.nn.utils.clip_grad_norm_(twin.parameters(), 1e9)
        opt.step(); ref.step()
        assert opt._flat["gather"] is None, "flat gradients were not used in place"
        assert abs(opt.grad_norm.item() - want_norm.item()) <= 2e-6 * want_norm.item()
        assert abs(optim.grad_norm(net.parameters()).item() - want_norm.item()) <= 2e-6 * want_norm.item()
    assert_same_state(net, twin, opt, ref)


def test_optimizer_state_dict_is_interchangeable_with_torch_adamw(tmp_path):
    """train.py:155-161 saves optimizer.state_dict(); train.py:85-87 loads it.  A FlatAdamW checkpoint must load into
    torch.optim.AdamW and the other way round, and training must continue identically."""
    from tinyrecurrentunet_b200 import optim
    net, twin = make_nets(2)
    opt = optim.FlatAdamW(net.parameters(), lr=4e-4)
    ref = torch.optim.AdamW(twin.parameters(), lr=4e-4)
    gen = torch.Generator(device="cuda").manual_seed(11)
    for _ in range(3):
        set_grads(net, twin, gen, 0.1)
        opt.step(); ref.step()
    # FlatAdamW -> file -> torch AdamW on a third copy of the model
    torch.save({"optimizer_state_dict": opt.state_dict(), "model_state_dict": net.state_dict()}, tmp_path / "3.pkl")
    ck = torch.load(tmp_path / "3.pkl", map_location="cpu")
    assert set(ck["optimizer_state_dict"]) == {"state", "param_groups"}
    assert set(ck["optimizer_state_dict"]["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    third = copy.deepcopy(twin)
    third.load_state_dict(ck["model_state_dict"])
    ref3 = torch.optim.AdamW(third.parameters(), lr=4e-4)
    ref3.load_state_dict(ck["optimizer_state_dict"])
    # torch AdamW -> FlatAdamW on a fourth copy
    from tinyrecurrentunet_b200 import network
    fourth = network.TRUNet().cuda()
    fourth.load_state_dict(twin.state_dict())
    opt4 = optim.FlatAdamW(fourth.parameters(), lr=4e-4)
    opt4.load_state_dict(copy.deepcopy(ref.state_dict()))
    for _ in range(3):
        gs = [torch.randn(p.shape, device="cuda", generator=gen) * 0.1 for p in net.parameters()]
        for model in (net, twin, third, fourth):
            for p, g in zip(model.parameters(), gs):
                p.grad = g.clone()
        for o in (opt, ref, ref3, opt4):
            o.step()
    assert_same_state(net, twin, opt, ref)
    assert_same_state(net, third, opt, ref3)
    assert_same_state(fourth, twin, opt4, ref)
    assert float(opt4.state[next(fourth.parameters())]["step"]) == 6.0


def test_flat_adamw_c_abi_argument_checks():
    from tinyrecurrentunet_b200 import _lib as L
    n = 64
    bufs = [torch.zeros(n, device="cuda") for _ in range(4)]
    desc = L.TruAdamWDesc(n, 1, 4e-4, 0.9, 0.999, 1e-8, 1e-2, 0.0)
    wsb = L.lib.tru_flat_adamw_workspace_bytes(C.byref(desc))
    ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    norm = torch.zeros((), device="cuda")

    def call(d, ws_bytes=wsb, p=bufs[0]):
        return L.lib.tru_flat_adamw_step(C.byref(d), p.data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(),
                                         norm.data_ptr(), ws.data_ptr(), ws_bytes, L.stream_ptr())
    assert call(desc) == 0
    assert call(L.TruAdamWDesc(n + 2, 1, 4e-4, 0.9, 0.999, 1e-8, 1e-2, 0.0)) == -1          # n not a multiple of 4
    assert call(L.TruAdamWDesc(n, 0, 4e-4, 0.9, 0.999, 1e-8, 1e-2, 0.0)) == -1              # step is 1-based
    assert call(L.TruAdamWDesc(n, 1, 4e-4, 1.0, 0.999, 1e-8, 1e-2, 0.0)) == -1              # beta1 out of range
    assert call(desc, ws_bytes=8) == -2                                                     # workspace too small
    assert call(desc, p=bufs[0][1:]) == -4                                                  # misaligned
    assert b"flat_adamw" in L.lib.tru_last_error()
    torch.cuda.synchronize()


def test_resume_from_checkpoint_continues_the_same_training(tmp_path):
    """train.py:70-104: four training iterations, with a checkpoint (util.save_checkpoint) after the second.  A fresh set
    of objects restored with util.load_checkpoint + LinearWarmupCosineDecay(iteration=n_iter) and fed the gradients of
    iterations 3 and 4 must land on bit-identical parameters: model, BatchNorm buffers, both moments, the step count
    and the learning-rate schedule all survive the round trip."""
    from oracle import tru_oracle as O
    from tinyrecurrentunet_b200 import network, optim, stft_loss, util
    mr = stft_loss.MultiResolutionSTFTLoss(fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240],
                                           win_lengths=[240, 600, 1200]).cuda()
    sched = dict(lr_max=4e-4, n_iter=10, divider=25, warmup_proportion=0.3)
    torch.manual_seed(5)
    net = network.TRUNet().cuda().train()
    opt = optim.FlatAdamW(net.parameters(), lr=4e-4, max_grad_norm=1e9)
    sch = util.LinearWarmupCosineDecay(opt, iteration=0, **sched)
    grads, lrs = [], []
    for i in range(4):
        clean, noisy = O.synthetic_batch(2, n=128 * 24, first=10 * i)
        opt.zero_grad()
        loss, _ = util.loss_fn(net, (clean.cuda(), noisy.cuda()), mrstftloss=mr)
        loss.backward()
        grads.append([p.grad.clone() for p in net.parameters()])
        lrs.append(sch.step())
        opt.step()
        if i == 1:
            util.save_checkpoint(str(tmp_path), i, net, opt, 5)
            at_save = {k: v.clone() for k, v in net.state_dict().items()}

    net_c = network.TRUNet().cuda().train()                # different random weights until the checkpoint is loaded
    opt_c = optim.FlatAdamW(net_c.parameters(), lr=4e-4, max_grad_norm=1e9)
    it, secs = util.load_checkpoint(str(tmp_path), "max", net_c, opt_c)
    assert (it, secs) == (1, 5)
    for k, v in net_c.state_dict().items():
        assert torch.equal(v, at_save[k]), k
    n_iter = it + 1                                        # train.py:98
    sch_c = util.LinearWarmupCosineDecay(opt_c, iteration=n_iter, **sched)
    for i in range(n_iter, 4):
        for p, g in zip(net_c.parameters(), grads[i]):
            p.grad = g.clone()
        assert sch_c.step() == lrs[i]
        opt_c.step()
    for (name, p), q in zip(net.named_parameters(), net_c.parameters()):
        assert torch.equal(p, q), name
        assert torch.equal(opt.state[p]["exp_avg"], opt_c.state[q]["exp_avg"]), name
        assert torch.equal(opt.state[p]["exp_avg_sq"], opt_c.state[q]["exp_avg_sq"]), name
    assert float(opt_c.state[next(net_c.parameters())]["step"]) == 4.0
