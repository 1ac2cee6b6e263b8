"""GPU: the data-parallel gradient path of the REAL TRUNet (distributed.py:95-147, train.py:133).

* NCCL, world size 1, in process: the hook-driven reduction takes the one-call flat-bucket path (one all_reduce of the
  whole 1.5 MB buffer incl. the loss tail), and leaves the gradients of a plain backward.
* gloo, world size 2, two processes sharing cuda:0 (NCCL refuses two ranks on one device; gloo reduces CUDA tensors through
  the host): after the all-reduce every rank holds the MEAN of the per-shard gradients, where each shard ran its own
  BatchNorm statistics - i.e. DP-2 == one process evaluating the two shards one after the other."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tru_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STFT = dict(fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240], win_lengths=[240, 600, 1200], sc_lambda=0.5, mag_lambda=0.5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_net(seed):
    from tinyrecurrentunet_b200 import network
    torch.manual_seed(seed)
    ref = O.randomize_bn(O.TRUNet(), seed)
    net = network.TRUNet(3, 64, 3, 128, [5, 3], [2, 1], 192)
    net.load_state_dict(ref.state_dict())
    return net.cuda().train()


def _backward(net, mr, clean, noisy, attach=False):
    from tinyrecurrentunet_b200 import distributed as D, util
    for p in net.parameters():
        p.grad = None
    loss, _ = util.loss_fn(net, (clean, noisy), ell_p=1, ell_p_lambda=1, stft_lambda=1, mrstftloss=mr)
    if attach:
        D.attach_loss(net, loss)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach(), torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()


def test_nccl_world1_real_trunet_takes_the_one_call_flat_path(monkeypatch):
    from tinyrecurrentunet_b200 import distributed as D, stft_loss
    mr = stft_loss.MultiResolutionSTFTLoss(**STFT).cuda()
    clean, noisy = O.synthetic_batch(2, n=128 * 40, first=3)
    clean, noisy = clean.cuda(), noisy.cuda()
    net = _make_net(2)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    loss0, g0 = _backward(net, mr, clean, noisy)                       # no hooks, no process group
    net.load_state_dict(state)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % _free_port(), world_size=1, rank=0)
    try:
        assert D.apply_gradient_allreduce(net) is net
        calls = []
        real = dist.all_reduce

        def counting(t, *a, **k):
            calls.append((t.numel(), t.data_ptr(), k.get("op", a[0] if a else None)))
            return real(t, *a, **k)
        monkeypatch.setattr(dist, "all_reduce", counting)
        loss1, g1 = _backward(net, mr, clean, noisy, attach=True)
        nparam = sum(p.numel() for p in net.parameters())
        assert nparam == 381472
        assert len(calls) == 1, calls                                  # ONE collective for 108 gradients + the loss scalar
        numel, ptr, op = calls[0]
        first = next(net.parameters())
        pad = sum((p.numel() + 3) // 4 * 4 for p in net.parameters())
        assert numel == pad + D.LOSS_TAIL and ptr == net._tru_flat_grad.data_ptr()
        assert op == dist.ReduceOp.AVG
        assert net._tru_flat_grad.data_ptr() <= first.grad.data_ptr() < ptr + 4 * numel
        # world size 1: the mean is the gradient itself (two runs agree to a few ulp, not bit for bit: a few fp32 atomics)
        assert (g1 - g0).abs().max().item() <= 1e-5 * g0.abs().max().item()
        assert abs(net.reduced_loss.item() - loss1.item()) == 0.0      # the piggy-backed logging scalar (train.py:133)
        assert abs(loss1.item() - loss0.item()) <= 1e-6 * abs(loss0.item())
    finally:
        dist.destroy_process_group()


def _dp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from tinyrecurrentunet_b200 import distributed as D, stft_loss
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, world_size=world, rank=rank)
    mr = stft_loss.MultiResolutionSTFTLoss(**STFT).cuda()
    net = _make_net(100 + rank)                                        # different weights per rank before the broadcast
    D.apply_gradient_allreduce(net)
    w = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    clean, noisy = O.synthetic_batch(2, n=128 * 40, first=20 + 2 * rank)   # this rank's shard
    loss, g = _backward(net, mr, clean.cuda(), noisy.cuda(), attach=True)
    q.put((rank, w.numpy().copy(), g.cpu().numpy().copy(), loss.item(), net.reduced_loss.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp2_real_trunet_equals_mean_of_per_shard_gradients():
    from tinyrecurrentunet_b200 import stft_loss
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = q.get(timeout=600)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w0, g0, l0, rl0 = res[0]
    w1, g1, l1, rl1 = res[1]
    w0, g0, w1, g1 = (torch.from_numpy(a) for a in (w0, g0, w1, g1))
    assert torch.equal(w0, w1)                                         # state broadcast from rank 0
    assert torch.equal(g0, g1)                                         # identical averaged gradients on both ranks
    assert abs(rl0 - 0.5 * (l0 + l1)) <= 1e-6 * abs(rl0) and rl0 == rl1
    # single process, the two shards one after the other with rank 0's weights (per-shard BatchNorm statistics)
    mr = stft_loss.MultiResolutionSTFTLoss(**STFT).cuda()
    net = _make_net(100)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    shard = []
    for r in range(2):
        net.load_state_dict(state)
        clean, noisy = O.synthetic_batch(2, n=128 * 40, first=20 + 2 * r)
        shard.append(_backward(net, mr, clean.cuda(), noisy.cuda())[1].cpu())
    want = (shard[0] + shard[1]) / 2
    err = ((g0 - want).abs().max() / want.abs().max()).item()
    print("DP-2 vs mean of per-shard gradients: max relative difference %.3e" % err)
    assert err <= 2e-6, err
