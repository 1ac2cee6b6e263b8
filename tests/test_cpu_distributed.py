"""CPU, gloo, world_size 2: the drop-in gradient all-reduce - state broadcast from
rank 0, hook-driven reduction after backward, flat-bucket fast path vs the reference path."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Flat(torch.autograd.Function):
    """Mimics network._TRUNetFn.backward: all grads are views of ONE flat buffer."""
    @staticmethod
    def forward(ctx, x, *params):
        ctx.save_for_backward(x, *params)
        return sum((p * p).sum() for p in params) * x.sum()

    @staticmethod
    def backward(ctx, g):
        x, *params = ctx.saved_tensors
        sizes = [p.numel() for p in params]
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot)
            tot += (n + 3) // 4 * 4
        flat = torch.zeros(tot)
        for o, n, p in zip(offs, sizes, params):
            flat[o:o + n] = (2 * p * x.sum() * g).reshape(-1)
        return (None,) + tuple(flat[o:o + n].view(p.shape) for o, n, p in zip(offs, sizes, params))


class Toy(torch.nn.Module):
    def __init__(self, flat):
        super().__init__()
        self.a = torch.nn.Parameter(torch.randn(5, 3))
        self.b = torch.nn.Parameter(torch.randn(7))
        self.bn = torch.nn.BatchNorm1d(3)
        self.flat = flat

    def forward(self, x):
        if self.flat:
            return Flat.apply(x, self.a, self.b, self.bn.weight, self.bn.bias)
        return sum((p * p).sum() for p in self.parameters()) * x.sum()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from tinyrecurrentunet_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    D.init_distributed(rank, world, "g", "gloo", "tcp://127.0.0.1:%d" % port)
    out = {}
    for flat in (True, False):
        torch.manual_seed(100 + rank)                      # different weights per rank before the broadcast
        net = Toy(flat)
        same = D.apply_gradient_allreduce(net)
        assert same is net                                 # no wrapper class (distributed.py:96-99)
        w = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        x = torch.full((4,), float(rank + 1))
        net(x).backward()
        g = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        calls = None
        if flat:
            bufs = [p.grad for p in net.parameters()]
            calls = D._single_flat_buffer([b.data for b in bufs]) is not None
        out[flat] = (w.numpy().copy(), g.numpy().copy(), calls)   # numpy: no fd passing after the worker exits
        loss_mean = D.reduce_tensor(torch.tensor(float(rank)), world)
        assert abs(loss_mean.item() - 0.5) < 1e-6
    q.put((rank, out))
    dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for flat in (True, False):
        w0, g0, c0 = res[0][flat]
        w1, g1, _ = res[1][flat]
        w0, g0, w1, g1 = (torch.from_numpy(a) for a in (w0, g0, w1, g1))
        assert torch.equal(w0, w1)                         # state broadcast from rank 0
        assert torch.equal(g0, g1)                         # identical averaged gradients on both ranks
        # mean over ranks of 2*p*sum(x): sum(x) = 4 on rank 0, 8 on rank 1 -> 2*p*6
        torch.testing.assert_close(g0, 2 * w0 * 6.0)
        if flat:
            assert c0 is True                              # the one-call flat-bucket path was taken
