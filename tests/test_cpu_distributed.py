"""CPU, gloo, world_size 2: the drop-in gradient all-reduce - state broadcast from
rank 0, hook-driven reduction after backward on the flat gradient bucket (one collective), the logging
scalar riding in the bucket's tail, and the loud failure for gradients that are not one flat buffer."""
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
        flat = torch.zeros(tot + 4)                        # + distributed.LOSS_TAIL, like network._TRUNetFn.backward
        Flat.owner._tru_flat_grad = flat
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
            Flat.owner = self
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
        loss = net(x)
        if not flat:
            # gradients that are separate tensors: there is no flatten / copy-back path any more - loud failure
            # (raised on every rank before any collective is issued, so nothing hangs)
            try:
                loss.backward()
            except RuntimeError as e:
                out[flat] = "not views of one flat buffer" in str(e)
            continue
        D.attach_loss(net, torch.tensor(float(rank)))      # train.py:133's logging scalar rides in the bucket
        n_calls = []
        real = dist.all_reduce
        dist.all_reduce = lambda t, *a, **k: (n_calls.append(t.numel()), real(t, *a, **k))[1]
        loss.backward()
        dist.all_reduce = real
        g = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        assert n_calls == [net._tru_flat_grad.numel()], n_calls          # ONE collective, gradients + tail
        assert abs(net.reduced_loss.item() - 0.5) < 1e-6
        out[flat] = (w.numpy().copy(), g.numpy().copy(), True)   # numpy: no fd passing after the worker exits
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
    w0, g0, c0 = res[0][True]
    w1, g1, _ = res[1][True]
    w0, g0, w1, g1 = (torch.from_numpy(a) for a in (w0, g0, w1, g1))
    assert torch.equal(w0, w1)                         # state broadcast from rank 0
    assert torch.equal(g0, g1)                         # identical averaged gradients on both ranks
    # mean over ranks of 2*p*sum(x): sum(x) = 4 on rank 0, 8 on rank 1 -> 2*p*6
    torch.testing.assert_close(g0, 2 * w0 * 6.0)
    assert c0 is True
    assert res[0][False] is True and res[1][False] is True


def test_flat_buffer_view_starts_at_the_first_gradient_and_rejects_gaps():
    """the all-reduce view must not cover foreign data in front of or between the gradients."""
    sys.path.insert(0, ROOT)
    from tinyrecurrentunet_b200 import distributed as D
    store = torch.arange(64, dtype=torch.float32)
    g = [store[8:14].view(2, 3), store[16:20], store[20:27]]          # padding of 2 floats after the first slice
    flat = D._single_flat_buffer(g)
    assert flat.data_ptr() == g[0].data_ptr() and flat.numel() == 19 and flat[0].item() == 8.0
    assert D._single_flat_buffer([store[8:14], store[24:30]]) is None          # a gap of 10 floats: foreign data
    assert D._single_flat_buffer([store[8:14], torch.zeros(4)]) is None        # different storages
    assert D._single_flat_buffer([store[16:20], store[8:14]]) is None          # not ascending
