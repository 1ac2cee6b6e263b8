"""Data-parallel gradient sync: drop-in for the reference's distributed.py:42-147
(init_distributed, apply_gradient_allreduce, reduce_tensor), SURVEY D12.

Same contract as the reference: the module is returned unchanged (no wrapper
class), state is broadcast from rank 0 once, and after every backward the gradients
are summed over ranks and divided by the world size, triggered by the reference's
own mechanism (per-parameter hook -> engine callback -> one reduction).

What is different is the data path.  TRUNet's backward (network._TRUNetFn) writes
all 108 gradients into ONE flat fp32 buffer (1,525,888 B) and hands autograd views
of it, so the reduction is a single in-place NCCL all-reduce of that buffer on the
backward stream: no flatten (torch.cat), no 108 copy-backs, one collective launch.
Modules whose grads are not views of one buffer take the reference path
(flatten -> all_reduce -> copy back)."""
import torch
import torch.distributed as dist
from torch.autograd import Variable


def reduce_tensor(tensor, num_gpus):
    """distributed.py:42-46."""
    rt = tensor.clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= num_gpus
    return rt


def init_distributed(rank, num_gpus, group_name, dist_backend, dist_url):
    """distributed.py:48-58 (one process per GPU, TCP rendezvous)."""
    if dist_backend == "nccl":
        assert torch.cuda.is_available(), "Distributed mode requires CUDA."
        torch.cuda.set_device(rank % torch.cuda.device_count())
    print("Initializing Distributed")
    dist.init_process_group(dist_backend, init_method=dist_url, world_size=num_gpus, rank=rank,
                            group_name=group_name)


def _flatten_dense_tensors(tensors):
    if len(tensors) == 1:
        return tensors[0].contiguous().view(-1)
    return torch.cat([t.contiguous().view(-1) for t in tensors], dim=0)


def _unflatten_dense_tensors(flat, tensors):
    outputs, offset = [], 0
    for tensor in tensors:
        numel = tensor.numel()
        outputs.append(flat.narrow(0, offset, numel).view_as(tensor))
        offset += numel
    return tuple(outputs)


def broadcast_state(module, src=0):
    """Initial weight sync (distributed.py:105-108: 177 broadcasts) as one broadcast
    per dtype of a flat buffer."""
    by_dtype = {}
    for t in module.state_dict().values():
        if torch.is_tensor(t):
            by_dtype.setdefault(t.dtype, []).append(t)
    for ts in by_dtype.values():
        flat = _flatten_dense_tensors([t.data for t in ts])
        dist.broadcast(flat, src)
        for t, s in zip(ts, _unflatten_dense_tensors(flat, ts)):
            t.data.copy_(s)


def _single_flat_buffer(grads):
    """The flat tensor all grads are views of, in order and densely packed, else None."""
    st = grads[0].untyped_storage()
    base = st.data_ptr()
    end = base
    for g in grads:
        if g.untyped_storage().data_ptr() != base or not g.is_contiguous() or g.data_ptr() < end:
            return None
        end = g.data_ptr() + g.numel() * g.element_size()
    n = (end - base) // grads[0].element_size()
    return torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(st, 0, (n,))


def allreduce_gradients(module):
    """One reduction of all gradients of `module` (mean over ranks).  Returns the number
    of collective calls issued (1 on the flat-buffer path)."""
    world = dist.get_world_size()
    buckets = {}
    for p in module.parameters():
        if p.requires_grad and p.grad is not None:
            buckets.setdefault(p.grad.dtype, []).append(p.grad.data)
    calls = 0
    for grads in buckets.values():
        flat = _single_flat_buffer(grads)
        if flat is not None:
            dist.all_reduce(flat)
            flat /= world
        else:                                   # reference path, distributed.py:127-134
            coalesced = _flatten_dense_tensors(grads)
            dist.all_reduce(coalesced)
            coalesced /= world
            for buf, synced in zip(grads, _unflatten_dense_tensors(coalesced, grads)):
                buf.copy_(synced)
        calls += 1
    return calls


def apply_gradient_allreduce(module):
    """distributed.py:95-147: broadcast state from rank 0, then all-reduce the gradients
    after every backward (hook -> engine callback).  Returns the same module object."""
    broadcast_state(module, 0)
    module.needs_reduction = False

    def allreduce_params():
        if module.needs_reduction:
            module.needs_reduction = False
            allreduce_gradients(module)

    def allreduce_hook(*unused):
        Variable._execution_engine.queue_callback(allreduce_params)

    for param in list(module.parameters()):
        if param.requires_grad:
            param.register_hook(allreduce_hook)

    def set_needs_reduction(self, input, output):
        self.needs_reduction = True

    module.register_forward_hook(set_needs_reduction)
    return module
