"""Data-parallel gradient sync: drop-in for the reference's distributed.py:42-147
(init_distributed, apply_gradient_allreduce, reduce_tensor), SURVEY D12.

Same contract as the reference: the module is returned unchanged (no wrapper
class), state is broadcast from rank 0 once, and after every backward the gradients
are averaged over the ranks, triggered by the reference's own mechanism
(per-parameter hook -> engine callback -> one reduction).

What is different is the data path.  TRUNet's backward (network._TRUNetFn) writes
all 108 gradients into ONE flat fp32 buffer (1,525,888 B + a 16-byte tail) and hands
autograd views of it, so the reduction is a single in-place all-reduce of that buffer
on the backward stream: no flatten (torch.cat), no 108 copy-backs, no separate divide
kernel (NCCL: ReduceOp.AVG, exact for power-of-two world sizes), and the logging scalar
of train.py:133 rides in the buffer's tail (``attach_loss``) instead of costing a second
collective and a host sync.  A module whose gradients are not views of one densely
packed buffer raises: this package has no second data path to fall back to."""
import torch
import torch.distributed as dist
from torch.autograd import Variable

LOSS_TAIL = 4           # floats appended to TRUNet's flat gradient buffer (network._TRUNetFn.backward): [loss, 0, 0, 0]


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


def broadcast_state(module, src=0):
    """Initial weight sync (distributed.py:105-108: 177 broadcasts) as one broadcast
    per dtype of a flat buffer."""
    by_dtype = {}
    for t in module.state_dict().values():
        if torch.is_tensor(t):
            by_dtype.setdefault(t.dtype, []).append(t)
    for ts in by_dtype.values():
        flat = torch.cat([t.data.contiguous().view(-1) for t in ts], dim=0)
        dist.broadcast(flat, src)
        offset = 0
        for t in ts:
            t.data.copy_(flat.narrow(0, offset, t.numel()).view_as(t))
            offset += t.numel()


def _single_flat_buffer(grads):
    """The flat tensor all ``grads`` are views of - same storage, ascending, contiguous, at most 3 elements (the 16-byte
    alignment padding) between neighbours - from the first gradient's offset to the end of the last one.  None if the
    gradients are not laid out like that."""
    st = grads[0].untyped_storage()
    base, es = st.data_ptr(), grads[0].element_size()
    start = grads[0].data_ptr()
    end = start
    for g in grads:
        if g.untyped_storage().data_ptr() != base or not g.is_contiguous() or g.data_ptr() < end or g.data_ptr() - end > 3 * es:
            return None
        end = g.data_ptr() + g.numel() * es
    return torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(st, (start - base) // es, ((end - start) // es,))


def attach_loss(module, loss):
    """Piggy-back the logging scalar of train.py:133 on the gradient bucket: call before ``loss.backward()``; after the
    backward ``module.reduced_loss`` is a 0-d device tensor holding the mean of ``loss`` over the ranks (no extra
    collective, no host sync until the caller reads it)."""
    module._tru_loss = loss.detach()


def allreduce_gradients(module):
    """One reduction of all gradients of ``module`` (mean over ranks).  Returns the number of collective calls (1)."""
    world = dist.get_world_size()
    grads = [p.grad.data for p in module.parameters() if p.requires_grad and p.grad is not None]
    if not grads:
        return 0
    flat = _single_flat_buffer(grads)
    if flat is None:
        raise RuntimeError("allreduce_gradients: the gradients are not views of one flat buffer (TRUNet's backward produces "
                           "them that way; gradient accumulation over several backwards is not supported)")
    # TRUNet's backward leaves its whole buffer (gradients + LOSS_TAIL floats) on the module: reduce the tail with it
    own = getattr(module, "_tru_flat_grad", None)
    loss = getattr(module, "_tru_loss", None)
    tail = None
    if own is not None and own.data_ptr() == flat.data_ptr() and 0 <= own.numel() - flat.numel() - LOSS_TAIL < 4:   # (< 4: padding of the last slice)
        flat, tail = own, own[-LOSS_TAIL:]
        if loss is not None:
            tail[0].copy_(loss)
    avg = dist.get_backend() == "nccl" and world & (world - 1) == 0       # 1/world is exact: same bits as SUM then divide
    dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM)
    if not avg:
        flat /= world
    if tail is not None and loss is not None:
        module.reduced_loss = tail[0]
    module._tru_loss = None
    return 1


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
