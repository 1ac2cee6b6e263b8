"""Optimizer step of the training iteration on flat buffers.

train.py:68 builds ``torch.optim.AdamW(net.parameters(), lr=...)``; train.py:138-140 then runs, every iteration,
``clip_grad_norm_(net.parameters(), 1e9)`` (= report the norm), ``scheduler.step()`` and ``optimizer.step()``.
With 108 small tensors that is three multi-tensor sweeps.  ``FlatAdamW`` is a drop-in ``torch.optim.AdamW``
(same constructor arguments, same ``state`` / ``param_groups`` / ``state_dict()`` layout, so the reference's
checkpoints load and vice versa) whose ``step()`` is ONE C call (tru_flat_adamw_step: two launches, no host
synchronisation): parameters, both moments and -- as delivered by TRUNet's backward -- the gradients live in one
flat fp32 buffer each, all in the same layout.

There is no CPU path: a step on non-CUDA parameters raises."""
import ctypes as C

import torch

from . import _lib as L


def _layout(params):
    """Offsets (in floats) of every parameter in the flat buffers: the layout network._TRUNetFn.backward uses for the
    gradients (every slice 16-byte aligned)."""
    offs, tot = [], 0
    for p in params:
        offs.append(tot)
        tot += (p.numel() + 3) // 4 * 4
    return offs, tot


class FlatAdamW(torch.optim.AdamW):
    """``torch.optim.AdamW`` (train.py:68) with a fused flat step.  ``max_grad_norm`` folds train.py:138 in: the L2
    norm of all gradients is left in ``self.grad_norm`` (a 0-d CUDA tensor, no sync) and, when it exceeds
    ``max_grad_norm``, the gradients are scaled by ``max_grad_norm / (norm + 1e-6)`` before the update exactly like
    ``clip_grad_norm_``.  ``None`` reports the norm without clipping.

    Construction re-points every ``p.data`` at a slice of one flat buffer (values preserved), so build the optimizer
    after ``net.cuda()`` as train.py does.  One parameter group."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False, max_grad_norm=None, **unused):
        if amsgrad or maximize:
            raise L.TruError("FlatAdamW implements the default AdamW (no amsgrad, no maximize)")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        if len(self.param_groups) != 1:
            raise L.TruError("FlatAdamW takes one parameter group (train.py:68 passes net.parameters())")
        self.max_grad_norm = max_grad_norm
        self.grad_norm = None
        self._flat = None
        self._state_moved = True

    # -- flat buffers ------------------------------------------------------------------------------------------
    def add_param_group(self, param_group):
        if self.param_groups:
            raise L.TruError("FlatAdamW takes one parameter group")
        super().add_param_group(param_group)

    def state_dict(self):
        """torch.optim.AdamW's layout.  The live states share one step counter; the dict gets a private copy per
        parameter, because torch's load_state_dict keeps "step" tensors by reference and its multi-tensor step
        increments each of them."""
        sd = super().state_dict()
        sd["state"] = {k: {kk: (vv.clone() if kk == "step" else vv) for kk, vv in st.items()} for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._state_moved = True               # the loaded moments are fresh tensors: fold them back into the flat buffers

    def _build(self, params):
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise L.TruError("FlatAdamW needs float32 CUDA parameters (there is no CPU path); got %s on %s"
                                 % (p.dtype, p.device))
        dev = params[0].device
        offs, tot = _layout(params)
        old = self._flat
        if old is not None and old["n"] == tot and old["p"].device == dev:
            fp, fm, fv = old["p"], old["m"], old["v"]
        else:
            fp, fm, fv = (torch.zeros(tot, device=dev, dtype=torch.float32) for _ in range(3))
        steps = set()
        for p, o in zip(params, offs):
            n = p.numel()
            view = fp[o:o + n].view(p.shape)
            if p.data.data_ptr() != view.data_ptr():
                view.copy_(p.data)
                p.data = view
            st = self.state[p]
            for key, flat in (("exp_avg", fm), ("exp_avg_sq", fv)):
                mv = flat[o:o + n].view(p.shape)
                cur = st.get(key)
                if cur is None:
                    mv.zero_()
                elif cur.data_ptr() != mv.data_ptr():
                    mv.copy_(cur)
                st[key] = mv
            steps.add(int(float(st["step"])) if "step" in st else 0)
        if len(steps) != 1:
            raise L.TruError("FlatAdamW: parameters disagree on the step count %s" % sorted(steps))
        step_t = torch.tensor(float(steps.pop()), dtype=torch.float32)      # one CPU scalar shared by all 108 states
        for p in params:
            self.state[p]["step"] = step_t
        desc = L.TruAdamWDesc(tot, 1, 0, 0, 0, 0, 0, 0)
        ws_bytes = L.lib.tru_flat_adamw_workspace_bytes(C.byref(desc))
        self._flat = dict(n=tot, offs=offs, p=fp, m=fm, v=fv, step=step_t, gather=None,
                          ws=torch.empty(ws_bytes, device=dev, dtype=torch.uint8), ws_bytes=ws_bytes,
                          norm=torch.zeros((), device=dev, dtype=torch.float32))
        self._state_moved = False
        return self._flat

    def _ensure_flat(self, params):
        fl = self._flat
        if fl is None or self._state_moved:
            return self._build(params)
        base = fl["p"].data_ptr()
        for p, o in zip(params, fl["offs"]):
            if p.data_ptr() != base + 4 * o:                 # someone moved a parameter (net.to(...), p.data = ...)
                return self._build(params)
        return fl

    def _grad_base(self, fl, params):
        """Device address of the flat gradient buffer; gathers the gradients into one if autograd did not deliver
        them as views of a single buffer in this layout."""
        grads = []
        for p in params:
            if p.grad is None:
                raise L.TruError("FlatAdamW.step: a parameter has no gradient (every parameter of the group must take part)")
            grads.append(p.grad)
        g0 = grads[0]
        base = g0.data_ptr()
        ok = base % 16 == 0
        if ok:
            for g, o in zip(grads, fl["offs"]):
                if g.data_ptr() != base + 4 * o or g.dtype != torch.float32 or not g.is_contiguous():
                    ok = False
                    break
        if ok:
            st = g0.untyped_storage()
            ok = st.data_ptr() + st.nbytes() >= base + 4 * fl["n"]
        if ok:
            return base
        if fl["gather"] is None:
            fl["gather"] = torch.zeros(fl["n"], device=fl["p"].device, dtype=torch.float32)
        views = [fl["gather"][o:o + g.numel()].view(g.shape) for g, o in zip(grads, fl["offs"])]
        torch._foreach_copy_(views, grads)
        return fl["gather"].data_ptr()

    # -- the step ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        params = group["params"]
        if group.get("amsgrad") or group.get("maximize"):
            raise L.TruError("FlatAdamW implements the default AdamW (no amsgrad, no maximize)")
        fl = self._ensure_flat(params)
        gbase = self._grad_base(fl, params)
        t = int(float(fl["step"])) + 1
        beta1, beta2 = group["betas"]
        desc = L.TruAdamWDesc(fl["n"], t, float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                              float(group["weight_decay"]), float(self.max_grad_norm or 0.0))
        L.run("tru_flat_adamw_step", fl["p"].device, C.byref(desc), L.ptr(fl["p"]), gbase, L.ptr(fl["m"]), L.ptr(fl["v"]),
              L.ptr(fl["norm"]), L.ptr(fl["ws"]), fl["ws_bytes"])
        fl["step"] += 1
        self.grad_norm = fl["norm"]
        return loss


@torch.no_grad()
def grad_norm(parameters):
    """L2 norm of all gradients (what train.py:138's clip_grad_norm_(..., 1e9) returns) as a 0-d CUDA tensor."""
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        raise L.TruError("grad_norm: no gradients")
    L.require_cuda(*grads)
    offs, tot = _layout(grads)
    flat = torch.zeros(tot, device=grads[0].device, dtype=torch.float32)
    torch._foreach_copy_([flat[o:o + g.numel()].view(g.shape) for g, o in zip(grads, offs)], grads)
    desc = L.TruAdamWDesc(tot, 1, 0, 0, 0, 0, 0, 0)
    ws_bytes = L.lib.tru_flat_adamw_workspace_bytes(C.byref(desc))
    ws = torch.empty(ws_bytes, device=flat.device, dtype=torch.uint8)
    out = torch.zeros((), device=flat.device, dtype=torch.float32)
    L.run("tru_flat_grad_norm", flat.device, tot, L.ptr(flat), L.ptr(out), L.ptr(ws), ws_bytes)
    return out
