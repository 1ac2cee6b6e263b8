"""Cosine-similarity loss: drop-in for the reference's cos_loss.py (CosSimLoss :4-56; imported by util.py:15, never called).

Same constructor and segment rule - slice i is ``[g[i-1], g[i])`` (the first starts at 0), the loss is the mean over the
slices of ``1 - cosine_similarity`` - on two CUDA kernels (csrc/cossim.cu).  The reference assembles the result with
``torch.FloatTensor(list)``, which only works for a single row and detaches it; here batch rows are averaged and the loss is
differentiable with respect to the prediction ``x``."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L


class _CosSimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, desc):
        stats = torch.empty((x.shape[0], desc.n_seg, 3), device=x.device, dtype=torch.float64)
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        L.run("tru_cossim_fwd", x.device, C.byref(desc), L.ptr(x), L.ptr(y), L.ptr(stats), L.ptr(loss))
        ctx.desc = desc
        ctx.save_for_backward(x, y, stats)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        x, y, stats = ctx.saved_tensors
        gx = torch.empty_like(x)
        gl = grad_loss.contiguous().float()
        L.run("tru_cossim_bwd", x.device, C.byref(ctx.desc), L.ptr(x), L.ptr(y), L.ptr(stats), L.ptr(gl), L.ptr(gx))
        return gx, None, None


class CosSimLoss(nn.Module):
    def __init__(self, eps=1e-5, g=[508, 1016, 2032, 4062]):
        super().__init__()
        self.eps = eps
        self.g = g
        self.m = len(self.g)
        if not 0 < self.m <= 8:
            raise ValueError("CosSimLoss supports 1..8 segments")

    def forward(self, x, y):
        """x (prediction), y (target): (B,N) float32 CUDA, N >= g[-1] -> 0-d loss."""
        L.require_cuda(x, y)
        if x.dim() != 2 or x.shape != y.shape or x.dtype != torch.float32 or y.dtype != torch.float32:
            raise L.TruError("CosSimLoss expects two float32 (B,N) tensors of the same shape")
        bounds = [0] + [int(v) for v in self.g]
        desc = L.TruCosSimDesc(x.shape[0], x.shape[1], self.m, (C.c_int * 9)(*(bounds + [0] * (9 - len(bounds)))), float(self.eps))
        return _CosSimFn.apply(x.contiguous(), y.contiguous().detach(), desc)
