"""TRU-Net: drop-in for the reference's network.py, running on libtru_b200.so.

The module tree, constructor signatures and state-dict keys are the reference's
(network.py:9-150), so checkpoints are interchangeable and
train.py / distributed.py work unchanged (parameters are ordinary nn.Parameter
leaves; gradients are delivered by autograd so per-parameter hooks fire).
TRUNet.forward does NOT call the sub-modules: the whole network is one C-ABI call
(csrc/trunet.cu) forward and one backward.  The block classes exist for the module
tree / checkpoint schema; called on their own they raise (there is no cuDNN path in
this package: the fused kernels only exist for the whole network).

Wiring follows the repair decisions D4/D10/D11 of SURVEY.md section 0.2 (the reference's
own forward does not run: defects X1-X6).
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L


def _no_standalone(self, *a, **k):
    raise L.TruError("%s: the block classes of tinyrecurrentunet_b200.network carry parameters for TRUNet's fused CUDA "
                     "forward / backward; they have no stand-alone forward (no cuDNN fallback in this package). "
                     "Call TRUNet(...)" % type(self).__name__)


class StandardConv1d(nn.Module):                     # network.py:9-21
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.StandardConv1d = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=stride // 2),
            nn.ReLU(inplace=True))

    forward = _no_standalone


class DepthwiseSeparableConv1d(nn.Module):           # network.py:24-43
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.DepthwiseSeparableConv1d = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, kernel_size=1),
            nn.BatchNorm1d(out_channels), nn.ReLU(inplace=True),
            nn.Conv1d(out_channels, out_channels, kernel_size, stride=stride, padding=kernel_size // 2,
                      groups=out_channels),
            nn.BatchNorm1d(out_channels), nn.ReLU(inplace=True))

    forward = _no_standalone


class GRUBlock(nn.Module):                           # network.py:45-58
    def __init__(self, in_channels, hidden_size, out_channels, bidirectional):
        super().__init__()
        self.GRU = nn.GRU(in_channels, hidden_size, batch_first=True, bidirectional=bidirectional)
        self.conv = nn.Sequential(
            nn.Conv1d(hidden_size * (2 if bidirectional == True else 1), out_channels, kernel_size=1),
            nn.BatchNorm1d(out_channels), nn.ReLU(inplace=True))

    forward = _no_standalone


def _tr(in_channels, out_channels, kernel_size, stride, last):
    mods = [nn.Conv1d(in_channels, out_channels, kernel_size=1), nn.BatchNorm1d(out_channels), nn.ReLU(inplace=True),
            nn.ConvTranspose1d(out_channels, out_channels, kernel_size, stride=stride, padding=stride // 2)]
    if not last:
        mods += [nn.BatchNorm1d(out_channels), nn.ReLU(inplace=True)]
    return nn.Sequential(*mods)


class FirstTrCNN(nn.Module):                         # network.py:60-76
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.FirstTrCNN = _tr(in_channels, out_channels, kernel_size, stride, False)

    forward = _no_standalone


class TrCNN(nn.Module):                              # network.py:79-100
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.TrCNN = _tr(in_channels, out_channels, kernel_size, stride, False)

    forward = _no_standalone                         # (pad / crop + concat of network.py:95-98: row maps in csrc/trunet.cu)


class LastTrCNN(nn.Module):                          # network.py:102-120
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.LastTrCNN = _tr(in_channels, out_channels, kernel_size, stride, True)

    forward = _no_standalone


def _param_order():
    names = ["encoder.0.StandardConv1d.0.weight", "encoder.0.StandardConv1d.0.bias"]
    for i in range(1, 6):
        p = "encoder.%d.DepthwiseSeparableConv1d." % i
        names += [p + s for s in ("0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias", "4.weight", "4.bias")]
    for d in range(6):
        cls = "FirstTrCNN" if d == 0 else ("LastTrCNN" if d == 5 else "TrCNN")
        p = "decoder.%d.%s." % (d, cls)
        names += [p + s for s in ("0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias")]
        if d < 5:
            names += [p + "4.weight", p + "4.bias"]
    g = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    names += ["FGRU.GRU." + s for s in g] + ["FGRU.GRU." + s + "_reverse" for s in g]
    names += ["FGRU.conv.0.weight", "FGRU.conv.0.bias", "FGRU.conv.1.weight", "FGRU.conv.1.bias"]
    names += ["TGRU.GRU." + s for s in g]
    names += ["TGRU.conv.0.weight", "TGRU.conv.0.bias", "TGRU.conv.1.weight", "TGRU.conv.1.bias"]
    return names


def _bn_order():
    names = []
    for i in range(1, 6):
        names += ["encoder.%d.DepthwiseSeparableConv1d.1" % i, "encoder.%d.DepthwiseSeparableConv1d.4" % i]
    for d in range(5):
        cls = "FirstTrCNN" if d == 0 else "TrCNN"
        names += ["decoder.%d.%s.1" % (d, cls), "decoder.%d.%s.4" % (d, cls)]
    names += ["decoder.5.LastTrCNN.1", "FGRU.conv.1", "TGRU.conv.1"]
    return names


PARAM_ORDER = _param_order()        # the order include/tru_b200.h documents for `params`
BN_ORDER = _bn_order()
assert len(PARAM_ORDER) == 108 and len(BN_ORDER) == 23


LOSS_TAIL = 4           # floats behind the last gradient in the flat buffer (distributed.attach_loss)


class _TRUNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h0, net, want_state, need_bwd, *params):
        B, T = x.shape[0], x.shape[1]
        training = net.training
        eps, momentum = net._bn_hyper()
        desc = L.TruNetDesc(B, T, int(training), eps, momentum)
        with torch.cuda.device(x.device):
            ws_bytes = L.lib.tru_trunet_workspace_bytes(C.byref(desc), int(need_bwd))
            ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
            out = torch.empty((B, T, 8, 257), device=x.device, dtype=torch.float32)
            given = torch.is_tensor(want_state)            # step(): the caller's own state buffer (CUDA-graph replay needs fixed addresses)
            if given:
                hl = want_state
                if hl.shape != (B * 16, 128) or hl.dtype != torch.float32 or hl.device != x.device or not hl.is_contiguous():
                    raise L.TruError("h_out must be a contiguous float32 (B*16,128) tensor on the input's device")
                if h0 is not None and hl.data_ptr() == h0.data_ptr():
                    raise L.TruError("h_out must not alias h (the step reads the old state while it writes the new one)")
            else:
                hl = torch.empty((B * 16, 128), device=x.device, dtype=torch.float32) if want_state else None
            pp = (C.c_void_p * 108)(*[p.data_ptr() for p in params])
            rm, rv, nb = net._bn_ptrs(x.device)
            L.check(L.lib.tru_trunet_forward(C.byref(desc), pp, rm, rv, nb, L.ptr(x), L.ptr(h0), L.ptr(out), L.ptr(hl),
                                             L.ptr(ws), ws_bytes, L.stream_ptr(x.device)), "tru_trunet_forward")
        if net._debug_keep_ws:                         # test aid (tests/test_gpu_network.py)
            net._last_ws, net._last_desc = ws, desc
        ctx.desc, ctx.ws_bytes, ctx.need_bwd, ctx.has_h0, ctx.net = desc, ws_bytes, need_bwd, h0 is not None, net
        ctx.shapes = [p.shape for p in params]
        ctx.save_for_backward(x, ws, *params)
        if given:
            return out                                 # the state went into the caller's buffer
        if want_state:
            ctx.mark_non_differentiable(hl)
            return out, hl
        return out

    @staticmethod
    def backward(ctx, gout, *unused):
        if not ctx.need_bwd:
            raise L.TruError("TRUNet backward needs a training-mode forward with grad enabled (eval-mode / frozen-BN "
                             "fine-tuning is not supported: the backward kernels use the batch statistics)")
        if ctx.has_h0:
            raise L.TruError("TRUNet backward with a carried TGRU state (h0) is not supported")
        x, ws = ctx.saved_tensors[:2]
        params = ctx.saved_tensors[2:]
        gout = gout.contiguous()
        sizes = [p.numel() for p in params]
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot)
            tot += (n + 3) // 4 * 4                      # keep every slice 16-byte aligned
        with torch.cuda.device(x.device):
            # one flat buffer for all gradients (+ the all-reduce's scalar tail).  It is a fresh zeroed block every time:
            # the previous step's .grad tensors may still be alive (gradient accumulation), so it cannot be reused in place.
            flat = torch.zeros(tot + LOSS_TAIL, device=x.device, dtype=torch.float32)
            base = flat.data_ptr()
            gp = (C.c_void_p * 108)(*[base + 4 * o for o in offs])
            pp = (C.c_void_p * 108)(*[p.data_ptr() for p in params])
            L.check(L.lib.tru_trunet_backward(C.byref(ctx.desc), pp, L.ptr(x), L.ptr(gout), gp, L.ptr(ws), ctx.ws_bytes,
                                              L.stream_ptr(x.device)), "tru_trunet_backward")
        ctx.net._tru_flat_grad = flat
        grads = [flat[o:o + n].view(s) for o, n, s in zip(offs, sizes, ctx.shapes)]
        return (None, None, None, None, None) + tuple(grads)


class TRUNet(nn.Module):
    """network.py:122-171.  The 7 reference kwargs are accepted and ignored exactly
    like the reference does (X4); ``in_channels`` must be 4 (X3, D1)."""

    def __init__(self, input_size=None, channels_input=None, channels_output=None, channels_hidden=None,
                 kernel_sizes=None, strides=None, tr_channels_input=None, in_channels=4):
        super().__init__()
        if in_channels != 4:
            raise NotImplementedError("the CUDA stem kernel is built for 4 input channels (README.md:50)")
        self.encoder = nn.ModuleList([
            StandardConv1d(in_channels, 64, 5, 2),
            DepthwiseSeparableConv1d(64, 128, 3, 1), DepthwiseSeparableConv1d(128, 128, 5, 2),
            DepthwiseSeparableConv1d(128, 128, 3, 1), DepthwiseSeparableConv1d(128, 128, 5, 2),
            DepthwiseSeparableConv1d(128, 128, 3, 2)])
        self.decoder = nn.ModuleList([
            FirstTrCNN(64, 64, 3, 2), TrCNN(192, 64, 5, 2), TrCNN(192, 64, 3, 1),
            TrCNN(192, 64, 5, 2), TrCNN(192, 64, 3, 1), LastTrCNN(128, 8, 5, 2)])
        self.FGRU = GRUBlock(128, 64, 64, bidirectional=True)
        self.TGRU = GRUBlock(64, 128, 64, bidirectional=False)
        self._debug_keep_ws = False

    # -- plumbing --
    def _ordered_params(self, device):
        ps = self.__dict__.get("_plist")
        if ps is None:                                 # Parameter objects keep their identity through .cuda() / .to() / load_state_dict
            d = dict(self.named_parameters())
            ps = self.__dict__["_plist"] = [d[n] for n in PARAM_ORDER]
        for n, p in zip(PARAM_ORDER, ps):
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise L.TruError("TRUNet parameter %s must be a contiguous float32 tensor on %s (it is %s on %s): the kernels "
                                 "take raw pointers; move the model with .cuda() / .float()" % (n, device, p.dtype, p.device))
        return ps

    def _bn_modules(self):
        bns = self.__dict__.get("_bnlist")
        if bns is None:
            mods = dict(self.named_modules())
            bns = self.__dict__["_bnlist"] = [mods[n] for n in BN_ORDER]
        return bns

    def _bn_hyper(self):
        """(eps, momentum) of the 23 BatchNorm1d layers: the kernels take one value for all of them."""
        bns = self._bn_modules()
        eps, mom = bns[0].eps, bns[0].momentum
        for n, b in zip(BN_ORDER, bns):
            if b.eps != eps or b.momentum != mom or b.momentum is None or not b.track_running_stats or not b.affine:
                raise L.TruError("TRUNet: BatchNorm %s has eps/momentum/track_running_stats/affine settings the fused kernels do "
                                 "not support (one eps and one numeric momentum for all layers, running statistics on)" % n)
        return float(eps), float(mom)

    def _bn_ptrs(self, device):
        bns = self._bn_modules()
        for n, b in zip(BN_ORDER, bns):
            for t, dt in ((b.running_mean, torch.float32), (b.running_var, torch.float32), (b.num_batches_tracked, torch.int64)):
                if t.device != device or t.dtype != dt or not t.is_contiguous():
                    raise L.TruError("TRUNet buffer of %s must be contiguous %s on %s" % (n, dt, device))
        rm = (C.c_void_p * 23)(*[b.running_mean.data_ptr() for b in bns])
        rv = (C.c_void_p * 23)(*[b.running_var.data_ptr() for b in bns])
        nb = (C.c_void_p * 23)(*[b.num_batches_tracked.data_ptr() for b in bns])
        return rm, rv, nb

    def _run(self, x, h0, want_state):
        L.require_cuda(x, h0)
        if x.dtype != torch.float32:
            raise L.TruError("TRUNet expects float32 features")
        squeeze = x.dim() == 3
        if squeeze:
            x = x.unsqueeze(0)
        if x.dim() != 4 or x.shape[2] != 4 or x.shape[3] != 257:
            raise L.TruError("TRUNet expects (T,4,257) or (B,T,4,257), got %s" % (tuple(x.shape),))
        x = x.contiguous()
        if h0 is not None:
            h0 = h0.reshape(-1, 128).contiguous()
            if h0.shape[0] != x.shape[0] * 16:
                raise L.TruError("h0 must hold B*16 states of size 128")
        if h0 is not None and (h0.device != x.device or h0.dtype != torch.float32):
            raise L.TruError("h0 must be float32 on the same device as x")
        params = self._ordered_params(x.device)
        need_bwd = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in params)
        res = _TRUNetFn.apply(x.detach(), h0, self, want_state, need_bwd, *params)
        if torch.is_tensor(want_state):
            out, hl = res, want_state
        else:
            out, hl = (res if want_state else (res, None))
        if squeeze:
            out = out.squeeze(0)
        return (out, hl) if (torch.is_tensor(want_state) or want_state) else out

    def forward(self, x, h0=None, return_state=False):
        """x (T,4,257) or (B,T,4,257) float32 CUDA -> (...,8,257)."""
        return self._run(x, h0, return_state)

    @torch.no_grad()
    def step(self, frame_feats, h, h_out=None):
        """Streaming (D11): frame_feats (S,4,257), h (S*16,128) -> (out (S,8,257), h').  ``h_out``: write h' into this
        (S*16,128) buffer (not ``h`` itself) instead of a new tensor - fixed addresses for CUDA-graph replay."""
        if self.training:
            raise L.TruError("step() is an inference call: put the model in eval() mode")
        out, h2 = self._run(frame_feats.unsqueeze(1), h, True if h_out is None else h_out)
        return out[:, 0], h2
