"""torch.autograd bindings of the C-ABI kernels (thin: shape checks, buffer
allocation through torch's caching allocator, one C call per op)."""
import ctypes as C

import torch

from . import _lib as L

NFFT, HOP, NBINS = 512, 128, 257
PCEN_DEFAULTS = dict(eps=1e-6, s=0.025, alpha=0.98, delta=2.0, r=0.5)   # dataset.py:57


def _f32c(t):
    if t.dtype != torch.float32:
        raise L.TruError("expected float32, got %s" % t.dtype)
    return t.contiguous()


# ------------------------------------------------------------------ front end
def _front_desc(batch, n, pcen):
    kw = dict(PCEN_DEFAULTS)
    kw.update(pcen or {})
    return L.TruFrontendDesc(batch, n, kw["eps"], kw["s"], kw["alpha"], kw["delta"], kw["r"])


def frontend(audio, pcen_state=None, return_state=False, pcen=None):
    """audio (B,N) -> feats (B,T',4,257); dataset.py:246-272 + pcenfunc :56-76."""
    L.require_cuda(audio, pcen_state)
    audio = _f32c(audio)
    B, N = audio.shape
    T = 1 + N // HOP
    d = _front_desc(B, N, pcen)
    feats = torch.empty((B, T, 4, NBINS), device=audio.device, dtype=torch.float32)
    ws_bytes = L.lib.tru_frontend_workspace_bytes(C.byref(d))
    ws = torch.empty(ws_bytes, device=audio.device, dtype=torch.uint8)
    st_in = _f32c(pcen_state) if pcen_state is not None else None
    st_out = torch.empty((B, NBINS), device=audio.device, dtype=torch.float32) if return_state else None
    L.run("tru_frontend_fwd", audio.device, C.byref(d), L.ptr(audio), L.ptr(st_in), L.ptr(feats), L.ptr(st_out),
                                   L.ptr(ws), ws_bytes)
    return (feats, st_out) if return_state else feats


def frontend_step(frames, pcen_state, pcen=None):
    """Streaming: frames (S,512), state (S,257) updated in place -> feats (S,4,257)."""
    L.require_cuda(frames, pcen_state)
    frames = _f32c(frames)
    if not pcen_state.is_contiguous() or pcen_state.dtype != torch.float32:
        raise L.TruError("pcen_state must be a contiguous float32 tensor (updated in place)")
    S = frames.shape[0]
    d = _front_desc(S, NFFT, pcen)
    feats = torch.empty((S, 4, NBINS), device=frames.device, dtype=torch.float32)
    L.run("tru_frontend_step", frames.device, C.byref(d), L.ptr(frames), L.ptr(pcen_state), L.ptr(feats))
    return feats


# ------------------------------------------------------------------- back end
class _MaskISTFT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net_out, chans, use_mask, beta):
        L.require_cuda(net_out)
        x = _f32c(net_out)
        B, T, Cn, F = x.shape
        if F != NBINS:
            raise L.TruError("last dim must be 257 bins")
        d = L.TruBackendDesc(B, T, Cn, chans[0], chans[1], chans[2], chans[3], chans[4], int(use_mask), beta)
        audio = torch.empty((B, HOP * (T - 1)), device=x.device, dtype=torch.float32)
        L.run("tru_backend_fwd", x.device, C.byref(d), L.ptr(x), L.ptr(audio))
        ctx.save_for_backward(x)
        ctx.desc = d
        return audio

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _f32c(g)
        gx = torch.empty_like(x)
        L.run("tru_backend_bwd", x.device, C.byref(ctx.desc), L.ptr(x), L.ptr(g), L.ptr(gx))
        return gx, None, None, None


def mask_istft(net_out, beta=0.5):
    """(B,T',8,257) -> (B,128(T'-1)): mod_phase x2 + beta-sigmoid mask + iSTFT (D5-D8)."""
    return _MaskISTFT.apply(net_out, (0, 2, 3, 6, 7), True, float(beta))


def mask_istft_step(net_out_frame, ola_state, frame_index, beta=0.5, flush=False):
    """Streaming back end (D11): net_out_frame (S,8,257) (None when flushing), ola_state (S,384) updated in place
    -> (S,128) = output block ``frame_index - 2`` of every stream (zeros while frame_index < 2)."""
    L.require_cuda(net_out_frame, ola_state)
    if not ola_state.is_contiguous() or ola_state.dtype != torch.float32 or ola_state.shape[-1] != 384:
        raise L.TruError("ola_state must be a contiguous float32 (S,384) tensor (updated in place)")
    S = ola_state.shape[0]
    x = None
    if not flush:
        x = _f32c(net_out_frame)
        if x.shape != (S, 8, NBINS):
            raise L.TruError("net_out_frame must be (S,8,257) with S = ola_state.shape[0]")
    d = L.TruBackendDesc(S, 1, 8, 0, 2, 3, 6, 7, 1, float(beta))
    audio = torch.empty((S, HOP), device=ola_state.device, dtype=torch.float32)
    L.run("tru_backend_step", ola_state.device, C.byref(d), L.ptr(x), L.ptr(ola_state), L.ptr(audio), int(frame_index),
                                   0 if flush else 1)
    return audio


def features_to_audio(feats3):
    """(B,T',3,257) [logmag, sin, cos] -> audio: ProcessAudio.backward, dataset.py:275-298."""
    return _MaskISTFT.apply(feats3, (0, 1, 2, 0, 0), False, 0.0)


# ----------------------------------------------------------------------- loss
class _MRSTFTL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, windows, cfg):
        L.require_cuda(x, y)
        x = _f32c(x)
        y = _f32c(y)
        if x.shape != y.shape or x.dim() != 2:
            raise L.TruError("loss expects x, y of identical shape (B,N)")
        B, N = x.shape
        nres = len(cfg["fft_sizes"])
        d = L.TruLossDesc(B, N, nres, (C.c_int * 3)(*(list(cfg["fft_sizes"]) + [0] * (3 - nres))),
                          (C.c_int * 3)(*(list(cfg["hop_sizes"]) + [0] * (3 - nres))),
                          (C.c_int * 3)(*(list(cfg["win_lengths"]) + [0] * (3 - nres))),
                          cfg["sc_lambda"], cfg["mag_lambda"])
        wins = [_f32c(w) for w in windows]
        L.require_cuda(*wins)
        wptr = (C.c_void_p * 3)(*([w.data_ptr() for w in wins] + [None] * (3 - nres)))
        sums = torch.empty(16, device=x.device, dtype=torch.float64)
        out = torch.empty(3, device=x.device, dtype=torch.float32)
        L.run("tru_loss_fwd", x.device, C.byref(d), L.ptr(x), L.ptr(y), wptr, L.ptr(sums), L.ptr(out))
        ctx.save_for_backward(x, y, sums, *wins)
        ctx.desc = d
        ctx.mark_non_differentiable(sums)
        return out[0], out[1], out[2], sums

    @staticmethod
    def backward(ctx, g_l1, g_sc, g_mag, _g_sums):
        x, y, sums = ctx.saved_tensors[:3]
        wins = ctx.saved_tensors[3:]
        d = ctx.desc
        zero = torch.zeros((), device=x.device, dtype=torch.float32)
        gout = torch.stack([g if g is not None else zero for g in (g_l1, g_sc, g_mag)]).float().contiguous()
        wptr = (C.c_void_p * 3)(*([w.data_ptr() for w in wins] + [None] * (3 - len(wins))))
        gx = torch.empty_like(x)
        L.run("tru_loss_bwd", x.device, C.byref(d), L.ptr(x), L.ptr(y), wptr, L.ptr(sums), L.ptr(gout), L.ptr(gx))
        return gx, None, None, None


def mrstft_l1(x, y, windows, cfg):
    """Fused (l1, sc_loss, mag_loss) of prediction x vs target y, both (B,N)."""
    l1, sc, mag, _ = _MRSTFTL1.apply(x, y, tuple(windows), cfg)
    return l1, sc, mag


def augment(noise, coef):
    """dataset.py:79-126 for B rows: noise (B,N) float32 CUDA, coef (B,59) (dataset.DataAugment.coefficients) -> (B,N)."""
    L.require_cuda(noise, coef)
    noise, coef = noise.contiguous(), coef.contiguous()
    if noise.dim() != 2 or coef.shape != (noise.shape[0], L.AUGMENT_NCOEF) or coef.dtype != torch.float32:
        raise L.TruError("augment: noise (B,N) and coef (B,%d) float32 expected" % L.AUGMENT_NCOEF)
    out = torch.empty_like(noise)
    L.run("tru_augment_fwd", noise.device, noise.shape[0], noise.shape[1], L.ptr(noise), L.ptr(coef), L.ptr(out))
    return out


def mix_crop(clean, aug_noise, clean_start, noise_start, n_out):
    """dataset.py:367-379: clean (B,Nc), aug_noise (B,Nn), int32 CUDA start offsets (B) -> (clean crop, noisy) (B,n_out)."""
    L.require_cuda(clean, aug_noise, clean_start, noise_start)
    if clean_start.dtype != torch.int32 or noise_start.dtype != torch.int32:
        raise L.TruError("mix_crop: start offsets must be int32")
    B = clean.shape[0]
    clean_out = torch.empty((B, n_out), device=clean.device, dtype=torch.float32)
    noisy_out = torch.empty_like(clean_out)
    L.run("tru_mix_crop", clean.device, B, clean.shape[1], aug_noise.shape[1], n_out, L.ptr(clean), L.ptr(aug_noise),
                               L.ptr(clean_start.contiguous()), L.ptr(noise_start.contiguous()), L.ptr(clean_out),
                               L.ptr(noisy_out))
    return clean_out, noisy_out
