"""ctypes binding of libtru_b200.so (include/tru_b200.h).

There is deliberately no fallback: if the shared library is missing the import
fails loudly, and every call that returns a non-zero status raises."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libtru_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "tinyrecurrentunet_b200: %s not found. Build it with "
        "`python tinyrecurrentunet_b200/build.py` (needs nvcc, targets sm_100a). "
        "There is no CPU/PyTorch fallback for this path." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

c_float_p = C.c_void_p     # raw device pointers are passed as integers
c_stream = C.c_void_p


class TruFrontendDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("n_samples", C.c_int),
                ("pcen_eps", C.c_double), ("pcen_s", C.c_double), ("pcen_alpha", C.c_double),
                ("pcen_delta", C.c_double), ("pcen_r", C.c_double)]


class TruBackendDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("n_frames", C.c_int), ("n_channels", C.c_int),
                ("ch_mag", C.c_int), ("ch_sin", C.c_int), ("ch_cos", C.c_int),
                ("ch_sin1", C.c_int), ("ch_cos1", C.c_int), ("use_mask", C.c_int),
                ("beta", C.c_double)]


class TruLossDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("n_samples", C.c_int), ("n_res", C.c_int),
                ("fft_size", C.c_int * 3), ("hop_size", C.c_int * 3), ("win_length", C.c_int * 3),
                ("sc_lambda", C.c_double), ("mag_lambda", C.c_double)]


class TruAdamWDesc(C.Structure):
    _fields_ = [("n", C.c_longlong), ("step", C.c_longlong), ("lr", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double), ("max_grad_norm", C.c_double)]


class TruCosSimDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("n_samples", C.c_int), ("n_seg", C.c_int), ("bounds", C.c_int * 9), ("eps", C.c_double)]


class TruNetDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("n_frames", C.c_int), ("training", C.c_int),
                ("bn_eps", C.c_double), ("bn_momentum", C.c_double)]


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


_sig("tru_abi_version", C.c_int, [])
_sig("tru_last_error", C.c_char_p, [])
_sig("tru_init", C.c_int, [])
_sig("tru_frontend_workspace_bytes", C.c_size_t, [C.POINTER(TruFrontendDesc)])
_sig("tru_frontend_fwd", C.c_int, [C.POINTER(TruFrontendDesc), c_float_p, c_float_p, c_float_p, c_float_p,
                                  C.c_void_p, C.c_size_t, c_stream])
_sig("tru_frontend_step", C.c_int, [C.POINTER(TruFrontendDesc), c_float_p, c_float_p, c_float_p, c_stream])
_sig("tru_backend_fwd", C.c_int, [C.POINTER(TruBackendDesc), c_float_p, c_float_p, c_stream])
_sig("tru_backend_step", C.c_int, [C.POINTER(TruBackendDesc), c_float_p, c_float_p, c_float_p, C.c_int, C.c_int, c_stream])
_sig("tru_backend_bwd", C.c_int, [C.POINTER(TruBackendDesc), c_float_p, c_float_p, c_float_p, c_stream])
_sig("tru_loss_fwd", C.c_int, [C.POINTER(TruLossDesc), c_float_p, c_float_p, C.POINTER(C.c_void_p),
                              C.c_void_p, c_float_p, c_stream])
_sig("tru_loss_bwd", C.c_int, [C.POINTER(TruLossDesc), c_float_p, c_float_p, C.POINTER(C.c_void_p),
                              C.c_void_p, c_float_p, c_float_p, c_stream])

_PP = C.POINTER(C.c_void_p)
_sig("tru_trunet_workspace_bytes", C.c_size_t, [C.POINTER(TruNetDesc), C.c_int])
_sig("tru_trunet_forward", C.c_int, [C.POINTER(TruNetDesc), _PP, _PP, _PP, _PP, c_float_p, c_float_p, c_float_p,
                                    c_float_p, C.c_void_p, C.c_size_t, c_stream])
_sig("tru_trunet_backward", C.c_int, [C.POINTER(TruNetDesc), _PP, c_float_p, c_float_p, _PP, C.c_void_p,
                                     C.c_size_t, c_stream])
_sig("tru_trunet_buffer_offset", C.c_longlong, [C.POINTER(TruNetDesc), C.c_char_p, C.c_int])

_sig("tru_launch_count", C.c_longlong, [])
_sig("tru_profile_enable", C.c_int, [C.c_int])
_sig("tru_profile_report", C.c_int, [C.c_char_p, C.c_size_t])
_sig("tru_flat_adamw_workspace_bytes", C.c_size_t, [C.POINTER(TruAdamWDesc)])
_sig("tru_flat_adamw_step", C.c_int, [C.POINTER(TruAdamWDesc), c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                     C.c_void_p, C.c_size_t, c_stream])
_sig("tru_augment_fwd", C.c_int, [C.c_int, C.c_int, c_float_p, c_float_p, c_float_p, c_stream])
_sig("tru_mix_crop", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p, C.c_void_p, C.c_void_p, c_float_p,
                              c_float_p, c_stream])
_sig("tru_cossim_fwd", C.c_int, [C.POINTER(TruCosSimDesc), c_float_p, c_float_p, C.c_void_p, c_float_p, c_stream])
_sig("tru_cossim_bwd", C.c_int, [C.POINTER(TruCosSimDesc), c_float_p, c_float_p, C.c_void_p, c_float_p, c_float_p, c_stream])
_sig("tru_flat_grad_norm", C.c_int, [C.c_longlong, c_float_p, c_float_p, C.c_void_p, C.c_size_t, c_stream])

EXPORTS = ["tru_launch_count", "tru_profile_enable", "tru_profile_report", "tru_abi_version", "tru_last_error", "tru_init", "tru_frontend_workspace_bytes",
           "tru_frontend_fwd", "tru_frontend_step", "tru_backend_fwd", "tru_backend_bwd", "tru_backend_step",
           "tru_loss_fwd", "tru_loss_bwd", "tru_trunet_workspace_bytes", "tru_trunet_forward",
           "tru_trunet_backward", "tru_trunet_buffer_offset", "tru_flat_adamw_workspace_bytes", "tru_flat_adamw_step",
           "tru_flat_grad_norm", "tru_augment_fwd", "tru_mix_crop", "tru_cossim_fwd",
           "tru_cossim_bwd"]
AUGMENT_CHUNK = 63        # TRU_AUGMENT_CHUNK
AUGMENT_NCOEF = 59        # TRU_AUGMENT_NCOEF


class TruError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib.tru_last_error()
        raise TruError("%s failed (%d): %s" % (what or "libtru_b200", rc, msg.decode() if msg else "?"))


def profile_enable(on=True):
    check(lib.tru_profile_enable(int(bool(on))), "tru_profile_enable")


def profile_report():
    """{kernel: dict(launches, ms, bytes, flops)} since the last report (synchronises)."""
    buf = C.create_string_buffer(1 << 16)
    check(lib.tru_profile_report(buf, len(buf)), "tru_profile_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms, b, f = line.split()
        out[name] = dict(launches=int(n), ms=float(ms), bytes=float(b), flops=float(f))
    return out


def ptr(t):
    """Raw device pointer of a contiguous fp32/fp64 CUDA tensor (or None)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    """The current torch stream of ``device`` (default: the current device).  Callers wrap the C call in
    ``torch.cuda.device(device)`` so that the library's own cudaGetDevice-based state matches."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def run(name, device, *args):
    """One C-ABI call on ``device``: that device is made current for the call (the library keeps per-device state keyed
    by cudaGetDevice), the work goes onto torch's current stream of that device, a non-zero status raises."""
    import torch
    with torch.cuda.device(device):
        check(getattr(lib, name)(*args, torch.cuda.current_stream(device).cuda_stream), name)


def same_device(*tensors):
    """All given CUDA tensors live on one device: returns it."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise TruError("tensors on different devices: %s vs %s" % (dev, t.device))
    return dev


def require_cuda(*tensors):
    """Every given tensor is a CUDA tensor and all of them are on one device."""
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TruError("tinyrecurrentunet_b200 ops need CUDA tensors (sm_100a); there is no CPU path")
    same_device(*tensors)
