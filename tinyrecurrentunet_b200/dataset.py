"""Feature front end / back end: drop-in for the hot-path part of the reference's
dataset.py (ProcessAudio :130-298, pcenfunc :56-76).  forward/backward run the
fused CUDA kernels (csrc/frontend.cu, csrc/backend.cu); the small elementwise
helpers keep the reference's names for API compatibility.  The wav-file dataset
and augmentation classes (dataset.py:79-126,301-412) are out of scope."""
import torch
import torch.nn as nn

from . import ops


def diff(x, axis):
    """dataset.py:24-34 as written (slices only the first two dims, X13)."""
    shape = list(x.shape)
    front = [0] * len(shape)
    front[axis] = 1
    size = list(shape)
    size[axis] -= 1
    a = x[front[0]:front[0] + size[0], front[1]:front[1] + size[1]]
    b = x[0:size[0], 0:size[1]]
    return a - b


def unwrap(p, axis=-1):
    """dataset.py:37-51 (an exact identity on the >=3-D tensors of the hot path, X13)."""
    pi = torch.tensor(torch.pi, dtype=p.dtype, device=p.device)
    dd = diff(p, axis=axis)
    ddmod = torch.remainder(dd + pi, 2.0 * pi) - pi
    idx = torch.logical_and(torch.eq(ddmod, -pi), torch.greater(dd, 0))
    ddmod = torch.where(idx, torch.ones_like(ddmod) * pi, ddmod)
    ph_correct = ddmod - dd
    ph_cumsum = torch.cumsum(ph_correct, axis=axis)
    return (p + ph_cumsum).squeeze(0)


def pcenfunc(x, eps=1e-6, s=0.025, alpha=0.98, delta=2, r=0.5, training=False):
    """Stand-alone PCEN of a magnitude tensor (1,T,F); dataset.py:56-76.  API
    helper only - on the hot path PCEN is fused into the front-end kernel."""
    frames = x.unbind(-2)
    m, ms = None, []
    for fr in frames:
        m = s * fr if m is None else (1 - s) * m + s * fr
        ms.append(m)
    M = torch.stack(ms, dim=-2)
    return (x / (M + eps).pow(alpha) + delta).pow(r) - delta ** r


class ProcessAudio(nn.Module):
    """dataset.py:130-298.  forward: audio -> (T',4,F) features [log-mag, PCEN,
    sin(phase), cos(phase)]; backward: features -> audio."""

    def __init__(self, n_fft=512, hop_length=128, sample_rate=48000, min_level_db=-100):
        super().__init__()
        if n_fft != 512 or hop_length != 128:
            raise NotImplementedError("the CUDA front end is built for n_fft=512, hop_length=128 (tiny.json)")
        self.n_fft = n_fft
        self.n_mels = n_fft // 2 + 1
        self.hop_length = hop_length
        self.sample_rate = sample_rate
        self.sr = sample_rate
        self.min_level_db = -100.
        self.ref_level_db = 25.

    # -- small elementwise helpers (reference names) --
    def get_mag_phase(self, spectrogram):
        return torch.abs(spectrogram).squeeze(0), torch.angle(spectrogram)

    def demod_phase(self, phase):
        d = unwrap(phase)
        return torch.sin(d), torch.cos(d)

    def mod_phase(self, magnitude, real_demod, imag_demod):
        wrap = torch.arctan2(real_demod, imag_demod)
        return (self.db_to_amp(self.de_norm(magnitude)) * torch.exp(1j * wrap)).unsqueeze(0)

    def amp_to_db(self, magnitude):
        return 20 * torch.log10(torch.clamp(magnitude, min=1e-7)) - self.ref_level_db

    def db_to_amp(self, db_spec):
        return torch.pow(10, db_spec / 20.0)

    def perm(self, tensor):
        return tensor.permute(2, 0, 1)

    def de_perm(self, tensor):
        return tensor.permute(1, 2, 0)

    def norm(self, db_spec):
        return torch.clamp((((db_spec - self.min_level_db) / -self.min_level_db) * 2.) - 1., -1, 1)

    def de_norm(self, norm_spec):
        return (((torch.clamp(norm_spec, -1, 1) + 1.) / 2.) * -self.min_level_db) + self.min_level_db \
            + self.ref_level_db

    # -- fused CUDA paths --
    def forward(self, audio):
        """(1,1,N) -> (T',4,257)  [reference call shape];  (B,N) -> (B,T',4,257)."""
        if audio.dim() == 3:
            if audio.shape[0] != 1 or audio.shape[1] != 1:
                raise ValueError("3-D input must be (1,1,N) like the reference (dataset.py:246-257)")
            return ops.frontend(audio.reshape(1, -1))[0]
        if audio.dim() == 2:
            return ops.frontend(audio)
        raise ValueError("audio must be (1,1,N) or (B,N)")

    def backward(self, denoised_features):
        """(T',3|4,257) -> (1,N);  (B,T',3|4,257) -> (B,N)   (dataset.py:275-298)."""
        f = denoised_features
        squeeze = f.dim() == 3
        if squeeze:
            f = f.unsqueeze(0)
        if f.shape[2] == 4:
            f = f[:, :, [0, 2, 3]]
        return ops.features_to_audio(f.contiguous())
