"""Feature front end / back end: drop-in for the hot-path part of the reference's
dataset.py (ProcessAudio :130-298, pcenfunc :56-76).  forward/backward run the
fused CUDA kernels (csrc/frontend.cu, csrc/backend.cu); the small elementwise
helpers keep the reference's names for API compatibility.  DataAugment
(dataset.py:79-126) and the crop + mix of CleanNoisyPairDataset.__getitem__
(dataset.py:367-379) run batched on the device (csrc/augment.cu);
reading wav files stays with the caller."""
import math
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
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


def _biquad_row(kind, sample_rate, cutoff_freq, Q, chunk):
    """Coefficient row of one biquad for tru_augment_fwd: torchaudio.functional.lowpass_biquad / highpass_biquad design
    (the SoX formulas, evaluated in float32 tensors exactly as torchaudio evaluates them, then divided by a0 as its
    lfilter does) followed by the chunk matrices A^(chunk * {1, 8, 16, 32, 64, 128}), A = [[-a1, -a2], [1, 0]], evaluated in
    float64 (the kernel's carry scan composes chunk end states with them)."""
    f32 = torch.float32
    cutoff_freq = torch.as_tensor(cutoff_freq, dtype=f32)
    Q = torch.as_tensor(Q, dtype=f32)
    w0 = 2 * math.pi * cutoff_freq / sample_rate
    cw = torch.cos(w0)
    alpha = torch.sin(w0) / 2 / Q
    if kind == "lowpass":
        b0 = (1 - cw) / 2
        b1 = 1 - cw
    elif kind == "highpass":
        b0 = (1 + cw) / 2
        b1 = -1 - cw
    else:
        raise ValueError(kind)
    a0 = 1 + alpha
    b = torch.stack([b0, b1, b0]) / a0
    a = torch.stack([-2 * cw, 1 - alpha]) / a0
    A = np.array([[-float(a[0]), -float(a[1])], [1.0, 0.0]], dtype=np.float64)
    powers = [np.linalg.matrix_power(A, chunk * p) for p in (1, 8, 16, 32, 64, 128)]
    return [float(v) for v in b] + [float(v) for v in a] + [float(v) for M in powers for v in M.reshape(-1)]


class DataAugment:
    """dataset.py:79-126.  Same parameter grids and the same ``random.choice`` draws (low-pass, high-pass, gain - in
    that order) as the reference, so a seeded run picks the same augmentation; the filtering itself is one CUDA kernel
    for the whole batch.  ``aug(x)`` takes a CUDA tensor (N,), (1,N) or (B,N) and draws one parameter set per row;
    ``aug(x, params)`` uses the given list of (gain_db, lp_cutoff, hp_cutoff) instead."""

    def __init__(self):
        self.min_gain, self.max_gain = -12.0, -5.0
        self.lp_min, self.lp_max = 7000, 10000
        self.hp_min, self.hp_max = 800, 1200
        self.sr = 48000                                            # dataset.py:110 (fixed, whatever the data rate is)
        self.Q = 0.7
        self.gains = torch.arange(self.min_gain, self.max_gain, 0.033)
        self.lp_freqs = torch.arange(self.lp_min, self.lp_max, 100)
        self.hp_freqs = torch.arange(self.hp_min, self.hp_max, 50)

    def sample_params(self):
        lp_cutoff = random.choice(self.lp_freqs)
        hp_cutoff = random.choice(self.hp_freqs)
        gain = random.choice(self.gains)
        return gain, lp_cutoff, hp_cutoff

    def coefficients(self, params):
        """(B, 59) float32 host tensor of kernel coefficient rows for a list of (gain_db, lp_cutoff, hp_cutoff)."""
        rows = []
        for gain_db, lp_cutoff, hp_cutoff in params:
            if float(gain_db) == 0:                                                      # F.gain
                ratio = 1.0
            elif torch.is_tensor(gain_db):          # the reference draws a 0-d float32 tensor: tensor arithmetic
                ratio = float(10 ** (gain_db / 20))
            else:                                   # a Python number: torchaudio evaluates it in double
                ratio = 10 ** (gain_db / 20)
            rows.append([ratio] + _biquad_row("lowpass", self.sr, lp_cutoff, self.Q, L.AUGMENT_CHUNK)
                        + _biquad_row("highpass", self.sr, hp_cutoff, self.Q, L.AUGMENT_CHUNK))
        return torch.tensor(rows, dtype=torch.float32)

    def __call__(self, x, params=None):
        L.require_cuda(x)
        if x.dtype != torch.float32 or x.dim() not in (1, 2):
            raise L.TruError("DataAugment expects a float32 CUDA tensor (N,), (1,N) or (B,N)")
        rows = x.reshape(-1, x.shape[-1]).contiguous()
        if params is None:
            params = [self.sample_params() for _ in range(rows.shape[0])]
        if len(params) != rows.shape[0]:
            raise L.TruError("DataAugment: %d parameter sets for %d rows" % (len(params), rows.shape[0]))
        coef = self.coefficients(params).to(x.device, non_blocking=True)
        return ops.augment(rows, coef).view(x.shape)


def assemble_batch(clean, noise, crop_length, aug=None, params=None, clean_start=None, noise_start=None):
    """The training pair of dataset.py:352-386 for a whole batch on the device: noise rows augmented over their full
    length (:360), clean rows cropped to ``crop_length`` samples at a random start (:367-373, ``np.random.randint`` like
    the reference), noisy = clean crop + augmented noise (:375).  The reference adds the uncropped noise, which only
    works when the noise file has exactly ``crop_length`` samples; that case is reproduced exactly, and other noise
    lengths are read from ``noise_start`` (default 0) and wrap around.
    clean (B,Nc), noise (B,Nn) float32 CUDA -> (clean (B,crop_length), noisy (B,crop_length))."""
    L.require_cuda(clean, noise)
    B, n_clean = clean.shape
    if noise.shape[0] != B:
        raise L.TruError("assemble_batch: %d clean rows, %d noise rows" % (B, noise.shape[0]))
    if not 0 < crop_length <= n_clean:
        raise L.TruError("assemble_batch: crop_length %d does not fit %d samples" % (crop_length, n_clean))
    aug = aug or DataAugment()
    aug_noise = aug(noise, params)
    if clean_start is None:
        clean_start = [int(np.random.randint(low=0, high=n_clean - crop_length + 1)) for _ in range(B)]
    if noise_start is None:
        noise_start = [0] * B
    if len(clean_start) != B or len(noise_start) != B:
        raise L.TruError("assemble_batch: need one start offset per row")
    if min(clean_start) < 0 or max(clean_start) > n_clean - crop_length or min(noise_start) < 0:
        raise L.TruError("assemble_batch: start offset out of range")
    starts = torch.tensor([list(clean_start), list(noise_start)], dtype=torch.int32).to(clean.device, non_blocking=True)
    return ops.mix_crop(clean.contiguous(), aug_noise, starts[0], starts[1], crop_length)
