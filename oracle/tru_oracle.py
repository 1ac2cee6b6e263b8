"""CPU oracle for the TRU-Net hot path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

This file is a plain PyTorch/CPU restatement of the reference algorithm
(Okrio/tinyrecurrentunet) for the path named in BASELINE.json.  It is only
ever imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
package ``tinyrecurrentunet_b200`` never imports it.

The reference is an unfinished work in progress: two of its
files do not parse and the model's forward is self-inconsistent, so the oracle
follows the frozen repair decisions D1-D12 of SURVEY.md section 0.2.  Every function
cites the reference lines it restates.

Parity pinning (see tests/test_oracle_vs_reference.py, oracle/make_golden.py):
  * front end ch 0/2/3, mod_phase, iSTFT, pcenfunc: pinned against the
    reference's own dataset.py executed in the build container.
  * multi-resolution STFT loss: pinned against the reference's stft_loss.py.
  * layer classes / parameter counts / state-dict keys: pinned against
    network.py:9-120 (executed with the 4 textual repairs) and docs/net.jpg.
  * gradient all-reduce: pinned against distributed.py (gloo, 2 processes).
  * augmentation (gain + two biquads) and the LR schedule: pinned against the
    reference's dataset.DataAugment / util.LinearWarmupCosineDecay executed
    with seeded ``random`` (tests/golden/augment_ref.npz, lr_schedule_ref.json).
  * model wiring (D4), output channel meaning (D5), mask wiring (D7) and the
    streaming step (D11) have NO runnable reference: **parity unpinned** for
    those; the oracle defines them.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

N_FFT = 512
HOP = 128
NBINS = 257
MIN_LEVEL_DB = -100.0
REF_LEVEL_DB = 25.0


# --------------------------------------------------------------------------
# synthetic inputs and weights
# --------------------------------------------------------------------------
def synthetic_clip(i, n=64000, sr=16000):
    """Deterministic (clean, noisy) pair number ``i``."""
    g = torch.Generator().manual_seed(1000 + i)
    t = torch.arange(n, dtype=torch.float64)
    env = 0.5 * (1.0 + torch.sin(2 * math.pi * 3 * t / n))
    clean = 0.1 * torch.randn(n, generator=g, dtype=torch.float32).double() * env
    for f0 in (220.0, 440.0, 1760.0):
        clean = clean + 0.05 * torch.sin(2 * math.pi * f0 * t / sr)
    noise = 0.03 * torch.randn(n, generator=g, dtype=torch.float32).double()
    clean = clean.float()
    noisy = (clean.double() + noise).float()
    return clean, noisy


def synthetic_batch(b, n=64000, first=0):
    pairs = [synthetic_clip(first + i, n) for i in range(b)]
    clean = torch.stack([p[0] for p in pairs])
    noisy = torch.stack([p[1] for p in pairs])
    return clean, noisy


def randomize_bn(net, seed=0):
    """Non-trivial BN affine/running stats so BN bugs cannot hide (section 8d)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return net


# --------------------------------------------------------------------------
# front end: dataset.py:56-76 (pcenfunc), 130-272 (ProcessAudio.forward)
# --------------------------------------------------------------------------
def pcen(x, eps=1e-6, s=0.025, alpha=0.98, delta=2.0, r=0.5, state=None):
    """PCEN over the time axis (dim -2) of ``x (..., T, F)``; D2.

    Follows dataset.py:56-76: M[0] = s*x[0] (equivalently M[-1] = 0),
    M[t] = (1-s) M[t-1] + s x[t]; out = (x/(M+eps)^alpha + delta)^r - delta^r.
    ``state`` (..., F) carries M between streaming calls (D11); returns
    (out, last M).
    """
    T = x.shape[-2]
    ms = []
    last = state
    for t in range(T):
        frame = x[..., t, :]
        if last is None:
            last = s * frame
        else:
            last = (1 - s) * last + s * frame
        ms.append(last)
    M = torch.stack(ms, dim=-2)
    out = (x / (M + eps).pow(alpha) + delta).pow(r) - delta ** r
    return out, last


def stft_rect(audio):
    """dataset.py:260-264: torch.stft with no window argument (rectangular),
    center=True, reflect padding, onesided. audio (B, N) -> (B, F, T') c64."""
    return torch.stft(audio, n_fft=N_FFT, hop_length=HOP, normalized=False,
                      return_complex=True)


def amp_to_db(mag):
    # dataset.py:207-211
    return 20.0 * torch.log10(torch.clamp(mag, min=1e-7)) - REF_LEVEL_DB


def norm_db(db):
    # dataset.py:229-235
    return torch.clamp(((db - MIN_LEVEL_DB) / -MIN_LEVEL_DB) * 2.0 - 1.0, -1, 1)


def de_norm(x):
    # dataset.py:238-243
    return ((torch.clamp(x, -1, 1) + 1.0) / 2.0) * -MIN_LEVEL_DB + MIN_LEVEL_DB + REF_LEVEL_DB


def db_to_amp(db):
    # dataset.py:214-218
    return torch.pow(10.0, db / 20.0)


def frontend(audio, pcen_state=None, return_state=False):
    """audio (B, N) f32 -> features (B, T', 4, 257) f32 (D1-D3).

    ch0 = norm(amp_to_db(|X|)), ch1 = PCEN(|X|), ch2 = sin(angle X),
    ch3 = cos(angle X).  dataset.py:246-272 produce ch 0/2/3 (unwrap is an
    exact identity on the 3-D tensors of that path, defect X13); ch1 is
    pcenfunc fed with the linear magnitude (D2).
    """
    spec = stft_rect(audio)                        # (B, F, T')
    mag = spec.abs()
    phase = torch.angle(spec)
    ch0 = norm_db(amp_to_db(mag))
    magT = mag.transpose(1, 2)                      # (B, T', F)
    ch1, last = pcen(magT, state=pcen_state)
    ch2 = torch.sin(phase).transpose(1, 2)
    ch3 = torch.cos(phase).transpose(1, 2)
    feats = torch.stack((ch0.transpose(1, 2), ch1, ch2, ch3), dim=2)
    if return_state:
        return feats, last
    return feats


# --------------------------------------------------------------------------
# model: network.py:9-120 (layer classes), 134-150 (layer list), wiring D4
# --------------------------------------------------------------------------
class StandardConv1d(nn.Module):
    # network.py:9-21: Conv1d(pad = stride//2) + ReLU
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.StandardConv1d = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride,
                      padding=stride // 2),
            nn.ReLU(inplace=True))

    def forward(self, x):
        return self.StandardConv1d(x)


class DepthwiseSeparableConv1d(nn.Module):
    # network.py:24-43: pw conv, BN, ReLU, depthwise conv (pad k//2), BN, ReLU
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.DepthwiseSeparableConv1d = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, 1),
            nn.BatchNorm1d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv1d(out_channels, out_channels, kernel_size, stride=stride,
                      padding=kernel_size // 2, groups=out_channels),
            nn.BatchNorm1d(out_channels),
            nn.ReLU(inplace=True))

    def forward(self, x):
        return self.DepthwiseSeparableConv1d(x)


class GRUBlock(nn.Module):
    # network.py:45-58: GRU (batch_first) -> transpose -> pw conv, BN, ReLU
    def __init__(self, in_channels, hidden_size, out_channels, bidirectional):
        super().__init__()
        self.GRU = nn.GRU(in_channels, hidden_size, batch_first=True,
                          bidirectional=bidirectional)
        self.conv = nn.Sequential(
            nn.Conv1d(hidden_size * (2 if bidirectional else 1), out_channels, 1),
            nn.BatchNorm1d(out_channels),
            nn.ReLU(inplace=True))

    def forward(self, x, h0=None, return_state=False):
        y, h = self.GRU(x, h0)
        y = self.conv(y.transpose(1, 2))
        if return_state:
            return y, h
        return y


def _tr_block(cin, cout, k, s, last=False):
    layers = [nn.Conv1d(cin, cout, 1), nn.BatchNorm1d(cout), nn.ReLU(inplace=True),
              nn.ConvTranspose1d(cout, cout, k, stride=s, padding=s // 2)]
    if not last:
        layers += [nn.BatchNorm1d(cout), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


def _skip_cat(x1, x2):
    # network.py:95-98 / 115-118: pad (negative = crop) then concat, x1 first
    d = x2.size(2) - x1.size(2)
    x1 = F.pad(x1, [d // 2, d - d // 2, 0, 0])
    return torch.cat((x1, x2), 1)


class FirstTrCNN(nn.Module):
    # network.py:60-76
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.FirstTrCNN = _tr_block(in_channels, out_channels, kernel_size, stride)

    def forward(self, x):
        return self.FirstTrCNN(x)


class TrCNN(nn.Module):
    # network.py:79-100
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.TrCNN = _tr_block(in_channels, out_channels, kernel_size, stride)

    def forward(self, x1, x2):
        return self.TrCNN(_skip_cat(x1, x2))


class LastTrCNN(nn.Module):
    # network.py:102-120
    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.LastTrCNN = _tr_block(in_channels, out_channels, kernel_size, stride, last=True)

    def forward(self, x1, x2):
        return self.LastTrCNN(_skip_cat(x1, x2))


class TRUNet(nn.Module):
    """network.py:122-171 repaired per D4/D10 (see module docstring).

    The 7 ctor kwargs are accepted and ignored exactly like the reference
    (defect X4); ``in_channels=4`` resolves X3.
    """

    def __init__(self, input_size=None, channels_input=None, channels_output=None,
                 channels_hidden=None, kernel_sizes=None, strides=None,
                 tr_channels_input=None, in_channels=4):
        super().__init__()
        self.encoder = nn.ModuleList([
            StandardConv1d(in_channels, 64, 5, 2),
            DepthwiseSeparableConv1d(64, 128, 3, 1),
            DepthwiseSeparableConv1d(128, 128, 5, 2),
            DepthwiseSeparableConv1d(128, 128, 3, 1),
            DepthwiseSeparableConv1d(128, 128, 5, 2),
            DepthwiseSeparableConv1d(128, 128, 3, 2)])
        self.decoder = nn.ModuleList([
            FirstTrCNN(64, 64, 3, 2),
            TrCNN(192, 64, 5, 2),
            TrCNN(192, 64, 3, 1),
            TrCNN(192, 64, 5, 2),
            TrCNN(192, 64, 3, 1),
            LastTrCNN(128, 8, 5, 2)])
        self.FGRU = GRUBlock(128, 64, 64, bidirectional=True)
        self.TGRU = GRUBlock(64, 128, 64, bidirectional=False)

    def forward(self, x, h0=None, return_state=False):
        """x (T,4,F) or (B,T,4,F) -> (...,8,F).  TGRU runs per (b, f') sequence
        over T with ``h0`` (1, B*16, 128) or zeros (D4, D10, D11)."""
        squeeze = x.dim() == 3
        if squeeze:
            x = x.unsqueeze(0)
        B, T = x.shape[0], x.shape[1]
        x = x.reshape(B * T, x.shape[2], x.shape[3])
        skips = []
        for blk in self.encoder:
            x = blk(x)
            skips.append(x)
        x = self.FGRU(x.transpose(1, 2))                       # (BT, 64, 16)
        L = x.shape[2]
        x = x.view(B, T, 64, L).permute(0, 3, 1, 2).reshape(B * L, T, 64)
        x, h = self.TGRU(x, h0, return_state=True)            # (B*L, 64, T)
        x = x.view(B, L, 64, T).permute(0, 3, 2, 1).reshape(B * T, 64, L)
        x = self.decoder[0](x)
        for i in range(1, 6):
            x = self.decoder[i](x, skips[5 - i])
        x = x.view(B, T, x.shape[1], x.shape[2])
        if squeeze:
            x = x.squeeze(0)
        if return_state:
            return x, h
        return x


# --------------------------------------------------------------------------
# back end: dataset.py:182-203 (mod_phase), phm.py:31-45, dataset.py:293-296
# --------------------------------------------------------------------------
def mod_phase(m, s, c):
    """D6: (norm log-mag, sin ch, cos ch) -> complex spectrum."""
    wrap = torch.arctan2(s, c)
    mag = db_to_amp(de_norm(m))
    return mag * torch.exp(1j * wrap)


def phase_aware_mask(mixture, estimated, beta=0.5):
    """phm.py:31-45 with the two undefined names fixed (X7, D7)."""
    mag_mixture = torch.abs(mixture)
    phase_mixture = torch.angle(mixture)
    phase_estimated = torch.angle(estimated)
    soft_mask = 1 / (1 + torch.exp(-beta * (phase_mixture - phase_estimated)))
    return soft_mask * mag_mixture


def istft_rect(spec):
    """dataset.py:293-296: torch.istft, no window, center=True. (B,F,T')->(B,N)."""
    return torch.istft(spec, n_fft=N_FFT, hop_length=HOP, normalized=False)


def backend(out, beta=0.5):
    """Network output (B,T',8,F) -> denoised audio (B, 128 (T'-1)) (D5-D8)."""
    o = out.transpose(1, 3)                 # (B, F, 8, T')
    spec0 = mod_phase(o[:, :, 0], o[:, :, 2], o[:, :, 3])
    spec1 = mod_phase(o[:, :, 4], o[:, :, 6], o[:, :, 7])
    mag = phase_aware_mask(spec0, spec1, beta)
    den = mag * torch.exp(1j * torch.angle(spec0))
    return istft_rect(den)


def features_to_audio(feats3):
    """ProcessAudio.backward (dataset.py:275-298) for (B,T',3,F) features."""
    o = feats3.transpose(1, 3)
    return istft_rect(mod_phase(o[:, :, 0], o[:, :, 1], o[:, :, 2]))


# --------------------------------------------------------------------------
# loss: stft_loss.py:9-166, util.py:239-250 (D9)
# --------------------------------------------------------------------------
STFT_CFG = dict(fft_sizes=(512, 1024, 2048), hop_sizes=(50, 120, 240),
                win_lengths=(240, 600, 1200), sc_lambda=0.5, mag_lambda=0.5)


def stft_mag(x, n_fft, hop, win):
    # stft_loss.py:9-30
    w = torch.hann_window(win, dtype=x.dtype, device=x.device)
    z = torch.stft(x, n_fft, hop, win, w, return_complex=True)
    p = z.real ** 2 + z.imag ** 2
    return torch.sqrt(torch.clamp(p, min=1e-7)).transpose(2, 1)


def mrstft_loss(x, y, cfg=STFT_CFG):
    """stft_loss.py:141-166, band='full'. x = prediction, y = target, (B,N)."""
    sc = 0.0
    mg = 0.0
    for n_fft, hop, win in zip(cfg["fft_sizes"], cfg["hop_sizes"], cfg["win_lengths"]):
        xm = stft_mag(x, n_fft, hop, win)
        ym = stft_mag(y, n_fft, hop, win)
        sc = sc + torch.norm(ym - xm, p="fro") / torch.norm(ym, p="fro")   # :50
        mg = mg + F.l1_loss(torch.log(ym), torch.log(xm))                   # :69
    n = len(cfg["fft_sizes"])
    return sc * cfg["sc_lambda"] / n, mg * cfg["mag_lambda"] / n


def loss_fn(net, clean, noisy, stft_lambda=1.0, cfg=STFT_CFG, beta=0.5):
    """util.py:186-251 repaired per D5-D9. clean/noisy (B,N)."""
    feats = frontend(noisy)
    out = net(feats)
    den = backend(out, beta)
    n = min(den.shape[-1], clean.shape[-1])
    l1 = torch.abs(F.l1_loss(den[..., :n], clean[..., :n]))
    sc, mg = mrstft_loss(den[..., :n], clean[..., :n], cfg)
    loss = l1 + (sc + mg) * stft_lambda
    return loss, {"l1": l1.detach(), "stft_sc": sc.detach() * stft_lambda,
                  "stft_mag": mg.detach() * stft_lambda}, den


# ---------------------------------------------------------------------------
# Batch assembly
# ---------------------------------------------------------------------------
def augment(noise, gain_db, lp_cutoff, hp_cutoff, sr=48000, q=0.7):
    """dataset.py:121-125: F.gain -> F.lowpass_biquad -> F.highpass_biquad, the calls of torchaudio.functional (the
    reference's third-party dependency for this step; unpinned in requirements.txt, 2.11.0 here) in the reference's
    order.  noise (..., N) CPU float32."""
    import torchaudio.functional as AF
    x = AF.gain(noise, gain_db=gain_db)
    x = AF.lowpass_biquad(x, sr, lp_cutoff, Q=q)
    return AF.highpass_biquad(x, sr, hp_cutoff, Q=q)


def assemble_batch(clean, noise, params, clean_start, noise_start, crop_length):
    """dataset.py:352-386 for a list of rows: augment the whole noise row (:360), crop the clean row (:367-373), add
    (:375).  Noise shorter / longer than the crop is read from noise_start and wraps (with equal lengths and start 0 this
    is the reference's ``clean_audio + noise_audio``)."""
    cs, ns = [], []
    for b in range(clean.shape[0]):
        a = augment(noise[b:b + 1], *params[b])[0]
        c = clean[b, clean_start[b]:clean_start[b] + crop_length]
        idx = (noise_start[b] + torch.arange(crop_length)) % a.shape[0]
        cs.append(c)
        ns.append(c + a[idx])
    return torch.stack(cs), torch.stack(ns)


def cos_sim_loss(x, y, eps=1e-5, g=(508, 1016, 2032, 4062)):
    """cos_loss.py:41-56: mean over the slices [g[i-1], g[i]) of 1 - nn.CosineSimilarity(dim=1, eps).  For one row this is the
    reference's value exactly; rows are averaged and the graph is kept (the reference's torch.FloatTensor(list) can do
    neither)."""
    terms, lo = [], 0
    for hi in g:
        terms.append(1 - F.cosine_similarity(x[:, lo:hi], y[:, lo:hi], dim=1, eps=eps))
        lo = hi
    return torch.stack(terms).mean()
