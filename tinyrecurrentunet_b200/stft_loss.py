"""Multi-resolution STFT loss: drop-in for the reference's stft_loss.py, backed
by the fused CUDA kernels in csrc/loss.cu (forward and backward, magnitudes never
materialised)."""
import torch
import torch.nn.functional as F

from . import ops


def stft(x, fft_size, hop_size, win_length, window):
    """Magnitude spectrogram (B, frames, bins); stft_loss.py:9-30.  Utility kept
    for API compatibility (plain torch.stft); the loss classes below do not use it."""
    z = torch.stft(x, fft_size, hop_size, win_length, window, return_complex=True)
    return torch.sqrt(torch.clamp(z.real ** 2 + z.imag ** 2, min=1e-7)).transpose(2, 1)


class SpectralConvergenceLoss(torch.nn.Module):
    def forward(self, x_mag, y_mag):          # stft_loss.py:40-50
        return torch.norm(y_mag - x_mag, p="fro") / torch.norm(y_mag, p="fro")


class LogSTFTMagnitudeLoss(torch.nn.Module):
    def forward(self, x_mag, y_mag):          # stft_loss.py:60-69
        return F.l1_loss(torch.log(y_mag), torch.log(x_mag))


def _check_band(band):
    if band != "full":
        raise NotImplementedError(
            "band=%r: only band='full' (config/tiny.json:33) has a CUDA kernel; the reference's "
            "'high' branch slices frames, not frequencies (stft_loss.py:103-106)" % (band,))


class STFTLoss(torch.nn.Module):
    """stft_loss.py:72-113."""

    def __init__(self, fft_size=1024, shift_size=120, win_length=600, window="hann_window", band="full"):
        super().__init__()
        self.fft_size = fft_size
        self.shift_size = shift_size
        self.win_length = win_length
        self.band = band
        self.spectral_convergence_loss = SpectralConvergenceLoss()
        self.log_stft_magnitude_loss = LogSTFTMagnitudeLoss()
        self.register_buffer("window", getattr(torch, window)(win_length))

    def forward(self, x, y):
        _check_band(self.band)
        cfg = dict(fft_sizes=[self.fft_size], hop_sizes=[self.shift_size], win_lengths=[self.win_length],
                   sc_lambda=1.0, mag_lambda=1.0)
        _, sc, mag = ops.mrstft_l1(x, y, [self.window], cfg)
        return sc, mag


class MultiResolutionSTFTLoss(torch.nn.Module):
    """stft_loss.py:116-166.  forward(x, y) -> (sc_loss, mag_loss)."""

    def __init__(self, fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240],
                 window="hann_window", sc_lambda=0.1, mag_lambda=0.1, band="full"):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.sc_lambda = sc_lambda
        self.mag_lambda = mag_lambda
        self.band = band
        self.stft_losses = torch.nn.ModuleList(
            [STFTLoss(fs, ss, wl, window, band) for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths)])

    def _cfg(self):
        return dict(fft_sizes=[f.fft_size for f in self.stft_losses],
                    hop_sizes=[f.shift_size for f in self.stft_losses],
                    win_lengths=[f.win_length for f in self.stft_losses],
                    sc_lambda=float(self.sc_lambda), mag_lambda=float(self.mag_lambda))

    def forward_with_l1(self, x, y):
        """(l1, sc_loss, mag_loss) from ONE fused launch (L1 of util.py:239 included)."""
        _check_band(self.band)
        if x.dim() == 3:
            x = x.reshape(-1, x.size(2))
            y = y.reshape(-1, y.size(2))
        if len(self.stft_losses) > 3:
            raise NotImplementedError("at most 3 resolutions per fused launch")
        return ops.mrstft_l1(x, y, [f.window for f in self.stft_losses], self._cfg())

    def forward(self, x, y):
        _, sc, mag = self.forward_with_l1(x, y)
        return sc, mag
