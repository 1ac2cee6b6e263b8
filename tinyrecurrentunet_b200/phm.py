"""Phase-aware beta-sigmoid mask: drop-in for the reference's phm.py.

The hot path never calls this module: the mask is fused with mod_phase and the
iSTFT in csrc/backend.cu (ops.mask_istft).  The class is kept, with the
reference's signature (phm.py:7-45), for code that builds the mask on its own
complex spectrograms; it is a handful of elementwise torch ops on the caller's
device with the two misspelt names of phm.py:41 fixed."""
import torch
import torch.nn as nn


class PhaseAwareMask(nn.Module):
    def __init__(self, beta=0.5):
        super().__init__()
        self.beta = beta

    def forward(self, mixture, estimated):
        mag_mixture = torch.abs(mixture)
        dphi = torch.angle(mixture) - torch.angle(estimated)
        return torch.sigmoid(self.beta * dphi) * mag_mixture
