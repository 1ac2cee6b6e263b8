"""tinyrecurrentunet_b200: B200 (sm_100a) implementation of the TRU-Net hot path
behind the reference's Python module API (network / phm / stft_loss / dataset /
util / distributed).  Importing the package loads libtru_b200.so; there is no
CPU or stock-PyTorch fallback for the hot path."""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)

__all__ = ["_lib", "ops", "network", "phm", "stft_loss", "cos_loss", "dataset", "util", "distributed", "optim"]
__version__ = "0.1.0"
