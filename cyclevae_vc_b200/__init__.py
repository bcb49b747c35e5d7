"""cyclevae_vc_b200 -- B200-native (sm_100a) GRU-VAE encoder/decoder hot path of patrickltobing/cyclevae-vc.

Importing the package loads libcyclevae_b200.so; there is no CPU or eager-PyTorch fallback."""
from . import _lib  # noqa: F401  (raises if the CUDA library is missing)
from .gru_vae import (GRU_RNN, TWFSEloss, TwoSidedDilConv1d, initialize, kl_per_utt, loss_vae, mcd_l1_per_utt,  # noqa: F401
                      reparam_concat, sampling_vae_batch)

__version__ = "0.1.0"
