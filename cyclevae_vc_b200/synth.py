"""Synthetic mel-cepstral batches with the layout of the reference's HDF5 features:
/feat_org_lf0 = [uv, log-F0, codeap x2, mcep x50] (feature_extract_vc.py:380-381) and
/cvuvlogf0fil_ap = [cv-uv, cv-log-F0, codeap x2] (feature_cv_extract_vc.py:133).  Host-side data
generation for the benchmark and examples (there is no dataset in the container)."""
from __future__ import annotations

import math

import numpy as np
import torch


def feature_stats(n_mcep: int = 50):
    """Generating moments; they double as the scale_in / scale_out statistics (train_*.py:344-347)."""
    mean = np.concatenate(([0.6, 5.3, -8.0, -8.0, 1.5], np.zeros(n_mcep - 1)))
    std = np.concatenate(([math.sqrt(0.24), 0.25, 4.0, 4.0, 1.5], 0.8 * 0.95 ** np.arange(1, n_mcep)))
    return mean, std


def make_batch(B: int, T: int, seed: int, n_spk: int = 2, n_mcep: int = 50, pin: bool = False):
    """h_src [B,T,4+n_mcep], cv_src [B,T,4], one-hot src/trg speaker codes [B,T,n_spk] (CPU fp32)."""
    g = torch.Generator().manual_seed(1234 + seed)
    mean, std = feature_stats(n_mcep)
    mean_t, std_t = torch.tensor(mean, dtype=torch.float32), torch.tensor(std, dtype=torch.float32)

    def feats(ncol):
        z = torch.randn(B, T, ncol, generator=g) * std_t[:ncol] + mean_t[:ncol]
        z[:, :, 0] = (torch.rand(B, T, generator=g) < 0.6).float()
        return z

    h, cv = feats(4 + n_mcep), feats(4)
    src = torch.zeros(B, T, n_spk)
    trg = torch.zeros(B, T, n_spk)
    spk = torch.arange(B) % n_spk
    src[torch.arange(B), :, spk] = 1.0
    trg[torch.arange(B), :, (spk + 1) % n_spk] = 1.0
    out = (h, cv, src, trg)
    if pin and torch.cuda.is_available():
        out = tuple(t.pin_memory() for t in out)
    return out


def build_models(hidden_units=1024, lat_dim=32, n_spk=2, n_mcep=50, stdim=4, do_prob=0.5, seed=1, device="cuda"):
    """Encoder / decoder exactly as the trainer builds them (train_*.py:310-347): reference `initialize`
    under a seed, then scale_in / scale_out overwritten by the feature statistics and frozen (:369-372)."""
    from .gru_vae import GRU_RNN, initialize
    torch.manual_seed(seed)
    enc = GRU_RNN(in_dim=stdim + n_mcep, out_dim=2 * lat_dim, hidden_units=hidden_units, do_prob=do_prob, scale_out_flag=False)
    dec = GRU_RNN(in_dim=lat_dim + n_spk, out_dim=n_mcep, hidden_units=hidden_units, do_prob=do_prob, scale_in_flag=False)
    enc.apply(initialize)
    dec.apply(initialize)
    mean, std = feature_stats(n_mcep)
    enc.scale_in.weight = torch.nn.Parameter(torch.diag(torch.tensor(1.0 / std, dtype=torch.float32)).unsqueeze(2))
    enc.scale_in.bias = torch.nn.Parameter(torch.tensor(-(mean / std), dtype=torch.float32))
    dec.scale_out.weight = torch.nn.Parameter(torch.diag(torch.tensor(std[stdim:], dtype=torch.float32)).unsqueeze(2))
    dec.scale_out.bias = torch.nn.Parameter(torch.tensor(mean[stdim:], dtype=torch.float32))
    for p in list(enc.scale_in.parameters()) + list(dec.scale_out.parameters()):
        p.requires_grad = False
    if device is not None:
        enc, dec = enc.to(device), dec.to(device)
    y0_dec = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1)   # train_*.py:359
    return enc, dec, y0_dec
