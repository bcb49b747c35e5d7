"""Drive the reference's OWN caller code through a `gru_vae`-compatible module (test infrastructure).

The trainer and the decoder scripts cannot be imported (h5py / dtw_c / pysptk / pyworld are absent), but their hot-path
call sites can be executed as written: the statements are cut out of the scripts' ASTs and exec'd in a namespace
that provides the names they use --

    train_gru_cyclevae_gauss_batch.py   train_generator (:45-149), save_checkpoint (:152-167) and the body of the
                                        frame-chunk branch of the training loop (:1293-1474: the 5 x n_cyc GRU_RNN
                                        passes with carried state, loss assembly, zero_grad / backward / optimizer.step)
    decode_gru-cyclevae_gauss.py        the `with torch.no_grad():` conversion block (:302-323)

once with the reference's classes on the CPU (oracle/make_golden.py -> tests/golden/callers.npz) and once with the
drop-in on the GPU (tests/test_gpu_parity.py).  The scripts are read from oracle/_ref/ (vendored, git-ignored copy made
by __graft_entry__.build()) or from /root/reference in the authoring container.

Stand-ins, because their originals are not part of the hot path: `dtw.calc_mcd` (mean mel-cepstral distortion of two
equally long sequences; only logged), `sampling_vae_batch` (a wrapper that draws its noise from a seeded CPU generator
so that both sides see the same noise, then calls the module's own function / formula), logging.
"""
from __future__ import annotations

import ast
import os
import types
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAINER = "train_gru_cyclevae_gauss_batch.py"
DECODER = "decode_gru-cyclevae_gauss.py"


def script_path(name: str):
    for d in (os.path.join(ROOT, "oracle", "_ref"), os.path.join(os.environ.get("CYCLEVAE_REFERENCE", "/root/reference"), "src", "bin")):
        p = os.path.join(d, name)
        if os.path.exists(p):
            return p
    return None


def _src(node, text):
    return ast.get_source_segment(text, node) or ""


def _function(tree, name):
    return next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name)


def _compile(stmts, tag):
    return compile(ast.Module(body=list(stmts), type_ignores=[]), tag, "exec")


def trainer_pieces():
    """-> (namespace with train_generator / save_checkpoint, code object of the frame-chunk branch body)"""
    path = script_path(TRAINER)
    text = open(path).read()
    tree = ast.parse(text)
    ns = {"np": np, "torch": torch, "os": os, "logging": _Quiet()}
    exec(_compile([_function(tree, "train_generator"), _function(tree, "save_checkpoint")], TRAINER), ns)
    branch = [n for n in ast.walk(tree) if isinstance(n, ast.If) and _src(n.test, text) == "args.batch_size > 0"
              and "optimizer.step()" in _src(n, text) and "prev_featfile_src == featfile_src" in _src(n, text)]
    assert len(branch) == 1, "the frame-chunk branch of the training loop was not found"
    return ns, _compile(branch[0].body, TRAINER + ":chunk-branch")


def decoder_block():
    """-> code object of the conversion block of decode_RNN (encoder, 300-sample latent mean, 3 decoder passes)"""
    path = script_path(DECODER)
    text = open(path).read()
    tree = ast.parse(text)
    blocks = [n for n in ast.walk(tree) if isinstance(n, ast.With) and "n_smpl_dec" in _src(n, text) and "cvmcep_trg" in _src(n, text)
              and "model_encoder" in _src(n, text)]
    inner = min(blocks, key=lambda n: len(_src(n, text)))
    return _compile([inner], DECODER + ":conversion")


class _Quiet:
    def info(self, *a, **k):
        pass

    warn = warning = debug = info


def calc_mcd(a, b):
    """Stand-in for dtw_c.calc_mcd (not in the reference tree): mean mel-cepstral distortion in dB of two aligned sequences."""
    d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    mcd = (10.0 / np.log(10.0)) * np.sqrt(2.0 * np.sum(d * d, axis=1))
    return float(np.mean(mcd)), mcd


class _AutoNS(dict):
    """Locals of the exec'd loop body: the trainer pre-creates dozens of per-cycle lists (train_*.py:380-455, 497-600);
    unknown names of those families materialise here on first use."""

    def __init__(self, n_cyc, *a, **k):
        super().__init__(*a, **k)
        self.n_cyc = n_cyc

    def __missing__(self, key):
        if key.startswith(("batch_", "y_in_", "h_in_")):
            v = [None] * self.n_cyc
        elif key.startswith(("loss_", "gv_", "mcdpow_", "mcd_", "lat_dist_")):
            v = [[] for _ in range(self.n_cyc)]
        else:
            raise KeyError(key)
        self[key] = v
        return v


def fake_loader_batch(x, cv, sc, tc, flens, pad):
    """One DataLoader batch in the layout of FeatureDatasetSingleVAE + padding (dataset.py:67-98): zero-padded to `pad`."""
    n = x.shape[0]

    def padded(t):
        out = torch.zeros(n, pad, t.shape[2])
        out[:, :t.shape[1]] = t
        return out

    spc = [np.arange(f) for f in flens]
    idx = torch.stack([torch.tensor(np.pad(s, (0, pad - len(s)))) for s in spc])
    xs = padded(x)
    for j, f in enumerate(flens):      # frames beyond an utterance's length are padding
        xs[j, f:] = 0
    return {"flen_src": torch.tensor(flens), "flen_spc_src": torch.tensor([len(s) for s in spc]),
            "flen_src_trg": torch.tensor(flens), "flen_spc_src_trg": torch.tensor([len(s) for s in spc]),
            "h_src": xs, "src_code": padded(sc), "trg_code": padded(tc), "cv_src": padded(cv), "h_src_trg": xs.clone(),
            "spcidx_src": idx, "spcidx_src_trg": idx.clone(),
            "featfile_src": ["spkA/utt%d.h5" % j for j in range(n)], "featfile_src_trg": ["spkB/utt%d.h5" % j for j in range(n)]}


def run_trainer_chunks(mod, enc, dec, device, *, x, cv, sc, tc, flens, lat_dim, n_cyc, y0_dec, sample, batch_size=80,
                       checkpoint_dir=None, checkpoint_after=None):
    """Run the reference trainer's chunk loop over ONE utterance batch with `mod`'s classes (`enc` / `dec` are GRU_RNN
    instances of that module, already on `device`).  Returns per-chunk losses and the final carried states."""
    ns_fn, body = trainer_pieces()
    n = x.shape[0]
    gen = ns_fn["train_generator"]([fake_loader_batch(x, cv, sc, tc, flens, max(flens) + 7)], device, batch_size=batch_size)
    args = Namespace(n_cyc=n_cyc, lat_dim=lat_dim, batch_size=batch_size, batch_size_utt=n, spk_src="spkA", epoch_count=1)
    stdim = 4
    params = [p for m in (enc, dec) for sub in (m.conv, m.gru, m.out_1) for p in sub.parameters()]   # train_*.py:373-376
    optimizer = torch.optim.Adam(params, lr=1e-4)                                                     # :377
    y_in_pp = torch.zeros(n, 1, 2 * lat_dim, device=device)                                           # :357-358
    y_in = y0_dec.to(device).repeat(n, 1, 1)
    g = {"np": np, "torch": torch, "os": os, "logging": _Quiet(), "time": __import__("time"), "Variable": torch.autograd.Variable,
         "dtw": types.SimpleNamespace(calc_mcd=calc_mcd)}
    loc = _AutoNS(n_cyc, args=args, model_encoder=enc, model_decoder=dec, optimizer=optimizer, criterion_mcd=mod.TWFSEloss(),
                  loss_vae=mod.loss_vae, sampling_vae_batch=sample, stdim=stdim, stdim_=stdim + 1, half_cyc=False,
                  y_in_pp=y_in_pp, y_in_src=y_in, y_in_trg=y_in, y_in_pp_mod=y_in_pp, y_in_src_mod=y_in, y_in_trg_mod=y_in,
                  iter_idx=0, iter_count=0, epoch_idx=0, prev_featfile_src=None, loss=[], total=[], start=0.0)
    names = ("batch_src", "batch_src_src_code", "batch_src_trg_code", "batch_src_trg", "batch_cv_src", "src_idx_s", "src_idx_e",
             "spcidx_src_s_idx", "spcidx_src_e_idx", "c_idx_src", "utt_idx_src", "spcidx_src", "spcidx_src_trg", "featfile_src",
             "featfile_src_trg", "flens_src", "flens_src_trg", "flens_spc_src", "flens_spc_src_trg", "select_utt_idx", "flen_acc",
             "n_batch_utt")
    losses = []
    while True:
        vals = next(gen)                       # the unpacking of train_*.py:671
        if vals[9] < 0:                        # c_idx_src < 0: end of the utterance batch
            break
        if loc["iter_count"] > 0:
            loc["prev_flens_src"] = loc["flens_src"]
        loc.update(dict(zip(names, vals)))
        exec(body, g, loc)
        losses.append(float(loc["batch_loss"].item()))
        if checkpoint_dir is not None and len(losses) == checkpoint_after:
            # the every-epoch round trip of save_checkpoint: .cpu(), state_dict, torch.save, .cuda() (train_*.py:152-167)
            saved = torch.Tensor.cuda
            if device.type == "cpu":
                torch.Tensor.cuda = lambda self, *a, **k: self
                torch.nn.Module.cuda = lambda self, *a, **k: self
            try:
                ns_fn["save_checkpoint"](checkpoint_dir, enc, dec, optimizer, np.random.get_state(), torch.get_rng_state(), len(losses))
            finally:
                if device.type == "cpu":
                    torch.Tensor.cuda = saved
                    del torch.nn.Module.cuda
    return {"losses": losses, "y_pp": loc["y_in_pp_src"][n_cyc - 1].detach().cpu().numpy(), "h_dec": loc["h_in_src_trg_src"][0].detach().cpu().numpy(),
            "trj": loc["batch_trj_src_trg_src"][n_cyc - 1].detach().cpu().numpy(), "iter_count": loc["iter_count"]}


def run_decoder_block(enc, dec, device, *, feat, feat_trg, lat_dim, n_smpl, y0_dec, sample):
    """Run decode_*.py:302-323 as written.  feat / feat_trg: numpy [T, 54]."""
    block = decoder_block()
    saved = torch.Tensor.cuda
    if device.type == "cpu":                   # the script hard-codes .cuda()
        torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        loc = {"model_encoder": enc, "model_decoder": dec, "feat": feat, "feat_trg": feat_trg, "sampling_vae_batch": sample,
               "config": Namespace(lat_dim=lat_dim), "args": Namespace(n_smpl_dec=n_smpl),
               "y_in_pp": torch.zeros(1, 1, 2 * lat_dim, device=device), "y_in_src": y0_dec.to(device), "y_in_trg": y0_dec.to(device)}
        exec(block, {"np": np, "torch": torch}, loc)
    finally:
        if device.type == "cpu":
            torch.Tensor.cuda = saved
    return {k: np.asarray(loc[k]) for k in ("cvmcep", "cvmcep_src", "cvmcep_trg")}


def seeded_sampler(mod, lat_dim_default, seed, device, native):
    """`sampling_vae_batch` stand-in: noise from a seeded CPU generator in call order (the same sequence on both sides)."""
    gen = torch.Generator().manual_seed(seed)

    def sample(param, lat_dim=None, training=False, relu_vae=False):
        lat = lat_dim or lat_dim_default
        eps = torch.randn(param.shape[:-1] + (lat,), generator=gen).to(device)
        if native:
            return mod.sampling_vae_batch(param, lat_dim=lat, eps=eps)
        return param[..., :lat] + torch.exp(param[..., lat:] / 2) * eps      # gru_vae.py:85-98 without its .cuda()

    return sample
