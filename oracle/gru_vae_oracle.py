"""CPU ORACLE for the GRU-VAE hot path of patrickltobing/cyclevae-vc.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product
(cyclevae_vc_b200/) never routes through this file and has no CPU fallback.

It is a from-scratch, functional restatement (plain torch CPU ops, explicit
parameter dict, explicit noise/dropout masks) of

    /root/reference/src/nets/gru_vae.py
        TwoSidedDilConv1d.forward   :53-66
        sampling_vae_batch          :85-98
        loss_vae                    :117-123
        GRU_RNN.forward             :322-455   (batched + unbatched layouts)
        TWFSEloss.forward           :521-534   (twf=None, rmse=False branch)
        initialize                  :21-33
    /root/reference/src/bin/train_gru_cyclevae_gauss_batch.py
        train_generator (chunking)  :70-134
        cyc graph (first chunk)     :1326-1338, (carried chunk) :1298-1311
        loss assembly               :1363-1410 (incl. the KL-cv cat quirk :1393)
    /root/reference/src/bin/decode_gru-cyclevae_gauss.py
        conversion composition      :302-323

Parity pin: the reference ships no tests/golden vectors, so the pin is
the reference module itself, imported in the authoring container by
oracle/make_golden.py, which writes tests/golden/*.npz; tests/test_oracle_golden.py
checks this restatement against those fixtures.

All functions are dtype-generic (fp32 for parity, fp64 for gradient checks) and
differentiable through torch autograd, so `backward` of the oracle is autograd of
this restatement (verified against the reference's autograd by make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LOG_VAR_FLOOR = -13.815510557964274104107948728106  # gru_vae.py:412 (ln 1e-6)
MCD_COEF = (10.0 / 2.3025850929940456840179914546844) * 1.4142135623730950488016887242097  # gru_vae.py:525

Params = Dict[str, torch.Tensor]


@dataclass(frozen=True)
class NetSpec:
    """Constructor arguments of GRU_RNN (gru_vae.py:282) that the callers use."""
    in_dim: int
    out_dim: int
    hidden_units: int = 1024
    kernel_size: int = 3
    dilation_size: int = 2        # number of conv layers (gru_vae.py:300)
    do_prob: float = 0.0
    scale_in: bool = True
    scale_out: bool = True

    @property
    def rec_field(self) -> int:          # gru_vae.py:44
        return self.kernel_size ** self.dilation_size

    @property
    def padding(self) -> int:            # gru_vae.py:45
        return (self.rec_field - 1) // 2

    @property
    def conv_dim(self) -> int:
        return self.in_dim * self.rec_field

    @property
    def tot_in_dim(self) -> int:         # gru_vae.py:302
        return self.conv_dim + self.out_dim

    def param_shapes(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """state_dict keys/shapes in the reference's registration order (gru_vae.py:295-320)."""
        k, H = self.kernel_size, self.hidden_units
        out: List[Tuple[str, Tuple[int, ...]]] = []
        if self.scale_in:
            out += [("scale_in.weight", (self.in_dim, self.in_dim, 1)), ("scale_in.bias", (self.in_dim,))]
        for i in range(self.dilation_size):
            cin = self.in_dim * k ** i
            out += [(f"conv.conv.{i}.weight", (cin * k, cin, k)), (f"conv.conv.{i}.bias", (cin * k,))]
        out += [("gru.weight_ih_l0", (3 * H, self.tot_in_dim)), ("gru.weight_hh_l0", (3 * H, H)),
                ("gru.bias_ih_l0", (3 * H,)), ("gru.bias_hh_l0", (3 * H,)),
                ("out_1.weight", (self.out_dim, H, 1)), ("out_1.bias", (self.out_dim,))]
        if self.scale_out:
            out += [("scale_out.weight", (self.out_dim, self.out_dim, 1)), ("scale_out.bias", (self.out_dim,))]
        return out

    def n_trainable(self) -> int:
        return sum(int(np.prod(s)) for n, s in self.param_shapes() if not n.startswith("scale_"))


def encoder_spec(in_dim=54, lat_dim=32, hidden_units=1024, kernel_size=3, dilation_size=2, do_prob=0.5) -> NetSpec:
    """train_gru_cyclevae_gauss_batch.py:310-318"""
    return NetSpec(in_dim, 2 * lat_dim, hidden_units, kernel_size, dilation_size, do_prob, True, False)


def decoder_spec(lat_dim=32, n_spk=2, out_dim=50, hidden_units=1024, kernel_size=3, dilation_size=2, do_prob=0.5) -> NetSpec:
    """train_gru_cyclevae_gauss_batch.py:320-328"""
    return NetSpec(lat_dim + n_spk, out_dim, hidden_units, kernel_size, dilation_size, do_prob, False, True)


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def init_params(spec: NetSpec, seed: int, *, gain: float = 1.0, bias_std: float = 0.0,
                mean: Optional[np.ndarray] = None, scale: Optional[np.ndarray] = None,
                dtype=torch.float32) -> Params:
    """Xavier-uniform weights / zero biases (gru_vae.py:21-33) from a numpy PCG64 stream,
    then scale_in / scale_out overwritten by statistics as the trainer does
    (train_gru_cyclevae_gauss_batch.py:344-347).  gain>1 / bias_std>0 give the
    "trained-like" stress set of SURVEY.md §8(d)."""
    rng = np.random.default_rng(seed)
    P: Params = {}
    for name, shape in spec.param_shapes():
        if name.endswith("weight") or "weight_" in name:
            if len(shape) == 3:
                fan_in, fan_out = shape[1] * shape[2], shape[0] * shape[2]
            else:
                fan_in, fan_out = shape[1], shape[0]
            bound = gain * math.sqrt(6.0 / (fan_in + fan_out))
            arr = rng.uniform(-bound, bound, size=shape)
        else:
            arr = rng.normal(0.0, bias_std, size=shape) if bias_std > 0 else np.zeros(shape)
        P[name] = torch.tensor(arr, dtype=dtype)
    if spec.scale_in:
        mu = np.zeros(spec.in_dim) if mean is None else np.asarray(mean, dtype=np.float64)
        sd = np.ones(spec.in_dim) if scale is None else np.asarray(scale, dtype=np.float64)
        P["scale_in.weight"] = torch.tensor(np.diag(1.0 / sd)[:, :, None], dtype=dtype)
        P["scale_in.bias"] = torch.tensor(-(mu / sd), dtype=dtype)
    if spec.scale_out:
        mu = np.zeros(spec.out_dim) if mean is None else np.asarray(mean, dtype=np.float64)
        sd = np.ones(spec.out_dim) if scale is None else np.asarray(scale, dtype=np.float64)
        P["scale_out.weight"] = torch.tensor(np.diag(sd)[:, :, None], dtype=dtype)
        P["scale_out.bias"] = torch.tensor(mu, dtype=dtype)
    return P


def params_checksum(P: Params) -> float:
    return float(sum(P[k].double().abs().sum().item() * (i + 1) for i, k in enumerate(sorted(P))))


# --------------------------------------------------------------------------------------
# front-end
# --------------------------------------------------------------------------------------
def frontend(P: Params, spec: NetSpec, x_btc: torch.Tensor) -> torch.Tensor:
    """scale_in (gru_vae.py:336) then the two-sided dilated stack (gru_vae.py:47-51,62-66).
    [B,T,in] -> [B,T,in*k^layers].  Layer 0 pads `padding` zeros both sides of the
    *normalised* signal; layer i>0 has dilation k^i and no padding."""
    x = x_btc.transpose(1, 2)
    if spec.scale_in:
        x = F.conv1d(x, P["scale_in.weight"], P["scale_in.bias"])
    k = spec.kernel_size
    for i in range(spec.dilation_size):
        x = F.conv1d(x, P[f"conv.conv.{i}.weight"], P[f"conv.conv.{i}.bias"],
                     dilation=k ** i, padding=spec.padding if i == 0 else 0)
    return x.transpose(1, 2)


# --------------------------------------------------------------------------------------
# GRU_RNN.forward
# --------------------------------------------------------------------------------------
def gru_rnn_forward(P: Params, spec: NetSpec, x: torch.Tensor, y_in: torch.Tensor,
                    h_in: Optional[torch.Tensor] = None, *,
                    mask_conv: Optional[torch.Tensor] = None, mask_gru: Optional[torch.Tensor] = None,
                    clamp_vae: bool = False, lat_dim: int = 16, internals: Optional[dict] = None):
    """GRU_RNN.forward (gru_vae.py:322-455) for the kwargs the callers use
    (do / clamp_vae / lat_dim / h_in; hidden_layers == 1).

    x: [B,T,in] or [T,in]; y_in: [B,1,out]; h_in: [1,B,H] or None.
    mask_conv [B,T,conv_dim] / mask_gru [B,T,H]: dropout masks already scaled by 1/(1-p)
    (the reference draws them inside nn.Dropout at :355/:369/:380; here they are inputs so
    both sides of a parity test can use the same mask).  None == eval / do=False.
    Returns (trj_out, y_last [B,1,out], h [1,B,H]) exactly like the reference.
    """
    batched = x.dim() > 2
    xb = x if batched else x.unsqueeze(0)
    B, T, _ = xb.shape
    H = spec.hidden_units
    C = spec.conv_dim
    x_conv = frontend(P, spec, xb)                                  # :354-357
    if mask_conv is not None:
        x_conv = x_conv * mask_conv
    W_ih, W_hh = P["gru.weight_ih_l0"], P["gru.weight_hh_l0"]
    b_ih, b_hh = P["gru.bias_ih_l0"], P["gru.bias_hh_l0"]
    W_o, b_o = P["out_1.weight"][:, :, 0], P["out_1.bias"]
    W_ih_t, W_hh_t, W_o_t = W_ih.t(), W_hh.t(), W_o.t()
    h = xb.new_zeros(B, H) if h_in is None else h_in[0]
    y = y_in[:, 0, :]
    ys = []
    hs = [] if internals is not None else None
    for t in range(T):                                              # :364-399
        gi = torch.addmm(b_ih, torch.cat((x_conv[:, t], y), 1), W_ih_t)
        gh = torch.addmm(b_hh, h, W_hh_t)
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])                   # gate order r,z,n
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        o = h if mask_gru is None else h * mask_gru[:, t]           # :369/:380 (h itself is not dropped)
        y = torch.addmm(b_o, o, W_o_t)                              # :371/:393
        ys.append(y)
        if hs is not None:
            hs.append(h)
    trj = torch.stack(ys, 1)                                        # [B,T,out]
    if internals is not None:
        internals["x_conv"] = x_conv
        internals["h_all"] = torch.stack(hs, 1)
        internals["y_all"] = trj
    if spec.scale_out:                                              # :402-406
        trj_out = F.conv1d(trj.transpose(1, 2), P["scale_out.weight"], P["scale_out.bias"]).transpose(1, 2)
    else:
        trj_out = trj
        if clamp_vae:                                               # :408-412 / :423-426
            trj_out = torch.cat((trj_out[:, :, :lat_dim],
                                 torch.clamp(trj_out[:, :, lat_dim:], min=LOG_VAR_FLOOR)), 2)
    if not batched:
        trj_out = trj_out.squeeze(0) if spec.scale_out else trj_out.reshape(-1, spec.out_dim)
    return trj_out, y.unsqueeze(1), h.unsqueeze(0)


# --------------------------------------------------------------------------------------
# VAE sampling and losses
# --------------------------------------------------------------------------------------
def sampling_vae_batch(param: torch.Tensor, eps: torch.Tensor, lat_dim: Optional[int] = None) -> torch.Tensor:
    """gru_vae.py:85-96 with the noise passed in (the reference draws eps with the CPU
    generator and calls .cuda(), :91/:94)."""
    if lat_dim is None:
        lat_dim = param.shape[-1] // 2
    return param[..., :lat_dim] + torch.exp(param[..., lat_dim:] / 2) * eps


def loss_vae(param: torch.Tensor, lat_dim: Optional[int] = None) -> torch.Tensor:
    """gru_vae.py:117-123: mean over frames of the diagonal-Gaussian KL (log-variance form)."""
    if lat_dim is None:
        lat_dim = param.shape[1] // 2
    mu, sigma = param[:, :lat_dim], param[:, lat_dim:]
    return torch.mean(0.5 * torch.sum(torch.exp(sigma) + mu * mu - sigma - 1.0, 1))


def mcd_l1(x: torch.Tensor, y: torch.Tensor):
    """TWFSEloss.forward(x, y, L2=False, GV=False) (gru_vae.py:525-528,534):
    per-frame (10/ln10)*sqrt(2)*sum_d|x-y|; returns (sum, mean, unbiased std)."""
    mcd = MCD_COEF * torch.sum(torch.abs(x - y), 1)
    return torch.sum(mcd), torch.mean(mcd), torch.std(mcd)


# --------------------------------------------------------------------------------------
# trainer bookkeeping (integers; bit-exact domain)
# --------------------------------------------------------------------------------------
def chunk_schedule(flens: Sequence[int], batch_size: int, spcidx: Optional[Sequence[Sequence[int]]] = None):
    """Frame-chunk schedule of train_generator (train_*.py:70-134): for every yielded chunk
    (src_idx_s, src_idx_e, spcidx_s_idx[], spcidx_e_idx[], flen_acc[], select_utt_idx[]).
    `spcidx[j]` are the speech-frame indices of utterance j (dataset.py:77); default = every
    frame is a speech frame.  Integer-only: tests demand bit-exact equality with the reference.

    Reference behaviours kept on purpose: the first chunk's src_idx_e is batch_size-1 even
    when the longest utterance is shorter (:72, slices clip later); flen_acc starts at
    batch_size (:77) and is lowered only when a later chunk crosses an utterance end (:107-108);
    an utterance is selected while its speech cursor has not reached its last index (:106)."""
    flens = [int(f) for f in flens]
    n = len(flens)
    if spcidx is None:
        spcidx = [list(range(f)) for f in flens]
    nspc = [len(s) for s in spcidx]
    max_flen = max(flens)
    cur_s, cur_e = [-1] * n, [-1] * n
    looking_for_start = [True] * n            # == (not s_flag) and e_flag in the reference

    def scan(j, lo, hi):                      # :79-98 / :109-128
        for i in range(cur_e[j] + 1, nspc[j]):
            v = spcidx[j][i]
            if looking_for_start[j] and v >= lo:
                if v > hi:
                    cur_s[j] = -1
                    return
                cur_s[j] = i
                looking_for_start[j] = False
                if i == nspc[j] - 1:
                    cur_e[j] = i
                    looking_for_start[j] = True
                    return
            elif (not looking_for_start[j]) and (v >= hi or i == nspc[j] - 1):
                cur_e[j] = i - 1 if v > hi else i
                looking_for_start[j] = True
                return

    s, e = 0, batch_size - 1
    flen_acc = [batch_size] * n
    for j in range(n):
        scan(j, s, e)
    out = [(s, e, list(cur_s), list(cur_e), list(flen_acc), list(range(n)))]
    while e < max_flen - 1:
        s = e + 1
        e = min(s + batch_size - 1, max_flen - 1)
        sel = []
        for j in range(n):
            if cur_e[j] < nspc[j] - 1:
                if e >= flens[j]:
                    flen_acc[j] = flens[j] - s
                scan(j, s, e)
                sel.append(j)
        out.append((s, e, list(cur_s), list(cur_e), list(flen_acc), sel))
    return out


# --------------------------------------------------------------------------------------
# the trainer's cyc graph and loss (one optimizer step's worth of model work)
# --------------------------------------------------------------------------------------
CYC_PASSES = ("pp_src", "src_src", "src_trg", "pp_src_trg", "src_trg_src")


def cyc_forward(Pe: Params, Pd: Params, enc: NetSpec, dec: NetSpec, *, x: torch.Tensor, cv: torch.Tensor,
                src_code: torch.Tensor, trg_code: torch.Tensor, n_cyc: int, lat_dim: int, stdim: int,
                y0_enc: torch.Tensor, y0_dec: torch.Tensor, eps: Sequence[Sequence[torch.Tensor]],
                masks: Optional[Sequence[Sequence[Tuple[torch.Tensor, torch.Tensor]]]] = None,
                state: Optional[dict] = None):
    """The 5 x n_cyc GRU_RNN passes of one chunk (train_*.py:1326-1338; with `state`
    = carried, detached (y, h) per pass it is :1298-1311).
    eps[i] = 3 noise tensors [B,T,lat]; masks[i][p] = (mask_conv, mask_gru) for pass p.
    Returns dict of per-cycle outputs and the new carried state."""
    out = {k: [] for k in ("lat_src", "trj_src_src", "trj_src_trg", "lat_src_trg", "trj_src_trg_src")}
    new_state = {}

    def run(P, spec, inp, name, i, p, **kw):
        if state is None:
            y0, h0 = (y0_enc if spec is enc else y0_dec), None
        else:
            y0, h0 = state[(name, i)]
            y0, h0 = y0.detach(), h0.detach()
        mc, mg = (None, None) if masks is None else masks[i][p]
        trj, y_last, h_last = gru_rnn_forward(P, spec, inp, y0, h0, mask_conv=mc, mask_gru=mg, **kw)
        new_state[(name, i)] = (y_last, h_last)
        return trj

    for i in range(n_cyc):
        enc_in = x if i == 0 else torch.cat((x[:, :, :stdim], out["trj_src_trg_src"][i - 1]), 2)
        lat_src = run(Pe, enc, enc_in, "pp_src", i, 0, clamp_vae=True, lat_dim=lat_dim)
        trj_src_src = run(Pd, dec, torch.cat((src_code, sampling_vae_batch(lat_src, eps[i][0], lat_dim)), 2), "src_src", i, 1)
        trj_src_trg = run(Pd, dec, torch.cat((trg_code, sampling_vae_batch(lat_src, eps[i][1], lat_dim)), 2), "src_trg", i, 2)
        lat_src_trg = run(Pe, enc, torch.cat((cv, trj_src_trg), 2), "pp_src_trg", i, 3, clamp_vae=True, lat_dim=lat_dim)
        trj_src_trg_src = run(Pd, dec, torch.cat((src_code, sampling_vae_batch(lat_src_trg, eps[i][2], lat_dim)), 2), "src_trg_src", i, 4)
        for k, v in (("lat_src", lat_src), ("trj_src_src", trj_src_src), ("trj_src_trg", trj_src_trg),
                     ("lat_src_trg", lat_src_trg), ("trj_src_trg_src", trj_src_trg_src)):
            out[k].append(v)
    return out, new_state


def cyc_loss(out: dict, x: torch.Tensor, *, n_cyc: int, lat_dim: int, stdim: int, flen_acc: Sequence[int],
             select_utt_idx: Sequence[int], kl_cv_quirk: bool = True):
    """Loss assembly of train_*.py:1363-1410 (not half_cyc).  With kl_cv_quirk the line-1393
    behaviour is reproduced: the per-cycle 'cv' vector is cat(batch_loss_lat_src[i], KL_cv of the
    LAST selected utterance) whenever more than one utterance is selected."""
    total = None
    parts = []
    for i in range(n_cyc):
        mcd_ss, mcd_sts, kl_s, kl_cv = [], [], [], []
        for j in select_utt_idx:
            F_ = int(flen_acc[j])
            tgt = x[j, :F_, stdim:]
            mcd_ss.append(mcd_l1(out["trj_src_src"][i][j, :F_], tgt)[1])
            mcd_sts.append(mcd_l1(out["trj_src_trg_src"][i][j, :F_], tgt)[1])
            kl_s.append(loss_vae(out["lat_src"][i][j, :F_], lat_dim))
            kl_cv.append(loss_vae(out["lat_src_trg"][i][j, :F_], lat_dim))
        s_ss, s_sts, s_kl = torch.stack(mcd_ss).sum(), torch.stack(mcd_sts).sum(), torch.stack(kl_s).sum()
        if kl_cv_quirk and len(select_utt_idx) > 1:
            s_cv = s_kl + kl_cv[-1]
        else:
            s_cv = torch.stack(kl_cv).sum()
        c = s_ss + s_sts + s_kl + s_cv
        total = c if total is None else total + c
        parts.append((s_ss, s_sts, s_kl, s_cv))
    return total, parts


# --------------------------------------------------------------------------------------
# stage-6 conversion (decode) composition
# --------------------------------------------------------------------------------------
def convert(Pe: Params, Pd: Params, enc: NetSpec, dec: NetSpec, feat: torch.Tensor, code: torch.Tensor, *,
            lat_dim: int, y0_enc: torch.Tensor, y0_dec: torch.Tensor, eps_mean: torch.Tensor) -> torch.Tensor:
    """decode_gru-cyclevae_gauss.py:303-305,318: encoder -> latent averaged over n_smpl_dec
    samples -> decoder with the target speaker code.  mean_k(mu + e^{s/2} eps_k) =
    mu + e^{s/2} mean_k(eps_k), so the average noise `eps_mean` [..,T,lat] is the input.
    feat: [T,in] (reference layout) or [B,T,in]."""
    with torch.no_grad():
        lat, _, _ = gru_rnn_forward(Pe, enc, feat, y0_enc, clamp_vae=True, lat_dim=lat_dim)
        lat_feat = sampling_vae_batch(lat, eps_mean, lat_dim)
        cv, _, _ = gru_rnn_forward(Pd, dec, torch.cat((code, lat_feat), -1), y0_dec)
    return cv


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synth_stats(n_mcep: int = 50) -> Tuple[np.ndarray, np.ndarray]:
    """Generating moments of the synthetic [uv, logF0, codeap x2, mcep x n] features."""
    mean = np.concatenate(([0.6, 5.3, -8.0, -8.0, 1.5], np.zeros(n_mcep - 1)))
    std = np.concatenate(([math.sqrt(0.24), 0.25, 4.0, 4.0, 1.5], 0.8 * 0.95 ** np.arange(1, n_mcep)))
    return mean, std


def synth_batch(B: int, T: int, seed: int, n_spk: int = 2, n_mcep: int = 50, dtype=torch.float32):
    """h_src [B,T,4+n_mcep], cv_src [B,T,4], one-hot src/trg codes [B,T,n_spk]."""
    g = torch.Generator().manual_seed(1234 + seed)
    mean, std = synth_stats(n_mcep)
    mean_t, std_t = torch.tensor(mean, dtype=torch.float64), torch.tensor(std, dtype=torch.float64)

    def feats(ncol):
        z = torch.randn(B, T, ncol, generator=g, dtype=torch.float64) * std_t[:ncol] + mean_t[:ncol]
        z[:, :, 0] = (torch.rand(B, T, generator=g, dtype=torch.float64) < 0.6).double()
        return z

    h = feats(4 + n_mcep).to(dtype)
    cv = feats(4).to(dtype)
    src = torch.zeros(B, T, n_spk, dtype=dtype)
    trg = torch.zeros(B, T, n_spk, dtype=dtype)
    spk = torch.arange(B) % n_spk
    src[torch.arange(B), :, spk] = 1.0
    trg[torch.arange(B), :, (spk + 1) % n_spk] = 1.0
    return h, cv, src, trg


def synth_noise(B: int, T: int, lat_dim: int, n_cyc: int, seed: int, dtype=torch.float32):
    g = torch.Generator().manual_seed(4321 + seed)
    return [[torch.randn(B, T, lat_dim, generator=g).to(dtype) for _ in range(3)] for _ in range(n_cyc)]


def synth_masks(B: int, T: int, enc: NetSpec, dec: NetSpec, n_cyc: int, seed: int, dtype=torch.float32):
    """Bernoulli keep-masks scaled by 1/(1-p) for the 5 passes of each cycle (enc,dec,dec,enc,dec)."""
    g = torch.Generator().manual_seed(9876 + seed)
    out = []
    for _ in range(n_cyc):
        cyc = []
        for spec in (enc, dec, dec, enc, dec):
            p = spec.do_prob
            if p <= 0:
                cyc.append((None, None))
                continue
            mc = (torch.rand(B, T, spec.conv_dim, generator=g) >= p).to(dtype) / (1.0 - p)
            mg = (torch.rand(B, T, spec.hidden_units, generator=g) >= p).to(dtype) / (1.0 - p)
            cyc.append((mc, mg))
        out.append(cyc)
    return out
