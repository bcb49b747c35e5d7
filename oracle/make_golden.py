"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference, which does not exist on the
GPU box):

    python oracle/make_golden.py            # writes tests/golden/

The fixtures pin oracle/gru_vae_oracle.py (tests/test_oracle_golden.py) and are the
reference-side truth for the CUDA parity tests (tests/test_gpu_*.py).  Nothing here is
copied from the reference: the module is imported from where it lies, its GRU_RNN /
loss_vae / TWFSEloss are called through their public signatures, and `train_generator`
is exec'd from the reference file's own AST (the trainer cannot be imported: h5py/dtw_c
are absent).  The only restated line is sampling (mu + exp(sigma/2)*eps) because
gru_vae.py:91/:94 hard-code .cuda().
"""
from __future__ import annotations

import ast
import os
import sys

sys.dont_write_bytecode = True
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CYCLEVAE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src", "nets"))
sys.path.insert(0, ROOT)
import gru_vae as ref  # noqa: E402  (the reference module)

from oracle import gru_vae_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class _MaskDrop(torch.nn.Module):
    """Stand-in for nn.Dropout that multiplies by supplied masks (one per call)."""

    def __init__(self, masks):
        super().__init__()
        self.masks, self.i = masks, 0

    def forward(self, x):
        m = self.masks[self.i]
        self.i += 1
        return x * m.reshape(x.shape)


def build_ref(spec: orc.NetSpec, P: orc.Params) -> "ref.GRU_RNN":
    m = ref.GRU_RNN(in_dim=spec.in_dim, out_dim=spec.out_dim, hidden_units=spec.hidden_units,
                    kernel_size=spec.kernel_size, dilation_size=spec.dilation_size, do_prob=spec.do_prob,
                    scale_in_flag=spec.scale_in, scale_out_flag=spec.scale_out)
    missing = m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(P[k].shape), k
    return m


def ref_forward(m, x, y_in, h_in=None, mask_conv=None, mask_gru=None, **kw):
    """Reference forward with injected dropout masks (do=True iff masks given)."""
    do = mask_conv is not None
    if do:
        T = x.shape[1] if x.dim() > 2 else x.shape[0]
        m.conv_drop = _MaskDrop([mask_conv])
        m.gru_drop = _MaskDrop([mask_gru[:, t:t + 1] for t in range(T)])
    return m(x, y_in, h_in=h_in, do=do, **kw)


def sub(t: torch.Tensor, stride: int) -> np.ndarray:
    """Sub-sample along the frame axis to keep fixtures small."""
    return t.detach().numpy()[..., ::stride, :].copy()


def stats_for(spec_in_dim: int):
    mean, std = orc.synth_stats(50)
    return mean[:spec_in_dim], std[:spec_in_dim]


# ----------------------------------------------------------------------------------------
def golden_tiny():
    """Tiny nets with the weights stored: forward (eval, dropout masks, h_in carry, unbatched),
    gradients from the reference's autograd, losses."""
    torch.manual_seed(0)
    enc = orc.NetSpec(in_dim=7, out_dim=6, hidden_units=20, do_prob=0.5, scale_in=True, scale_out=False)
    dec = orc.NetSpec(in_dim=5, out_dim=4, hidden_units=20, do_prob=0.5, scale_in=False, scale_out=True)
    rng = np.random.default_rng(7)
    Pe = orc.init_params(enc, 11, gain=2.0, bias_std=0.1, mean=rng.normal(size=7), scale=rng.uniform(0.5, 2, size=7))
    Pd = orc.init_params(dec, 12, gain=2.0, bias_std=0.1, mean=rng.normal(size=4), scale=rng.uniform(0.5, 2, size=4))
    # a non-diagonal scale_in / scale_out too: the reference stores full matrices (gru_vae.py:297,318)
    Pe["scale_in.weight"] = Pe["scale_in.weight"] + 0.05 * torch.randn(7, 7, 1)
    Pd["scale_out.weight"] = Pd["scale_out.weight"] + 0.05 * torch.randn(4, 4, 1)
    # log-variance channels at / below the floor so the clamp branch (gru_vae.py:412) and its
    # zero-gradient region are exercised
    Pe["out_1.bias"][4] = -15.0
    Pe["out_1.bias"][5] = -13.8
    me, md = build_ref(enc, Pe), build_ref(dec, Pd)
    B, T, lat = 3, 13, 3
    x = torch.randn(B, T, 7)
    y0e = torch.randn(B, 1, 6) * 0.3
    h0e = torch.randn(1, B, 20) * 0.3
    xd = torch.randn(B, T, 5)
    y0d = torch.randn(B, 1, 4) * 0.3
    h0d = torch.randn(1, B, 20) * 0.3
    mce = (torch.rand(B, T, enc.conv_dim) >= 0.5).float() * 2
    mge = (torch.rand(B, T, 20) >= 0.5).float() * 2
    mcd = (torch.rand(B, T, dec.conv_dim) >= 0.5).float() * 2
    mgd = (torch.rand(B, T, 20) >= 0.5).float() * 2
    g = {"B": B, "T": T, "lat": lat}
    for k, v in Pe.items():
        g["Pe/" + k] = v.numpy()
    for k, v in Pd.items():
        g["Pd/" + k] = v.numpy()
    g.update(x=x.numpy(), y0e=y0e.numpy(), h0e=h0e.numpy(), xd=xd.numpy(), y0d=y0d.numpy(), h0d=h0d.numpy(),
             mce=mce.numpy(), mge=mge.numpy(), mcd=mcd.numpy(), mgd=mgd.numpy())
    me.eval(); md.eval()
    with torch.no_grad():
        o, y, h = me(x, y0e, clamp_vae=True, lat_dim=lat)
        g.update(enc_eval_trj=o.numpy(), enc_eval_y=y.numpy(), enc_eval_h=h.numpy())
        o, y, h = me(x, y0e, h_in=h0e, clamp_vae=True, lat_dim=lat)
        g.update(enc_eval_hin_trj=o.numpy(), enc_eval_hin_y=y.numpy(), enc_eval_hin_h=h.numpy())
        o, y, h = md(xd, y0d, h_in=h0d)
        g.update(dec_eval_trj=o.numpy(), dec_eval_y=y.numpy(), dec_eval_h=h.numpy())
        # unbatched layout [T,C] (gru_vae.py:339-346,405-406,423-426)
        o, y, h = me(x[1], y0e[1:2], clamp_vae=True, lat_dim=lat)
        g.update(enc_unb_trj=o.numpy(), enc_unb_y=y.numpy(), enc_unb_h=h.numpy())
        o, y, h = md(xd[2], y0d[2:3])
        g.update(dec_unb_trj=o.numpy(), dec_unb_y=y.numpy(), dec_unb_h=h.numpy())
        # front-end alone
        g["enc_xconv"] = me.conv(me.scale_in(x.transpose(1, 2))).transpose(1, 2).numpy()
        g["dec_xconv"] = md.conv(xd.transpose(1, 2)).transpose(1, 2).numpy()
    # training mode with masks + gradients (loss = weighted sums so every output matters)
    me.train(); md.train()
    for net, m, spec, xin, y0, h0, mc, mg, kw in (("enc", me, enc, x, y0e, h0e, mce, mge, dict(clamp_vae=True, lat_dim=lat)),
                                                   ("dec", md, dec, xd, y0d, h0d, mcd, mgd, {})):
        xin = xin.clone().requires_grad_(True)
        y0 = y0.clone().requires_grad_(True)
        h0 = h0.clone().requires_grad_(True)
        m.zero_grad()
        o, y, h = ref_forward(m, xin, y0, h0, mc, mg, **kw)
        wt = torch.linspace(-1, 1, o.numel()).reshape(o.shape)
        loss = (o * wt).sum() + 0.7 * (y * y).sum() + 0.3 * h.sum()
        loss.backward()
        g.update({f"{net}_tr_trj": o.detach().numpy(), f"{net}_tr_y": y.detach().numpy(), f"{net}_tr_h": h.detach().numpy(),
                  f"{net}_tr_loss": loss.item(), f"{net}_tr_dx": xin.grad.numpy(), f"{net}_tr_dy0": y0.grad.numpy(),
                  f"{net}_tr_dh0": h0.grad.numpy()})
        for k, p in m.named_parameters():
            if p.grad is not None:
                g[f"{net}_tr_grad/{k}"] = p.grad.numpy()
    # losses (gru_vae.py:117-123, 525-528)
    crit = ref.TWFSEloss()
    lat_p = torch.randn(11, 6)
    a, b = torch.randn(11, 4), torch.randn(11, 4)
    s, mn, sd = crit(a, b, L2=False, GV=False)
    g.update(loss_lat_in=lat_p.numpy(), loss_kl=ref.loss_vae(lat_p, lat_dim=3).item(),
             mcd_a=a.numpy(), mcd_b=b.numpy(), mcd_sum=s.item(), mcd_mean=mn.item(), mcd_std=sd.item())
    np.savez_compressed(os.path.join(OUT, "tiny.npz"), **g)
    print("tiny.npz", len(g), "arrays")


# ----------------------------------------------------------------------------------------
def ref_cyc_step(me, md, enc, dec, x, cv, sc, tc, n_cyc, lat, stdim, y0e, y0d, eps, masks, flen_acc, select):
    """The trainer's first-chunk cyc graph + loss (train_*.py:1326-1338,1363-1410) driven through
    the reference's own GRU_RNN / loss_vae / TWFSEloss objects."""
    crit = ref.TWFSEloss()
    out = {k: [None] * n_cyc for k in ("lat_src", "trj_src_src", "trj_src_trg", "lat_src_trg", "trj_src_trg_src")}

    def samp(p, e):
        return p[:, :, :lat] + torch.exp(p[:, :, lat:] / 2) * e

    for i in range(n_cyc):
        mk = masks[i] if masks is not None else [(None, None)] * 5
        ein = x if i == 0 else torch.cat((x[:, :, :stdim], out["trj_src_trg_src"][i - 1]), 2)
        out["lat_src"][i], _, _ = ref_forward(me, ein, y0e, None, *mk[0], clamp_vae=True, lat_dim=lat)
        out["trj_src_src"][i], _, _ = ref_forward(md, torch.cat((sc, samp(out["lat_src"][i], eps[i][0])), 2), y0d, None, *mk[1])
        out["trj_src_trg"][i], _, _ = ref_forward(md, torch.cat((tc, samp(out["lat_src"][i], eps[i][1])), 2), y0d, None, *mk[2])
        out["lat_src_trg"][i], _, _ = ref_forward(me, torch.cat((cv, out["trj_src_trg"][i]), 2), y0e, None, *mk[3], clamp_vae=True, lat_dim=lat)
        out["trj_src_trg_src"][i], _, _ = ref_forward(md, torch.cat((sc, samp(out["lat_src_trg"][i], eps[i][2])), 2), y0d, None, *mk[4])
    total = None
    for i in range(n_cyc):
        v_ss = v_sts = v_kl = v_cv = None
        for k, j in enumerate(select):
            F_ = flen_acc[j]
            _, a, _ = crit(out["trj_src_src"][i][j, :F_], x[j, :F_, stdim:], L2=False, GV=False)
            _, b, _ = crit(out["trj_src_trg_src"][i][j, :F_], x[j, :F_, stdim:], L2=False, GV=False)
            c = ref.loss_vae(out["lat_src"][i][j, :F_], lat_dim=lat)
            d = ref.loss_vae(out["lat_src_trg"][i][j, :F_], lat_dim=lat)
            if k > 0:
                v_ss = torch.cat((v_ss, a.unsqueeze(0)))
                v_sts = torch.cat((v_sts, b.unsqueeze(0)))
                v_kl = torch.cat((v_kl, c.unsqueeze(0)))
                v_cv = torch.cat((v_kl, d.unsqueeze(0)))          # the reference's line 1393 verbatim behaviour
            else:
                v_ss, v_sts, v_kl, v_cv = a.unsqueeze(0), b.unsqueeze(0), c.unsqueeze(0), d.unsqueeze(0)
        c = v_ss.sum() + v_sts.sum() + v_kl.sum() + v_cv.sum()
        total = c if total is None else total + c
    return out, total


def golden_cfg0():
    """BASELINE.json configs[0]: hu128 ld16 cyc1, B=8 T=200, dropout masks injected, fwd+bwd."""
    lat, stdim, n_cyc, B, T = 16, 4, 1, 8, 200
    mean, std = orc.synth_stats(50)
    enc = orc.encoder_spec(54, lat, 128)
    dec = orc.decoder_spec(lat, 2, 50, 128)
    Pe = orc.init_params(enc, 101, mean=mean, scale=std)
    Pd = orc.init_params(dec, 102, mean=mean[stdim:], scale=std[stdim:])
    me, md = build_ref(enc, Pe).train(), build_ref(dec, Pd).train()
    x, cv, sc, tc = orc.synth_batch(B, T, 0)
    eps = orc.synth_noise(B, T, lat, n_cyc, 0)
    masks = orc.synth_masks(B, T, enc, dec, n_cyc, 0)
    y0e = torch.zeros(B, 1, 2 * lat)
    y0d = torch.tensor(((0 - mean[stdim:]) / std[stdim:]), dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    flen_acc = [T, T, 150, T, 77, T, T, 199]
    out, total = ref_cyc_step(me, md, enc, dec, x, cv, sc, tc, n_cyc, lat, stdim, y0e, y0d, eps, masks, flen_acc, list(range(B)))
    total.backward()
    g = {"loss": total.item(), "flen_acc": np.array(flen_acc), "pe_sum": orc.params_checksum(Pe), "pd_sum": orc.params_checksum(Pd)}
    for k, v in out.items():
        g[k] = sub(v[0], 9)
    for net, m in (("enc", me), ("dec", md)):
        for k, p in m.named_parameters():
            if p.grad is not None:
                gr = p.grad.detach().numpy()
                g[f"gnorm/{net}/{k}"] = np.sqrt((gr.astype(np.float64) ** 2).sum())
                g[f"gsamp/{net}/{k}"] = gr.reshape(-1)[:: max(1, gr.size // 64)][:64].copy()
    np.savez_compressed(os.path.join(OUT, "cfg0_cyc1.npz"), **g)
    print("cfg0_cyc1.npz loss", total.item())


def golden_flagship():
    """hu1024 ld32 (configs[1]/[2] shapes): eval conversion at T=800 unbatched (decode), a B=2 T=80
    training-mode cyc2 step (loss + sub-sampled outputs), 'trained-like' weights for drift stress."""
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    enc = orc.encoder_spec(54, lat, 1024)
    dec = orc.decoder_spec(lat, 2, 50, 1024)
    y0d1 = torch.tensor(((0 - mean[stdim:]) / std[stdim:]), dtype=torch.float32).reshape(1, 1, -1)
    g = {}
    for tag, gain, bstd in (("init", 1.0, 0.0), ("trained", 3.0, 0.05)):
        Pe = orc.init_params(enc, 201, gain=gain, bias_std=bstd, mean=mean, scale=std)
        Pd = orc.init_params(dec, 202, gain=gain, bias_std=bstd, mean=mean[stdim:], scale=std[stdim:])
        me, md = build_ref(enc, Pe).eval(), build_ref(dec, Pd).eval()
        g[f"{tag}/pe_sum"], g[f"{tag}/pd_sum"] = orc.params_checksum(Pe), orc.params_checksum(Pd)
        # stage-6 conversion, reference layout: unbatched [T,54] (decode_*.py:303-305,318)
        T = 800
        x, _, sc, tc = orc.synth_batch(1, T, 1)
        eps_mean = orc.synth_noise(1, T, lat, 1, 1)[0][0] / np.sqrt(300.0)   # mean of 300 N(0,1) draws
        with torch.no_grad():
            lat_src, _, _ = me(x[0], torch.zeros(1, 1, 2 * lat), clamp_vae=True, lat_dim=lat)
            lat_feat = lat_src[:, :lat] + torch.exp(lat_src[:, lat:] / 2) * eps_mean[0]
            cvm, _, _ = md(torch.cat((tc[0], lat_feat), 1), y0d1)
        g[f"{tag}/dec800_lat"] = lat_src.numpy()[::5].copy()
        g[f"{tag}/dec800_cvmcep"] = cvm.numpy()[::5].copy()
        # batched eval B=3, T=80, two consecutive chunks with carried state (TBPTT carry, eval numerics)
        B, T = 3, 80
        x, cv, sc, tc = orc.synth_batch(B, 2 * T, 2)
        with torch.no_grad():
            o1, y1, h1 = me(x[:, :T], torch.zeros(B, 1, 2 * lat), clamp_vae=True, lat_dim=lat)
            o2, y2, h2 = me(x[:, T:], y1, h_in=h1, clamp_vae=True, lat_dim=lat)
            zin = torch.cat((sc, torch.cat((o1, o2), 1)[:, :, :lat]), 2)
            d1, yd1, hd1 = md(zin[:, :T], y0d1.repeat(B, 1, 1))
            d2, yd2, hd2 = md(zin[:, T:], yd1, h_in=hd1)
        g[f"{tag}/carry_lat"] = sub(torch.cat((o1, o2), 1), 4)
        g[f"{tag}/carry_mcep"] = sub(torch.cat((d1, d2), 1), 4)
        g[f"{tag}/carry_h_enc"] = h2.numpy()[:, :, ::8].copy()
        g[f"{tag}/carry_h_dec"] = hd2.numpy()[:, :, ::8].copy()
    # training-mode cyc2 step, B=2 T=80 (configs[1] shapes, small B so the CPU run is seconds)
    Pe = orc.init_params(enc, 201, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, mean=mean[stdim:], scale=std[stdim:])
    me, md = build_ref(enc, Pe).train(), build_ref(dec, Pd).train()
    B, T, n_cyc = 2, 80, 2
    x, cv, sc, tc = orc.synth_batch(B, T, 3)
    eps = orc.synth_noise(B, T, lat, n_cyc, 3)
    masks = orc.synth_masks(B, T, enc, dec, n_cyc, 3)
    out, total = ref_cyc_step(me, md, enc, dec, x, cv, sc, tc, n_cyc, lat, stdim, torch.zeros(B, 1, 2 * lat),
                              y0d1.repeat(B, 1, 1), eps, masks, [T, 61], [0, 1])
    total.backward()
    g["cyc2/loss"] = total.item()
    for k, v in out.items():
        for i in range(n_cyc):
            g[f"cyc2/{k}/{i}"] = sub(v[i], 8)
    for net, m in (("enc", me), ("dec", md)):
        for k, p in m.named_parameters():
            if p.grad is not None:
                gr = p.grad.detach().numpy()
                g[f"cyc2/gnorm/{net}/{k}"] = np.sqrt((gr.astype(np.float64) ** 2).sum())
                g[f"cyc2/gsamp/{net}/{k}"] = gr.reshape(-1)[:: max(1, gr.size // 64)][:64].copy()
    np.savez_compressed(os.path.join(OUT, "flagship.npz"), **g)
    print("flagship.npz cyc2 loss", total.item())


def golden_spk4():
    """configs[3]: 4-speaker one-hot code (decoder in_dim = lat+4), eval forward B=2 T=40."""
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    dec = orc.decoder_spec(lat, 4, 50, 1024)
    Pd = orc.init_params(dec, 302, mean=mean[stdim:], scale=std[stdim:])
    md = build_ref(dec, Pd).eval()
    B, T = 2, 40
    g0 = torch.Generator().manual_seed(55)
    z = torch.randn(B, T, lat, generator=g0)
    code = torch.zeros(B, T, 4)
    code[0, :, 2] = 1
    code[1, :, 3] = 1
    y0 = torch.tensor(((0 - mean[stdim:]) / std[stdim:]), dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    with torch.no_grad():
        o, y, h = md(torch.cat((code, z), 2), y0)
    np.savez_compressed(os.path.join(OUT, "spk4.npz"), pd_sum=orc.params_checksum(Pd), trj=o.numpy(), y=y.numpy(),
                        h=h.numpy()[:, :, ::8].copy())
    print("spk4.npz")


def golden_spk4_cyc2():
    """configs[3]: 4-speaker many-to-many CycleVAE (decoder in_dim = lat + 4, one-hot codes), training-mode cyc2 step at
    B=2 T=80 with injected dropout masks: loss, sub-sampled outputs, gradient norms + samples from the reference's autograd."""
    lat, stdim, n_spk = 32, 4, 4
    mean, std = orc.synth_stats(50)
    enc = orc.encoder_spec(54, lat, 1024)
    dec = orc.decoder_spec(lat, n_spk, 50, 1024)
    Pe = orc.init_params(enc, 301, mean=mean, scale=std)
    Pd = orc.init_params(dec, 302, mean=mean[stdim:], scale=std[stdim:])
    me, md = build_ref(enc, Pe).train(), build_ref(dec, Pd).train()
    B, T, n_cyc = 2, 80, 2
    x, cv, sc, tc = orc.synth_batch(B, T, 13, n_spk=n_spk)
    sc[1], tc[1] = 0.0, 0.0
    sc[1, :, 2], tc[1, :, 0] = 1.0, 1.0          # utterance 1: speaker 2 -> speaker 0 (codes beyond the 2-speaker pattern)
    eps = orc.synth_noise(B, T, lat, n_cyc, 13)
    masks = orc.synth_masks(B, T, enc, dec, n_cyc, 13)
    y0d1 = torch.tensor(((0 - mean[stdim:]) / std[stdim:]), dtype=torch.float32).reshape(1, 1, -1)
    out, total = ref_cyc_step(me, md, enc, dec, x, cv, sc, tc, n_cyc, lat, stdim, torch.zeros(B, 1, 2 * lat),
                              y0d1.repeat(B, 1, 1), eps, masks, [T, 70], [0, 1])
    total.backward()
    g = {"loss": total.item(), "src_code": sc.numpy(), "trg_code": tc.numpy(), "pe_sum": orc.params_checksum(Pe), "pd_sum": orc.params_checksum(Pd)}
    for k, v in out.items():
        for i in range(n_cyc):
            g[f"{k}/{i}"] = sub(v[i], 8)
    for net, m in (("enc", me), ("dec", md)):
        for k, p in m.named_parameters():
            if p.grad is not None:
                gr = p.grad.detach().numpy()
                g[f"gnorm/{net}/{k}"] = np.sqrt((gr.astype(np.float64) ** 2).sum())
                g[f"gsamp/{net}/{k}"] = gr.reshape(-1)[:: max(1, gr.size // 64)][:64].copy()
    np.savez_compressed(os.path.join(OUT, "spk4_cyc2.npz"), **g)
    print("spk4_cyc2.npz loss", total.item())


DECODE512_ROWS = [0, 37, 100, 127, 128, 200, 255, 256, 300, 383, 384, 450, 500, 511, 64, 320]


def golden_decode512():
    """configs[2]: stage-6 batch conversion of 512 utterances x 800 frames.  The reference converts ONE utterance per call
    (decode_*.py:303-305,318, unbatched [T,54] layout), so the fixture is the reference run on 16 of the 512 synthetic
    utterances (4 in each 128-row slice of the batched launch), every 5th frame of the latent and of the converted mcep."""
    lat, stdim, B, T = 32, 4, 512, 800
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024)
    Pe = orc.init_params(enc, 201, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, mean=mean[stdim:], scale=std[stdim:])
    me, md = build_ref(enc, Pe).eval(), build_ref(dec, Pd).eval()
    y0d1 = torch.tensor(((0 - mean[stdim:]) / std[stdim:]), dtype=torch.float32).reshape(1, 1, -1)
    x, _, sc, tc = orc.synth_batch(B, T, 21)
    eps_mean = orc.synth_noise(B, T, lat, 1, 21)[0][0] / np.sqrt(300.0)
    lats, cvms = [], []
    with torch.no_grad():
        for r in DECODE512_ROWS:
            lat_src, _, _ = me(x[r], torch.zeros(1, 1, 2 * lat), clamp_vae=True, lat_dim=lat)
            lat_feat = lat_src[:, :lat] + torch.exp(lat_src[:, lat:] / 2) * eps_mean[r]
            cvm, _, _ = md(torch.cat((tc[r], lat_feat), 1), y0d1)
            lats.append(lat_src.numpy()[::5].copy())
            cvms.append(cvm.numpy()[::5].copy())
    np.savez_compressed(os.path.join(OUT, "decode512.npz"), rows=np.array(DECODE512_ROWS), lat=np.stack(lats), cvmcep=np.stack(cvms),
                        x_sum=float(x.double().abs().sum()), eps_sum=float(eps_mean.double().abs().sum()))
    print("decode512.npz", len(DECODE512_ROWS), "rows")


def golden_callers():
    """The reference's own caller code, cut out of the scripts' ASTs (tests/ref_callers.py), run with the reference's own
    classes on the CPU: three 80-frame chunks of one utterance batch through the trainer's frame-chunk branch
    (train_*.py:1293-1474, carried state, its loss assembly, its torch.optim.Adam) with a save_checkpoint round trip after
    the second chunk, and the conversion block of the decoder (decode_*.py:302-323).  do_prob = 0 and seeded noise, so
    the GPU run of the same code against the drop-in must reproduce the numbers."""
    import tempfile
    from tests import ref_callers as rc
    lat, stdim, n_cyc = 32, 4, 2
    mean, std = orc.synth_stats(50)
    enc = orc.NetSpec(54, 2 * lat, 1024, 3, 2, 0.0, True, False)
    dec = orc.NetSpec(lat + 2, 50, 1024, 3, 2, 0.0, False, True)
    Pe = orc.init_params(enc, 401, mean=mean, scale=std)
    Pd = orc.init_params(dec, 402, mean=mean[stdim:], scale=std[stdim:])
    me, md = build_ref(enc, Pe).train(), build_ref(dec, Pd).train()
    for m in (me, md):
        for k, p in m.named_parameters():
            p.requires_grad_(not k.startswith("scale_"))
    flens = [200, 170]
    x, cv, sc, tc = orc.synth_batch(2, 200, 31)
    y0d1 = torch.tensor(((0 - mean[stdim:]) / std[stdim:]), dtype=torch.float32).reshape(1, 1, -1)
    dev = torch.device("cpu")
    with tempfile.TemporaryDirectory() as td:
        r = rc.run_trainer_chunks(ref, me, md, dev, x=x, cv=cv, sc=sc, tc=tc, flens=flens, lat_dim=lat, n_cyc=n_cyc, y0_dec=y0d1,
                                  sample=rc.seeded_sampler(ref, lat, 5, dev, native=False), checkpoint_dir=td, checkpoint_after=2)
    g = {"losses": np.array(r["losses"]), "y_pp": r["y_pp"], "h_dec": r["h_dec"][:, :, ::8].copy(), "trj": r["trj"][:, ::4].copy(),
         "iter_count": r["iter_count"], "flens": np.array(flens)}
    # decoder block on fresh (untrained) copies, eval mode
    me, md = build_ref(enc, Pe).eval(), build_ref(dec, Pd).eval()
    T = 120
    f, _, _, _ = orc.synth_batch(2, T, 33)
    d = rc.run_decoder_block(me, md, dev, feat=f[0].numpy(), feat_trg=f[1].numpy(), lat_dim=lat, n_smpl=300, y0_dec=y0d1,
                             sample=rc.seeded_sampler(ref, lat, 6, dev, native=False))
    for k, v in d.items():
        g["dec_" + k] = v[::3].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "callers.npz"), **g)
    print("callers.npz losses", r["losses"])


def golden_chunks():
    """Bit-exact integer bookkeeping: exec the reference's own train_generator on fake loader batches."""
    src = open(os.path.join(REF, "src", "bin", "train_gru_cyclevae_gauss_batch.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "train_generator")
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train_generator", "exec"), ns)
    rng = np.random.default_rng(5)
    cases = []
    for case in range(6):
        n = int(rng.integers(1, 6))
        flens = rng.integers(30, 400, size=n)
        pad = 420
        spc = []
        for f in flens:
            if case % 2 == 0:
                idx = np.arange(f)
            else:                                   # speech frames = a strict subset (silence trimmed)
                lo, hi = int(rng.integers(0, f // 4)), int(f - rng.integers(0, f // 4))
                idx = np.array(sorted(rng.choice(np.arange(lo, hi), size=max(2, (hi - lo) * 3 // 4), replace=False)))
            spc.append(idx)
        bs = int(rng.choice([20, 80, 64]))
        batch = {
            "flen_src": torch.tensor(flens), "flen_spc_src": torch.tensor([len(s) for s in spc]),
            "flen_src_trg": torch.tensor(flens), "flen_spc_src_trg": torch.tensor([len(s) for s in spc]),
            "h_src": torch.zeros(n, pad, 2), "src_code": torch.zeros(n, pad, 2), "trg_code": torch.zeros(n, pad, 2),
            "cv_src": torch.zeros(n, pad, 2), "h_src_trg": torch.zeros(n, pad, 2),
            "spcidx_src": torch.stack([torch.tensor(np.pad(s, (0, pad - len(s)))) for s in spc]),
            "spcidx_src_trg": torch.stack([torch.tensor(np.pad(s, (0, pad - len(s)))) for s in spc]),
            "featfile_src": ["a"] * n, "featfile_src_trg": ["b"] * n,
        }
        gen = ns["train_generator"]([batch], torch.device("cpu"), batch_size=bs)
        rows = []
        while True:
            y = next(gen)
            if y[9] < 0:
                break
            s, e, ss, ee, sel, acc = y[5], y[6], y[7], y[8], y[19], y[20]
            rows.append((int(s), int(e), [int(v) for v in ss], [int(v) for v in ee], [int(v) for v in acc], [int(v) for v in sel]))
        cases.append({"flens": [int(f) for f in flens], "bs": bs, "spc": [[int(v) for v in s] for s in spc], "rows": rows})
    import json
    with open(os.path.join(OUT, "chunks.json"), "w") as f:
        json.dump(cases, f)
    print("chunks.json", sum(len(c["rows"]) for c in cases), "chunks")


def golden_init():
    """`.apply(initialize)` under a seed: the drop-in module must give identical numbers
    (same module tree => same RNG consumption order).  Store per-parameter checksums."""
    g = {}
    for tag, kw in (("enc", dict(in_dim=54, out_dim=32, hidden_units=128, scale_out_flag=False, do_prob=0.5)),
                    ("dec", dict(in_dim=18, out_dim=50, hidden_units=128, scale_in_flag=False, do_prob=0.5))):
        torch.manual_seed(1)
        m = ref.GRU_RNN(**kw)
        m.apply(ref.initialize)
        for k, v in m.state_dict().items():
            g[f"{tag}/{k}"] = np.array([v.double().sum().item(), v.double().abs().sum().item(), float(v.reshape(-1)[v.numel() // 2])])
        g[f"{tag}/keys"] = np.array(list(m.state_dict().keys()))
    np.savez_compressed(os.path.join(OUT, "init.npz"), **g)
    print("init.npz")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["tiny", "cfg0", "flagship", "spk4", "spk4_cyc2", "decode512", "callers", "chunks", "init"]
    for w in which:
        globals()["golden_" + w]()
