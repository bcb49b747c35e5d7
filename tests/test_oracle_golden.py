"""Pin the CPU oracle to fixtures produced by the unmodified reference (oracle/make_golden.py).

The reference ships no tests; these fixtures are outputs of
/root/reference/src/nets/gru_vae.py itself.  Tolerances: fp32 restatement vs fp32 reference,
different summation order only -> 2e-5 absolute on O(1) values; integers bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gru_vae_oracle as orc

TOL = 2e-5


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _params(g, prefix):
    return {k[len(prefix):]: torch.tensor(g[k]) for k in g.files if k.startswith(prefix)}


def _t(a):
    return torch.tensor(np.asarray(a))


@pytest.fixture(scope="module")
def tiny(golden_dir):
    g = _load(golden_dir, "tiny.npz")
    enc = orc.NetSpec(in_dim=7, out_dim=6, hidden_units=20, do_prob=0.5, scale_in=True, scale_out=False)
    dec = orc.NetSpec(in_dim=5, out_dim=4, hidden_units=20, do_prob=0.5, scale_in=False, scale_out=True)
    return g, enc, dec, _params(g, "Pe/"), _params(g, "Pd/")


def test_param_shapes_match_reference_state_dict(tiny):
    g, enc, dec, Pe, Pd = tiny
    assert [k for k, _ in enc.param_shapes()] == list(Pe.keys())
    assert [k for k, _ in dec.param_shapes()] == list(Pd.keys())
    for k, s in enc.param_shapes():
        assert tuple(Pe[k].shape) == s
    # flagship parameter counts
    assert orc.encoder_spec().n_trainable() == 5170160
    assert orc.decoder_spec().n_trainable() == 4401202
    assert orc.decoder_spec(n_spk=4).n_trainable() == 4469122


def test_frontend(tiny):
    g, enc, dec, Pe, Pd = tiny
    assert np.abs(orc.frontend(Pe, enc, _t(g["x"])).numpy() - g["enc_xconv"]).max() < TOL
    assert np.abs(orc.frontend(Pd, dec, _t(g["xd"])).numpy() - g["dec_xconv"]).max() < TOL


def test_forward_eval_and_layouts(tiny):
    g, enc, dec, Pe, Pd = tiny
    lat = int(g["lat"])
    o, y, h = orc.gru_rnn_forward(Pe, enc, _t(g["x"]), _t(g["y0e"]), clamp_vae=True, lat_dim=lat)
    for a, k in ((o, "enc_eval_trj"), (y, "enc_eval_y"), (h, "enc_eval_h")):
        assert a.shape == g[k].shape and np.abs(a.numpy() - g[k]).max() < TOL, k
    assert (g["enc_eval_trj"][:, :, lat:] == np.float32(orc.LOG_VAR_FLOOR)).any(), "clamp branch not exercised"
    o, y, h = orc.gru_rnn_forward(Pe, enc, _t(g["x"]), _t(g["y0e"]), _t(g["h0e"]), clamp_vae=True, lat_dim=lat)
    for a, k in ((o, "enc_eval_hin_trj"), (y, "enc_eval_hin_y"), (h, "enc_eval_hin_h")):
        assert np.abs(a.numpy() - g[k]).max() < TOL, k
    o, y, h = orc.gru_rnn_forward(Pd, dec, _t(g["xd"]), _t(g["y0d"]), _t(g["h0d"]))
    for a, k in ((o, "dec_eval_trj"), (y, "dec_eval_y"), (h, "dec_eval_h")):
        assert np.abs(a.numpy() - g[k]).max() < TOL, k
    # unbatched [T,C] layout
    o, y, h = orc.gru_rnn_forward(Pe, enc, _t(g["x"][1]), _t(g["y0e"][1:2]), clamp_vae=True, lat_dim=lat)
    for a, k in ((o, "enc_unb_trj"), (y, "enc_unb_y"), (h, "enc_unb_h")):
        assert a.shape == g[k].shape and np.abs(a.numpy() - g[k]).max() < TOL, k
    o, y, h = orc.gru_rnn_forward(Pd, dec, _t(g["xd"][2]), _t(g["y0d"][2:3]))
    for a, k in ((o, "dec_unb_trj"), (y, "dec_unb_y"), (h, "dec_unb_h")):
        assert a.shape == g[k].shape and np.abs(a.numpy() - g[k]).max() < TOL, k


@pytest.mark.parametrize("net", ["enc", "dec"])
def test_forward_backward_with_dropout_masks(tiny, net):
    g, enc, dec, Pe, Pd = tiny
    lat = int(g["lat"])
    spec, P0, x, y0, h0, mc, mg, kw = ((enc, Pe, "x", "y0e", "h0e", "mce", "mge", dict(clamp_vae=True, lat_dim=lat))
                                      if net == "enc" else (dec, Pd, "xd", "y0d", "h0d", "mcd", "mgd", {}))
    P = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
    x, y0, h0 = (_t(g[n]).requires_grad_(True) for n in (x, y0, h0))
    o, y, h = orc.gru_rnn_forward(P, spec, x, y0, h0, mask_conv=_t(g[mc]), mask_gru=_t(g[mg]), **kw)
    wt = torch.linspace(-1, 1, o.numel()).reshape(o.shape)
    loss = (o * wt).sum() + 0.7 * (y * y).sum() + 0.3 * h.sum()
    loss.backward()
    assert np.abs(o.detach().numpy() - g[f"{net}_tr_trj"]).max() < TOL
    assert abs(loss.item() - float(g[f"{net}_tr_loss"])) < 1e-3
    for a, k in ((x.grad, "dx"), (y0.grad, "dy0"), (h0.grad, "dh0")):
        ref = g[f"{net}_tr_{k}"]
        assert np.abs(a.numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), k
    for k, p in P.items():
        key = f"{net}_tr_grad/{k}"
        if key in g.files:
            ref = g[key]
            assert np.abs(p.grad.numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), k


def test_losses(tiny):
    g = tiny[0]
    assert abs(orc.loss_vae(_t(g["loss_lat_in"]), 3).item() - float(g["loss_kl"])) < 1e-5
    s, m, sd = orc.mcd_l1(_t(g["mcd_a"]), _t(g["mcd_b"]))
    assert abs(s.item() - float(g["mcd_sum"])) < 1e-3
    assert abs(m.item() - float(g["mcd_mean"])) < 1e-4
    assert abs(sd.item() - float(g["mcd_std"])) < 1e-4


def test_chunk_schedule_bit_exact(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "chunks.json")))
    assert cases
    for c in cases:
        rows = orc.chunk_schedule(c["flens"], c["bs"], c["spc"])
        assert len(rows) == len(c["rows"])
        for got, want in zip(rows, c["rows"]):
            s, e, ss, ee, acc, sel = got
            assert [s, e, ss, ee, acc, sel] == want


def _cyc_inputs(B, T, lat, n_cyc, seed, enc, dec):
    x, cv, sc, tc = orc.synth_batch(B, T, seed)
    return x, cv, sc, tc, orc.synth_noise(B, T, lat, n_cyc, seed), orc.synth_masks(B, T, enc, dec, n_cyc, seed)


def test_cfg0_cyc1_step(golden_dir):
    """BASELINE.json configs[0]: hu128 ld16 cyc1 on 8x200x50 synthetic mcep, fwd + bwd."""
    g = _load(golden_dir, "cfg0_cyc1.npz")
    lat, stdim, B, T = 16, 4, 8, 200
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 128), orc.decoder_spec(lat, 2, 50, 128)
    Pe = orc.init_params(enc, 101, mean=mean, scale=std)
    Pd = orc.init_params(dec, 102, mean=mean[stdim:], scale=std[stdim:])
    assert orc.params_checksum(Pe) == pytest.approx(float(g["pe_sum"]), rel=1e-12)
    assert orc.params_checksum(Pd) == pytest.approx(float(g["pd_sum"]), rel=1e-12)
    for P in (Pe, Pd):
        for k, v in P.items():
            if not k.startswith("scale_"):
                v.requires_grad_(True)
    x, cv, sc, tc, eps, masks = _cyc_inputs(B, T, lat, 1, 0, enc, dec)
    y0e = torch.zeros(B, 1, 2 * lat)
    y0d = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    out, _ = orc.cyc_forward(Pe, Pd, enc, dec, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=1, lat_dim=lat, stdim=stdim,
                             y0_enc=y0e, y0_dec=y0d, eps=eps, masks=masks)
    total, _ = orc.cyc_loss(out, x, n_cyc=1, lat_dim=lat, stdim=stdim, flen_acc=g["flen_acc"].tolist(),
                            select_utt_idx=list(range(B)))
    total.backward()
    assert total.item() == pytest.approx(float(g["loss"]), rel=2e-6)
    for k in out:
        assert np.abs(out[k][0].detach().numpy()[:, ::9] - g[k]).max() < 5e-5, k
    for net, P in (("enc", Pe), ("dec", Pd)):
        for k, v in P.items():
            if v.grad is None:
                continue
            gr = v.grad.numpy()
            assert np.sqrt((gr.astype(np.float64) ** 2).sum()) == pytest.approx(float(g[f"gnorm/{net}/{k}"]), rel=1e-4), k
            samp = gr.reshape(-1)[:: max(1, gr.size // 64)][:64]
            ref = g[f"gsamp/{net}/{k}"]
            assert np.abs(samp - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), k


@pytest.mark.parametrize("tag,gain,bstd", [("init", 1.0, 0.0), ("trained", 3.0, 0.05)])
def test_flagship_decode_and_carry(golden_dir, tag, gain, bstd):
    """hu1024 ld32: stage-6 conversion of an 800-frame utterance (unbatched layout) and two chunks
    with carried (y, h) state; both with reference-init and 'trained-like' weights."""
    g = _load(golden_dir, "flagship.npz")
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024)
    Pe = orc.init_params(enc, 201, gain=gain, bias_std=bstd, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, gain=gain, bias_std=bstd, mean=mean[stdim:], scale=std[stdim:])
    assert orc.params_checksum(Pe) == pytest.approx(float(g[f"{tag}/pe_sum"]), rel=1e-12)
    y0d1 = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1)
    T = 800
    x, _, sc, tc = orc.synth_batch(1, T, 1)
    eps_mean = orc.synth_noise(1, T, lat, 1, 1)[0][0] / np.sqrt(300.0)
    with torch.no_grad():
        lat_src, _, _ = orc.gru_rnn_forward(Pe, enc, x[0], torch.zeros(1, 1, 2 * lat), clamp_vae=True, lat_dim=lat)
        cvm = orc.convert(Pe, Pd, enc, dec, x[0], tc[0], lat_dim=lat, y0_enc=torch.zeros(1, 1, 2 * lat), y0_dec=y0d1,
                          eps_mean=eps_mean[0])
    assert np.abs(lat_src.numpy()[::5] - g[f"{tag}/dec800_lat"]).max() < 1e-4
    assert np.abs(cvm.numpy()[::5] - g[f"{tag}/dec800_cvmcep"]).max() < 1e-4
    B, T = 3, 80
    x, cv, sc, tc = orc.synth_batch(B, 2 * T, 2)
    with torch.no_grad():
        o1, y1, h1 = orc.gru_rnn_forward(Pe, enc, x[:, :T], torch.zeros(B, 1, 2 * lat), clamp_vae=True, lat_dim=lat)
        o2, y2, h2 = orc.gru_rnn_forward(Pe, enc, x[:, T:], y1, h1, clamp_vae=True, lat_dim=lat)
        zin = torch.cat((sc, torch.cat((o1, o2), 1)[:, :, :lat]), 2)
        d1, yd1, hd1 = orc.gru_rnn_forward(Pd, dec, zin[:, :T], y0d1.repeat(B, 1, 1))
        d2, yd2, hd2 = orc.gru_rnn_forward(Pd, dec, zin[:, T:], yd1, hd1)
    assert np.abs(torch.cat((o1, o2), 1).numpy()[:, ::4] - g[f"{tag}/carry_lat"]).max() < 1e-4
    assert np.abs(torch.cat((d1, d2), 1).numpy()[:, ::4] - g[f"{tag}/carry_mcep"]).max() < 1e-4
    assert np.abs(h2.numpy()[:, :, ::8] - g[f"{tag}/carry_h_enc"]).max() < 1e-4
    assert np.abs(hd2.numpy()[:, :, ::8] - g[f"{tag}/carry_h_dec"]).max() < 1e-4


def test_flagship_cyc2_step(golden_dir):
    """configs[1] shapes (hu1024 ld32 cyc2, T=80) at B=2 with ragged flen_acc and the KL-cv quirk."""
    g = _load(golden_dir, "flagship.npz")
    lat, stdim, B, T, n_cyc = 32, 4, 2, 80, 2
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024)
    Pe = orc.init_params(enc, 201, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, mean=mean[stdim:], scale=std[stdim:])
    for P in (Pe, Pd):
        for k, v in P.items():
            if not k.startswith("scale_"):
                v.requires_grad_(True)
    x, cv, sc, tc, eps, masks = _cyc_inputs(B, T, lat, n_cyc, 3, enc, dec)
    y0d = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    out, _ = orc.cyc_forward(Pe, Pd, enc, dec, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=n_cyc, lat_dim=lat,
                             stdim=stdim, y0_enc=torch.zeros(B, 1, 2 * lat), y0_dec=y0d, eps=eps, masks=masks)
    total, _ = orc.cyc_loss(out, x, n_cyc=n_cyc, lat_dim=lat, stdim=stdim, flen_acc=[T, 61], select_utt_idx=[0, 1])
    total.backward()
    assert total.item() == pytest.approx(float(g["cyc2/loss"]), rel=2e-6)
    for k in out:
        for i in range(n_cyc):
            assert np.abs(out[k][i].detach().numpy()[:, ::8] - g[f"cyc2/{k}/{i}"]).max() < 1e-4, (k, i)
    for net, P in (("enc", Pe), ("dec", Pd)):
        for k, v in P.items():
            if v.grad is None:
                continue
            gr = v.grad.numpy()
            assert np.sqrt((gr.astype(np.float64) ** 2).sum()) == pytest.approx(float(g[f"cyc2/gnorm/{net}/{k}"]), rel=2e-4), k


def test_spk4(golden_dir):
    g = _load(golden_dir, "spk4.npz")
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    dec = orc.decoder_spec(lat, 4, 50, 1024)
    Pd = orc.init_params(dec, 302, mean=mean[stdim:], scale=std[stdim:])
    assert orc.params_checksum(Pd) == pytest.approx(float(g["pd_sum"]), rel=1e-12)
    B, T = 2, 40
    z = torch.randn(B, T, lat, generator=torch.Generator().manual_seed(55))
    code = torch.zeros(B, T, 4)
    code[0, :, 2] = 1
    code[1, :, 3] = 1
    y0 = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    with torch.no_grad():
        o, y, h = orc.gru_rnn_forward(Pd, dec, torch.cat((code, z), 2), y0)
    assert np.abs(o.numpy() - g["trj"]).max() < 1e-4
    assert np.abs(y.numpy() - g["y"]).max() < 1e-4
    assert np.abs(h.numpy()[:, :, ::8] - g["h"]).max() < 1e-4


def test_spk4_cyc2_step(golden_dir):
    """configs[3]: 4-speaker one-hot codes (decoder in_dim 36), training-mode cyc2 step at B=2 T=80 vs the reference."""
    g = _load(golden_dir, "spk4_cyc2.npz")
    lat, stdim, B, T, n_cyc, n_spk = 32, 4, 2, 80, 2, 4
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, n_spk, 50, 1024)
    Pe = orc.init_params(enc, 301, mean=mean, scale=std)
    Pd = orc.init_params(dec, 302, mean=mean[stdim:], scale=std[stdim:])
    assert orc.params_checksum(Pd) == pytest.approx(float(g["pd_sum"]), rel=1e-12)
    for P in (Pe, Pd):
        for k, v in P.items():
            if not k.startswith("scale_"):
                v.requires_grad_(True)
    x, cv, _, _ = orc.synth_batch(B, T, 13, n_spk=n_spk)
    sc, tc = _t(g["src_code"]), _t(g["trg_code"])
    eps = orc.synth_noise(B, T, lat, n_cyc, 13)
    masks = orc.synth_masks(B, T, enc, dec, n_cyc, 13)
    y0d = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    out, _ = orc.cyc_forward(Pe, Pd, enc, dec, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=n_cyc, lat_dim=lat,
                             stdim=stdim, y0_enc=torch.zeros(B, 1, 2 * lat), y0_dec=y0d, eps=eps, masks=masks)
    total, _ = orc.cyc_loss(out, x, n_cyc=n_cyc, lat_dim=lat, stdim=stdim, flen_acc=[T, 70], select_utt_idx=[0, 1])
    total.backward()
    assert total.item() == pytest.approx(float(g["loss"]), rel=2e-6)
    for k in out:
        for i in range(n_cyc):
            assert np.abs(out[k][i].detach().numpy()[:, ::8] - g[f"{k}/{i}"]).max() < 1e-4, (k, i)
    for net, P in (("enc", Pe), ("dec", Pd)):
        for k, v in P.items():
            if v.grad is not None:
                gr = v.grad.numpy()
                assert np.sqrt((gr.astype(np.float64) ** 2).sum()) == pytest.approx(float(g[f"gnorm/{net}/{k}"]), rel=2e-4), k


def test_decode512_rows(golden_dir):
    """configs[2]: two of the 16 reference-converted utterances of the 512 x 800 synthetic batch (the GPU test checks all 16)."""
    g = _load(golden_dir, "decode512.npz")
    lat, stdim, B, T = 32, 4, 512, 800
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024)
    Pe = orc.init_params(enc, 201, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, mean=mean[stdim:], scale=std[stdim:])
    x, _, sc, tc = orc.synth_batch(B, T, 21)
    eps_mean = orc.synth_noise(B, T, lat, 1, 21)[0][0] / np.sqrt(300.0)
    assert float(x.double().abs().sum()) == pytest.approx(float(g["x_sum"]), rel=1e-12)
    y0d = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1)
    rows = g["rows"].tolist()
    for i in (1, 12):
        r = rows[i]
        with torch.no_grad():
            lat_src, _, _ = orc.gru_rnn_forward(Pe, enc, x[r], torch.zeros(1, 1, 2 * lat), clamp_vae=True, lat_dim=lat)
            cvm = orc.convert(Pe, Pd, enc, dec, x[r], tc[r], lat_dim=lat, y0_enc=torch.zeros(1, 1, 2 * lat), y0_dec=y0d, eps_mean=eps_mean[r])
        assert np.abs(lat_src.numpy()[::5] - g["lat"][i]).max() < 1e-4
        assert np.abs(cvm.numpy()[::5] - g["cvmcep"][i]).max() < 1e-4


def test_oracle_gradcheck_fp64():
    """BPTT of the restatement is exact in fp64 (finite differences), incl. masks, h_in, y_in."""
    spec = orc.NetSpec(in_dim=3, out_dim=4, hidden_units=5, do_prob=0.5, scale_in=True, scale_out=False)
    P = orc.init_params(spec, 3, gain=2.0, bias_std=0.2, dtype=torch.float64)
    g = torch.Generator().manual_seed(0)
    B, T = 2, 6
    x = torch.randn(B, T, 3, generator=g, dtype=torch.float64, requires_grad=True)
    y0 = torch.randn(B, 1, 4, generator=g, dtype=torch.float64, requires_grad=True)
    h0 = torch.randn(1, B, 5, generator=g, dtype=torch.float64, requires_grad=True)
    mc = (torch.rand(B, T, spec.conv_dim, generator=g) > 0.5).double() * 2
    mg = (torch.rand(B, T, 5, generator=g) > 0.5).double() * 2
    w = P["gru.weight_hh_l0"].requires_grad_(True)

    def f(x, y0, h0, w):
        Q = dict(P)
        Q["gru.weight_hh_l0"] = w
        o, y, h = orc.gru_rnn_forward(Q, spec, x, y0, h0, mask_conv=mc, mask_gru=mg, clamp_vae=True, lat_dim=2)
        return o.sum() + (y * y).sum() + h.sum()

    assert torch.autograd.gradcheck(f, (x, y0, h0, w), eps=1e-6, atol=1e-6)


def test_dtw_oracle_properties():
    """oracle/dtw_oracle.py (the checker of the device metrics; dtw_c is not in the reference tree): identity alignment,
    recovery of a pure time-stretch, MCD against its closed form."""
    from oracle import dtw_oracle as dto
    rng = np.random.default_rng(1)
    x = np.cumsum(rng.normal(size=(60, 8)), axis=0)
    al, path, mean, steps, cost = dto.dtw_org_to_trg(x, x)
    assert (path == np.arange(60)).all() and mean == 0.0 and steps == 60 and cost == 0.0
    stretched = np.repeat(x, 2, axis=0)
    al, path, mean, steps, cost = dto.dtw_org_to_trg(stretched, x)
    assert mean == 0.0 and (path == 2 * np.arange(60) + 1).all()      # the LAST source frame paired with each target frame
    y = x + 0.5
    m, s = dto.calc_mcd(x, y)
    assert m == pytest.approx((10 / np.log(10)) * np.sqrt(2 * 8 * 0.25)) and s == pytest.approx(0.0, abs=1e-9)
