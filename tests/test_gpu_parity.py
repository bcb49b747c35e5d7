"""Parity of the CUDA path (through the drop-in module -> ctypes -> C ABI) against the CPU oracle and
the reference-generated golden fixtures.  Everything here needs a B200: `pytest -m gpu`.

Tolerances (fp32 arithmetic, different summation order than the reference):
  * forward outputs (mcep / latent / states): max-abs <= 1e-4  (BASELINE.json north_star)
  * gradients: 1e-4 relative to max(1, max|ref|) on tiny nets, 2e-4 relative on gradient norms at hu1024
  * integers (chunk schedule) bit-exact -- covered on CPU in test_host_cpu.py
"""
import os

import numpy as np
import pytest
import torch

from oracle import gru_vae_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-4
PATH_FP32, PATH_TC, PATH_TC_FOLDED = 0, 1, 2   # include/cyclevae_b200.h CVB_PATH_*


def _paths():
    """(forward, backward) recurrence kernels of the last calls: a silent fall to the fp32-FMA kernels must fail a test."""
    from cyclevae_vc_b200._lib import lib
    return lib.cvb_last_recurrence_path(0), lib.cvb_last_recurrence_path(1)


@pytest.fixture(scope="module")
def cvb():
    import cyclevae_vc_b200 as pkg  # raises if libcyclevae_b200.so is missing: no silent fallback
    assert torch.cuda.is_available()
    return pkg


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _params(g, prefix):
    return {k[len(prefix):]: torch.tensor(g[k]) for k in g.files if k.startswith(prefix)}


def _module(cvb, spec: orc.NetSpec, P):
    m = cvb.GRU_RNN(in_dim=spec.in_dim, out_dim=spec.out_dim, hidden_units=spec.hidden_units, kernel_size=spec.kernel_size,
                    dilation_size=spec.dilation_size, do_prob=spec.do_prob, scale_in_flag=spec.scale_in,
                    scale_out_flag=spec.scale_out)
    res = m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return m.cuda()


def _c(a):
    return torch.tensor(np.asarray(a)).cuda()


def _maxabs(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


@pytest.fixture(scope="module")
def tiny(golden_dir, cvb):
    g = _load(golden_dir, "tiny.npz")
    enc = orc.NetSpec(in_dim=7, out_dim=6, hidden_units=20, do_prob=0.5, scale_in=True, scale_out=False)
    dec = orc.NetSpec(in_dim=5, out_dim=4, hidden_units=20, do_prob=0.5, scale_in=False, scale_out=True)
    Pe, Pd = _params(g, "Pe/"), _params(g, "Pd/")
    return g, enc, dec, Pe, Pd, _module(cvb, enc, Pe), _module(cvb, dec, Pd)


def test_gemm_matches_torch(cvb):
    from cyclevae_vc_b200._lib import check, lib, ptr
    torch.manual_seed(0)
    for (M, N, K, ta, tb) in [(37, 29, 53, 0, 1), (64, 128, 256, 1, 0), (5, 7, 1000, 0, 0), (130, 66, 31, 1, 1)]:
        A = torch.randn((K, M) if ta else (M, K), device="cuda")
        Bm = torch.randn((N, K) if tb else (K, N), device="cuda")
        Cm = torch.randn(M, N, device="cuda")
        ref = 0.5 * (A.t() if ta else A).double() @ (Bm.t() if tb else Bm).double() + 2.0 * Cm.double()
        check(lib.cvb_gemm(ta, tb, M, N, K, 0.5, ptr(A), A.shape[1], ptr(Bm), Bm.shape[1], 2.0, ptr(Cm), N,
                           torch.cuda.current_stream().cuda_stream))
        assert _maxabs(Cm.double(), ref) < 1e-3 * max(1.0, ref.abs().max().item())


def test_tiny_forward_eval_layouts(tiny):
    g, enc, dec, Pe, Pd, me, md = tiny
    lat = int(g["lat"])
    me.eval(); md.eval()
    with torch.no_grad():
        o, y, h = me(_c(g["x"]), _c(g["y0e"]), clamp_vae=True, lat_dim=lat)
        for a, k in ((o, "enc_eval_trj"), (y, "enc_eval_y"), (h, "enc_eval_h")):
            assert _maxabs(a, g[k]) < TOL, k
        assert (o.cpu().numpy()[:, :, lat:] == np.float32(orc.LOG_VAR_FLOOR)).any(), "clamp branch not exercised"
        o, y, h = me(_c(g["x"]), _c(g["y0e"]), h_in=_c(g["h0e"]), clamp_vae=True, lat_dim=lat)
        for a, k in ((o, "enc_eval_hin_trj"), (y, "enc_eval_hin_y"), (h, "enc_eval_hin_h")):
            assert _maxabs(a, g[k]) < TOL, k
        o, y, h = md(_c(g["xd"]), _c(g["y0d"]), h_in=_c(g["h0d"]))
        for a, k in ((o, "dec_eval_trj"), (y, "dec_eval_y"), (h, "dec_eval_h")):
            assert _maxabs(a, g[k]) < TOL, k
        # unbatched [T,C] layout (gru_vae.py:339-346,405-406,423-426)
        o, y, h = me(_c(g["x"][1]), _c(g["y0e"][1:2]), clamp_vae=True, lat_dim=lat)
        for a, k in ((o, "enc_unb_trj"), (y, "enc_unb_y"), (h, "enc_unb_h")):
            assert _maxabs(a, g[k]) < TOL, k
        o, y, h = md(_c(g["xd"][2]), _c(g["y0d"][2:3]))
        for a, k in ((o, "dec_unb_trj"), (y, "dec_unb_y"), (h, "dec_unb_h")):
            assert _maxabs(a, g[k]) < TOL, k


def test_tiny_frontend(tiny):
    import ctypes as C
    from cyclevae_vc_b200._lib import check, lib, ptr
    g, enc, dec, Pe, Pd, me, md = tiny
    for m, xk, ok in ((me, "x", "enc_xconv"), (md, "xd", "dec_xconv")):
        x = _c(g[xk])
        B, T, _ = x.shape
        net = m._net_struct(m._param_list())
        ws = torch.empty(lib.cvb_frontend_ws_floats(C.byref(net), B, T), device="cuda")
        Cd = m.in_dim * m.receptive_field
        xc = torch.empty(T, B, Cd, device="cuda")
        check(lib.cvb_frontend_fwd(C.byref(net), B, T, ptr(x), None, ptr(ws), ptr(xc), torch.cuda.current_stream().cuda_stream))
        assert _maxabs(xc.transpose(0, 1), g[ok]) < TOL


@pytest.mark.parametrize("net", ["enc", "dec"])
def test_tiny_forward_backward_with_dropout_masks(tiny, net):
    g, enc, dec, Pe, Pd, me, md = tiny
    lat = int(g["lat"])
    m, xk, y0k, h0k, mck, mgk, kw = ((me, "x", "y0e", "h0e", "mce", "mge", dict(clamp_vae=True, lat_dim=lat)) if net == "enc"
                                     else (md, "xd", "y0d", "h0d", "mcd", "mgd", {}))
    m.train()
    m.zero_grad()
    x, y0, h0 = (_c(g[n]).requires_grad_(True) for n in (xk, y0k, h0k))
    m.inject_dropout_masks(_c(g[mck]), _c(g[mgk]))
    o, y, h = m(x, y0, h_in=h0, do=True, **kw)
    wt = torch.linspace(-1, 1, o.numel()).reshape(o.shape).cuda()
    loss = (o * wt).sum() + 0.7 * (y * y).sum() + 0.3 * h.sum()
    loss.backward()
    assert _maxabs(o, g[f"{net}_tr_trj"]) < TOL
    assert _maxabs(y, g[f"{net}_tr_y"]) < TOL
    assert _maxabs(h, g[f"{net}_tr_h"]) < TOL
    assert abs(loss.item() - float(g[f"{net}_tr_loss"])) < 1e-3
    for a, k in ((x.grad, "dx"), (y0.grad, "dy0"), (h0.grad, "dh0")):
        ref = g[f"{net}_tr_{k}"]
        assert _maxabs(a, ref) < 1e-4 * max(1.0, np.abs(ref).max()), k
    n_checked = 0
    for k, p in m.named_parameters():
        key = f"{net}_tr_grad/{k}"
        if key in g.files:
            ref = g[key]
            assert p.grad is not None, k
            assert _maxabs(p.grad, ref) < 1e-4 * max(1.0, np.abs(ref).max()), k
            n_checked += 1
    assert n_checked >= 10


def test_losses_and_sampling(tiny, cvb):
    g = tiny[0]
    kl = cvb.loss_vae(_c(g["loss_lat_in"]), lat_dim=3)
    assert abs(kl.item() - float(g["loss_kl"])) < 1e-5
    crit = cvb.TWFSEloss()
    s, m, sd = crit(_c(g["mcd_a"]), _c(g["mcd_b"]), L2=False, GV=False)
    assert abs(s.item() - float(g["mcd_sum"])) < 1e-3
    assert abs(m.item() - float(g["mcd_mean"])) < 1e-4
    assert abs(sd.item() - float(g["mcd_std"])) < 1e-4
    # gradients of both losses vs autograd of the oracle, ragged lengths, batched entry points
    gen = torch.Generator().manual_seed(3)
    B, T, lat, D = 4, 23, 5, 9
    latp = torch.randn(B, T, 2 * lat, generator=gen)
    a, b = torch.randn(B, T, D, generator=gen), torch.randn(B, T, D + 4, generator=gen)
    flens = [23, 0, 7, 1]
    lc, ac = latp.clone().requires_grad_(True), a.clone().requires_grad_(True)
    ref = sum(orc.loss_vae(lc[j, :f], lat) * (j + 1) for j, f in enumerate(flens) if f > 0) + \
        sum(orc.mcd_l1(ac[j, :f], b[j, :f, 4:])[1] * (j + 2) + 0.1 * orc.mcd_l1(ac[j, :f], b[j, :f, 4:])[0]
            for j, f in enumerate(flens) if f > 0)
    ref.backward()
    lg, ag = latp.cuda().requires_grad_(True), a.cuda().requires_grad_(True)
    fl = torch.tensor(flens, dtype=torch.int32, device="cuda")
    wk = torch.arange(1, B + 1, device="cuda", dtype=torch.float32)
    s, m, sd = cvb.mcd_l1_per_utt(ag, b.cuda(), fl, 0, 4)
    out = (cvb.kl_per_utt(lg, fl, lat) * wk).sum() + (m * (wk + 1)).sum() + 0.1 * s.sum()
    out.backward()
    assert abs(out.item() - ref.item()) < 1e-3 * max(1.0, abs(ref.item()))
    assert _maxabs(lg.grad, lc.grad) < 1e-5 * max(1.0, lc.grad.abs().max().item())
    assert _maxabs(ag.grad, ac.grad) < 1e-5 * max(1.0, ac.grad.abs().max().item())
    # reparameterise + concat, given noise: forward and backward
    eps = torch.randn(B, T, lat, generator=gen)
    code = torch.randn(B, T, 2, generator=gen)
    lc = latp.clone().requires_grad_(True)
    zr = torch.cat((code, orc.sampling_vae_batch(lc, eps, lat)), 2)
    wz = torch.linspace(-1, 1, zr.numel()).reshape(zr.shape)
    (zr * wz).sum().backward()
    lg = latp.cuda().requires_grad_(True)
    z = cvb.reparam_concat(lg, code.cuda(), eps.cuda(), lat)
    (z * wz.cuda()).sum().backward()
    assert _maxabs(z, zr) < 1e-5
    assert _maxabs(lg.grad, lc.grad) < 1e-5 * max(1.0, lc.grad.abs().max().item())
    assert _maxabs(cvb.sampling_vae_batch(latp.cuda(), lat_dim=lat, eps=eps.cuda()), orc.sampling_vae_batch(latp, eps, lat)) < 1e-5


def test_device_rng_statistics(cvb):
    """In-kernel Philox draws: dropout keep-rate and N(0,1) moments; seeded reproducibility."""
    from cyclevae_vc_b200.gru_vae import draw_dropout_masks
    torch.manual_seed(11)
    mc, mg = draw_dropout_masks(16, 40, 486, 1024, 0.5, torch.device("cuda"))
    for m in (mc, mg):
        vals = torch.unique(m)
        assert vals.tolist() == [0.0, 2.0]
        assert abs((m > 0).float().mean().item() - 0.5) < 5e-3
    latp = torch.zeros(64, 100, 64, device="cuda")
    z = cvb.sampling_vae_batch(latp, lat_dim=32)       # mu=0, log-var=0 -> z = eps
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1.0) < 0.01
    assert abs((z ** 4).mean().item() - 3.0) < 0.1
    torch.manual_seed(11)
    mc2, _ = draw_dropout_masks(16, 40, 486, 1024, 0.5, torch.device("cuda"))
    assert torch.equal(mc, mc2)


def test_reference_noise_stream_is_the_references_cpu_generator(cvb):
    """Opt-in CVB_REFERENCE_NOISE / gru_vae.REFERENCE_NOISE_STREAM: sampling_vae_batch draws eps as gru_vae.py:91 does
    (torch.randn on the CPU generator, copied to the device), so a seeded run sees the reference's own noise and leaves
    the CPU generator in the reference's state; off by default (device-side Philox, nothing taken per element)."""
    from cyclevae_vc_b200 import gru_vae as gv
    g = torch.Generator().manual_seed(2)
    latp = torch.randn(5, 17, 64, generator=g)
    assert not gv.reference_noise_stream()
    gv.REFERENCE_NOISE_STREAM = True
    try:
        torch.manual_seed(123)
        z = cvb.sampling_vae_batch(latp.cuda(), lat_dim=32)
        after = torch.rand(1)
    finally:
        gv.REFERENCE_NOISE_STREAM = None
    torch.manual_seed(123)
    eps = torch.randn(5, 17, 32)                      # the reference's draw
    after_ref = torch.rand(1)
    assert _maxabs(z, orc.sampling_vae_batch(latp, eps, 32)) < 1e-5
    assert torch.equal(after, after_ref)              # same position of the CPU generator afterwards


def test_adam_matches_torch(cvb):
    from cyclevae_vc_b200.cycle import FlatAdam
    torch.manual_seed(5)
    ps = [torch.nn.Parameter(torch.randn(37, 11, device="cuda")), torch.nn.Parameter(torch.randn(501, device="cuda"))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref = torch.optim.Adam(qs, lr=1e-2)
    opt = FlatAdam(ps, lr=1e-2)
    for it in range(5):
        gs = [torch.randn_like(p) for p in ps]
        opt.zero_grad()
        for p, q, g_ in zip(ps, qs, gs):
            p.grad.add_(g_)
            q.grad = g_.clone()
        opt.step()
        ref.step()
    for p, q in zip(ps, qs):
        assert _maxabs(p, q) < 1e-5


def _cyc_setup(cvb, hu, lat, B, T, n_cyc, seed, pe_seed, pd_seed, n_spk=2):
    stdim = 4
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, hu), orc.decoder_spec(lat, n_spk, 50, hu)
    Pe = orc.init_params(enc, pe_seed, mean=mean, scale=std)
    Pd = orc.init_params(dec, pd_seed, mean=mean[stdim:], scale=std[stdim:])
    me, md = _module(cvb, enc, Pe).train(), _module(cvb, dec, Pd).train()
    for m in (me, md):
        for k, p in m.named_parameters():
            p.requires_grad_(not k.startswith("scale_"))
    x, cv, sc, tc = orc.synth_batch(B, T, seed)
    eps = orc.synth_noise(B, T, lat, n_cyc, seed)
    masks = orc.synth_masks(B, T, enc, dec, n_cyc, seed)
    y0e = torch.zeros(B, 1, 2 * lat)
    y0d = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    return enc, dec, me, md, x, cv, sc, tc, eps, masks, y0e, y0d


def _run_cyc(cvb, me, md, x, cv, sc, tc, eps, masks, y0e, y0d, n_cyc, lat, flen_acc, select):
    from cyclevae_vc_b200 import cycle
    cu = lambda t: t.cuda()
    out, st = cycle.cyc_forward(me, md, x=cu(x), cv=cu(cv), src_code=cu(sc), trg_code=cu(tc), n_cyc=n_cyc, lat_dim=lat,
                                stdim=4, y0_enc=cu(y0e), y0_dec=cu(y0d), do=True,
                                eps=[[cu(e) for e in ec] for ec in eps],
                                masks=[[(cu(a), cu(b)) for a, b in mc] for mc in masks])
    total, parts = cycle.cyc_loss(out, cu(x), n_cyc=n_cyc, lat_dim=lat, stdim=4, flen_acc=flen_acc, select_utt_idx=select)
    return out, total


def test_cfg0_cyc1_step_vs_reference_golden(golden_dir, cvb):
    """BASELINE.json configs[0] (hu128 ld16 cyc1, 8x200x50) -- on the GPU path, vs fixtures produced by the
    unmodified reference: outputs, loss, gradient norms and samples."""
    g = _load(golden_dir, "cfg0_cyc1.npz")
    lat, B, T = 16, 8, 200
    enc, dec, me, md, x, cv, sc, tc, eps, masks, y0e, y0d = _cyc_setup(cvb, 128, lat, B, T, 1, 0, 101, 102)
    out, total = _run_cyc(cvb, me, md, x, cv, sc, tc, eps, masks, y0e, y0d, 1, lat, g["flen_acc"].tolist(), list(range(B)))
    total.backward()
    assert total.item() == pytest.approx(float(g["loss"]), rel=5e-6)
    for k in out:
        assert _maxabs(out[k][0][:, ::9], g[k]) < TOL, k
    for net, m in (("enc", me), ("dec", md)):
        for k, p in m.named_parameters():
            if p.grad is None:
                continue
            gr = p.grad.detach().cpu().numpy()
            assert np.sqrt((gr.astype(np.float64) ** 2).sum()) == pytest.approx(float(g[f"gnorm/{net}/{k}"]), rel=2e-4), k
            samp = gr.reshape(-1)[:: max(1, gr.size // 64)][:64]
            ref = g[f"gsamp/{net}/{k}"]
            assert np.abs(samp - ref).max() < 2e-4 * max(1.0, np.abs(ref).max()), k


def test_flagship_cyc2_step_vs_reference_golden(golden_dir, cvb):
    """configs[1] shapes (hu1024 ld32 cyc2, T=80) at B=2, ragged flen_acc, KL-cv quirk."""
    g = _load(golden_dir, "flagship.npz")
    lat, B, T, n_cyc = 32, 2, 80, 2
    enc, dec, me, md, x, cv, sc, tc, eps, masks, y0e, y0d = _cyc_setup(cvb, 1024, lat, B, T, n_cyc, 3, 201, 202)
    out, total = _run_cyc(cvb, me, md, x, cv, sc, tc, eps, masks, y0e, y0d, n_cyc, lat, [T, 61], [0, 1])
    total.backward()
    torch.cuda.synchronize()
    assert _paths() == (PATH_TC, PATH_TC), "hu1024 training must run the tcgen05 recurrence kernels, not a fallback"
    assert total.item() == pytest.approx(float(g["cyc2/loss"]), rel=5e-6)
    for k in out:
        for i in range(n_cyc):
            assert _maxabs(out[k][i][:, ::8], g[f"cyc2/{k}/{i}"]) < TOL, (k, i)
    for net, m in (("enc", me), ("dec", md)):
        for k, p in m.named_parameters():
            if p.grad is None:
                continue
            gr = p.grad.detach().cpu().numpy()
            assert np.sqrt((gr.astype(np.float64) ** 2).sum()) == pytest.approx(float(g[f"cyc2/gnorm/{net}/{k}"]), rel=3e-4), k
            samp = gr.reshape(-1)[:: max(1, gr.size // 64)][:64]
            ref = g[f"cyc2/gsamp/{net}/{k}"]
            assert np.abs(samp - ref).max() < 3e-4 * max(1.0, np.abs(ref).max()), k


@pytest.mark.parametrize("tag,gain,bstd", [("init", 1.0, 0.0), ("trained", 3.0, 0.05)])
def test_flagship_decode_and_carry(golden_dir, cvb, tag, gain, bstd):
    """hu1024 ld32: stage-6 conversion of an 800-frame utterance in the reference's unbatched layout and
    two 80-frame chunks with carried (y, h); reference-init and 'trained-like' (x3 weights) parameters."""
    from cyclevae_vc_b200 import cycle
    g = _load(golden_dir, "flagship.npz")
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024)
    Pe = orc.init_params(enc, 201, gain=gain, bias_std=bstd, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, gain=gain, bias_std=bstd, mean=mean[stdim:], scale=std[stdim:])
    me, md = _module(cvb, enc, Pe).eval(), _module(cvb, dec, Pd).eval()
    y0d1 = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).cuda()
    T = 800
    x, _, sc, tc = orc.synth_batch(1, T, 1)
    eps_mean = (orc.synth_noise(1, T, lat, 1, 1)[0][0] / np.sqrt(300.0)).cuda()
    with torch.no_grad():
        lat_src, _, _ = me(x[0].cuda(), torch.zeros(1, 1, 2 * lat).cuda(), clamp_vae=True, lat_dim=lat)
        cvm = cycle.convert(me, md, x[0].cuda(), tc[0].cuda(), lat_dim=lat, y0_enc=torch.zeros(1, 1, 2 * lat).cuda(),
                            y0_dec=y0d1, eps_mean=eps_mean[0])
    assert lat_src.shape == (T, 2 * lat) and cvm.shape == (T, 50)
    assert _paths()[0] == PATH_TC_FOLDED, "hu1024 inference must run the folded tcgen05 kernel, not a fallback"
    assert _maxabs(lat_src[::5], g[f"{tag}/dec800_lat"]) < TOL
    if tag == "init":
        assert _maxabs(cvm[::5], g[f"{tag}/dec800_cvmcep"]) < TOL
    else:
        # Stress set (weights x3, |mcep| up to 21): the fp32 REFERENCE is itself 1.4e-4 away from exact (fp64)
        # arithmetic after 800 recurrent steps, so "within 1e-4 of the reference" is below fp32 noise here.
        # Bar: no further from the fp64 oracle than 1.5x the reference's own distance, and within 1e-4 relative to the
        # output scale of the reference.  (Round 1 sat at 2.3x: tools/split_error_budget.py traced it to the truncating
        # accumulation inside long tcgen05 chains and to unscaled fp16 lo planes; both are fixed in the kernels.)
        P64e, P64d = ({k: v.double() for k, v in P.items()} for P in (Pe, Pd))
        exact = orc.convert(P64e, P64d, enc, dec, x[0].double(), tc[0].double(), lat_dim=lat,
                            y0_enc=torch.zeros(1, 1, 2 * lat, dtype=torch.float64), y0_dec=y0d1.cpu().double(),
                            eps_mean=eps_mean[0].cpu().double()).numpy()[::5]
        ref = g[f"{tag}/dec800_cvmcep"]
        ref_vs_exact = np.abs(ref - exact).max()
        mine_vs_exact = np.abs(cvm[::5].cpu().numpy() - exact).max()
        assert mine_vs_exact <= max(TOL, 1.5 * ref_vs_exact), (mine_vs_exact, ref_vs_exact)
        assert _maxabs(cvm[::5], ref) < TOL * max(1.0, np.abs(ref).max())
    B, T = 3, 80
    x, cv, sc, tc = (t.cuda() for t in orc.synth_batch(B, 2 * T, 2))
    with torch.no_grad():
        o1, y1, h1 = me(x[:, :T], torch.zeros(B, 1, 2 * lat).cuda(), clamp_vae=True, lat_dim=lat)
        o2, y2, h2 = me(x[:, T:], y1, h_in=h1, clamp_vae=True, lat_dim=lat)
        zin = torch.cat((sc, torch.cat((o1, o2), 1)[:, :, :lat]), 2)
        d1, yd1, hd1 = md(zin[:, :T], y0d1.repeat(B, 1, 1))
        d2, yd2, hd2 = md(zin[:, T:], yd1, h_in=hd1)
    assert _maxabs(torch.cat((o1, o2), 1)[:, ::4], g[f"{tag}/carry_lat"]) < TOL
    ref = g[f"{tag}/carry_mcep"]
    if tag == "init":
        assert _maxabs(torch.cat((d1, d2), 1)[:, ::4], ref) < TOL
    else:   # stress set: same bar as above (distance to the fp64 oracle vs the reference's own distance)
        P64e, P64d = ({k: v.double() for k, v in P.items()} for P in (Pe, Pd))
        xc, scc = x.cpu().double(), sc.cpu().double()
        e1, ey1, eh1 = orc.gru_rnn_forward(P64e, enc, xc[:, :T], torch.zeros(B, 1, 2 * lat, dtype=torch.float64), clamp_vae=True, lat_dim=lat)
        e2, _, eh2 = orc.gru_rnn_forward(P64e, enc, xc[:, T:], ey1, eh1, clamp_vae=True, lat_dim=lat)
        zin64 = torch.cat((scc, torch.cat((e1, e2), 1)[:, :, :lat]), 2)
        f1, fy1, fh1 = orc.gru_rnn_forward(P64d, dec, zin64[:, :T], y0d1.cpu().double().repeat(B, 1, 1))
        f2, _, fh2 = orc.gru_rnn_forward(P64d, dec, zin64[:, T:], fy1, fh1)
        exact = torch.cat((f1, f2), 1).numpy()[:, ::4]
        ref_vs_exact = np.abs(ref - exact).max()
        mine_vs_exact = np.abs(torch.cat((d1, d2), 1)[:, ::4].cpu().numpy() - exact).max()
        assert mine_vs_exact <= max(TOL, 1.5 * ref_vs_exact), (mine_vs_exact, ref_vs_exact)
        assert _maxabs(torch.cat((d1, d2), 1)[:, ::4], ref) < TOL * max(1.0, np.abs(ref).max())
    if tag == "init":
        assert _maxabs(h2[:, :, ::8], g[f"{tag}/carry_h_enc"]) < TOL
        assert _maxabs(hd2[:, :, ::8], g[f"{tag}/carry_h_dec"]) < TOL
    else:   # stress set: the carried states against the fp64 oracle, same 1.5x bar
        for mine, exact_h, key in ((h2, eh2, "carry_h_enc"), (hd2, fh2, "carry_h_dec")):
            ex = exact_h.numpy()[:, :, ::8]
            ref_d = np.abs(g[f"{tag}/{key}"] - ex).max()
            assert np.abs(mine[:, :, ::8].cpu().numpy() - ex).max() <= max(TOL, 1.5 * ref_d), key


def test_spk4_decoder(golden_dir, cvb):
    """configs[3]: 4-speaker one-hot code (decoder in_dim 36)."""
    g = _load(golden_dir, "spk4.npz")
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    dec = orc.decoder_spec(lat, 4, 50, 1024)
    Pd = orc.init_params(dec, 302, mean=mean[stdim:], scale=std[stdim:])
    md = _module(cvb, dec, Pd).eval()
    B, T = 2, 40
    z = torch.randn(B, T, lat, generator=torch.Generator().manual_seed(55))
    code = torch.zeros(B, T, 4)
    code[0, :, 2] = 1
    code[1, :, 3] = 1
    y0 = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    with torch.no_grad():
        o, y, h = md(torch.cat((code, z), 2).cuda(), y0.cuda())
    assert _paths()[0] == PATH_TC_FOLDED
    assert _maxabs(o, g["trj"]) < TOL
    assert _maxabs(y, g["y"]) < TOL
    assert _maxabs(h[:, :, ::8], g["h"]) < TOL


def test_full_size_properties(cvb):
    """BASELINE.json full sizes (hu1024, B=80 x T=80 training chunk; wide decode batch), checked through
    size-independent properties: run-to-run determinism (no float atomics), batch-row independence
    (row j of a batched call == the unbatched call on row j), a seeded oracle spot-check of a few rows,
    and zero gradient to padded frames' inputs when the loss masks them out of reach of the receptive field."""
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    enc = orc.encoder_spec(54, lat, 1024)
    Pe = orc.init_params(enc, 201, gain=2.0, bias_std=0.02, mean=mean, scale=std)
    me = _module(cvb, enc, Pe).eval()
    B, T = 80, 80
    x, _, _, _ = orc.synth_batch(B, T, 7)
    xg = x.cuda()
    y0 = torch.zeros(B, 1, 2 * lat).cuda()
    with torch.no_grad():
        o1, y1, h1 = me(xg, y0, clamp_vae=True, lat_dim=lat)
        o2, y2, h2 = me(xg, y0, clamp_vae=True, lat_dim=lat)
        assert torch.equal(o1, o2) and torch.equal(h1, h2) and torch.equal(y1, y2)
        for j in (0, 37, 79):
            oj, yj, hj = me(xg[j], y0[j:j + 1], clamp_vae=True, lat_dim=lat)
            assert _maxabs(oj, o1[j]) < 5e-5 and _maxabs(hj[0, 0], h1[0, j]) < 5e-5   # batched gx: tensor-core GEMM, unbatched: cuBLAS
        # live oracle calls run in float64 (torch's CPU float32 kernels occasionally return ~1e-4 less accurate results on the
        # GPU boxes' hosts; the fixtures under tests/golden were generated once and are not affected)
        Pe64 = {k: v.double() for k, v in Pe.items()}
        ref, _, href = orc.gru_rnn_forward(Pe64, enc, x[[3, 64]].double(), torch.zeros(2, 1, 2 * lat, dtype=torch.float64), clamp_vae=True,
                                           lat_dim=lat)
    assert _maxabs(o1[[3, 64]].double(), ref) < TOL
    assert _maxabs(h1[0, [3, 64]].double(), href[0]) < TOL
    # decode-side wide batch: 256 utterances x 200 frames (two batch tiles of the persistent kernel)
    Bd, Td = 256, 200
    xd, _, _, _ = orc.synth_batch(Bd, Td, 8)
    with torch.no_grad():
        od, _, hd = me(xd.cuda(), torch.zeros(Bd, 1, 2 * lat).cuda(), clamp_vae=True, lat_dim=lat)
        refd, _, hrefd = orc.gru_rnn_forward(Pe64, enc, xd[[0, 127, 128, 255]].double(), torch.zeros(4, 1, 2 * lat, dtype=torch.float64),
                                             clamp_vae=True, lat_dim=lat)
    assert _maxabs(od[[0, 127, 128, 255]].double(), refd) < TOL
    assert torch.isfinite(od).all()


@pytest.mark.parametrize("feedback", ["default", "grid", "cluster"])
@pytest.mark.parametrize("net,B", [("enc", 80), ("dec", 80), ("dec", 5), ("enc", 128)])
def test_tensor_core_recurrence_vs_exact_kernels(cvb, net, B, feedback):
    """The tcgen05 recurrence kernels (forward: fp16 hi/lo split; BPTT: bf16 hi/lo split over thread-block
    clusters) against the fp32-FMA persistent kernels (CVB_RECURRENCE=exact, themselves checked against the
    oracle above) at the BASELINE.json chunk size hu1024 x T=80, dropout masks, carried h_in / y_in,
    gradients w.r.t. inputs, carried state and every parameter.  Both generations of the training kernels are checked
    at every shape: two grid-wide exchanges per step (CVB_TC_FEEDBACK=grid: gru_tc.cu / gru_tc_bwd.cu) and one
    (cluster: gru_tc2.cu / gru_tc2_bwd.cu, fixed-point totals of the y feedback); the default picks per shape."""
    lat, stdim, T = 32, 4, 80
    mean, std = orc.synth_stats(50)
    if net == "enc":
        spec = orc.encoder_spec(54, lat, 1024)
        P = orc.init_params(spec, 201, gain=1.5, bias_std=0.02, mean=mean, scale=std)
    else:
        spec = orc.decoder_spec(lat, 2, 50, 1024)
        P = orc.init_params(spec, 202, gain=1.5, bias_std=0.02, mean=mean[stdim:], scale=std[stdim:])
    m = _module(cvb, spec, P).train()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, T, spec.in_dim, generator=g).cuda()
    y0 = (0.3 * torch.randn(B, 1, spec.out_dim, generator=g)).cuda()
    h0 = (0.5 * torch.randn(1, B, 1024, generator=g)).cuda()
    mc = ((torch.rand(B, T, spec.conv_dim, generator=g) >= 0.5).float() * 2).cuda()
    mg = ((torch.rand(B, T, 1024, generator=g) >= 0.5).float() * 2).cuda()
    w_o = torch.randn(B, T, spec.out_dim, generator=g).cuda()
    w_y = torch.randn(B, 1, spec.out_dim, generator=g).cuda()
    w_h = torch.randn(1, B, 1024, generator=g).cuda()

    def run(mode):
        # "exact" = the all-fp32 path: fp32-FMA recurrence kernels AND cuBLAS fp32 for every dense product (so the
        # split-precision GEMM, the tap-fused conv products and the in-kernel bias gradients are A/B-checked too)
        if mode:
            os.environ["CVB_RECURRENCE"] = mode
            os.environ["CVB_GEMM"] = "cublas"
        else:
            os.environ.pop("CVB_RECURRENCE", None)
            os.environ.pop("CVB_GEMM", None)
            if feedback != "default":
                os.environ["CVB_TC_FEEDBACK"] = feedback
        try:
            xs, ys, hs = (t.clone().requires_grad_(True) for t in (x, y0, h0))
            for p in m.parameters():
                p.grad = None
            m.inject_dropout_masks(mc, mg)
            o, yl, hl = m(xs, ys, h_in=hs, do=True, clamp_vae=(net == "enc"), lat_dim=lat)
            ((o * w_o).sum() + (yl * w_y).sum() + (hl * w_h).sum()).backward()
            torch.cuda.synchronize()
            grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
            # no silent fallback: the tensor-core kernels (path 1) must be the ones that ran by default
            from cyclevae_vc_b200._lib import lib
            want = 0 if mode else 1
            assert lib.cvb_last_recurrence_path(0) == want and lib.cvb_last_recurrence_path(1) == want
            hops = (lib.cvb_last_recurrence_hops(0), lib.cvb_last_recurrence_hops(1))
            if mode:
                assert hops == (0, 0)
            else:   # default: the one-exchange forward kernel at every shape, the one-exchange BPTT kernel up to 32 rows
                assert hops == {"grid": (2, 2), "cluster": (1, 1), "default": (1, 1 if B <= 32 else 2)}[feedback], hops
            return (o.detach(), yl.detach(), hl.detach()), (xs.grad, ys.grad, hs.grad), grads
        finally:
            os.environ.pop("CVB_RECURRENCE", None)
            os.environ.pop("CVB_GEMM", None)
            os.environ.pop("CVB_TC_FEEDBACK", None)

    out_e, gin_e, gp_e = run("exact")
    out_t, gin_t, gp_t = run(None)
    for a, b in zip(out_t, out_e):
        assert _maxabs(a, b) < 5e-5 * max(1.0, float(b.abs().max()))
    for a, b in zip(gin_t, gin_e):
        assert _maxabs(a, b) < 1e-4 * max(1e-3, float(b.abs().max())), (float(b.abs().max()))
    assert set(gp_t) == set(gp_e)
    for k in gp_e:
        assert _maxabs(gp_t[k], gp_e[k]) < 1e-4 * max(1e-3, float(gp_e[k].abs().max())), k


@pytest.mark.parametrize("f16", [1, 0])
def test_split_precision_tensor_core_gemm(cvb, f16):
    """cvb_gemm_tc (fp16 / bf16 hi+lo operands, 3 products, fp32 TMEM accumulation) against float64: ragged M/N/K,
    all four transpose combinations, accumulate and bias epilogues, leading dimensions wider than the rows."""
    from cyclevae_vc_b200._lib import check, lib, ptr
    g = torch.Generator().manual_seed(3)
    tol = 5e-6 if f16 else 6e-5   # x sqrt(K): operand split 2^-22 / 2^-17 per product, max over the outputs (~4 sigma);
    #                             K-slices of 128 are summed in fp32 registers, so the error does not grow ~K
    shapes = [(128, 128, 64, 0, 1), (6400, 3072, 486, 0, 1), (3072, 1024, 6400, 1, 0), (6400, 306, 3072, 0, 0),
              (200, 50, 100, 1, 1), (130, 66, 31, 1, 1), (257, 129, 65, 0, 0), (64, 1030, 777, 1, 0),
              # output rows 16-byte aligned (ldc % 4 == 0) but N % 4 != 0: the last 16-byte group of a row straddles N;
              # 8-byte aligned rows (ldc even) with an odd N; aligned rows wider than one column tile
              (300, 306, 200, 0, 1, 308), (300, 305, 200, 1, 0, 306), (140, 486, 130, 0, 0, 488)]
    for shape in shapes:
        M, N, K, ta, tb = shape[:5]
        lda = (M if ta else K) + 3
        ldb = (K if tb else N) + 5
        ldc = shape[5] if len(shape) > 5 else N + 1
        A = torch.randn((K if ta else M), lda, generator=g).cuda()
        Bm = torch.randn((N if tb else K), ldb, generator=g).cuda()
        bias = torch.randn(N, generator=g).cuda()
        C0 = torch.randn(M, ldc, generator=g).cuda()
        Aeff = (A[:, :M].t() if ta else A[:, :K]).double()
        Beff = (Bm[:, :K].t() if tb else Bm[:, :N]).double()
        prod = Aeff @ Beff
        scale = float(np.sqrt(K))
        for beta1, use_bias in ((0, 0), (1, 1)):
            Cm = C0.clone()
            check(lib.cvb_gemm_tc(ta, tb, M, N, K, ptr(A), lda, ptr(Bm), ldb, beta1, ptr(bias) if use_bias else None, ptr(Cm), ldc,
                                  f16, torch.cuda.current_stream().cuda_stream), "cvb_gemm_tc")
            ref = prod + (C0[:, :N].double() if beta1 else 0) + (bias.double() if use_bias else 0)
            err = (Cm[:, :N].double() - ref).abs().max().item()
            assert err < tol * scale, (M, N, K, ta, tb, beta1, err)
            assert torch.equal(Cm[:, N:], C0[:, N:])   # nothing written past the row


def test_tensor_core_forward_any_row_count(cvb):
    """Batch-row counts that are not multiples of the 8-row operand groups, the widest single launch, and a wide batch that
    the host slices (cvb_recurrence_max_rows): tensor-core forward == fp32-FMA forward, eval mode, two passes each (the
    exchange buffers are reused across launches, so stale contents must never leak into live rows)."""
    torch.manual_seed(0)
    enc = cvb.GRU_RNN(in_dim=54, out_dim=64, hidden_units=1024, do_prob=0.5, scale_out_flag=False).cuda().eval()
    enc.apply(cvb.initialize)
    for B, T in ((1, 1), (3, 1), (17, 2), (17, 9), (43, 12), (86, 12), (128, 6), (300, 5)):
        x = torch.randn(B, T, 54, device="cuda")
        y0 = 0.1 * torch.randn(B, 1, 64, device="cuda")
        h0 = 0.3 * torch.randn(1, B, 1024, device="cuda")
        with torch.no_grad():
            os.environ["CVB_RECURRENCE"] = "exact"
            try:
                o_e, y_e, h_e = enc(x, y0, h_in=h0, clamp_vae=True, lat_dim=32)
            finally:
                os.environ.pop("CVB_RECURRENCE", None)
            from cyclevae_vc_b200._lib import lib
            assert lib.cvb_last_recurrence_path(0) == 0   # fp32-FMA kernels
            for _ in range(2):
                o_t, y_t, h_t = enc(x, y0, h_in=h0, clamp_vae=True, lat_dim=32)
                assert lib.cvb_last_recurrence_path(0) == 2   # the folded inference kernel, not a fallback
                assert _maxabs(o_t, o_e) < 2e-5 and _maxabs(h_t, h_e) < 2e-5 and _maxabs(y_t, y_e) < 2e-5, (B, T)


def test_ragged_batch_conversion_equals_per_utterance(cvb):
    """stage-6 conversion of utterances of different lengths packed into padded row groups
    (cycle.convert_utterances) == the reference's one-utterance-per-call composition (decode_*.py:303-305,318), checked
    against (i) the fp64 oracle run utterance by utterance and (ii) this library's own unbatched conversion.  The pads must
    be invisible: the encoder's pads are the scale_in mean (zeros of the normalised domain, gru_vae.py:336,357), the
    decoder's are zeros, and the recurrence is causal.  Plus the GV post-filter (decode_*.py:419-420)."""
    from cyclevae_vc_b200 import cycle
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024)
    Pe = orc.init_params(enc, 201, gain=1.0, bias_std=0.0, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, gain=1.0, bias_std=0.0, mean=mean[stdim:], scale=std[stdim:])
    me, md = _module(cvb, enc, Pe).eval(), _module(cvb, dec, Pd).eval()
    assert _maxabs(cycle.scale_in_mean(me), mean.astype(np.float32)) < 1e-5
    y0e = torch.zeros(1, 1, 2 * lat)
    y0d = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1)
    lens = [37, 120, 64, 5, 200, 81, 9]
    feats, codes, epss = [], [], []
    for i, T in enumerate(lens):
        x, _, sc, tc = orc.synth_batch(1, T, 40 + i)
        feats.append(x[0])
        codes.append(tc[0])
        epss.append(orc.synth_noise(1, T, lat, 1, 50 + i)[0][0][0] / np.sqrt(300.0))   # [T, lat]
    trg_code = codes[0][0]
    assert all(torch.equal(c, trg_code.expand_as(c)) for c in codes)
    with torch.no_grad():
        got = cycle.convert_utterances(me, md, [f.cuda() for f in feats], trg_code.cuda(), lat_dim=lat, y0_enc=y0e.cuda(),
                                       y0_dec=y0d.cuda(), eps_means=[e.cuda() for e in epss], rows_per_group=3)
        alone = [cycle.convert(me, md, f.cuda(), c.cuda(), lat_dim=lat, y0_enc=y0e.cuda(), y0_dec=y0d.cuda(), eps_mean=e.cuda())
                 for f, c, e in zip(feats, codes, epss)]
    P64e, P64d = ({k: v.double() for k, v in P.items()} for P in (Pe, Pd))
    for i, T in enumerate(lens):
        assert got[i].shape == (T, 50)
        assert _maxabs(got[i], alone[i]) < 2e-5, i
        exact = orc.convert(P64e, P64d, enc, dec, feats[i].double(), codes[i].double(), lat_dim=lat, y0_enc=y0e.double(),
                            y0_dec=y0d.double(), eps_mean=epss[i].double())
        assert _maxabs(got[i], exact.float()) < TOL, i
    # GV post-filter, decode_*.py:419-420 restated in numpy
    cv = got[4].double().cpu().numpy()
    rng = np.random.default_rng(0)
    gv_trg, cvgv = rng.uniform(0.5, 2.0, 49), rng.uniform(0.5, 2.0, 49)
    datamean = np.mean(cv[:, 1:], axis=0)
    ref = np.c_[cv[:, 0], np.sqrt(gv_trg / cvgv) * (cv[:, 1:] - datamean) + datamean]
    mine = cycle.gv_postfilter(got[4], torch.tensor(gv_trg, dtype=torch.float32).cuda(), torch.tensor(cvgv, dtype=torch.float32).cuda())
    assert _maxabs(mine, ref.astype(np.float32)) < 1e-5


def test_tensor_core_recurrence_shortest_sequences(cvb):
    """Edge shapes of the training kernels: a single frame (T = 1: the prologue and one step, nothing carried inside the
    launch) and a single utterance (B = 1), tensor-core path vs the all-fp32 path, outputs and every gradient."""
    from cyclevae_vc_b200._lib import lib
    lat = 32
    mean, std = orc.synth_stats(50)
    spec = orc.encoder_spec(54, lat, 1024)
    P = orc.init_params(spec, 201, gain=1.5, bias_std=0.02, mean=mean, scale=std)
    m = _module(cvb, spec, P).train()
    g = torch.Generator().manual_seed(17)
    for B, T in ((1, 1), (1, 7), (9, 1), (2, 3)):
        x = torch.randn(B, T, spec.in_dim, generator=g).cuda()
        y0 = (0.3 * torch.randn(B, 1, spec.out_dim, generator=g)).cuda()
        h0 = (0.5 * torch.randn(1, B, 1024, generator=g)).cuda()
        mc = ((torch.rand(B, T, spec.conv_dim, generator=g) >= 0.5).float() * 2).cuda()
        mg = ((torch.rand(B, T, 1024, generator=g) >= 0.5).float() * 2).cuda()
        w_o = torch.randn(B, T, spec.out_dim, generator=g).cuda()

        def run(exact):
            if exact:
                os.environ["CVB_RECURRENCE"] = "exact"
                os.environ["CVB_GEMM"] = "cublas"
            try:
                xs, ys, hs = (t.clone().requires_grad_(True) for t in (x, y0, h0))
                for p in m.parameters():
                    p.grad = None
                m.inject_dropout_masks(mc, mg)
                o, yl, hl = m(xs, ys, h_in=hs, do=True, clamp_vae=True, lat_dim=lat)
                ((o * w_o).sum() + yl.sum() + hl.sum()).backward()
                torch.cuda.synchronize()
                assert lib.cvb_last_recurrence_path(0) == (0 if exact else 1) and lib.cvb_last_recurrence_path(1) == (0 if exact else 1)
                return [o.detach(), yl.detach(), hl.detach(), xs.grad, ys.grad, hs.grad] + \
                       [p.grad.clone() for p in m.parameters() if p.grad is not None]
            finally:
                os.environ.pop("CVB_RECURRENCE", None)
                os.environ.pop("CVB_GEMM", None)

        ref, got = run(True), run(False)
        assert len(ref) == len(got)
        for a, b in zip(got, ref):
            assert _maxabs(a, b) < 1e-4 * max(1e-3, float(b.abs().max())), (B, T, tuple(b.shape))


def test_forward_only_kernel_choice(cvb):
    """Forward-only calls (torch.no_grad) and the kernel each one takes (cvb_last_recurrence_path): no dropout -> the folded
    one-exchange kernel; dropout masks -> the two-exchange kernel (the fold needs o_t = h_t); CVB_EVAL_FOLD=0 -> the
    two-exchange kernel without masks.  All against the all-fp32 path on the same masks, wide batch included (sliced at
    the row count of the kernel actually used)."""
    from cyclevae_vc_b200._lib import lib
    lat = 32
    mean, std = orc.synth_stats(50)
    spec = orc.decoder_spec(lat, 2, 50, 1024)
    P = orc.init_params(spec, 202, gain=1.5, bias_std=0.02, mean=mean[4:], scale=std[4:])
    m = _module(cvb, spec, P)
    g = torch.Generator().manual_seed(3)
    for B, T in ((24, 16), (150, 7)):
        x = torch.randn(B, T, spec.in_dim, generator=g).cuda()
        y0 = (0.3 * torch.randn(B, 1, spec.out_dim, generator=g)).cuda()
        h0 = (0.5 * torch.randn(1, B, 1024, generator=g)).cuda()
        mc = ((torch.rand(B, T, spec.conv_dim, generator=g) >= 0.5).float() * 2).cuda()
        mg = ((torch.rand(B, T, 1024, generator=g) >= 0.5).float() * 2).cuda()

        def run(do, env):
            os.environ.update(env)
            try:
                with torch.no_grad():
                    if do:
                        m.train()
                        m.inject_dropout_masks(mc, mg)
                    else:
                        m.eval()
                    out = m(x, y0, h_in=h0, do=do)
                torch.cuda.synchronize()
                return out, lib.cvb_last_recurrence_path(0)
            finally:
                for k in env:
                    os.environ.pop(k, None)

        for do in (False, True):
            ref, path = run(do, {"CVB_RECURRENCE": "exact", "CVB_GEMM": "cublas"})
            assert path == 0
            cases = [({}, 1 if do else 2)] + ([] if do else [({"CVB_EVAL_FOLD": "0"}, 1)])
            for env, want in cases:
                got, path = run(do, env)
                assert path == want, (B, do, env, path)
                for a, b in zip(got, ref):
                    assert _maxabs(a, b) < 5e-5 * max(1.0, float(b.abs().max())), (B, do, env)


def test_persistent_kernels_are_deterministic(cvb):
    """Every cross-CTA sum of the persistent kernels is taken in a fixed order (no float atomics, fixed K walk per
    cluster), so repeated calls must agree BIT FOR BIT — which also makes this the race detector for their exchange
    buffers (a slot overwritten while a peer still reads it shows up as a changed bit): folded inference kernel at the
    widest launch over 300 steps, training forward + BPTT at the bench shape."""
    torch.manual_seed(1)
    enc = cvb.GRU_RNN(in_dim=54, out_dim=64, hidden_units=1024, do_prob=0.5, scale_out_flag=False).cuda()
    enc.apply(cvb.initialize)
    with torch.no_grad():
        for p in enc.parameters():
            if p.dim() > 1 and p.requires_grad:
                p.mul_(2.0)   # livelier dynamics than the reference init: rounding differences would not die out
    B, T = 128, 300
    x = torch.randn(B, T, 54, device="cuda")
    y0 = 0.1 * torch.randn(B, 1, 64, device="cuda")
    h0 = 0.3 * torch.randn(1, B, 1024, device="cuda")
    enc.eval()
    with torch.no_grad():
        first = None
        for _ in range(5):
            out = enc(x, y0, h_in=h0, clamp_vae=True, lat_dim=32)
            if first is None:
                first = [t.clone() for t in out]
            else:
                assert all(torch.equal(a, b) for a, b in zip(out, first))
    from cyclevae_vc_b200._lib import lib
    assert lib.cvb_last_recurrence_path(0) == 2
    B, T = 80, 80
    x = torch.randn(B, T, 54, device="cuda")
    y0 = 0.1 * torch.randn(B, 1, 64, device="cuda")
    h0 = 0.3 * torch.randn(1, B, 1024, device="cuda")
    mc = (torch.rand(B, T, enc.in_dim * enc.receptive_field, device="cuda") >= 0.5).float() * 2
    mg = (torch.rand(B, T, 1024, device="cuda") >= 0.5).float() * 2
    w = torch.randn(B, T, 64, device="cuda")
    enc.train()
    # both generations of the training kernels; the one-exchange kernels sum the y feedback with integer (fixed-point)
    # atomics, whose result does not depend on the arrival order
    for feedback in ("grid", "cluster"):
        os.environ["CVB_TC_FEEDBACK"] = feedback
        try:
            first = None
            for _ in range(3):
                xs, hs = x.clone().requires_grad_(True), h0.clone().requires_grad_(True)
                for p in enc.parameters():
                    p.grad = None
                enc.inject_dropout_masks(mc, mg)
                o, yl, hl = enc(xs, y0, h_in=hs, do=True, clamp_vae=True, lat_dim=32)
                ((o * w).sum() + hl.sum()).backward()
                got = [o.detach().clone(), hl.detach().clone(), xs.grad.clone(), hs.grad.clone()] + \
                      [p.grad.clone() for p in enc.parameters() if p.grad is not None]
                if first is None:
                    first = got
                else:
                    assert all(torch.equal(a, b) for a, b in zip(got, first)), feedback
            assert lib.cvb_last_recurrence_path(0) == 1 and lib.cvb_last_recurrence_path(1) == 1
            assert lib.cvb_last_recurrence_hops(0) == lib.cvb_last_recurrence_hops(1) == (2 if feedback == "grid" else 1)
        finally:
            os.environ.pop("CVB_TC_FEEDBACK", None)


@pytest.mark.parametrize("gscale", [1e-9, 1.0, 1e7])
def test_one_exchange_bptt_any_gradient_scale(cvb, gscale):
    """The one-exchange BPTT kernel adds the cluster partials of the y feedback as 64-bit fixed-point numbers whose scale is
    chosen per launch from max |dY|, |dh_last|: gradients of a loss scaled by 1e-9 or 1e7 must come out as accurate
    (relative to their own size) as at scale 1 -- against the all-fp32 path run on the same scaled loss."""
    lat, T, B = 32, 40, 8
    mean, std = orc.synth_stats(50)
    spec = orc.encoder_spec(54, lat, 1024)
    P = orc.init_params(spec, 203, gain=1.5, bias_std=0.02, mean=mean, scale=std)
    m = _module(cvb, spec, P).train()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(B, T, spec.in_dim, generator=g).cuda()
    y0 = (0.3 * torch.randn(B, 1, spec.out_dim, generator=g)).cuda()
    mc = ((torch.rand(B, T, spec.conv_dim, generator=g) >= 0.5).float() * 2).cuda()
    mg = ((torch.rand(B, T, 1024, generator=g) >= 0.5).float() * 2).cuda()
    w_o = (gscale * torch.randn(B, T, spec.out_dim, generator=g)).cuda()
    from cyclevae_vc_b200._lib import lib

    def run(exact):
        if exact:
            os.environ["CVB_RECURRENCE"] = "exact"
            os.environ["CVB_GEMM"] = "cublas"
        else:
            os.environ["CVB_TC_FEEDBACK"] = "cluster"
        try:
            xs = x.clone().requires_grad_(True)
            for p in m.parameters():
                p.grad = None
            m.inject_dropout_masks(mc, mg)
            o, yl, hl = m(xs, y0, do=True, clamp_vae=True, lat_dim=lat)
            (o * w_o).sum().backward()
            torch.cuda.synchronize()
            assert lib.cvb_last_recurrence_hops(1) == (0 if exact else 1)
            return xs.grad.clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        finally:
            for k in ("CVB_RECURRENCE", "CVB_GEMM", "CVB_TC_FEEDBACK"):
                os.environ.pop(k, None)

    gx_e, gp_e = run(True)
    gx_t, gp_t = run(False)
    assert torch.isfinite(gx_t).all() and float(gx_e.abs().max()) > 0
    assert _maxabs(gx_t, gx_e) < 1e-4 * float(gx_e.abs().max())
    for k in gp_e:
        assert _maxabs(gp_t[k], gp_e[k]) < 1e-4 * max(1e-3 * gscale, float(gp_e[k].abs().max())), k


@pytest.mark.parametrize("cluster8", ["0", "1"])
def test_tensor_core_recurrence_hu512_both_cluster_shapes(cvb, cluster8):
    """hidden_units = 512 (64 CTAs): the BPTT kernel's 8-CTA cluster shape is co-resident here (8 clusters), so both K-split
    shapes are checked against the all-fp32 path, forward and gradients."""
    lat, T, B = 16, 24, 40
    spec = orc.NetSpec(in_dim=20, out_dim=2 * lat, hidden_units=512, do_prob=0.5, scale_in=True, scale_out=False)
    rng = np.random.default_rng(3)
    P = orc.init_params(spec, 77, gain=1.5, bias_std=0.02, mean=rng.normal(size=20), scale=rng.uniform(0.5, 2.0, size=20))
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, 20, generator=g).cuda()
    y0 = (0.3 * torch.randn(B, 1, 2 * lat, generator=g)).cuda()
    h0 = (0.5 * torch.randn(1, B, 512, generator=g)).cuda()
    mc = ((torch.rand(B, T, spec.conv_dim, generator=g) >= 0.5).float() * 2).cuda()
    mg = ((torch.rand(B, T, 512, generator=g) >= 0.5).float() * 2).cuda()
    w_o = torch.randn(B, T, 2 * lat, generator=g).cuda()

    def run(exact):
        env = {"CVB_RECURRENCE": "exact", "CVB_GEMM": "cublas"} if exact else {"CVB_TC_CLUSTER8": cluster8}
        os.environ.update(env)
        try:
            m = _module(cvb, spec, P).train()   # fresh module: the cluster shape is chosen once per shape and process-cached
            xs, ys, hs = (t.clone().requires_grad_(True) for t in (x, y0, h0))
            m.inject_dropout_masks(mc, mg)
            o, yl, hl = m(xs, ys, h_in=hs, do=True, clamp_vae=True, lat_dim=lat)
            ((o * w_o).sum() + hl.sum()).backward()
            torch.cuda.synchronize()
            return o.detach(), hl.detach(), xs.grad, hs.grad, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        finally:
            for k in env:
                os.environ.pop(k, None)

    o_e, h_e, dx_e, dh_e, gp_e = run(True)
    o_t, h_t, dx_t, dh_t, gp_t = run(False)
    assert _maxabs(o_t, o_e) < 5e-5 * max(1.0, float(o_e.abs().max())) and _maxabs(h_t, h_e) < 5e-5
    for a, b in ((dx_t, dx_e), (dh_t, dh_e)):
        assert _maxabs(a, b) < 1e-4 * max(1e-3, float(b.abs().max()))
    for k in gp_e:
        assert _maxabs(gp_t[k], gp_e[k]) < 1e-4 * max(1e-3, float(gp_e[k].abs().max())), k


@pytest.mark.parametrize("net", ["enc", "dec"])
def test_bench_shape_training_pass_vs_fp64_oracle(cvb, net):
    """The bench workload's pass shape (hu1024, 80 utterances x 80 frames, dropout masks, carried y_in / h_in) on the
    tcgen05 kernels against the FLOAT64 ORACLE (not against this library's own fp32 kernels): outputs, gradients w.r.t.
    x / y_in / h_in and every parameter gradient (all elements + norm).  gru_vae.py:364-399, SURVEY.md A.3."""
    lat, stdim, B, T = 32, 4, 80, 80
    mean, std = orc.synth_stats(50)
    if net == "enc":
        spec = orc.encoder_spec(54, lat, 1024)
        P = orc.init_params(spec, 201, gain=1.0, bias_std=0.02, mean=mean, scale=std)
    else:
        spec = orc.decoder_spec(lat, 2, 50, 1024)
        P = orc.init_params(spec, 202, gain=1.0, bias_std=0.02, mean=mean[stdim:], scale=std[stdim:])
    g = torch.Generator().manual_seed(23)
    if net == "enc":
        x = orc.synth_batch(B, T, 31)[0]
    else:
        x = torch.cat((orc.synth_batch(B, T, 31)[2], torch.randn(B, T, lat, generator=g)), 2)
    y0 = 0.3 * torch.randn(B, 1, spec.out_dim, generator=g)
    h0 = 0.5 * torch.randn(1, B, 1024, generator=g)
    mc = (torch.rand(B, T, spec.conv_dim, generator=g) >= 0.5).float() * 2
    mg = (torch.rand(B, T, 1024, generator=g) >= 0.5).float() * 2
    w_o = torch.randn(B, T, spec.out_dim, generator=g)
    w_y = torch.randn(B, 1, spec.out_dim, generator=g)
    w_h = torch.randn(1, B, 1024, generator=g)
    kw = dict(clamp_vae=True, lat_dim=lat) if net == "enc" else {}
    # float64 oracle + autograd
    P64 = {k: v.double().requires_grad_(not k.startswith("scale_")) for k, v in P.items()}
    xr, yr, hr = (t.double().requires_grad_(True) for t in (x, y0, h0))
    o_r, y_r, h_r = orc.gru_rnn_forward(P64, spec, xr, yr, hr, mask_conv=mc.double(), mask_gru=mg.double(), **kw)
    ((o_r * w_o.double()).sum() + (y_r * w_y.double()).sum() + (h_r * w_h.double()).sum()).backward()
    # device
    m = _module(cvb, spec, P).train()
    for k, p in m.named_parameters():
        p.requires_grad_(not k.startswith("scale_"))
    xs, ys, hs = (t.cuda().requires_grad_(True) for t in (x, y0, h0))
    m.inject_dropout_masks(mc.cuda(), mg.cuda())
    o, yl, hl = m(xs, ys, h_in=hs, do=True, **kw)
    ((o * w_o.cuda()).sum() + (yl * w_y.cuda()).sum() + (hl * w_h.cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert _paths() == (PATH_TC, PATH_TC)
    for a, b, name in ((o, o_r, "trj"), (yl, y_r, "y_last"), (hl, h_r, "h_last")):
        assert _maxabs(a.double(), b.detach()) < TOL * max(1.0, float(b.abs().max())), name
    for a, b, name in ((xs.grad, xr.grad, "dx"), (ys.grad, yr.grad, "dy_in"), (hs.grad, hr.grad, "dh_in")):
        assert _maxabs(a.double(), b) < 2e-4 * max(1e-3, float(b.abs().max())), (name, float(b.abs().max()))
    n_checked = 0
    for k, p in m.named_parameters():
        if k.startswith("scale_"):
            continue
        ref = P64[k].grad
        assert ref is not None and p.grad is not None, k
        assert _maxabs(p.grad.double(), ref) < 2e-4 * max(1e-3, float(ref.abs().max())), (k, float(ref.abs().max()))
        assert float(p.grad.double().norm()) == pytest.approx(float(ref.norm()), rel=2e-4), k
        n_checked += 1
    assert n_checked == 10


def test_spk4_cyc2_step_vs_reference_golden(golden_dir, cvb):
    """configs[3]: 4-speaker one-hot codes (decoder in_dim 36: conv 324, GRU input 374 -- ragged for 64-wide K chunks),
    training-mode cyc2 step at B=2 T=80 against fixtures produced by the unmodified reference."""
    g = _load(golden_dir, "spk4_cyc2.npz")
    lat, B, T, n_cyc = 32, 2, 80, 2
    enc, dec, me, md, x, cv, _, _, eps, masks, y0e, y0d = _cyc_setup(cvb, 1024, lat, B, T, n_cyc, 13, 301, 302, n_spk=4)
    x, cv, _, _ = orc.synth_batch(B, T, 13, n_spk=4)
    sc, tc = torch.tensor(g["src_code"]), torch.tensor(g["trg_code"])
    out, total = _run_cyc(cvb, me, md, x, cv, sc, tc, eps, masks, y0e, y0d, n_cyc, lat, [T, 70], [0, 1])
    total.backward()
    torch.cuda.synchronize()
    assert md.in_dim == 36 and _paths() == (PATH_TC, PATH_TC)
    assert total.item() == pytest.approx(float(g["loss"]), rel=5e-6)
    for k in out:
        for i in range(n_cyc):
            assert _maxabs(out[k][i][:, ::8], g[f"{k}/{i}"]) < TOL, (k, i)
    for net, m in (("enc", me), ("dec", md)):
        for k, p in m.named_parameters():
            if p.grad is None:
                continue
            gr = p.grad.detach().cpu().numpy()
            assert np.sqrt((gr.astype(np.float64) ** 2).sum()) == pytest.approx(float(g[f"gnorm/{net}/{k}"]), rel=3e-4), k
            samp = gr.reshape(-1)[:: max(1, gr.size // 64)][:64]
            ref = g[f"gsamp/{net}/{k}"]
            assert np.abs(samp - ref).max() < 3e-4 * max(1.0, np.abs(ref).max()), k


def test_decode512_batch_vs_reference_golden(golden_dir, cvb):
    """configs[2] at full size: stage-6 conversion of 512 utterances x 800 frames in one batched call (4 launches of 128
    rows per network) against the reference run one utterance at a time on 16 of them, 4 per 128-row slice."""
    from cyclevae_vc_b200 import cycle
    g = _load(golden_dir, "decode512.npz")
    lat, stdim, B, T = 32, 4, 512, 800
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024)
    Pe = orc.init_params(enc, 201, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, mean=mean[stdim:], scale=std[stdim:])
    me, md = _module(cvb, enc, Pe).eval(), _module(cvb, dec, Pd).eval()
    x, _, sc, tc = orc.synth_batch(B, T, 21)
    eps_mean = orc.synth_noise(B, T, lat, 1, 21)[0][0] / np.sqrt(300.0)
    assert float(x.double().abs().sum()) == pytest.approx(float(g["x_sum"]), rel=1e-12)
    y0d = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1).cuda()
    y0e = torch.zeros(B, 1, 2 * lat).cuda()
    with torch.no_grad():
        lat_src, _, _ = me(x.cuda(), y0e, clamp_vae=True, lat_dim=lat)
        assert _paths()[0] == PATH_TC_FOLDED
        cvm = cycle.convert(me, md, x.cuda(), tc.cuda(), lat_dim=lat, y0_enc=y0e, y0_dec=y0d, eps_mean=eps_mean.cuda())
    assert cvm.shape == (B, T, 50) and torch.isfinite(cvm).all()
    rows = g["rows"].tolist()
    assert len(set(r // 128 for r in rows)) == 4
    assert _maxabs(lat_src[rows][:, ::5], g["lat"]) < TOL
    assert _maxabs(cvm[rows][:, ::5], g["cvmcep"]) < TOL


def test_wide_latent_encoder_eval_and_training(cvb):
    """out_dim > 64 (encoder at lat_dim = 50, egs/one-to-one/run.sh lists 50 and 64): inference takes the folded tcgen05
    kernel (any out_dim: y comes from a product after the launch) and its scratch must be sized for it; training falls
    to the fp32-FMA kernels (the two-exchange tcgen05 kernels hold out <= 64).  Both against the float64 oracle."""
    lat, B, T = 50, 8, 100
    mean, std = orc.synth_stats(50)
    spec = orc.encoder_spec(54, lat, 1024)
    P = orc.init_params(spec, 211, gain=1.0, bias_std=0.02, mean=mean, scale=std)
    P64 = {k: v.double() for k, v in P.items()}
    x = orc.synth_batch(B, T, 41)[0]
    y0 = torch.zeros(B, 1, 2 * lat)
    m = _module(cvb, spec, P).eval()
    with torch.no_grad():
        o, y, h = m(x.cuda(), y0.cuda(), clamp_vae=True, lat_dim=lat)
        o1, _, _ = m(x[0].cuda(), y0[:1].cuda(), clamp_vae=True, lat_dim=lat)     # B = 1: stage-6 decode of one utterance
        o_r, y_r, h_r = orc.gru_rnn_forward(P64, spec, x.double(), y0.double(), clamp_vae=True, lat_dim=lat)
    assert _paths()[0] == PATH_TC_FOLDED
    assert _maxabs(o.double(), o_r) < TOL and _maxabs(h.double(), h_r) < TOL and _maxabs(y.double(), y_r) < TOL
    assert _maxabs(o1.double(), o_r[0]) < TOL
    m.train()
    g = torch.Generator().manual_seed(5)
    mc = (torch.rand(B, T, spec.conv_dim, generator=g) >= 0.5).float() * 2
    mg = (torch.rand(B, T, 1024, generator=g) >= 0.5).float() * 2
    Pg = {k: v.double().requires_grad_(not k.startswith("scale_")) for k, v in P.items()}
    o_r, _, h_r = orc.gru_rnn_forward(Pg, spec, x.double(), y0.double(), mask_conv=mc.double(), mask_gru=mg.double(), clamp_vae=True, lat_dim=lat)
    (o_r.square().sum() + h_r.sum()).backward()
    for k, p in m.named_parameters():
        p.requires_grad_(not k.startswith("scale_"))
    m.inject_dropout_masks(mc.cuda(), mg.cuda())
    o, _, h = m(x.cuda(), y0.cuda(), do=True, clamp_vae=True, lat_dim=lat)
    (o.square().sum() + h.sum()).backward()
    torch.cuda.synchronize()
    assert _paths() == (PATH_FP32, PATH_FP32)
    assert _maxabs(o.double(), o_r.detach()) < TOL
    ref = Pg["gru.weight_hh_l0"].grad
    assert _maxabs(m.gru.weight_hh_l0.grad.double(), ref) < 2e-4 * max(1.0, float(ref.abs().max()))


def test_fused_step_driver_graph_replay(cvb):
    """SURVEY.md §8(f)-1, cycle.CycleStep: one optimisation step (10 GRU_RNN passes, losses, BPTT, Adam) replayed as a CUDA
    graph.  (i) a replay draws FRESH dropout masks / latent noise (device-resident Philox state) and applies the right Adam
    bias correction (device-resident step count); (ii) from the same generator state and parameters the replayed graph and
    the kernel-by-kernel launch give bit-identical losses and parameters; (iii) gradients accumulated by the kernels
    straight into the flat buffer equal autograd's own accumulation."""
    from cyclevae_vc_b200 import cycle, synth
    lat, stdim, B, T, n_cyc = 32, 4, 6, 20, 2
    dev = torch.device("cuda")

    def build():
        enc, dec, y0d1 = synth.build_models(1024, lat, 2, 50, stdim, seed=1, device=dev)
        enc.train(); dec.train()
        return enc, dec, y0d1

    x, cv, sc, tc = (t.to(dev) for t in synth.make_batch(B, T, 5, 2, 50))
    y0e = torch.zeros(B, 1, 2 * lat, device=dev)
    runs = {}
    for mode in ("graph", "eager"):
        enc, dec, y0d1 = build()
        y0d = y0d1.to(dev).repeat(B, 1, 1).contiguous()
        opt = cycle.FlatAdam(cycle.trainable_parameters(enc, dec), lr=1e-3)
        torch.manual_seed(77)
        cs = cycle.CycleStep(enc, dec, opt, B=B, T=T, n_cyc=n_cyc, lat_dim=lat, stdim=stdim, n_spk=2, y0_enc=y0e, y0_dec=y0d,
                             graph=(mode == "graph"))
        assert (cs.graph is not None) == (mode == "graph")
        cs.rng.state[1:3] = 0                      # the graph's warm-up steps consumed counters: same start for both modes
        losses = []
        for _ in range(3):
            losses.append(float(cs.step(x, cv, sc, tc).item()))
        torch.cuda.synchronize()
        runs[mode] = (losses, opt.flat.clone(), cs.rng.state.clone())
    lg, pg, sg = runs["graph"]
    le, pe, se = runs["eager"]
    assert len(set(lg)) == 3, "every replay must draw new masks / noise"
    assert lg == le and torch.equal(pg, pe) and torch.equal(sg, se)
    assert int(sg[2]) == 3 and int(sg[1]) > 0
    # (iii) sink gradients == autograd-accumulated gradients on identical masks / noise
    enc_a, dec_a, y0d1 = build()
    enc_b, dec_b, _ = build()
    y0d = y0d1.to(dev).repeat(B, 1, 1).contiguous()
    for m in (enc_a, dec_a, enc_b, dec_b):
        for k, p in m.named_parameters():
            p.requires_grad_(not k.startswith("scale_"))
    opt = cycle.FlatAdam(cycle.trainable_parameters(enc_b, dec_b), lr=1e-3)
    eps = [[e.to(dev) for e in ec] for ec in orc.synth_noise(B, T, lat, n_cyc, 9)]
    masks = [[(a.to(dev), b.to(dev)) for a, b in mc] for mc in orc.synth_masks(B, T, orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024), n_cyc, 9)]
    grads = []
    for enc, dec in ((enc_a, dec_a), (enc_b, dec_b)):
        if enc is enc_b:
            opt.zero_grad()
        out, _ = cycle.cyc_forward(enc, dec, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=n_cyc, lat_dim=lat, stdim=stdim, y0_enc=y0e,
                                   y0_dec=y0d, do=True, eps=eps, masks=masks)
        loss, _ = cycle.cyc_loss(out, x, n_cyc=n_cyc, lat_dim=lat, stdim=stdim, flen_acc=[T] * B, select_utt_idx=list(range(B)))
        loss.backward()
        grads.append([p.grad.clone() for p in cycle.trainable_parameters(enc, dec)])
    for ga, gb in zip(*grads):
        assert _maxabs(ga, gb) <= 2e-6 * max(1e-3, float(ga.abs().max()))


def test_device_stager_side_stream(cvb):
    """features.DeviceStager with a side copy stream: the consumer stream must see complete batches (it waits on the copy
    event) even when it is busy, and alternating pinned slots must not be overwritten while their copies are in flight."""
    from cyclevae_vc_b200.features import DeviceStager
    dev = torch.device("cuda")
    st = DeviceStager(dev, stream=torch.cuda.Stream())
    busy = torch.randn(4096, 4096, device=dev)
    ok = True
    for it in range(6):
        batch = {"h_src": torch.full((8, 2200, 54), float(it)), "flen_src": torch.full((8,), it), "featfile_src": ["a"] * 8,
                 "spcidx_src": torch.full((8, 2200), it, dtype=torch.int64)}
        busy = busy @ busy * 1e-4          # keep the consumer stream occupied while the copies run on the side stream
        d = st.put(batch)
        ok = ok and bool((d["h_src"] == float(it)).all()) and bool((d["spcidx_src"] == it).all()) and d["h_src"].is_cuda
        assert d["featfile_src"] == ["a"] * 8 and not d["flen_src"].is_cuda
    torch.cuda.synchronize()
    assert ok


def test_reference_caller_code_through_the_dropin(golden_dir, cvb, tmp_path):
    """The reference's OWN caller code -- the frame-chunk branch of the training loop (train_*.py:1293-1474: 5 x n_cyc
    GRU_RNN passes with carried, detached state, per-utterance loss assembly, its torch.optim.Adam) with train_generator
    and a save_checkpoint round trip (.cpu() / torch.save / .cuda(), :152-167) after the second chunk, and the conversion
    block of decode_*.py:302-323 -- cut out of the scripts' ASTs (tests/ref_callers.py) and exec'd AS WRITTEN against the
    drop-in module on the GPU.  Fixture: the same code with the reference's own classes on the CPU."""
    from tests import ref_callers as rc
    if rc.script_path(rc.TRAINER) is None:
        pytest.skip("oracle/_ref/ (vendored by __graft_entry__.build() where /root/reference exists) is not present")
    import cyclevae_vc_b200.dropin.gru_vae as mod       # what `import gru_vae` resolves to with the drop-in on PYTHONPATH
    g = _load(golden_dir, "callers.npz")
    lat, stdim, n_cyc = 32, 4, 2
    mean, std = orc.synth_stats(50)
    enc = orc.NetSpec(54, 2 * lat, 1024, 3, 2, 0.0, True, False)
    dec = orc.NetSpec(lat + 2, 50, 1024, 3, 2, 0.0, False, True)
    Pe = orc.init_params(enc, 401, mean=mean, scale=std)
    Pd = orc.init_params(dec, 402, mean=mean[stdim:], scale=std[stdim:])
    dev = torch.device("cuda")

    def build():
        out = []
        for spec, P in ((enc, Pe), (dec, Pd)):
            m = mod.GRU_RNN(in_dim=spec.in_dim, out_dim=spec.out_dim, hidden_units=1024, do_prob=0.0, scale_in_flag=spec.scale_in,
                            scale_out_flag=spec.scale_out)
            m.load_state_dict({k: v.clone() for k, v in P.items()})
            for k, p in m.named_parameters():
                p.requires_grad_(not k.startswith("scale_"))
            out.append(m.cuda())
        return out

    me, md = build()
    me.train(); md.train()
    flens = g["flens"].tolist()
    x, cv, sc, tc = orc.synth_batch(2, 200, 31)
    y0d1 = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1)
    r = rc.run_trainer_chunks(mod, me, md, dev, x=x, cv=cv, sc=sc, tc=tc, flens=flens, lat_dim=lat, n_cyc=n_cyc, y0_dec=y0d1,
                              sample=rc.seeded_sampler(mod, lat, 5, dev, native=True), checkpoint_dir=str(tmp_path), checkpoint_after=2)
    assert _paths() == (PATH_TC, PATH_TC)
    assert r["iter_count"] == int(g["iter_count"]) == 3
    ref_l = g["losses"].tolist()
    assert r["losses"][0] == pytest.approx(ref_l[0], rel=1e-5)          # before any update: pure forward / loss parity
    for a, b in zip(r["losses"][1:], ref_l[1:]):                         # after the trainer's own Adam steps + checkpoint round trip
        assert a == pytest.approx(b, rel=2e-4)
    assert _maxabs(r["trj"][:, ::4], g["trj"]) < 5e-3 and _maxabs(r["h_dec"][:, :, ::8], g["h_dec"]) < 5e-3
    # the checkpoint the reference's save_checkpoint wrote from the drop-in modules loads back into fresh ones
    ck = torch.load(str(tmp_path / "checkpoint-2.pkl"), weights_only=False)
    assert set(ck) == {"model_encoder", "model_decoder", "optimizer", "numpy_random_state", "torch_random_state", "iterations"}
    m2, d2 = build()
    m2.load_state_dict(ck["model_encoder"])
    d2.load_state_dict(ck["model_decoder"])
    assert all(not v.is_cuda for v in ck["model_encoder"].values())
    # stage-6 conversion block, as written, on untrained copies
    me, md = build()
    me.eval(); md.eval()
    f, _, _, _ = orc.synth_batch(2, 120, 33)
    d = rc.run_decoder_block(me, md, dev, feat=f[0].numpy(), feat_trg=f[1].numpy(), lat_dim=lat, n_smpl=300, y0_dec=y0d1,
                             sample=rc.seeded_sampler(mod, lat, 6, dev, native=True))
    assert _paths()[0] == PATH_TC_FOLDED
    for k in ("cvmcep", "cvmcep_src", "cvmcep_trg"):
        assert d[k].dtype == np.float64 and _maxabs(d[k][::3].astype(np.float32), g["dec_" + k]) < TOL, k


def test_device_mcd_and_dtw_metrics(cvb):
    """the evaluation metrics of the training loop on the device against oracle/dtw_oracle.py
    (dtw_c itself is not in the reference tree: parity unpinned, the published definitions are what is checked)."""
    from cyclevae_vc_b200 import cycle
    from oracle import dtw_oracle as dto
    rng = np.random.default_rng(0)
    x, y = rng.normal(size=(300, 50)).astype(np.float32), rng.normal(size=(300, 50)).astype(np.float32)
    m = cycle.mcd_aligned(torch.tensor(x).cuda(), torch.tensor(y).cuda()).cpu().numpy()
    rm, rs = dto.calc_mcd(x, y)
    assert abs(m[0] - rm) < 1e-4 * rm and abs(m[1] - rs) < 1e-3 * rs
    idx = np.sort(rng.choice(300, size=120, replace=False))
    big = torch.tensor(np.c_[np.zeros((300, 4), np.float32), x]).cuda()            # [T, 54]: the call sites slice [:, stdim:]
    m = cycle.mcd_aligned(big[:, 4:], torch.tensor(y).cuda(), torch.tensor(idx).cuda(), torch.tensor(idx).cuda()).cpu().numpy()
    assert abs(m[0] - dto.calc_mcd(x[idx], y[idx])[0]) < 1e-4 * rm
    # DTW: a time-warped, noisy copy must be aligned back; cost and path against the numpy dynamic programme
    N, M = 173, 141
    trg = np.cumsum(rng.normal(size=(M, 50)).astype(np.float32) * 0.3, axis=0)
    warp = np.clip(np.round(np.linspace(0, M - 1, N) + rng.normal(size=N) * 1.5), 0, M - 1).astype(int)
    warp.sort()
    org = trg[warp] + 0.02 * rng.normal(size=(N, 50)).astype(np.float32)
    al, path, st = cycle.dtw_org_to_trg(torch.tensor(org).cuda(), torch.tensor(trg).cuda())
    path, st = path.cpu().numpy(), st.cpu().numpy()
    ral, rpath, rmean, rsteps, rcost = dto.dtw_org_to_trg(org, trg)
    assert path[0] >= 0 and path[-1] == N - 1 and (np.diff(path) >= 0).all()
    assert st[2] == pytest.approx(rcost, rel=1e-4) and st[0] == pytest.approx(rmean, rel=1e-3)
    assert (path == rpath).mean() > 0.98 and int(st[1]) == rsteps
    assert _maxabs(al, org[path]) == 0.0
    # exact case (integer-valued distances: no rounding differences, ties resolved in the same order)
    a = rng.integers(0, 4, size=(40, 2)).astype(np.float32)
    b = rng.integers(0, 4, size=(33, 2)).astype(np.float32)
    _, p2, s2 = cycle.dtw_org_to_trg(torch.tensor(a).cuda(), torch.tensor(b).cuda())
    _, rp2, _, rsteps2, _ = dto.dtw_org_to_trg(a, b)
    assert (p2.cpu().numpy() == rp2).all() and int(s2[1]) == rsteps2


@pytest.mark.parametrize("in_dim,B,T", [(7, 3, 11), (54, 80, 80)])
def test_two_sided_dil_conv_stand_alone(cvb, in_dim, B, T):
    """TwoSidedDilConv1d.forward on its own (gru_vae.py:53-66: x [B,C,T] -> [B,9C,T]) and its gradients against the same
    stack of torch Conv1d layers evaluated in float64 on the CPU; the small shape runs the per-tap products, the bench shape
    the composed / tap-fused tensor-core products."""
    torch.manual_seed(3)
    conv = cvb.TwoSidedDilConv1d(in_dim=in_dim, kernel_size=3, layers=2)
    conv.apply(cvb.initialize)
    for c in conv.conv:
        torch.nn.init.normal_(c.bias, std=0.1)
    ref = [torch.nn.Conv1d(c.in_channels, c.out_channels, 3, dilation=c.dilation, padding=c.padding).double() for c in conv.conv]
    for r, c in zip(ref, conv.conv):
        r.weight.data.copy_(c.weight.data.double())
        r.bias.data.copy_(c.bias.data.double())
    conv = conv.cuda()
    x = torch.randn(B, in_dim, T)
    w = torch.randn(B, in_dim * 9, T)
    xr = x.double().requires_grad_(True)
    yr = ref[1](ref[0](xr))
    (yr * w.double()).sum().backward()
    xd = x.cuda().requires_grad_(True)
    yd = conv(xd)
    assert yd.shape == (B, in_dim * 9, T)
    (yd * w.cuda()).sum().backward()
    scale = max(1.0, float(yr.abs().max()))
    assert _maxabs(yd.double(), yr.detach()) < TOL * scale
    assert _maxabs(xd.grad.double(), xr.grad) < TOL * max(1.0, float(xr.grad.abs().max()))
    for r, c in zip(ref, conv.conv):
        assert _maxabs(c.weight.grad.double(), r.weight.grad) < 2e-4 * max(1.0, float(r.weight.grad.abs().max()))
        assert _maxabs(c.bias.grad.double(), r.bias.grad) < 2e-4 * max(1.0, float(r.bias.grad.abs().max()))


def test_concat_features_matches_torch_cat(cvb):
    """The encoder-input concatenations of the trainer (train_*.py:1304,1307) through cvb_concat2_fwd: a strided slice as the
    first part (x[:, :, :stdim]), values and the split gradient bit-exact against torch.cat."""
    from cyclevae_vc_b200 import gru_vae as gv
    torch.manual_seed(5)
    x = torch.randn(6, 13, 54, device="cuda")
    rec = torch.randn(6, 13, 50, device="cuda", requires_grad=True)
    a = x[:, :, :4]
    out = gv.concat_features(a, rec)
    ref = torch.cat((a, rec.detach()), 2)
    assert torch.equal(out.detach(), ref)
    g = torch.randn_like(out)
    out.backward(g)
    assert torch.equal(rec.grad, g[:, :, 4:])
    cv = torch.randn(6, 13, 4, device="cuda", requires_grad=True)
    out2 = gv.concat_features(cv, rec.detach())
    out2.backward(g)
    assert torch.equal(out2.detach(), torch.cat((cv.detach(), rec.detach()), 2)) and torch.equal(cv.grad, g[:, :, :4])
