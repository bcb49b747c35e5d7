"""Per-term error budget of the split-precision tensor-core recurrence (CPU emulation, no GPU needed).

The tcgen05 recurrence kernels form W h as  h_hi W_hi + h_hi W_lo + h_lo W_hi  with 16-bit hi/lo operands and fp32
accumulation in TMEM.  On the "trained-like" stress set (weights x3) round 1 measured the tensor-core path 2.3x further
from the float64 oracle than the fp32 reference itself.  This tool separates the candidate causes by emulating the
decoder's folded inference recurrence (gru_tc_eval.cu arithmetic: 4 K-slices of 256, chains of 16 accumulating MMAs of
K = 16, partial sums added in fp32) on the CPU, one term switched at a time, over the same 800 frames as the fixture
`trained/dec800_cvmcep`:

    lo planes    unscaled fp16 residual (subnormal below 6.1e-5: absolute resolution 2^-24)  vs  residual x 2^11
    chain        accumulation inside a tcgen05 chain truncates (measured growth ~2.2e-7 per K=16 step on a long chain);
                 modelled as round-toward-zero to fp32 after each of `trunc` partial adds per MMA; chains of 16 / 8 / 4
    lo x lo      the dropped fourth product

Every variant sees the same gx (float64 front-end, rounded to fp32) and differs only in how  W h  is formed; "fp32" is
the reference's own arithmetic (torch CPU float32 matvec).  Output: max-abs distance of the de-normalised mcep from
the float64 run.  Usage:  python tools/split_error_budget.py [T]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gru_vae_oracle as orc  # noqa: E402

H, OUT, LAT = 1024, 50, 32


def rz32(x):
    """float64 -> float32 rounded toward zero (returned as float64)."""
    r = x.astype(np.float32)
    big = np.abs(r.astype(np.float64)) > np.abs(x)
    r[big] = np.nextafter(r[big], np.float32(0))
    return r.astype(np.float64)


def rn32(x):
    return x.astype(np.float32).astype(np.float64)


def split16(x, scaled):
    """x (float32 values in float64) -> (hi, lo) as float64 holding fp16-representable values; lo is the TRUE value of
    the stored residual (i.e. already divided by the scale)."""
    hi = x.astype(np.float16).astype(np.float64)
    r = x - hi
    if scaled:
        lo = (r * 2048.0).astype(np.float16).astype(np.float64) / 2048.0
    else:
        lo = r.astype(np.float16).astype(np.float64)
    return hi, lo


def matvec_emulated(Wh, Wl, h, *, scaled, chain, trunc, lolo, lohi_corr=None):
    """One W h of the folded kernel: rows [R], K = H in 4 slices of 256; per slice chains of `chain` K=16 MMAs (0 = exact
    accumulation), partial sums of the slices and of the (hi x hi | corrections) halves added in fp32."""
    hh, hl = split16(h, scaled)
    if lohi_corr is None:
        lohi_corr = scaled
    R = Wh.shape[0]
    nb = H // 16
    W3h, W3l = Wh.reshape(R, nb, 16), Wl.reshape(R, nb, 16)
    h3h, h3l = hh.reshape(nb, 16), hl.reshape(nb, 16)
    p_hh = np.einsum("rbk,bk->rb", W3h, h3h)          # exact products / block sums in float64
    p_hl = np.einsum("rbk,bk->rb", W3l, h3h)          # h_hi W_lo
    p_lh = np.einsum("rbk,bk->rb", W3h, h3l)          # h_lo W_hi
    p_ll = np.einsum("rbk,bk->rb", W3l, h3l)
    total = np.zeros(R)
    for s in range(4):
        blocks = range(s * 16, (s + 1) * 16)
        if chain == 0:
            main = p_hh[:, blocks].sum(1) + p_lh[:, blocks].sum(1)
            corr = p_hl[:, blocks].sum(1)
            if lolo:
                corr = corr + p_ll[:, blocks].sum(1)
            part = rn32(rn32(main) + rn32(corr))
        else:
            part = np.zeros(R)
            for c0 in range(0, 16, chain):
                main = np.zeros(R)
                corr = np.zeros(R)
                for b in list(blocks)[c0:c0 + chain]:
                    if lohi_corr:   # both corrections in their own accumulator (scaled lo planes: truncation there is x 2^-11)
                        main = rz32(main + p_hh[:, b]) if trunc else rn32(main + p_hh[:, b])
                        corr = corr + p_hl[:, b] + p_lh[:, b] if scaled else rz32(rz32(corr + p_hl[:, b]) + p_lh[:, b])
                    else:        # round 1: h_lo W_hi is accumulated onto the main columns by a second MMA
                        main = rz32(main + p_hh[:, b]) if trunc else rn32(main + p_hh[:, b])
                        main = rz32(main + p_lh[:, b]) if trunc else rn32(main + p_lh[:, b])
                        corr = corr + p_hl[:, b]
                    if trunc > 1:   # a second truncation per MMA (the hardware adds the K=16 products in two groups)
                        main = rz32(main * (1.0 + 0.0))
                if lolo:
                    corr = corr + p_ll[:, list(blocks)[c0:c0 + chain]].sum(1)
                part = rn32(part + rn32(rn32(main) + rn32(corr)))
        total = rn32(total + part)
    return total


def run(T=800):
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, LAT, H), orc.decoder_spec(LAT, 2, OUT, H)
    Pe = orc.init_params(enc, 201, gain=3.0, bias_std=0.05, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, gain=3.0, bias_std=0.05, mean=mean[4:], scale=std[4:])
    x, _, sc, tc = orc.synth_batch(1, T, 1)
    eps_mean = (orc.synth_noise(1, T, LAT, 1, 1)[0][0] / np.sqrt(300.0))[0].double()
    P64e, P64d = ({k: v.double() for k, v in P.items()} for P in (Pe, Pd))
    y0d = torch.tensor((0 - mean[4:]) / std[4:]).reshape(1, 1, -1)
    with torch.no_grad():
        lat, _, _ = orc.gru_rnn_forward(P64e, enc, x[0].double(), torch.zeros(1, 1, 2 * LAT, dtype=torch.float64), clamp_vae=True, lat_dim=LAT)
        z = torch.cat((tc[0].double(), orc.sampling_vae_batch(lat, eps_mean, LAT)), 1)
        xc = orc.frontend(P64d, dec, z.unsqueeze(0))[0]                       # [T, C] float64
    C = dec.conv_dim
    W_ih, W_hh = P64d["gru.weight_ih_l0"].numpy(), P64d["gru.weight_hh_l0"].numpy()
    b_ih, b_hh = P64d["gru.bias_ih_l0"].numpy(), P64d["gru.bias_hh_l0"].numpy()
    W_o, b_o = P64d["out_1.weight"][:, :, 0].numpy(), P64d["out_1.bias"].numpy()
    W_x, W_y = W_ih[:, :C], W_ih[:, C:]
    so_w, so_b = np.diag(P64d["scale_out.weight"][:, :, 0].numpy()), P64d["scale_out.bias"].numpy()
    gx64 = xc.numpy() @ W_x.T + b_ih
    y_in = y0d.numpy()[0, 0]

    def recur(mode, **kw):
        """mode 'f64' | 'f32' (reference arithmetic: unfolded, float32) | 'tc' (folded, emulated split products)."""
        if mode == "f64":
            h, y, ys = np.zeros(H), y_in.copy(), []
            for t in range(T):
                gi = gx64[t] + W_y @ y
                gh = W_hh @ h + b_hh
                r = 1 / (1 + np.exp(-(gi[:H] + gh[:H])))
                zt = 1 / (1 + np.exp(-(gi[H:2 * H] + gh[H:2 * H])))
                n = np.tanh(gi[2 * H:] + r * gh[2 * H:])
                h = (1 - zt) * n + zt * h
                y = W_o @ h + b_o
                ys.append(y)
            return np.array(ys) * so_w + so_b
        if mode == "f32":
            f = np.float32
            Wy, Whh, Wo = W_y.astype(f), W_hh.astype(f), W_o.astype(f)
            gx, bhh, bo = gx64.astype(f), b_hh.astype(f), b_o.astype(f)
            h, y, ys = np.zeros(H, f), y_in.astype(f), []
            for t in range(T):
                gi = gx[t] + Wy @ y
                gh = Whh @ h + bhh
                r = f(1) / (f(1) + np.exp(-(gi[:H] + gh[:H])))
                zt = f(1) / (f(1) + np.exp(-(gi[H:2 * H] + gh[H:2 * H])))
                n = np.tanh(gi[2 * H:] + r * gh[2 * H:])
                h = (f(1) - zt) * n + zt * h
                y = Wo @ h + bo
                ys.append(y.astype(np.float64))
            return np.array(ys) * so_w + so_b
        # folded kernel (gru_tc_eval.cu): W_fb = W_y W_o (float64 product rounded to fp32), rows [r' | z' | hn | in']
        W_fb = rn32(rn32(W_y) @ rn32(W_o))
        c_fb = rn32(rn32(W_y) @ rn32(b_o))
        Whh32 = rn32(W_hh)
        W4 = np.concatenate((rn32(Whh32[:H] + W_fb[:H]), rn32(Whh32[H:2 * H] + W_fb[H:2 * H]), Whh32[2 * H:], W_fb[2 * H:]), 0)
        Wh, Wl = split16(W4, kw["scaled"])
        gx = rn32(gx64 + c_fb)
        bhh = rn32(b_hh)
        h, ys = np.zeros(H), []
        Wo32, bo32 = rn32(W_o), rn32(b_o)
        for t in range(T):
            a = matvec_emulated(Wh, Wl, h, **kw)
            g0 = gx[t].copy()
            if t == 0:   # first step: the caller's y_in instead of W_o h_in + b_o
                g0 = rn32(g0 + rn32(W_y) @ rn32(y_in) - c_fb)
            ar = rn32(g0[:H] + a[:H] + bhh[:H])
            az = rn32(g0[H:2 * H] + a[H:2 * H] + bhh[H:2 * H])
            ghn = rn32(a[2 * H:3 * H] + bhh[2 * H:])
            r = rn32(1 / (1 + np.exp(-ar)))
            zt = rn32(1 / (1 + np.exp(-az)))
            n = rn32(np.tanh(rn32(g0[2 * H:] + a[3 * H:] + r * ghn)))
            h = rn32((1 - zt) * n + zt * h)
            ys.append(rn32(Wo32 @ h + bo32))
        return np.array(ys) * so_w + so_b

    exact = recur("f64")
    rows = [("fp32 reference arithmetic (unfolded)", recur("f32"))]
    variants = [
        ("round 1: unscaled lo, chain 16 truncating, lo.hi on main", dict(scaled=False, chain=16, trunc=1, lolo=False)),
        ("unscaled lo, exact accumulation", dict(scaled=False, chain=0, trunc=0, lolo=False)),
        ("scaled lo (x2^11), exact accumulation", dict(scaled=True, chain=0, trunc=0, lolo=False)),
        ("scaled lo, exact accumulation, + lo.lo", dict(scaled=True, chain=0, trunc=0, lolo=True)),
        ("scaled lo, chain 16 truncating", dict(scaled=True, chain=16, trunc=1, lolo=False)),
        ("scaled lo, chain 8 truncating", dict(scaled=True, chain=8, trunc=1, lolo=False)),
        ("scaled lo, chain 4 truncating", dict(scaled=True, chain=4, trunc=1, lolo=False)),
        ("scaled lo, chain 16 round-to-nearest", dict(scaled=True, chain=16, trunc=0, lolo=False)),
        ("unscaled lo, chain 16 truncating, lo.hi on corrections", dict(scaled=False, chain=16, trunc=1, lolo=False, lohi_corr=True)),
        ("unscaled lo, chain 8 truncating, lo.hi on corrections", dict(scaled=False, chain=8, trunc=1, lolo=False, lohi_corr=True)),
        ("unscaled lo, chain 8 truncating, lo.hi on main", dict(scaled=False, chain=8, trunc=1, lolo=False)),
    ]
    for name, kw in variants:
        rows.append((name, recur("tc", **kw)))
    print(f"decoder hu1024, trained-like weights (x3), T={T}, |mcep| max {np.abs(exact).max():.1f}; max-abs distance from float64:")
    ref_err = None
    for name, ys in rows:
        e = np.abs(ys - exact).max()
        if ref_err is None:
            ref_err = e
        print(f"  {name:58s} {e:.3e}   ({e / ref_err:.2f} x the fp32 reference)")


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 800)
