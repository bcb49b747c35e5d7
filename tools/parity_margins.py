"""Max-abs distance of the CUDA path from the reference-generated fixtures (hu1024, T=800 decode and carried chunks) for the
tensor-core and the fp32-FMA recurrence kernels / GEMM back-ends: prints the margins behind the 1e-4 parity bar (GPU box only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cyclevae_vc_b200 as cvb  # noqa: E402
from cyclevae_vc_b200 import cycle  # noqa: E402
from oracle import gru_vae_oracle as orc  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "flagship.npz"))


def module(spec, P):
    m = cvb.GRU_RNN(in_dim=spec.in_dim, out_dim=spec.out_dim, hidden_units=spec.hidden_units, kernel_size=spec.kernel_size,
                    dilation_size=spec.dilation_size, do_prob=spec.do_prob, scale_in_flag=spec.scale_in, scale_out_flag=spec.scale_out)
    m.load_state_dict({k: v.clone() for k, v in P.items()})
    return m.cuda().eval()


def run(tag, gain, bstd):
    lat, stdim = 32, 4
    mean, std = orc.synth_stats(50)
    enc, dec = orc.encoder_spec(54, lat, 1024), orc.decoder_spec(lat, 2, 50, 1024)
    Pe = orc.init_params(enc, 201, gain=gain, bias_std=bstd, mean=mean, scale=std)
    Pd = orc.init_params(dec, 202, gain=gain, bias_std=bstd, mean=mean[stdim:], scale=std[stdim:])
    me, md = module(enc, Pe), module(dec, Pd)
    y0d1 = torch.tensor((0 - mean[stdim:]) / std[stdim:], dtype=torch.float32).reshape(1, 1, -1).cuda()
    T = 800
    x, _, sc, tc = orc.synth_batch(1, T, 1)
    eps_mean = (orc.synth_noise(1, T, lat, 1, 1)[0][0] / np.sqrt(300.0)).cuda()
    with torch.no_grad():
        lat_src, _, _ = me(x[0].cuda(), torch.zeros(1, 1, 2 * lat).cuda(), clamp_vae=True, lat_dim=lat)
        cvm = cycle.convert(me, md, x[0].cuda(), tc[0].cuda(), lat_dim=lat, y0_enc=torch.zeros(1, 1, 2 * lat).cuda(), y0_dec=y0d1,
                            eps_mean=eps_mean[0])
    e_lat = np.abs(lat_src[::5].cpu().numpy() - g[f"{tag}/dec800_lat"]).max()
    e_cvm = np.abs(cvm[::5].cpu().numpy() - g[f"{tag}/dec800_cvmcep"]).max()
    return e_lat, e_cvm, float(np.abs(g[f"{tag}/dec800_cvmcep"]).max())


for rec in ("exact", "tc"):
    for gemm in ("cublas", "tc"):
        if rec == "exact":
            os.environ["CVB_RECURRENCE"] = "exact"
        else:
            os.environ.pop("CVB_RECURRENCE", None)
        if gemm == "cublas":
            os.environ["CVB_GEMM"] = "cublas"
        else:
            os.environ.pop("CVB_GEMM", None)
        for tag, gain, bstd in (("init", 1.0, 0.0), ("trained", 3.0, 0.05)):
            e_lat, e_cvm, sc = run(tag, gain, bstd)
            print(f"recurrence={rec:5s} gemm={gemm:6s} {tag:8s}: lat max-abs {e_lat:.2e}   converted mcep max-abs {e_cvm:.2e} (|mcep| max {sc:.1f})")
