"""Forward parity of the tensor-core recurrence vs the fp32-FMA kernels over batch-row counts (GPU box only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cyclevae_vc_b200 as cvb  # noqa: E402

torch.manual_seed(0)
enc = cvb.GRU_RNN(in_dim=54, out_dim=64, hidden_units=1024, do_prob=0.5, scale_out_flag=False).cuda().eval()
enc.apply(cvb.initialize)
T = 40
for B in [int(a) for a in sys.argv[1:]] or [80, 86, 88, 96, 72, 43, 5]:
    x = torch.randn(B, T, 54, device="cuda")
    y0 = torch.zeros(B, 1, 64, device="cuda")
    with torch.no_grad():
        os.environ["CVB_RECURRENCE"] = "exact"
        o_e, _, h_e = enc(x, y0, clamp_vae=True, lat_dim=32)
        os.environ.pop("CVB_RECURRENCE")
        o_t, _, h_t = enc(x, y0, clamp_vae=True, lat_dim=32)
    torch.cuda.synchronize()
    d = (o_e - o_t).abs()
    bad_rows = (d.amax(dim=(1, 2)) > 1e-4).nonzero().flatten().tolist()
    print(f"B={B:3d}: max|out diff| {d.max().item():.2e}  max|h diff| {(h_e - h_t).abs().max().item():.2e}  finite {bool(torch.isfinite(o_t).all())}  bad rows {bad_rows[:12]}")
