"""Which piece of the step invalidates a CUDA-graph capture?  (GPU box only; debugging aid)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclevae_vc_b200 import cycle, synth  # noqa: E402
from cyclevae_vc_b200 import gru_vae as gv  # noqa: E402

dev = torch.device("cuda")
lat, stdim, B, T = 32, 4, 6, 20
enc, dec, y0d1 = synth.build_models(1024, lat, 2, 50, stdim, seed=1, device=dev)
enc.train(); dec.train()
x, cv, sc, tc = (t.to(dev) for t in synth.make_batch(B, T, 5, 2, 50))
y0e = torch.zeros(B, 1, 2 * lat, device=dev)
y0d = y0d1.to(dev).repeat(B, 1, 1).contiguous()
rng = gv.DeviceRng(dev)
opt = cycle.FlatAdam(cycle.trainable_parameters(enc, dec), lr=1e-3)
flens = torch.full((B,), T, dtype=torch.int32, device=dev)


def try_capture(name, fn):
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        print(f"OK    {name}")
    except Exception as e:  # noqa: BLE001
        print(f"FAIL  {name}: {str(e).splitlines()[0][:120]}")
        torch.cuda.synchronize()


def f_masks():
    with gv.device_rng(rng):
        gv.draw_dropout_masks(B, T, 486, 1024, 0.5, dev)
    rng.end_step()


def f_enc_eval():
    with torch.no_grad():
        enc.eval()
        enc(x, y0e, clamp_vae=True, lat_dim=lat)
        enc.train()


def f_enc_train_fwd():
    with gv.device_rng(rng):
        enc(x, y0e, clamp_vae=True, lat_dim=lat, do=True)
    rng.end_step()


def f_enc_train_fwd_bwd():
    opt.zero_grad()
    with gv.device_rng(rng):
        o, _, _ = enc(x, y0e, clamp_vae=True, lat_dim=lat, do=True)
        o.sum().backward()
    rng.end_step()


def f_losses():
    lat_t = torch.randn(B, T, 2 * lat, device=dev, requires_grad=True)
    gv.kl_per_utt(lat_t, flens, lat).sum().backward()


def f_full():
    opt.zero_grad()
    with gv.device_rng(rng):
        out, _ = cycle.cyc_forward(enc, dec, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=2, lat_dim=lat, stdim=stdim, y0_enc=y0e, y0_dec=y0d, do=True)
        loss, _ = cycle.cyc_loss(out, x, n_cyc=2, lat_dim=lat, stdim=stdim, flen_acc=None, select_utt_idx=list(range(B)), flens_dev=flens)
        loss.backward()
    rng.end_step()


for nm, fn in (("masks", f_masks), ("enc eval", f_enc_eval), ("enc train fwd", f_enc_train_fwd), ("enc train fwd+bwd", f_enc_train_fwd_bwd),
               ("losses", f_losses), ("full step", f_full)):
    try_capture(nm, fn)


zx, zcv, zsc, ztc = (torch.zeros_like(t) for t in (x, cv, sc, tc))
loss_buf = torch.zeros((), device=dev)


def mk_full(e, d, o, xx, cc, ss, tt, copy_loss):
    def fn():
        o.zero_grad()
        with gv.device_rng(rng):
            out, _ = cycle.cyc_forward(e, d, x=xx, cv=cc, src_code=ss, trg_code=tt, n_cyc=2, lat_dim=lat, stdim=stdim, y0_enc=y0e, y0_dec=y0d, do=True)
            loss, _ = cycle.cyc_loss(out, xx, n_cyc=2, lat_dim=lat, stdim=stdim, flen_acc=None, select_utt_idx=list(range(B)), flens_dev=flens)
            loss.backward()
        rng.end_step()
        if copy_loss:
            loss_buf.copy_(loss.detach())
    return fn


try_capture("full step, zero inputs", mk_full(enc, dec, opt, zx, zcv, zsc, ztc, False))
try_capture("full step + loss copy", mk_full(enc, dec, opt, x, cv, sc, tc, True))
enc3, dec3, _ = synth.build_models(1024, lat, 2, 50, stdim, seed=1, device=dev)
enc3.train(); dec3.train()
opt3 = cycle.FlatAdam(cycle.trainable_parameters(enc3, dec3), lr=1e-3)
try_capture("full step, fresh modules", mk_full(enc3, dec3, opt3, x, cv, sc, tc, False))
enc4, dec4, _ = synth.build_models(1024, lat, 2, 50, stdim, seed=1, device=dev)
enc4.train(); dec4.train()
opt4 = cycle.FlatAdam(cycle.trainable_parameters(enc4, dec4), lr=1e-3)
cs4 = cycle.CycleStep(enc4, dec4, opt4, B=B, T=T, n_cyc=2, lat_dim=lat, stdim=stdim, n_spk=2, y0_enc=y0e, y0_dec=y0d, graph=False)
try_capture("CycleStep._body", cs4._body)
cs4.x.copy_(x); cs4.cv.copy_(cv); cs4.sc.copy_(sc); cs4.tc.copy_(tc)
try_capture("CycleStep._body, real inputs", cs4._body)
