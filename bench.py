#!/usr/bin/env python
"""Throughput of the GRU-VAE hot path in mcep frames/sec -- BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|decode|spk4] [--batch-utt B] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

Workloads (config.workload names the one that ran):
  train   configs[1]: SF1<->TF1 CycleVAE hu1024 ld32 ks3 ds2 cyc2 -- one optimisation step = 4 encoder + 6 decoder
          GRU_RNN passes forward (dropout + noise drawn on device), losses, BPTT, gradient all-reduce at N>1, Adam
          (train_*.py:1298-1420) on 80-frame chunks (--batch_size 80, train_*.py:70-134) x B utterances per GPU.
          Default B = 80 (the literal "bs80"); the recipe's own batch_size_utt values are --batch-utt 1 / 8.
  spk4    configs[3]: the same step with 4-speaker one-hot codes (decoder in_dim 36), default B = 8 per GPU.
  decode  configs[2]: stage-6 conversion (decode_*.py:303-305,318: ENC -> latent mean -> DEC) of 512 utterances x 800
          frames per GPU, eval mode; config.full_stage6 adds the 2 ENC + 3 DEC of decode_*.py:303-323.
Weak scaling: the per-GPU batch is fixed.

`--impl reference` times the reference's own CPU implementation of the same workload at the SAME batch size: the
unmodified src/nets/gru_vae.py (vendored by __graft_entry__.build() into the git-ignored oracle/_ref/, kind
"reference"; the CPU restatement oracle/gru_vae_oracle.py, kind "port", when that copy is absent) driven by the
trainer's / decoder's step composition restated from train_*.py:1298-1420 / decode_*.py:303-318 (the scripts themselves
need h5py / dtw_c / pysptk, absent here).  It never imports cyclevae_vc_b200.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HIDDEN, LAT, NMCEP, STDIM, NCYC, TCHUNK = 1024, 32, 50, 4, 2, 80
DEC_UTT, DEC_T, N_SMPL = 512, 800, 300
METRIC = "mcep frames/sec (cyc2 enc+dec fwd+bwd)"
METRIC_DECODE = "mcep frames/sec (stage-6 conversion, enc+dec fwd)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="train", choices=["train", "decode", "spk4"])
    ap.add_argument("--batch-utt", type=int, default=None, help="utterances per GPU (default: 80 train, 8 spk4, 512 decode)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels one by one instead of replaying the captured CUDA graph")
    a = ap.parse_args()
    if a.batch_utt is None:
        a.batch_utt = {"train": 80, "spk4": 8, "decode": DEC_UTT}[a.workload]
    a.n_spk = 4 if a.workload == "spk4" else 2
    return a


def config_of(args, world):
    """The `config` object -- identical in the native and the reference arm."""
    B = args.batch_utt
    if args.workload == "decode":
        return {"workload": f"stage-6 decode: batch conversion {B} utts x {DEC_T} frames x {NMCEP} mcep per GPU, hu{HIDDEN} ld{LAT}, "
                            f"1 ENC + 1 DEC forward, latent = mean of {N_SMPL} samples (configs[2])",
                "frames_per_step_per_gpu": B * DEC_T, "parallelism": f"dp{world} (utterance shards, no collective)"}
    name = "SF1<->TF1 CycleVAE" if args.workload == "train" else "4-speaker many-to-many CycleVAE (one-hot spk code, dec in 36)"
    return {"workload": f"{name} hu{HIDDEN} ld{LAT} ks3 ds2 cyc{NCYC}, {TCHUNK}-frame chunk x {B} utterances per GPU"
                        + (" (bs80, configs[1])" if args.workload == "train" and B == 80 else ""),
            "frames_per_step_per_gpu": B * TCHUNK, "passes_per_step": "4 ENC + 6 DEC fwd+bwd",
            "parallelism": f"dp{world}", "optimizer": "Adam (in timed region)"}


# ------------------------------------------------------------------------------------------------
# algorithmic work
def recurrence_flops_per_frame(out_dim, backward):
    fwd = 2 * 3 * HIDDEN * (HIDDEN + out_dim) + 2 * out_dim * HIDDEN          # W_hh h + W_y y ; W_o o
    if not backward:
        return fwd
    return 2 * 3 * HIDDEN * HIDDEN + 2 * 3 * HIDDEN * out_dim + 2 * out_dim * HIDDEN   # dgh W_hh ; dgi W_y ; dy W_o


def dense_flops_per_pass(in_dim, out_dim, B, T, backward):
    """Algorithmic FLOPs of the dense products of ONE GRU_RNN pass outside the recurrence kernels (k_gemm_tc): the composed
    9-tap conv product on the padded grid and gx = xc W_x^T forward; dW_hh, dW_ih, dxc, dW_o, dW_y and the two front-end
    products backward."""
    cl, rows, rp = 9 * in_dim, B * T, B * (T + 8)
    fwd = 2 * rp * cl * cl + 2 * rows * 3 * HIDDEN * cl
    if not backward:
        return fwd
    return (2 * 3 * HIDDEN * HIDDEN * rows + 2 * 3 * HIDDEN * cl * rows + 2 * rows * cl * 3 * HIDDEN + 2 * out_dim * HIDDEN * rows
            + 2 * 3 * HIDDEN * out_dim * rows + 2 * 2 * rp * cl * cl)


def folded_flops_per_frame():
    return 2 * 4 * HIDDEN * HIDDEN        # [r' | z' | hn | in'] rows of the folded inference recurrence (gru_tc_eval.cu)


def streaming_rooflines(enc, dec, opt, B, T, n_spk, dev, peak_hbm):
    """Achieved HBM GB/s of the path's streaming kernels (bytes = inputs + outputs once), each timed
    alone with CUDA events after an L2 flush (a 256 MB write), median of 9 launches, at the step's own sizes."""
    import torch
    from cyclevae_vc_b200 import gru_vae as gv
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    lat = torch.randn(B, T, 2 * LAT, device=dev)
    code = torch.zeros(B, T, n_spk, device=dev)
    trj = torch.randn(B, T, NMCEP, device=dev)
    x = torch.randn(B, T, STDIM + NMCEP, device=dev)
    flens = torch.full((B,), T, dtype=torch.int32, device=dev)
    conv_dim, frames = 9 * (STDIM + NMCEP), B * T
    cases = [
        ("k_adam (cvb_adam_step: Adam over the flat parameter buffer)", lambda: opt.step(), opt.n * 28.0,
         "read p, g, m, v + write p, m, v = 28 B per parameter"),
        ("k_dropout_mask (conv + GRU-output masks of one encoder pass)", lambda: gv.draw_dropout_masks(B, T, conv_dim, HIDDEN, 0.5, dev),
         frames * (conv_dim + HIDDEN) * 4.0, "4 B per mask element written"),
        ("k_reparam_concat_fwd (sampling_vae_batch + speaker-code concat, noise drawn in-kernel)",
         lambda: gv.reparam_concat(lat, code, None, LAT), frames * (2 * LAT + n_spk + (n_spk + LAT) + LAT) * 4.0,
         "read [mu|log-var] + code, write [code|z] + the stored noise"),
        ("k_kl_fwd (loss_vae per utterance)", lambda: gv.kl_per_utt(lat, flens, LAT), frames * 2 * LAT * 4.0, "256 B per frame read"),
        ("k_mcd_fwd (TWFSEloss L1 per utterance)", lambda: gv.mcd_l1_per_utt(trj, x, flens, 0, STDIM), frames * 2 * NMCEP * 4.0,
         "400 B per frame read"),
    ]
    out = []
    with torch.no_grad():
        for name, fn, nbytes, what in cases:
            fn()
            ts = []
            for _ in range(9):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            gbs = nbytes / (ms * 1e-3) / 1e9
            e = {"bound": "hbm", "kernel": name, "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                 "ms": ms, "algorithmic_bytes": nbytes, "bytes_rule": what}
            if nbytes / (peak_hbm * 1e9) * 1e3 < 0.002:   # a full-bandwidth pass would take less than 2 us: launch latency, not bandwidth
                e["note"] = (f"{nbytes / 1e6:.1f} MB at this batch: {nbytes / (peak_hbm * 1e9) * 1e6:.2f} us at the HBM peak -- the launch is "
                             "latency-bound at this size, not bandwidth-bound")
            out.append(e)
    return out


def frontend_bytes_per_frame(in_dim):
    return 4 * in_dim + 4 * 9 * in_dim    # read x, write xc (2160 B ENC, 1360 B DEC)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference module (oracle/_ref) or the port (oracle/gru_vae_oracle.py).  Nothing here imports
# cyclevae_vc_b200; the oracle is the thing TIMED here, never a product path.
def _load_reference_module():
    """The unmodified reference src/nets/gru_vae.py, vendored into oracle/_ref/ by build(); None when absent."""
    path = os.path.join(ROOT, "oracle", "_ref", "gru_vae.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("cyclevae_reference_gru_vae", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cpu_models(ref, n_spk):
    """Encoder / decoder as the trainer builds them (train_*.py:310-347) from the reference's own classes."""
    import torch
    from oracle import gru_vae_oracle as orc
    mean, std = orc.synth_stats(NMCEP)
    torch.manual_seed(1)
    enc = ref.GRU_RNN(in_dim=STDIM + NMCEP, out_dim=2 * LAT, hidden_units=HIDDEN, do_prob=0.5, scale_out_flag=False)
    dec = ref.GRU_RNN(in_dim=LAT + n_spk, out_dim=NMCEP, hidden_units=HIDDEN, do_prob=0.5, scale_in_flag=False)
    enc.apply(ref.initialize)
    dec.apply(ref.initialize)
    enc.scale_in.weight = torch.nn.Parameter(torch.diag(torch.tensor(1.0 / std, dtype=torch.float32)).unsqueeze(2))
    enc.scale_in.bias = torch.nn.Parameter(torch.tensor(-(mean / std), dtype=torch.float32))
    dec.scale_out.weight = torch.nn.Parameter(torch.diag(torch.tensor(std[STDIM:], dtype=torch.float32)).unsqueeze(2))
    dec.scale_out.bias = torch.nn.Parameter(torch.tensor(mean[STDIM:], dtype=torch.float32))
    for p in list(enc.scale_in.parameters()) + list(dec.scale_out.parameters()):
        p.requires_grad = False
    return enc, dec, mean, std


def cpu_train_step_factory(B, n_spk):
    """One cyc2 fwd+bwd+Adam step on the CPU at B utterances x 80 frames -> (step(), kind)."""
    import torch
    from oracle import gru_vae_oracle as orc
    ref = _load_reference_module()
    x, cv, sc, tc = orc.synth_batch(B, TCHUNK, 100, n_spk=n_spk)
    mean, std = orc.synth_stats(NMCEP)
    y0e = torch.zeros(B, 1, 2 * LAT)
    y0d = torch.tensor((0 - mean[STDIM:]) / std[STDIM:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)
    if ref is not None:
        # the reference's own modules, dropout drawn by their nn.Dropout, losses by its loss_vae / TWFSEloss in the
        # trainer's per-utterance loops (train_*.py:1326-1338, 1363-1410), torch.optim.Adam (:377)
        enc, dec, _, _ = _cpu_models(ref, n_spk)
        enc.train(); dec.train()
        crit = ref.TWFSEloss()
        train = [p for m in (enc, dec) for sub in (m.conv, m.gru, m.out_1) for p in sub.parameters()]
        opt = torch.optim.Adam(train, lr=1e-4)

        def samp(p):   # sampling_vae_batch (gru_vae.py:85-98) without its hard-coded .cuda()
            return p[:, :, :LAT] + torch.exp(p[:, :, LAT:] / 2) * torch.randn(B, TCHUNK, LAT)

        def step():
            rec = None
            total = 0.0
            for i in range(NCYC):
                ein = x if i == 0 else torch.cat((x[:, :, :STDIM], rec), 2)
                lat_src, _, _ = enc(ein, y0e, clamp_vae=True, lat_dim=LAT, do=True)
                trj_ss, _, _ = dec(torch.cat((sc, samp(lat_src)), 2), y0d, do=True)
                trj_st, _, _ = dec(torch.cat((tc, samp(lat_src)), 2), y0d, do=True)
                lat_st, _, _ = enc(torch.cat((cv, trj_st), 2), y0e, clamp_vae=True, lat_dim=LAT, do=True)
                rec, _, _ = dec(torch.cat((sc, samp(lat_st)), 2), y0d, do=True)
                for j in range(B):
                    total = total + crit(trj_ss[j], x[j, :, STDIM:], L2=False, GV=False)[1] + crit(rec[j], x[j, :, STDIM:], L2=False, GV=False)[1] \
                        + ref.loss_vae(lat_src[j], lat_dim=LAT) + ref.loss_vae(lat_st[j], lat_dim=LAT)
            opt.zero_grad()
            total.backward()
            opt.step()
            return float(total.detach())

        return step, "reference"
    enc_s, dec_s = orc.encoder_spec(STDIM + NMCEP, LAT, HIDDEN), orc.decoder_spec(LAT, n_spk, NMCEP, HIDDEN)
    Pe = orc.init_params(enc_s, 1, mean=mean, scale=std)
    Pd = orc.init_params(dec_s, 2, mean=mean[STDIM:], scale=std[STDIM:])
    train = []
    for P in (Pe, Pd):
        for k, v in P.items():
            if not k.startswith("scale_"):
                v.requires_grad_(True)
                train.append(v)
    opt = torch.optim.Adam(train, lr=1e-4)

    def step():
        eps = [[torch.randn(B, TCHUNK, LAT) for _ in range(3)] for _ in range(NCYC)]
        masks = [[((torch.rand(B, TCHUNK, s.conv_dim) >= 0.5).float() * 2, (torch.rand(B, TCHUNK, HIDDEN) >= 0.5).float() * 2)
                  for s in (enc_s, dec_s, dec_s, enc_s, dec_s)] for _ in range(NCYC)]
        opt.zero_grad()
        out, _ = orc.cyc_forward(Pe, Pd, enc_s, dec_s, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM,
                                 y0_enc=y0e, y0_dec=y0d, eps=eps, masks=masks)
        loss, _ = orc.cyc_loss(out, x, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM, flen_acc=[TCHUNK] * B, select_utt_idx=list(range(B)))
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, "port"


def cpu_decode_step_factory(n_utt):
    """Stage-6 conversion of n_utt utterances x 800 frames the way the reference does it: ONE utterance per call, unbatched
    [T,54] layout, 300 latent samples averaged (decode_*.py:303-305,318) -> (step(), kind)."""
    import torch
    from oracle import gru_vae_oracle as orc
    ref = _load_reference_module()
    x, _, _, tc = orc.synth_batch(n_utt, DEC_T, 200)
    mean, std = orc.synth_stats(NMCEP)
    y0e = torch.zeros(1, 1, 2 * LAT)
    y0d = torch.tensor((0 - mean[STDIM:]) / std[STDIM:], dtype=torch.float32).reshape(1, 1, -1)
    if ref is not None:
        enc, dec, _, _ = _cpu_models(ref, 2)
        enc.eval(); dec.eval()

        def step():
            with torch.no_grad():
                for u in range(n_utt):
                    lat_src, _, _ = enc(x[u], y0e, clamp_vae=True, lat_dim=LAT)
                    rep = lat_src.unsqueeze(0).repeat(N_SMPL, 1, 1)                       # decode_*.py:304
                    smp = rep[:, :, :LAT] + torch.exp(rep[:, :, LAT:] / 2) * torch.randn(N_SMPL, DEC_T, LAT)
                    lat_feat = torch.mean(smp, 0)                                         # :305
                    dec(torch.cat((tc[u], lat_feat), 1), y0d)                             # :318
            return 0.0

        return step, "reference"
    enc_s, dec_s = orc.encoder_spec(STDIM + NMCEP, LAT, HIDDEN), orc.decoder_spec(LAT, 2, NMCEP, HIDDEN)
    Pe = orc.init_params(enc_s, 1, mean=mean, scale=std)
    Pd = orc.init_params(dec_s, 2, mean=mean[STDIM:], scale=std[STDIM:])

    def step():
        for u in range(n_utt):
            eps_mean = torch.randn(DEC_T, LAT) / N_SMPL ** 0.5
            orc.convert(Pe, Pd, enc_s, dec_s, x[u], tc[u], lat_dim=LAT, y0_enc=y0e, y0_dec=y0d, eps_mean=eps_mean)
        return 0.0

    return step, "port"


def time_cpu(step, steps, warmup):
    import torch
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    W = max(1, args.warmup)
    if args.workload == "decode":
        # a step = a bounded sample of the workload: 4 of the utterances, converted one at a time as the reference does
        n = 4
        step, kind = cpu_decode_step_factory(n)
        frames = n * DEC_T
        sample = f"{n} of the {args.batch_utt} utterances x {DEC_T} frames per step, one utterance per call (the reference's batch = 1)"
        metric = METRIC_DECODE
    else:
        step, kind = cpu_train_step_factory(args.batch_utt, args.n_spk)
        frames = args.batch_utt * TCHUNK
        sample = f"the full per-GPU step: {args.batch_utt} utterances x {TCHUNK} frames, cyc2 fwd+bwd+Adam"
        metric = METRIC
    sec, cores = time_cpu(step, args.steps, W)
    fps = frames / sec
    src = "unmodified reference src/nets/gru_vae.py (oracle/_ref)" if kind == "reference" else "oracle/gru_vae_oracle.py (CPU restatement)"
    sample += f"; median of {args.steps} steps; {src}"
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_of(args, args.gpus),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def run_native(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import cyclevae_vc_b200 as cvb  # noqa: F401  (raises when libcyclevae_b200.so is missing: no fallback)
    from cyclevae_vc_b200 import cycle, synth
    from cyclevae_vc_b200._lib import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cyclevae_vc_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)        # plumbing: barriers, max-over-ranks of the timings, id exchange
        comm = cycle.NativeComm(rank, world)                  # the data-path collective: the library's own cvb_allreduce_sum
    decode = args.workload == "decode"
    B, T = args.batch_utt, (DEC_T if decode else TCHUNK)
    enc, dec, y0d1 = synth.build_models(HIDDEN, LAT, args.n_spk, NMCEP, STDIM, seed=1, device=dev)
    torch.manual_seed(1000 + rank)                                       # per-rank noise / dropout streams
    host = synth.make_batch(B, T, 100 + rank, args.n_spk, NMCEP, pin=True)   # this rank's utterance shard
    y0e = torch.zeros(B, 1, 2 * LAT, device=dev)
    y0d = y0d1.to(dev).repeat(B, 1, 1).contiguous()

    if decode:
        enc.eval(); dec.eval()
        host = (host[0], host[3])                                        # features, target-speaker code
        out_host = torch.empty(B, T, NMCEP).pin_memory()

        def step(x, tc):
            return cycle.convert(enc, dec, x, tc, lat_dim=LAT, y0_enc=y0e, y0_dec=y0d, n_smpl=N_SMPL)

        def read_back(res):
            out_host.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(out_host[0, 0, 0])

        d2h = B * T * NMCEP * 4
    else:
        enc.train(); dec.train()
        opt = cycle.FlatAdam(cycle.trainable_parameters(enc, dec), lr=1e-4)
        # the fused step driver: forward + losses + BPTT replayed as one CUDA graph, all-reduce, Adam
        cs = cycle.CycleStep(enc, dec, opt, B=B, T=T, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM, n_spk=args.n_spk, y0_enc=y0e, y0_dec=y0d,
                             graph=not args.no_graph, comm=comm)

        def step(x, cv, sc, tc):
            return cs.step(x, cv, sc, tc)

        def read_back(res):
            return float(res.item())

        d2h = 4
    if decode:
        devb = [t.to(dev) for t in host]
        stage = [torch.empty_like(t, device=dev) for t in host]
    else:   # the step driver's own static input buffers: resident inputs are used in place, host inputs are copied into them
        stage = [cs.x, cs.cv, cs.sc, cs.tc]
        for d, h in zip(stage, host):
            d.copy_(h)
        devb = stage

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    W = max(3, args.warmup)
    for _ in range(W):
        step(*devb)
    barrier()
    # ---- device-resident timing (inputs already in HBM) ----------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.cvb_profile_reset()
    lib.cvb_profile_enable(1 if decode or args.no_graph else 0)
    n0 = lib.cvb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(*devb)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = (lib.cvb_launch_count() - n0) // args.steps + (0 if decode else cs.kernels_per_replay)
    lib.cvb_profile_enable(0)
    prof_steps = args.steps
    if not (decode or args.no_graph):
        # Per-kernel durations: a replayed graph has no host-side launches to bracket, so the SAME step (same kernels, same
        # buffers) is run kernel by kernel for a few extra steps with CUDA events around every recurrence launch and
        # front-end chain on the launching stream.  Outside the timed region above; only feeds the roofline objects.
        eager = cycle.CycleStep(enc, dec, opt, B=B, T=T, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM, n_spk=args.n_spk, y0_enc=y0e, y0_dec=y0d,
                                graph=False, comm=comm)
        for d, h in zip((eager.x, eager.cv, eager.sc, eager.tc), host):
            d.copy_(h)
        eager.step(eager.x, eager.cv, eager.sc, eager.tc)
        barrier()
        lib.cvb_profile_reset()
        lib.cvb_profile_enable(2)   # level 2 also brackets every dense-product launch (roofline_gemm)
        prof_steps = 3
        for _ in range(prof_steps):
            eager.step(eager.x, eager.cv, eager.sc, eager.tc)
        barrier()
        lib.cvb_profile_enable(0)
    prof = {}
    for kind, name in ((0, "k_gru_fwd"), (1, "k_gru_bwd"), (3, "frontend_fwd"), (2, "k_gemm_tc")):
        tot, n = C.c_float(0), C.c_int(0)
        lib.cvb_profile_summary(kind, C.byref(tot), C.byref(n))
        prof[name] = (tot.value, n.value)
    lib.cvb_profile_reset()
    # ---- end-to-end timing: host (pinned) inputs -> H2D -> step -> result read back ---------------
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    for _ in range(args.steps):
        for d, h in zip(stage, host):
            d.copy_(h, non_blocking=True)
        last = read_back(step(*stage))
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3) / args.steps)
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(t.numel() * t.element_size() for t in host)
    full_ms = None
    if decode:
        # the complete model work of stage 6 (decode_*.py:303-323): 2 ENC (source, target features) + 3 DEC
        x, tc = devb
        sc = torch.roll(tc, 1, 2)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(max(1, args.steps // 2)):
            with torch.no_grad():
                lat_s, _, _ = enc(x, y0e, clamp_vae=True, lat_dim=LAT)
                lat_t, _, _ = enc(x, y0e, clamp_vae=True, lat_dim=LAT)
                eps = torch.randn(B, T, LAT, device=dev) / N_SMPL ** 0.5
                from cyclevae_vc_b200 import gru_vae as gv
                dec(gv.reparam_concat(lat_s, tc, eps, LAT), y0d)
                dec(gv.reparam_concat(lat_s, sc, eps, LAT), y0d)
                dec(gv.reparam_concat(lat_t, tc, eps, LAT), y0d)
        f1.record()
        barrier()
        full_ms = max_over_ranks(f0.elapsed_time(f1) / max(1, args.steps // 2))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained")
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
        if not peak_tf:
            peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
        peak_hbm = peaks.get("hbm_gbs") or 6400.0
        # dominant kernel = the recurrence kernel with the larger share of the step
        rec = {k: v for k, v in prof.items() if k.startswith("k_gru")}
        dom = max(rec, key=lambda k: rec[k][0])
        tot_ms, n_l = rec[dom]
        avg_ms = tot_ms / max(1, n_l)
        rows = min(B, 128)                                               # batch rows of one launch (wider batches are sliced)
        if decode:
            fl = rows * T * folded_flops_per_frame()
            note = ("folded inference recurrence (gru_tc_eval.cu): one exchange of h per step, [4H x H] fp16 hi+lo operand resident in "
                    "shared memory; `achieved` counts algorithmic fp32-equivalent FLOPs (issued 16-bit MMA FLOPs are 3x)")
        else:
            bwd = dom == "k_gru_bwd"
            # launches alternate encoder (out 64) / decoder (out 50) passes: 4 enc + 6 dec per step
            fl = rows * T * (4 * recurrence_flops_per_frame(2 * LAT, bwd) + 6 * recurrence_flops_per_frame(NMCEP, bwd)) / 10.0
            note = ("split-precision tcgen05 recurrence (fp16/bf16 hi+lo operands, fp32 TMEM accumulation); `achieved` counts ALGORITHMIC "
                    "fp32-equivalent FLOPs (issued 16-bit MMA FLOPs are 3x) against the dense bf16 peak; not tensor-bound: a recurrent step is "
                    "0.54 GFLOP (0.4 us at the peak) behind a grid-wide exchange -- the forward kernel (gru_tc2.cu) is bound by its one hop + "
                    "the chain totals -> operand swap -> gates -> partial swap -> release, the BPTT kernel at B >= 48 by the L2 -> SM broadcast "
                    "of dgh (245 KB per CTA and step, the same lines wanted by 32 clusters) plus one hop (DESIGN.md section 4)")
        achieved = fl / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        # which generation of the training kernels ran (gru_tc2*.cu: one grid-wide exchange per step, gru_tc*.cu: two), and both
        # directions side by side
        hops = {"k_gru_fwd": lib.cvb_last_recurrence_hops(0), "k_gru_bwd": lib.cvb_last_recurrence_hops(1)}

        def kname(k):
            return k + ("_tc_eval" if decode else "_tc2" if hops[k] == 1 else "_tc")

        both = []
        for k, (t_ms, n_k) in rec.items():
            if not n_k:
                continue
            a_ms = t_ms / n_k
            if decode:
                f_k = rows * T * folded_flops_per_frame()
            else:
                f_k = rows * T * (4 * recurrence_flops_per_frame(2 * LAT, k == "k_gru_bwd") + 6 * recurrence_flops_per_frame(NMCEP, k == "k_gru_bwd")) / 10.0
            both.append({"kernel": kname(k), "grid_wide_exchanges_per_step": hops[k], "avg_launch_ms": a_ms, "us_per_recurrent_step": a_ms * 1e3 / T,
                         "achieved": f_k / (a_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "frac": f_k / (a_ms * 1e-3) / 1e12 / peak_tf,
                         "share_of_step": t_ms / prof_steps / ms})
        traffic, traffic_src = None, None
        try:
            for line in open(os.path.join(ROOT, "profiles", "r02_ncu_summary.csv")):
                c = line.strip().split(",")
                if c[0] == kname(dom) and int(c[1]) == B:
                    traffic = (float(c[2]) + float(c[3])) * 1e6
                    traffic_src = "profiles/r02_ncu_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
        except Exception:
            pass
        cfg = config_of(args, world)
        cfg["l2"] = "per-step working set (saved gate activations / gx, > 1 GB at the default batch) exceeds the 126 MB L2; no flush needed"
        cfg["last_result"] = last
        if world > 1 and not decode:
            cfg["collective"] = f"one cvb_allreduce_sum per step (ncclAllReduce SUM fp32 over the flat gradient buffer, {opt.n * 4 / 1e6:.1f} MB)"
        if full_ms is not None:
            cfg["full_stage6"] = {"composition": "2 ENC + 3 DEC (decode_*.py:303-323)", "ms_per_step": full_ms,
                                  "frames_per_s": B * T * world / (full_ms * 1e-3)}
        fe_ms, fe_n = prof.get("frontend_fwd", (0.0, 0))
        roof_fe = None
        if fe_n:
            n_enc = fe_n * (1 if decode else 4) // (2 if decode else 10)
            by = rows * T * (n_enc * frontend_bytes_per_frame(STDIM + NMCEP) + (fe_n - n_enc) * frontend_bytes_per_frame(LAT + args.n_spk)) / fe_n
            gbs = by / (fe_ms / fe_n * 1e-3) / 1e9
            roof_fe = {"bound": "hbm", "kernel": "front-end chain (scale_in + two-sided dilated conv + dropout -> xc)", "achieved": gbs,
                       "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm, "avg_ms": fe_ms / fe_n, "launches_timed": fe_n,
                       "algorithmic_bytes_per_call": by,
                       "note": "bytes = x in + xc out once (SURVEY.md §8d); at fp32-parity precision the conv is compute-bound "
                               "(243 FLOP/B ENC), so the tensor pipe, not HBM, binds this chain"}
        res = {
            "metric": METRIC_DECODE if decode else METRIC, "value": B * T * world / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": B * T * world / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": kname(dom), "achieved": achieved, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                         "avg_launch_ms": avg_ms, "launches_timed": n_l, "algorithmic_flops_per_launch": fl, "peak_source": peak_src,
                         "share_of_step": {k: v[0] / prof_steps / ms for k, v in prof.items() if k != "k_gemm_tc"},
                         "us_per_recurrent_step": avg_ms * 1e3 / T, "rows_per_launch": rows, "note": note},
            "roofline_recurrence": both,
        }
        if roof_fe:
            res["roofline_frontend"] = roof_fe
        gm_ms, gm_n = prof.get("k_gemm_tc", (0.0, 0))
        if gm_n and not decode:
            enc_in, dec_in = STDIM + NMCEP, LAT + args.n_spk
            fl_g = sum(4 * dense_flops_per_pass(enc_in, 2 * LAT, B, T, bw) + 6 * dense_flops_per_pass(dec_in, NMCEP, B, T, bw) for bw in (False, True))
            g_tf = fl_g / (gm_ms / prof_steps * 1e-3) / 1e12
            res["roofline_gemm"] = {
                "bound": "tensor", "kernel": "k_gemm_tc (every dense product of the step outside the recurrence kernels)", "achieved": g_tf,
                "issued": 3 * g_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": g_tf / peak_tf, "frac_issued": 3 * g_tf / peak_tf,
                "ms_per_step": gm_ms / prof_steps, "launches_per_step": gm_n // prof_steps, "algorithmic_flops_per_step": fl_g,
                "share_of_step": gm_ms / prof_steps / ms,
                "note": "split-precision products: 3 16-bit MMAs per algorithmic fp32 product (`issued`); the mainloop sits at the "
                        "L2 -> SM delivery ceiling (64 KB per 128x128x64 stage in 1280-1560 cycles = 6.0-6.4 KB/clk chip-wide against 813 "
                        "cycles of MMAs), DESIGN.md section 4; times are CUDA events around each group's launches in the eager profiling steps (k_gemm_tc + its split-K reductions with their launch gaps, operand pass excluded): an upper bound of the kernel time inside the replayed graph (profiles/*_step_timeline.txt: 4.1 ms)"}
        if not decode:
            res["roofline_streaming"] = streaming_rooflines(enc, dec, opt, B, T, args.n_spk, dev, peak_hbm)
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            if decode:
                n = 4
                cstep, kind = cpu_decode_step_factory(n)
                sec, cores = time_cpu(cstep, 2, 1)
                val, what = n * DEC_T / sec, f"{n} of the {B} utterances x {DEC_T} frames, one utterance per call as the reference decodes"
            else:
                cstep, kind = cpu_train_step_factory(B, args.n_spk)
                sec, cores = time_cpu(cstep, 2, 1)
                val, what = B * T / sec, f"the same step: {B} utterances x {T} frames"
            res["cpu_baseline"] = {"value": val, "unit": "frames/s", "cores": cores, "kind": kind,
                                   "sample": f"{what}; median of 2 steps after 1 warm-up ({sec:.2f} s/step); "
                                             + ("unmodified reference gru_vae.py (oracle/_ref)" if kind == "reference" else "oracle/gru_vae_oracle.py")
                                             + " on the host CPU"}
        print(json.dumps(res))
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
