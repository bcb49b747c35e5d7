#!/usr/bin/env python
"""mcep frames/sec of one CycleVAE optimisation step (cyc2: 4 encoder + 6 decoder GRU_RNN passes forward,
losses, BPTT, gradient all-reduce at N>1, Adam) -- BASELINE.json's metric on configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch-utt B] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
Workload (config.workload): hu1024 ld32 ks3 ds2 cyc2, 80-frame chunks (the reference's --batch_size 80 frames,
train_*.py:70-134) x B utterances per GPU (default 80: the literal "bs80" of configs[1]; the recipe's own
batch_size_utt values 1 / 8 are reachable with --batch-utt).  Weak scaling: B per GPU is fixed.

`--impl reference` times the CPU restatement of the reference's path (oracle/gru_vae_oracle.py -- the reference
is pure Python and cannot travel to the GPU box) on the host cores with the same step
composition, on a bounded sample of the workload.
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

HIDDEN, LAT, NSPK, NMCEP, STDIM, NCYC, TCHUNK = 1024, 32, 2, 50, 4, 2, 80
METRIC = "mcep frames/sec (cyc2 enc+dec fwd+bwd)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch-utt", type=int, default=80, help="utterances per GPU (80-frame chunk each)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--ref-batch-utt", type=int, default=8, help="utterances in the CPU sample of --impl reference / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(B):
    return f"SF1<->TF1 CycleVAE hu{HIDDEN} ld{LAT} ks3 ds2 cyc{NCYC}, {TCHUNK}-frame chunk x {B} utterances per GPU (bs80)"


# ------------------------------------------------------------------------------------------------
# algorithmic work of the recurrence kernels: per frame and per network pass
def recurrence_flops_per_frame(out_dim, backward):
    fwd = 2 * 3 * HIDDEN * (HIDDEN + out_dim) + 2 * out_dim * HIDDEN          # W_hh h + W_y y ; W_o o
    if not backward:
        return fwd
    return 2 * 3 * HIDDEN * HIDDEN + 2 * 3 * HIDDEN * out_dim + 2 * out_dim * HIDDEN   # dgh W_hh ; dgi W_y ; dy W_o


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
def cpu_reference_step_factory(state_enc, state_dec, B):
    """One cyc2 fwd+bwd+Adam step of the CPU restatement (the checker, used here only as the timed baseline)."""
    import torch
    from oracle import gru_vae_oracle as orc
    from cyclevae_vc_b200 import synth
    enc, dec = orc.encoder_spec(STDIM + NMCEP, LAT, HIDDEN), orc.decoder_spec(LAT, NSPK, NMCEP, HIDDEN)
    Pe = {k: v.detach().cpu().clone() for k, v in state_enc.items()}
    Pd = {k: v.detach().cpu().clone() for k, v in state_dec.items()}
    train = []
    for P in (Pe, Pd):
        for k, v in P.items():
            if not k.startswith("scale_"):
                v.requires_grad_(True)
                train.append(v)
    opt = torch.optim.Adam(train, lr=1e-4)
    x, cv, sc, tc = synth.make_batch(B, TCHUNK, 0, NSPK, NMCEP)
    mean, std = synth.feature_stats(NMCEP)
    y0e = torch.zeros(B, 1, 2 * LAT)
    y0d = torch.tensor((0 - mean[STDIM:]) / std[STDIM:], dtype=torch.float32).reshape(1, 1, -1).repeat(B, 1, 1)

    def step():
        eps = [[torch.randn(B, TCHUNK, LAT) for _ in range(3)] for _ in range(NCYC)]          # gru_vae.py:91
        masks = [[((torch.rand(B, TCHUNK, s.conv_dim) >= 0.5).float() * 2, (torch.rand(B, TCHUNK, HIDDEN) >= 0.5).float() * 2)
                  for s in (enc, dec, dec, enc, dec)] for _ in range(NCYC)]                      # nn.Dropout draws
        opt.zero_grad()
        out, _ = orc.cyc_forward(Pe, Pd, enc, dec, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM,
                                 y0_enc=y0e, y0_dec=y0d, eps=eps, masks=masks)
        loss, _ = orc.cyc_loss(out, x, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM, flen_acc=[TCHUNK] * B, select_utt_idx=list(range(B)))
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


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
    from cyclevae_vc_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.ref_batch_utt
    enc, dec, _ = synth.build_models(HIDDEN, LAT, NSPK, NMCEP, STDIM, device=None)
    step = cpu_reference_step_factory(enc.state_dict(), dec.state_dict(), B)
    sec, cores = time_cpu(step, args.steps, max(1, args.warmup))
    fps = B * TCHUNK / sec
    sample = f"{B} utterances x {TCHUNK} frames per step (bounded sample of the workload), median of {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(1, args.warmup), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args.batch_utt), "cpu_sample": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    import cyclevae_vc_b200 as cvb
    from cyclevae_vc_b200 import cycle, synth
    from cyclevae_vc_b200._lib import lib
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cyclevae_vc_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T = args.batch_utt, TCHUNK
    enc, dec, y0d1 = synth.build_models(HIDDEN, LAT, NSPK, NMCEP, STDIM, seed=1, device=dev)
    enc.train(); dec.train()
    opt = cycle.FlatAdam(cycle.trainable_parameters(enc, dec), lr=1e-4)
    torch.manual_seed(1000 + rank)                                  # per-rank noise / dropout streams
    host = synth.make_batch(B, T, 100 + rank, NSPK, NMCEP, pin=True)   # this rank's utterance shard
    devb = [t.to(dev) for t in host]
    stage = [torch.empty_like(t, device=dev) for t in host]
    y0e = torch.zeros(B, 1, 2 * LAT, device=dev)
    y0d = y0d1.to(dev).repeat(B, 1, 1).contiguous()
    flens = torch.full((B,), T, dtype=torch.int32, device=dev)
    sel = list(range(B))

    def step(x, cv, sc, tc):
        opt.zero_grad()
        out, _ = cycle.cyc_forward(enc, dec, x=x, cv=cv, src_code=sc, trg_code=tc, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM,
                                   y0_enc=y0e, y0_dec=y0d, do=True)
        loss, _ = cycle.cyc_loss(out, x, n_cyc=NCYC, lat_dim=LAT, stdim=STDIM, flen_acc=None, select_utt_idx=sel, flens_dev=flens)
        loss.backward()
        cycle.allreduce_grads(opt.grad)
        opt.step()
        return loss

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
    lib.cvb_profile_enable(1)
    n0 = lib.cvb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(*devb)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = (lib.cvb_launch_count() - n0) // args.steps
    lib.cvb_profile_enable(0)
    prof = {}
    for kind, name in ((0, "k_gru_fwd"), (1, "k_gru_bwd")):
        tot, n = C.c_float(0), C.c_int(0)
        lib.cvb_profile_summary(kind, C.byref(tot), C.byref(n))
        prof[name] = (tot.value, n.value)
    lib.cvb_profile_reset()
    # ---- end-to-end timing: host (pinned) inputs -> H2D -> step -> loss read back ----------------
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    for _ in range(args.steps):
        for d, h in zip(stage, host):
            d.copy_(h, non_blocking=True)
        last = float(step(*stage).item())
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3) / args.steps)
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(t.numel() * t.element_size() for t in host)

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
        # dominant kernel = the recurrence kernel with the larger share of the step
        dom = max(prof, key=lambda k: prof[k][0])
        tot_ms, n_l = prof[dom]
        avg_ms = tot_ms / max(1, n_l)
        bwd = dom == "k_gru_bwd"
        # launches alternate encoder (out 64) / decoder (out 50) passes: 4 enc + 6 dec per step
        fl = B * T * (4 * recurrence_flops_per_frame(2 * LAT, bwd) + 6 * recurrence_flops_per_frame(NMCEP, bwd)) / 10.0
        achieved = fl / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (B=80 T=80 ENC pass);
        # only quoted for the workload it was captured on
        traffic, traffic_src = None, None
        if B == 80 and T == 80:
            try:
                for line in open(os.path.join(ROOT, "profiles", "r01_ncu_tc_summary.csv")):
                    c = line.strip().split(",")
                    if c[0] == dom + "_tc":
                        traffic = (float(c[8]) + float(c[9])) * 1e6
                        traffic_src = "profiles/r01_ncu_tc_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
            except Exception:
                pass
        res = {
            "metric": METRIC, "value": B * T * world / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(B), "frames_per_step_per_gpu": B * T, "passes_per_step": "4 ENC + 6 DEC fwd+bwd",
                       "parallelism": f"dp{world}", "optimizer": "Adam (fused, in timed region)",
                       "l2": "per-step working set (saved gate activations ~1.3 GB at B=80) exceeds the 126 MB L2; no flush needed",
                       "last_loss": last},
            "e2e": {"value": B * T * world / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": traffic, "traffic_source": traffic_src, "avg_launch_ms": avg_ms, "launches_timed": n_l,
                         "algorithmic_flops_per_launch": fl, "peak_source": peak_src,
                         "share_of_step": {k: v[0] / args.steps / ms for k, v in prof.items()},
                         "us_per_recurrent_step": avg_ms * 1e3 / T,
                         "note": "split-precision tcgen05 recurrence (fp16/bf16 hi+lo operands, 2 MMAs per K step, fp32 TMEM "
                                 "accumulation); `achieved` counts ALGORITHMIC fp32-equivalent FLOPs (the issued 16-bit MMA FLOPs are "
                                 "3x) against the dense bf16 peak; the kernel is bound by the two grid-wide exchanges of every "
                                 "recurrent step (us_per_recurrent_step), not by the tensor pipe"},
        }
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            Bc = args.ref_batch_utt
            cstep = cpu_reference_step_factory(enc.state_dict(), dec.state_dict(), Bc)
            sec, cores = time_cpu(cstep, 3, 1)
            res["cpu_baseline"] = {"value": Bc * T / sec, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "sample": f"{Bc} utterances x {T} frames per step, median of 3 steps after 1 warm-up "
                                             f"({sec:.2f} s/step); oracle/gru_vae_oracle.py on the host CPU"}
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
