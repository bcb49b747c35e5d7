"""Data-parallel correctness on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

SURVEY.md §8(e): utterances shard over ranks, every rank walks the SAME chunk schedule (computed from the global longest
utterance), a rank whose utterances have ended contributes zero loss exactly as select_utt_idx does in the trainer
(train_*.py:107-133), and ONE all-reduce(SUM) of the flat gradient buffer makes every rank's gradient equal to the
gradient of a single process that holds all utterances (the reference sums per-utterance losses, train_*.py:1403,1408).
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LAT, STDIM, NCYC, T, HID = 32, 4, 2, 80, 1024


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads_for(rows, chunk, dev, seed_models=1):
    """Flat gradient of the cyc2 step of chunk `chunk` (0-based) for the utterances `rows` of the global batch."""
    from cyclevae_vc_b200 import cycle, synth
    from oracle import gru_vae_oracle as orc
    flens_all = [160, 150, 75, 70]                      # both utterances of the second shard end inside the first chunk
    n_all = len(flens_all)
    sched = cycle.chunk_schedule(flens_all, T)          # the GLOBAL schedule: same number of steps on every rank
    lo, hi, _, _, flen_acc, selected = sched[chunk]
    enc, dec, y0d1 = synth.build_models(HID, LAT, 2, 50, STDIM, seed=seed_models, device=dev)
    enc.train(); dec.train()
    opt = cycle.FlatAdam(cycle.trainable_parameters(enc, dec), lr=1e-4)
    x, cv, sc, tc = (t[:, lo:hi + 1] for t in orc.synth_batch(n_all, 160, 77))
    eps = orc.synth_noise(n_all, T, LAT, NCYC, 78)
    masks = orc.synth_masks(n_all, T, orc.encoder_spec(54, LAT, HID), orc.decoder_spec(LAT, 2, 50, HID), NCYC, 79)
    rows = list(rows)
    pick = lambda t: t[rows].contiguous().to(dev)
    sel_local = [k for k, j in enumerate(rows) if j in selected]
    opt.zero_grad()
    if sel_local:                                        # a rank with no selected utterance runs no step, its gradient stays zero
        out, _ = cycle.cyc_forward(enc, dec, x=pick(x), cv=pick(cv), src_code=pick(sc), trg_code=pick(tc), n_cyc=NCYC, lat_dim=LAT,
                                   stdim=STDIM, y0_enc=torch.zeros(len(rows), 1, 2 * LAT, device=dev),
                                   y0_dec=y0d1.to(dev).repeat(len(rows), 1, 1), do=True,
                                   eps=[[pick(e) for e in ec] for ec in eps], masks=[[(pick(a), pick(b)) for a, b in mc] for mc in masks])
        loss, _ = cycle.cyc_loss(out, pick(x), n_cyc=NCYC, lat_dim=LAT, stdim=STDIM, flen_acc=[flen_acc[j] for j in rows],
                                 select_utt_idx=sel_local, kl_cv_quirk=False)
        loss.backward()
    torch.cuda.synchronize()
    return opt.grad, selected


def _worker(rank, world, port, out_q):
    import torch.distributed as dist
    from cyclevae_vc_b200 import cycle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    res = {}
    comm = cycle.NativeComm(rank, world)                 # the library's own communicator (cvb_comm_init through the C ABI)
    for chunk in (0, 1):
        lo, hi = cycle.shard_utterances(4, rank, world)
        g, selected = _grads_for(range(lo, hi), chunk, dev)
        g_lib = g.clone()
        cycle.allreduce_grads(g)                         # the only collective: SUM (torch.distributed's NCCL group)
        cycle.allreduce_grads(g_lib, comm)               # the same through cvb_allreduce_sum
        torch.cuda.synchronize()
        if rank == 0:
            full, _ = _grads_for(range(4), chunk, dev)
            res[chunk] = (float((g - full).abs().max()), float(full.abs().max()), list(selected), float((g_lib - g).abs().max()))
    comm.close()
    if rank == 0:
        out_q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_summed_shard_gradients_equal_single_gpu_gradient():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][2] == [0, 1, 2, 3] and res[1][2] == [0, 1]        # second chunk: rank 1 has no utterance left (zero gradient, still reduces)
    for chunk in (0, 1):
        err, scale, _, lib_vs_torch = res[chunk]
        assert err <= 2e-4 * max(1.0, scale), (chunk, err, scale)
        assert lib_vs_torch <= 1e-6 * max(1.0, scale), (chunk, lib_vs_torch)   # two-rank sums: the same additions either way
