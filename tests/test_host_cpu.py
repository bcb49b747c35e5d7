"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
the drop-in module mirrors the reference's surface, integer bookkeeping is bit-exact, and the
data-parallel sharding + SUM all-reduce reproduces the single-process gradient (gloo, world_size 2)."""
import json
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from oracle import gru_vae_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    ge.build()
    import cyclevae_vc_b200
    return cyclevae_vc_b200


def test_abi_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "cyclevae_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cvb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 19
    from cyclevae_vc_b200 import _lib
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.cvb_abi_version() == _lib.ABI_VERSION
    m = re.search(r"#define CVB_ABI_VERSION (\d+)", hdr)
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_workspace_sizing_is_host_arithmetic(pkg):
    import ctypes as C
    from cyclevae_vc_b200 import _lib
    m = pkg.GRU_RNN(in_dim=54, out_dim=64, hidden_units=1024, scale_out_flag=False, do_prob=0.5)
    net = _lib.CvbNet()
    net.in_dim, net.out_dim, net.hidden, net.kernel_size, net.n_conv = 54, 64, 1024, 3, 2
    B, T = 8, 80
    fe = _lib.lib.cvb_frontend_ws_floats(C.byref(net), B, T)
    R = B * (T + 8)
    # composed front-end (every shape the tensor-core GEMM takes): the normalised padded grid + 9 slack rows, and xc;
    # the layer-by-layer path of tiny shapes keeps every layer's grid and the repacked weights
    assert R * 54 + 9 * 54 + B * T * 486 <= fe < R * (54 + 162) + B * T * 486
    net.in_dim, net.hidden = 3, 8
    fe_tiny = _lib.lib.cvb_frontend_ws_floats(C.byref(net), 2, 5)
    assert fe_tiny >= 2 * 13 * (3 + 9 + 27) + 3 * (9 * 3 + 27 * 9) + 2 * 5 * 27
    net.in_dim, net.hidden = 54, 1024
    rec = _lib.lib.cvb_recurrent_ws_floats(C.byref(net), B, T, 1, 1)
    assert rec >= (T + 1) * B * (1024 + 64) + 5 * T * B * 1024
    # the forward scratch holds gx (+ the folded-feedback matrix of the inference kernel); training adds the BPTT scratch (max of both)
    fwd_only = _lib.lib.cvb_scratch_floats(C.byref(net), B, T, 0)
    assert fwd_only >= T * B * 3 * 1024 and _lib.lib.cvb_scratch_floats(C.byref(net), B, T, 1) >= fwd_only


def test_module_surface_matches_reference(pkg, golden_dir):
    g = np.load(os.path.join(golden_dir, "init.npz"), allow_pickle=False)
    for tag, kw in (("enc", dict(in_dim=54, out_dim=32, hidden_units=128, scale_out_flag=False, do_prob=0.5)),
                    ("dec", dict(in_dim=18, out_dim=50, hidden_units=128, scale_in_flag=False, do_prob=0.5))):
        torch.manual_seed(1)
        m = pkg.GRU_RNN(**kw)
        m.apply(pkg.initialize)
        assert list(m.state_dict().keys()) == [str(k) for k in g[f"{tag}/keys"]]
        for k, v in m.state_dict().items():
            want = g[f"{tag}/{k}"]
            got = np.array([v.double().sum().item(), v.double().abs().sum().item(), float(v.reshape(-1)[v.numel() // 2])])
            assert np.allclose(got, want, rtol=0, atol=1e-9), (tag, k)
    m = pkg.GRU_RNN(in_dim=54, out_dim=64, hidden_units=128, scale_out_flag=False, do_prob=0.5)
    assert (m.in_dim, m.out_dim, m.receptive_field, m.tot_in_dim) == (54, 64, 9, 54 * 9 + 64)
    # attribute access patterns of the trainer (train_*.py:344-347, 369-376)
    m.scale_in.weight = torch.nn.Parameter(torch.diag(torch.ones(54)).unsqueeze(2))
    m.scale_in.bias = torch.nn.Parameter(torch.zeros(54))
    assert len(list(m.conv.parameters())) == 4 and len(list(m.gru.parameters())) == 4 and len(list(m.out_1.parameters())) == 2
    import inspect
    sig = inspect.signature(m.forward)
    assert list(sig.parameters) == ["x", "y_in", "softmax", "sigmoid", "exp", "h_in", "noise", "res", "res_stdim", "res_endim",
                                    "do", "clamp_vae", "relu_vae", "lat_dim", "clamp_vae_laplace"]
    assert list(inspect.signature(pkg.GRU_RNN.__init__).parameters)[1:] == [
        "in_dim", "out_dim", "hidden_units", "hidden_layers", "kernel_size", "dilation_size", "do_prob", "scale_in_flag",
        "scale_out_flag", "scale_in_out_flag"]


def test_no_cpu_fallback_and_unused_kwargs_raise(pkg):
    m = pkg.GRU_RNN(in_dim=5, out_dim=4, hidden_units=8)
    x, y = torch.zeros(2, 3, 5), torch.zeros(2, 1, 4)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(x, y)
    for kw in (dict(softmax=True), dict(res=True), dict(noise=0.1), dict(relu_vae=True), dict(clamp_vae_laplace=True)):
        with pytest.raises(NotImplementedError):
            m(x, y, **kw)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pkg.TwoSidedDilConv1d(in_dim=5)(torch.zeros(2, 5, 7))    # stand-alone conv: same rule, no eager path
    with pytest.raises(NotImplementedError):
        pkg.GRU_RNN(hidden_layers=2)
    with pytest.raises(NotImplementedError):
        pkg.TWFSEloss()(torch.zeros(3, 4), torch.zeros(3, 4))   # default L2=True, GV=True is never used by the scripts


def test_dropin_shim_imports_as_gru_vae(pkg):
    sys.path.insert(0, os.path.join(ROOT, "cyclevae_vc_b200", "dropin"))
    try:
        sys.modules.pop("gru_vae", None)
        import gru_vae
        for n in ("GRU_RNN", "initialize", "TWFSEloss", "sampling_vae_batch", "loss_vae"):
            assert hasattr(gru_vae, n)
        assert gru_vae.GRU_RNN is pkg.GRU_RNN
    finally:
        sys.path.pop(0)
        sys.modules.pop("gru_vae", None)


def test_chunk_schedule_bit_exact(pkg, golden_dir):
    from cyclevae_vc_b200.cycle import chunk_schedule
    cases = json.load(open(os.path.join(golden_dir, "chunks.json")))
    assert cases
    for c in cases:
        rows = chunk_schedule(c["flens"], c["bs"], c["spc"])
        assert len(rows) == len(c["rows"])
        for got, want in zip(rows, c["rows"]):
            assert [got[0], got[1], got[2], got[3], got[4], got[5]] == want
    # default speech-frame index = every frame; same answers as the oracle restatement on ragged inputs
    for flens, bs in (([5], 80), ([81, 80, 79, 1], 80), ([200, 333, 17], 20)):
        assert chunk_schedule(flens, bs) == orc.chunk_schedule(flens, bs)


def test_pack_utterances_and_gv_postfilter(pkg):
    """Host logic of the batched stage-6 front door: padding with the front-end's own pad value, and
    the GV post-filter against decode_*.py:419-420 restated in numpy."""
    from cyclevae_vc_b200 import cycle
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(T, 6, generator=g) for T in (5, 1, 9)]
    pad = torch.arange(6, dtype=torch.float32)
    x, lens = cycle.pack_utterances(feats, pad)
    assert lens == [5, 1, 9] and x.shape == (3, 9, 6)
    for i, f in enumerate(feats):
        assert torch.equal(x[i, :lens[i]], f)
        assert torch.equal(x[i, lens[i]:], pad.expand(9 - lens[i], 6))
    cv = torch.randn(40, 50, generator=g, dtype=torch.float64).numpy()
    rng = np.random.default_rng(1)
    gv_trg, cvgv = rng.uniform(0.5, 2.0, 49), rng.uniform(0.5, 2.0, 49)
    datamean = np.mean(cv[:, 1:], axis=0)
    ref = np.c_[cv[:, 0], np.sqrt(gv_trg / cvgv) * (cv[:, 1:] - datamean) + datamean]
    mine = cycle.gv_postfilter(torch.tensor(cv), torch.tensor(gv_trg), torch.tensor(cvgv)).numpy()
    assert np.abs(mine - ref).max() < 1e-12
    # GV statistics over a set of converted utterances, calc_cvgv_*.py:203,320-321 restated in numpy
    convs = [torch.randn(T, 50, generator=g, dtype=torch.float64) for T in (30, 12, 51)]
    cvlist = [np.var(c.numpy()[:, 1:], axis=0) for c in convs]
    m, v = cycle.cvgv_stats(convs)
    assert np.abs(m.numpy() - np.mean(np.array(cvlist), axis=0)).max() < 1e-12
    assert np.abs(v.numpy() - np.var(np.array(cvlist), axis=0)).max() < 1e-12


def test_feature_pack_roundtrip_and_trainer_batch(pkg, tmp_path):
    """the flat feature cache returns the arrays bit for bit, PairDataset items equal the reference's
    FeatureDatasetSingleVAE.__getitem__ semantics (dataset.py:67-101: zero padding to pad_len with `padding()`,
    dataset.py:18-26; 2-way speaker codes keyed on the directory's speaker), and collate_trimmed equals the DataLoader's
    default collate followed by the trainer's trimming (train_*.py:47-63)."""
    from cyclevae_vc_b200 import features
    rng = np.random.default_rng(5)
    utts = []
    for spk in ("SF1", "TF1"):
        for k, T in enumerate((31, 17, 44)):
            n_spc = T - 6
            utts.append({"name": f"{spk}/utt{k}", "spk": spk, "feat_org_lf0": rng.normal(size=(T + (spk == "TF1") * 3, 54)),
                         "cvuvlogf0fil_ap": rng.normal(size=(T + (spk == "TF1") * 3, 4)),
                         "spcidx_range": np.sort(rng.choice(T, n_spc, replace=False))[None, :]})
    path = str(tmp_path / "feats.cvbfeat")
    features.write_pack(path, utts)
    pack = features.FeaturePack(path)
    assert len(pack) == 6 and pack.names == [u["name"] for u in utts]
    for i, u in enumerate(utts):
        assert np.array_equal(pack.array(i, "feat_org_lf0"), u["feat_org_lf0"].astype(np.float32))
        assert np.array_equal(pack.array(i, "cvuvlogf0fil_ap"), u["cvuvlogf0fil_ap"].astype(np.float32))
        assert np.array_equal(pack.array(i, "spcidx_range"), u["spcidx_range"][0])
    src = [f"SF1/utt{k}" for k in range(3)] + [f"TF1/utt{k}" for k in range(3)]
    src_trg = [f"TF1/utt{k}" for k in range(3)] + [f"SF1/utt{k}" for k in range(3)]
    ds = features.PairDataset(pack, src, src_trg, spk_src="SF1", pad_len=50)
    it = ds[4]   # a TF1 utterance: codes swap (dataset.py:77-80)
    u, v = utts[4], utts[1]
    T = u["feat_org_lf0"].shape[0]
    assert it["flen_src"] == T and it["flen_src_trg"] == v["feat_org_lf0"].shape[0] and it["flen_spc_src"] == u["spcidx_range"].shape[1]
    assert it["h_src"].shape == (50, 54) and it["h_src"].dtype == torch.float32 and it["spcidx_src"].dtype == torch.int64
    assert torch.equal(it["h_src"][:T], torch.tensor(u["feat_org_lf0"].astype(np.float32))) and float(it["h_src"][T:].abs().sum()) == 0
    assert torch.equal(it["src_code"][:T], torch.tensor([[0.0, 1.0]]).expand(T, 2)) and float(it["src_code"][T:].abs().sum()) == 0
    assert torch.equal(it["trg_code"][:T], torch.tensor([[1.0, 0.0]]).expand(T, 2))
    assert torch.equal(it["spcidx_src"][:it["flen_spc_src"]], torch.tensor(u["spcidx_range"][0]))
    assert torch.equal(it["h_src_trg"][:it["flen_src_trg"]], torch.tensor(v["feat_org_lf0"].astype(np.float32)))
    # trainer batch: default collate of the padded items, then the trimming of train_*.py:47-63
    idxs = [0, 4, 2]
    ref = torch.utils.data.default_collate([ds[i] for i in idxs])
    got = features.collate_trimmed(ds, idxs)
    for k, lk in (("h_src", "flen_src"), ("src_code", "flen_src"), ("trg_code", "flen_src"), ("cv_src", "flen_src"),
                  ("spcidx_src", "flen_spc_src"), ("h_src_trg", "flen_src_trg"), ("spcidx_src_trg", "flen_spc_src_trg")):
        mx = int(ref[lk].max())
        assert torch.equal(got[lk], ref[lk])
        assert got[k].dtype == ref[k].dtype and torch.equal(got[k], ref[k][:, :mx]), k
    assert got["featfile_src"] == ref["featfile_src"]
    staged = features.DeviceStager(torch.device("cpu")).put(got)   # CPU: pass-through copies through the slot buffers
    assert all(torch.equal(staged[k], got[k]) for k in got if torch.is_tensor(got[k]))


def test_shard_utterances(pkg):
    from cyclevae_vc_b200.cycle import shard_utterances
    for n, w in ((80, 8), (7, 2), (5, 4), (3, 8)):
        spans = [shard_utterances(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert [hi - lo for lo, hi in spans] == [len(s) for s in np.array_split(np.arange(n), w)]


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cyclevae_vc_b200.cycle import allreduce_grads, shard_utterances
    spec = orc.NetSpec(in_dim=6, out_dim=4, hidden_units=12, do_prob=0.0, scale_in=False, scale_out=False)
    P = {k: v.requires_grad_(True) for k, v in orc.init_params(spec, 9, gain=2.0, bias_std=0.1).items()}
    g = torch.Generator().manual_seed(0)
    B, T = 5, 9
    x, y0 = torch.randn(B, T, 6, generator=g), torch.randn(B, 1, 4, generator=g)
    tgt = torch.randn(B, T, 4, generator=g)
    lo, hi = shard_utterances(B, rank, world)
    o, _, _ = orc.gru_rnn_forward(P, spec, x[lo:hi], y0[lo:hi])
    loss = sum(orc.mcd_l1(o[j], tgt[lo + j])[1] for j in range(hi - lo))   # per-utterance means, SUMMED (train_*.py:1403)
    loss.backward()
    flat = torch.cat([P[k].grad.reshape(-1) for k in sorted(P)])
    allreduce_grads(flat)
    if rank == 0:
        q.put(flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_sum_allreduce_equals_single_process(pkg):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = orc.NetSpec(in_dim=6, out_dim=4, hidden_units=12, do_prob=0.0, scale_in=False, scale_out=False)
    P = {k: v.requires_grad_(True) for k, v in orc.init_params(spec, 9, gain=2.0, bias_std=0.1).items()}
    g = torch.Generator().manual_seed(0)
    B, T = 5, 9
    x, y0 = torch.randn(B, T, 6, generator=g), torch.randn(B, 1, 4, generator=g)
    tgt = torch.randn(B, T, 4, generator=g)
    o, _, _ = orc.gru_rnn_forward(P, spec, x, y0)
    sum(orc.mcd_l1(o[j], tgt[j])[1] for j in range(B)).backward()
    want = torch.cat([P[k].grad.reshape(-1) for k in sorted(P)]).numpy()
    assert np.abs(got - want).max() < 1e-5 * max(1.0, np.abs(want).max())
