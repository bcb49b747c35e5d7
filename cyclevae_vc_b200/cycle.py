"""Host-side composition of one CycleVAE optimisation step on top of the drop-in GRU_RNN:
the caller-side logic of src/bin/train_gru_cyclevae_gauss_batch.py restated so the benchmark and
the parity tests drive the hot path exactly as the reference's trainer does (the trainer itself
cannot run here: h5py / dtw_c / pysptk are absent).

    chunk_schedule   train_generator's frame-chunk bookkeeping      train_*.py:70-134   (integers)
    cyc_forward      the 5 x n_cyc GRU_RNN passes of one chunk      train_*.py:1298-1311 / 1326-1338
    cyc_loss         loss assembly incl. the KL-cv cat quirk        train_*.py:1363-1410
    FlatAdam         Adam over conv+gru+out_1 of both nets          train_*.py:373-377, 1418-1420
    allreduce_grads  the one data-parallel collective (SUM)         new
    convert          stage-6 conversion composition                 decode_*.py:303-305,318
    convert_utterances  the same for a list of ragged utterances, packed    new
    gv_postfilter    global-variance post-filter on the device      decode_*.py:419-420
    mcd_aligned      mean / std MCD of aligned frames on the device  train_*.py:1435-1439 (dtw.calc_mcd)
    dtw_org_to_trg   DTW of a trajectory onto the target, on device train_*.py:679-688   (dtw.dtw_org_to_trg)
    cvgv_stats       GV statistics of converted utterances          calc_cvgv_*.py:203,320-321
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import gru_vae as gv
from ._lib import check, lib, ptr


# ------------------------------------------------------------------------------------------------
def chunk_schedule(flens: Sequence[int], batch_size: int, spcidx: Optional[Sequence[Sequence[int]]] = None):
    """Frame-chunk schedule of train_generator (train_*.py:70-134).  Returns one tuple per yielded
    chunk: (src_idx_s, src_idx_e, spcidx_s_idx[], spcidx_e_idx[], flen_acc[], select_utt_idx[]).
    Integer-only; the reference's behaviours are kept: the first chunk always spans batch_size
    frames (:72), flen_acc starts at batch_size (:77) and shrinks only when a later chunk crosses the
    utterance end (:107-108), an utterance stays selected until its last speech frame is consumed (:106).
    """
    flens = [int(f) for f in flens]
    n_utt = len(flens)
    if spcidx is None:
        spcidx = [range(f) for f in flens]
    n_spc = [len(s) for s in spcidx]
    max_flen = max(flens)
    start_i = [-1] * n_utt
    end_i = [-1] * n_utt
    seeking_start = [True] * n_utt

    def advance(j: int, lo: int, hi: int) -> None:
        """Move utterance j's speech-frame cursors to cover the chunk [lo, hi] (:79-98, :109-128)."""
        idx = spcidx[j]
        i = end_i[j] + 1
        while i < n_spc[j]:
            v = int(idx[i])
            if seeking_start[j]:
                if v >= lo:
                    if v > hi:            # no speech frame inside this chunk
                        start_i[j] = -1
                        return
                    start_i[j] = i
                    seeking_start[j] = False
                    if i == n_spc[j] - 1:  # the very last speech frame opens and closes the span
                        end_i[j] = i
                        seeking_start[j] = True
                        return
            elif v >= hi or i == n_spc[j] - 1:
                end_i[j] = i - 1 if v > hi else i
                seeking_start[j] = True
                return
            i += 1

    lo, hi = 0, batch_size - 1
    flen_acc = [batch_size] * n_utt
    for j in range(n_utt):
        advance(j, lo, hi)
    rows = [(lo, hi, list(start_i), list(end_i), list(flen_acc), list(range(n_utt)))]
    while hi < max_flen - 1:
        lo = hi + 1
        hi = min(lo + batch_size - 1, max_flen - 1)
        selected = []
        for j in range(n_utt):
            if end_i[j] >= n_spc[j] - 1:
                continue
            if hi >= flens[j]:
                flen_acc[j] = flens[j] - lo
            advance(j, lo, hi)
            selected.append(j)
        rows.append((lo, hi, list(start_i), list(end_i), list(flen_acc), selected))
    return rows


# ------------------------------------------------------------------------------------------------
PASS_NAMES = ("pp_src", "src_src", "src_trg", "pp_src_trg", "src_trg_src")
OUT_KEYS = ("lat_src", "trj_src_src", "trj_src_trg", "lat_src_trg", "trj_src_trg_src")


def cyc_forward(enc: gv.GRU_RNN, dec: gv.GRU_RNN, *, x, cv, src_code, trg_code, n_cyc: int, lat_dim: int, stdim: int,
                y0_enc, y0_dec, do: bool = True, eps=None, masks=None, state: Optional[dict] = None):
    """The cyc graph of one chunk: per cycle ENC -> DEC(src) / DEC(trg) -> ENC(cv | converted) -> DEC(src)
    (first chunk: train_*.py:1326-1338 with the initial y_in and h_in=None; later chunks: :1298-1311
    with the carried, detached (y, h) in `state`).  eps[i][0..2] / masks[i][p] = (mask_conv, mask_gru)
    inject the noise / dropout masks for parity tests; None draws them on the device."""
    out: Dict[str, List[torch.Tensor]] = {k: [] for k in OUT_KEYS}
    new_state = {}

    def run(net, inp, name, i, p, **kw):
        if state is None:
            y0, h0 = (y0_enc if net is enc else y0_dec), None
        else:
            y0, h0 = state[(name, i)]
            y0, h0 = y0.detach(), h0.detach()
        if masks is not None and masks[i][p][0] is not None:
            net.inject_dropout_masks(*masks[i][p])
        trj, y_last, h_last = net(inp, y0, h_in=h0, do=do, **kw)
        new_state[(name, i)] = (y_last, h_last)
        return trj

    def e(i, k):
        return None if eps is None else eps[i][k]

    for i in range(n_cyc):
        enc_in = x if i == 0 else gv.concat_features(x[:, :, :stdim], out["trj_src_trg_src"][i - 1])
        lat_src = run(enc, enc_in, "pp_src", i, 0, clamp_vae=True, lat_dim=lat_dim)
        trj_src_src = run(dec, gv.reparam_concat(lat_src, src_code, e(i, 0), lat_dim), "src_src", i, 1)
        trj_src_trg = run(dec, gv.reparam_concat(lat_src, trg_code, e(i, 1), lat_dim), "src_trg", i, 2)
        lat_src_trg = run(enc, gv.concat_features(cv, trj_src_trg), "pp_src_trg", i, 3, clamp_vae=True, lat_dim=lat_dim)
        trj_src_trg_src = run(dec, gv.reparam_concat(lat_src_trg, src_code, e(i, 2), lat_dim), "src_trg_src", i, 4)
        for k, v in zip(OUT_KEYS, (lat_src, trj_src_src, trj_src_trg, lat_src_trg, trj_src_trg_src)):
            out[k].append(v)
    return out, new_state


def cyc_loss(out: dict, x, *, n_cyc: int, lat_dim: int, stdim: int, flen_acc: Sequence[int],
             select_utt_idx: Sequence[int], kl_cv_quirk: bool = True, flens_dev: Optional[torch.Tensor] = None):
    """train_*.py:1363-1410 with one kernel per loss term instead of a Python loop over utterances:
    sum over selected utterances of mean-L1-MCD(src_src), mean-L1-MCD(src_trg_src), KL(lat_src) and the
    'cv' KL term, which reproduces line 1393 (cat onto batch_loss_lat_src) when kl_cv_quirk is set."""
    B = x.shape[0]
    if flens_dev is None:
        fl = [0] * B
        for j in select_utt_idx:
            fl[j] = int(flen_acc[j])
        flens_dev = torch.tensor(fl, dtype=torch.int32, device=x.device)
    last = int(select_utt_idx[-1])
    total = None
    parts = []
    for i in range(n_cyc):
        _, m_ss, _ = gv.mcd_l1_per_utt(out["trj_src_src"][i], x, flens_dev, 0, stdim)
        _, m_sts, _ = gv.mcd_l1_per_utt(out["trj_src_trg_src"][i], x, flens_dev, 0, stdim)
        kl_s = gv.kl_per_utt(out["lat_src"][i], flens_dev, lat_dim)
        kl_cv = gv.kl_per_utt(out["lat_src_trg"][i], flens_dev, lat_dim)
        s_ss, s_sts, s_kl = m_ss.sum(), m_sts.sum(), kl_s.sum()
        s_cv = s_kl + kl_cv[last] if (kl_cv_quirk and len(select_utt_idx) > 1) else kl_cv.sum()
        c = s_ss + s_sts + s_kl + s_cv
        total = c if total is None else total + c
        parts.append((s_ss, s_sts, s_kl, s_cv))
    return total, parts


# ------------------------------------------------------------------------------------------------
def trainable_parameters(enc: gv.GRU_RNN, dec: gv.GRU_RNN) -> List[torch.nn.Parameter]:
    """The parameter list the trainer hands to Adam (train_*.py:373-376): conv, gru, out_1 of both nets
    (scale_in / scale_out are frozen statistics, :369-372)."""
    ps: List[torch.nn.Parameter] = []
    for net in (enc, dec):
        ps += list(net.conv.parameters()) + list(net.gru.parameters()) + list(net.out_1.parameters())
    return ps


class FlatAdam:
    """torch.optim.Adam(lr=1e-4) of train_*.py:377 over ONE flat fp32 buffer: parameters and their
    .grad become views of two contiguous tensors, so the data-parallel exchange is a single all-reduce
    and the update is a single kernel (cvb_adam_step)."""

    ALIGN = 16   # floats

    def __init__(self, params: Sequence[torch.nn.Parameter], lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        params = list(params)
        dev = params[0].device
        # every parameter starts on a 64-byte boundary (the kernels read weight rows as float4); the pad elements stay
        # zero in grad / exp_avg / exp_avg_sq, so the single Adam kernel and the single all-reduce run over them harmlessly
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, off in zip(params, offs):
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            p.grad = self.grad[off:off + k].view_as(p.data)
            p.requires_grad_(True)
            p._cvb_grad_sink = True   # GRU_RNN's backward adds its gradients straight into p.grad (gru_vae._GruRnnFn)
        self.params, self.n = params, n
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0

    def zero_grad(self):
        self.grad.zero_()

    def step(self, grad_scale: float = 1.0, dev_state: Optional[torch.Tensor] = None):
        """dev_state (gru_vae.DeviceRng.state): take the step count from the device (state[2] + 1) instead of the host --
        the form a captured CUDA graph needs; the caller advances it (DeviceRng.end_step(step_delta=1))."""
        self.t += 1
        with torch.cuda.device(self.flat.device):
            check(lib.cvb_adam_step(self.n, ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.lr,
                                    self.betas[0], self.betas[1], self.eps, self.t, None if dev_state is None else dev_state.data_ptr(),
                                    grad_scale, torch.cuda.current_stream().cuda_stream), "cvb_adam_step")


class NativeComm:
    """The library's own NCCL communicator (cvb_comm_init / cvb_allreduce_sum of the C ABI): one per rank, created on the
    current CUDA device.  The 128-byte NCCL id is made by rank 0 and handed to the other ranks by the caller's side
    channel -- here torch.distributed's object broadcast when a process group is up (the plumbing; the collective on
    the data path is the library's)."""

    def __init__(self, rank: int, world: int, unique_id: Optional[bytes] = None):
        import ctypes as C
        if unique_id is None:
            import torch.distributed as dist
            box = [None]
            if rank == 0:
                buf = C.create_string_buffer(128)
                check(lib.cvb_comm_unique_id(buf), "cvb_comm_unique_id")
                box[0] = buf.raw
            dist.broadcast_object_list(box, src=0)
            unique_id = box[0]
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        check(lib.cvb_comm_init(C.byref(self._h), C.c_char_p(unique_id), rank, world), "cvb_comm_init")

    def allreduce_sum(self, flat: torch.Tensor) -> None:
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
        with torch.cuda.device(flat.device):
            check(lib.cvb_allreduce_sum(self._h, flat.data_ptr(), flat.numel(), torch.cuda.current_stream().cuda_stream), "cvb_allreduce_sum")

    def close(self) -> None:
        if self._h:
            check(lib.cvb_comm_destroy(self._h), "cvb_comm_destroy")
            self._h = None


def allreduce_grads(flat_grad: torch.Tensor, comm: Optional[NativeComm] = None) -> None:
    """The only collective of the data-parallel path: SUM (not mean -- the reference sums the
    per-utterance losses, train_*.py:1403,1408, so summed shard gradients equal the gradient of one
    process holding every shard's utterances).  comm: the library's own communicator (cvb_allreduce_sum);
    None: torch.distributed's NCCL process group when one is initialised."""
    if comm is not None:
        if comm.world > 1:
            comm.allreduce_sum(flat_grad)
        return
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)


class CycleStep:
    """one optimisation step of the trainer (train_*.py:1298-1420: zero_grad, the 5 x n_cyc GRU_RNN
    passes, loss assembly, backward, [gradient all-reduce], Adam) as ONE call.  Forward + losses + BPTT are captured in a
    CUDA graph and replayed; dropout masks and latent noise are drawn from a device-resident Philox state
    (gru_vae.DeviceRng), the Adam step count lives on the device too, parameter gradients are accumulated by the kernels
    straight into the flat gradient buffer, and the only host work per step is one graph launch, the all-reduce (N > 1)
    and the Adam launch.  Inputs are copied into static device buffers (the copies are part of the caller's stream).

    state=None runs first-chunk semantics (initial y_in, h_in = None: train_*.py:1326-1338) -- the benchmark workload.
    flens: int32 [B] frames of each utterance that count in the losses (flen_acc; 0 = utterance not selected)."""

    def __init__(self, enc: gv.GRU_RNN, dec: gv.GRU_RNN, opt: FlatAdam, *, B: int, T: int, n_cyc: int, lat_dim: int, stdim: int,
                 n_spk: int, y0_enc: torch.Tensor, y0_dec: torch.Tensor, graph: bool = True, kl_cv_quirk: bool = True,
                 comm: Optional[NativeComm] = None):
        dev = opt.flat.device
        self.enc, self.dec, self.opt, self.comm = enc, dec, opt, comm
        self.n_cyc, self.lat_dim, self.stdim, self.kl_cv_quirk = n_cyc, lat_dim, stdim, kl_cv_quirk
        self.x = torch.zeros(B, T, enc.in_dim, device=dev)
        self.cv = torch.zeros(B, T, stdim, device=dev)
        self.sc = torch.zeros(B, T, n_spk, device=dev)
        self.tc = torch.zeros(B, T, n_spk, device=dev)
        self.flens = torch.full((B,), T, dtype=torch.int32, device=dev)
        self.y0_enc, self.y0_dec = y0_enc.to(dev).contiguous(), y0_dec.to(dev).contiguous()
        self.loss = torch.zeros((), device=dev)
        self.rng = gv.DeviceRng(dev)
        self.sel = list(range(B))
        self.graph = None
        self.kernels_per_replay = 0
        if graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):   # warm-up on a side stream: sizes every workspace before the capture
                for _ in range(2):
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            n0 = lib.cvb_launch_count()
            with torch.cuda.graph(self.graph):
                self._body()
            self.kernels_per_replay = int(lib.cvb_launch_count() - n0)   # this library's kernels inside one replay

    def _body(self):
        # the parameters are new at every step: the first product that uses a parameter matrix refreshes its 16-bit image
        # (inside the captured graph too), the other passes of the step reuse it
        lib.cvb_weights_changed()
        self.opt.zero_grad()
        with gv.device_rng(self.rng):
            out, _ = cyc_forward(self.enc, self.dec, x=self.x, cv=self.cv, src_code=self.sc, trg_code=self.tc, n_cyc=self.n_cyc,
                                 lat_dim=self.lat_dim, stdim=self.stdim, y0_enc=self.y0_enc, y0_dec=self.y0_dec, do=True)
            loss, _ = cyc_loss(out, self.x, n_cyc=self.n_cyc, lat_dim=self.lat_dim, stdim=self.stdim, flen_acc=None,
                               select_utt_idx=self.sel, kl_cv_quirk=self.kl_cv_quirk, flens_dev=self.flens)
            loss.backward()
        self.rng.end_step()
        self.loss.copy_(loss.detach())

    def step(self, x, cv, src_code, trg_code, flens: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on this batch; returns the (device) loss scalar of the step."""
        for dst, src in ((self.x, x), (self.cv, cv), (self.sc, src_code), (self.tc, trg_code)):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        if flens is not None:
            self.flens.copy_(flens, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        allreduce_grads(self.opt.grad, self.comm)
        self.opt.step(dev_state=self.rng.state)
        check(lib.cvb_state_advance(self.rng.state.data_ptr(), 0, 1, torch.cuda.current_stream().cuda_stream), "cvb_state_advance")
        return self.loss


def shard_utterances(n_utt: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous utterance shard [lo, hi) of rank `rank` (np.array_split semantics, the split
    decode_*.py:190 applies to file lists)."""
    base, rem = divmod(n_utt, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def convert(enc: gv.GRU_RNN, dec: gv.GRU_RNN, feat, code, *, lat_dim: int, y0_enc, y0_dec, eps_mean=None, n_smpl: int = 300):
    """decode_*.py:303-305,318: ENC -> latent averaged over n_smpl samples -> DEC with the target
    code.  mean_k(mu + e^{s/2} eps_k) == mu + e^{s/2} mean_k(eps_k): the averaged noise is drawn
    (or given as eps_mean) once instead of materialising [n_smpl, T, 2*lat] as the reference does.
    feat: [T,in] (reference layout) or [B,T,in]."""
    lat, _, _ = enc(feat, y0_enc, clamp_vae=True, lat_dim=lat_dim)
    if eps_mean is None:
        shp = lat.shape[:-1] + (lat_dim,)
        eps_mean = torch.randn(shp, device=lat.device) / float(n_smpl) ** 0.5
    z = gv.reparam_concat(lat, code, eps_mean, lat_dim)
    cvm, _, _ = dec(z, y0_dec)
    return cvm


# ------------------------------------------------------------------------------------------------
def scale_in_mean(enc: gv.GRU_RNN) -> torch.Tensor:
    """The feature mean the encoder's frozen scale_in layer subtracts (train_*.py:344-345: weight = diag(1/std),
    bias = -mean/std): x = mean is what the zero padding of the conv front-end looks like in the input domain
    (gru_vae.py:336,357: pads are zeros of the NORMALISED features)."""
    if not enc.scale_in_flag:
        return torch.zeros(enc.in_dim, device=enc.out_1.weight.device)
    w = enc.scale_in.weight.detach()[:, :, 0]
    return -enc.scale_in.bias.detach() / torch.diagonal(w)


def pack_utterances(feats: Sequence[torch.Tensor], pad_value: torch.Tensor):
    """[T_i, C] utterances -> ([N, T_max, C] padded with pad_value [C], lengths).  Padding frames hold the value the
    front-end's own zero padding stands for, so the two-sided convolution sees at an utterance's end exactly what it sees
    when the utterance is run alone; the recurrence is causal, so frames t < T_i never depend on the padding after them."""
    lens = [int(f.shape[0]) for f in feats]
    x = pad_value.to(feats[0].device, torch.float32).repeat(len(feats), max(lens), 1)
    for i, f in enumerate(feats):
        x[i, :lens[i]] = f
    return x, lens


def convert_utterances(enc: gv.GRU_RNN, dec: gv.GRU_RNN, feats: Sequence[torch.Tensor], trg_code: torch.Tensor, *,
                       lat_dim: int, y0_enc, y0_dec, eps_means: Optional[Sequence[torch.Tensor]] = None,
                       n_smpl: int = 300, rows_per_group: Optional[int] = None) -> List[torch.Tensor]:
    """Stage-6 conversion (decode_*.py:303-305,318) of MANY utterances of different lengths at once — the reference
    converts one utterance per call.  Utterances are sorted by length and packed into groups of the row count one
    persistent launch holds (so a group runs only to ITS longest utterance), padded as `pack_utterances` describes; the
    decoder input of padding frames is zeroed (its front-end has no scale_in: zero IS its padding).  Returns the
    converted mcep [T_i, out] per utterance in the caller's order; equal to converting each utterance alone.
    trg_code: [n_spk] one-hot; y0_enc [1,1,2*lat], y0_dec [1,1,out]; eps_means[i]: [T_i, lat] averaged noise (else drawn)."""
    n = len(feats)
    if n == 0:
        return []
    dev = feats[0].device
    if rows_per_group is None:
        rows_per_group = min(enc.max_rows_per_launch(0), dec.max_rows_per_launch(0))
    order = sorted(range(n), key=lambda i: -int(feats[i].shape[0]))
    mean = scale_in_mean(enc)
    out: List[Optional[torch.Tensor]] = [None] * n
    code = trg_code.to(dev, torch.float32).reshape(1, 1, -1)
    for g0 in range(0, n, rows_per_group):
        idx = order[g0:g0 + rows_per_group]
        x, lens = pack_utterances([feats[i] for i in idx], mean)
        B, T = x.shape[0], x.shape[1]
        valid = (torch.arange(T, device=dev).unsqueeze(0) < torch.tensor(lens, device=dev).unsqueeze(1)).unsqueeze(2)
        lat, _, _ = enc(x, y0_enc.expand(B, -1, -1).contiguous(), clamp_vae=True, lat_dim=lat_dim)
        if eps_means is None:
            eps = torch.randn(B, T, lat_dim, device=dev) / float(n_smpl) ** 0.5
        else:
            eps = torch.zeros(B, T, lat_dim, device=dev)
            for r, i in enumerate(idx):
                eps[r, :lens[r]] = eps_means[i]
        z = gv.reparam_concat(lat, code.expand(B, T, -1).contiguous(), eps, lat_dim) * valid
        cvm, _, _ = dec(z, y0_dec.expand(B, -1, -1).contiguous())
        for r, i in enumerate(idx):
            out[i] = cvm[r, :lens[r]]
    return out  # type: ignore[return-value]


def gv_postfilter(cvmcep: torch.Tensor, gv_mean_trg: torch.Tensor, cvgv_mean: torch.Tensor) -> torch.Tensor:
    """decode_*.py:419-420 on the device: coefficients 1.. are rescaled around their utterance mean by
    sqrt(gv_mean_trg / cvgv_mean) (target-speaker global variance over the variance of converted training data);
    c0 passes through.  cvmcep [T, D]; the statistics [D-1]."""
    rest = cvmcep[:, 1:]
    m = rest.mean(0, keepdim=True)
    return torch.cat((cvmcep[:, :1], torch.sqrt(gv_mean_trg / cvgv_mean).to(rest) * (rest - m) + m), 1)


def cvgv_stats(converted: Sequence[torch.Tensor]):
    """calc_cvgv_*.py:203,320-321: per utterance the (population) variance over frames of coefficients 1.., then mean and
    variance of those vectors over the utterances -> (cvgv_mean, cvgv_var), each [D-1]; `cvgv_mean` is what
    `gv_postfilter` divides the target speaker's GV by.  Stays on the device of the inputs (float64 accumulation)."""
    per_utt = torch.stack([c[:, 1:].double().var(dim=0, unbiased=False) for c in converted])
    return per_utt.mean(0), per_utt.var(dim=0, unbiased=False)


# ------------------------------------------------------------------------------------------------
def mcd_aligned(x: torch.Tensor, y: torch.Tensor, idx_x: Optional[torch.Tensor] = None, idx_y: Optional[torch.Tensor] = None):
    """Mean and population std [dB] of the frame-wise mel-cepstral distortion of two aligned [T, D] CUDA tensors -- the
    device-side replacement of `dtw.calc_mcd` at train_*.py:1435-1439 (no trajectory leaves the GPU).  idx_x / idx_y:
    optional int64 frame indices (the speech-frame index_select of the call sites).  Returns a [2] tensor {mean, std}."""
    assert x.is_cuda and y.is_cuda and x.shape[1] == y.shape[1]
    xs, ys = x if x.stride(1) == 1 else x.contiguous(), y if y.stride(1) == 1 else y.contiguous()
    n = int(idx_x.numel()) if idx_x is not None else int(xs.shape[0])
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.cvb_mcd_aligned(n, int(xs.shape[1]), xs.data_ptr(), int(xs.stride(0)), ys.data_ptr(), int(ys.stride(0)),
                                  None if idx_x is None else idx_x.contiguous().data_ptr(), None if idx_y is None else idx_y.contiguous().data_ptr(),
                                  ptr(out), torch.cuda.current_stream().cuda_stream), "cvb_mcd_aligned")
    return out


def dtw_org_to_trg(org: torch.Tensor, trg: torch.Tensor):
    """Dynamic time warping of org [N, D] onto the time axis of trg [M, D] on the device (replaces `dtw.dtw_org_to_trg`,
    train_*.py:679-688): MCD frame distance, symmetric step pattern.  Returns (aligned_org [M, D], path [M] int32 = the
    source frame paired with each target frame, stats [3] = {mean MCD over target frames [dB], path steps, accumulated cost})."""
    assert org.is_cuda and trg.is_cuda and org.shape[1] == trg.shape[1]
    o, t = org.contiguous().float(), trg.contiguous().float()
    N, M, D = int(o.shape[0]), int(t.shape[0]), int(o.shape[1])
    ws = torch.empty(int(lib.cvb_dtw_ws_bytes(N, M)), dtype=torch.uint8, device=o.device)
    path = torch.empty(M, dtype=torch.int32, device=o.device)
    stats = torch.empty(3, dtype=torch.float32, device=o.device)
    with torch.cuda.device(o.device):
        check(lib.cvb_dtw_mcd(N, M, D, ptr(o), D, ptr(t), D, ws.data_ptr(), path.data_ptr(), ptr(stats),
                              torch.cuda.current_stream().cuda_stream), "cvb_dtw_mcd")
    return o[path.long()], path, stats
