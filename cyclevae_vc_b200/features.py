"""HDF5-free feature cache for the callers of the hot path.

The reference keeps every utterance in its own HDF5 file and re-opens three datasets per item
(`src/utils/dataset.py:71-87`: `/feat_org_lf0`, `/cvuvlogf0fil_ap`, `/spcidx_range`), pads each to
`--pad_len` = 2200 frames with `padding()` (`dataset.py:18-26`, `train_*.py:456`) and lets the trainer trim
the batch back to its longest utterance and copy it to the device (`train_*.py:47-63`).  h5py is not part
of this image, and per-item HDF5 opens + 2200-frame pads become the next bottleneck once the step takes
30 ms, so the same three arrays are kept here in ONE flat little-endian file that is memory-mapped:

    "CVBFEAT1" | u64 header bytes | header JSON (utterances: name, spk, per-array dtype/shape/offset) | arrays (64-byte aligned)

`PairDataset.__getitem__` returns the same dictionary as `FeatureDatasetSingleVAE.__getitem__`
(`dataset.py:67-101`, same keys, dtypes, padding); `collate_trimmed` builds the trainer's trimmed batch
directly (no 2200-frame pads are ever materialised) and `DeviceStager` moves it through two pinned host
buffers with non-blocking copies.  Pure host code: nothing here touches the CUDA library.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

MAGIC = b"CVBFEAT1"
KEYS = ("feat_org_lf0", "cvuvlogf0fil_ap", "spcidx_range")
_ALIGN = 64


def write_pack(path: str, utterances: Sequence[dict]) -> None:
    """utterances: dicts with `name`, `spk` and the arrays of KEYS (feat_org_lf0 [T,D] float, cvuvlogf0fil_ap [T,4]
    float, spcidx_range [n] or [1,n] int — the reference stores it as [1,n] and reads `[0]`, dataset.py:78)."""
    metas, blobs, off = [], [], 0
    for u in utterances:
        m = {"name": str(u["name"]), "spk": str(u["spk"]), "arrays": {}}
        for k in KEYS:
            a = np.asarray(u[k])
            if k == "spcidx_range":
                a = a.reshape(-1).astype("<i8")
            else:
                a = np.ascontiguousarray(a, dtype="<f4")
            off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
            m["arrays"][k] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": off}
            blobs.append((off, a.tobytes()))
            off += a.nbytes
        metas.append(m)
    header = json.dumps({"utterances": metas}).encode()
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(header)))
        f.write(header)
        base = (f.tell() + _ALIGN - 1) // _ALIGN * _ALIGN
        for o, b in blobs:
            f.seek(base + o)
            f.write(b)
    os.replace(tmp, path)


class FeaturePack:
    """Memory-mapped reader of a file written by `write_pack`."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path}: not a cyclevae_vc_b200 feature pack")
            (n,) = struct.unpack("<Q", f.read(8))
            self._meta = json.loads(f.read(n).decode())["utterances"]
            self._base = (f.tell() + _ALIGN - 1) // _ALIGN * _ALIGN
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        self._index = {m["name"]: i for i, m in enumerate(self._meta)}
        if len(self._index) != len(self._meta):
            raise ValueError(f"{path}: duplicate utterance names")

    def __len__(self) -> int:
        return len(self._meta)

    @property
    def names(self) -> List[str]:
        return [m["name"] for m in self._meta]

    def index(self, name: str) -> int:
        return self._index[name]

    def spk(self, i: int) -> str:
        return self._meta[i]["spk"]

    def array(self, i: int, key: str) -> np.ndarray:
        a = self._meta[i]["arrays"][key]
        dt = np.dtype(a["dtype"])
        n = int(np.prod(a["shape"])) if a["shape"] else 1
        start = self._base + a["offset"]
        return self._mm[start:start + n * dt.itemsize].view(dt).reshape(a["shape"])


def padding(x: np.ndarray, flen: int, value=0) -> np.ndarray:
    """dataset.py:18-26: pad with `value` at the end up to flen rows (longer inputs pass through unchanged)."""
    diff = flen - x.shape[0]
    if diff > 0:
        pad = np.full((diff,) + x.shape[1:], value, dtype=np.float64)
        x = np.concatenate([x, pad])
    return x


class PairDataset(torch.utils.data.Dataset):
    """`FeatureDatasetSingleVAE` (dataset.py:55-101) over a FeaturePack: item idx pairs utterance `src[idx]` with
    `src_trg[idx]` (the same sentence of the other speaker); speaker codes are 2-way one-hot rows keyed on
    `spk == spk_src` (dataset.py:73-80)."""

    def __init__(self, pack: FeaturePack, src: Sequence[str], src_trg: Sequence[str], spk_src: str, pad_len: int = 2200):
        if len(src) != len(src_trg):
            raise ValueError("src and src_trg lists differ in length")
        self.pack, self.src, self.src_trg, self.spk_src, self.pad_len = pack, list(src), list(src_trg), spk_src, int(pad_len)

    def __len__(self) -> int:
        return len(self.src)

    def raw(self, idx: int) -> dict:
        """Un-padded arrays of one item (what `collate_trimmed` consumes)."""
        i, j = self.pack.index(self.src[idx]), self.pack.index(self.src_trg[idx])
        h_src = self.pack.array(i, "feat_org_lf0")
        flen = h_src.shape[0]
        src_code = np.zeros((flen, 2), dtype=np.float32)
        trg_code = np.zeros((flen, 2), dtype=np.float32)
        own = 0 if self.pack.spk(i) == self.spk_src else 1
        src_code[:, own] = 1
        trg_code[:, 1 - own] = 1
        return {"h_src": h_src, "src_code": src_code, "trg_code": trg_code, "cv_src": self.pack.array(i, "cvuvlogf0fil_ap"),
                "spcidx_src": self.pack.array(i, "spcidx_range"), "h_src_trg": self.pack.array(j, "feat_org_lf0"),
                "spcidx_src_trg": self.pack.array(j, "spcidx_range"), "featfile_src": self.src[idx], "featfile_src_trg": self.src_trg[idx]}

    def __getitem__(self, idx: int) -> dict:
        r = self.raw(idx)

        def fl(a):
            return torch.FloatTensor(padding(np.asarray(a), self.pad_len, 0.0))

        def lg(a):
            return torch.LongTensor(padding(np.asarray(a), self.pad_len, 0.0))

        return {"h_src": fl(r["h_src"]), "flen_src": r["h_src"].shape[0], "src_code": fl(r["src_code"]), "trg_code": fl(r["trg_code"]),
                "cv_src": fl(r["cv_src"]), "spcidx_src": lg(r["spcidx_src"]), "flen_spc_src": r["spcidx_src"].shape[0],
                "h_src_trg": fl(r["h_src_trg"]), "flen_src_trg": r["h_src_trg"].shape[0], "spcidx_src_trg": lg(r["spcidx_src_trg"]),
                "flen_spc_src_trg": r["spcidx_src_trg"].shape[0], "featfile_src": r["featfile_src"],
                "featfile_src_trg": r["featfile_src_trg"]}


_FLOAT_KEYS = (("h_src", "flen_src"), ("src_code", "flen_src"), ("trg_code", "flen_src"), ("cv_src", "flen_src"),
               ("h_src_trg", "flen_src_trg"))
_LONG_KEYS = (("spcidx_src", "flen_spc_src"), ("spcidx_src_trg", "flen_spc_src_trg"))


def collate_trimmed(ds: PairDataset, idxs: Sequence[int]) -> dict:
    """The batch `train_generator` works on (train_*.py:47-63): every tensor zero-padded to the longest utterance OF
    THE BATCH (`batch[k][:, :max_flen]`), plus the length vectors.  Equal to DataLoader(default collate) over
    `PairDataset.__getitem__` followed by the trainer's trimming, without building the 2200-frame pads."""
    raws = [ds.raw(i) for i in idxs]
    out: Dict[str, object] = {}
    lens = {"flen_src": [r["h_src"].shape[0] for r in raws], "flen_src_trg": [r["h_src_trg"].shape[0] for r in raws],
            "flen_spc_src": [r["spcidx_src"].shape[0] for r in raws], "flen_spc_src_trg": [r["spcidx_src_trg"].shape[0] for r in raws]}
    for k, v in lens.items():
        out[k] = torch.tensor(v, dtype=torch.int64)
    for keys, dtype in ((_FLOAT_KEYS, torch.float32), (_LONG_KEYS, torch.int64)):
        for k, lk in keys:
            T = max(lens[lk])
            first = np.asarray(raws[0][k])
            buf = torch.zeros((len(raws), T) + first.shape[1:], dtype=dtype)
            for b, r in enumerate(raws):
                a = np.asarray(r[k])
                buf[b, :a.shape[0]] = torch.from_numpy(np.array(a)).to(dtype)   # np.array: private copy of the read-only map
            out[k] = buf
    out["featfile_src"] = [r["featfile_src"] for r in raws]
    out["featfile_src_trg"] = [r["featfile_src_trg"] for r in raws]
    return out


class DeviceStager:
    """Host -> device staging of collated batches through two alternating sets of pinned buffers: the copy of batch
    k+1 is enqueued (non-blocking, on `stream`) while batch k is being consumed.  Pinned buffers grow on demand and are
    reused; a slot is recycled only after the copies issued from it have completed (event per slot)."""

    def __init__(self, device: torch.device, stream: Optional[torch.cuda.Stream] = None):
        self.device = torch.device(device)
        self.stream = stream
        self._slots: List[Dict[str, torch.Tensor]] = [{}, {}]
        self._events: List[Optional[torch.cuda.Event]] = [None, None]
        self._turn = 0

    def _pinned(self, slot: Dict[str, torch.Tensor], key: str, like: torch.Tensor) -> torch.Tensor:
        buf = slot.get(key)
        if buf is None or buf.dtype != like.dtype or buf.numel() < like.numel():
            buf = torch.empty(max(like.numel(), 1), dtype=like.dtype, pin_memory=self.device.type == "cuda")
            slot[key] = buf
        return buf[:like.numel()].view(like.shape)

    def put(self, batch: dict) -> dict:
        """Enqueue the copies of `batch` and return the device-side dictionary.  With a side `stream` the CONSUMER stream
        (the stream current at the call) is made to wait for the copies before it may touch the tensors, and the tensors
        are recorded on it so the caching allocator does not hand their memory out while it still reads them."""
        s = self._turn
        self._turn ^= 1
        if self._events[s] is not None:
            self._events[s].synchronize()
        out = {}
        cuda = self.device.type == "cuda"
        consumer = torch.cuda.current_stream(self.device) if cuda else None
        side = self.stream if (cuda and self.stream is not None) else None
        if side is not None:
            side.wait_stream(consumer)   # the pinned slot / earlier device work of the consumer is ordered before the copies
        with (torch.cuda.stream(side) if side is not None else _null()):
            for k, v in batch.items():
                if torch.is_tensor(v) and v.dim() > 1:
                    h = self._pinned(self._slots[s], k, v)
                    h.copy_(v)
                    d = h.to(self.device, non_blocking=True)
                    if side is not None:
                        d.record_stream(consumer)
                    out[k] = d
                else:
                    out[k] = v
            if cuda:
                ev = torch.cuda.Event()
                ev.record()
                self._events[s] = ev
        if side is not None:
            consumer.wait_event(self._events[s])
        return out


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
