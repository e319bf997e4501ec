"""Memory-mappable sequence store: the on-disk format either side of the hot path.

The reference keeps one pickle per sample -- `<id>_data.pkl` = {"frames": uint8 [N, H, W, 3], "actions": float64 [N, 7],
"timesteps": ...}, written by generate_dataset.py:194-199 -- and `BaseSequenceRetriever.get_sequence`
(data_loader/sequence_retriver.py:29-33) unpickles the WHOLE sample for every `__getitem__`: a 186-step sample is 28 MB of
frames copied through the pickle machinery per access, on every epoch, in every DataLoader worker.

Here a dataset is ONE file that is opened once and `np.memmap`-ed:

    offset 0      magic  b"VCADSEQ1"
           8      uint64 header_bytes (JSON header length)
           16     JSON header: version, H, W, C, act_dim, frame dtype, action dtype, samples = [{id, n, frame_off, action_off,
                  time_off}], optional per-sample timesteps
           4096-aligned blobs: frames of sample 0 (n0*H*W*C bytes), ..., actions of every sample, timesteps of every sample

`get_sequence(idx)` returns zero-copy views (frames [N,H,W,3] uint8, actions [N,7] float64) into the page cache: no
deserialisation, the OS reads only the pages a window of the sequence touches, and all DataLoader workers share them.
`MmapSequenceRetriever` has the interface of the reference's `BaseSequenceRetriever` (`get_sequence(idx)` ->
(frames, actions, base_file_id), `__len__`), so `DatasetBase` can take it in place of the pickle retriever
(data_loader.py:205-207 `load_retriever`).  Values are bit-identical to the pickles (tests/test_sequence_store.py).
"""
from __future__ import annotations

import json
import os
import pickle
import struct
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

MAGIC = b"VCADSEQ1"
ALIGN = 4096


def _align(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


def sample_id_of(data_file: str) -> str:
    """`0001234_data.pkl` -> `0001234` (the id the reference derives from file names, data_loader.py:302-310)."""
    return os.path.basename(data_file).split("_")[0]


def convert_pickles(data_files: Sequence[str], out_path: str, ids: Optional[Sequence[str]] = None) -> dict:
    """Write the store from the reference's per-sample pickles (streams one sample at a time).  Returns the header."""
    data_files = list(data_files)
    ids = list(ids) if ids is not None else [sample_id_of(f) for f in data_files]
    if len(ids) != len(data_files):
        raise ValueError("ids and data_files must have the same length")
    if len(set(ids)) != len(ids):
        dup = sorted({i for i in ids if ids.count(i) > 1})
        raise ValueError(f"duplicate sample ids {dup[:5]}: samples are looked up by id")
    # pass 1: shapes (a sample is unpickled twice; conversion is a one-off)
    metas = []
    H = W = C = act_dim = None
    fdtype = adtype = None
    for f in data_files:
        with open(f, "rb") as fh:
            d = pickle.load(fh)
        fr, ac = np.asarray(d["frames"]), np.asarray(d["actions"])
        if fr.ndim != 4 or ac.ndim != 2 or fr.shape[0] != ac.shape[0]:
            raise ValueError(f"{f}: frames {fr.shape} / actions {ac.shape} are not [N,H,W,C] / [N,A] with equal N")
        shape = (fr.shape[1], fr.shape[2], fr.shape[3], ac.shape[1], str(fr.dtype), str(ac.dtype))
        if H is None:
            H, W, C, act_dim, fdtype, adtype = shape
        elif shape != (H, W, C, act_dim, fdtype, adtype):
            raise ValueError(f"{f}: sample layout {shape} differs from the first sample's {(H, W, C, act_dim, fdtype, adtype)}")
        ts = np.asarray(d.get("timesteps", np.arange(fr.shape[0])), dtype=np.float64).reshape(-1)
        metas.append(dict(n=int(fr.shape[0]), nt=int(ts.shape[0])))
    fbytes = np.dtype(fdtype).itemsize * H * W * C
    abytes = np.dtype(adtype).itemsize * act_dim
    header = dict(version=1, H=H, W=W, C=C, act_dim=act_dim, frame_dtype=fdtype, action_dtype=adtype, samples=[])
    # offsets are relative to the start of the data area so that the header length does not feed back into them
    off = 0
    for sid, m in zip(ids, metas):
        header["samples"].append(dict(id=sid, n=m["n"], nt=m["nt"], frame_off=off))
        off = _align(off + m["n"] * fbytes)
    for s in header["samples"]:
        s["action_off"] = off
        off = _align(off + s["n"] * abytes)
    for s in header["samples"]:
        s["time_off"] = off
        off = _align(off + s["nt"] * 8)
    blob = json.dumps(header).encode()
    data_start = _align(16 + len(blob))
    tmp = out_path + ".tmp"
    with open(tmp, "wb") as out:
        out.write(MAGIC)
        out.write(struct.pack("<Q", len(blob)))
        out.write(blob)
        out.truncate(data_start + off)
        for f, s in zip(data_files, header["samples"]):
            with open(f, "rb") as fh:
                d = pickle.load(fh)
            out.seek(data_start + s["frame_off"])
            out.write(np.ascontiguousarray(d["frames"], dtype=fdtype).tobytes())
            out.seek(data_start + s["action_off"])
            out.write(np.ascontiguousarray(d["actions"], dtype=adtype).tobytes())
            out.seek(data_start + s["time_off"])
            ts = np.asarray(d.get("timesteps", np.arange(s["n"])), dtype=np.float64).reshape(-1)
            out.write(ts.tobytes())
    os.replace(tmp, out_path)
    return header


def convert_dataset_dir(dataset_path: str, out_path: str) -> dict:
    """Convert every `*_data.pkl` under `dataset_path`, in the order DatasetBase.create_data_and_image_files uses (sorted paths)."""
    files = []
    for root, _, names in os.walk(dataset_path):
        files += [os.path.join(root, n) for n in names if n.endswith("_data.pkl")]
    files.sort()
    if not files:
        raise ValueError(f"no *_data.pkl under {dataset_path}")
    return convert_pickles(files, out_path)


class SequenceStore:
    """Read side: one np.memmap over the file, zero-copy views per sample."""

    def __init__(self, path: str):
        self._open(path)

    # DataLoader workers started with `spawn` (and anything else that pickles the dataset) must re-open the mapping: pickling an
    # np.memmap would serialise the whole file
    def __getstate__(self):
        return {"path": self.path}

    def __setstate__(self, state):
        self._open(state["path"])

    def _open(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{path}: not a videocad_b200 sequence store")
            (n,) = struct.unpack("<Q", f.read(8))
            self.header = json.loads(f.read(n).decode())
        if self.header.get("version") != 1:
            raise ValueError(f"{path}: unsupported store version {self.header.get('version')}")
        self._data_start = _align(16 + n)
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        h = self.header
        self._fshape = (h["H"], h["W"], h["C"])
        self._fdtype, self._adtype = np.dtype(h["frame_dtype"]), np.dtype(h["action_dtype"])
        self.ids: List[str] = [s["id"] for s in h["samples"]]
        self._index = {sid: i for i, sid in enumerate(self.ids)}
        per = int(np.prod(self._fshape)) * self._fdtype.itemsize
        end = max([0] + [max(s["frame_off"] + s["n"] * per, s["action_off"] + s["n"] * h["act_dim"] * self._adtype.itemsize,
                             s["time_off"] + s["nt"] * 8) for s in h["samples"]])
        if self._data_start + end > self._mm.shape[0]:
            raise ValueError(f"{path}: truncated store ({self._mm.shape[0]} bytes, the index needs {self._data_start + end})")

    def __len__(self) -> int:
        return len(self.ids)

    def index_of(self, sample_id: str) -> int:
        return self._index[sample_id]

    def length(self, idx: int) -> int:
        return self.header["samples"][idx]["n"]

    def _view(self, off: int, count: int, dtype: np.dtype, shape: Tuple[int, ...]) -> np.ndarray:
        start = self._data_start + off
        return self._mm[start:start + count * dtype.itemsize].view(dtype).reshape(shape)

    def frames(self, idx: int, lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
        """uint8 [hi-lo, H, W, C] view of sample idx (only the touched pages are read from disk)."""
        s = self.header["samples"][idx]
        hi = s["n"] if hi is None else min(hi, s["n"])
        lo = max(0, min(lo, hi))
        per = int(np.prod(self._fshape))
        return self._view(s["frame_off"] + lo * per * self._fdtype.itemsize, (hi - lo) * per, self._fdtype, (hi - lo,) + self._fshape)

    def actions(self, idx: int) -> np.ndarray:
        s = self.header["samples"][idx]
        return self._view(s["action_off"], s["n"] * self.header["act_dim"], self._adtype, (s["n"], self.header["act_dim"]))

    def timesteps(self, idx: int) -> np.ndarray:
        s = self.header["samples"][idx]
        return self._view(s["time_off"], s["nt"], np.dtype(np.float64), (s["nt"],))

    def get(self, idx: int):
        return self.frames(idx), self.actions(idx), self.ids[idx]


class MmapSequenceRetriever:
    """Drop-in for the reference's BaseSequenceRetriever (data_loader/sequence_retriver.py:25-36): same constructor
    arguments and return values, backed by a SequenceStore instead of one pickle per sample.  `data_files` / `image_files`
    keep their meaning (the sample order); samples are looked up in the store by id."""

    def __init__(self, data_files: Iterable[str], image_files: Iterable[str], store: "SequenceStore | str"):
        self.data_files = list(data_files)
        self.image_files = list(image_files)
        self.store = SequenceStore(store) if isinstance(store, str) else store
        self._rows = [self.store.index_of(sample_id_of(f)) for f in self.data_files]

    def get_sequence(self, idx):
        frames, actions, _ = self.store.get(self._rows[idx])
        base_file_id = os.path.basename(self.image_files[idx]).split("_")[0]
        return frames, actions, base_file_id

    def __len__(self):
        return len(self.data_files)
