"""Frame ingestion: the data path INTO the hot path.

The reference's loader turns every stored frame into a model input on the CPU, one PIL image at a time
(/root/reference/data_loader/data_loader.py:434-446 with the transform of main.py:103-108:
Resize((224, 224)) -> Grayscale(1) -> ToTensor() -> Normalize([0.5], [0.5])), collates fp32 tensors, pins them, and the
trainer copies them to the GPU synchronously (`prepare_batch`, trainer.py:307-311).  Here

  * `FrameTransform`   runs that transform on the GPU (libvideocad_b200.so: `vc_frames_rgb_u8_ingest`, bit-exact with
                       Pillow's 8-bit arithmetic), so a batch crosses PCIe as the bytes the dataset stores -- 3 B per pixel,
                       no CPU transform -- instead of 4 B per grey pixel after it;
  * `cad_to_gray_u8`   keeps the reference's own cv2 code for the ONE target image per sample (BGR -> grey -> resize, uint8;
                       data_loader.py:468-474) and leaves `/ 255` + Normalize to the device (`vc_frames_u8_normalize`);
  * `RawSequenceDataset` is `DatasetBase.__getitem__` (data_loader.py:434-508) minus the CPU transforms: the stored bytes of a sample;
  * `collate_u8`       is `DatasetBase.collate_with_padding` (data_loader.py:319-366) for uint8 samples: pads with the frame
                       count instead of -1 values (the -1 fill is applied on the device after normalisation);
  * `DevicePrefetcher` double-buffers pinned host batches -> device on a copy stream while the previous step computes, and
                       hands the trainer batches that are already on the device and already normalised (its `.to(device)`
                       calls become no-ops).

Nothing here falls back to the CPU for the transform: without the CUDA library `FrameTransform` raises.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Iterator, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as L

PRECISION_BITS = 32 - 8 - 2  # Pillow: 8-bit resampling accumulates in int32 with 22 fractional bits


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pillow's bilinear resampling taps for one axis: precompute_coeffs (triangle filter, support 1.0 widened by the
    down-scaling factor = the antialiasing PIL always applies) followed by normalize_coeffs_8bpc (fixed point, round half away
    from zero).  -> (kk int32 [out, ksize], bounds int32 [out, 2] = (first input index, tap count))."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)      # C (int) cast: truncation; arguments are >= -0.5
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((x + xmin[:, None] - center[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w = np.where(x < xmax[:, None], w, 0.0)
    ww = w.sum(axis=1, keepdims=True)
    w = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    kk = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS)).astype(np.int64).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return np.ascontiguousarray(kk), np.ascontiguousarray(bounds)


class FrameTransform:
    """The reference's `frame_transform` (main.py:103-108) on the device.

        ft = FrameTransform((224, 224))
        x = ft(frames_u8)        # uint8 [..., H, W, 3] RGB on the GPU  ->  fp32 [..., 1, 224, 224] in [-1, 1]
    """

    def __init__(self, size: Tuple[int, int] = (224, 224), mean: float = 0.5, std: float = 0.5, _lib=None):
        self.size, self.mean, self.std = (int(size[0]), int(size[1])), float(mean), float(std)
        self._lib = _lib  # tests: CPU emulation library
        self._tables: Dict[tuple, tuple] = {}

    def _table(self, in_size: int, out_size: int, device):
        key = (in_size, out_size, str(device))
        t = self._tables.get(key)
        if t is None:
            kk, bounds = resample_coeffs(in_size, out_size)
            t = (torch.from_numpy(kk).to(device), torch.from_numpy(bounds).to(device), kk.shape[1])
            self._tables[key] = t
        return t

    def __call__(self, frames_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() < 3 or frames_u8.shape[-1] != 3:
            raise ValueError(f"FrameTransform expects uint8 RGB frames [..., H, W, 3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        if not frames_u8.is_cuda and self._lib is None:
            raise RuntimeError("videocad_b200.ingest.FrameTransform runs on CUDA tensors only; there is no CPU fallback")
        lib = self._lib if self._lib is not None else L.load()
        src = frames_u8.contiguous()
        lead, (Hin, Win) = tuple(src.shape[:-3]), src.shape[-3:-1]
        Hout, Wout = self.size
        n = int(np.prod(lead)) if lead else 1
        dev = src.device
        if out is None:
            out = torch.empty(*lead, 1, Hout, Wout, dtype=torch.float32, device=dev)
        elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != n * Hout * Wout:
            raise ValueError("FrameTransform: `out` must be a contiguous fp32 tensor of n * Hout * Wout elements")
        if n == 0:
            return out
        kk_h = b_h = kk_v = b_v = tmp = None
        ks_h = ks_v = 0
        if Win != Wout:
            kk_h, b_h, ks_h = self._table(Win, Wout, dev)
            tmp = torch.empty(n, Hin, Wout, 3, dtype=torch.uint8, device=dev)
        if Hin != Hout:
            kk_v, b_v, ks_v = self._table(Hin, Hout, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
        L.check(lib.vc_frames_rgb_u8_ingest(src.data_ptr(), n, Hin, Win, Hout, Wout, L.ptr(kk_h), L.ptr(b_h), ks_h, L.ptr(kk_v), L.ptr(b_v),
                                            ks_v, L.ptr(tmp), self.mean, self.std, out.data_ptr(), stream), lib)
        return out


def cad_to_gray_u8(cad_bgr: np.ndarray, image_size: Tuple[int, int] = (224, 224)) -> np.ndarray:
    """The reference's CPU steps for the target CAD image up to (not including) the float conversion
    (data_loader.py:468-470): cv2 BGR -> grey, cv2.resize (bilinear).  One image per sample; `/ 255` and Normalize happen on
    the device (`normalize_u8`)."""
    import cv2

    g = cv2.cvtColor(cad_bgr, cv2.COLOR_BGR2GRAY)
    return cv2.resize(g, image_size)


def normalize_u8(img_u8: torch.Tensor, mean: float = 0.5, std: float = 0.5, _lib=None) -> torch.Tensor:
    """uint8 grey images [..., H, W] on the GPU -> fp32 (u / 255 - mean) / std, bit-exact with ToTensor + Normalize."""
    if img_u8.dtype != torch.uint8:
        raise ValueError("normalize_u8 expects uint8")
    if not img_u8.is_cuda and _lib is None:
        raise RuntimeError("videocad_b200.ingest.normalize_u8 runs on CUDA tensors only; there is no CPU fallback")
    lib = _lib if _lib is not None else L.load()
    src = img_u8.contiguous()
    out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    stream = torch.cuda.current_stream(src.device).cuda_stream if src.is_cuda else None
    L.check(lib.vc_frames_u8_normalize(src.data_ptr(), src.numel(), float(mean), float(std), out.data_ptr(), stream), lib)
    return out


class RawSequenceDataset(torch.utils.data.Dataset):
    """`DatasetBase.__getitem__` (data_loader.py:434-508) without its CPU transforms: a sample is the bytes the dataset stores --
    {"frames": uint8 [n, H, W, 3] (a zero-copy view when the retriever is a `MmapSequenceRetriever`), "actions": [n, 7],
    "cad_image": uint8 [S, S] grey (the reference's own cv2 steps, `cad_to_gray_u8`)} -- for `collate_u8` + `DevicePrefetcher`.

    `sequence_retriever`: anything with the reference's retriever interface (`get_sequence(idx) -> (frames, actions, base_file_id)`,
    `__len__`; sequence_retriver.py:25-36); `image_loader`: the reference's `ImageLoader` (`get_image(base_file_id)` -> BGR array,
    image_loader.py:30-43).  Multi-view samples are not covered (the reference loads them with a third code path, :416-429)."""

    def __init__(self, sequence_retriever, image_loader, image_size: Tuple[int, int] = (224, 224)):
        self.sequence_retriever, self.image_loader, self.image_size = sequence_retriever, image_loader, tuple(image_size)

    def __len__(self) -> int:
        return len(self.sequence_retriever)

    def __getitem__(self, idx: int) -> dict:
        if idx < 0 or idx >= len(self):
            raise IndexError("Index out of range")  # as DatasetBase.__getitem__ (data_loader.py:435-436)
        frames, actions, base_file_id = self.sequence_retriever.get_sequence(idx)
        cad = self.image_loader.get_image(base_file_id)
        if cad is None:
            raise ValueError(f"Missing CAD image for sample {base_file_id}")
        return {"frames": frames, "actions": actions, "cad_image": cad_to_gray_u8(cad, self.image_size)}


def collate_u8(samples: Sequence[dict], pin: bool = True) -> dict:
    """`collate_with_padding` (data_loader.py:319-366) for raw samples {"frames": uint8 [n, H, W, 3], "actions": [n, 7],
    "cad_image": uint8 [S, S]}: pads every sequence to the longest one of the batch.  Frames are padded with zeros and the true
    lengths are returned (the reference's -1 fill is applied after normalisation, on the device); actions are padded with -1
    exactly as the reference does.  Buffers are pinned so that the H2D copies can be asynchronous."""
    max_len = max(int(s["frames"].shape[0]) for s in samples)
    B = len(samples)
    H, W = samples[0]["frames"].shape[1:3]
    frames = torch.zeros(B, max_len, H, W, 3, dtype=torch.uint8)
    actions = torch.full((B, max_len, samples[0]["actions"].shape[1]), -1.0, dtype=torch.float32)
    lengths = torch.zeros(B, dtype=torch.int64)
    cad = torch.stack([torch.from_numpy(np.array(s["cad_image"], dtype=np.uint8)) for s in samples])
    for i, s in enumerate(samples):
        n = int(s["frames"].shape[0])
        np.copyto(frames[i, :n].numpy(), s["frames"])  # straight from the (read-only, memory-mapped) store into the batch buffer
        np.copyto(actions[i, :n].numpy(), s["actions"], casting="same_kind")
        lengths[i] = n
    out = {"frames_u8": frames, "actions": actions, "cad_u8": cad, "lengths": lengths,
           "timesteps": torch.arange(max_len).unsqueeze(0).repeat(B, 1)}
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def transform_batch(b: dict, transform: FrameTransform, mean: float = 0.5, std: float = 0.5, _lib=None) -> dict:
    """A `collate_u8` batch (already on the compute device) -> the dict the reference trainer consumes: frames through the device
    transform, padded positions filled with -1 (the reference pads AFTER its transform, data_loader.py:311-317), CAD image
    normalised."""
    frames = transform(b["frames_u8"])                                        # [B, L, 1, S, S]
    pad = torch.arange(frames.shape[1], device=frames.device)[None, :] >= b["lengths"][:, None]
    frames.masked_fill_(pad[:, :, None, None, None], -1.0)
    return {"frames": frames, "actions": b["actions"], "cad_image": normalize_u8(b["cad_u8"], mean, std, _lib=_lib).unsqueeze(1),
            "timesteps": b["timesteps"]}


class DevicePrefetcher:
    """Iterate over host batches produced by `collate_u8`, yielding the dict the reference trainer consumes
    ({"frames": fp32 [B, L, 1, S, S], "actions", "cad_image": fp32 [B, 1, S, S], "timesteps"}) already on the device:

        for batch in DevicePrefetcher(loader, device):       # loader: any iterable of collate_u8 batches
            loss, metrics = trainer._process_batch(batch)     # prepare_batch's .to(device) calls are no-ops now

    Batch i + 1 is copied (pinned -> device, non_blocking) and transformed on a second stream while step i computes; the
    consumer's stream waits on an event, never on the host."""

    def __init__(self, batches: Iterable[dict], device, size: Tuple[int, int] = (224, 224), mean: float = 0.5, std: float = 0.5):
        self.batches, self.device = batches, torch.device(device)
        self.transform = FrameTransform(size, mean, std)
        self.mean, self.std = mean, std
        self.stream = torch.cuda.Stream(device=self.device)

    def _stage(self, hb: dict):
        with torch.cuda.stream(self.stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in hb.items()}
            batch = transform_batch(dev, self.transform, self.mean, self.std)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return batch, ev

    def __iter__(self) -> Iterator[dict]:
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            batch, ev = nxt
            try:
                nxt = self._stage(next(it))  # overlaps with the consumer's work on `batch`
            except StopIteration:
                nxt = None
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for v in batch.values():
                v.record_stream(cur)
            yield batch
