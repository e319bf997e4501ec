"""Frame ingestion (videocad_b200/ingest.py + csrc/ingest.cu) against the reference's own loader code:
torchvision's Resize -> Grayscale -> ToTensor -> Normalize on PIL images (the `frame_transform` of main.py:103-108) and, end to
end, the unmodified `DatasetBase` + `collate_with_padding` (data_loader/data_loader.py:205-508) on an on-disk dataset.
Everything is byte / integer work up to the last two fp32 operations: the bar is bit-exact."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from oracle import reference_model as rm
from videocad_b200 import ingest
from videocad_b200 import lib as L

SIZES = [(224, 224), (37, 53), (448, 448), (300, 500), (224, 100), (100, 224), (225, 223), (720, 1280)]


def _reference_transform():
    from torchvision import transforms

    return transforms.Compose([transforms.Resize((224, 224)), transforms.Grayscale(1), transforms.ToTensor(),
                               transforms.Normalize([0.5], [0.5])])  # main.py:103-108


def _check_transform(device, lib=None):
    from PIL import Image

    ref_t = _reference_transform()
    ft = ingest.FrameTransform((224, 224), _lib=lib)
    rng = np.random.default_rng(0)
    for (h, w) in SIZES:
        n = 3 if h * w < 500 * 500 else 1
        fr = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        fr[0, : h // 2] = 255  # saturated / flat regions: the clip8 and rounding paths
        fr[-1, :, : w // 3] = 0
        want = torch.stack([ref_t(Image.fromarray(f)) for f in fr])
        got = ft(torch.from_numpy(fr).to(device)).cpu()
        assert got.shape == want.shape == (n, 1, 224, 224)
        assert torch.equal(got, want), f"{(h, w)}: {(got - want).abs().max().item()} max|d|, {(got != want).sum().item()} pixels differ"
    # leading batch dimensions and the empty batch
    fr = torch.from_numpy(rng.integers(0, 256, size=(2, 3, 64, 96, 3), dtype=np.uint8)).to(device)
    assert ft(fr).shape == (2, 3, 1, 224, 224)
    assert ft(fr[:0]).shape == (0, 3, 1, 224, 224)
    with pytest.raises(ValueError):
        ft(fr.float())
    with pytest.raises(ValueError):
        ft(fr[..., :2])


def test_coefficient_tables_follow_pillow():
    kk, b = ingest.resample_coeffs(448, 224)  # exact 2x down-scaling: 5 taps, window [2x-1, 2x+3) clipped
    assert kk.shape == (224, 5) and b[0].tolist() == [0, 3] and b[100].tolist() == [199, 4]
    assert np.all(np.abs(kk.sum(1) - (1 << 22)) <= 3)  # fixed-point weights sum to one (up to rounding of each tap)
    kk, b = ingest.resample_coeffs(100, 224)  # up-scaling: support 1 -> 3 taps
    assert kk.shape == (224, 3) and int(b[:, 1].max()) <= 3


def test_frame_transform_cpu_restatement_is_bit_exact():
    from oracle import build_emu

    _check_transform("cpu", lib=L.load(build_emu.build(), require_cuda_build=False))


def test_frame_transform_refuses_cpu_tensors_without_the_test_hook():
    with pytest.raises(RuntimeError):
        ingest.FrameTransform()(torch.zeros(1, 32, 32, 3, dtype=torch.uint8))


@pytest.mark.gpu
def test_frame_transform_gpu_is_bit_exact():
    _check_transform("cuda")


def _write_dataset(root, lengths, H, W, seed=0):
    """<root>/<id[:4]>/<id>_data.pkl + <id>_frame.png, the layout DatasetBase walks (data_loader.py:302-317, image_loader.py:30-43)."""
    import cv2

    rng = np.random.default_rng(seed)
    for i, n in enumerate(lengths):
        sid = f"{i + 3:08d}"
        d = os.path.join(root, sid[:4])
        os.makedirs(d, exist_ok=True)
        frames = rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)
        actions = np.full((n, 7), -1.0)
        actions[:, 0] = rng.integers(0, 5, size=n)
        actions[:, 1:3] = rng.integers(0, 1000, size=(n, 2))
        actions[0] = 0.0
        with open(os.path.join(d, f"{sid}_data.pkl"), "wb") as f:
            pickle.dump({"frames": frames, "actions": actions, "timesteps": np.arange(n)}, f)
        cv2.imwrite(os.path.join(d, f"{sid}_frame.png"), rng.integers(0, 256, size=(150 + 10 * i, 200, 3), dtype=np.uint8))


def _end_to_end(tmp_path, device, lib=None, H=224, W=224):
    """reference: DatasetBase.__getitem__ + collate_with_padding with main.py's transforms;
    here: SequenceStore -> RawSequenceDataset -> collate_u8 -> transform_batch (device)."""
    from torchvision import transforms

    from videocad_b200.sequence_store import MmapSequenceRetriever, convert_dataset_dir

    rm._prepare_path()
    for mod in [m for m in sys.modules if m == "data_loader" or m.startswith("data_loader.")]:
        sys.modules.pop(mod)
    from data_loader.data_loader import DatasetBase  # type: ignore  (the reference's module)

    root = str(tmp_path / "ds")
    _write_dataset(root, [3, 7, 5], H, W)
    ds = DatasetBase(root, frame_transform=_reference_transform(), image_transform=transforms.Normalize(mean=[0.5], std=[0.5]),
                     image_size=(224, 224), image_dir=root)
    want = ds.collate_with_padding([ds[i] for i in range(len(ds))])
    store_path = str(tmp_path / "ds.vcseq")
    convert_dataset_dir(root, store_path)
    retr = MmapSequenceRetriever(ds.data_files, ds.image_files, store_path)
    raw = ingest.RawSequenceDataset(retr, ds.image_loader, (224, 224))  # the reference's own ImageLoader finds the CAD images
    assert len(raw) == len(ds)
    with pytest.raises(IndexError):
        raw[len(raw)]
    samples = [raw[i] for i in range(len(raw))]
    assert samples[1]["frames"].dtype == np.uint8 and not samples[1]["frames"].flags.owndata  # still a view into the mapping
    hb = ingest.collate_u8(samples, pin=False)
    loader = torch.utils.data.DataLoader(raw, batch_size=len(raw), collate_fn=lambda b: ingest.collate_u8(b, pin=False))
    for k, v in next(iter(loader)).items():  # the same batch through a torch DataLoader
        assert torch.equal(v, hb[k]), k
    got = ingest.transform_batch({k: v.to(device) for k, v in hb.items()}, ingest.FrameTransform((224, 224), _lib=lib), _lib=lib)
    for k in ("frames", "actions", "cad_image", "timesteps"):
        assert got[k].shape == want[k].shape, k
        assert torch.equal(got[k].cpu(), want[k]), k
    return hb, want


@pytest.mark.skipif(not rm.available(), reason="reference sources neither under /root/reference nor staged in oracle/_ref")
@pytest.mark.parametrize("hw", [(224, 224), (96, 160)])
def test_data_path_equals_reference_dataset_cpu_emulation(tmp_path, hw):
    from oracle import build_emu

    _end_to_end(tmp_path, "cpu", lib=L.load(build_emu.build(), require_cuda_build=False), H=hw[0], W=hw[1])


@pytest.mark.gpu
@pytest.mark.skipif(not rm.available(), reason="reference sources neither under /root/reference nor staged in oracle/_ref")
def test_data_path_equals_reference_dataset_gpu_and_prefetcher(tmp_path):
    hb, want = _end_to_end(tmp_path, "cuda")
    # the double-buffered prefetcher delivers the same batches, in order, already on the device
    pinned = {k: v.pin_memory() for k, v in hb.items()}
    seen = 0
    for batch in ingest.DevicePrefetcher([pinned, pinned, pinned], "cuda"):
        assert batch["frames"].is_cuda and batch["frames"].dtype == torch.float32
        for k in ("frames", "actions", "cad_image"):
            assert torch.equal(batch[k].cpu(), want[k]), k
        seen += 1
    assert seen == 3


def test_frame_transform_random_sizes_property():
    """Any stored frame size: the device transform's arithmetic (restated on the CPU in the emulation library behind the same host
    code: Pillow's coefficient tables, fixed-point passes, rgb2l, ToTensor, Normalize) equals torchvision on PIL images bit for bit --
    up- and down-scaling, odd sizes, single rows / columns."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from PIL import Image

    from oracle import build_emu

    lib = L.load(build_emu.build(), require_cuda_build=False)
    ref_t = _reference_transform()
    ft = ingest.FrameTransform((224, 224), _lib=lib)

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(h=st.integers(1, 500), w=st.integers(1, 500), seed=st.integers(0, 2 ** 16))
    def check(h, w, seed):
        fr = np.random.default_rng(seed).integers(0, 256, size=(1, h, w, 3), dtype=np.uint8)
        want = ref_t(Image.fromarray(fr[0])).unsqueeze(0)
        got = ft(torch.from_numpy(fr))
        assert torch.equal(got, want), f"{(h, w)}: {(got != want).sum().item()} pixels differ"

    check()
