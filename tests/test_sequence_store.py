"""Memory-mappable sequence store (videocad_b200/sequence_store.py) against the reference's own
pickle path: `BaseSequenceRetriever.get_sequence` (data_loader/sequence_retriver.py:25-36) on `*_data.pkl` files written the
way generate_dataset.py:194-199 writes them must return bit-identical frames / actions / ids."""
import os
import pickle
import sys

import numpy as np
import pytest

from oracle import reference_model as rm
from videocad_b200.sequence_store import MmapSequenceRetriever, SequenceStore, convert_dataset_dir, convert_pickles


def _write_dataset(root, lengths, H=32, W=32, seed=0):
    """data_resized layout: <root>/<id[:4]>/<id>_data.pkl + <id>_frame.png placeholder (only names matter to the retriever)."""
    rng = np.random.default_rng(seed)
    data_files, image_files, truth = [], [], []
    for i, n in enumerate(lengths):
        sid = f"{i + 17:08d}"
        d = os.path.join(root, sid[:4])
        os.makedirs(d, exist_ok=True)
        frames = rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)
        actions = np.full((n, 7), -1.0)
        actions[:, 0] = rng.integers(0, 5, size=n)
        actions[:, 1:3] = rng.integers(0, 1000, size=(n, 2))
        actions[0] = 0.0
        timesteps = np.arange(n) * 3
        path = os.path.join(d, f"{sid}_data.pkl")
        with open(path, "wb") as f:
            pickle.dump({"frames": frames, "actions": actions, "timesteps": timesteps}, f)
        data_files.append(path)
        image_files.append(os.path.join(d, f"{sid}_frame.png"))
        truth.append((frames, actions, timesteps, sid))
    return data_files, image_files, truth


def test_round_trip_is_bit_exact_and_zero_copy(tmp_path):
    data_files, image_files, truth = _write_dataset(str(tmp_path / "ds"), [1, 5, 186, 2])
    out = str(tmp_path / "ds.vcseq")
    header = convert_dataset_dir(str(tmp_path / "ds"), out)
    assert [s["n"] for s in header["samples"]] == [1, 5, 186, 2]
    st = SequenceStore(out)
    assert len(st) == 4
    for i, (frames, actions, timesteps, sid) in enumerate(truth):
        f, a, got_id = st.get(i)
        assert got_id == sid and f.dtype == np.uint8 and a.dtype == np.float64
        assert f.shape == frames.shape and np.array_equal(f, frames)
        assert np.array_equal(a, actions) and np.array_equal(st.timesteps(i), timesteps.astype(np.float64))
        assert not f.flags.owndata and not f.flags.writeable  # a view into the mapping, read-only
        assert np.array_equal(st.frames(i, 1, 3), frames[1:3])  # windowed access touches only those pages
    assert st.frames(2, 180, 400).shape[0] == 6 and st.frames(0, 5, 9).shape[0] == 0  # ragged / empty windows


def test_errors(tmp_path):
    data_files, image_files, _ = _write_dataset(str(tmp_path / "ds"), [3, 3])
    with open(data_files[1], "wb") as f:  # a sample with another frame size must be refused, not silently mis-indexed
        pickle.dump({"frames": np.zeros((3, 16, 16, 3), np.uint8), "actions": np.zeros((3, 7))}, f)
    with pytest.raises(ValueError, match="differs"):
        convert_pickles(data_files, str(tmp_path / "bad.vcseq"))
    with pytest.raises(ValueError, match="no \\*_data.pkl"):
        convert_dataset_dir(str(tmp_path / "empty"), str(tmp_path / "x.vcseq"))
    p = tmp_path / "junk.vcseq"
    p.write_bytes(b"not a store at all")
    with pytest.raises(ValueError, match="not a videocad_b200 sequence store"):
        SequenceStore(str(p))
    good_files, _, _ = _write_dataset(str(tmp_path / "ds2"), [3, 4])
    with pytest.raises(ValueError, match="duplicate"):  # samples are looked up by id: two files with one id cannot both be stored
        convert_pickles(good_files, str(tmp_path / "dup.vcseq"), ids=["a", "a"])
    out = str(tmp_path / "ok.vcseq")
    convert_pickles(good_files, out)
    size = os.path.getsize(out)
    with open(out, "r+b") as f:
        f.truncate(size - 4096)  # a copy that was cut short must be refused at open time, not fail inside a DataLoader worker
    with pytest.raises(ValueError, match="truncated"):
        SequenceStore(out)


def test_store_pickles_as_a_path_not_as_the_data(tmp_path):
    """DataLoader workers started with `spawn` pickle the dataset (and its retriever): the store must travel as its path and
    re-open the mapping, not serialise the mapped file."""
    data_files, image_files, truth = _write_dataset(str(tmp_path / "ds"), [6, 40])
    out = str(tmp_path / "ds.vcseq")
    convert_pickles(data_files, out)
    r = MmapSequenceRetriever(data_files, image_files, out)
    blob = pickle.dumps(r)
    assert len(blob) < 4096 < os.path.getsize(out)
    r2 = pickle.loads(blob)
    for i, (frames, actions, _, sid) in enumerate(truth):
        f, a, got = r2.get_sequence(i)
        assert got == sid and np.array_equal(f, frames) and np.array_equal(a, actions) and not f.flags.owndata


@pytest.mark.skipif(not rm.available(), reason="reference sources neither under /root/reference nor staged in oracle/_ref")
def test_matches_the_reference_retriever(tmp_path):
    """Same constructor arguments, same return values as the unmodified BaseSequenceRetriever on the same files."""
    rm._prepare_path()
    sys.modules.pop("data_loader", None)
    from data_loader.sequence_retriver import BaseSequenceRetriever  # type: ignore  (the reference's module)

    data_files, image_files, _ = _write_dataset(str(tmp_path / "ds"), [4, 9, 2, 31], seed=3)
    order = [2, 0, 3, 1]  # a split hands the retriever its own ordering of the files
    df, imf = [data_files[i] for i in order], [image_files[i] for i in order]
    out = str(tmp_path / "ds.vcseq")
    convert_pickles(data_files, out)
    ref, got = BaseSequenceRetriever(df, imf), MmapSequenceRetriever(df, imf, out)
    assert len(ref) == len(got) == 4
    for i in range(4):
        rf, ra, rid = ref.get_sequence(i)
        gf, ga, gid = got.get_sequence(i)
        assert rid == gid
        assert rf.dtype == gf.dtype and np.array_equal(rf, gf)
        assert ra.dtype == ga.dtype and np.array_equal(ra, ga)
