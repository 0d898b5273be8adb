"""CPU: the step-front data path's host side — LMDB file walker, latent dataset, rank-strided batches, seed batches."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskdit_b200 import data as D  # noqa: E402
from maskdit_b200.sampler import rank_seed_batches, write_png  # noqa: E402


def test_mdb_roundtrip_small_and_overflow_values(tmp_path):
    rng = np.random.default_rng(0)
    items = {b"length": b"3", b"": b"empty-key"}
    for i in range(700):                                   # enough keys for several leaf pages and a branch level
        items[f"k-{i}".encode()] = rng.bytes(int(rng.integers(0, 200)))
    items[b"big-1"] = rng.bytes(32768)                     # one latent: 8*32*32 fp32 -> overflow pages
    items[b"big-2"] = rng.bytes(4080)                      # exactly one overflow page
    items[b"big-3"] = rng.bytes(4081)                      # spills into a second page
    D.write_mdb(str(tmp_path / "db"), items)
    rd = D.MdbReader(str(tmp_path / "db"))
    assert rd.entries == len(items) and rd.depth >= 2
    for k, v in items.items():
        assert bytes(rd.get(k)) == v, k
    for k in (b"k-", b"k-700", b"zzz", b"a", b"big-", b"k-1000"):
        assert rd.get(k) is None
    rd.close()


def test_mdb_three_levels(tmp_path):
    items = {f"key-{i:07d}-{'x' * 180}".encode(): str(i).encode() for i in range(6000)}
    D.write_mdb(str(tmp_path / "db"), items)
    rd = D.MdbReader(str(tmp_path / "db"))
    assert rd.depth >= 3
    for i in (0, 1, 17, 2999, 5998, 5999):
        k = f"key-{i:07d}-{'x' * 180}".encode()
        assert bytes(rd.get(k)) == str(i).encode()
    assert rd.get(b"key-0006000") is None


def test_latent_dataset_and_rank_strided_batches(tmp_path):
    n, C, R, ncls = 40, 4, 8, 10
    rng = np.random.default_rng(1)
    moments = rng.standard_normal((n, 2 * C, R, R)).astype(np.float32)
    labels = rng.integers(0, ncls, n)
    D.write_latent_lmdb(str(tmp_path), moments, labels)
    ds = D.ImageNetLatentDataset(str(tmp_path), resolution=R, num_channels=C, num_classes=ncls)
    assert len(ds) == n
    z, y = ds[7]
    assert z.shape == (2 * C, R, R) and np.array_equal(z, moments[7])            # datasets.py:287-290
    assert y.dtype == np.float32 and y.sum() == 1 and y[labels[7]] == 1          # helper.get_one_hot
    with pytest.raises(FileNotFoundError):
        D.ImageNetLatentDataset(str(tmp_path / "nope"))
    # world 2, batch 4: step k of rank r reads items (2k + r) * 4 .. +4; 40 // 8 = 5 steps per epoch, then wraps
    seen = []
    for r in range(2):
        it = D.batches(ds, 4, rank=r, world=2, pin=False)
        for k in range(6):
            zb, yb = next(it)
            base = ((k % 5) * 2 + r) * 4
            assert torch.equal(zb, torch.from_numpy(moments[base:base + 4]))
            assert yb.argmax(1).tolist() == labels[base:base + 4].tolist()
            if k < 5:
                seen += list(range(base, base + 4))
    assert sorted(seen) == list(range(n))
    it = D.batches(ds, 4, rank=1, world=2, start=3, pin=False)                   # resume at step 3
    zb, _ = next(it)
    assert torch.equal(zb, torch.from_numpy(moments[28:32]))


def test_rank_seed_batches_and_png(tmp_path):
    seeds = list(range(100, 170))
    per_rank = [rank_seed_batches(seeds, 8, r, 4) for r in range(4)]
    assert sorted(s for rb in per_rank for b in rb for s in b) == seeds
    assert per_rank[1][0] == list(range(106, 112))        # torch.tensor_split(12)[1::4]: sample.py:232-235
    assert rank_seed_batches([], 8) == [] and rank_seed_batches([5], 8, 1, 2) == [[]]
    img = (np.arange(16 * 8 * 3) % 256).astype(np.uint8).reshape(16, 8, 3)
    write_png(str(tmp_path / "a.png"), img)
    import struct
    import zlib
    raw = open(tmp_path / "a.png", "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and struct.unpack(">II", raw[16:24]) == (8, 16)
    i = raw.index(b"IDAT")
    n = struct.unpack(">I", raw[i - 4:i])[0]
    rows = np.frombuffer(zlib.decompress(raw[i + 4:i + 4 + n]), np.uint8).reshape(16, 1 + 8 * 3)
    assert np.array_equal(rows[:, 1:].reshape(16, 8, 3), img) and not rows[:, 0].any()


def test_webdataset_shards_roundtrip(tmp_path):
    """lmdb2wds.py:26 layout (`<key>.latent` pickle, `<key>.cls` ASCII) -> train_wds.py:58-64 decoding, shards split by rank."""
    rng = np.random.default_rng(2)
    moments = rng.standard_normal((12, 8, 4, 4)).astype(np.float32)
    labels = rng.integers(0, 10, 12)
    paths = []
    for s in range(3):
        p = str(tmp_path / f"latent-{s:04d}.tar")
        D.write_wds_shard(p, moments[4 * s:4 * s + 4], labels[4 * s:4 * s + 4], start=4 * s)
        paths.append(p)
    it = D.wds_samples(paths, rank=0, world=1, num_classes=10)
    for i in range(14):                                       # 12 samples, then the epoch wraps around
        z, y = next(it)
        assert np.array_equal(z, moments[i % 12]) and y.argmax() == labels[i % 12] and y.sum() == 1
    it1 = D.wds_samples(paths, rank=1, world=2, num_classes=10)          # data_list[rank::world] -> shard 1 only
    assert np.array_equal(next(it1)[0], moments[4])
    zb, yb = next(D.wds_batches(paths, 5, num_classes=10, pin=False))
    assert torch.equal(zb, torch.from_numpy(moments[:5])) and yb.argmax(1).tolist() == labels[:5].tolist()
    with pytest.raises(ValueError):
        next(D.wds_samples(paths[:1], rank=1, world=2))


# ---- property tests (hypothesis): host logic against its specification over random inputs -------------------------------
from hypothesis import given, settings, strategies as st_  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(n=st_.integers(1, 400), mb=st_.integers(1, 64), size=st_.integers(1, 8))
def test_rank_seed_batches_equals_tensor_split(n, mb, size):
    """sample.py:232-235 for arbitrary (seed count, max batch, world): same batches as torch.tensor_split(...)[rank::size],
    every seed exactly once, no batch above max_batch_size, every rank the same number of batches."""
    seeds = list(range(1000, 1000 + n))
    num_batches = ((n - 1) // (mb * size) + 1) * size
    ref = [b.tolist() for b in torch.as_tensor(seeds).tensor_split(num_batches)]
    got = [rank_seed_batches(seeds, mb, r, size) for r in range(size)]
    for r in range(size):
        assert got[r] == ref[r::size]
    assert sorted(s for g in got for b in g for s in b) == seeds
    assert all(len(b) <= mb for g in got for b in g) and len({len(g) for g in got}) == 1


@settings(max_examples=25, deadline=None)
@given(data=st_.dictionaries(st_.binary(min_size=1, max_size=40), st_.binary(min_size=0, max_size=6000), min_size=1,
                             max_size=120))
def test_mdb_roundtrip_random_tables(tmp_path_factory, data):
    """Any table of byte keys / values (inline and overflow-page values mixed) written by `write_mdb` is read back exactly
    by the page walker, and absent keys are reported absent."""
    d = tmp_path_factory.mktemp("mdb")
    D.write_mdb(str(d), data)
    rd = D.MdbReader(str(d))
    assert rd.entries == len(data)
    for k, v in data.items():
        assert bytes(rd.get(k)) == v
    for k in list(data)[:5]:
        probe = k + b"\x00zz"
        if probe not in data:
            assert rd.get(probe) is None
    rd.close()
