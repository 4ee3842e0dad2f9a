"""CPU: host-side mirror of the Faiss surface (argument handling, CPU-side collection, sharding
arithmetic). No compute."""
import numpy as np
import pytest

from keds_b200 import faiss_compat as faiss
from keds_b200.sharded import packed_layout, shard_bounds


def test_flat_index_collects_rows_but_cannot_search():
    ix = faiss.IndexFlatL2(8)
    assert ix.d == 8 and ix.ntotal == 0 and ix.is_trained and ix.metric_type == faiss.METRIC_L2
    ix.add(np.zeros((3, 8), np.float32))
    ix.add(np.ones((2, 8), np.float32))
    assert ix.ntotal == 5
    with pytest.raises(RuntimeError, match="no CPU search"):
        ix.search(np.zeros((1, 8), np.float32), 1)
    ix.reset()
    assert ix.ntotal == 0
    assert faiss.IndexFlatIP(4).metric_type == faiss.METRIC_INNER_PRODUCT


def test_add_argument_checks_follow_faiss():
    ix = faiss.IndexFlatIP(8)
    with pytest.raises(AssertionError):
        ix.add(np.zeros((3, 7), np.float32))
    with pytest.raises(TypeError):
        ix.add(np.zeros((3, 8), np.float64))
    with pytest.raises(ValueError):
        ix.add(np.zeros(8, np.float32))
    ix.add(np.asfortranarray(np.zeros((3, 8), np.float32)))  # coerced to C order
    assert ix._blocks[0].flags["C_CONTIGUOUS"]


def test_cloning_to_gpu_without_a_gpu_fails_loudly():
    if faiss.get_num_gpus() > 0:
        pytest.skip("a GPU is present")
    ix = faiss.IndexFlatL2(8)
    with pytest.raises(RuntimeError):
        faiss.index_cpu_to_gpu(faiss.StandardGpuResources(), 0, ix)
    with pytest.raises(RuntimeError):
        faiss.index_cpu_to_all_gpus(ix)


@pytest.mark.parametrize("n,world", [(10, 1), (10, 2), (10, 3), (7, 8), (0, 4), (8_000_000, 8)])
def test_shard_bounds_partition_the_rows(n, world):
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
        assert b == c and a <= b and c <= d
    per = -(-n // world)
    assert all(hi - lo <= per for lo, hi in spans)


def test_packed_layout_is_8_byte_aligned():
    for nq, k in [(1, 1), (128, 16), (3, 5), (4096, 64)]:
        total, off_i, d_bytes = packed_layout(nq, k)
        assert off_i % 8 == 0 and total % 8 == 0 and d_bytes == nq * k * 4
        assert total == off_i + nq * k * 8


def test_database_sampling_and_pair_loading(tmp_path):
    """database.py:14-36 + LoadDataBase (src/data.py:636-671): sample names, keep complete pairs."""
    import torch
    from keds_b200 import database as kdb

    names = [f"{i:05d}.pt" for i in range(20)]
    picked = kdb.sample_pairs(names, 8, seed=3)
    assert len(picked) == 8 and len(set(picked)) == 8 and set(picked) <= set(names)
    assert picked == kdb.sample_pairs(names, 8, seed=3)
    assert len(kdb.sample_pairs(names, 100, seed=1)) == 20
    img_dir, txt_dir = tmp_path / "image_feature_database", tmp_path / "text_feature_database"
    img_dir.mkdir()
    txt_dir.mkdir()
    for i, n in enumerate(picked):
        torch.save(torch.full((1, 6), float(i)), img_dir / n)          # [1, d] like a CLIP output
        if i != 2:                                                      # one pair lacks its text half
            torch.save(torch.full((6,), -float(i)), txt_dir / n)
    img, txt, kept = kdb.load_pair_folders(str(tmp_path), picked)
    assert img.shape == (7, 6) and txt.shape == (7, 6) and picked[2] not in kept
    assert img.dtype == torch.float32 and float(img[0, 0]) == 0.0 and float(txt[1, 0]) == -1.0
