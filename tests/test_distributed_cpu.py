"""CPU gate for the N > 1 path: world_size-2 (and 3, ragged) gloo processes shard a global utterance
range, 'score' their slice with a deterministic stand-in, all-gather, and must reproduce the global
vector in the reference's utterance order and the same EER as the un-sharded oracle."""
import os
import socket

import numpy as np
import pytest

from dfs_b200 import distributed as dd

torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_scores(lo, hi):
    idx = np.arange(lo, hi, dtype=np.int64)
    return (np.sin(idx * 0.37) * 0.25 + 0.5).astype(np.float32)


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = dd.shard_range(n, rank, world)
        local = torch.from_numpy(_fake_scores(lo, hi))
        g = dd.gather_scores(local, n_total=n)
        g2 = dd.gather_scores(local)          # sizes exchanged instead of derived
        assert torch.equal(g, g2)
        np.save(os.path.join(out_dir, f"gathered_{rank}.npy"), g.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1000), (2, 1001), (3, 10)])
def test_shard_gather_roundtrip(tmp_path, world, n):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    ref = _fake_scores(0, n)
    from oracle import eer as oeer
    labels = (np.arange(n) % 3 == 0).astype(np.int64)
    for r in range(world):
        g = np.load(tmp_path / f"gathered_{r}.npy")
        assert np.array_equal(g, ref)                                   # rank order == utterance order
        assert oeer.calculate_eer(g, labels) == oeer.calculate_eer(ref, labels)


def test_shard_ranges_cover_without_overlap():
    for n in (0, 1, 7, 208, 1_000_000):
        for world in (1, 2, 3, 8):
            spans = [dd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert sum(dd.shard_sizes(n, world)) == n
    with pytest.raises(ValueError):
        dd.shard_range(10, 2, 2)
