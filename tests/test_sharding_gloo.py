"""CPU test of the N>1 host path: world_size-2 gloo process group, id scatter + PCM gather with a
stand-in synthesiser (the real one needs a GPU)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_bounds_cover_and_balance():
    from mimic3_b200.shard import shard_bounds, shard_by_cost
    for n in (1, 7, 32, 256, 257):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    costs = [80, 5, 1800, 40, 41, 39, 900, 7]
    parts = shard_by_cost(costs, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(costs)))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= 1800 + 41  # longest-first greedy: no rank far above the biggest item


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from mimic3_b200.shard import gather_pcm, scatter_ids, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, T = 7, 11
    rng = np.random.default_rng(0)
    ids = rng.integers(4, 50, size=(B, T)).astype(np.int64)
    lengths = rng.integers(1, T + 1, size=B).astype(np.int64)
    sid = (np.arange(B) % 3).astype(np.int64)
    got = scatter_ids(ids if rank == 0 else None, lengths if rank == 0 else None, sid if rank == 0 else None, "cpu")
    lo, hi = shard_bounds(B, world, rank)
    assert np.array_equal(got[0].numpy(), ids[lo:hi]) and np.array_equal(got[1].numpy(), lengths[lo:hi])
    assert np.array_equal(got[2].numpy(), sid[lo:hi])
    # stand-in synthesiser: utterance b -> lengths[b]*4 samples of value ids[b,0]
    pcm, offs = [], [0]
    for i in range(hi - lo):
        n = int(got[1][i]) * 4
        pcm.append(np.full(n, int(got[0][i, 0]), dtype=np.int16))
        offs.append(offs[-1] + n)
    out = gather_pcm(torch.from_numpy(np.concatenate(pcm)), offs, "cpu")
    if rank == 0:
        bufs, offsets = out
        flat = []
        for r in range(world):
            for i in range(len(offsets[r]) - 1):
                flat.append(bufs[r][offsets[r][i]:offsets[r][i + 1]].numpy())
        assert len(flat) == B
        for b in range(B):
            assert flat[b].shape[0] == lengths[b] * 4 and (flat[b] == ids[b, 0]).all()
        Path(tmp, "ok").write_text("ok")
    else:
        assert out is None
    dist.destroy_process_group()


def test_scatter_gather_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
