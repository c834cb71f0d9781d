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


def _worker(rank, world, port, tmp, transport="nccl"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from mimic3_b200.shard import HostPcmCollector, IdScatter, PcmCollector, make_groups, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg, mg = make_groups("cpu")
    B, T = 7, 11
    scatter = IdScatter(max_rows=9, max_t=16, device="cpu", payload_group=pg, meta_group=mg)
    per = (9 + world - 1) // world
    if transport == "host":   # every rank writes its own slice of one shared-memory segment; no payload collective
        collector = HostPcmCollector(capacity_samples=per * 16 * 4, max_rows_per_rank=per, device="cpu", meta_group=mg, depth=2)
    else:
        collector = PcmCollector(capacity_samples=per * 16 * 4, max_rows_per_rank=per, device="cpu", payload_group=pg,
                                 meta_group=mg, depth=2)
    tickets, want = [], []
    for step in range(5):      # more steps than slots: buffers are reused, nothing is allocated per step
        rng = np.random.default_rng(step)
        Bs = B if step != 3 else 2          # a batch smaller than the world still works (rank 1 gets 1 row or none)
        Bs = 1 if step == 4 else Bs
        ids = rng.integers(4, 50, size=(Bs, T)).astype(np.int64)
        lengths = rng.integers(1, T + 1, size=Bs).astype(np.int64)
        sid = (np.arange(Bs) % 3).astype(np.int64) if step % 2 == 0 else None
        got = scatter(ids if rank == 0 else None, lengths if rank == 0 else None, sid if rank == 0 else None)
        lo, hi = shard_bounds(Bs, world, rank)
        assert got[0].shape == (hi - lo, 16)
        assert np.array_equal(got[0][:, :T].numpy(), ids[lo:hi]) and not got[0][:, T:].any()
        assert np.array_equal(got[1], lengths[lo:hi])
        assert (got[2] is None) if sid is None else np.array_equal(got[2], sid[lo:hi])
        # stand-in synthesiser: utterance b -> lengths[b] frames of 4 samples, every sample = ids[b,0] + step
        buf = collector.send_buffer()
        n, frames = 0, []
        for i in range(hi - lo):
            k = int(got[1][i]) * 4
            buf[n:n + k] = int(got[0][i, 0]) + step
            n += k
            frames.append(int(got[1][i]))
        tickets.append(collector.submit(n, frames))
        want.append((ids, lengths, step))
        if len(tickets) == 2:                  # collect one step behind, like the bench's pipelined loop
            _check(collector.collect(tickets.pop(0)), want.pop(0), rank, world)
    while tickets:
        _check(collector.collect(tickets.pop(0)), want.pop(0), rank, world)
    collector.drain()
    if hasattr(collector, "close"):
        collector.close()
    if rank == 0:
        Path(tmp, "ok").write_text("ok")
    dist.destroy_process_group()


def _check(out, want, rank, world):
    ids, lengths, step = want
    if rank != 0:
        assert out is None
        return
    pcm, frames = out
    if isinstance(pcm, list):     # HostPcmCollector: one view per rank, rank order == row order
        assert len(pcm) == world and all(a.dtype == np.int16 for a in pcm)
        pcm = np.concatenate(pcm)
    assert len(frames) == world
    flat = np.concatenate([np.asarray(f) for f in frames])
    assert np.array_equal(flat, lengths)                      # rank order == row order
    assert pcm.dtype == np.int16 and pcm.shape[0] == int(lengths.sum()) * 4
    off = 0
    for b in range(len(lengths)):
        k = int(lengths[b]) * 4
        assert (pcm[off:off + k] == ids[b, 0] + step).all()
        off += k


def test_scatter_gather_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_scatter_gather_world3_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    assert (tmp_path / "ok").exists()


def test_host_segment_collector_world3_gloo(tmp_path):
    """HostPcmCollector: same steps, same checks, PCM through one POSIX shared-memory segment instead of send/recv."""
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(3, port, str(tmp_path), "host"), nprocs=3, join=True)
    assert (tmp_path / "ok").exists()


def _worker_two_in_flight(rank, world, port, tmp):
    """The bench's N>1 loop: two 'engine calls' per rank hold send buffers at once (the second one taken with ahead=1),
    three collector slots, two id-scatter buffer sets, collect two submits behind."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from mimic3_b200.shard import HostPcmCollector, IdScatter, make_groups, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg, mg = make_groups("cpu")
    T, per = 11, (9 + world - 1) // world
    scats = [IdScatter(max_rows=9, max_t=16, device="cpu", payload_group=pg, meta_group=mg) for _ in range(2)]
    coll = HostPcmCollector(capacity_samples=per * 16 * 4, max_rows_per_rank=per, device="cpu", meta_group=mg, depth=3)

    def batch(step):
        rng = np.random.default_rng(100 + step)
        Bs = [7, 5, 2, 9, 1, 7, 3][step]
        return (rng.integers(4, 50, size=(Bs, T)).astype(np.int64), rng.integers(1, T + 1, size=Bs).astype(np.int64))

    def scatter(step):
        ids, lengths = batch(step)
        return scats[step & 1](ids if rank == 0 else None, lengths if rank == 0 else None, None)

    def synth(buf, shard, step):      # the stand-in engine call: fills ITS buffer, possibly long after a later call took the next one
        n, frames = 0, []
        for i in range(shard[0].shape[0]):
            k = int(shard[1][i]) * 4
            buf[n:n + k] = int(shard[0][i, 0]) + step
            n += k
            frames.append(int(shard[1][i]))
        return n, frames

    steps, inflight, tickets, want = 7, [], [], []
    nxt = scatter(0)
    for step in range(steps):
        shard = nxt
        buf = coll.send_buffer(len(inflight))          # 0 or 1 calls already hold a buffer
        inflight.append((buf, (shard[0].clone(), shard[1].copy()), step))
        if len(inflight) > 1:                          # the older call "finishes" only now, after the newer one took its buffer
            b, sh, st = inflight.pop(0)
            tickets.append(coll.submit(*synth(b, sh, st)))
            want.append(batch(st) + (st,))
        while len(tickets) >= 2:
            _check(coll.collect(tickets.pop(0)), want.pop(0), rank, world)
        if step + 1 < steps:
            nxt = scatter(step + 1)                    # into the other buffer set, while `shard`'s call is still "running"
    while inflight:
        b, sh, st = inflight.pop(0)
        tickets.append(coll.submit(*synth(b, sh, st)))
        want.append(batch(st) + (st,))
    while tickets:
        _check(coll.collect(tickets.pop(0)), want.pop(0), rank, world)
    coll.drain()
    coll.close()
    if rank == 0:
        Path(tmp, "ok").write_text("ok")
    dist.destroy_process_group()


def test_two_calls_in_flight_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_worker_two_in_flight, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
