"""Batch sharding for multi-GPU synthesis (one process per GPU).

Utterances are independent (the reference already issues one ``run`` per sentence,
``mimic3_tts/tts.py:474-513``), so the batch is partitioned by rows and there is NO collective
on the compute path.  The only exchanges are the ones BASELINE.json names: scatter the padded
phoneme-id tensor from rank 0, gather int16 PCM back to rank 0 (NCCL on GPUs; the same code
runs over gloo on CPU tensors for the tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of ``rank`` (first ``n_rows % world`` ranks get one extra)."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cost(costs: Sequence[int], world: int) -> List[List[int]]:
    """Ragged batches: longest-first greedy assignment of rows to ranks (cost ~ phoneme count)."""
    order = sorted(range(len(costs)), key=lambda i: -int(costs[i]))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(costs[i])
    return [sorted(rows) for rows in out]


def scatter_ids(ids, lengths, sid, device, group=None):
    """Rank 0 holds (ids int64 [B,T], lengths [B], sid [B] or None); every rank returns its shard
    as tensors on ``device``.  Uses torch.distributed.scatter (NCCL on cuda, gloo on cpu)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == 0:
        meta = torch.tensor([ids.shape[0], ids.shape[1], 0 if sid is None else 1], dtype=torch.int64, device=device)
    dist.broadcast(meta, src=0, group=group)
    B, T, has_sid = (int(v) for v in meta.tolist())
    lo, hi = shard_bounds(B, world, rank)
    per = max(shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world))
    cols = T + 2  # ids | length | sid packed in one tensor -> one scatter
    mine = torch.zeros((per, cols), dtype=torch.int64, device=device)
    chunks = None
    if rank == 0:
        full = torch.zeros((B, cols), dtype=torch.int64, device=device)
        full[:, :T] = torch.as_tensor(ids, device=device)
        full[:, T] = torch.as_tensor(lengths, device=device)
        if sid is not None:
            full[:, T + 1] = torch.as_tensor(sid, device=device)
        chunks = []
        for r in range(world):
            a, b = shard_bounds(B, world, r)
            c = torch.zeros((per, cols), dtype=torch.int64, device=device)
            c[: b - a] = full[a:b]
            chunks.append(c)
    dist.scatter(mine, chunks, src=0, group=group)
    n = hi - lo
    return mine[:n, :T].contiguous(), mine[:n, T].contiguous(), (mine[:n, T + 1].contiguous() if has_sid else None)


def gather_pcm(pcm, sample_offsets, device, group=None):
    """Every rank passes its packed int16 PCM (tensor on ``device``) and per-utterance offsets;
    rank 0 gets (list of per-rank PCM tensors, list of per-rank offset arrays); others get None."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([pcm.numel(), len(sample_offsets)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    max_s = max(int(s[0]) for s in sizes)
    max_o = max(int(s[1]) for s in sizes)
    buf = torch.zeros(max_s, dtype=torch.int16, device=device)
    buf[: pcm.numel()] = pcm.reshape(-1)
    off = torch.zeros(max_o, dtype=torch.int64, device=device)
    off[: len(sample_offsets)] = torch.as_tensor(np.asarray(sample_offsets), device=device)
    # payload travels as raw bytes (gloo has no int16 collectives; NCCL does not care)
    bufs = [torch.zeros(max_s * 2, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    offs = [torch.zeros(max_o, dtype=torch.int64, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf.view(torch.uint8), bufs, dst=0, group=group)
    if rank == 0:
        bufs = [b.view(torch.int16) for b in bufs]
    dist.gather(off, offs, dst=0, group=group)
    if rank != 0:
        return None
    return ([b[: int(s[0])] for b, s in zip(bufs, sizes)], [o[: int(s[1])].cpu().numpy() for o, s in zip(offs, sizes)])
