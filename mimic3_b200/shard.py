"""Batch sharding for multi-GPU synthesis (one process per GPU).

Utterances are independent (the reference already issues one ``run`` per sentence,
``mimic3_tts/tts.py:474-513``), so the batch is partitioned by rows and there is NO collective
on the compute path.  The only exchanges are the ones BASELINE.json names: scatter the padded
phoneme-id tensor from rank 0, gather int16 PCM back to rank 0.

Round-2 shape of the two exchanges (round 1 measured 153 ms per step at 8 GPUs against 16.6 ms of
compute, all of it in per-step allocations, ``.item()`` syncs, a padded ``dist.gather`` and a pageable
``.cpu()`` on rank 0):

* :class:`IdScatter` -- buffers allocated once; rank 0 stages the ids in pinned host memory, one async
  H2D copy, one NCCL ``scatter`` of equal-sized row blocks; the two small per-row vectors the engine
  wants on the HOST (lengths, speaker ids) travel host-to-host over a gloo side group, so no rank has
  to read anything back from its GPU to start computing.
* :class:`PcmCollector` -- exact-size grouped NCCL ``send``/``recv`` into a pre-allocated device
  staging buffer on rank 0 (sizes are known on the host from the engine call and exchanged over the
  gloo side group: no device sync), then ONE D2H copy into pinned host memory on a copy stream.
  Slots are double-buffered: step k's gather + D2H run while step k+1 computes (372 MB through one
  PCIe Gen5 link is ~7 ms, under the 16.6 ms of compute it hides behind).

* :class:`HostPcmCollector` -- the gather WITHOUT a payload collective: every rank copies its own PCM
  device -> host over its OWN PCIe link into its slice of one POSIX shared-memory segment (pinned with
  ``cudaHostRegister``) that rank 0 reads.  With the NCCL gather all of the step's PCM (46 MB per GPU, 373 MB at
  8 GPUs) funnels through rank 0's GPU and its single PCIe link while that GPU is also computing its own shard:
  measured 0.82 of linear end to end at 4 GPUs (r02p).  Per-rank copies are 1 ms each, in parallel, hidden under
  the next step.  Same interface as :class:`PcmCollector`; ``collect`` returns one int16 view per rank.

The same classes run over a single gloo group on CPU tensors (``tests/test_sharding_gloo.py``).
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


def _is_cuda(device) -> bool:
    import torch
    return torch.device(device).type == "cuda"


class IdScatter:
    """Scatter of one padded id batch per call from rank 0; all buffers are allocated at construction.

    ``max_rows`` / ``max_t`` bound the GLOBAL batch (rows, ids per row).  ``payload_group`` carries the id
    tensor (NCCL on GPUs), ``meta_group`` the host-side vectors (gloo); with one gloo group and
    ``device="cpu"`` both are the same group.
    """

    def __init__(self, max_rows: int, max_t: int, device, payload_group=None, meta_group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch = dist, torch
        self.pg, self.mg = payload_group, meta_group
        self.world, self.rank = dist.get_world_size(payload_group), dist.get_rank(payload_group)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.per = (max_rows + self.world - 1) // self.world
        self.max_t = max_t
        self.mine = torch.zeros((self.per, max_t), dtype=torch.int64, device=self.device)
        self.meta_mine = torch.zeros(3 + 2 * self.per, dtype=torch.int64)       # B, T, has_sid | lengths | sid
        if self.rank == 0:
            self.h_full = torch.zeros((self.world * self.per, max_t), dtype=torch.int64, pin_memory=self.cuda)
            self.d_full = torch.zeros((self.world * self.per, max_t), dtype=torch.int64, device=self.device)
            self.meta_full = torch.zeros((self.world, 3 + 2 * self.per), dtype=torch.int64)

    def __call__(self, ids: Optional[np.ndarray], lengths: Optional[np.ndarray], sid: Optional[np.ndarray]):
        """Rank 0 passes (ids int64 [B,T], lengths [B], sid [B] or None), the others ``None``.  Returns this rank's
        shard: (ids tensor on ``device`` [n, max_t] -- a view of the persistent receive buffer (row stride ``max_t``,
        columns past the batch's T are zero), valid until the next call --, lengths numpy int64 [n], sid numpy
        int64 [n] or None).  On CUDA the ids are complete on the
        device when this returns (the current stream has been synchronised with the collective)."""
        torch, dist = self.torch, self.dist
        chunks = meta_chunks = None
        if self.rank == 0:
            B, T = ids.shape
            if B > self.world * self.per or T > self.max_t:
                raise ValueError(f"batch {B} x {T} exceeds the scatter buffers ({self.world * self.per} x {self.max_t})")
            self.meta_full.zero_()
            self.h_full[:, :T].zero_()
            for r in range(self.world):
                a, b = shard_bounds(B, self.world, r)
                n = b - a
                self.h_full[r * self.per: r * self.per + n, :T] = torch.from_numpy(np.ascontiguousarray(ids[a:b]))
                m = self.meta_full[r]
                m[0], m[1], m[2] = B, T, 0 if sid is None else 1
                m[3: 3 + n] = torch.from_numpy(np.ascontiguousarray(lengths[a:b], dtype=np.int64))
                if sid is not None:
                    m[3 + self.per: 3 + self.per + n] = torch.from_numpy(np.ascontiguousarray(sid[a:b], dtype=np.int64))
            self.d_full.copy_(self.h_full, non_blocking=True)                     # the step's H2D: pinned -> device
            chunks = list(self.d_full.view(self.world, self.per, self.max_t).unbind(0))
            meta_chunks = list(self.meta_full.unbind(0))
        dist.scatter(self.meta_mine, meta_chunks, src=0, group=self.mg)          # host -> host
        dist.scatter(self.mine, chunks, src=0, group=self.pg)                    # device -> device (NCCL)
        B, T, has_sid = (int(v) for v in self.meta_mine[:3])
        lo, hi = shard_bounds(B, self.world, self.rank)
        n = hi - lo
        if self.cuda:
            torch.cuda.current_stream(self.device).synchronize()                 # engine runs on its own stream
        lens = self.meta_mine[3: 3 + n].numpy().copy()
        sids = self.meta_mine[3 + self.per: 3 + self.per + n].numpy().copy() if has_sid else None
        return self.mine[:n], lens, sids


class PcmTicket:
    __slots__ = ("slot", "total", "counts", "frames", "event", "works")


class PcmCollector:
    """Gather of packed int16 PCM to pinned host memory on rank 0, double-buffered (module docstring).

    Every rank: ``buf = collector.send_buffer(n)`` (persistent device buffer the engine writes its PCM into),
    ``t = collector.submit(n_samples, frames_per_utterance)``; rank 0 later: ``pcm, per_utt = collector.collect(t)``.
    """

    def __init__(self, capacity_samples: int, max_rows_per_rank: int, device, payload_group=None, meta_group=None,
                 depth: int = 2):
        import torch
        import torch.distributed as dist
        self.dist, self.torch = dist, torch
        self.pg, self.mg = payload_group, meta_group
        self.world, self.rank = dist.get_world_size(payload_group), dist.get_rank(payload_group)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.depth, self.k = depth, 0
        self.per = max_rows_per_rank
        self.cap = int(capacity_samples)
        self.src = [torch.zeros(self.cap, dtype=torch.int16, device=self.device) for _ in range(depth)]
        self.pending: List[Optional[PcmTicket]] = [None] * depth
        self.meta_mine = torch.zeros(2 + self.per, dtype=torch.int64)           # samples, utterances | frames
        if self.rank == 0:
            self.meta_all = [torch.zeros(2 + self.per, dtype=torch.int64) for _ in range(self.world)]
            total_cap = self.cap * self.world
            # rank 0's own PCM is written straight into its slice of the staging buffer
            self.stage = [torch.zeros(total_cap, dtype=torch.int16, device=self.device) for _ in range(depth)]
            self.host = [torch.zeros(total_cap, dtype=torch.int16, pin_memory=self.cuda) for _ in range(depth)]
            self.copy_stream = torch.cuda.Stream(self.device) if self.cuda else None
            self.events = [torch.cuda.Event() for _ in range(depth)] if self.cuda else None

    # -- buffers ---------------------------------------------------------------------------------
    def _slot(self) -> int:
        return self.k % self.depth

    def send_buffer(self):
        """Device int16 buffer for this step's PCM (rank 0: the head of the staging buffer).  Waits -- on the
        host, microseconds in steady state -- until the previous use of the slot has left the GPU."""
        s = self._slot()
        t = self.pending[s]
        if t is not None:
            self._finish(t)
            self.pending[s] = None
        return self.stage[s][: self.cap] if self.rank == 0 else self.src[s]

    def _finish(self, t: PcmTicket):
        if t.event is not None:
            t.event.synchronize()
        for w in t.works or ():
            w.wait()
        t.works = None

    # -- one step ----------------------------------------------------------------------------------
    def submit(self, n_samples: int, frames) -> PcmTicket:
        torch, dist = self.torch, self.dist
        s = self._slot()
        self.k += 1
        if n_samples > self.cap:
            raise ValueError(f"{n_samples} samples exceed the collector capacity {self.cap}")
        frames = np.asarray(frames, dtype=np.int64)
        self.meta_mine.zero_()
        self.meta_mine[0], self.meta_mine[1] = int(n_samples), len(frames)
        self.meta_mine[2: 2 + len(frames)] = torch.from_numpy(frames)
        dist.gather(self.meta_mine, self.meta_all if self.rank == 0 else None, dst=0, group=self.mg)   # host side
        t = PcmTicket()
        t.slot, t.event, t.works = s, None, None
        if self.rank != 0:
            if self.world > 1 and n_samples:
                op = dist.P2POp(dist.isend, self.src[s][:n_samples].view(torch.uint8), 0, self.pg)
                t.works = dist.batch_isend_irecv([op])
            t.total, t.counts, t.frames = n_samples, None, None
            self.pending[s] = t
            return t
        counts = [int(m[0]) for m in self.meta_all]
        t.counts = counts
        t.frames = [m[2: 2 + int(m[1])].numpy().copy() for m in self.meta_all]
        t.total = sum(counts)
        stage = self.stage[s]
        ops, off = [], counts[0]
        for r in range(1, self.world):
            if counts[r]:
                # payload travels as raw bytes (gloo has no int16 collectives; NCCL does not care)
                ops.append(dist.P2POp(dist.irecv, stage[off: off + counts[r]].view(torch.uint8), r, self.pg))
            off += counts[r]
        works = dist.batch_isend_irecv(ops) if ops else []
        if self.cuda:
            with torch.cuda.stream(self.copy_stream):
                for w in works:
                    w.wait()                                  # the copy stream waits for the receives (no host wait)
                # pack: ranks wrote at their exact offsets already, rank 0's own samples sit at the head
                self.host[s][: t.total].copy_(stage[: t.total], non_blocking=True)
                t.event = self.events[s]
                t.event.record(self.copy_stream)
        else:
            for w in works:
                w.wait()
            self.host[s][: t.total].copy_(stage[: t.total])
        self.pending[s] = t
        return t

    def collect(self, t: PcmTicket):
        """Rank 0: (int16 numpy view of the pinned host buffer [total], list of per-rank frame-count arrays).
        The view is valid until the slot is reused (``depth`` submits later)."""
        self._finish(t)
        if self.rank != 0:
            return None
        return self.host[t.slot][: t.total].numpy(), t.frames

    def drain(self):
        for i, t in enumerate(self.pending):
            if t is not None:
                self._finish(t)
                self.pending[i] = None
        if self.cuda:
            self.torch.cuda.synchronize(self.device)


class HostPcmCollector:
    """PCM of every rank in ONE shared pinned host segment that rank 0 reads; no payload collective (module docstring).

    Every rank: ``buf = c.send_buffer()`` (persistent device buffer the engine writes its PCM into),
    ``t = c.submit(n_samples, frames_per_utterance)`` (async D2H of this rank's samples into its slice + a host-side
    gather of the counts), and -- every rank, same order -- ``c.collect(t)``: waits for the own copy, then a host
    barrier; rank 0 gets ``([int16 view per rank], [frame counts per rank])``, valid until the slot is reused.
    """

    def __init__(self, capacity_samples: int, max_rows_per_rank: int, device, meta_group=None, depth: int = 2):
        import torch
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self.dist, self.torch = dist, torch
        self.mg = meta_group
        self.world, self.rank = dist.get_world_size(meta_group), dist.get_rank(meta_group)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.depth, self.k = depth, 0
        self.per = max_rows_per_rank
        self.cap = int(capacity_samples)
        nbytes = 2 * self.cap * self.world * depth
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=0, group=meta_group)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
        # layout [rank][slot][cap]: a rank's slices are one contiguous region (one cudaHostRegister)
        self.all = torch.from_numpy(np.ndarray((self.world, depth, self.cap), dtype=np.int16, buffer=self.shm.buf))
        self.mine = self.all[self.rank]
        self._registered = None
        if self.cuda:
            ptr, size = self.mine.data_ptr(), self.mine.numel() * 2
            rc = torch.cuda.cudart().cudaHostRegister(ptr, size, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister of the shared PCM segment failed ({rc})")
            self._registered = ptr
            self.copy_stream = torch.cuda.Stream(self.device)
            self.events = [torch.cuda.Event() for _ in range(depth)]
        self.src = [torch.zeros(self.cap, dtype=torch.int16, device=self.device) for _ in range(depth)]
        self.pending: List[Optional[PcmTicket]] = [None] * depth
        self.meta_mine = torch.zeros(2 + self.per, dtype=torch.int64)
        self.meta_all = [torch.zeros(2 + self.per, dtype=torch.int64) for _ in range(self.world)] if self.rank == 0 else None
        dist.barrier(group=meta_group)

    def _slot(self) -> int:
        return self.k % self.depth

    def send_buffer(self, ahead: int = 0):
        """Device buffer of the next call's PCM.  ``ahead`` = engine calls that already hold a buffer but have not been
        ``submit``-ted yet (two calls in flight: the second one asks with ``ahead=1``); needs ``depth >= ahead + 2`` for
        the slot's previous copy to be long finished."""
        s = (self.k + ahead) % self.depth
        t = self.pending[s]
        if t is not None:          # the slot's previous D2H must have left the device buffer
            if t.event is not None:
                t.event.synchronize()
            self.pending[s] = None
        return self.src[s]

    def submit(self, n_samples: int, frames) -> PcmTicket:
        torch, dist = self.torch, self.dist
        s = self._slot()
        self.k += 1
        if n_samples > self.cap:
            raise ValueError(f"{n_samples} samples exceed the collector capacity {self.cap}")
        frames = np.asarray(frames, dtype=np.int64)
        t = PcmTicket()
        t.slot, t.event, t.works, t.total = s, None, None, int(n_samples)
        if n_samples:
            if self.cuda:
                with torch.cuda.stream(self.copy_stream):       # the engine call has completed: src[s] is final
                    self.mine[s][:n_samples].copy_(self.src[s][:n_samples], non_blocking=True)
                    t.event = self.events[s]
                    t.event.record(self.copy_stream)
            else:
                self.mine[s][:n_samples].copy_(self.src[s][:n_samples])
        self.meta_mine.zero_()
        self.meta_mine[0], self.meta_mine[1] = int(n_samples), len(frames)
        self.meta_mine[2: 2 + len(frames)] = torch.from_numpy(frames)
        dist.gather(self.meta_mine, self.meta_all, dst=0, group=self.mg)   # host side: counts + frames per utterance
        if self.rank == 0:
            t.counts = [int(m[0]) for m in self.meta_all]
            t.frames = [m[2: 2 + int(m[1])].numpy().copy() for m in self.meta_all]
            t.total = sum(t.counts)
        else:
            t.counts = t.frames = None
        self.pending[s] = t
        return t

    def collect(self, t: PcmTicket):
        """EVERY rank calls this, in submit order.  Rank 0: ([int16 numpy view of rank r's samples], frames)."""
        if t.event is not None:
            t.event.synchronize()
        self.dist.barrier(group=self.mg)                          # every rank's copy has landed in the segment
        if self.rank != 0:
            return None
        return [self.all[r][t.slot][: t.counts[r]].numpy() for r in range(self.world)], t.frames

    def drain(self):
        for i, t in enumerate(self.pending):
            if t is not None and t.event is not None:
                t.event.synchronize()
            self.pending[i] = None
        if self.cuda:
            self.torch.cuda.synchronize(self.device)

    def close(self):
        self.drain()
        if self._registered is not None:
            self.torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None
        self.all = self.mine = None
        try:
            self.dist.barrier(group=self.mg)
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:  # pragma: no cover
            pass


def make_groups(device):
    """(payload_group, meta_group): on CUDA a dedicated NCCL communicator for the PCM gather (so a 46 MB receive
    never sits in front of the next step's id scatter on one NCCL stream) and a gloo group for host-side metadata;
    on CPU both are the default (gloo) group."""
    import torch.distributed as dist
    if _is_cuda(device):
        return dist.new_group(backend="nccl"), dist.new_group(backend="gloo")
    return None, None
