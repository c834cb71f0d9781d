#!/usr/bin/env python
"""Benchmark of the ids->waveform hot path (contract: see the task description / DESIGN.md §5).

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine (libm3b200)
    python bench.py --impl reference --gpus N --steps K ...    # the CPU path (oracle port), rank 0 only

Workload = BASELINE.json configs[2] (the one the metric is quoted on): an en_US/vctk_low-shaped
voice (109 speakers, synthetic random weights -- no real voice is reachable offline), batch 256,
80 ids per utterance, sid[b] = b mod 109, PCG64(1234).  Utterances are independent, so every rank
runs this workload on its own rows with no collective on the compute path: by default per-GPU
work is fixed (weak scaling, global batch 256 x N, as the bench contract prescribes for a path
that partitions); ``--scaling strong`` shards ONE batch of 256 over the N ranks instead
(BASELINE configs[2] literally: 32 utterances per GPU at N = 8).  A "step" = one pass of the hot
path over the whole (global) batch.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GLOBAL_BATCH = 256
IDS_PER_UTT = 80
N_SPEAKERS = 109
NUM_SYMBOLS = 50
VOICE_SEED = 22
METRIC = "audio samples/s (en_US/vctk_low batch=256)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_voice(root: Path) -> Path:
    """Synthetic vctk_low-shaped voice; written by local rank 0, shared through the filesystem."""
    from mimic3_b200 import synth_voice as sv
    d = root / "en_US" / "vctk_low"
    marker = d / ".complete"
    if not marker.exists():
        sv.write_voice(d, sv.low_config(n_speakers=N_SPEAKERS, num_symbols=NUM_SYMBOLS), seed=VOICE_SEED)
        marker.write_text("ok")
    return d


def make_inputs(batch: int = GLOBAL_BATCH):
    rng = np.random.Generator(np.random.PCG64(1234))
    ids = rng.integers(4, NUM_SYMBOLS, size=(batch, IDS_PER_UTT)).astype(np.int64)  # pad/bos/eos/blank excluded
    lengths = np.full(batch, IDS_PER_UTT, dtype=np.int64)
    sid = (np.arange(batch) % N_SPEAKERS).astype(np.int64)
    return ids, lengths, sid


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(gpu_index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # nvidia-smi missing
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_flops_per_frame(cfg) -> dict:
    """SURVEY.md §8(d) general formula, per mel frame, from the voice's config.  Also returns the
    per-stage split so the fused last-stage kernel can be charged with exactly its own work."""
    I = cfg.inter_channels
    C0 = cfg.upsample_initial_channel
    out = {"conv_pre": 2 * I * C0 * 7, "ups": 0, "mrf": 0, "ups_stage": [], "mrf_stage": []}
    L = 1
    c = C0
    per = 2 if cfg.resblock == "1" else 1
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, cout = c, c // 2
        up = 2 * cin * cout * k * L          # each input row touches k taps
        L *= u
        mrf = sum(2 * cout * cout * rk * len(dil) * per * L
                  for rk, dil in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes))
        out["ups"] += up
        out["mrf"] += mrf
        out["ups_stage"].append(up)
        out["mrf_stage"].append(mrf)
        c = cout
    out["post"] = 2 * c * 7 * L
    Hf = cfg.hidden_channels
    out["flow"] = 4 * 2 * ((I // 2) * Hf + 4 * (Hf * 2 * Hf * 5) + 3 * Hf * 2 * Hf + Hf * Hf + Hf * (I // 2))
    return out


def run_reference(args, rank: int, world: int):
    """CPU arm: the oracle port (kind "port": onnxruntime + generator.onnx are unobtainable offline)
    on rank 0's host cores, B=1 loop like the reference (voice.py:180-181), bounded sample per step."""
    if rank != 0:
        return
    import torch
    from oracle.vits_oracle import VitsOracle, audio_float_to_int16
    with tempfile.TemporaryDirectory() as d:
        vd = make_voice(Path(d))
        orc = VitsOracle(str(vd))
        ids, lengths, sid = make_inputs()
        sample = max(1, args.ref_utts)
        scales = orc.defaults
        cores, _tried = pick_cpu_threads(orc, ids, sid, scales)

        def step(k):
            n = 0
            for j in range(sample):
                b = (k * sample + j) % GLOBAL_BATCH
                n += audio_float_to_int16(orc.infer(ids[b], scales, sid=int(sid[b]), seed=1234, row=b)).size
            return n
        for k in range(args.warmup):
            step(k)
        t0 = time.perf_counter()
        total = sum(step(args.warmup + k) for k in range(args.steps))
        dt = time.perf_counter() - t0
    value = total / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: vctk_low-shaped synthetic voice, 109 speakers, batch=256 x 80 ids",
                   "sample": f"{sample} utterances of the batch per step, batch-1 loop"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} utterances/step x {args.steps} steps, torch fp32 CPU, {cores} threads"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-utts", type=int, default=32, help="utterances per step for the CPU arm")
    ap.add_argument("--cpu-baseline-utts", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--deterministic", action="store_true", help="noise scales 0 (parity settings)")
    ap.add_argument("--batch", type=int, default=GLOBAL_BATCH, help="batch override (profiling runs only)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch utterances PER GPU (default); strong: --batch utterances sharded over all GPUs")
    ap.add_argument("--profile-only", action="store_true", help="skip stage/e2e/cpu passes (ncu runs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world:
        log(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE {world}; using WORLD_SIZE")
    from mimic3_b200.engine import B200Session, STAGES
    from mimic3_b200 import synth_voice as sv

    # -- voice: written once (rank 0), read by every rank ---------------------------------
    shared = Path(os.environ.get("M3B200_BENCH_DIR", tempfile.gettempdir())) / f"m3b200_bench_voice_{os.environ.get('MASTER_PORT', 'single')}"
    if local_rank == 0:
        make_voice(shared)
    if distributed:
        dist.barrier()
    vd = make_voice(shared)
    sess = B200Session(str(vd), device=local_rank)
    cfg = sv.low_config(n_speakers=N_SPEAKERS, num_symbols=NUM_SYMBOLS)
    scales = (0.0, 1.0, 0.0) if args.deterministic else (sess.info.noise_scale, sess.info.length_scale, sess.info.noise_w)

    # -- inputs: rank r owns rows [r*B/N, (r+1)*B/N) ------------------------------------------
    GB = args.batch * world if args.scaling == "weak" else args.batch  # global batch of one step
    ids, lengths, sid = make_inputs(GB)
    per = GB // world
    lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else GB
    my_lengths, my_sid = lengths[lo:hi], sid[lo:hi]
    if distributed:
        # NCCL scatter of the padded id tensor from rank 0 (north_star: the only collectives on the
        # path are the id scatter and the PCM gather)
        from mimic3_b200.shard import gather_pcm, scatter_ids, shard_bounds
        lo, hi = shard_bounds(GB, world, rank)
        d_ids, d_len, d_sid = scatter_ids(ids if rank == 0 else None, lengths if rank == 0 else None,
                                          sid if rank == 0 else None, torch.device("cuda", local_rank))
        my_lengths, my_sid = d_len.cpu().numpy(), d_sid.cpu().numpy()
    else:
        d_ids = torch.from_numpy(ids[lo:hi]).cuda()
    h_ids = np.ascontiguousarray(ids[lo:hi])
    torch.cuda.synchronize()

    def step_resident(seed, timing=False):
        return sess.infer(IDS_PER_UTT, my_lengths, scales, my_sid, seed=seed, host_copy=False,
                          device_ids_ptr=d_ids.data_ptr(), stage_timing=timing)

    pcm_dev = torch.empty(1, dtype=torch.int16, device="cuda")

    def step_e2e(seed):
        """Host ids in, int16 PCM back on the host.  N>1: rank 0 owns the host buffers; ids are
        scattered and PCM gathered over NCCL, then rank 0 copies to (pinned) host memory."""
        nonlocal pcm_dev
        if not distributed:
            r = sess.infer(h_ids, my_lengths, scales, my_sid, seed=seed, copy=False)  # PCM lands in pinned host memory
            n = r.total_samples
            assert r.pcm.shape[0] == n
            r.close()
            return n
        di, dl, ds = scatter_ids(ids if rank == 0 else None, lengths if rank == 0 else None,
                                 sid if rank == 0 else None, torch.device("cuda", local_rank))
        cap = int(dl.sum().item()) * 64 * sess.info.hop_length  # generous bound on samples
        if pcm_dev.numel() < cap:
            pcm_dev = torch.empty(cap, dtype=torch.int16, device="cuda")
        r = sess.infer(IDS_PER_UTT, dl.cpu().numpy(), scales, ds.cpu().numpy(), seed=seed, host_copy=False,
                       device_ids_ptr=di.data_ptr(), device_pcm_out=pcm_dev)
        out = gather_pcm(pcm_dev[: r.total_samples], r.sample_offsets, torch.device("cuda", local_rank))
        if rank == 0:
            host = [b.cpu() for b in out[0]]
            return sum(int(h.numel()) for h in host)
        return 0

    def sync_all():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    for k in range(args.warmup):
        step_resident(1000 + k)
    # ---- timed region: device-resident inputs --------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sync_all()
    t0 = time.perf_counter()
    samples = launches = 0
    dev_ms = 0.0
    frames = 0
    for k in range(args.steps):
        r = step_resident(2000 + k)
        samples += r.total_samples
        frames += int(r.frames.sum())
        launches += r.launches
        dev_ms += r.device_ms
    sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    # ---- stage times (separate pass; event pairs around each stage on the engine's stream) -----
    stage_ms = {s: 0.0 for s in STAGES}
    st_frames = 0
    if args.profile_only:
        if rank == 0:
            print(json.dumps({"profile_only": True, "samples_per_step": samples / args.steps,
                              "ms_per_step": wall / args.steps * 1e3}), flush=True)
        return
    for k in range(args.steps):
        r = step_resident(2000 + k, timing=True)
        st_frames += int(r.frames.sum())
        for s in STAGES:
            if "ms:" + s in r.tensors:
                stage_ms[s] += float(r.tensors["ms:" + s][0, 0])

    # ---- end to end through the public call: host ids in, int16 PCM back on the host -------------
    for k in range(2):
        step_e2e(3000 + k)
    sync_all()
    t1 = time.perf_counter()
    e2e_samples = 0
    for k in range(args.steps):
        e2e_samples += step_e2e(2000 + k)
    sync_all()
    e2e_wall = time.perf_counter() - t1

    # ---- reduce over ranks: max time, summed samples ------------------------------------------------
    vec = torch.tensor([wall, dev_ms / 1e3, e2e_wall], dtype=torch.float64, device="cuda")
    if distributed and rank != 0:
        e2e_samples = 0  # rank 0 already counted every gathered sample
    cnt = torch.tensor([samples, e2e_samples, launches, frames], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    wall_max, dev_max, e2e_max = [float(x) for x in vec.tolist()]
    tot_samples, tot_e2e, tot_launch, tot_frames = [float(x) for x in cnt.tolist()]

    if rank == 0:
        value = tot_samples / wall_max
        fl = algorithmic_flops_per_frame(cfg)
        frames_ps = st_frames / args.steps
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        dec_ms = stage_ms.get("dec_last", 0.0) / args.steps
        mrf_ms = stage_ms["mrf"] / args.steps
        if dec_ms > 0:   # dominant kernel: the fused last generator stage (upsample + MRF + conv_post)
            k_name = "dec_fused_kernel (ConvTranspose + MRF + conv_post, last generator stage)"
            k_flops = (fl["ups_stage"][-1] + fl["mrf_stage"][-1] + fl["post"]) * frames_ps
            k_ms = dec_ms
            traffic = 397.0e6 * (GB / 64.0) / world  # dram read+write of one launch, ncu --set full at batch 64 (profiles/r01b_ncu_full_ws_kernels.md), scaled to this batch
        else:
            k_name, k_flops, k_ms, traffic = "MRF stages (mrf_tc_kernel x3)", fl["mrf"] * frames_ps, mrf_ms, None
        achieved_tf = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        mrf_all_tf = fl["mrf"] * frames_ps / ((mrf_ms + dec_ms) * 1e-3) / 1e12 if (mrf_ms + dec_ms) > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_max / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "fp16 tensor-core operands + fp32 accumulate/residual (flow, decoder); f32 (text encoder, durations)",
            "data": "synthetic",
            "config": {"workload": "configs[2]: vctk_low-shaped synthetic voice (109 speakers, random weights), "
                                   + (f"{args.batch} x 80 ids per GPU (weak scaling: global batch {GB}), sid=b%109"
                                      if args.scaling == "weak" else
                                      f"one batch of {GB} x 80 ids sharded by rows over the ranks (strong scaling), sid=b%109"),
                       "global_batch": GB, "ids_per_utterance": IDS_PER_UTT,
                       "scales": [float(s) for s in scales], "parallelism": f"batch-shard x{world}",
                       "frames_per_step": tot_frames / args.steps, "samples_per_step": tot_samples / args.steps,
                       "l2": "no flush: per-step activations (GBs) exceed the 126 MB L2 many times over",
                       "timing": "wall clock between barrier+synchronize pairs (>= CUDA-event time), max over ranks",
                       "device_event_ms_per_step": dev_max / args.steps * 1e3},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "roofline": {"kernel": k_name, "bound": "tensor", "achieved": achieved_tf,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                         "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (fp16 = same tensor rate)" if peaks else "fallback 1400",
                         "kernel_ms": k_ms, "kernel_algorithmic_tflop": k_flops / 1e12,
                         "all_mrf_stages_tflops": mrf_all_tf,
                         "note": "smem operand fetch caps SS-mode MMAs at N=32 to 40 % of the tensor peak (DESIGN.md §3)"},
            "e2e": {"value": tot_e2e / e2e_max, "unit": "samples/s",
                    "h2d_bytes_per_step": int(GB * IDS_PER_UTT * 8 + GB * 16),
                    "d2h_bytes_per_step": int(tot_e2e / args.steps * 2)},
            "gpu_launches": int(tot_launch),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(vd, ids, sid, scales, args.cpu_baseline_utts)
            except Exception as e:  # the CPU leg must never cost the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": 0, "kind": "port",
                                        "sample": f"failed: {e!r}"}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def pick_cpu_threads(orc, ids, sid, scales, probe_utts: int = 4):
    """The oracle port is a batch-1 stream of small ops: more intra-op threads is not always faster (64 threads
    lose to 8-16 on these shapes).  Probe a few counts on a handful of utterances and keep the fastest, so the CPU
    arm is reported at the best this port does on the box; returns (threads, {threads: samples/s})."""
    import torch
    from oracle.vits_oracle import audio_float_to_int16
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    tried = {}
    for t in sorted({t for t in (4, 8, 16, 32, cores) if 1 <= t <= cores}):
        torch.set_num_threads(t)
        orc.infer(ids[0, :20], scales, sid=int(sid[0]))  # thread pool warm-up at this size
        t0 = time.perf_counter()
        n = sum(audio_float_to_int16(orc.infer(ids[b], scales, sid=int(sid[b]), seed=1234, row=b)).size
                for b in range(probe_utts))
        tried[t] = n / (time.perf_counter() - t0)
    best = max(tried, key=tried.get)
    torch.set_num_threads(best)
    return best, tried


def cpu_baseline(vd, ids, sid, scales, n_utts):
    from oracle.vits_oracle import VitsOracle, audio_float_to_int16
    orc = VitsOracle(str(vd))
    threads, tried = pick_cpu_threads(orc, ids, sid, scales)
    t0 = time.perf_counter()
    n = 0
    for b in range(n_utts):
        n += audio_float_to_int16(orc.infer(ids[b], scales, sid=int(sid[b]), seed=1234, row=b)).size
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{n_utts} utterances of the batch, batch-1 loop, torch fp32 CPU ({dt:.1f} s), {threads} intra-op threads "
                      f"(fastest of {sorted(tried)} on this host)"}


if __name__ == "__main__":
    main()
