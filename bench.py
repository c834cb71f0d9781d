#!/usr/bin/env python
"""Benchmark of the ids->waveform hot path (contract: see the task description / DESIGN.md §5).

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine (libm3b200)
    python bench.py --impl reference --gpus N --steps K ...    # the CPU path, rank 0 only

Default workload = BASELINE.json configs[2] (the one the metric is quoted on): an en_US/vctk_low-shaped
voice (109 speakers, synthetic random weights -- no real voice is reachable offline), batch 256,
80 ids per utterance, sid[b] = b mod 109, PCG64(1234).  ``--workload cfg2|cfg4|cfg5`` select the other
BASELINE configs (profiles/ lines, not the headline).  Utterances are independent, so every rank runs
its own rows with no collective on the compute path: by default per-GPU work is fixed (weak scaling,
global batch 256 x N, as the bench contract prescribes for a path that partitions); ``--scaling strong``
shards ONE batch over the N ranks instead (BASELINE configs[2] literally: 32 utterances per GPU at N = 8).
A "step" = one pass of the hot path over the whole (global) batch.  Prints ONE JSON line on rank 0.

Two timed regions:
  value  ids already resident in HBM, PCM left in HBM (kernel-resident throughput);
  e2e    host ids -> engine -> int16 PCM in (pinned) host memory on rank 0.  N>1: rank 0 scatters the ids
         (pinned H2D + NCCL) and gathers the PCM (NCCL send/recv + one D2H) every step
         (mimic3_b200/shard.py); the gather + D2H of step k overlap the compute of step k+1 (two slots),
         every byte of every step has landed on the host before the closing timestamp.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
from concurrent.futures import ThreadPoolExecutor
import time
from collections import deque
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

# name -> (sub-directory, speakers, symbols, seed): synthetic stand-ins with the shipped voices' shapes
VOICES = {
    "vctk_low": ("en_US/vctk_low", N_SPEAKERS, NUM_SYMBOLS, VOICE_SEED),
    "thorsten_low": ("de_DE/thorsten_low", 1, 56, 23),
    "siwis_low": ("fr_FR/siwis_low", 1, 44, 24),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_voice(root: Path, name: str = "vctk_low") -> Path:
    """Synthetic <name>-shaped voice; written by local rank 0, shared through the filesystem."""
    from mimic3_b200 import synth_voice as sv
    sub, nspk, nsym, seed = VOICES[name]
    d = root / sub
    marker = d / ".complete"
    if not marker.exists():
        sv.write_voice(d, sv.low_config(n_speakers=nspk, num_symbols=nsym), seed=seed)
        marker.write_text("ok")
    return d


def make_inputs(batch: int = GLOBAL_BATCH, ids_per_utt: int = IDS_PER_UTT, num_symbols: int = NUM_SYMBOLS,
                n_speakers: int = N_SPEAKERS, all_speakers: bool = True):
    """PCG64(1234), fixed-length rows, ids uniform over the real symbols (pad/bos/eos/blank = 0..3
    excluded), sid = b mod n_speakers (cfg3) or 0 (cfg2: p239 is index 0, voices.json:735-736)."""
    rng = np.random.Generator(np.random.PCG64(1234))
    ids = rng.integers(4, num_symbols, size=(batch, ids_per_utt)).astype(np.int64)
    lengths = np.full(batch, ids_per_utt, dtype=np.int64)
    sid = (np.arange(batch) % n_speakers).astype(np.int64) if all_speakers else np.zeros(batch, dtype=np.int64)
    return ids, lengths, sid


def make_jobs(workload: str, batch: int):
    """One step = these engine calls, in order: dicts(voice, ids, lengths, sid|None, scales|None=voice defaults)."""
    if workload == "cfg3":
        ids, lengths, sid = make_inputs(batch)
        return [dict(voice="vctk_low", ids=ids, lengths=lengths, sid=sid, scales=None)]
    if workload == "cfg2":   # B=32 x 100 ids, speaker p239 (index 0)
        ids, lengths, sid = make_inputs(batch, 100, all_speakers=False)
        return [dict(voice="vctk_low", ids=ids, lengths=lengths, sid=sid, scales=None)]
    if workload == "cfg4":   # long form: B=8 x 1800 ids, single-speaker voice, three length scales per step
        ids, lengths, _ = make_inputs(batch, 1800, VOICES["thorsten_low"][2])
        return [dict(voice="thorsten_low", ids=ids, lengths=lengths, sid=None, scales=(0.667, ls, 0.8))
                for ls in (0.8, 1.0, 1.2)]
    if workload == "cfg5":   # mixed voices, rows grouped by voice (43/43/42 of 128), all three resident
        names = ["vctk_low", "thorsten_low", "siwis_low"]
        base, extra = divmod(batch, 3)
        jobs = []
        for i, n in enumerate(names):
            rows = base + (1 if i < extra else 0)
            ids, lengths, sid = make_inputs(rows, IDS_PER_UTT, VOICES[n][2], VOICES[n][1])
            jobs.append(dict(voice=n, ids=ids, lengths=lengths, sid=sid if VOICES[n][1] > 1 else None, scales=None))
        return jobs
    raise SystemExit(f"unknown workload {workload}")


WORKLOAD_TEXT = {
    "cfg3": "configs[2]: vctk_low-shaped synthetic voice (109 speakers, random weights), {b} x 80 ids, sid=b%109",
    "cfg2": "configs[1]: vctk_low-shaped synthetic voice, speaker index 0 (p239), {b} x 100 ids",
    "cfg4": "configs[3]: thorsten_low-shaped synthetic voice (single speaker), {b} x 1800 ids, one pass per length_scale in (0.8, 1.0, 1.2) per step",
    "cfg5": "configs[4]: mixed batch of {b} x 80 ids over three resident voices (vctk_low 109 spk / thorsten_low / siwis_low), rows grouped by voice, one engine call per voice per step",
}
DEFAULT_BATCH = {"cfg3": 256, "cfg2": 32, "cfg4": 8, "cfg5": 128}


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


def probe_reference_stack() -> dict:
    """BASELINE.md §3.1 / SURVEY §7 step 0, performed at RUN TIME on the box the numbers come from: is the
    reference's own arithmetic (onnxruntime + a real generator.onnx) reachable here?  The result is printed in the
    bench line; the CPU arm uses onnxruntime when the probe finds it, else the oracle port (kind "port")."""
    out = {"onnxruntime": None, "baseline_ref": False, "mimic3_tts": False, "voices_found": [], "searched": []}
    ref = ROOT / "baseline" / "_ref"
    out["baseline_ref"] = ref.is_dir()
    if ref.is_dir() and str(ref) not in sys.path:
        sys.path.append(str(ref))
    try:
        import onnxruntime  # noqa: F401
        out["onnxruntime"] = getattr(onnxruntime, "__version__", "?")
    except Exception as e:
        out["onnxruntime_error"] = type(e).__name__
    try:
        import importlib.util
        out["mimic3_tts"] = importlib.util.find_spec("mimic3_tts") is not None
    except Exception:
        pass
    dirs = []
    home = os.environ.get("XDG_DATA_HOME") or os.path.join(os.path.expanduser("~"), ".local", "share")
    dirs.append(Path(home) / "mycroft" / "mimic3" / "voices")                # const.py:25-27
    for d in (os.environ.get("XDG_DATA_DIRS") or "/usr/local/share:/usr/share").split(":"):
        if d:
            dirs.append(Path(d) / "mycroft" / "mimic3" / "voices")          # tts.py:160-181
    dirs.append(ref / "voices")
    for d in dirs:
        out["searched"].append(str(d))
        try:
            if d.is_dir():
                out["voices_found"] += [str(p.parent) for p in sorted(d.glob("*/*/generator.onnx"))][:8]
        except OSError:
            pass
    return out


def ort_cpu_baseline(probe: dict, n_utts: int):
    """The reference's own CPU path, if the probe found it: onnxruntime CPUExecutionProvider, default session
    options, B=1 loop, timed like voice.py:229-232 (run + audio_float_to_int16).  Returns None when unavailable."""
    if not probe.get("onnxruntime") or not probe.get("voices_found"):
        return None
    import onnxruntime
    vd = Path(next((v for v in probe["voices_found"] if v.endswith("vctk_low")), probe["voices_found"][0]))
    cfgj = json.loads((vd / "config.json").read_text())
    sess = onnxruntime.InferenceSession(str(vd / "generator.onnx"), providers=["CPUExecutionProvider"])
    names = {i.name for i in sess.get_inputs()}
    nsym = int(cfgj["model"]["num_symbols"])
    nspk = int(cfgj["model"].get("n_speakers", 1))
    inf = cfgj.get("inference", {})
    scales = np.array([inf.get("noise_scale", 0.667), inf.get("length_scale", 1.0), inf.get("noise_w", 0.8)], dtype=np.float32)
    ids, lengths, sid = make_inputs(n_utts, IDS_PER_UTT, nsym, max(nspk, 1))

    def one(b):
        feed = {"input": ids[b:b + 1], "input_lengths": lengths[b:b + 1], "scales": scales}
        if "sid" in names:
            feed["sid"] = sid[b:b + 1]
        audio = sess.run(None, feed)[0].squeeze()
        peak = max(0.01, float(np.max(np.abs(audio))))
        return np.clip(audio * (32767.0 / peak), -32767, 32767).astype("int16").size
    for b in range(min(3, n_utts)):
        one(b)
    t0 = time.perf_counter()
    n = sum(one(b) for b in range(n_utts))
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": f"onnxruntime {probe['onnxruntime']} CPUExecutionProvider on {vd}, {n_utts} utterances x 80 ids, "
                      f"batch-1 loop ({dt:.1f} s), default session options"}


def run_reference(args, rank: int, world: int):
    """CPU arm: onnxruntime if the run-time probe finds it (kind "reference"), else the oracle port (kind "port")
    on rank 0's host cores, B=1 loop like the reference (voice.py:180-181), bounded sample per step."""
    if rank != 0:
        return
    probe = probe_reference_stack()
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
    base = {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{sample} utterances/step x {args.steps} steps, torch fp32 CPU, {cores} threads"}
    try:
        ort = ort_cpu_baseline(probe, sample * args.steps)
    except Exception as e:
        ort = None
        probe["ort_run_error"] = repr(e)
    if ort:
        base, value = ort, ort["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: vctk_low-shaped synthetic voice, 109 speakers, batch=256 x 80 ids",
                   "sample": f"{sample} utterances of the batch per step, batch-1 loop"},
        "cpu_baseline": base, "reference_probe": probe,
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measured_traffic(kernel: str):
    """DRAM read+write bytes per processed audio sample of `kernel`, from the newest committed ncu capture
    (profiles/traffic.json, written by tools/summarize_ncu.py --traffic from a `ncu --set full` run of this
    bench command); None when no capture has been summarised."""
    try:
        t = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        e = t.get(kernel)
        return (float(e["dram_bytes_per_sample"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--ref-utts", type=int, default=32, help="utterances per step for the CPU arm")
    ap.add_argument("--cpu-baseline-utts", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--deterministic", action="store_true", help="noise scales 0 (parity settings)")
    ap.add_argument("--batch", type=int, default=0, help="batch override (0 = the workload's own batch)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch utterances PER GPU (default); strong: --batch utterances sharded over all GPUs")
    ap.add_argument("--profile-only", action="store_true", help="skip stage/e2e/cpu passes (ncu runs)")
    ap.add_argument("--e2e-depth", type=int, default=2, help="slots of the PCM collector (1 = no overlap of gather/D2H with compute)")
    ap.add_argument("--gather", choices=("host", "nccl"), default="host",
                    help="N>1 end-to-end PCM path: 'host' = every rank copies its PCM over its own PCIe link into one shared pinned "
                         "host segment that rank 0 reads (no payload collective); 'nccl' = send/recv to rank 0's GPU, one D2H there")
    args = ap.parse_args()
    if not args.batch:
        args.batch = DEFAULT_BATCH[args.workload]

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
    from mimic3_b200.shard import shard_bounds

    # -- jobs of one step: every rank builds the same global jobs, then keeps its rows ----------------------
    GB = args.batch * world if args.scaling == "weak" else args.batch  # global batch of one step
    jobs = make_jobs(args.workload, GB)

    # -- voices: written once (local rank 0), read by every rank; all stay resident -------------------------
    shared = Path(os.environ.get("M3B200_BENCH_DIR", tempfile.gettempdir())) / f"m3b200_bench_voice_{os.environ.get('MASTER_PORT', 'single')}"
    names = sorted({j["voice"] for j in jobs})
    if local_rank == 0:
        for n in names:
            make_voice(shared, n)
    if distributed:
        dist.barrier()
    sessions = {n: B200Session(str(make_voice(shared, n)), device=local_rank) for n in names}
    vd = make_voice(shared, "vctk_low") if "vctk_low" in names else None
    cfg = sv.low_config(n_speakers=N_SPEAKERS, num_symbols=NUM_SYMBOLS)

    dev = torch.device("cuda", local_rank)
    for j in jobs:
        s = sessions[j["voice"]]
        j["sess"] = s
        if j["scales"] is None:
            j["scales"] = (s.info.noise_scale, s.info.length_scale, s.info.noise_w)
        if args.deterministic:
            j["scales"] = (0.0, j["scales"][1], 0.0)
        lo, hi = shard_bounds(j["ids"].shape[0], world, rank)
        j["lo"], j["hi"] = lo, hi
        j["my_lengths"] = j["lengths"][lo:hi]
        j["my_sid"] = None if j["sid"] is None else j["sid"][lo:hi]
        j["h_ids"] = np.ascontiguousarray(j["ids"][lo:hi])
        j["d_ids"] = torch.from_numpy(j["h_ids"]).cuda() if hi > lo else None
        j["T"] = int(j["ids"].shape[1])
    torch.cuda.synchronize()

    def step_resident(seed, timing=False):
        out = []
        for j in jobs:
            if j["hi"] == j["lo"]:
                continue
            out.append(j["sess"].infer(j["T"], j["my_lengths"], j["scales"], j["my_sid"], seed=seed, host_copy=False,
                                       device_ids_ptr=j["d_ids"].data_ptr(), stage_timing=timing))
        return out

    # -- end to end: host ids in, int16 PCM back on the host ---------------------------------------------------
    scat = coll = None
    if distributed and not args.profile_only:
        from mimic3_b200.shard import HostPcmCollector, IdScatter, PcmCollector, make_groups
        gather_pg, meta_pg = make_groups(dev)
        max_rows = max(j["ids"].shape[0] for j in jobs)
        max_t = max(j["T"] for j in jobs)
        per = (max_rows + world - 1) // world
        # two scatter buffer sets: the ids of call k+1 are scattered while call k runs
        scats = [IdScatter(max_rows, max_t, dev, payload_group=None, meta_group=meta_pg) for _ in range(2)]
        scat = scats[0]
        # capacity per rank: ids x 7 frames/id x hop (the synthetic voices give 4.2-4.7 frames per id; the whole segment is pinned,
        # so a looser bound costs real /dev/shm pages: world x depth x cap x 2 bytes)
        cap = per * max_t * 7 * 256
        gather = args.gather
        if gather == "host":  # the shared segment must fit /dev/shm (a tmpfs that is too small kills the process with SIGBUS)
            ok = [True]
            if rank == 0:
                try:
                    st = os.statvfs("/dev/shm")
                    need = 2 * cap * world * max(3, args.e2e_depth)
                    ok[0] = st.f_bavail * st.f_frsize > need + (64 << 20)
                    if not ok[0]:
                        print(f"[bench] /dev/shm has {st.f_bavail * st.f_frsize >> 20} MiB free, the PCM segment needs {need >> 20} MiB: "
                              "falling back to the NCCL gather", file=sys.stderr, flush=True)
                except OSError:
                    ok[0] = False
            dist.broadcast_object_list(ok, src=0, group=meta_pg)
            if not ok[0]:
                gather = "nccl"
        args.gather = gather
        if gather == "host":  # two engine calls in flight per rank (below) + one copy in flight: three slots
            coll = HostPcmCollector(cap, per, dev, meta_group=meta_pg, depth=max(3, args.e2e_depth))
        else:
            coll = PcmCollector(cap, per, dev, payload_group=gather_pg, meta_group=meta_pg, depth=max(1, args.e2e_depth))
    tickets = deque()
    trace = {"scatter": 0.0, "infer": 0.0, "submit": 0.0, "collect": 0.0, "n": 0} if os.environ.get("M3B200_BENCH_TRACE") else None

    def e2e_collect(t):
        got = coll.collect(t)
        if got is None:
            return 0
        pcm, frames = got
        n = int(sum(int(np.sum(f)) for f in frames)) * 256
        have = sum(a.shape[0] for a in pcm) if isinstance(pcm, list) else pcm.shape[0]
        assert have == n, (have, n)
        return n

    # N=1: up to e2e_depth `m3_infer` calls in flight from host threads (the C library releases the GIL and gives every call
    # its own context and stream): call k's device->host copy of the PCM runs under call k+1's kernels, as in the N>1 path
    pool = ThreadPoolExecutor(max_workers=max(1, args.e2e_depth)) if not distributed and args.e2e_depth > 1 else None

    def one_call(j, seed):
        r = j["sess"].infer(j["h_ids"], j["my_lengths"], j["scales"], j["my_sid"], seed=seed, copy=False)
        n = r.total_samples
        assert r.pcm.shape[0] == n
        r.close()
        return n

    dpool = ThreadPoolExecutor(max_workers=2) if distributed else None
    dstate = {"next": None, "k": 0}
    inflight = deque()

    def do_scatter(ji):
        jj = jobs[ji]
        sc = scats[dstate["k"] & 1]
        dstate["k"] += 1
        return sc(jj["ids"] if rank == 0 else None, jj["lengths"] if rank == 0 else None, jj["sid"] if rank == 0 else None)

    def step_e2e(seed):
        """N=1: `m3_infer` with host ids, PCM lands in pinned host memory.  N>1: rank 0 owns the host buffers."""
        n = 0
        for j in jobs:
            if not distributed:
                if pool is None:
                    n += one_call(j, seed)
                else:
                    tickets.append(pool.submit(one_call, j, seed))
                    while len(tickets) >= args.e2e_depth:
                        n += tickets.popleft().result()
                continue
            # N>1, pipelined like N=1: up to two engine calls of this rank in flight in worker threads (the C library releases
            # the GIL; call k+1's kernels fill the tails of call k's persistent kernels), while this thread submits the PCM of
            # the call that just finished (D2H into the shared segment), finishes the one before (wait for its copy, host-side
            # gather of the counts, barrier) and scatters the ids of the next call into the other buffer set.  All collectives
            # stay on this thread, in the same order on every rank.  (The NCCL gather keeps one call in flight.)
            ta = time.perf_counter()
            ji = jobs.index(j)
            if dstate["next"] is None:
                dstate["next"] = do_scatter(ji)
            d_ids, lens, sids = dstate["next"]
            two = args.gather == "host"
            buf = coll.send_buffer(len(inflight)) if two else coll.send_buffer()
            tb = time.perf_counter()
            if len(lens):
                inflight.append(dpool.submit(j["sess"].infer, d_ids.stride(0), lens, j["scales"], sids, seed=seed, host_copy=False,
                                             device_ids_ptr=d_ids.data_ptr(), device_pcm_out=buf))
            else:
                inflight.append(None)
            tc_ = time.perf_counter()
            while len(inflight) > (1 if two else 0):   # the older call: its PCM is final, start its copy
                f = inflight.popleft()
                if f is not None:
                    r = f.result()
                    tickets.append(coll.submit(r.total_samples, r.frames))
                else:
                    tickets.append(coll.submit(0, []))
            td = time.perf_counter()
            while len(tickets) >= 2:
                n += e2e_collect(tickets.popleft())
            te = time.perf_counter()
            dstate["next"] = do_scatter((ji + 1) % len(jobs))   # (the last one of the run is never used)
            if trace is not None:
                tf = time.perf_counter()
                trace["scatter"] += (tb - ta) + (tf - te)
                trace["infer"] += td - tc_
                trace["submit"] += 0.0
                trace["collect"] += te - td
                trace["n"] += 1
        return n

    def e2e_drain():
        n = 0
        while distributed and inflight:
            f = inflight.popleft()
            if f is not None:
                r = f.result()
                tickets.append(coll.submit(r.total_samples, r.frames))
            else:
                tickets.append(coll.submit(0, []))
        while tickets:
            t = tickets.popleft()
            n += t.result() if pool is not None else e2e_collect(t)
        if coll:
            coll.drain()
        return n

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
        for r in step_resident(2000 + k):
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
        for r in step_resident(2000 + k, timing=True):
            st_frames += int(r.frames.sum())
            for s in STAGES:
                if "ms:" + s in r.tensors:
                    stage_ms[s] += float(r.tensors["ms:" + s][0, 0])

    # ---- end to end through the public call ---------------------------------------------------------------
    for k in range(2):
        step_e2e(3000 + k)
    e2e_drain()
    sync_all()
    t1 = time.perf_counter()
    e2e_samples = 0
    for k in range(args.steps):
        e2e_samples += step_e2e(2000 + k)
    e2e_samples += e2e_drain()
    sync_all()
    e2e_wall = time.perf_counter() - t1
    if trace is not None and trace["n"]:
        log(f"[bench trace] rank {rank}: per engine call (ms): " + ", ".join(
            f"{k} {trace[k] / trace['n'] * 1e3:.3f}" for k in ("scatter", "infer", "submit", "collect")))

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
        traffic = traffic_src = None
        if dec_ms > 0:   # dominant kernel: the fused last generator stage (upsample + MRF + conv_post)
            k_name = "dec_planes_kernel (ConvTranspose + MRF + conv_post, last generator stage; dec_fused_kernel with M3B200_DEC_V2=1)"
            k_flops = (fl["ups_stage"][-1] + fl["mrf_stage"][-1] + fl["post"]) * frames_ps
            k_ms = dec_ms
            per_sample, traffic_src = measured_traffic("dec_fused_kernel" if os.environ.get("M3B200_DEC_V2") else "dec_planes_kernel")
            if per_sample is not None:
                traffic = per_sample * frames_ps * 256
        else:
            k_name, k_flops, k_ms = "MRF stages (mrf_tc_kernel x3)", fl["mrf"] * frames_ps, mrf_ms
        achieved_tf = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        mrf_all_tf = fl["mrf"] * frames_ps / ((mrf_ms + dec_ms) * 1e-3) / 1e12 if (mrf_ms + dec_ms) > 0 else 0.0
        h2d = sum(int(j["ids"].size * 8 + j["ids"].shape[0] * 16) for j in jobs)
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_max / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "fp16 tensor-core operands + fp32 accumulate/residual (flow, decoder); f32 (text encoder, durations)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload].format(b=GB)
                                   + (f"; {args.batch} rows per GPU (weak scaling: global batch {GB})" if args.scaling == "weak"
                                      else f"; one batch of {GB} rows sharded over the ranks (strong scaling)"),
                       "global_batch": GB, "ids_per_utterance": jobs[0]["T"],
                       "scales": [[float(s) for s in j["scales"]] for j in jobs] if len(jobs) > 1 else [float(s) for s in jobs[0]["scales"]],
                       "parallelism": f"batch-shard x{world}",
                       "frames_per_step": tot_frames / args.steps, "samples_per_step": tot_samples / args.steps,
                       "l2": "no flush: per-step activations (GBs) exceed the 126 MB L2 many times over",
                       "timing": "wall clock between barrier+synchronize pairs (>= CUDA-event time), max over ranks",
                       "device_event_ms_per_step": dev_max / args.steps * 1e3,
                       "e2e_pipeline": ((f"N>1: id scatter (pinned H2D + NCCL) every step, two engine calls per rank in flight in worker threads, ids of call k+1 scattered meanwhile; PCM: every rank copies its own int16 PCM over its own "
                                         f"PCIe link into its slice of ONE shared pinned host segment that rank 0 reads (no payload collective), "
                                         f"{max(1, args.e2e_depth)} slots: step k's copies overlap step k+1's compute"
                                         if args.gather == "host" else
                                         f"N>1: id scatter (pinned H2D + NCCL) and PCM gather (NCCL send/recv + one D2H to pinned memory on "
                                         f"rank 0) every step, {max(1, args.e2e_depth)} slots: step k's gather/D2H overlap step k+1's compute"))
                                       if distributed else (f"N=1: m3_infer with host ids, PCM copied to pinned host memory inside the call; {max(1, args.e2e_depth)} calls in "
                                                            f"flight from host threads (call k's D2H under call k+1's kernels)")},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "roofline": {"kernel": k_name, "bound": "tensor", "achieved": achieved_tf,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (fp16 = same tensor rate)" if peaks else "fallback 1400",
                         "kernel_ms": k_ms, "kernel_algorithmic_tflop": k_flops / 1e12,
                         "all_mrf_stages_tflops": mrf_all_tf,
                         "note": "smem operand fetch caps SS-mode MMAs at N=32 to 40 % of the tensor peak (DESIGN.md §3)"},
            "e2e": {"value": tot_e2e / e2e_max, "unit": "samples/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(tot_e2e / args.steps * 2),
                    "ms_per_step": e2e_max / args.steps * 1e3,
                    "note": (None if distributed or args.e2e_depth <= 1 else
                             f"{args.e2e_depth} m3_infer calls in flight: call k+1's kernels also fill the tails of call k's persistent kernels and "
                             "the host round trip of the duration predictor, so this can exceed `value`, which times one call at a time")},
            "gpu_launches": int(tot_launch),
            "clocks": clocks,
            "reference_probe": probe_reference_stack(),
        }
        if not args.no_cpu_baseline and vd is not None:
            j0 = next(j for j in jobs if j["voice"] == "vctk_low")
            try:
                ort = ort_cpu_baseline(line["reference_probe"], args.cpu_baseline_utts)
            except Exception as e:
                ort = None
                line["reference_probe"]["ort_run_error"] = repr(e)
            try:
                line["cpu_baseline"] = ort or cpu_baseline(vd, j0["ids"], j0["sid"], j0["scales"], args.cpu_baseline_utts)
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
    n_utts = min(n_utts, ids.shape[0])
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
