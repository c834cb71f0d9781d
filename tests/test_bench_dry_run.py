"""bench.py's GPU arm executed on the CPU with a fake engine: every line of the arm runs (argument handling, the
resident / stage-timing / end-to-end passes, the JSON assembly with roofline, e2e, clocks and cpu_baseline), so a typo
in a code path that only a GPU box reaches cannot cost the round's benchmark line.  Numbers are meaningless here."""
import io
import json
import sys
from contextlib import redirect_stdout
from types import SimpleNamespace
from unittest import mock

import numpy as np
import pytest


class FakeResult:
    def __init__(self, batch, stage_timing):
        from mimic3_b200.engine import STAGES
        self.frames = np.full(batch, 355, dtype=np.int64)
        self.sample_offsets = np.concatenate([[0], np.cumsum(self.frames * 256)])
        self.total_samples = int(self.sample_offsets[-1])
        self.launches, self.device_ms = 110, 16.5
        self.pcm = np.zeros(self.total_samples, dtype=np.int16)
        self.tensors = {"ms:" + s: np.array([[1.0]], dtype=np.float32) for s in STAGES} if stage_timing else {}

    def close(self):
        pass


class FakeSession:
    def __init__(self, path, device=0, **kw):
        self.info = SimpleNamespace(noise_scale=0.667, length_scale=1.0, noise_w=0.8, hop_length=256)

    def infer(self, ids, lengths, scales, sid, stage_timing=False, **kw):
        return FakeResult(len(lengths), stage_timing)


@pytest.mark.parametrize("extra", [[], ["--scaling", "strong", "--no-cpu-baseline"], ["--profile-only"],
                                   ["--workload", "cfg2", "--no-cpu-baseline"], ["--workload", "cfg4", "--no-cpu-baseline"],
                                   ["--workload", "cfg5", "--no-cpu-baseline"]])
def test_gpu_arm_runs_end_to_end_with_a_fake_engine(extra, monkeypatch, tmp_path):
    import torch
    import bench
    import mimic3_b200.engine as engine

    real_empty, real_tensor = torch.empty, torch.tensor
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch, "empty", lambda *a, device=None, **k: real_empty(*a, **k))
    monkeypatch.setattr(torch, "tensor", lambda *a, device=None, **k: real_tensor(*a, **k))
    monkeypatch.setattr(engine, "B200Session", FakeSession)
    monkeypatch.setattr(bench, "cpu_baseline", lambda vd, ids, sid, scales, n: {"value": 1.0, "unit": "samples/s", "cores": 1,
                                                                              "kind": "port", "sample": "fake"})
    monkeypatch.setenv("M3B200_BENCH_DIR", str(tmp_path))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "1", "--batch", "8"] + extra)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.main()
    lines = [l for l in out.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    if "--profile-only" in extra:
        assert d["profile_only"] is True
        return
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "stage_ms_per_step"):
        assert key in d, key
    assert d["scaling"] == ("strong" if "strong" in extra else "weak") and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["config"]["global_batch"] == 8 and "workload" in d["config"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    calls = 3 if ("cfg4" in extra or "cfg5" in extra) else 1      # engine calls per step
    assert d["gpu_launches"] == 2 * 110 * calls and d["e2e"]["value"] > 0
    assert "reference_probe" in d and "onnxruntime" in d["reference_probe"]
    assert ("cpu_baseline" in d) == ("--no-cpu-baseline" not in extra)
