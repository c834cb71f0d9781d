"""Regenerates the fixtures in tests/golden/ (run in the build container).

1. ``int16_reference.npz`` -- inputs/outputs of the REAL reference function
   ``audio_float_to_int16`` imported from /root/reference/mimic3_tts/utils.py:237-244
   (the only arithmetic of the hot path that lives in the reference repo itself;
   that module imports standalone).  Pins the oracle's restatement.
2. ``golden_wav_stats.json`` -- properties of the reference's own golden WAVs
   (tests/apope_sample_*.wav): length multiple of hop 256, peak == 32767.
3. ``oracle_tiny.npz`` -- oracle outputs for seeded synthetic voices: a regression pin
   that the oracle computes the same thing on the GPU box as here (NOT a reference pin:
   generator.onnx + onnxruntime are unobtainable, parity at that boundary is unpinned).
"""
import importlib.util
import json
import sys
import tempfile
import wave
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
HERE = Path(__file__).resolve().parent


def main():
    spec = importlib.util.spec_from_file_location("m3utils", "/root/reference/mimic3_tts/utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(20260922)
    cases = {}
    for i, (n, amp) in enumerate([(1000, 0.3), (777, 1.0), (64, 0.004), (5000, 0.05), (3, 0.9), (4096, 2.5)]):
        x = (np.tanh(rng.standard_normal(n)) * amp).astype(np.float32) if amp <= 1 else (rng.standard_normal(n) * amp).astype(np.float32)
        cases[f"in{i}"] = x
        cases[f"out{i}"] = ref.audio_float_to_int16(x)
    cases["in6"] = np.zeros(16, dtype=np.float32)
    cases["out6"] = ref.audio_float_to_int16(cases["in6"])
    np.savez_compressed(HERE / "int16_reference.npz", **cases)

    stats = {}
    for arch in ("amd64", "arm64", "armv7"):
        with wave.open(f"/root/reference/tests/apope_sample_{arch}.wav", "rb") as w:
            n = w.getnframes()
            a = np.frombuffer(w.readframes(n), dtype="<i2")
            stats[arch] = {"frames": n, "rate": w.getframerate(), "width": w.getsampwidth(),
                           "channels": w.getnchannels(), "peak": int(np.abs(a.astype(np.int32)).max()),
                           "mod_hop256": int(n % 256)}
    (HERE / "golden_wav_stats.json").write_text(json.dumps(stats, indent=1))

    from mimic3_b200 import synth_voice as sv
    from oracle.vits_oracle import VitsOracle, audio_float_to_int16
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, cfg, seed in (("tiny", sv.tiny_config(), 11), ("tiny_ms", sv.tiny_config(n_speakers=3), 12),
                                ("tiny_rb1_dp", sv.tiny_config(n_speakers=2, resblock="1", use_sdp=False), 13)):
            sv.write_voice(Path(d) / name, cfg, seed=seed)
            o = VitsOracle(str(Path(d) / name))
            ids = np.random.default_rng(5).integers(4, cfg.num_symbols, size=23).astype(np.int64)
            audio, inter = o.infer(ids, (0.0, 1.0, 0.0), sid=1, return_intermediates=True)
            out[f"{name}_ids"] = ids
            out[f"{name}_durations"] = inter["durations"]
            out[f"{name}_audio"] = audio
            out[f"{name}_pcm"] = audio_float_to_int16(audio)
            noisy = o.infer(ids, (0.667, 1.1, 0.8), sid=1, seed=77, row=0)
            out[f"{name}_noisy_audio"] = noisy
    np.savez_compressed(HERE / "oracle_tiny.npz", **out)
    print("wrote", sorted(p.name for p in HERE.glob("*")))


if __name__ == "__main__":
    main()
