"""CPU tests: the C-ABI library loads and exports every declared symbol; host-side mirror of
Mimic3Voice behaves like the reference (mimic3_tts/voice.py:154-243); loader error paths."""
import re
import ctypes
from pathlib import Path

import numpy as np
import pytest

from mimic3_b200 import engine
from mimic3_b200.voice import B200Voice, VoiceConfig, load_phoneme_ids

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(built_library):
    header = (ROOT / "include" / "m3b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(m3_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(engine.API_SYMBOLS)
    lib = ctypes.CDLL(str(built_library))
    for name in declared:
        assert hasattr(lib, name), name
    lib.m3_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.m3_version()


def test_library_is_sm100a_only(built_library):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(built_library)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_loader_errors_without_gpu_are_specific(built_library, voices, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    lib = engine.load_library()
    with pytest.raises(engine.B200EngineError, match="no CUDA device|no CPU fallback"):
        engine.B200Session(str(voices("tiny_ms_folded")))  # parsed + bound fine, then: no GPU
    with pytest.raises(FileNotFoundError):
        engine.B200Session(str(tmp_path / "nope"))
    # a voice whose generator.onnx lacks a tensor -> model error naming the tensor
    from mimic3_b200 import synth_voice as sv
    cfg = sv.tiny_config()
    params = sv.make_params(cfg, 1)
    sv.write_voice(tmp_path / "bad", cfg, seed=1)
    del params["dec.conv_post.weight"]
    sv.write_generator_onnx(tmp_path / "bad" / "generator.onnx", cfg, params)
    with pytest.raises(engine.B200EngineError, match="dec.conv_post.weight"):
        engine.B200Session(str(tmp_path / "bad"))
    # shape mismatch between config.json and the graph
    params = sv.make_params(cfg, 1)
    params["enc_p.proj.weight"] = params["enc_p.proj.weight"][:-2]
    sv.write_generator_onnx(tmp_path / "bad" / "generator.onnx", cfg, params)
    with pytest.raises(engine.B200EngineError, match="enc_p.proj.weight"):
        engine.B200Session(str(tmp_path / "bad"))
    (tmp_path / "bad" / "generator.onnx").write_bytes(b"\x08\x07garbage")
    with pytest.raises(engine.B200EngineError):
        engine.B200Session(str(tmp_path / "bad"))


class FakeSession:
    """Records what the host layer sends to the engine."""

    class info:
        has_speaker_embedding = 1

    def __init__(self):
        self.calls = []

    def infer(self, text, lengths, scales, sid, seed=0, **kw):
        self.calls.append(dict(text=text.copy(), lengths=lengths.copy(), scales=scales.copy(),
                               sid=None if sid is None else sid.copy(), seed=seed))

        class R:
            total_samples = int(lengths.sum()) * 4
            def utterance_pcm(self_, b):
                return np.zeros(int(lengths[b]) * 4, dtype=np.int16)
        return R()


def _voice(multispeaker=True, speaker_map=None):
    cfg = VoiceConfig({"model": {"n_speakers": 3 if multispeaker else 1},
                       "inference": {"length_scale": 1.2, "noise_scale": 0.5, "noise_w": 0.7}})
    return B200Voice(cfg, FakeSession(), {}, None, speaker_map)


def test_ids_to_audio_builds_reference_inputs():
    v = _voice(speaker_map={"p239": 2, "alias": 1})
    v.ids_to_audio([3, 4, 5, 6], speaker="p239", rate=2.0)
    c = v.onnx_model.calls[-1]
    assert c["text"].dtype == np.int64 and c["text"].shape == (1, 4)          # voice.py:180
    assert c["lengths"].tolist() == [4]                                        # voice.py:181
    np.testing.assert_allclose(c["scales"], [0.5, 0.6, 0.7])                   # [noise, length/rate, noise_w] :182-189, :170
    assert c["scales"].dtype == np.float32 and c["sid"].tolist() == [2]
    v.ids_to_audio([3], speaker="1")               # not in map -> int()  (voice.py:203-205)
    assert v.onnx_model.calls[-1]["sid"].tolist() == [1]
    v.ids_to_audio([3], speaker="nobody")          # warning + first speaker (voice.py:206-211)
    assert v.onnx_model.calls[-1]["sid"].tolist() == [0]
    v.ids_to_audio([3], speaker=2, length_scale=0.9, noise_scale=0.0, noise_w=0.0, rate=0)
    c = v.onnx_model.calls[-1]
    np.testing.assert_allclose(c["scales"], [0.0, 0.9, 0.0])                   # rate <= 0: no scaling (voice.py:170)
    assert c["sid"].tolist() == [2] and c["seed"] == 0
    single = _voice(multispeaker=False)
    single.ids_to_audio([1, 2], speaker="p239")
    assert single.onnx_model.calls[-1]["sid"] is None                          # no "sid" key (voice.py:197)


def test_batch_padding_and_lengths():
    v = _voice(speaker_map={"a": 1})
    out = v.ids_to_audio_batch([[1, 2, 3], [4], [5, 6]], speakers=["a", None, 2], noise_scale=0, noise_w=0)
    c = v.onnx_model.calls[-1]
    assert c["text"].shape == (3, 3) and c["lengths"].tolist() == [3, 1, 2]
    assert c["text"][1].tolist() == [4, 0, 0] and c["sid"].tolist() == [1, 0, 2]
    assert [len(o) for o in out] == [12, 4, 8]


def test_voice_directory_files(voices):
    d = voices("tiny_ms")
    with open(d / "phonemes.txt", encoding="utf-8") as f:
        p2i = load_phoneme_ids(f)
    assert p2i["_"] == 0 and len(p2i) == 20
    cfg = VoiceConfig.load(open(d / "config.json"))
    assert cfg.is_multispeaker and cfg.audio.sample_rate == 22050
    assert cfg.inference.noise_w == pytest.approx(0.8)


def test_entry_points_reject_null_arguments_without_a_gpu(built_library):
    """Argument errors are reported, never dereferenced: no GPU needed to see that."""
    lib = engine.load_library()
    res = ctypes.c_void_p()
    assert lib.m3_infer_ex(None, None, None, 1, 1, None, None, None, ctypes.byref(res)) == engine.M3_ERR_INVALID
    assert b"NULL" in lib.m3_last_error()
    assert lib.m3_infer(None, None, None, 1, 1, None, None, 0, 0, ctypes.byref(res)) == engine.M3_ERR_INVALID
    n = ctypes.c_int64(-1)
    assert not lib.m3_result_stream(None, ctypes.byref(n)) and n.value == 0
    assert lib.m3_result_batch(None) == 0 and lib.m3_result_kernel_launches(None) == 0
    lib.m3_result_free(None)
    lib.m3_voice_free(None)
    assert ctypes.sizeof(engine.InferOpts) == 56          # layout of m3_infer_opts (include/m3b200.h)
    assert engine.InferOpts.row_scales.offset == 16 and engine.InferOpts.wav_header.offset == 48


def test_header_is_plain_c_and_links(built_library, tmp_path):
    """include/m3b200.h must be consumable from C (the reference-side binding is ctypes / cgo-style FFI): compile a
    C99 translation unit against it, link the shared library, call the GPU-free entry points."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "m3b200.h"
_Static_assert(sizeof(m3_infer_opts) == 56, "m3_infer_opts layout");
_Static_assert(sizeof(m3_voice_info) == 64, "m3_voice_info layout");
int main(void) {
  unsigned char h[44];
  m3_voice* v = 0;
  if (m3_wav_header(22050, 1000, h) != M3_OK || memcmp(h, "RIFF", 4) != 0) return 1;
  if (m3_voice_load("/nonexistent/voice", 0, &v) == M3_OK || v != 0) return 2;
  if (strlen(m3_last_error()) == 0) return 3;
  printf("%s|%d\n", m3_version(), (int)m3_device_count());
  return 0;
}
''')
    exe = tmp_path / "abi"
    lib = Path(built_library)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                    str(lib), f"-Wl,-rpath,{lib.parent}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out
    assert "sm_100a" in out.stdout


def test_loader_survives_corrupted_voice_files(built_library):
    """ORT raises on a bad model; the native loader must do the same (error code -> exception), never crash the
    server process (SURVEY.md §8b "Errors").  300 corrupted generator.onnx + 75 corrupted config.json variants."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, str(ROOT / "tests" / "fuzz_loader.py"), "7", "300"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, (out.returncode, out.stderr[-2000:])
    assert "return codes:" in out.stdout
    codes = eval(out.stdout.split("return codes:")[1].strip())
    assert set(codes) <= {engine.M3_ERR_MODEL, engine.M3_ERR_IO, engine.M3_ERR_NOGPU, engine.M3_OK}, codes
    assert codes.get(engine.M3_ERR_MODEL, 0) > 100


def test_loader_is_clean_under_asan(tmp_path):
    """The protobuf / JSON readers and the binder compiled with -fsanitize=address,undefined and fed 120 corrupted
    voices: no out-of-bounds read, no UB -- an error return that happens to survive is not enough for a server."""
    import shutil
    import subprocess
    from mimic3_b200 import synth_voice as sv
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    csrc = ROOT / "mimic3_b200" / "csrc"
    exe = tmp_path / "harness"
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-I", str(csrc), "-I", str(ROOT / "include"), str(ROOT / "tests" / "asan_loader_harness.cc"),
                            str(csrc / "onnx_reader.cc"), str(csrc / "voice_model.cc"), "-o", str(exe)],
                           capture_output=True, text=True)
    if build.returncode != 0 and "asan" in build.stderr.lower():
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    sv.write_voice(tmp_path / "good", sv.tiny_config(n_speakers=2), seed=1)
    good = (tmp_path / "good" / "generator.onnx").read_bytes()
    cfg = (tmp_path / "good" / "config.json").read_bytes()
    rng = np.random.default_rng(11)
    paths = [str(tmp_path / "good")]
    for it in range(120):
        d = tmp_path / f"v{it}"
        d.mkdir()
        b, c = bytearray(good), bytearray(cfg)
        m = it % 5
        if m == 0:
            b = b[: int(rng.integers(0, len(b)))]
        elif m == 1:
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, min(len(b), 4096) if rng.random() < 0.7 else len(b)))] = int(rng.integers(0, 256))
        elif m == 2:
            pos = int(rng.integers(0, len(b)))
            b[pos:pos] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 64)), dtype=np.uint8))
        elif m == 3:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] |= 0x80
        else:
            c = c[: int(rng.integers(0, len(c)))]
        (d / "generator.onnx").write_bytes(bytes(b))
        (d / "config.json").write_bytes(bytes(c))
        paths.append(str(d))
    run = subprocess.run([str(exe)] + paths, capture_output=True, text=True, timeout=600,
                         env={"ASAN_OPTIONS": "detect_leaks=0", "PATH": "/usr/bin:/bin"})
    assert run.returncode == 0, run.stderr[-3000:]
    assert "ERROR: AddressSanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-3000:]
    ok, bad = (int(x) for x in run.stdout.split()[1::2])
    assert ok >= 1 and bad >= 40 and ok + bad == 121
