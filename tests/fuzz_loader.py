"""Fuzzer for the native voice loader (protobuf reader, JSON reader, binder): corrupted `generator.onnx` / `config.json`
must produce an error code, never a crash.  `m3_voice_load` parses, binds and packs on the host before it touches
the GPU, so this runs on any machine.  Run by tests/test_cabi_and_host.py in a subprocess (a segfault must not take
pytest down):    python tests/fuzz_loader.py <seed> <iterations>"""
import sys, tempfile, shutil, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mimic3_b200 import synth_voice as sv, engine
lib = engine.load_library()
import ctypes as C
d = Path(tempfile.mkdtemp())
sv.write_voice(d / "v", sv.tiny_config(n_speakers=2), seed=1)
good = (d / "v" / "generator.onnx").read_bytes()
cfg_good = (d / "v" / "config.json").read_bytes()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
codes = {}
for it in range(N):
    b = bytearray(good)
    mode = it % 4
    if mode == 0:   # truncate
        b = b[: int(rng.integers(0, len(b)))]
    elif mode == 1: # flip bytes in the first 4 KB (headers / node structure) and anywhere
        for _ in range(int(rng.integers(1, 8))):
            pos = int(rng.integers(0, min(len(b), 4096) if rng.random() < 0.7 else len(b)))
            b[pos] = int(rng.integers(0, 256))
    elif mode == 2: # splice random garbage
        pos = int(rng.integers(0, len(b)))
        b[pos:pos] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 64)), dtype=np.uint8))
    else:           # corrupt varint lengths: set high bits
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(b)))
            b[pos] |= 0x80
    (d / "v" / "generator.onnx").write_bytes(bytes(b))
    h = C.c_void_p()
    rc = lib.m3_voice_load(str(d / "v").encode(), 0, C.byref(h))
    codes[rc] = codes.get(rc, 0) + 1
    if rc == 0:
        lib.m3_voice_free(h)
# config.json corruption
(d / "v" / "generator.onnx").write_bytes(good)
for it in range(max(10, N // 4)):
    b = bytearray(cfg_good)
    if it % 2 == 0:
        b = b[: int(rng.integers(0, len(b)))]
    else:
        for _ in range(3):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(32, 127))
    (d / "v" / "config.json").write_bytes(bytes(b))
    h = C.c_void_p()
    rc = lib.m3_voice_load(str(d / "v").encode(), 0, C.byref(h))
    codes[rc] = codes.get(rc, 0) + 1
    if rc == 0:
        lib.m3_voice_free(h)
print("return codes:", codes)
shutil.rmtree(d)
