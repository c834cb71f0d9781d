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
# semantic corruption: fields that are well-formed JSON / protobuf but out of range -- zeros, negatives, huge
# values, NaN -- must come back as error codes too (n_heads = 0 used to be a SIGFPE, deep nesting a stack overflow)
import json
base_cfg = json.loads(cfg_good)
sem = 0
for key in ("n_heads", "n_layers", "hidden_channels", "inter_channels", "filter_channels", "kernel_size",
            "upsample_initial_channel", "num_symbols", "n_speakers", "gin_channels"):
    for val in (0, -1, -2 ** 31, 2 ** 31 - 1, 1e300, float("nan"), 3.7, "x", None, [1], True):
        c = json.loads(cfg_good)
        c["model"][key] = val
        (d / "v" / "config.json").write_text(json.dumps(c))
        h = C.c_void_p()
        rc = lib.m3_voice_load(str(d / "v").encode(), 0, C.byref(h))
        codes[rc] = codes.get(rc, 0) + 1
        sem += 1
        if rc == 0:
            lib.m3_voice_free(h)
for key, vals in (("upsample_rates", ([0, 1], [-4, 4], [], [1e9, 2], [2] * 40, ["a"])),
                  ("upsample_kernel_sizes", ([0, 0], [-3, 4], [10 ** 9, 4])),
                  ("resblock_kernel_sizes", ([0, 0], [2, 4], [-3, 5], [])),
                  ("resblock_dilation_sizes", ([[0], [0]], [[-1, 2], [2, 6]], [[], []], [[10 ** 9], [1]]))):
    for val in vals:
        c = json.loads(cfg_good)
        c["model"][key] = val
        (d / "v" / "config.json").write_text(json.dumps(c))
        h = C.c_void_p()
        rc = lib.m3_voice_load(str(d / "v").encode(), 0, C.byref(h))
        codes[rc] = codes.get(rc, 0) + 1
        sem += 1
        if rc == 0:
            lib.m3_voice_free(h)
for doc in ("[" * 2_000_000, "{\"a\":" * 500_000, '{"model": ' + "[" * 100_000 + "]" * 100_000 + "}"):
    (d / "v" / "config.json").write_text(doc)
    h = C.c_void_p()
    rc = lib.m3_voice_load(str(d / "v").encode(), 0, C.byref(h))
    assert rc != 0
    codes[rc] = codes.get(rc, 0) + 1
    sem += 1
# weight_g / weight_v pairs with an empty leading dimension (division by dims[0] in the fusion)
(d / "v" / "config.json").write_bytes(cfg_good)
from mimic3_b200 import onnx_writer as ow
params = sv.make_params(sv.tiny_config(n_speakers=2), 1)
inits = [ow.tensor_proto(k, v) for k, v in params.items() if not k.startswith("flow.flows.0.enc.in_layers.0.weight")]
inits.append(ow.tensor_proto("flow.flows.0.enc.in_layers.0.weight_g", np.zeros((0,), np.float32)))
inits.append(ow.tensor_proto("flow.flows.0.enc.in_layers.0.weight_v", np.zeros((0, 4, 5), np.float32)))
(d / "v" / "generator.onnx").write_bytes(ow.model_proto([], inits, [], []))
h = C.c_void_p()
rc = lib.m3_voice_load(str(d / "v").encode(), 0, C.byref(h))
assert rc != 0, "a voice whose in_layers weight is empty must not load"
codes[rc] = codes.get(rc, 0) + 1
sem += 1
print("semantic cases:", sem)
print("return codes:", codes)
shutil.rmtree(d)
