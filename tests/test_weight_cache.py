"""Packed-weight cache (include/m3b200.h m3_voice_load_ex / m3_weight_cache_*).

CPU: the GPU-free halves -- conversion, complete validation of a blob, damage / staleness detection, the sha256 the
blob records against hashlib (= mimic3_tts/utils.py file_sha256_sum, what download.py:108-117 compares with the
voices.json entry).  GPU: a voice loaded from its blob is the same voice (bit-identical PCM), the manifest check
raises on a wrong digest, and the load times of both paths are printed."""
import hashlib
import os
import shutil

import numpy as np
import pytest

from mimic3_b200 import engine


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def test_sha256_matches_hashlib(built_library, tmp_path):
    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 1 << 20, (1 << 20) + 17):
        p = tmp_path / f"f{n}"
        p.write_bytes(np.random.default_rng(n).integers(0, 256, size=n, dtype=np.uint8).tobytes())
        assert engine.sha256_file(p) == _sha(p), n
    with pytest.raises(FileNotFoundError):
        engine.sha256_file(tmp_path / "absent")


def test_build_check_and_key(built_library, voices, tmp_path):
    vd = voices("tiny_ms")
    cache = tmp_path / "cache"
    blob = engine.weight_cache_build(vd, cache)
    assert blob.parent == cache and blob.suffix == ".m3w" and blob.stat().st_size > 512
    # the blob records the digest the voice registry lists for generator.onnx
    assert engine.weight_cache_check(blob) == _sha(vd / "generator.onnx")
    # same voice, same switches -> same key; the voice directory and the generator.onnx inside it are the same voice
    assert engine.weight_cache_build(vd / "generator.onnx", cache) == blob
    # another voice -> another blob
    assert engine.weight_cache_build(voices("tiny"), cache) != blob
    # manifest check at conversion time
    with pytest.raises(engine.B200EngineError, match="voice registry"):
        engine.weight_cache_build(vd, cache, expected_sha256="0" * 64)
    with pytest.raises(ValueError):
        engine.weight_cache_build(vd, cache, expected_sha256="xyz")
    engine.weight_cache_build(vd, cache, expected_sha256=_sha(vd / "generator.onnx").upper())


def test_pack_switches_are_part_of_the_key(built_library, voices, tmp_path, monkeypatch):
    vd = voices("tiny_ms")
    a = engine.weight_cache_build(vd, tmp_path)
    monkeypatch.setenv("M3B200_TC_FORMAT", "bf16")
    b = engine.weight_cache_build(vd, tmp_path)
    assert a != b and a.read_bytes() != b.read_bytes()


def test_damaged_blobs_are_rejected_not_trusted(built_library, voices, tmp_path):
    blob = engine.weight_cache_build(voices("tiny_ms"), tmp_path)
    good = blob.read_bytes()
    rng = np.random.default_rng(3)
    bad = tmp_path / "bad.m3w"
    cases = [good[:100], good[:len(good) // 2], good + b"\0" * 64, b"", b"M3B200WC" + b"\xff" * 600]
    for _ in range(40):   # single flipped bytes anywhere: header fields, meta section, slabs
        pos = int(rng.integers(0, len(good)))
        cases.append(good[:pos] + bytes([good[pos] ^ (1 + int(rng.integers(0, 255)))]) + good[pos + 1:])
    accepted = 0
    for data in cases:
        bad.write_bytes(data)
        try:
            engine.weight_cache_check(bad)
            accepted += 1   # only flips inside unused header padding / strings after their NUL can pass
        except engine.B200EngineError:
            pass
    assert accepted <= 12, accepted
    with pytest.raises(FileNotFoundError):
        engine.weight_cache_check(tmp_path / "absent.m3w")


@pytest.mark.gpu
def test_voice_from_blob_is_the_same_voice(built_library, voices, tmp_path):
    from mimic3_b200.engine import B200Session
    for name, nsym in (("low_ms", 50), ("tiny_rb1_dp", 20)):
        vd = tmp_path / name
        shutil.copytree(voices(name), vd)
        cache = tmp_path / f"cache_{name}"
        digest = _sha(vd / "generator.onnx")
        plain = B200Session(str(vd))
        assert not plain.load_stats["from_cache"] and plain.load_stats["cache_file"] == ""
        first = B200Session(str(vd), cache_dir=cache, expected_sha256=digest)       # miss: converts + writes
        assert not first.load_stats["from_cache"] and first.load_stats["cache_written"]
        assert first.load_stats["onnx_sha256"] == digest
        second = B200Session(str(vd), cache_dir=cache, expected_sha256=digest)      # hit
        assert second.load_stats["from_cache"] and second.load_stats["onnx_sha256"] == digest
        assert second.load_stats["hash_ms"] == 0.0                                   # recorded digest, no re-hash
        third = B200Session(str(vd), cache_dir=cache, verify_sha256=True)           # hit + re-hash
        assert third.load_stats["from_cache"] and third.load_stats["hash_ms"] > 0.0
        print(f"{name}: plain load {plain.load_stats['total_ms']:.0f} ms "
              f"(parse {plain.load_stats['parse_ms']:.0f} + pack {plain.load_stats['pack_ms']:.0f} + upload "
              f"{plain.load_stats['upload_ms']:.0f}); first with cache {first.load_stats['total_ms']:.0f} ms; "
              f"from blob {second.load_stats['total_ms']:.0f} ms (read+validate {second.load_stats['cache_read_ms']:.0f}"
              f" + upload {second.load_stats['upload_ms']:.0f})")
        rng = np.random.default_rng(4)
        lens = np.array([31, 12, 20], dtype=np.int64)
        ids = np.zeros((3, 31), dtype=np.int64)
        for b, L in enumerate(lens):
            ids[b, :L] = rng.integers(4, nsym, size=L)
        sid = np.array([0, 1, 1], dtype=np.int64)
        outs = [s.infer(ids, lens, (0.667, 1.0, 0.8), sid, seed=11, keep_float=True) for s in (plain, first, second, third)]
        for o in outs[1:]:
            np.testing.assert_array_equal(o.frames, outs[0].frames)
            np.testing.assert_array_equal(o.audio, outs[0].audio)
            np.testing.assert_array_equal(o.pcm, outs[0].pcm)
        assert second.info.n_params == plain.info.n_params and second.info.n_speakers == plain.info.n_speakers
        # wrong registry digest: refused on the hit path too
        with pytest.raises(engine.B200EngineError, match="voice registry"):
            B200Session(str(vd), cache_dir=cache, expected_sha256="1" * 64)
        # generator.onnx replaced (size / mtime differ): the old blob is not picked up
        data = (vd / "generator.onnx").read_bytes()
        (vd / "generator.onnx").write_bytes(data)
        os.utime(vd / "generator.onnx", ns=(1, 1))
        again = B200Session(str(vd), cache_dir=cache)
        assert not again.load_stats["from_cache"]
        for s in (plain, first, second, third, again):
            s.close()


def test_registry_lookup_matches_reference_voices_json(tmp_path):
    """registry_sha256 reads the reference's voices.json layout (mimic3_tts/_resources.py:35-51)."""
    import json
    from mimic3_b200.voice import registry_sha256
    reg = {"en_US/vctk_low": {"files": {"generator.onnx": {"size_bytes": 5, "sha256_sum": "ab" * 32},
                                          "config.json": {"size_bytes": 2, "sha256_sum": "cd" * 32}}}}
    f = tmp_path / "voices.json"
    f.write_text(json.dumps(reg))
    assert registry_sha256("/x/y/en_US/vctk_low", f) == "ab" * 32
    assert registry_sha256("/x/y/en_US/vctk_low", reg, "config.json") == "cd" * 32
    assert registry_sha256("/x/y/en_US/other", reg) is None
    assert registry_sha256("/x/y/en_US/vctk_low", reg, "missing.bin") is None
