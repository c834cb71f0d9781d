"""The drop-in boundary exercised from the REFERENCE side: the unmodified ``mimic3_tts`` package (installed from
/root/reference into baseline/_ref; third-party imports that do not exist offline replaced as described in
tests/ref_stubs.py) runs its own ``Mimic3Voice.load_from_directory`` / ``ids_to_audio`` (voice.py:154-243, 245-376)
and its own ``Mimic3TextToSpeechSystem`` queue (tts.py:337-551) with libm3b200 where onnxruntime stood.

CPU tests: selection logic of mimic3_b200.plugin, the onnxruntime-shaped shim, result types are opentts_abc's.
GPU tests: int16 audio from the reference's own code path == B200Voice's engine-native path, bit for bit.
"""
import sys

import numpy as np
import pytest

import ref_stubs

ref = ref_stubs.import_reference()
needs_ref = pytest.mark.skipif(ref is None, reason="no copy of the reference package reachable (baseline/_ref)")


@needs_ref
def test_result_types_are_the_references_own():
    """mimic3_http/synthesis.py:60-85 dispatches on isinstance(result, AudioResult / MarkResult): the batched queue
    must hand out opentts_abc's classes, not look-alikes."""
    import opentts_abc
    from mimic3_b200 import tts
    assert tts.AudioResult is opentts_abc.AudioResult
    assert tts.MarkResult is opentts_abc.MarkResult
    assert tts.BaseResult is opentts_abc.BaseResult
    a = tts.AudioResult(sample_rate_hz=22050, sample_width_bytes=2, num_channels=1, audio_bytes=b"\x00\x00")
    assert isinstance(a, opentts_abc.AudioResult) and a.to_wav_bytes()[:4] == b"RIFF"


@needs_ref
def test_plugin_patch_selects_by_provider_and_never_falls_back(tmp_path, monkeypatch):
    import mimic3_tts.voice as rv
    from mimic3_b200 import plugin
    from mimic3_b200.engine import B200EngineError
    calls = []
    monkeypatch.setattr(rv.Mimic3Voice, "_load_model", staticmethod(lambda p, **kw: calls.append((p, kw)) or "ort-session"))
    plugin.patch_reference(rv)
    plugin.patch_reference(rv)   # idempotent
    monkeypatch.delenv("M3B200_PROVIDER", raising=False)
    assert rv.Mimic3Voice._load_model(tmp_path / "generator.onnx", providers=["CPUExecutionProvider"]) == "ort-session"
    assert rv.Mimic3Voice._load_model(tmp_path / "generator.onnx") == "ort-session" and len(calls) == 2
    # selected: the engine is constructed -- and fails LOUDLY here (no voice file / no GPU), never via onnxruntime
    with pytest.raises((B200EngineError, FileNotFoundError, ValueError)):
        rv.Mimic3Voice._load_model(tmp_path / "generator.onnx", providers=[("B200ExecutionProvider", {})])
    monkeypatch.setenv("M3B200_PROVIDER", "1")
    with pytest.raises((B200EngineError, FileNotFoundError, ValueError)):
        rv.Mimic3Voice._load_model(tmp_path / "generator.onnx", providers=["CUDAExecutionProvider"])
    assert len(calls) == 2


def test_onnxruntime_shim_surface():
    from mimic3_b200 import plugin
    m = plugin.install_as_onnxruntime(force=True) if "onnxruntime" not in sys.modules or getattr(
        sys.modules["onnxruntime"], "__m3b200_shim__", False) else None
    if m is None:
        pytest.skip("a real onnxruntime is loaded in this process")
    so = m.SessionOptions()
    so.graph_optimization_level = m.GraphOptimizationLevel.ORT_DISABLE_ALL     # voice.py:395-398
    so.use_deterministic_compute = True                                        # voice.py:401
    assert m.get_available_providers() == ["B200ExecutionProvider"]
    with pytest.raises(TypeError):
        m.InferenceSession(b"model bytes")


@needs_ref
@pytest.mark.gpu
def test_unmodified_mimic3voice_on_the_engine_matches_b200voice(voices, built_library):
    """Reference class, reference method, engine session injected exactly where voice.py:74-86 takes it."""
    import mimic3_tts.voice as rv
    from mimic3_b200 import plugin
    from mimic3_b200.voice import B200Voice
    plugin.patch_reference(rv)
    for name, speakers in (("tiny_ms", ["p201", 2, None]), ("low_ms", ["p239x", 57, "108"]), ("low", [None])):
        vdir = voices(name)
        mine = B200Voice.load_from_directory(vdir)
        theirs = rv.Mimic3Voice.load_from_directory(vdir, providers=["B200ExecutionProvider"])   # unmodified loader
        assert type(theirs).__name__ == "SymbolsVoice" and theirs.onnx_model.get_providers() == ["B200ExecutionProvider"]
        rng = np.random.default_rng(5)
        for spk in speakers:
            ids = rng.integers(4, mine.onnx_model.info.num_symbols, size=int(rng.integers(5, 60))).tolist()
            for kw in (dict(noise_scale=0.0, noise_w=0.0), dict(noise_scale=0.0, noise_w=0.0, length_scale=1.3, rate=0.8)):
                a = theirs.ids_to_audio(ids, speaker=spk, **kw)          # voice.py:154-243, run() at :230
                b = mine.ids_to_audio(ids, speaker=spk, **kw)            # m3_infer, int16 on the device
                assert a.dtype == np.int16 and a.shape == b.shape
                np.testing.assert_array_equal(a, b)
            # same phoneme -> id conversion from both sides (voice.py:126-152)
            words = [["a", "b"], ["c"]]
            assert theirs.phonemes_to_ids(words) == mine.phonemes_to_ids(words)


@needs_ref
@pytest.mark.gpu
def test_unmodified_tts_system_end_to_end_on_the_engine(voices, built_library, tmp_path):
    """Mimic3TextToSpeechSystem (tts.py) untouched: voice lookup, SymbolsVoice.text_to_phonemes, phonemes_to_ids,
    the result queue, add_break, volume -- with every sentence synthesised by libm3b200 -- against the batched
    B200UtteranceQueue fed the same phonemes."""
    import shutil
    import mimic3_tts.voice as rv
    from mimic3_tts.tts import Mimic3Settings, Mimic3TextToSpeechSystem
    from opentts_abc import AudioResult, MarkResult
    from mimic3_b200 import plugin
    from mimic3_b200.tts import B200Settings, B200UtteranceQueue
    from mimic3_b200.voice import B200Voice
    plugin.patch_reference(rv)
    vroot = tmp_path / "voices"
    shutil.copytree(voices("tiny_ms"), vroot / "xx_XX" / "tiny_low")
    import os
    os.environ["M3B200_PROVIDER"] = "1"
    try:
        settings = Mimic3Settings(voice="xx_XX/tiny_low", voices_directories=[vroot], speaker="p201",
                                  noise_scale=0.0, noise_w=0.0, length_scale=1.1, no_download=True)
        tts = Mimic3TextToSpeechSystem(settings)
        tts.begin_utterance()
        tts.speak_text("abc#de")
        tts.add_break(120)
        tts.set_mark("m1")
        tts.volume = 50
        tts.speak_text("fgh")
        results = list(tts.end_utterance())
    finally:
        del os.environ["M3B200_PROVIDER"]
    kinds = [type(r).__name__ for r in results]
    assert kinds.count("AudioResult") == 3 and kinds.count("MarkResult") == 1
    voice = B200Voice.load_from_directory(vroot / "xx_XX" / "tiny_low")
    q = B200UtteranceQueue(B200Settings(voice="xx_XX/tiny_low", speaker="p201", noise_scale=0.0, noise_w=0.0,
                                        length_scale=1.1), get_voice=lambda key: voice)
    q.begin_utterance()
    q.speak_phonemes([list("abc"), list("de")], is_utterance=True)
    q.add_break(120)
    q.set_mark("m1")
    q.settings.volume = 50
    q.speak_phonemes([list("fgh")], is_utterance=True)
    mine = list(q.end_utterance())
    assert [type(r).__name__ for r in mine] == kinds
    for a, b in zip(results, mine):
        if isinstance(a, AudioResult):
            assert a.sample_rate_hz == b.sample_rate_hz and a.audio_bytes == b.audio_bytes
        elif isinstance(a, MarkResult):
            assert a.name == b.name
