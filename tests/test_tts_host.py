"""CPU tests of the callers' side: sentence grouping of end_utterance against
goldens produced by the reference's own code, the batched queue over a recording fake engine, phonemes->ids."""
import json
from pathlib import Path

import numpy as np
import pytest

from mimic3_b200 import tts
from mimic3_b200.phonemes import load_phoneme_ids, load_phoneme_map, phonemes2ids
from mimic3_b200.voice import B200Voice, VoiceConfig

GOLDEN = Path(__file__).resolve().parent / "golden" / "end_utterance_plans.json"


def _settings(d):
    return tts.B200Settings(voice=d["voice"], speaker=d["speaker"], length_scale=d["length_scale"],
                            volume=d["volume"], rate=d["rate"])


def test_plan_sentences_matches_reference_end_utterance():
    """tests/golden/make_golden_end_utterance.py ran the unmodified mimic3_tts/tts.py:470-515 over 120 random
    queues; plan_sentences must yield the same items, sentences and settings objects."""
    cases = json.loads(GOLDEN.read_text())["cases"]
    assert len(cases) >= 100
    spoken = 0
    for case in cases:
        queue = []
        for q in case["queue"]:
            if q["kind"] == "phonemes":
                queue.append(tts.B200Phonemes(current_settings=_settings(q["settings"]), phonemes=q["phonemes"],
                                              is_utterance=q["is_utterance"]))
            elif q["kind"] == "break":
                queue.append(tts.AudioResult(sample_rate_hz=22050, sample_width_bytes=2, num_channels=1,
                                                 audio_bytes=bytes(int(q["ms"] / 1000.0 * 22050) * 2)))
            else:
                queue.append(tts.MarkResult(name=q["name"]))
        plan = tts.plan_sentences(queue)
        assert len(plan) == len(case["yielded"])
        for got, want in zip(plan, case["yielded"]):
            if want["kind"] == "speak":
                spoken += 1
                assert isinstance(got, tts._Sentence) and got.phonemes == want["phonemes"]
                assert (got.settings is None) == (want["settings"] is None)
                if got.settings is not None:
                    assert got.settings == _settings(want["settings"])
            elif want["kind"] == "break":
                assert isinstance(got, tts.AudioResult) and len(got.audio_bytes) == want["n_bytes"]
            else:
                assert isinstance(got, tts.MarkResult) and got.name == want["name"]
    assert spoken > 200


class FakeSession:
    class info:
        has_speaker_embedding = 1
        hop_length = 4

    def __init__(self):
        self.calls = []

    def infer(self, text, lengths, scales, sid, seed=0, **kw):
        self.calls.append(dict(text=text.copy(), lengths=lengths.copy(), scales=scales, sid=None if sid is None else sid.copy(),
                               seed=seed, **kw))
        fake = self

        class R:
            total_samples = int(lengths.sum()) * 4

            def utterance_pcm(self_, b):  # length and content identify the row
                return np.full(int(lengths[b]) * 4, int(text[b, 0]), dtype=np.int16)

            def stream_bytes(self_):
                return b"STREAM%d" % len(fake.calls)
        return R()


def _voice(sample_rate=22050):
    cfg = VoiceConfig({"model": {"n_speakers": 3}, "audio": {"sample_rate": sample_rate},
                       "inference": {"length_scale": 1.0, "noise_scale": 0.667, "noise_w": 0.8},
                       "phonemes": {"blank": "#", "blank_between": "words", "bos": "^", "eos": "$", "auto_bos_eos": True}})
    p2i = {"_": 0, "^": 1, "$": 2, "#": 3, "a": 4, "b": 5, "c": 6, ",": 7, ".": 8}
    return B200Voice(cfg, FakeSession(), p2i, None, {"spk": 2})


def test_queue_batches_one_call_per_voice_and_keeps_order():
    voices = {"en/a": _voice(), "de/b": _voice(16000)}
    q = tts.B200UtteranceQueue(tts.B200Settings(voice="en/a"), voices.__getitem__)
    q.speak_phonemes([["a"], ["b"]])                      # sentence 0 (settings: queue default, None -> self.settings)
    q.settings.length_scale, q.settings.volume = 1.5, 50.0
    q.add_break(100)
    q.speak_phonemes([["b", "c"]])                        # sentence 1: spoken with the PREVIOUS item's settings
    q.voice = "de/b#spk"
    q.set_mark("here")
    q.speak_phonemes([["c"]], is_utterance=False)
    q.speak_phonemes([["a"]])                             # joins the open sentence; settings changed -> see reference rule
    out = list(q.end_utterance())
    kinds = [type(r).__name__ for r in out]
    assert kinds == ["AudioResult", "AudioResult", "AudioResult", "MarkResult", "AudioResult"]
    assert len(out[1].audio_bytes) == 2 * 2205                                # the 100 ms break, untouched
    a, b = voices["en/a"].onnx_model.calls, voices["de/b"].onnx_model.calls
    assert len(a) == 1 and len(b) == 1                                        # one engine call per voice
    # Reference rules (tts.py:470-515, 525): sentence 0 is spoken with settings=None -> the system's settings AT
    # end_utterance time (voice de/b#spk, length 1.5, volume 50); sentence 1 with the settings of the item queued
    # before it (en/a defaults); the last sentence with those of the non-utterance part before it (de/b).
    ca, cb = a[0], b[0]
    assert ca["text"].shape[0] == 1 and ca["scales"] is None
    ids1 = voices["en/a"].phonemes_to_ids([["b", "c"]])
    assert ca["text"][0, : len(ids1)].tolist() == ids1 and ca["lengths"][0] == len(ids1)
    np.testing.assert_allclose(ca["row_scales"], [[0.667, 1.0, 0.8]], rtol=1e-6)
    assert ca["volume"] is None and ca["sid"].tolist() == [0]
    assert cb["text"].shape[0] == 2                                           # both de/b sentences in ONE call
    ids0, ids2 = voices["de/b"].phonemes_to_ids([["a"], ["b"]]), voices["de/b"].phonemes_to_ids([["c"], ["a"]])
    assert cb["text"][0, : len(ids0)].tolist() == ids0 and cb["text"][1, : len(ids2)].tolist() == ids2
    np.testing.assert_allclose(cb["row_scales"], [[0.667, 1.5, 0.8]] * 2, rtol=1e-6)
    np.testing.assert_allclose(cb["volume"], [0.5, 0.5])
    assert cb["sid"].tolist() == [2, 2]
    assert [r.sample_rate_hz for r in out if isinstance(r, tts.AudioResult)] == [16000, 22050, 22050, 16000]
    assert len(out[0].audio_bytes) == 2 * 4 * len(ids0) and len(out[4].audio_bytes) == 2 * 4 * len(ids2)
    assert not q._results
    # byte-for-byte what results_to_wav_bytes of the reference's HTTP path would assemble
    wav = tts.results_to_wav_bytes(out)
    assert wav[:4] == b"RIFF" and len(wav) == 44 + sum(len(r.audio_bytes) for r in out if isinstance(r, tts.AudioResult))


def test_end_utterance_wav_folds_breaks_into_silences():
    v = _voice()
    q = tts.B200UtteranceQueue(tts.B200Settings(voice="en/a", noise_scale=0.0, noise_w=0.0), lambda key: v)
    q.add_break(10)
    q.speak_phonemes([["a"]])
    q.add_break(20)
    q.set_mark("m")
    q.add_break(30)
    q.speak_phonemes([["b"]])
    assert q.end_utterance_wav() == b"STREAM1"
    c = v.onnx_model.calls[0]
    assert c["wav_header"] is True and c["seed"] == 0
    assert c["lead_silence"].tolist() == [220, 0] and c["trail_silence"].tolist() == [441 + 661, 0]
    assert c["sid"].tolist() == [0, 0] and c["volume"] is None
    # two voices -> host assembly like the reference
    v2 = _voice()
    q = tts.B200UtteranceQueue(tts.B200Settings(voice="x"), {"x": v, "y": v2}.__getitem__)
    q.speak_phonemes([["a"]])
    q.voice = "y"
    q.speak_phonemes([["b"]])
    q.speak_phonemes([["c"]])
    wav = q.end_utterance_wav()
    assert wav[:4] == b"RIFF" and wav[8:12] == b"WAVE"


def test_phonemes2ids_documented_behaviour():
    p2i = {"_": 0, "^": 1, "$": 2, "#": 3, "a": 4, "b": 5, "ˈ": 6, ",": 7, ".": 8, "1": 9, "ma": 10}
    words = [["a", "b"], ["b"]]
    assert phonemes2ids(words, p2i) == [4, 5, 5]
    assert phonemes2ids(words, p2i, blank="#") == [3, 4, 5, 3, 5, 3]                                  # words
    assert phonemes2ids(words, p2i, blank="#", blank_between="tokens") == [3, 4, 3, 5, 3, 5, 3]
    assert phonemes2ids(words, p2i, blank="#", blank_at_start=False, blank_at_end=False) == [4, 5, 3, 5]
    assert phonemes2ids(words, p2i, blank="#", bos="^", eos="$", auto_bos_eos=True) == [3, 1, 3, 4, 5, 3, 5, 3, 2, 3]
    assert phonemes2ids([["a", "?"], [";"]], p2i, simple_punctuation=True) == [4, 8, 7]
    assert phonemes2ids([["a", "?"]], p2i, simple_punctuation=False) == [4]                             # unknown dropped
    assert phonemes2ids([["ˈa"]], p2i, separate={"ˈ"}) == [6, 4]
    assert phonemes2ids([["ma1"]], p2i, separate_tones=True) == [10, 9]
    assert phonemes2ids([["ma1"]], p2i, separate_tones=True, tone_before=True) == [9, 10]
    assert phonemes2ids([["x"]], p2i, phoneme_map={"x": ["a", "b"]}) == [4, 5]
    assert phonemes2ids([], p2i, blank="#") == [] and phonemes2ids([["zz"]], p2i, blank="#") == []
    with pytest.raises(KeyError):
        phonemes2ids([["zz"]], p2i, fail_on_missing=True)
    assert load_phoneme_ids(["0 _\n", "# comment\n", "\n", "3 #\n", "4  \n"]) == {"_": 0, "#": 3, " ": 4}
    assert load_phoneme_map(["x a b\n", "\n", "y c\n"]) == {"x": ["a", "b"], "y": ["c"]}


def test_voice_phonemes_to_ids_uses_config_and_injected_function():
    v = _voice()
    assert v.phonemes_to_ids([["a"], ["b", "c"]]) == [3, 1, 3, 4, 3, 5, 6, 3, 2, 3]
    seen = {}
    v.phonemes_to_ids_fn = lambda **kw: seen.update(kw) or [42]
    assert v.phonemes_to_ids([["a"]]) == [42]
    assert seen["blank"] == "#" and seen["auto_bos_eos"] is True and seen["fail_on_missing"] is False  # voice.py:133-152


def test_ids_to_audio_inputs_match_reference_goldens():
    """tests/golden/make_golden_ids_to_audio.py called the unmodified Mimic3Voice.ids_to_audio (voice.py:154-243) with a
    recording fake in place of the ORT session over 192 argument combinations; B200Voice must hand the engine the
    same `input`, `input_lengths`, `scales` and `sid` (names map to m3_infer's arguments, voice.py:180-218)."""
    g = json.loads((GOLDEN.parent / "ids_to_audio_inputs.json").read_text())
    assert len(g["cases"]) >= 150
    for c in g["cases"]:
        cfg = VoiceConfig({"model": {"n_speakers": 3 if c["multispeaker"] else 1}, "inference": g["defaults"]})
        assert cfg.is_multispeaker == c["multispeaker"]
        v = B200Voice(cfg, FakeSession(), {}, None, c["speaker_map"])
        v.ids_to_audio(c["ids"], speaker=c["speaker"], length_scale=c["length_scale"], noise_scale=c["noise_scale"],
                       noise_w=c["noise_w"], rate=c["rate"])
        call, want = v.onnx_model.calls[-1], c["inputs"]
        assert call["text"].dtype == np.int64 and list(call["text"].shape) == want["input"]["shape"]
        assert call["text"].reshape(-1).tolist() == want["input"]["values"]
        assert call["lengths"].dtype == np.int64 and call["lengths"].tolist() == want["input_lengths"]["values"]
        assert call["scales"].dtype == np.float32 and want["scales"]["dtype"] == "float32"
        np.testing.assert_array_equal(call["scales"], np.array(want["scales"]["values"], dtype=np.float32))
        if "sid" in want:
            assert call["sid"].dtype == np.int64 and call["sid"].tolist() == want["sid"]["values"]
        else:
            assert call["sid"] is None


def test_queue_results_match_reference_speak_sentence_goldens():
    """tests/golden/make_golden_speak_sentence.py ran the reference's end_utterance AND _speak_sentence_phonemes
    (tts.py:470-551) over 80 random queues with fake voices whose output depends on every argument they receive.
    The batched queue, given the same fakes behind `ids_to_audio_rows`, must yield byte-identical AudioResults:
    same voice, ids, speaker, scales, rate, and the reference's audioop.mul volume step."""
    import hashlib
    import sys
    sys.path.insert(0, str(GOLDEN.parent))
    from fakes import fake_audio, fake_ids, sample_rate_of
    from oracle.post_chain import audioop_mul_int16

    class FakeBatchVoice:
        def __init__(self, key, log):
            from types import SimpleNamespace
            self.key, self.log = key, log
            self.config = SimpleNamespace(audio=SimpleNamespace(sample_rate=sample_rate_of(key)))

        def phonemes_to_ids(self, phonemes):
            return fake_ids(phonemes)

        def ids_to_audio_rows(self, batch_ids, speakers=None, length_scales=None, noise_scales=None, noise_ws=None,
                              rates=None, volumes=None, seed=None):
            self.log.append(len(batch_ids))
            out = []
            for i, ids in enumerate(batch_ids):
                a = fake_audio(self.key, ids, speakers[i], length_scales[i], noise_scales[i], noise_ws[i], rates[i])
                out.append(a if volumes[i] == 100.0 else audioop_mul_int16(a, volumes[i] / 100.0))  # device post chain
            return out

    cases = json.loads((GOLDEN.parent / "speak_sentence_results.json").read_text())["cases"]
    n_audio = n_calls = 0
    for case in cases:
        log, voices = [], {}
        q = tts.B200UtteranceQueue(tts.B200Settings(voice="v0"), lambda key: voices.setdefault(key, FakeBatchVoice(key, log)))
        for item in case["queue"]:
            if item["kind"] == "phonemes":
                s = item["settings"]
                q._results.append(tts.B200Phonemes(
                    current_settings=tts.B200Settings(voice=s["voice"], speaker=s["speaker"], length_scale=s["length_scale"],
                                                      noise_scale=s["noise_scale"], noise_w=s["noise_w"], volume=s["volume"],
                                                      rate=s["rate"]),
                    phonemes=item["phonemes"], is_utterance=item["is_utterance"]))
            elif item["kind"] == "break":
                q.add_break(item["ms"])
            else:
                q.set_mark(item["name"])
        e = case["end_settings"]
        q.settings = tts.B200Settings(voice=e["voice"], speaker=e["speaker"], length_scale=e["length_scale"],
                                      noise_scale=e["noise_scale"], noise_w=e["noise_w"], volume=e["volume"], rate=e["rate"])
        chunk = (None, 1, 2, 3)[cases.index(case) % 4]   # the streaming mode must yield the very same results
        got = list(q.end_utterance(max_batch_sentences=chunk))
        assert len(got) == len(case["yielded"])
        for g, want in zip(got, case["yielded"]):
            if want["kind"] == "audio":
                n_audio += 1
                assert isinstance(g, tts.AudioResult) and g.sample_rate_hz == want["sample_rate"]
                assert len(g.audio_bytes) == want["n_bytes"]
                assert hashlib.sha256(g.audio_bytes).hexdigest() == want["sha256"]
            else:
                assert isinstance(g, tts.MarkResult) and g.name == want["name"]
        assert sum(log) == len(case["calls"])                    # same sentences synthesised ...
        if chunk is None:
            assert len(log) == len({c["voice"] for c in case["calls"]})   # ... in ONE engine call per voice
            n_calls += len(log)
        else:
            assert max(log, default=0) <= chunk and len(log) <= len(case["calls"])
    assert n_audio > 250 and n_calls < 80


def test_streaming_end_utterance_is_lazy_and_reads_settings_per_group():
    """max_batch_sentences=1 is the reference's generator (tts.py:470-515): a sentence is synthesised when the consumer asks
    for it, with the settings in force at that moment for sentences that carry none of their own."""
    calls = []

    class V:
        def __init__(self, key):
            from types import SimpleNamespace
            self.key = key
            self.config = SimpleNamespace(audio=SimpleNamespace(sample_rate=22050))

        def phonemes_to_ids(self, phonemes):
            return [len(w) for w in phonemes]

        def ids_to_audio_rows(self, batch_ids, speakers=None, length_scales=None, **kw):
            calls.append((self.key, len(batch_ids), list(length_scales)))
            return [np.full(4, len(ids), dtype=np.int16) for ids in batch_ids]

    q = tts.B200UtteranceQueue(tts.B200Settings(voice="a"), lambda key: V(key))
    for n in (1, 2, 3):
        q._results.append(tts.B200Phonemes(current_settings=None, phonemes=[["x"] * n], is_utterance=True))
        q.set_mark(f"m{n}")
    gen = q.end_utterance(max_batch_sentences=1)
    assert calls == []                                  # nothing runs until it is iterated
    first = next(gen)
    assert isinstance(first, tts.AudioResult) and len(calls) == 1 and calls[0][1] == 1
    q.settings.length_scale = 2.5                       # changed while the utterance is being spoken
    rest = list(gen)
    assert [type(r).__name__ for r in [first] + rest] == ["AudioResult", "MarkResult"] * 3
    assert len(calls) == 3 and calls[1][2] == [2.5] and calls[2][2] == [2.5]
    assert q._results == []

