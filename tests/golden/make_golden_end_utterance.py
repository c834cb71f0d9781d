"""Golden vectors for the sentence-grouping rules of Mimic3TextToSpeechSystem.end_utterance
(mimic3_tts/tts.py:470-515), produced BY THE REFERENCE ITSELF.

`import mimic3_tts` fails here on third-party packages that are not installed (xdgenvpy, gruut_ipa,
dataclasses_json, phonemes2ids, onnxruntime, ...).  None of them is touched by end_utterance, so this
script stubs them, imports the unmodified reference from /root/reference, replaces only
`_speak_sentence_phonemes` (the synthesis call) by a recorder, and runs the real `end_utterance` over
random queues.  Output: tests/golden/end_utterance_plans.json (committed; /root/reference is not needed
to run the tests).        python tests/golden/make_golden_end_utterance.py
"""
import json
import random
import sys
import types
from copy import deepcopy
from pathlib import Path
from unittest import mock

sys.path.insert(0, "/root/reference")


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Val:
    def __init__(self, v):
        self.value = v


stub("xdgenvpy", XDG=mock.MagicMock())
stub("gruut_ipa", IPA=types.SimpleNamespace(BREAK_MINOR=_Val("|"), BREAK_MAJOR=_Val("‖"), graphemes=lambda s: list(s)))
stub("dataclasses_json", DataClassJsonMixin=type("DataClassJsonMixin", (), {}))
for name in ("phonemes2ids", "onnxruntime", "espeak_phonemizer", "epitran", "gruut", "gruut.const", "gruut.utils",
             "gruut.text_processor", "requests", "tqdm", "tqdm.auto"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            stub(name, __getattr__=lambda attr: mock.MagicMock())

import mimic3_tts.tts as ref  # noqa: E402  (the unmodified reference)
from opentts_abc import AudioResult, MarkResult  # noqa: E402


def random_queue(rng):
    tts = ref.Mimic3TextToSpeechSystem(ref.Mimic3Settings(voice="v0"))
    spoken = []

    def fake_speak(sent_phonemes, settings=None):
        spoken.append(None)
        return ("SPEAK", deepcopy(sent_phonemes), None if settings is None else
                dict(voice=settings.voice, speaker=settings.speaker, length_scale=settings.length_scale,
                     volume=settings.volume, rate=settings.rate))
    tts._speak_sentence_phonemes = fake_speak
    queue = []
    for _ in range(rng.randint(0, 9)):
        kind = rng.random()
        if kind < 0.65:
            if rng.random() < 0.4:      # settings change between sentences (SSML <voice> / <prosody>)
                tts.settings.voice = rng.choice(["v0", "v1"])
                tts.settings.length_scale = rng.choice([None, 0.8, 1.2])
                tts.settings.volume = rng.choice([100.0, 50.0])
            words = [[rng.choice("abcde") for _ in range(rng.randint(1, 3))] for _ in range(rng.randint(0, 3))]
            is_utt = rng.random() < 0.7
            tts._results.append(ref.Mimic3Phonemes(current_settings=deepcopy(tts.settings), phonemes=words, is_utterance=is_utt))
            s = tts.settings
            queue.append(dict(kind="phonemes", phonemes=words, is_utterance=is_utt,
                              settings=dict(voice=s.voice, speaker=s.speaker, length_scale=s.length_scale,
                                            volume=s.volume, rate=s.rate)))
        elif kind < 0.85:
            ms = rng.choice([0, 10, 250])
            tts.add_break(ms)
            queue.append(dict(kind="break", ms=ms))
        else:
            tts.set_mark("m%d" % len(queue))
            queue.append(dict(kind="mark", name="m%d" % len(queue)))
    out = []
    for r in tts.end_utterance():
        if isinstance(r, tuple):
            out.append(dict(kind="speak", phonemes=r[1], settings=r[2]))
        elif isinstance(r, AudioResult):
            out.append(dict(kind="break", n_bytes=len(r.audio_bytes), sample_rate=r.sample_rate_hz))
        elif isinstance(r, MarkResult):
            out.append(dict(kind="mark", name=r.name))
    assert not tts._results
    return dict(queue=queue, yielded=out)


if __name__ == "__main__":
    rng = random.Random(20260923)
    cases = [random_queue(rng) for _ in range(120)]
    path = Path(__file__).resolve().parent / "end_utterance_plans.json"
    path.write_text(json.dumps(dict(source="mimic3_tts/tts.py:470-515 executed with stubbed third-party imports",
                                    cases=cases), separators=(",", ":")))
    n = sum(1 for c in cases for y in c["yielded"] if y["kind"] == "speak")
    print(f"{len(cases)} queues, {n} spoken sentences -> {path} ({path.stat().st_size} bytes)")
