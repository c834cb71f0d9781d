"""Golden vectors for Mimic3TextToSpeechSystem.end_utterance + _speak_sentence_phonemes (mimic3_tts/tts.py:470-551)
THROUGH to the bytes of every AudioResult, produced BY THE REFERENCE ITSELF.

Same technique as make_golden_end_utterance.py (unmodified reference imported with its missing third-party
packages stubbed), but here `_speak_sentence_phonemes` runs for real: only `_get_or_load_voice` is replaced, by fake
voices whose `phonemes_to_ids` / `ids_to_audio` are deterministic functions that record their arguments.  The golden
holds, per yielded result: which voice was asked, with which ids and which keyword arguments (speaker, length_scale,
noise_scale, noise_w, rate), and the sha256 of `audio_bytes` after the reference's own volume step
(`audioop.mul`, tts.py:540-543).  Output: tests/golden/speak_sentence_results.json
"""
import hashlib
import json
import random
import sys
from copy import deepcopy
from pathlib import Path
from types import SimpleNamespace

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden_end_utterance as base  # noqa: E402  (stubs + imports the reference)

ref = base.ref
AudioResult, MarkResult = base.AudioResult, base.MarkResult


from fakes import fake_audio, fake_ids, sample_rate_of  # noqa: E402


class FakeVoice:
    def __init__(self, key, log):
        self.key, self.log = key, log
        self.config = SimpleNamespace(audio=SimpleNamespace(sample_rate=sample_rate_of(key)))

    def phonemes_to_ids(self, phonemes):
        return fake_ids(phonemes)

    def ids_to_audio(self, ids, speaker=None, length_scale=None, noise_scale=None, noise_w=None, rate=1.0):
        self.log.append(dict(voice=self.key, ids=list(ids), speaker=speaker, length_scale=length_scale,
                             noise_scale=noise_scale, noise_w=noise_w, rate=rate))
        return fake_audio(self.key, ids, speaker, length_scale, noise_scale, noise_w, rate)


def one_case(rng):
    tts = ref.Mimic3TextToSpeechSystem(ref.Mimic3Settings(voice="v0"))
    log = []
    voices = {}
    tts._get_or_load_voice = lambda key: voices.setdefault(key, FakeVoice(key, log))
    queue = []
    for _ in range(rng.randint(1, 8)):
        kind = rng.random()
        if kind < 0.7:
            if rng.random() < 0.5:
                s = tts.settings
                s.voice = rng.choice(["v0", "v1"])
                s.speaker = rng.choice([None, "p1", 2])
                s.length_scale = rng.choice([None, 0.8, 1.2])
                s.noise_scale = rng.choice([None, 0.0, 0.5])
                s.noise_w = rng.choice([None, 0.0, 0.9])
                s.volume = rng.choice([100.0, 50.0, 150.0, 0.0])
                s.rate = rng.choice([1.0, 1.5, 0.5])
            words = [[rng.choice("abcde") for _ in range(rng.randint(1, 3))] for _ in range(rng.randint(1, 3))]
            is_utt = rng.random() < 0.7
            tts._results.append(ref.Mimic3Phonemes(current_settings=deepcopy(tts.settings), phonemes=words, is_utterance=is_utt))
            s = tts.settings
            queue.append(dict(kind="phonemes", phonemes=words, is_utterance=is_utt,
                              settings=dict(voice=s.voice, speaker=s.speaker, length_scale=s.length_scale,
                                            noise_scale=s.noise_scale, noise_w=s.noise_w, volume=s.volume, rate=s.rate)))
        elif kind < 0.88:
            ms = rng.choice([0, 10, 250])
            tts.add_break(ms)
            queue.append(dict(kind="break", ms=ms))
        else:
            tts.set_mark("m%d" % len(queue))
            queue.append(dict(kind="mark", name="m%d" % len(queue)))
    s = tts.settings
    end_settings = dict(voice=s.voice, speaker=s.speaker, length_scale=s.length_scale, noise_scale=s.noise_scale,
                        noise_w=s.noise_w, volume=s.volume, rate=s.rate)
    out = []
    for r in tts.end_utterance():
        if isinstance(r, AudioResult):
            out.append(dict(kind="audio", sample_rate=r.sample_rate_hz, n_bytes=len(r.audio_bytes),
                            sha256=hashlib.sha256(r.audio_bytes).hexdigest()))
        elif isinstance(r, MarkResult):
            out.append(dict(kind="mark", name=r.name))
    return dict(queue=queue, end_settings=end_settings, calls=[dict(voice=c["voice"]) for c in log], yielded=out)


if __name__ == "__main__":
    rng = random.Random(77)
    cases = [one_case(rng) for _ in range(80)]
    path = Path(__file__).resolve().parent / "speak_sentence_results.json"
    path.write_text(json.dumps(dict(source="mimic3_tts/tts.py:470-551 executed with stubbed third-party imports and fake voices",
                                    cases=cases), separators=(",", ":")))
    print(f"{len(cases)} queues, {sum(len(c['calls']) for c in cases)} synthesis calls -> {path} ({path.stat().st_size} bytes)")
