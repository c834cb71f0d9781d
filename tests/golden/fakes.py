"""Deterministic stand-ins shared by make_golden_speak_sentence.py (runs the reference) and tests/test_tts_host.py
(runs mimic3_b200.tts): a fake phonemes->ids and a fake int16 'synthesis' that depends on every argument."""
import hashlib

import numpy as np


def fake_ids(phonemes):
    return [3 + (ord(p) - ord("a")) for w in phonemes for p in w] + [2]


def fake_audio(voice_key, ids, speaker, length_scale, noise_scale, noise_w, rate):
    h = hashlib.sha256(repr((voice_key, list(ids), speaker, length_scale, noise_scale, noise_w, rate)).encode()).digest()
    rng = np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))
    return rng.integers(-32768, 32768, size=16 * len(ids)).astype(np.int16)


def sample_rate_of(voice_key):
    return 22050 if voice_key == "v0" else 16000
