"""TEST INFRASTRUCTURE (CPU oracle) -- restatement of the PCM post chain the reference applies after
``ids_to_audio``; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

* :func:`audioop_mul_int16` -- ``audioop.mul(audio_bytes, 2, settings.volume / 100.0)`` of
  ``Mimic3TextToSpeechSystem._speak_sentence_phonemes`` (mimic3_tts/tts.py:540-543).  The arithmetic lives in
  CPython's ``Modules/audioop.c`` (``audioop_mul_impl`` + ``fbound``), not in the reference tree: double
  multiply, clamp (> 32767 -> 32767, < -32767 -> -32768), floor.  Pinned against the real ``audioop`` module
  in tests/test_post_chain.py (Python < 3.13 ships it).
* :func:`break_samples` -- ``add_break`` (tts.py:452-465): ``int((time_ms / 1000.0) * sample_rate)`` samples.
* :func:`wav_bytes` -- ``AudioResult.to_wav_bytes`` (opentts_abc/__init__.py:117-127) through the ``wave`` module.
"""
from __future__ import annotations

import io
import wave

import numpy as np


def audioop_mul_int16(pcm: np.ndarray, factor: float) -> np.ndarray:
    val = pcm.astype(np.float64) * float(factor)
    out = np.where(val > 32767.0, 32767.0, np.where(val < -32767.0, -32768.0, val))
    return np.floor(out).astype(np.int16)


def break_samples(time_ms: float, sample_rate: int) -> int:
    return int((time_ms / 1000.0) * sample_rate)


def wav_bytes(pcm: np.ndarray, sample_rate: int) -> bytes:
    with io.BytesIO() as wav_io:
        wav_file = wave.open(wav_io, "wb")
        with wav_file:
            wav_file.setframerate(sample_rate)
            wav_file.setsampwidth(2)
            wav_file.setnchannels(1)
            wav_file.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())
        return wav_io.getvalue()


def assemble_stream(utterances, sample_rate, volume=None, lead=None, trail=None, wav=False) -> bytes:
    """What a caller of the reference builds on the host: per sentence volume scaling, silence results in
    between, everything concatenated (mimic3_http/synthesis.py / __main__.py), optionally WAV-framed."""
    parts = []
    for b, u in enumerate(utterances):
        if lead is not None and lead[b]:
            parts.append(np.zeros(int(lead[b]), dtype=np.int16))
        parts.append(audioop_mul_int16(u, volume[b]) if volume is not None else u)
        if trail is not None and trail[b]:
            parts.append(np.zeros(int(trail[b]), dtype=np.int16))
    pcm = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int16)
    return wav_bytes(pcm, sample_rate) if wav else pcm.astype("<i2").tobytes()
