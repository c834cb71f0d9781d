"""Batched ``end_utterance``: the callers' side of the ids->audio path.

``Mimic3TextToSpeechSystem.end_utterance`` (``mimic3_tts/tts.py:470-515``) walks its queue of pending
results and speaks ONE sentence per ``InferenceSession.run`` (``_speak_sentence_phonemes``,
``tts.py:519-551``).  :class:`B200UtteranceQueue` keeps the queue, the result types and the sentence
grouping rules of the reference -- including which settings object a sentence is spoken with -- but
sends all sentences of an utterance that share a voice to the engine in one call
(``B200Voice.ids_to_audio_rows``: per-sentence speaker / scales / rate / volume through ``m3_infer_ex``),
then yields the results in the reference's order.  :func:`results_to_wav_bytes` is the WAV assembly of
``mimic3_http/synthesis.py:60-85``; :meth:`B200UtteranceQueue.end_utterance_wav` produces the same bytes with
the silences and the RIFF header written by the engine (PCM post chain, §8f rank 2).

Out of scope, as everywhere in this repo: turning text into phonemes (``speak_text`` / ``speak_tokens``,
``tts.py:337-449`` -- gruut / espeak-ng / epitran).  Callers queue phonemes (or ids) directly.
"""
from __future__ import annotations

import io
import typing
import wave
from copy import deepcopy
from dataclasses import dataclass, field

import numpy as np

from .voice import DEFAULT_RATE, DEFAULT_VOLUME, B200Voice

PHONEMES_LIST_TYPE = typing.List[typing.List[str]]
DEFAULT_VOICE = "en_UK/apope_low"   # mimic3_tts/const.py:20


try:
    # The plugin surface's own result types (``opentts_abc/__init__.py:84-141``, stdlib only; installed with Mimic 3):
    # the HTTP server dispatches on ``isinstance(result, AudioResult)`` (``mimic3_http/synthesis.py:60-85``), so the
    # queue must hand out THESE classes wherever the reference is installed.
    from opentts_abc import AudioResult, BaseResult, MarkResult
except ImportError:  # stand-alone use without Mimic 3 on the path: structurally identical local definitions

    @dataclass
    class BaseResult:
        """opentts_abc.BaseResult (``opentts_abc/__init__.py:84-93``)."""

        tag: typing.Optional[typing.Any] = None

    @dataclass
    class AudioResult(BaseResult):
        """opentts_abc.AudioResult (``opentts_abc/__init__.py:96-127``): raw 16-bit mono audio of one sentence or break."""

        sample_rate_hz: int = 22050
        sample_width_bytes: int = 2
        num_channels: int = 1
        audio_bytes: bytes = b""

        def to_wav_bytes(self) -> bytes:
            with io.BytesIO() as wav_io:
                wav_file: wave.Wave_write = wave.open(wav_io, "wb")
                with wav_file:
                    wav_file.setframerate(self.sample_rate_hz)
                    wav_file.setsampwidth(self.sample_width_bytes)
                    wav_file.setnchannels(self.num_channels)
                    wav_file.writeframes(self.audio_bytes)
                return wav_io.getvalue()

    @dataclass
    class MarkResult(BaseResult):
        """opentts_abc.MarkResult (``opentts_abc/__init__.py:130-141``): an SSML <mark> was reached."""

        name: str = ""


@dataclass
class B200Settings:
    """The fields of ``Mimic3Settings`` (``tts.py:63-128``) that reach the ids->audio path."""

    voice: typing.Optional[str] = None
    speaker: typing.Optional[typing.Union[str, int]] = None
    length_scale: typing.Optional[float] = None
    noise_scale: typing.Optional[float] = None
    noise_w: typing.Optional[float] = None
    sample_rate: int = 22050          # of add_break() silence
    volume: float = DEFAULT_VOLUME    # [0, 100]
    rate: float = DEFAULT_RATE


@dataclass
class B200Phonemes:
    """``Mimic3Phonemes`` (``tts.py:131-142``): pending sentence (part) with the settings current when queued."""

    current_settings: B200Settings
    phonemes: PHONEMES_LIST_TYPE = field(default_factory=list)
    is_utterance: bool = True


@dataclass
class _Sentence:
    phonemes: PHONEMES_LIST_TYPE
    settings: typing.Optional[B200Settings]


def plan_sentences(results: typing.Iterable[typing.Any], phonemes_type=B200Phonemes) -> typing.List[typing.Any]:
    """The grouping loop of ``end_utterance`` (``tts.py:470-515``) without the synthesis: returns the
    sequence the reference would yield, with a :class:`_Sentence` wherever it would call
    ``_speak_sentence_phonemes(sent_phonemes, settings=last_settings)``.  Faithful to the reference, a
    sentence is spoken with the settings of the PREVIOUS queued item (``last_settings`` is only updated
    after the item is handled), ``None`` meaning the system's current settings."""
    plan: typing.List[typing.Any] = []
    last_settings = None
    sent: PHONEMES_LIST_TYPE = []
    for result in results:
        if isinstance(result, phonemes_type):
            if result.is_utterance:
                if sent and (last_settings is not None) and (result.current_settings != last_settings):
                    plan.append(_Sentence(list(sent), last_settings))
                    sent.clear()
                sent.extend(result.phonemes)
                if sent:
                    plan.append(_Sentence(list(sent), last_settings))
                    sent.clear()
            else:
                sent.extend(result.phonemes)
            last_settings = result.current_settings
        else:
            if sent:
                plan.append(_Sentence(list(sent), last_settings))
                sent.clear()
            plan.append(result)
    if sent:
        plan.append(_Sentence(list(sent), last_settings))
    return plan


def results_to_wav_bytes(results: typing.Iterable[typing.Any]) -> bytes:
    """``mimic3_http/synthesis.py:60-85``: all AudioResults of a request in one WAV, parameters of the first."""
    with io.BytesIO() as wav_io:
        wav_file: wave.Wave_write = wave.open(wav_io, "wb")
        params_set = False
        with wav_file:
            for result in results:
                if isinstance(result, AudioResult):
                    if not params_set:
                        wav_file.setframerate(result.sample_rate_hz)
                        wav_file.setsampwidth(result.sample_width_bytes)
                        wav_file.setnchannels(result.num_channels)
                        params_set = True
                    wav_file.writeframes(result.audio_bytes)
            if not params_set:
                wav_file.setframerate(22050)
                wav_file.setsampwidth(2)
                wav_file.setnchannels(1)
        return wav_io.getvalue()


class B200UtteranceQueue:
    """The queue half of ``Mimic3TextToSpeechSystem`` (``tts.py:144-152, 452-515``) over :class:`B200Voice`."""

    def __init__(self, settings: B200Settings, get_voice: typing.Callable[[str], B200Voice]):
        self.settings = settings
        self._get_voice = get_voice
        self._results: typing.List[typing.Any] = []

    # ---- the voice property of the reference (tts.py:314-330) -------------------------------------------
    @property
    def voice(self) -> str:
        return self.settings.voice or DEFAULT_VOICE

    @voice.setter
    def voice(self, new_voice: str):
        if new_voice != self.settings.voice:
            self.settings.speaker = None  # clear speaker on voice change
        self.settings.voice = new_voice or DEFAULT_VOICE
        if "#" in self.settings.voice:  # <voice>#<speaker>
            voice, speaker = self.settings.voice.split("#", maxsplit=1)
            self.settings.voice, self.settings.speaker = voice, speaker

    # ---- queueing (tts.py:364-468) ---------------------------------------------------------------------
    def begin_utterance(self):
        pass  # tts.py:364-365

    def speak_phonemes(self, phonemes: PHONEMES_LIST_TYPE, is_utterance: bool = True):
        """What ``speak_text`` / ``speak_tokens`` append after phonemisation (``tts.py:397-403, 444-450``)."""
        if phonemes:
            self._results.append(B200Phonemes(current_settings=deepcopy(self.settings), phonemes=[list(w) for w in phonemes],
                                              is_utterance=is_utterance))

    def add_break(self, time_ms: int):
        num_samples = int((time_ms / 1000.0) * self.settings.sample_rate)  # tts.py:454
        self._results.append(AudioResult(sample_rate_hz=self.settings.sample_rate, audio_bytes=bytes(num_samples * 2),
                                         sample_width_bytes=2, num_channels=1))

    def set_mark(self, name: str):
        self._results.append(MarkResult(name=name))

    # ---- synthesis ---------------------------------------------------------------------------------------
    def _rows(self, plan):
        """(voice key, voice, ids, settings) of every sentence of the plan, in plan order."""
        rows = []
        for item in plan:
            if isinstance(item, _Sentence):
                settings = item.settings or self.settings           # tts.py:525
                key = settings.voice or self.voice                   # tts.py:526
                voice = self._get_voice(key)
                rows.append((key, voice, voice.phonemes_to_ids(item.phonemes), settings))
            else:
                rows.append(None)
        return rows

    def end_utterance(self, max_batch_sentences: typing.Optional[int] = None) -> typing.Iterable[typing.Any]:
        """Same results in the same order as ``tts.py:470-515``; one engine call per voice instead of one per
        sentence.  (A generator like the reference's: nothing runs until it is iterated.)

        ``max_batch_sentences`` bounds how many sentences are synthesised before anything is yielded: the plan is cut
        into groups of at most that many sentences (marks and breaks stay with the sentence they follow), each
        group is one engine call per voice, and its results are yielded before the next group runs.  ``None`` (default)
        = the whole utterance in one batch (highest throughput); ``1`` = the reference's behaviour, one sentence per
        yield with ``self.settings`` / ``self.voice`` read when that sentence is spoken (lowest first-audio latency)."""
        if max_batch_sentences is not None and max_batch_sentences < 1:
            raise ValueError("max_batch_sentences must be >= 1")
        plan = plan_sentences(self._results)
        start = 0
        while start < len(plan) or start == 0:
            stop, n = start, 0
            while stop < len(plan):          # extend the group up to the sentence budget (+ trailing marks / breaks)
                if isinstance(plan[stop], _Sentence):
                    if max_batch_sentences is not None and n == max_batch_sentences:
                        break
                    n += 1
                stop += 1
            group = plan[start:stop]
            rows = self._rows(group)         # settings / voice are read here, i.e. when this group is spoken
            by_voice: typing.Dict[str, typing.List[int]] = {}
            for i, row in enumerate(rows):
                if row is not None:
                    by_voice.setdefault(row[0], []).append(i)
            audio: typing.Dict[int, np.ndarray] = {}
            for key, idxs in by_voice.items():
                voice = rows[idxs[0]][1]
                st = [rows[i][3] for i in idxs]
                outs = voice.ids_to_audio_rows(
                    [rows[i][2] for i in idxs], speakers=[s.speaker for s in st], length_scales=[s.length_scale for s in st],
                    noise_scales=[s.noise_scale for s in st], noise_ws=[s.noise_w for s in st], rates=[s.rate for s in st],
                    volumes=[s.volume for s in st])
                audio.update(zip(idxs, outs))
            for i, item in enumerate(group):
                if rows[i] is None:
                    yield item
                else:
                    yield AudioResult(sample_rate_hz=rows[i][1].config.audio.sample_rate, audio_bytes=audio[i].tobytes(),
                                      sample_width_bytes=2, num_channels=1)
            if stop >= len(plan):
                break
            start = stop
        self._results.clear()

    def end_utterance_wav(self) -> bytes:
        """``results_to_wav_bytes(end_utterance())`` with the post chain on the device: when every sentence uses the
        same voice and the breaks share its sample rate, the silences, the volume scaling and the 44-byte
        header come out of ONE ``m3_infer_ex`` call as one buffer (no host pass over the samples)."""
        plan = plan_sentences(self._results)
        rows = self._rows(plan)
        sent_idx = [i for i, r in enumerate(rows) if r is not None]
        keys = {rows[i][0] for i in sent_idx}
        single = bool(sent_idx) and len(keys) == 1 and all(
            not isinstance(it, AudioResult) or (it.sample_rate_hz == rows[sent_idx[0]][1].config.audio.sample_rate
                                                and it.sample_width_bytes == 2 and it.num_channels == 1)
            for it in plan)
        if not sent_idx or not single:
            return results_to_wav_bytes(self.end_utterance())
        voice = rows[sent_idx[0]][1]
        lead = [0] * len(sent_idx)
        trail = [0] * len(sent_idx)
        pos = -1  # index into sent_idx of the last sentence seen
        for i, item in enumerate(plan):
            if rows[i] is not None:
                pos += 1
            elif isinstance(item, AudioResult):
                n = len(item.audio_bytes) // 2
                if pos < 0:
                    lead[0] += n
                else:
                    trail[pos] += n
        st = [rows[i][3] for i in sent_idx]
        ids = [rows[i][2] for i in sent_idx]
        lengths = np.array([len(p) for p in ids], dtype=np.int64)
        text = np.zeros((len(ids), max(1, int(lengths.max()))), dtype=np.int64)
        for b, p in enumerate(ids):
            text[b, : len(p)] = np.asarray(p, dtype=np.int64)
        row_scales = np.stack([voice._scales(s.length_scale, s.noise_scale, s.noise_w, s.rate) for s in st]).astype(np.float32)
        sid = None
        if voice.config.is_multispeaker and voice.onnx_model.info.has_speaker_embedding:
            sid = np.array([voice._resolve_speaker(s.speaker) for s in st], dtype=np.int64)
        volume = None
        if any(s.volume != DEFAULT_VOLUME for s in st):
            volume = np.array([s.volume / 100.0 for s in st], dtype=np.float64)
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if (row_scales[:, 0].any() or row_scales[:, 2].any()) else 0
        r = voice.onnx_model.infer(text, lengths, None, sid, seed=seed, row_scales=row_scales, volume=volume,
                                   lead_silence=np.array(lead, dtype=np.int64), trail_silence=np.array(trail, dtype=np.int64),
                                   wav_header=True)
        self._results.clear()
        return r.stream_bytes()
