"""Host-side mirror of ``Mimic3Voice`` for the ids->audio hot path.

Same names, argument meaning and error behaviour as the reference
(``mimic3_tts/voice.py``): :meth:`B200Voice.ids_to_audio` follows ``voice.py:154-243``
line by line, :meth:`B200Voice.load_from_directory` reads the same voice directory
(``voice.py:246-321``), and sessions are shared per ``generator.onnx`` path under a
lock (``voice.py:71-72,277-299``).  The text front-ends (``text_to_phonemes`` --
gruut / espeak-ng / epitran, ``voice.py:413-775``) are out of scope:
phonemes or ids are the input.  Additions: :meth:`ids_to_audio_batch` (one setting for all rows) and
:meth:`ids_to_audio_rows` (per-sentence settings + volume in ONE engine call, what the batched
``end_utterance`` of ``mimic3_b200.tts`` uses).
"""
from __future__ import annotations

import csv
import json
import logging
import sys
import threading
import time
import typing
from pathlib import Path
from types import SimpleNamespace

import numpy as np

from .engine import B200Session

_LOGGER = logging.getLogger(__name__)

try:
    # the package the reference itself calls (voice.py:33, 133-152, 268-271, 302-307): whenever it is installed it
    # is the id conversion, so ids -- and audio -- are the reference's by construction
    from phonemes2ids import load_phoneme_ids, load_phoneme_map, phonemes2ids
    if getattr(sys.modules.get("phonemes2ids"), "__name__", "") == "mimic3_b200.phonemes":
        raise ImportError  # a test registered the restatement under that name
    PHONEMES2IDS_SOURCE = "phonemes2ids"
except ImportError:
    # without the package: load_from_directory builds a native table (csrc/phonemes.cc through libm3b200.so, see
    # NativePhonemeTable); the Python restatement below is what a bare B200Voice(...) falls back to and what the
    # native code is fuzzed against
    from .phonemes import load_phoneme_ids, load_phoneme_map, phonemes2ids
    PHONEMES2IDS_SOURCE = "mimic3_b200.phonemes"
    _LOGGER.info("phonemes2ids is not installed: phonemes -> ids runs natively in libm3b200 (parity with the package is "
                 "pinned by its documented behaviour only)")

DEFAULT_RATE = 1.0
DEFAULT_VOLUME = 100.0   # mimic3_tts/const.py
B200_PROVIDER = "B200ExecutionProvider"


def _ns(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: _ns(v) for k, v in d.items()})
    if isinstance(d, list):
        return [_ns(v) for v in d]
    return d


class VoiceConfig:
    """The parts of ``TrainingConfig`` (``mimic3_tts/config.py:274-327``) the path reads."""

    def __init__(self, data: dict):
        self.raw = data
        audio = {"sample_rate": 22050, "hop_length": 256, **data.get("audio", {})}
        inference = {"length_scale": 1.0, "noise_scale": 0.667, "noise_w": 0.8, **data.get("inference", {})}
        model = {"n_speakers": 1, **data.get("model", {})}
        self.audio = _ns(audio)
        self.inference = _ns(inference)
        self.model = _ns(model)
        # PhonemesConfig defaults, config.py:147-176
        phonemes = {"pad": "_", "bos": None, "eos": None, "blank": "#", "blank_word": None, "blank_between": "words",
                    "blank_at_start": True, "blank_at_end": True, "simple_punctuation": True, "punctuation_map": None,
                    "separate": None, "separate_graphemes": False, "separate_tones": False, "tone_before": False,
                    "phoneme_map": None, "auto_bos_eos": False, **(data.get("phonemes") or {})}
        self.phonemes = SimpleNamespace(**phonemes)  # maps stay dicts
        self.datasets = [_ns(d) for d in data.get("datasets", [])]
        self.phonemizer = data.get("phonemizer")
        self.text_language = data.get("text_language")

    @property
    def is_multispeaker(self) -> bool:  # config.py:316-318
        return self.model.n_speakers > 1 or any(getattr(d, "multispeaker", False) for d in self.datasets)

    @staticmethod
    def load(config_file: typing.TextIO) -> "VoiceConfig":
        return VoiceConfig(json.loads(config_file.read()))


def registry_sha256(voice_dir, voices_registry, relative_path: str = "generator.onnx") -> typing.Optional[str]:
    """``sha256_sum`` the reference's voice registry lists for a file of this voice, or None if the voice / file is not
    listed.  ``voices_registry`` is ``mimic3_tts/voices.json`` as a path or the parsed dict
    ``{"<lang>/<voice>": {"files": {"<relative path>": {"size_bytes", "sha256_sum"}}}}`` (``_resources.py:35-51``);
    the voice key is the last two components of the voice directory (``voice_dir = voices_dir / voice_key``, ``download.py:86``)."""
    if not isinstance(voices_registry, dict):
        with open(voices_registry, "r", encoding="utf-8") as f:
            voices_registry = json.load(f)
    voice_dir = Path(voice_dir)
    entry = voices_registry.get(f"{voice_dir.parent.name}/{voice_dir.name}")
    if not entry:
        return None
    info = (entry.get("files") or {}).get(relative_path)
    return (info or {}).get("sha256_sum") or None


class B200Voice:
    """Drop-in for the ids->audio half of ``Mimic3Voice`` backed by libm3b200."""

    _SHARED_MODELS: typing.Dict[str, B200Session] = {}
    _SHARED_MODELS_LOCK = threading.Lock()

    def __init__(self, config, onnx_model, phoneme_to_id, phoneme_map=None, speaker_map=None,
                 phonemes_to_ids_fn=None):
        self.config = config
        self.onnx_model = onnx_model
        self.phoneme_to_id = phoneme_to_id
        self.phoneme_map = phoneme_map
        self.speaker_map = speaker_map
        # any callable with the phonemes2ids.phonemes2ids keyword signature (voice.py:133-152); default: the real
        # phonemes2ids package when it is installed, else the restatement in mimic3_b200.phonemes (module top)
        self.phonemes_to_ids_fn = phonemes_to_ids_fn or phonemes2ids

    # -- voice.py:126-152 ---------------------------------------------------------------
    def phonemes_to_ids(self, phonemes) -> typing.List[int]:
        """Convert phonemes to ids for a voice model (see phonemes.txt)."""
        pc = self.config.phonemes
        phoneme_map = self.phoneme_map or pc.phoneme_map
        return self.phonemes_to_ids_fn(
            word_phonemes=phonemes, phoneme_to_id=self.phoneme_to_id, pad=pc.pad, bos=pc.bos, eos=pc.eos,
            auto_bos_eos=pc.auto_bos_eos, blank=pc.blank, blank_word=pc.blank_word, blank_between=pc.blank_between,
            blank_at_start=pc.blank_at_start, blank_at_end=pc.blank_at_end, simple_punctuation=pc.simple_punctuation,
            punctuation_map=pc.punctuation_map, separate=pc.separate, separate_graphemes=pc.separate_graphemes,
            separate_tones=pc.separate_tones, tone_before=pc.tone_before, phoneme_map=phoneme_map,
            fail_on_missing=False)

    # -- voice.py:89 ------------------------------------------------------------------
    def text_to_phonemes(self, text, text_language=None):
        raise NotImplementedError(
            "text front-ends are outside the B200 hot path; use the reference Mimic3Voice "
            "subclasses and pass their ids to ids_to_audio (see INTEGRATION.md)")

    def _resolve_speaker(self, speaker) -> int:
        """voice.py:196-215."""
        speaker_id = 0
        if isinstance(speaker, str):
            if self.speaker_map:
                maybe = self.speaker_map.get(speaker)
                if maybe is None:
                    try:
                        speaker_id = int(speaker)
                    except ValueError:
                        _LOGGER.warning(
                            "Unable to find a speaker with the name '%s'. Falling back to first speaker.", speaker)
                else:
                    speaker_id = maybe
            # (reference: a str speaker without a speaker_map is ignored -> id 0)
        elif speaker is not None:
            speaker_id = speaker
        return int(speaker_id)

    def _scales(self, length_scale, noise_scale, noise_w, rate):
        """voice.py:166-178."""
        if length_scale is None:
            length_scale = self.config.inference.length_scale
        if rate > 0:
            length_scale /= rate
        if noise_scale is None:
            noise_scale = self.config.inference.noise_scale
        if noise_w is None:
            noise_w = self.config.inference.noise_w
        return np.array([noise_scale, length_scale, noise_w], dtype=np.float32)

    # -- voice.py:154-243 -------------------------------------------------------------------
    def ids_to_audio(self, phoneme_ids, speaker=None, length_scale=None, noise_scale=None, noise_w=None,
                     rate: float = DEFAULT_RATE, seed: typing.Optional[int] = None) -> np.ndarray:
        """Synthesize int16 audio from phoneme ids (one utterance)."""
        return self.ids_to_audio_batch([phoneme_ids], [speaker], length_scale, noise_scale, noise_w, rate, seed)[0]

    def ids_to_audio_batch(self, batch_ids, speakers=None, length_scale=None, noise_scale=None, noise_w=None,
                           rate: float = DEFAULT_RATE, seed: typing.Optional[int] = None,
                           ) -> typing.List[np.ndarray]:
        """Batched extension: every utterance keeps batch-1 edge semantics
        and its own peak normalisation, so ``out[i]`` equals ``ids_to_audio(batch_ids[i])``."""
        scales = self._scales(length_scale, noise_scale, noise_w, rate)
        n = len(batch_ids)
        lengths = np.array([len(p) for p in batch_ids], dtype=np.int64)
        text = np.zeros((n, max(1, int(lengths.max()) if n else 1)), dtype=np.int64)
        for i, p in enumerate(batch_ids):
            text[i, : len(p)] = np.asarray(p, dtype=np.int64)
        sid = None
        if self.config.is_multispeaker:
            speakers = speakers if speakers is not None else [None] * n
            sid = np.array([self._resolve_speaker(s) for s in speakers], dtype=np.int64)
        _LOGGER.debug("TTS settings: speaker-id=%s, length-scale=%s, noise-scale=%s, noise-w=%s",
                      None if sid is None else sid.tolist(), scales[1], scales[0], scales[2])
        start_time = time.perf_counter()
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) if (scales[0] or scales[2]) else 0
        if sid is not None and not self.onnx_model.info.has_speaker_embedding:
            sid = None  # config says multispeaker but the graph has no "sid" input
        r = self.onnx_model.infer(text, lengths, scales, sid, seed=seed)
        audios = [r.utterance_pcm(b) for b in range(n)]
        end_time = time.perf_counter()
        audio_sec = r.total_samples / self.config.audio.sample_rate
        _LOGGER.debug("RTF: %s", (end_time - start_time) / audio_sec if audio_sec > 0 else 0.0)
        return audios

    def ids_to_audio_rows(self, batch_ids, speakers=None, length_scales=None, noise_scales=None, noise_ws=None,
                          rates=None, volumes=None, seed: typing.Optional[int] = None,
                          ) -> typing.List[np.ndarray]:
        """One engine call for sentences that each carry their own settings (``Mimic3Settings`` of the
        sentence: speaker / length_scale / noise_scale / noise_w / rate / volume, ``tts.py:519-543``).
        ``volumes`` are the reference's 0-100 values; rows whose volume is not 100 go through the device
        equivalent of ``audioop.mul(audio, 2, volume / 100)`` (``tts.py:540-543``).  ``out[i]`` is what
        ``_speak_sentence_phonemes`` would hand to ``AudioResult`` for sentence i."""
        n = len(batch_ids)
        pick = lambda seq, i, default=None: default if seq is None else seq[i]
        rows = np.stack([self._scales(pick(length_scales, i), pick(noise_scales, i), pick(noise_ws, i),
                                      pick(rates, i, DEFAULT_RATE)) for i in range(n)]) if n else np.zeros((0, 3), np.float32)
        lengths = np.array([len(p) for p in batch_ids], dtype=np.int64)
        text = np.zeros((n, max(1, int(lengths.max()) if n else 1)), dtype=np.int64)
        for i, p in enumerate(batch_ids):
            text[i, : len(p)] = np.asarray(p, dtype=np.int64)
        sid = None
        if self.config.is_multispeaker and self.onnx_model.info.has_speaker_embedding:
            sid = np.array([self._resolve_speaker(pick(speakers, i)) for i in range(n)], dtype=np.int64)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) if (rows[:, 0].any() or rows[:, 2].any()) else 0
        volume = None
        if volumes is not None and any(v != DEFAULT_VOLUME for v in volumes):
            volume = np.array([v / 100.0 for v in volumes], dtype=np.float64)
        r = self.onnx_model.infer(text, lengths, None, sid, seed=seed, row_scales=rows.astype(np.float32), volume=volume)
        return [r.utterance_pcm(b) for b in range(n)]

    # -- voice.py:245-376 ------------------------------------------------------------------------
    @staticmethod
    def load_from_directory(voice_dir, session_options=None, providers=None, share_models: bool = True,
                            use_deterministic_compute: bool = False, device: typing.Optional[int] = None,
                            weight_cache_dir=None, voices_registry=None) -> "B200Voice":
        """``weight_cache_dir``: directory of packed-weight blobs (one-time conversion of generator.onnx, re-read on
        later loads; ``None`` = ``$M3B200_WEIGHT_CACHE``).  ``voices_registry``: the reference's ``voices.json``
        (path or parsed dict, ``mimic3_tts/_resources.py:35-51``): when the voice is listed there, generator.onnx
        must have the registry's ``sha256_sum`` (the check ``download.py:108-117`` makes) or loading raises."""
        voice_dir = Path(voice_dir)
        expected_sha256 = registry_sha256(voice_dir, voices_registry) if voices_registry is not None else None
        _LOGGER.debug("Loading voice from %s", voice_dir)
        with open(voice_dir / "config.json", "r", encoding="utf-8") as config_file:
            config = VoiceConfig.load(config_file)
        with open(voice_dir / "phonemes.txt", "r", encoding="utf-8") as ids_file:
            phoneme_to_id = load_phoneme_ids(ids_file)
        generator_path = voice_dir / "generator.onnx"
        if share_models:
            with B200Voice._SHARED_MODELS_LOCK:
                model_key = f"{generator_path.absolute()}@{device}"
                onnx_model = B200Voice._SHARED_MODELS.get(model_key)
                if onnx_model is None:
                    onnx_model = B200Voice._load_model(generator_path, session_options, providers,
                                                       use_deterministic_compute, device,
                                                       weight_cache_dir, expected_sha256)
                    B200Voice._SHARED_MODELS[model_key] = onnx_model
                else:
                    _LOGGER.debug("Using shared B200 model (%s)", model_key)
        else:
            onnx_model = B200Voice._load_model(generator_path, session_options, providers,
                                               use_deterministic_compute, device, weight_cache_dir, expected_sha256)
        phoneme_map = None
        phoneme_map_path = voice_dir / "phoneme_map.txt"
        if phoneme_map_path.is_file():
            with open(phoneme_map_path, "r", encoding="utf-8") as map_file:
                phoneme_map = load_phoneme_map(map_file)
        speaker_map = None
        speaker_map_path = voice_dir / "speaker_map.csv"
        if speaker_map_path.is_file():
            with open(speaker_map_path, "r", encoding="utf-8") as map_file:
                reader = csv.reader(map_file, delimiter="|")
                speaker_map = {}
                for row in reader:
                    if not row:
                        continue
                    speaker_id = int(row[0])
                    for alias in row[2:]:
                        speaker_map[alias] = speaker_id
        fn = None
        if PHONEMES2IDS_SOURCE != "phonemes2ids":
            # native table lookup: phonemes.txt / phoneme_map.txt read by the C library
            from .phonemes import NativePhonemeTable
            table = NativePhonemeTable.from_files(voice_dir / "phonemes.txt",
                                                  phoneme_map_path if phoneme_map_path.is_file() else None)
            if phoneme_map is None:
                table.set_phoneme_map(getattr(config.phonemes, "phoneme_map", None))
            fn = table.phonemes2ids
        return B200Voice(config=config, onnx_model=onnx_model, phoneme_to_id=phoneme_to_id,
                         phoneme_map=phoneme_map, speaker_map=speaker_map, phonemes_to_ids_fn=fn)

    # -- voice.py:378-407 ------------------------------------------------------------------------
    @staticmethod
    def _load_model(generator_path, session_options=None, providers=None,
                    use_deterministic_compute: bool = False, device: typing.Optional[int] = None,
                    weight_cache_dir=None, expected_sha256: typing.Optional[str] = None) -> B200Session:
        _LOGGER.debug("Loading model from %s", generator_path)
        return B200Session(str(generator_path), sess_options=session_options, providers=providers, device=device,
                           cache_dir=weight_cache_dir, expected_sha256=expected_sha256)
