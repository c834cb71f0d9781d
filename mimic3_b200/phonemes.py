"""phonemes -> ids, the step immediately before the engine.

The reference does this in ``Mimic3Voice.phonemes_to_ids`` (``mimic3_tts/voice.py:126-152``) by calling the
third-party package ``phonemes2ids`` (``requirements.txt``: ``phonemes2ids>=1.2,<2``; rhasspy/phonemes2ids,
absent from the reference tree and not installable offline) with the ``PhonemesConfig`` of the voice
(``mimic3_tts/config.py:147-176``).  This module restates that package's published algorithm
(``phonemes2ids.phonemes2ids``, v1.2) so id-level batches can be built without it:

1. every phoneme goes through ``phoneme_map`` (one phoneme -> list of phonemes), then -- with
   ``simple_punctuation`` -- through ``punctuation_map`` (default ``; : -> ,`` and ``? ! -> .``);
2. optional splitting: ``separate`` (symbols such as stress marks split off as their own phonemes),
   ``separate_graphemes`` (every grapheme cluster its own phoneme), ``separate_tones`` (trailing tone
   digits/letters split off, before the phoneme when ``tone_before``);
3. lookup in ``phoneme_to_id``; unknown phonemes are dropped (``fail_on_missing=False`` at voice.py:151);
4. ``bos`` / ``eos`` are added as words of their own when ``auto_bos_eos``; the ``blank`` id is
   interleaved between tokens and/or words per ``blank_between`` (``blank_word`` id between words when
   set) and added at the start / end per ``blank_at_start`` / ``blank_at_end``.

**Parity unpinned**: neither the package nor a voice's ``phonemes.txt`` is available here and the
reference holds no id-level golden for this step, so the tests pin this module's own documented
behaviour only.  ``B200Voice`` therefore accepts any callable with this signature
(``phonemes_to_ids_fn``), e.g. the real ``phonemes2ids.phonemes2ids`` where it is installed.
"""
from __future__ import annotations

import re
import typing
import unicodedata
from enum import Enum

PHONEME_ID_TYPE = int
WORD_PHONEMES_TYPE = typing.Sequence[typing.Sequence[str]]

DEFAULT_PUNCTUATION_MAP = {";": ",", ":": ",", "?": ".", "!": "."}


class BlankBetween(str, Enum):
    """``mimic3_tts/config.py`` / phonemes2ids: where the blank symbol goes."""

    TOKENS = "tokens"
    WORDS = "words"
    TOKENS_AND_WORDS = "tokens_and_words"


def load_phoneme_ids(ids_file: typing.Iterable[str]) -> typing.Dict[str, int]:
    """``phonemes.txt``: ``<id><space><phoneme>`` per line (``phonemes2ids.load_phoneme_ids``, voice.py:268-271).
    Blank lines and ``#`` comment lines are skipped; the phoneme keeps inner/trailing spaces."""
    out: typing.Dict[str, int] = {}
    for line in ids_file:
        line = line.rstrip("\r\n")
        if not line.strip() or line.startswith("#"):
            continue
        pid, _, phoneme = line.partition(" ")
        out[phoneme] = int(pid)
    return out


def load_phoneme_map(map_file: typing.Iterable[str]) -> typing.Dict[str, typing.List[str]]:
    """``phoneme_map.txt``: ``<from> <to> [<to> ...]`` per line (voice.py:302-307)."""
    out: typing.Dict[str, typing.List[str]] = {}
    for line in map_file:
        parts = line.strip("\r\n").split(" ")
        if len(parts) >= 2 and parts[0]:
            out[parts[0]] = parts[1:]
    return out


def _graphemes(text: str) -> typing.List[str]:
    """Grapheme clusters: a base code point followed by its combining marks (and ZWJ sequences)."""
    out: typing.List[str] = []
    for ch in text:
        if out and (unicodedata.combining(ch) or unicodedata.category(ch) in ("Mn", "Me", "Mc") or ch == "‍"
                    or out[-1].endswith("‍")):
            out[-1] += ch
        else:
            out.append(ch)
    return out


_TONE = re.compile(r"^(.*?)([0-9]+|[˥˦˧˨˩]+)$")


def _split(phoneme: str, separate: typing.Collection[str], separate_graphemes: bool, separate_tones: bool,
           tone_before: bool) -> typing.List[str]:
    parts = [phoneme]
    if separate_tones:
        m = _TONE.match(phoneme)
        if m and m.group(1):
            parts = [m.group(2), m.group(1)] if tone_before else [m.group(1), m.group(2)]
    if separate:
        seps = sorted(separate, key=len, reverse=True)
        pattern = re.compile("(" + "|".join(re.escape(s) for s in seps) + ")")
        parts = [p for part in parts for p in pattern.split(part) if p]
    if separate_graphemes:
        parts = [g for part in parts for g in _graphemes(part)]
    return parts


def phonemes2ids(word_phonemes: WORD_PHONEMES_TYPE, phoneme_to_id: typing.Mapping[str, int],
                 pad: typing.Optional[str] = None, bos: typing.Optional[str] = None, eos: typing.Optional[str] = None,
                 auto_bos_eos: bool = False, blank: typing.Optional[str] = None,
                 blank_word: typing.Optional[str] = None,
                 blank_between: typing.Union[str, BlankBetween] = BlankBetween.WORDS, blank_at_start: bool = True,
                 blank_at_end: bool = True, simple_punctuation: bool = False,
                 punctuation_map: typing.Optional[typing.Mapping[str, str]] = None,
                 separate: typing.Optional[typing.Collection[str]] = None, separate_graphemes: bool = False,
                 separate_tones: bool = False, tone_before: bool = False,
                 phoneme_map: typing.Optional[typing.Mapping[str, typing.Sequence[str]]] = None,
                 fail_on_missing: bool = False,
                 missing_func: typing.Optional[typing.Callable[[str], typing.Optional[typing.List[int]]]] = None,
                 ) -> typing.List[PHONEME_ID_TYPE]:
    """Same keyword arguments as ``phonemes2ids.phonemes2ids`` as called at ``voice.py:133-152``."""
    if punctuation_map is None:
        punctuation_map = DEFAULT_PUNCTUATION_MAP
    blank_between = BlankBetween(blank_between)
    separate = set(separate or ())
    phoneme_map = phoneme_map or {}

    def lookup(symbol: typing.Optional[str]) -> typing.Optional[int]:
        return None if symbol is None else phoneme_to_id.get(symbol)

    blank_id = lookup(blank)
    blank_word_id = lookup(blank_word)
    if blank_word_id is None:
        blank_word_id = blank_id

    words: typing.List[typing.List[int]] = []
    if auto_bos_eos and lookup(bos) is not None:
        words.append([typing.cast(int, lookup(bos))])
    for word in word_phonemes:
        ids: typing.List[int] = []
        for phoneme in word:
            if not phoneme:
                continue
            mapped = phoneme_map.get(phoneme)
            if mapped is None:
                mapped_list = [phoneme]
            elif isinstance(mapped, str):
                mapped_list = [mapped]
            else:
                mapped_list = list(mapped)
            for p in mapped_list:
                if simple_punctuation:
                    p = punctuation_map.get(p, p)
                for sub in _split(p, separate, separate_graphemes, separate_tones, tone_before):
                    pid = phoneme_to_id.get(sub)
                    if pid is not None:
                        ids.append(pid)
                    elif missing_func is not None:
                        ids.extend(missing_func(sub) or [])
                    elif fail_on_missing:
                        raise KeyError(sub)
        if ids:
            words.append(ids)
    if auto_bos_eos and lookup(eos) is not None:
        words.append([typing.cast(int, lookup(eos))])
    if not words:
        return []

    out: typing.List[int] = []
    if blank_at_start and blank_id is not None:
        out.append(blank_id)
    between_tokens = blank_between in (BlankBetween.TOKENS, BlankBetween.TOKENS_AND_WORDS) and blank_id is not None
    between_words = blank_between in (BlankBetween.WORDS, BlankBetween.TOKENS_AND_WORDS) and blank_word_id is not None
    for wi, ids in enumerate(words):
        last_word = wi + 1 == len(words)
        for ti, pid in enumerate(ids):
            out.append(pid)
            last_token = ti + 1 == len(ids)
            if between_tokens and not (last_token and (last_word or between_words)):
                out.append(typing.cast(int, blank_id))
        if between_words and not last_word:
            out.append(typing.cast(int, blank_word_id))
    if blank_at_end and blank_id is not None:
        out.append(blank_id)
    return out


class NativePhonemeTable:
    """phonemes -> ids through ``libm3b200.so`` (``m3_phoneme_table_*`` / ``m3_phonemes_to_ids``, csrc/phonemes.cc):
    ``phonemes.txt`` and ``phoneme_map.txt`` are read natively into hash tables (``voice.py:268-271, 302-307``) and
    :meth:`phonemes2ids` takes the keyword arguments of ``phonemes2ids.phonemes2ids`` as
    ``Mimic3Voice.phonemes_to_ids`` passes them (``voice.py:133-152``).  The table belongs to the object: the
    ``phoneme_to_id`` / ``phoneme_map`` keywords of a call are accepted for signature compatibility and must be the
    mappings the table was made from (``None`` / the voice's own)."""

    def __init__(self):
        import ctypes as C
        from . import engine
        self._C = C
        self._engine = engine
        self._lib = engine.load_library()
        lib = self._lib
        if not getattr(lib, "_m3_ph_bound", False):
            vp, i32 = C.c_void_p, C.c_int32
            lib.m3_phoneme_table_create.restype = i32
            lib.m3_phoneme_table_create.argtypes = [C.POINTER(vp)]
            lib.m3_phoneme_table_free.restype = None
            lib.m3_phoneme_table_free.argtypes = [vp]
            for name in ("m3_phoneme_table_load_ids", "m3_phoneme_table_load_map"):
                getattr(lib, name).restype = i32
                getattr(lib, name).argtypes = [vp, C.c_char_p]
            lib.m3_phoneme_table_add.restype = i32
            lib.m3_phoneme_table_add.argtypes = [vp, C.c_char_p, C.c_int64]
            lib.m3_phoneme_table_add_map.restype = i32
            lib.m3_phoneme_table_add_map.argtypes = [vp, C.c_char_p, C.POINTER(C.c_char_p), i32]
            lib.m3_phoneme_table_size.restype = C.c_int64
            lib.m3_phoneme_table_size.argtypes = [vp]
            lib.m3_phoneme_table_lookup.restype = i32
            lib.m3_phoneme_table_lookup.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
            lib.m3_phonemes_to_ids.restype = i32
            lib.m3_phonemes_to_ids.argtypes = [vp, C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(i32), i32,
                                               C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)]
            lib._m3_ph_bound = True
        h = C.c_void_p()
        self._check(lib.m3_phoneme_table_create(C.byref(h)))
        self._h = h
        self._has_map = False

    def _check(self, rc):
        if rc != 0:
            self._engine._raise(self._lib, rc)

    @classmethod
    def from_files(cls, phonemes_txt, phoneme_map_txt=None) -> "NativePhonemeTable":
        t = cls()
        t._check(t._lib.m3_phoneme_table_load_ids(t._h, str(phonemes_txt).encode()))
        if phoneme_map_txt is not None:
            t._check(t._lib.m3_phoneme_table_load_map(t._h, str(phoneme_map_txt).encode()))
            t._has_map = True
        return t

    @classmethod
    def from_mappings(cls, phoneme_to_id, phoneme_map=None) -> "NativePhonemeTable":
        t = cls()
        for k, v in phoneme_to_id.items():
            t._check(t._lib.m3_phoneme_table_add(t._h, k.encode("utf-8"), int(v)))
        t.set_phoneme_map(phoneme_map)
        return t

    def set_phoneme_map(self, phoneme_map) -> None:
        """``phoneme_map = self.phoneme_map or self.config.phonemes.phoneme_map`` (voice.py:130): entries are added."""
        C = self._C
        for k, v in (phoneme_map or {}).items():
            to = [v] if isinstance(v, str) else list(v)
            arr = (C.c_char_p * max(1, len(to)))(*[s.encode("utf-8") for s in to])
            self._check(self._lib.m3_phoneme_table_add_map(self._h, k.encode("utf-8"), arr, len(to)))
            self._has_map = True

    def __len__(self) -> int:
        return int(self._lib.m3_phoneme_table_size(self._h))

    def __getitem__(self, phoneme: str) -> int:
        out = self._C.c_int64()
        rc = self._lib.m3_phoneme_table_lookup(self._h, phoneme.encode("utf-8"), self._C.byref(out))
        if rc != 0:
            raise KeyError(phoneme)
        return int(out.value)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.m3_phoneme_table_free(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def phonemes2ids(self, word_phonemes, phoneme_to_id=None, pad=None, bos=None, eos=None, auto_bos_eos=False,
                     blank=None, blank_word=None, blank_between=BlankBetween.WORDS, blank_at_start=True,
                     blank_at_end=True, simple_punctuation=False, punctuation_map=None, separate=None,
                     separate_graphemes=False, separate_tones=False, tone_before=False, phoneme_map=None,
                     fail_on_missing=False, missing_func=None) -> typing.List[int]:
        C = self._C
        if fail_on_missing or missing_func is not None:
            raise NotImplementedError("the native table drops unknown phonemes (voice.py:151 passes fail_on_missing=False)")
        from .engine import _PhonemeOpts
        enc = lambda s: None if s is None else s.encode("utf-8")
        o = _PhonemeOpts()
        o.struct_size = C.sizeof(_PhonemeOpts)
        o.flags = ((1 if auto_bos_eos else 0) | (2 if blank_at_start else 0) | (4 if blank_at_end else 0)
                   | (8 if simple_punctuation else 0) | (16 if separate_graphemes else 0)
                   | (32 if separate_tones else 0) | (64 if tone_before else 0))
        o.blank_between = {"tokens": 0, "words": 1, "tokens_and_words": 2}[BlankBetween(blank_between).value]
        o.bos, o.eos, o.blank, o.blank_word = enc(bos), enc(eos), enc(blank), enc(blank_word)
        keep = []
        if punctuation_map is None:
            o.n_punctuation = -1
        else:
            items = list(punctuation_map.items())
            pf = (C.c_char_p * max(1, len(items)))(*[k.encode("utf-8") for k, _ in items])
            pt = (C.c_char_p * max(1, len(items)))(*[v.encode("utf-8") for _, v in items])
            keep += [pf, pt]
            o.punctuation_from, o.punctuation_to, o.n_punctuation = pf, pt, len(items)
        seps = sorted(set(separate or ()), key=len, reverse=True)
        if seps:
            sa = (C.c_char_p * len(seps))(*[s.encode("utf-8") for s in seps])
            keep.append(sa)
            o.separate, o.n_separate = sa, len(seps)
        flat = [p.encode("utf-8") for w in word_phonemes for p in w]
        lens = [len(w) for w in word_phonemes]
        ph = (C.c_char_p * max(1, len(flat)))(*flat)
        wl = (C.c_int32 * max(1, len(lens)))(*lens)
        cap = 4 * len(flat) + 2 * len(lens) + 16
        n = C.c_int64()
        while True:
            out = (C.c_int64 * cap)()
            rc = self._lib.m3_phonemes_to_ids(self._h, C.byref(o), ph, wl, len(lens), out, cap, C.byref(n))
            if rc != 0 and n.value > cap:   # splitting / mapping produced more ids than estimated
                cap = int(n.value)
                continue
            self._check(rc)
            return list(out[: n.value])
