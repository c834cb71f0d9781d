"""Native phonemes -> ids (csrc/phonemes.cc) against the Python restatement of phonemes2ids
(mimic3_b200/phonemes.py) on randomised voices, and -- whenever tests/golden/phonemes2ids_golden.json exists (made by
tests/golden/make_golden_phonemes2ids.py on a machine that has the real package) -- against the package itself."""
import json
from pathlib import Path

import numpy as np
import pytest

from mimic3_b200 import phonemes as ph

GOLDEN = Path(__file__).parent / "golden" / "phonemes2ids_golden.json"

ALPHABET = ["a", "b", "ɛ", "ŋ", "ˈ", "ˌ", "ː", "aː", "ˈa", "t͡ʃ", "é", "é", "1", "2", "12", "˥", "˥˩", "a1", "ɕ˥˩",
            ",", ".", ";", "?", "!", ":", "#", "_", "^", "$", " ", "x", "yy", "👩‍🚀", "ñ̀", "b2˥"]


def _random_case(rng):
    n_sym = int(rng.integers(5, len(ALPHABET)))
    syms = list(rng.choice(ALPHABET, size=n_sym, replace=False))
    p2i = {s: i for i, s in enumerate(syms)}
    pmap = {}
    for _ in range(int(rng.integers(0, 4))):
        pmap[str(rng.choice(ALPHABET))] = [str(s) for s in rng.choice(ALPHABET, size=int(rng.integers(1, 4)))]
    words = [[str(s) for s in rng.choice(ALPHABET + ["", "zz"], size=int(rng.integers(0, 6)))]
             for _ in range(int(rng.integers(0, 6)))]
    opt = lambda choices: choices[int(rng.integers(0, len(choices)))]
    kw = dict(pad=opt([None, "_"]), bos=opt([None, "^"]), eos=opt([None, "$"]), auto_bos_eos=bool(rng.integers(0, 2)),
              blank=opt([None, "#", "_"]), blank_word=opt([None, " ", "#"]),
              blank_between=opt(["tokens", "words", "tokens_and_words"]), blank_at_start=bool(rng.integers(0, 2)),
              blank_at_end=bool(rng.integers(0, 2)), simple_punctuation=bool(rng.integers(0, 2)),
              punctuation_map=opt([None, {"?": ",", ";": "."}, {}]),
              separate=opt([None, ["ˈ"], ["ˈ", "ˌ", "ː"], ["aː", "a", "˥˩"]]),
              separate_graphemes=bool(rng.integers(0, 2)), separate_tones=bool(rng.integers(0, 2)),
              tone_before=bool(rng.integers(0, 2)))
    return p2i, pmap, words, kw


def test_native_equals_restatement_on_random_voices(built_library):
    rng = np.random.default_rng(2024)
    for case in range(600):
        p2i, pmap, words, kw = _random_case(rng)
        table = ph.NativePhonemeTable.from_mappings(p2i, pmap)
        want = ph.phonemes2ids(words, p2i, phoneme_map=pmap, **kw)
        got = table.phonemes2ids(words, phoneme_to_id=p2i, phoneme_map=pmap, **kw)
        assert got == want, (case, words, kw, p2i, pmap)
        table.close()


def test_native_loaders_read_the_voice_files_like_the_restatement(built_library, tmp_path):
    txt = "# comment\n0 _\n1 ^\n2 $\n3 #\n\n4 a\n5 b c\r\n6  \n7 ˈ\n8 t͡ʃ\n9\n"
    (tmp_path / "phonemes.txt").write_text(txt, encoding="utf-8")
    (tmp_path / "phoneme_map.txt").write_text("x a b\ny ˈ\nbad\n z  a\nw a  b\n", encoding="utf-8")
    with open(tmp_path / "phonemes.txt", encoding="utf-8") as f:
        p2i = ph.load_phoneme_ids(f)
    with open(tmp_path / "phoneme_map.txt", encoding="utf-8") as f:
        pmap = ph.load_phoneme_map(f)
    table = ph.NativePhonemeTable.from_files(tmp_path / "phonemes.txt", tmp_path / "phoneme_map.txt")
    assert len(table) == len(p2i)
    for k, v in p2i.items():
        assert table[k] == v, repr(k)
    words = [["x", "y", "w"], ["a", "b c", " ", "t͡ʃ", "", "z"]]
    for kw in (dict(blank="#"), dict(blank="#", blank_between="tokens", auto_bos_eos=True, bos="^", eos="$")):
        assert table.phonemes2ids(words, **kw) == ph.phonemes2ids(words, p2i, phoneme_map=pmap, **kw)
    with pytest.raises(FileNotFoundError):
        ph.NativePhonemeTable.from_files(tmp_path / "absent.txt")
    (tmp_path / "broken.txt").write_text("x a\n")
    with pytest.raises(Exception, match="not an integer"):
        ph.NativePhonemeTable.from_files(tmp_path / "broken.txt")


def test_b200voice_uses_the_native_table_when_the_package_is_absent(built_library, voices):
    from mimic3_b200 import voice as bv
    if bv.PHONEMES2IDS_SOURCE == "phonemes2ids":
        pytest.skip("the real phonemes2ids package is installed: it is the id conversion")
    import inspect
    src = inspect.getsource(bv.B200Voice.load_from_directory)
    assert "NativePhonemeTable.from_files" in src
    vd = voices("tiny")
    table = ph.NativePhonemeTable.from_files(vd / "phonemes.txt")
    with open(vd / "phonemes.txt", encoding="utf-8") as f:
        p2i = ph.load_phoneme_ids(f)
    syms = [s for s in p2i if s]
    words = [syms[:3], syms[3:7], ["?"]]
    kw = dict(pad="_", blank="#", blank_between="words", simple_punctuation=True)
    assert table.phonemes2ids(words, **kw) == ph.phonemes2ids(words, p2i, **kw)


@pytest.mark.skipif(not GOLDEN.exists(), reason="no golden from the real phonemes2ids package (cannot be imported offline)")
def test_native_matches_the_real_package_goldens(built_library):
    for case in json.loads(GOLDEN.read_text(encoding="utf-8")):
        table = ph.NativePhonemeTable.from_mappings(case["phoneme_to_id"], case["phoneme_map"])
        assert table.phonemes2ids(case["words"], **case["kwargs"]) == case["ids"], case
        assert ph.phonemes2ids(case["words"], case["phoneme_to_id"], phoneme_map=case["phoneme_map"], **case["kwargs"]) == case["ids"]
