"""Pins csrc/phonemes.cc and mimic3_b200/phonemes.py on the REAL phonemes2ids package (rhasspy/phonemes2ids,
`phonemes2ids>=1.2,<2` in the reference's requirements.txt; called at mimic3_tts/voice.py:133-152).  The package cannot
be installed offline, so run this wherever it imports:

    python tests/golden/make_golden_phonemes2ids.py        # writes tests/golden/phonemes2ids_golden.json

tests/test_phonemes_native.py::test_native_matches_the_real_package_goldens consumes the file when it exists."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    try:
        import phonemes2ids
    except ImportError:
        print("phonemes2ids is not importable here: no golden written (parity with the package stays unpinned)")
        return 1
    from test_phonemes_native import _random_case
    rng = np.random.default_rng(77)
    cases = []
    for _ in range(400):
        p2i, pmap, words, kw = _random_case(rng)
        ids = phonemes2ids.phonemes2ids(word_phonemes=words, phoneme_to_id=p2i, phoneme_map=pmap, fail_on_missing=False, **kw)
        cases.append({"phoneme_to_id": p2i, "phoneme_map": pmap, "words": words, "kwargs": kw, "ids": list(ids)})
    out = Path(__file__).parent / "phonemes2ids_golden.json"
    out.write_text(json.dumps(cases, ensure_ascii=False), encoding="utf-8")
    print(out, len(cases), "cases from phonemes2ids", getattr(phonemes2ids, "__version__", "?"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
