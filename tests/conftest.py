import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
# the unmodified reference package (mimic3_tts, opentts_abc), installed with --no-deps into baseline/_ref:
# on the path from the start so that `opentts_abc` resolves to the reference's own module everywhere
_REF = ROOT / "baseline" / "_ref"
if (_REF / "opentts_abc").is_dir() and str(_REF) not in sys.path:
    sys.path.append(str(_REF))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_library():
    """libm3b200.so, built in-tree if absent (nvcc cross-compiles without a GPU)."""
    from mimic3_b200 import build
    from mimic3_b200.engine import library_path

    if not library_path().exists():
        build.build()
    return library_path()


@pytest.fixture(scope="session")
def voices(tmp_path_factory):
    """Synthetic voice directories, written once per session (name -> path)."""
    from mimic3_b200 import synth_voice as sv

    root = tmp_path_factory.mktemp("voices")
    made = {}

    def get(name):
        if name in made:
            return made[name]
        d = root / name
        if name == "tiny":
            sv.write_voice(d, sv.tiny_config(), seed=11)
        elif name == "tiny_ms":
            sv.write_voice(d, sv.tiny_config(n_speakers=3), seed=12)
        elif name == "tiny_ms_folded":
            sv.write_voice(d, sv.tiny_config(n_speakers=3), seed=12, style="folded", alt_encoding=True)
        elif name == "tiny_ms_wn":
            sv.write_voice(d, sv.tiny_config(n_speakers=3), seed=12, style="weightnorm")
        elif name == "tiny_rb1_dp":
            sv.write_voice(d, sv.tiny_config(n_speakers=2, resblock="1", use_sdp=False), seed=13)
        elif name == "low":
            sv.write_voice(d, sv.low_config(), seed=21)
        elif name == "low_ms":
            sv.write_voice(d, sv.low_config(n_speakers=109), seed=22)
        else:
            raise KeyError(name)
        made[name] = d
        return d

    return get


def rand_ids(rng, num_symbols, length):
    return rng.integers(4, num_symbols, size=length).astype(np.int64)
