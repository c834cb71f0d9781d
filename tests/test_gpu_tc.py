"""tcgen05 / TMEM building-block self-tests (descriptor conventions, tap-shifted A operand,
accumulate on top of tcgen05.st) run on the GPU through the C ABI."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu

CASES = {0: "bf16 K32 N32 k3", 1: "bf16 K64 N64 k7 d12", 2: "bf16 K128 N128 k5 d6 +init",
         3: "fp16 K32 N32 k3 d2 +init", 4: "bf16 K192 N256 1x1", 6: "bf16 K96 N192 1x1",
         7: "fp16 K64 N64 k7 d3 +init"}


@pytest.mark.parametrize("which", sorted(CASES))
def test_tcgen05_blocks(built_library, which):
    from mimic3_b200.engine import load_library
    lib = load_library()
    err = ctypes.c_double()
    assert lib.m3_selftest(which, ctypes.byref(err)) == 0
    print(f"selftest {which} ({CASES[which]}): max abs err {err.value:.3e}")
    assert 0 <= err.value < 2e-3, f"{CASES[which]}: {err.value}"
