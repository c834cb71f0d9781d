"""Prints the tcgen05/TMEM micro-benchmarks of libm3b200 (cycles per repetition)."""
import ctypes, sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
from mimic3_b200.engine import load_library
lib = load_library()
names = {100: "tcgen05.ld x16 + wait, 4 warps", 101: "same, 8 warps", 102: "same, 1 warp", 103: "4x ld x16 per wait, 4 warps",
         104: "4x ld x16 per wait, 8 warps", 105: "tcgen05.st x16, 4 warps", 106: "tcgen05.st x16, 8 warps",
         110: "SS MMA M128 N32 K16 stream", 111: "SS MMA N64", 112: "SS MMA N128", 113: "SS MMA N256",
         114: "elect-issued MMA N32, 1 accumulator", 115: "elect-issued MMA N128, 1 accumulator",
         116: "elect-issued MMA N32, 4 accumulators", 117: "elect-issued MMA N64, 4 accumulators",
         118: "elect-issued MMA N128, 4 accumulators", 119: "elect-issued MMA M64 N32, 4 accumulators",
         120: "MMA+commit+wait N32", 121: "MMA+commit+wait N128", 130: "st/sync/6xMMA/commit/wait/ld N32",
         131: "same N128", 140: "__syncthreads (256 thr)"}
for k, n in names.items():
    v = ctypes.c_double()
    rc = lib.m3_selftest(k, ctypes.byref(v))
    print(f"{k}: {n:40s} rc={rc} cycles/rep = {v.value:.1f}")
