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

print("\nMMA stream as the conv kernels issue it (elect-issued, 4 accumulators, 512 MMAs); variant bits: 1 A start +16 B, "
      "2 odd pitch (133 rows), 4 four A start rows, 8 four B blocks, 16 LSU traffic (4 warps STS/LDS.128), "
      "32 8 KB bulk copies into smem, 64 A start +64 B")
for n, base in ((32, 200), (64, 300), (128, 400)):
    for var in (0, 1, 64, 2, 3, 4, 6, 14, 16, 32, 48, 15 + 16, 15 + 32, 15 + 48):
        v = ctypes.c_double()
        rc = lib.m3_selftest(base + var, ctypes.byref(v))
        print(f"N={n:3d} variant {var:2d} ({var:06b}): rc={rc} cycles/MMA = {v.value:.1f}")

print("\nN=64 MMA stream with a tcgen05.commit every P MMAs (5xx), plus a completed-mbarrier wait + fence per stage (6xx)")
for base in (500, 600):
    for period in (1, 2, 4, 6, 12, 24, 60):
        v = ctypes.c_double()
        rc = lib.m3_selftest(base + period, ctypes.byref(v))
        print(f"{base + period}: period {period:2d}: rc={rc} cycles/MMA = {v.value:.1f}")

print("\n7 warps x (3 tcgen05.ld x16 per round): 700 wait after each load, 701 one wait per round, 702/703 the same under an N=32 MMA stream")
for k in (700, 701, 702, 703):
    v = ctypes.c_double()
    rc = lib.m3_selftest(k, ctypes.byref(v))
    print(f"{k}: rc={rc} cycles/round = {v.value:.1f}")
