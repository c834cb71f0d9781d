"""In-tree build of libm3b200.so (sm_100a only) with plain nvcc.

``python -m mimic3_b200.build`` or ``__graft_entry__.build()``.  Objects go to
``build/``, the library to ``mimic3_b200/libm3b200.so`` (git-ignored; it travels to
the GPU box with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OUT = PKG / "libm3b200.so"
OBJ = ROOT / "build" / "obj"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "-I", str(ROOT / "include"), "-I", str(CSRC)]
if os.environ.get("M3B200_KERNEL_PROFILE"):  # per-role cycle counters in the persistent kernels (debug builds only):
    COMMON.append("-DM3B200_KERNEL_PROFILE")   # a separate library (select it with M3B200_LIBRARY) and object directory
    OUT = PKG / "libm3b200_prof.so"
    OBJ = ROOT / "build" / "obj_prof"


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found")
    return exe


def sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cc")))


def _stamp(src: Path, flags) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def compile_one(src: Path, verbose: bool = False, extra=()) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    obj = OBJ / (src.name + ".o")
    flags = ARCH + COMMON + list(extra)
    if src.suffix == ".cu":
        flags = flags + ["-Xptxas", "-v"] if verbose else flags
    stamp = OBJ / (src.name + ".stamp")
    sig = _stamp(src, flags)
    if obj.exists() and stamp.exists() and stamp.read_text() == sig:
        return obj
    cmd = [nvcc()] + flags + ["-x", "cu" if src.suffix == ".cu" else "c++", "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose or r.stderr.strip():
        sys.stderr.write(r.stderr)
    stamp.write_text(sig)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    if force and OBJ.exists():
        shutil.rmtree(OBJ)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(lambda s: compile_one(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if OUT.exists() and OUT.stat().st_mtime >= newest and not force:
        return OUT
    cmd = [nvcc()] + ARCH + ["-shared", "-o", str(OUT)] + [str(o) for o in objs] + ["-Xlinker", "--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
