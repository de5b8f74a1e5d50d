"""In-tree build of the sm_100a C-ABI library (`libsd_b200.so`) with nvcc.

The library is built next to this file so it travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).  nvcc cross-compiles for
sm_100a without a GPU.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libsd_b200.so"                 # fp16 operands (the parity build)
LIB_BF16 = HERE / "libsd_b200_bf16.so"       # same sources with -DSD_BF16 (bf16 operands)
SOURCES = ["sd_api.cu", "seg_kernels.cu", "unet.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing",
] + os.environ.get("SD_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; cannot build libsd_b200.so")


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "sd_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, dtype: str = "f16") -> Path:
    """Compiles csrc/*.cu for sm_100a and links the library of the given operand type (no-op if up to date)."""
    if dtype not in ("f16", "bf16"):
        raise ValueError(f"dtype {dtype!r}: f16 or bf16")
    lib = LIB if dtype == "f16" else LIB_BF16
    extra = [] if dtype == "f16" else ["-DSD_BF16"]
    bdir = CSRC / "build" / dtype
    stamp = bdir / "stamp"
    digest = _digest()
    if not force and lib.exists() and stamp.exists() and stamp.read_text() == digest:
        return lib
    nvcc = _nvcc()
    bdir.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = bdir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out, file=sys.stderr)
    cmd = [nvcc, "-shared", "-o", str(lib), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    stamp.write_text(digest)
    return lib


if __name__ == "__main__":
    for dt in ("f16", "bf16"):
        print(build(force="--force" in sys.argv, verbose=True, dtype=dt))
