"""Builds libagf_b200.so (sm_100a only) in-tree with nvcc.  No torch, no JIT cache."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent
SOURCES = ["common.cu", "gram.cu", "apply.cu", "pairs.cu", "featgram.cu", "synth.cu", "augment.cu"]
HEADERS = ["common.cuh", "frame_pipe.cuh", "panel.cuh", "philox.cuh", "../../include/agf_b200.h"]
LIB = CSRC / "libagf_b200.so"
STAMP = CSRC / ".libagf_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libagf_b200.so cannot be built")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    fp = _fingerprint()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == fp:
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB), *[str(CSRC / s) for s in SOURCES], "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    STAMP.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
