"""Builds libagf_b200.so (sm_100a only) in-tree with nvcc.  No torch, no JIT cache."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent
SOURCES = ["common.cu", "gram.cu", "apply.cu", "pairs.cu", "featgram.cu", "synth.cu", "augment.cu", "mapval.cu", "peak.cu", "qp.cu", "peer.cu", "gram_i8.cu", "gram_i8t.cu", "apply_i8.cu"]
HEADERS = ["common.cuh", "i8.cuh", "i8_digits.cuh", "frame_pipe.cuh", "panel.cuh", "philox.cuh", "../../include/agf_b200.h"]
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
    """Compile every source to an object file (in parallel) and link the shared library."""
    from concurrent.futures import ThreadPoolExecutor

    fp = _fingerprint()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == fp:
        return LIB
    nvcc = _nvcc()
    objdir = CSRC / "_build"
    objdir.mkdir(exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "--shared"]

    def compile_one(name: str):
        obj = objdir / (Path(name).stem + ".o")
        cmd = [nvcc, *compile_flags, "-c", "-o", str(obj), str(CSRC / name)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        return name, obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = [(n, r) for n, _, r in results if r.returncode != 0]
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(f"[{n}]\n{r.stdout}{r.stderr}" for n, r in failed))
    if verbose:
        for _, _, r in results:
            print(r.stderr)
    link = subprocess.run([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                           "-o", str(LIB), *[str(o) for _, o, _ in results], "-lcudart"],
                          capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + link.stdout + link.stderr)
    STAMP.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
