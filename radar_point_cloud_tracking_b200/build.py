"""Build ``libradarb200.so`` (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc."""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libradarb200.so"
SOURCES = ["context.cu", "spoke.cu", "land.cu", "dbscan.cu", "fuse.cu", "synth.cu", "pipeline.cu", "csv.cu", "plyfmt.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-extended-lambda",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--fmad=false",            # no silent FMA contraction: the float32/float64 chains are bit-exact contracts
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; cannot build libradarb200.so")
    return cand


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "common.cuh", INCLUDE / "radarb200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-shared", *NVCC_FLAGS, "-I", str(INCLUDE), "-I", str(CSRC),
           "-o", str(LIB), *[str(CSRC / s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    (PKG / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libradarb200.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
