"""Build the C oracle (``oracle/liboracle_stdbscan.so``). Test infrastructure only.

The reference's hot path is pure Python (numpy / scikit-learn) — there is no C/C++ reference
source to compile, so there is no ``oracle/_ref``; the pin against the real reference is the
golden fixtures under ``tests/golden/`` (made by ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "stdbscan_ref.c"
LIB = HERE / "liboracle_stdbscan.so"


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", str(LIB), str(SRC), "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
