"""Build recipe for the CPU oracle (test infrastructure, not product code).

Compiles oracle/wr_oracle.c + oracle/wr_oracle_blend.c -> oracle/_build/libwr_oracle.so with gcc.  `-ffp-contract=off`
is REQUIRED: the raster contract is a sequence of individually rounded
fp32 operations and a fused multiply-add would change coverage at snap boundaries.

There is no `oracle/_ref`: the reference (/root/reference) contains no native source for this
path -- its arithmetic lives in the un-vendored nvdiffrast dependency (requirements.txt:55),
so nothing of the reference can be compiled here (recorded in DESIGN.md section 6).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "wr_oracle.c"), os.path.join(HERE, "wr_oracle_blend.c")]
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libwr_oracle.so")

CFLAGS = ["-O2", "-fPIC", "-shared", "-std=c11", "-fopenmp", "-ffp-contract=off",
          "-fno-fast-math", "-Wall", "-Wextra"]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in SRCS):
        return LIB
    cmd = ["gcc", *CFLAGS, *SRCS, "-o", LIB, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
