#!/usr/bin/env python
"""Compile the C oracle (TEST INFRASTRUCTURE) into oracle/liboracle.so with gcc.

-ffp-contract=off is part of the oracle's arithmetic contract (see splat_oracle.c header):
every fused multiply-add is spelled out with fmaf(); -mfma only makes fmaf() a single instruction.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "splat_oracle.c")
OUT = os.path.join(HERE, "liboracle.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-fopenmp", "-shared", "-fPIC",
           "-Wall", "-Wextra", "-o", OUT, SRC, "-lm"]
    print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
