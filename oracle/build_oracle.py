#!/usr/bin/env python3
"""Compiles the plain-C oracle (oracle/oracle.c) into oracle/liboracle.so.  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle.c")
OUT = os.path.join(HERE, "liboracle.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", SRC, "-o", OUT, "-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
