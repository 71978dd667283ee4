"""Build recipe for the oracle's C restatement (test infrastructure, see oracle/__init__.py).

`oracle/_ref/` (the real reference compiled from its own sources) does not exist for this
project: the reference repo is pure Python and the native kernels it runs live in the
torchvision wheel, whose sources are not on disk. The real reference
arithmetic is instead reachable by importing torchvision's CPU ops, which both the build
container and the GPU box can do; tests/test_oracle_pin.py pins this restatement to them.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "oracle_kernels.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
