"""Build the C part of the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

`python -m oracle.build` or `oracle.build.build()` compiles nms_greedy.c into
oracle/_build/liboracle_nms.so with strict fp32 semantics (no contraction, no fast-math).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_nms.so")
SRC = os.path.join(HERE, "nms_greedy.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", SRC, "-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
