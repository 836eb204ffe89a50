"""Builds the plain-C oracle (test infrastructure) with gcc into oracle/_build/.  Called by
__graft_entry__.build(); building the checker is not using it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liballreduce_oracle.so")


def build() -> str:
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "allreduce_oracle.c")
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    # -ffp-contract=off / no fast-math: the oracle's fp32 arithmetic must be plain IEEE
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off",
                           "-fno-fast-math", "-Wall", "-o", LIB, src, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build())
