"""Builds the oracle's C chess engine (test infrastructure, see oracle/__init__.py)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "chess_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libchess_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        tmp = OUT + ".tmp.%d" % os.getpid()
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", tmp, SRC])
        os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
