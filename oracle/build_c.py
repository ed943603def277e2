"""gcc build of the plain-C part of the oracle (oracle/qil_oracle_c.c) -> oracle/build/libqil_oracle_c.so.
Test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "qil_oracle_c.c")
OUT = os.path.join(HERE, "build", "libqil_oracle_c.so")


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-shared", "-fPIC", "-o", OUT, SRC], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
