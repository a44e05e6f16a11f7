"""TEST INFRASTRUCTURE ONLY -- builds oracle/_build/librrl_oracle.so from oracle/rrl_oracle.c.

The reference (pure Python/PyTorch) has no compilable sources, so there is no oracle/_ref/
build; the reference itself is imported in the build container by oracle/ref_loader.py.
Flags: -ffp-contract=off keeps every product and sum separately rounded (the reference's
ATen kernels do not fuse), -fopenmp parallelises over lines.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rrl_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "librrl_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
           "-fopenmp", "-Wall", "-Wextra", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
