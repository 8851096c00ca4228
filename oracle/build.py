"""Builds the oracle's C restatement (oracle/gaussel_c.c -> oracle/_build/libgaussel_c.so).

Test infrastructure only: called by `__graft_entry__.build()` and by `bench.py --impl reference` /
`bench.py`'s `cpu_baseline` leg.  Nothing here touches the product package (`cans_b200`)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gaussel_c.c")
LIB = os.path.join(HERE, "_build", "libgaussel_c.so")


def build_gaussel_c(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", LIB, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle: building gaussel_c.c failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_gaussel_c(True))
