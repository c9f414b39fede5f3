"""Builds the oracle's C backend (``oracle/_build/liboracle_c.so``).  Test infrastructure only.

The reference itself is pure Python over SciPy's compiled ``cKDTree`` (no C/C++ sources under
``/root/reference``), so there is nothing to compile into ``oracle/_ref``; the reference's own CPU
path is exercised through the oracle's ``"scipy"`` backend instead (same SciPy calls).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "oracle_c.c")
    out_dir = os.path.join(HERE, "_build")
    out = os.path.join(out_dir, "liboracle_c.so")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(out_dir, exist_ok=True)
    # -ffp-contract=off: the distance must be a plain rounded subtraction, nothing fused.
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-o", out, src, "-lm"]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force=True))
