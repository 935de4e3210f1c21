"""ORACLE build recipe (test infrastructure): compiles oracle/c/dial_port.c twice with gcc —
REAL=double (validated against the NumPy oracle) and REAL=float (the timed fp32 CPU baseline) —
into oracle/_build/ (git-ignored, travels to the GPU box like the other in-tree .so files).
`/root/reference` holds no compilable source for this path (pure Python on un-vendored JAX/MJX),
so there is no oracle/_ref to build."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "dial_port.c")
HDR = os.path.join(os.path.dirname(HERE), "include", "dial_b200.h")
OUT = os.path.join(HERE, "_build")
FLAGS = ["-O3", "-march=x86-64-v2", "-fno-math-errno", "-std=c11", "-fopenmp", "-shared", "-fPIC"]


def lib_path(real: str) -> str:
    return os.path.join(OUT, f"libdial_port_{'f32' if real == 'float' else 'f64'}.so")


def build(force: bool = False) -> None:
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for f in (SRC, HDR):
        h.update(open(f, "rb").read())
    h.update(" ".join(FLAGS).encode())
    digest = h.hexdigest()
    for real in ("double", "float"):
        so = lib_path(real)
        side = so + ".sha256"
        if not force and os.path.exists(so) and os.path.exists(side) and open(side).read().strip() == digest:
            continue
        subprocess.check_call(["gcc"] + FLAGS + [f"-DREAL={real}", "-o", so, SRC, "-lm"])
        with open(side, "w") as f:
            f.write(digest)


if __name__ == "__main__":
    build()
