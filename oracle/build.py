"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see gs_oracle.c header).

Produces oracle/_build/libgs_oracle_f32.so and libgs_oracle_f64.so with gcc.
`oracle/_ref/` (a build of the reference's own rasteriser) cannot exist: the reference
does not vendor that source, so there is nothing to compile.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "gs_oracle.c")

VARIANTS = {"f32": [], "f64": ["-DGSO_FP64"]}


def lib_path(variant: str) -> str:
    return os.path.join(OUT, f"libgs_oracle_{variant}.so")


def build(force: bool = False, verbose: bool = False) -> None:
    os.makedirs(OUT, exist_ok=True)
    for variant, defs in VARIANTS.items():
        out = lib_path(variant)
        if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC):
            continue
        cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11", "-Wall", "-Wextra",
               *defs, SRC, "-o", out, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
