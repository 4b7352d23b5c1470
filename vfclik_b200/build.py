"""Build ``libvfk.so`` in-tree with nvcc for sm_100a (``python -m vfclik_b200.build``)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvfk.so")
SOURCES = [os.path.join(CSRC, "vfk_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("vfk_kernels.cuh", "vfk_math.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "vfk.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]   # cudart linked statically (nvcc default)


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), out: str = OUT) -> str:
    if not force and out == OUT and up_to_date():
        return OUT
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (%d): %s" % (res.returncode, " ".join(cmd)))
    return out


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[len("--out="):] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build_library(force="--force" in sys.argv or bool(extra) or bool(outs), verbose="-v" in sys.argv, extra_flags=extra,
                        out=outs[0] if outs else OUT))
