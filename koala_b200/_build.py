"""Build recipe for libpv_koala_b200.so (nvcc, sm_100a only, in-tree so the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libpv_koala_b200.so")
SOURCES = ["engine.cu", "koala_abi.cu"]
HEADERS = ["exports.map", "engine.h", "koala_common.cuh", "stft_kernels.cuh", "masknet_fp32.cuh", "tcgen05_common.cuh", "masknet_fused.cuh", "masknet_i8.cuh",
           os.path.join("..", "..", "include", "pv_koala_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-cudart", "static", "--shared",
    "-Xlinker", "--version-script=" + os.path.join(CSRC, "exports.map"),
]


def _nvcc() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB_PATH) -> str:
    """Compile every CUDA source into koala_b200/lib/libpv_koala_b200.so.  Cross-compiles without a GPU.
    `defines` / `out` build a tuning variant (-DNAME=VALUE ...) next to it (tools/variant_bench.py)."""
    if not force and not defines and out == LIB_PATH and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else LIB_PATH))
