"""Build libadept_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m adept_b200.build [--verbose] [--force]

The shared library lands next to this file (adept_b200/libadept_b200.so); objects go to adept_b200/csrc/build/.
"""

from __future__ import annotations

import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libadept_b200.so"
SOURCES = ["api.cu", "push.cu", "vdfdx_tma.cu", "vdfdx_dual.cu", "rowops.cu", "field.cu", "collide.cu", "step.cu", "backward.cu", "vrow.cu", "bluestein.cu"]
NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> Path:
    nvcc = _nvcc()
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "adept_b200.h"]
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str):
        obj = objdir / (src.replace(".cu", ".o"))
        if force or _stale(obj, [CSRC / src, *headers]):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
            if verbose:
                sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


def xla_include_dir():
    """Directory holding xla/ffi/api/ffi.h: $XLA_FFI_INCLUDE_DIR, else jax.ffi.include_dir() when jax is importable."""
    env = os.environ.get("XLA_FFI_INCLUDE_DIR")
    if env and (Path(env) / "xla" / "ffi" / "api" / "ffi.h").exists():
        return env
    try:
        import jax.ffi  # noqa: PLC0415  (absent from the B200 image; present on a maintainer's machine)

        return jax.ffi.include_dir()
    except Exception:
        return None


def build_xla() -> Path | None:
    """Layer 2: libadept_b200_xla.so with the XLA FFI handlers of csrc/xla_ffi.cc, built iff the XLA FFI
    headers are found; returns None (and builds nothing) otherwise."""
    inc = xla_include_dir()
    if inc is None:
        return None
    build()
    out = HERE / "libadept_b200_xla.so"
    cmd = [_nvcc(), "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-I", inc, str(CSRC / "xla_ffi.cc"), "-o",
           str(out), "-L", str(HERE), "-ladept_b200", "-Xlinker", "-rpath=$ORIGIN", "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"building the XLA FFI layer failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--xla", action="store_true", help="also build libadept_b200_xla.so (needs the XLA FFI headers)")
    a = ap.parse_args()
    print(build(a.verbose, a.force))
    if a.xla:
        print(build_xla() or "XLA FFI headers not found (set XLA_FFI_INCLUDE_DIR or install jax): layer 2 not built")
