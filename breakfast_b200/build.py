"""Build libbreakfast_b200.so (sm_100a only) in-tree with nvcc.

The shared library carries the C ABI of include/breakfast_b200.h; cudart is linked statically so
the library loads (and exports its symbols) on a box without a GPU driver, and fails loudly with
BF_ERR_NO_DEVICE on the first compute call there.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libbreakfast_b200.so"
SOURCES = [CSRC / "api.cu", CSRC / "peaks.cu"]
HEADERS = [CSRC / "kernels.cuh", PKG_DIR.parent / "include" / "breakfast_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-shared",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libbreakfast_b200.so")
    return exe


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)])


class _BuildLock:
    """Several processes (torchrun ranks, pytest-xdist workers) may find the library stale at the same time: one
    builds, the others wait and then see it fresh.  The compiler writes a temporary file that is renamed into place."""

    def __init__(self, target: Path):
        self.path = target.with_suffix(target.suffix + ".lock")

    def __enter__(self):
        import fcntl
        self.fh = open(self.path, "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()


def _compile(cmd: list, target: Path, what: str) -> None:
    tmp = target.with_name(f".{target.name}.{os.getpid()}.tmp")
    res = subprocess.run(cmd + ["-o", str(tmp)], capture_output=True, text=True)
    if res.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"{what} failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, target)
    return res


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    with _BuildLock(LIB_PATH):
        if not force and not needs_build():   # another process built it while we waited
            return LIB_PATH
        cmd = [_nvcc(), *NVCC_FLAGS]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += [str(s) for s in SOURCES]
        res = _compile(cmd, LIB_PATH, "nvcc")
        if verbose:
            sys.stderr.write(res.stdout + res.stderr)
    return LIB_PATH


SYNTH_SRC = CSRC / "synth.c"
SYNTH_LIB = PKG_DIR / "libbfsynth.so"


def build_synth(force: bool = False) -> Path:
    """gcc build of the synthetic-profile generator (benchmark/test input only)."""
    if not force and SYNTH_LIB.exists() and SYNTH_LIB.stat().st_mtime >= SYNTH_SRC.stat().st_mtime:
        return SYNTH_LIB
    gcc = shutil.which("gcc") or "gcc"
    with _BuildLock(SYNTH_LIB):
        if not force and SYNTH_LIB.exists() and SYNTH_LIB.stat().st_mtime >= SYNTH_SRC.stat().st_mtime:
            return SYNTH_LIB
        _compile([gcc, "-O2", "-fPIC", "-shared", "-std=c11", str(SYNTH_SRC), "-lm"], SYNTH_LIB, "gcc")
    return SYNTH_LIB


HOST_SRC = CSRC / "host_parse.cpp"
HOST_LIB = PKG_DIR / "libbfhost.so"


def build_host(force: bool = False) -> Path:
    """g++ build of the native host-side parser (tokenise / filter / dedup / CSR); no CUDA involved."""
    if not force and HOST_LIB.exists() and HOST_LIB.stat().st_mtime >= HOST_SRC.stat().st_mtime:
        return HOST_LIB
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found: cannot build libbfhost.so")
    with _BuildLock(HOST_LIB):
        if not force and HOST_LIB.exists() and HOST_LIB.stat().st_mtime >= HOST_SRC.stat().st_mtime:
            return HOST_LIB
        _compile([gxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-pthread", str(HOST_SRC)], HOST_LIB, "g++")
    return HOST_LIB


if __name__ == "__main__":
    build_host(force="--force" in sys.argv)
    build_synth(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
