"""Build libtae_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

    python -m tae_b200.build [--force] [--verbose]

Each csrc/*.cu is compiled to an object (in parallel), then linked into tae_b200/libtae_b200.so.
"""
from __future__ import annotations

import argparse
import fcntl
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "csrc" / "build"
LIB_PATH = PKG_DIR / "libtae_b200.so"
INCLUDE_DIR = PKG_DIR.parent / "include"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(INCLUDE_DIR),
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE_DIR / "tae_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    # flags WITHOUT the checkout's absolute include path: the same sources must hash the same wherever the repo lives
    # (the GPU box runs a copy under another path and must not rebuild a library that is already current)
    h.update(" ".join(f for f in NVCC_FLAGS if f != str(INCLUDE_DIR)).encode())
    return h.hexdigest()


def fingerprint() -> str:
    return _fingerprint()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = BUILD_DIR / "fingerprint.txt"
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB_PATH
    if not Path(NVCC).exists():
        # no toolchain: a prebuilt library is only usable if it was built from these very sources — the loader
        # (tae_b200._lib.load) compares the fingerprint embedded in the .so and raises on a mismatch
        if LIB_PATH.exists():
            return LIB_PATH
        raise RuntimeError(f"nvcc not found at {NVCC} and {LIB_PATH} is missing")
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    # One builder at a time (torchrun starts one process per GPU, all of which import the package): the others block
    # on the lock, re-check the stamp and return.  The library is linked under a temporary name and renamed into
    # place, so a concurrent dlopen never sees a half-written file.
    with open(BUILD_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == fp:
                return LIB_PATH
            return _build_locked(fp, stamp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(fp: str, stamp: Path, verbose: bool) -> Path:
    def compile_one(src: Path) -> Path:
        obj = BUILD_DIR / (src.stem + ".o")
        cmd = [NVCC, *NVCC_FLAGS, f'-DTAE_SRC_FINGERPRINT="{fp}"', "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB_PATH.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [NVCC, "-shared", "-o", str(tmp), *[str(o) for o in objs], "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    stamp.write_text(fp)
    return LIB_PATH


def build_variant(name: str, defines: list[str], verbose: bool = False) -> Path:
    """Developer A/B builds: the same sources with extra -D flags, linked as tae_b200/libtae_b200.<name>.so (objects
    under csrc/build/<name>/).  `TAE_B200_LIB=<path>` makes `tae_b200._lib` load it instead of the default library."""
    out_dir = BUILD_DIR / name
    out_dir.mkdir(parents=True, exist_ok=True)
    lib = PKG_DIR / f"libtae_b200.{name}.so"
    fp = _fingerprint()

    def compile_one(src: Path) -> Path:
        obj = out_dir / (src.stem + ".o")
        cmd = [NVCC, *NVCC_FLAGS, f'-DTAE_SRC_FINGERPRINT="{fp}"', *[f"-D{d}" for d in defines], "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(" ".join(cmd), res.stderr, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    res = subprocess.run([NVCC, "-shared", "-o", str(lib), *[str(o) for o in objs], "-lcudart"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--variant", default=None, help="A/B build name: writes libtae_b200.<name>.so")
    ap.add_argument("-D", dest="defines", action="append", default=[], help="extra macro for a --variant build")
    a = ap.parse_args()
    if a.variant:
        print(build_variant(a.variant, a.defines, verbose=a.verbose))
    else:
        print(build(force=a.force, verbose=a.verbose))
    sys.exit(0)
