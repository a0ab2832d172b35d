"""In-tree build of libsbir_b200.so (sm_100a only) with plain nvcc — no torch extension machinery.

`python -m art_sbir_b200._build` or `art_sbir_b200._build.build()`; the .so lands in
art_sbir_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import concurrent.futures
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
BUILD_DIR = PKG / "lib" / "obj"
LIB_PATH = LIB_DIR / "libsbir_b200.so"
SOURCES = ["dist_topk_bf16_euclidean.cu", "dist_topk_bf16_cosine.cu", "dist_topk_f32_euclidean.cu", "dist_topk_f32_cosine.cu",
           "api.cu", "rowops.cu", "dist_topk.cu", "finalize.cu", "batch_hard.cu", "host_path.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _flags() -> list:
    """SBIR_BUILD_DIAG=1 at BUILD time compiles the diagnostic switches / cycle counters of K1 in (-DSBIR_DIAG)."""
    return NVCC_FLAGS + (["-DSBIR_DIAG"] if os.environ.get("SBIR_BUILD_DIAG") == "1" else [])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the sbir_b200 CUDA library cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*")) + [PKG.parent / "include" / "sbir_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path."""
    LIB_DIR.mkdir(exist_ok=True)
    BUILD_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "build.sha256"
    digest = _digest()

    def fresh() -> bool:
        return LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == digest

    if not force and fresh():
        return LIB_PATH
    # Several ranks of one torchrun job may get here at once: one builds, the others wait on the
    # lock and then find the library fresh.  The .so is linked under a temporary name and renamed
    # into place, so a reader never sees a half-written file.
    with open(LIB_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return LIB_PATH
            return _build_locked(digest, stamp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(digest: str, stamp: Path, verbose: bool) -> Path:
    nvcc = _nvcc()

    def compile_one(src: str) -> tuple[str, str]:
        obj = BUILD_DIR / (src + ".o")
        cmd = [nvcc, *_flags(), "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return str(obj), r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), max(4, os.cpu_count() or 4))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    (LIB_DIR / "ptxas.log").write_text("\n".join(log for _, log in results))
    tmp = LIB_DIR / f".libsbir_b200.{os.getpid()}.so"
    link = [nvcc, "-shared", "-o", str(tmp), *[o for o, _ in results],
            "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    stamp.write_text(digest)
    if verbose:
        print(f"built {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
