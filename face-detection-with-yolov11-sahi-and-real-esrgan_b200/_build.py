"""In-tree nvcc build of the C-ABI library (csrc/*.cu -> libfsd_b200.so, sm_100a only).

The built .so is git-ignored but travels to the GPU box with the gpurun snapshot, so nothing is JIT-built
under ~/.cache.  `build()` is incremental (mtime based) and safe to call from several processes.
"""
from __future__ import annotations

import fcntl
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libfsd_b200.so"
OBJ_DIR = PKG_DIR / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-fmad=false",  # parity: no implicit FMA contraction (kernels use explicit fmaf where a fused op is intended)
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the fsd_b200 CUDA library cannot be built (there is no CPU fallback)")
    return exe


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return max(p.stat().st_mtime for p in hdrs)


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    lib_m = LIB_PATH.stat().st_mtime
    if _deps_mtime() > lib_m:
        return True
    return any(s.stat().st_mtime > lib_m for s in sources())


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libfsd_b200.so next to this file."""
    OBJ_DIR.mkdir(exist_ok=True)
    with open(OBJ_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():
            return LIB_PATH
        nvcc = _nvcc()
        hdr_m = _deps_mtime()
        jobs = []
        for src in sources():
            obj = OBJ_DIR / (src.stem + ".o")
            if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m):
                cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
                if ptxas_info:
                    cmd[1:1] = ["-Xptxas", "-v"]
                jobs.append(cmd)

        def run(cmd):
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
            if verbose or ptxas_info:
                print(" ".join(cmd[-4:]), "\n", r.stderr)
            return r

        with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
            list(ex.map(run, jobs))
        objs = [str(OBJ_DIR / (s.stem + ".o")) for s in sources()]
        tmp = LIB_PATH.with_suffix(".so.tmp")
        link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *objs, "-ldl"]
        run(link)
        os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
