"""In-tree build of libb2retr.so (hand-written sm_100a CUDA behind a C ABI).

`python -m movie_recommender_demo_b200.build [--force] [--ptxas-v]` or `build()` from
`__graft_entry__.py`.  nvcc cross-compiles without a GPU; the .so is git-ignored but travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libb2retr.so"

SOURCES = ["scan_tc.cu", "scan_pair.cu", "ingest.cu", "select.cu", "index.cu", "tower.cu", "tower_mlp.cu", "tower_fused.cu", "ranker.cu", "ivf.cu", "ivfpq.cu", "persist.cu", "peer.cu"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (needed to build libb2retr.so for sm_100a)")


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "b2retr.h"]
    return max(p.stat().st_mtime for p in hdrs if p.exists())


def build(force: bool = False, verbose: bool = False, ptxas_v: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libb2retr.so. Returns its path."""
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    hdr_m = _deps_mtime()
    # The objects do not travel to the GPU box (.gpurunignore) but the library does: a library that is newer than
    # every source and header is up to date whether or not its objects are still around.
    if (not force and not ptxas_v and LIB_PATH.exists()
            and LIB_PATH.stat().st_mtime >= max([hdr_m] + [s.stat().st_mtime for s in srcs])):
        return LIB_PATH
    jobs = []
    objs = []
    for src in srcs:
        obj = OBJ_DIR / (src.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if ptxas_v:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or ptxas_v or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    need_link = force or bool(jobs) or not LIB_PATH.exists() or any(
        o.stat().st_mtime > LIB_PATH.stat().st_mtime for o in objs)
    if need_link:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
               *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("link of libb2retr.so failed")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_v="--ptxas-v" in sys.argv)
    print(p)
