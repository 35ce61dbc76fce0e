"""Builds libmon_b200.so (the CUDA core behind include/mon_c.h) in-tree with nvcc for sm_100a.

    python -m ro_map_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so travels with the tree to the GPU box.
Objects are cached under ro_map_b200/_build/ and rebuilt when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libmon_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
              "-Xptxas", "-v"]
if (CSRC / "kernels_mlp_tc.cu").exists():
    NVCC_FLAGS.append("-DMON_HAVE_TC")
NVCC_FLAGS += os.environ.get("MON_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the B200 core cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers() -> list[Path]:
    return sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "mon_c.h"]


def _compile(src: Path, verbose: bool) -> tuple[Path, str]:
    obj = BUILD / (src.stem + ".o")
    cmd = [_nvcc(), *ARCH, *NVCC_FLAGS, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    (BUILD / (src.stem + ".ptxas.log")).write_text(log)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{log}")
    if verbose:
        print(log)
    return obj, log


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    srcs = sources()
    newest_header = max(p.stat().st_mtime for p in _headers())
    todo, objs = [], []
    for s in srcs:
        obj = BUILD / (s.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(s.stat().st_mtime, newest_header):
            todo.append(s)
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [_nvcc(), *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    return LIB


HOST = PKG / "host"
HOST_LIB = PKG / "libMON.so"
HOST_BIN = PKG / "offline_nerf"
REPLAY_BIN = PKG / "online_replay"


def build_host(force: bool = False) -> Path:
    """libMON.so: the C++ facade (nerf::NerfManagerOffline/Online, NeRF) over the C ABI, plus the headless
    offline_nerf driver.  Plain g++: no CUDA, Eigen or OpenCV needed (mon_compat.h falls back to shims)."""
    build()
    srcs = [HOST / "nerf_host.cpp"]
    deps = srcs + sorted(HOST.glob("*.h")) + [ROOT / "include" / "mon_c.h"]
    newest = max(p.stat().st_mtime for p in deps)
    cxx = os.environ.get("CXX", "g++")
    common = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-pthread", "-I", str(ROOT / "include"), "-I", str(HOST)]
    if force or not HOST_LIB.exists() or HOST_LIB.stat().st_mtime < newest:
        cmd = [cxx, *common, "-shared", "-o", str(HOST_LIB), *map(str, srcs), "-L", str(PKG), "-lmon_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("libMON.so build failed:\n" + proc.stdout + proc.stderr)
    for main_src, binary in ((HOST / "offline_nerf.cpp", HOST_BIN), (HOST / "online_replay.cpp", REPLAY_BIN)):
        if force or not binary.exists() or binary.stat().st_mtime < max(newest, main_src.stat().st_mtime):
            cmd = [cxx, *common, "-o", str(binary), str(main_src), "-L", str(PKG), "-lMON", "-lmon_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                raise RuntimeError(f"{binary.name} build failed:\n" + proc.stdout + proc.stderr)
    return HOST_LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
    print(build_host(force="--force" in sys.argv))
