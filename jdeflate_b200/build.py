"""Build recipes for the B200-native jdeflate library.

``build_cuda()``   nvcc (sm_100a, -lineinfo) + gcc -std=c99  ->  jdeflate_b200/lib/libjdeflate.so
``build_oracle()`` make -C oracle  (CPU restatement, helper tools and, when
                   /root/reference exists, the compiled reference) -> oracle/_ref/
``build_emu()``    TEST INFRASTRUCTURE: the same kernel sources compiled with g++
                   against tests/simt/simt_emu.h -> tests/simt/_build/libjdeflate_emu.so

Everything is built in-tree so the shared objects travel with the repository
snapshot to the GPU box; nothing is installed into site-packages.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "jdeflate_b200"
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
INCLUDE = ROOT / "include"
EMU_DIR = ROOT / "tests" / "simt"
EMU_OUT = EMU_DIR / "_build"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CC = os.environ.get("CC", "gcc")
CXX = os.environ.get("CXX", "g++")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
] + os.environ.get("JDB_NVCC_EXTRA", "").split()
HOST_CFLAGS = ["-std=c99", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra",
               "-Wno-unused-parameter", "-D_POSIX_C_SOURCE=200809L"]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, log: Path | None = None):
    res = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if log is not None:
        log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(" ".join(str(c) for c in cmd) + "\n" + res.stdout + res.stderr)
        raise RuntimeError(f"command failed: {cmd[0]} ... {cmd[-1]}")
    return res


def _headers():
    return (list((CSRC / "device").glob("*.h")) + list((CSRC / "device").glob("*.cuh")) +
            list((CSRC / "host").glob("*.h")) + list(INCLUDE.rglob("*.h")))


def device_sources():
    return sorted((CSRC / "device").glob("*.cu"))


def host_sources():
    return sorted((CSRC / "host").glob("*.c"))


def lib_path() -> Path:
    return LIBDIR / "libjdeflate.so"


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA translation unit for sm_100a and link libjdeflate.so."""
    OBJDIR.mkdir(parents=True, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in device_sources():
        obj = OBJDIR / (src.stem + ".cu.o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append(([NVCC, *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC / "device",
                          "-c", src, "-o", obj], OBJDIR / (src.stem + ".ptxas.log")))
    for src in host_sources():
        obj = OBJDIR / (src.stem + ".c.o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append(([CC, *HOST_CFLAGS, "-I", INCLUDE, "-I", CSRC / "host",
                          "-c", src, "-o", obj], None))
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(lambda j: _run(j[0], j[1]), jobs))
    out = lib_path()
    if force or jobs or _newer(out, objs):
        _run([NVCC, "-shared", "-o", out, *objs, "-Xlinker", "-Bsymbolic",
              "-Xlinker", "--exclude-libs,ALL", "-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    if verbose:
        for log in sorted(OBJDIR.glob("*.ptxas.log")):
            sys.stdout.write(log.read_text())
    return out


def build_oracle(verbose: bool = False) -> Path:
    res = subprocess.run(["make", "-C", str(ROOT / "oracle"), "all"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("oracle build failed")
    if verbose:
        sys.stdout.write(res.stdout)
    return ROOT / "oracle" / "_ref"


def emu_lib_path() -> Path:
    return EMU_OUT / "libjdeflate_emu.so"


def build_emu(force: bool = False, sanitize: bool = False) -> Path:
    """TEST INFRASTRUCTURE: kernels compiled for the CPU SIMT emulator."""
    EMU_OUT.mkdir(parents=True, exist_ok=True)
    hdrs = _headers() + [EMU_DIR / "simt_emu.h"]
    san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"] if sanitize else []
    cxxflags = ["-std=c++17", "-O1", "-g", "-fPIC", "-DJDB_SIMT_EMU", "-Wall", "-Wno-unused-function",
                "-Wno-unknown-pragmas", "-Wno-unused-variable", "-Wno-sign-compare",
                "-Wno-unused-but-set-variable", *san,
                "-I", EMU_DIR, "-I", INCLUDE, "-I", CSRC / "device",
                *os.environ.get("JDB_EMU_EXTRA", "").split()]
    jobs, objs = [], []
    for src in device_sources():
        if src.name == "runtime.cu":
            continue
        obj = EMU_OUT / (src.stem + ".emu.o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append(([CXX, *cxxflags, "-x", "c++", "-c", src, "-o", obj], None))
    for src in sorted(EMU_DIR.glob("*.cpp")):
        obj = EMU_OUT / (src.stem + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append(([CXX, *cxxflags, "-c", src, "-o", obj], None))
    for src in host_sources():
        obj = EMU_OUT / (src.stem + ".c.o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append(([CC, *HOST_CFLAGS, "-g", *san, "-I", INCLUDE, "-I", CSRC / "host",
                          "-c", src, "-o", obj], None))
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(lambda j: _run(j[0], j[1]), jobs))
    out = emu_lib_path()
    if force or jobs or _newer(out, objs):
        _run([CXX, "-shared", "-o", out, *objs, *san, "-Wl,-Bsymbolic", "-lpthread"])
    return out


if __name__ == "__main__":
    what = sys.argv[1:] or ["cuda", "oracle"]
    if "cuda" in what:
        print(build_cuda(force="--force" in what, verbose="-v" in what))
    if "oracle" in what:
        print(build_oracle())
    if "emu" in what:
        print(build_emu(force="--force" in what, sanitize="--asan" in what))
