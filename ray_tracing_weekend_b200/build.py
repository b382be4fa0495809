"""Build librtw_cuda.so (the C-ABI library of include/rtw.h) in-tree with nvcc for sm_100a.

Five translation units: kernels_f32.cu (fast path; -fmad=false, its fused multiply-adds are explicit fmaf calls so that every
FP32 kernel rounds alike), kernels_f64.cu (reference-exact path, -fmad=false), bvh_device.cu (device LBVH), capi.cu (extern "C"
surface + host BVH builder), host/rtw_host_capi.cpp (host mirror helpers).  cudart is linked statically so the
.so loads next to torch's own runtime without LD_LIBRARY_PATH games.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "_obj")
LIB_PATH = os.path.join(LIB_DIR, "librtw_cuda.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
UNITS = [
    ("kernels_f32.cu", ["-fmad=false"]),    # FMAs are explicit (fmaf) so every kernel rounds alike
    ("kernels_f64.cu", ["-fmad=false"]),
    ("bvh_device.cu", ["-fmad=false"]),     # device-side LBVH construction (CUB radix sort + hand-written tree kernels)
    ("capi.cu", []),
    ("../host/rtw_host_capi.cpp", []),
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build librtw_cuda.so")
    return exe


def _sources():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    host = os.path.join(HERE, "host")
    out += [os.path.join(host, f) for f in os.listdir(host)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "rtw_host.h"))
    out.append(os.path.join(os.path.dirname(HERE), "include", "rtw.h"))
    out.append(os.path.abspath(__file__))
    return out


HASH_PATH = os.path.join(LIB_DIR, ".source_hash")


def source_hash() -> str:
    """Content hash of every source and of the build recipe.  Staleness is decided by content, not by mtime:
    the built library travels to the GPU box inside a repo snapshot whose mtimes mean nothing, and N torchrun
    ranks importing the package at once must not all decide to rebuild."""
    h = hashlib.sha256()
    for p in sorted(_sources()):
        h.update(os.path.relpath(p, HERE).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    # one builder at a time (torchrun ranks, pytest-xdist workers); the others wait and then find it fresh
    with open(os.path.join(LIB_DIR, ".build_lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():
            return LIB_PATH
        return _build_locked(verbose)


def build_variant(name: str, defines=(), f32_flags=None, verbose: bool = False) -> str:
    """Tuning build (scripts/variant_bench.py): the same library with extra -D defines on every unit and, optionally, other flags for
    kernels_f32.cu, written to lib/variants/librtw_cuda_<name>.so.  Load it with RTW_LIBRARY=<path>."""
    out_dir = os.path.join(LIB_DIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    units = [(src, (list(f32_flags) if (f32_flags is not None and src == "kernels_f32.cu") else list(extra)) + [f"-D{d}" for d in defines])
             for src, extra in UNITS]
    return _compile_and_link(units, os.path.join(OBJ_DIR, "variant_" + name), os.path.join(out_dir, f"librtw_cuda_{name}.so"), verbose)


def _build_locked(verbose: bool) -> str:
    _compile_and_link(UNITS, OBJ_DIR, LIB_PATH, verbose)
    # the reference's `bin` with --backend cuda (host mirror CLI)
    cli = [shutil.which("g++") or "g++", "-O2", "-std=c++17", os.path.join(HERE, "host", "rtw_bin.cpp"), "-o",
           os.path.join(LIB_DIR, "rtw_bin"), "-L" + LIB_DIR, "-lrtw_cuda", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cli, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("rtw_bin build failed: " + " ".join(cli) + "\n" + r.stdout)
    with open(HASH_PATH, "w") as f:
        f.write(source_hash())
    return LIB_PATH


def _compile_and_link(units, obj_dir: str, lib_path: str, verbose: bool) -> str:
    nvcc = _nvcc()
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src, extra in units:
        obj = os.path.join(obj_dir, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out, file=sys.stderr)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + (out or ""))
    link = [nvcc, *ARCH, "-shared", "-cudart", "static", "-o", lib_path, *objs, "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: " + " ".join(link) + "\n" + r.stdout)
    return lib_path


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
