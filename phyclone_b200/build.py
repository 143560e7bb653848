"""Compile the CUDA extension in-tree:  python -m phyclone_b200.build"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "pcb_api.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", f) for f in ("pcb_kernels.cuh", "pcb_device.cuh", "pcb_sweep.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "phyclone_b200.h")
]
OUT_DIR = os.path.join(HERE, "lib")
OUT = os.path.join(OUT_DIR, "libphyclone_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


HOST_SRC = os.path.join(HERE, "csrc", "pcb_host.cpp")
HOST_OUT = os.path.join(OUT_DIR, "libpcb_host.so")


def build_host(force=False):
    """Host-side sampling draws (csrc/pcb_host.cpp): plain g++, linked against numpy's own libnpyrandom.a."""
    if not force and os.path.exists(HOST_OUT) and os.path.getmtime(HOST_OUT) >= os.path.getmtime(HOST_SRC):
        return HOST_OUT
    import sysconfig

    import numpy

    np_dir = os.path.dirname(numpy.__file__)
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared",
           "-I" + os.path.join(np_dir, "_core", "include"), "-I" + sysconfig.get_paths()["include"],
           "-o", HOST_OUT, HOST_SRC, os.path.join(np_dir, "random", "lib", "libnpyrandom.a"), "-lm"]
    os.makedirs(OUT_DIR, exist_ok=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building %s" % HOST_OUT)
    return HOST_OUT


MOD_SRC = os.path.join(HERE, "csrc", "pcb_hostmod.cpp")


def build_hostmod(force=False):
    """`_pcbhost` CPython extension (tile-store bookkeeping): plain g++ against Python.h."""
    import sysconfig

    out = os.path.join(OUT_DIR, "_pcbhost" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(MOD_SRC):
        return out
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I" + sysconfig.get_paths()["include"], "-o", out, MOD_SRC]
    os.makedirs(OUT_DIR, exist_ok=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building %s" % out)
    return out


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    build_host(force)
    build_hostmod(force)
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building %s" % OUT)
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
