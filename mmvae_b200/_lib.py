"""Build and load ``libcmmvae_b200.so`` (the C-ABI library of include/cmmvae_b200.h).

The library is compiled in-tree with nvcc for sm_100a only and bound with ctypes.  There is no
fallback of any kind: if the library is missing or does not export a symbol, importing the ops
raises.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libcmmvae_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "cmmvae_b200.h")
ABI_VERSION = 2   # must equal CMMVAE_ABI_VERSION of the header the library was built from

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library (seconds per file; no GPU needed)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(CSRC, "build", os.path.basename(src) + ".o")
        os.makedirs(os.path.dirname(obj), exist_ok=True)
        objs.append(obj)
        cmd = [nvcc] + flags + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode()))
        if verbose and out:
            print(out.decode())
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def exported_symbols_in_header():
    """Names of every function include/cmmvae_b200.h declares."""
    import re

    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmmvae_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or library fallback for the cmmvae_b200 kernels)")
        lib = ctypes.CDLL(LIB_PATH)
        for name in exported_symbols_in_header():
            if not hasattr(lib, name):
                raise RuntimeError(f"libcmmvae_b200.so does not export {name}")
        lib.cmmvae_last_error.restype = ctypes.c_char_p
        lib.cmmvae_launch_count.restype = ctypes.c_longlong
        if lib.cmmvae_abi_version() != ABI_VERSION:
            raise RuntimeError("libcmmvae_b200.so ABI version mismatch")
        _lib = lib
    return _lib
