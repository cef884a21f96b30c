"""In-tree build of libvitmarl_b200.so (hand-written sm_100a CUDA + C ABI) with nvcc.

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.
nvcc cross-compiles without a GPU, so `build()` is also the CPU-side "does it build" check.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libvitmarl_b200.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA path cannot be built (there is no CPU fallback)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link the C-ABI shared library."""
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(_HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(_HERE, "..", "include", "*.h"))
    nvcc = _nvcc()
    objs, logs, procs = [], [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, *os.environ.get("VITMARL_NVCC_EXTRA", "").split(), "-c", src, "-o", obj]
            procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, cmd, p in procs:
        out, _ = p.communicate()
        logs.append(out)
        if verbose:
            print(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed for {src}: {' '.join(cmd)}")
    if procs:
        with open(os.path.join(obj_dir, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-cudart", "static"]
        subprocess.check_call(cmd)
    return LIB_PATH


def xla_ffi_include_dir():
    """Directory holding xla/ffi/api/ffi.h when JAX / jaxlib is installed, else None (this image: None)."""
    try:
        import jax.ffi
        return jax.ffi.include_dir()
    except Exception:
        return None


def build_ffi(include_dir: str = None) -> str:
    """Compile csrc/ffi/vitmarl_ffi.cc (the XLA FFI handlers) against the library; only possible where the XLA FFI headers
    exist.  Returns the .so path, or raises if the headers are missing."""
    include_dir = include_dir or xla_ffi_include_dir()
    if not include_dir or not os.path.exists(os.path.join(include_dir, "xla", "ffi", "api", "ffi.h")):
        raise RuntimeError("XLA FFI headers not found (JAX is not installed): the jax.ffi shim cannot be built here")
    build()
    out = os.path.join(LIB_DIR, "libvitmarl_ffi.so")
    src = os.path.join(CSRC, "ffi", "vitmarl_ffi.cc")
    cmd = [_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", include_dir, src, "-o", out,
           "-L", LIB_DIR, "-lvitmarl_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
