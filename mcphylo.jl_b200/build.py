"""In-tree build of libmcphylo_b200.so for sm_100a (nvcc cross-compiles without a GPU).

The library is several translation units -- the host side (C ABI, planner, multi-GPU group) and one
unit per state count K for the walk kernels -- compiled in parallel to objects under lib/obj/ and linked
with the CUDA runtime statically.  Only stale objects are recompiled: an edit of the host side does not
rebuild any kernel."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libmcphylo_b200.so")
HEADER = os.path.join(_HERE, "..", "include", "mcphylo_b200.h")

_COMMON = ["schedule.hpp", "device_layout.cuh", "smem_layout.cuh", "kernel_api.hpp"]
_KERNEL = _COMMON + ["walk_inst.cuh", "model_const.cuh", "device_math.cuh", "kernel_walk.cuh", "epilogue_prior.cuh",
                     "kernel_levels.cuh"]
# translation unit -> the headers it includes
UNITS = {
    "mcphylo_b200.cu": _COMMON + ["host_state.hpp", "planner.hpp", "nccl_dyn.hpp", "kernel_tables.cuh",
                                  "epilogue_prior.cuh", "kernel_finalize.cuh", HEADER],
    "walk_k2.cu": _KERNEL, "walk_k3.cu": _KERNEL, "walk_k4.cu": _KERNEL, "walk_k5.cu": _KERNEL, "walk_k6.cu": _KERNEL,
    "walk_generic.cu": _COMMON + ["model_const.cuh", "device_math.cuh", "kernel_generic.cuh", "kernel_mma.cuh"],
}
SOURCES = list(UNITS)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-cudart", "static",
              "-Xlinker", "--no-undefined", "-ldl", "-lpthread"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libmcphylo_b200.so")


def _path(name: str) -> str:
    return name if os.path.isabs(name) else os.path.join(SRC_DIR, name)


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")


def _unit_stale(src: str) -> bool:
    obj = _obj(src)
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(_path(d)) > t for d in [src] + UNITS[src])


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_unit_stale(s) or os.path.getmtime(_obj(s)) > t for s in SOURCES)


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = nvcc_path()
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)

    def compile_unit(src: str):
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", _obj(src), _path(src)]
        return src, subprocess.run(cmd, capture_output=True, text=True, env=env)

    todo = [s for s in SOURCES if force or _unit_stale(s)]
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as pool:
        for src, res in pool.map(compile_unit, todo):
            if verbose:
                print(f"==== {src}\n{res.stderr}")
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
    res = subprocess.run([nvcc] + LINK_FLAGS + ["-o", LIB] + [_obj(s) for s in SOURCES],
                         capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB
