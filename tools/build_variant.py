#!/usr/bin/env python
"""Builds a VARIANT of libmcphylo_b200.so (extra nvcc flags, e.g. -DMCP_RING_LDGSTS=1) into build_exp/<tag>/ for
A/B runs on one GPU box: `MCPHYLO_B200_LIB=build_exp/<tag>/libmcphylo_b200.so python tools/ab_ring.py ...`.
Not part of the product build.

    python tools/build_variant.py ldgsts -DMCP_RING_LDGSTS=1
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcphylo_jl_b200 import build as B  # noqa: E402


def main():
    tag, flags = sys.argv[1], sys.argv[2:]
    out = os.path.join(ROOT, "build_exp", tag)
    os.makedirs(out, exist_ok=True)
    nvcc = B.nvcc_path()

    def cc(src):
        obj = os.path.join(out, os.path.splitext(src)[0] + ".o")
        r = subprocess.run([nvcc] + B.NVCC_FLAGS + flags + ["-c", "-o", obj, os.path.join(B.SRC_DIR, src)],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        objs = list(pool.map(cc, B.SOURCES))
    lib = os.path.join(out, "libmcphylo_b200.so")
    r = subprocess.run([nvcc] + B.LINK_FLAGS + ["-o", lib] + objs, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    print(lib)


if __name__ == "__main__":
    main()
