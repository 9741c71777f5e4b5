#!/usr/bin/env python
"""Host-to-device bandwidth of the alignment upload: contiguous 1-D copies vs pitched 2-D copies
(cudaMemcpy2DAsync, source pitch = full alignment width) vs one 1-D copy per leaf row, from pinned memory.
Decides how mcp_eval_streamed moves a site block of an (n_leaves, S) row-major host array."""
import ctypes
import json
import sys

import torch

rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name)
        break
    except OSError:
        pass
if rt is None:
    import glob, os
    import nvidia.cuda_runtime
    rt = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(nvidia.cuda_runtime.__file__), "lib", "libcudart.so*"))[0])
vp, sz = ctypes.c_void_p, ctypes.c_size_t
rt.cudaMemcpy2DAsync.argtypes = [vp, sz, vp, sz, sz, sz, ctypes.c_int, vp]
rt.cudaMemcpyAsync.argtypes = [vp, vp, sz, ctypes.c_int, vp]


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    n_leaves, S = 1000, 1_000_000
    host = torch.empty((n_leaves, S), dtype=torch.uint8).pin_memory()
    host.fill_(1)
    st = vp(torch.cuda.current_stream().cuda_stream)
    out = []
    for width in (37888, 151552, 500000, 1000000):
        stride = (width + 1023) // 1024 * 1024
        dev = torch.empty(n_leaves * stride + 1024, dtype=torch.uint8, device="cuda")
        contig = torch.empty((n_leaves, width), dtype=torch.uint8).pin_memory()
        nbytes = n_leaves * width

        def pitched():
            assert rt.cudaMemcpy2DAsync(vp(dev.data_ptr()), stride, vp(host.data_ptr()), S, width, n_leaves, 1, st) == 0

        def contiguous_src_2d():
            assert rt.cudaMemcpy2DAsync(vp(dev.data_ptr()), stride, vp(contig.data_ptr()), width, width, n_leaves, 1, st) == 0

        def flat_1d():
            assert rt.cudaMemcpyAsync(vp(dev.data_ptr()), vp(contig.data_ptr()), nbytes, 1, st) == 0

        def per_row():
            for r in range(n_leaves):
                rt.cudaMemcpyAsync(vp(dev.data_ptr() + r * stride), vp(host.data_ptr() + r * S), width, 1, st)

        row = {"width": width}
        for name, fn in (("pitched_src_pitch_S", pitched), ("pitched_src_contiguous", contiguous_src_2d),
                         ("flat_1d", flat_1d), ("one_copy_per_row", per_row)):
            ms = timed(fn)
            row[name] = {"ms": ms, "GBps": nbytes / ms / 1e6}
        out.append(row)
        print(row, file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
