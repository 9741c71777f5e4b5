"""The C ABI used from plain C (examples/c_abi_example.c: dlopen, no Python, no torch), checked
against the oracle on the same 4-taxon JC problem."""
import os
import subprocess

import numpy as np
import pytest

import mcphylo_jl_b200 as mcp
from mcphylo_jl_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_example")
    subprocess.run(["/usr/bin/gcc", "-O2", "-o", exe, os.path.join(ROOT, "examples", "c_abi_example.c"), "-ldl", "-lm"],
                   check=True)
    return exe


def test_c_example_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    res = subprocess.run([_build(tmp_path), capi.LIB_PATH], capture_output=True, text=True)
    assert res.returncode == 1 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_c_example_matches_oracle(tmp_path, oracle):
    res = subprocess.run([_build(tmp_path), capi.LIB_PATH], capture_output=True, text=True, check=True)
    vals = np.array([float(v) for v in res.stdout.split()])
    tree = mcp.ParseNewick("((a:0.1,b:0.2)e:0.05,(c:0.3,d:0.1)f:0.2)g;")
    ft = mcp.flatten(tree)
    assert list(ft.postorder_num) == [1, 2, 5, 3, 4, 6, 7]
    codes = np.array([[0, 1, 2, 3, 0, 0, 4, 2], [0, 1, 2, 3, 1, 0, 2, 2],
                      [0, 1, 3, 3, 0, 4, 2, 1], [0, 2, 2, 3, 0, 0, 2, 2]], dtype=np.uint8)
    x = oracle.codes_to_dense(codes, ft.leaf_nums, 4, ft.NN)
    U, D, Uinv, mu = mcp.JC(np.full(4, 0.25), [1.0])
    ll, g = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, np.ones(1),
                               np.full(4, 0.25), True, 1)
    assert abs(vals[0] - ll) <= 1e-10 * abs(ll)
    assert np.allclose(vals[1:], g, rtol=1e-8, atol=0)
