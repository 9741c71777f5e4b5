import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, f"{name}.json")) as fh:
        return json.load(fh)


def golden_case(name):
    """(tree, x_dense, codes, leaf_nums, fixture) for a likelihood golden."""
    import mcphylo_jl_b200 as mcp

    fx = load_golden(name)
    tree = mcp.ParseNewick(fx["newick"])
    df = np.array([list(r) for r in fx["rows"]], dtype="<U1")
    x = mcp.datafortree(df, fx["taxa"], tree, fx["symbols"], fx["gap"], fx["missing"])
    codes, leaf_nums = mcp.codesfortree(df, fx["taxa"], tree, fx["symbols"], fx["gap"], fx["missing"])
    return tree, x, codes, leaf_nums, fx


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc

    orc.build()
    return orc
