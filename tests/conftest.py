import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand with g++."""
    from oracle import hgo
    hgo.build()
    hgo.lib()
    return hgo


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def load_case(name, oracle, golden_dir=None, seed=None):
    """(params, inputs uint64, bounds, segments, num_vars, oracle preprocessing, per-row lookup index)."""
    import numpy as np
    import hyper_greco_b200  # noqa: F401
    from hyper_greco_b200 import params, witness
    P = params.PARAMS[name]
    if seed is None:
        inp = np.load(os.path.join(golden_dir or os.path.join(ROOT, "tests", "golden"), f"lasso_inputs_{name}.npz"))["inputs"]
    else:
        inp = np.array(witness.lasso_inputs(P, witness.synth_witness(P, seed)), dtype=np.uint64)
    bounds = witness.lasso_lookup_bounds(P)
    segs = witness.lasso_lookup_segments(P)
    nv = witness.lasso_num_vars(P)
    opp = oracle.Preprocessing(bounds)
    rows = np.concatenate([np.full(l, opp.lookup_index(b), np.int32) for b, l in segs])
    return P, inp, bounds, segs, nv, opp, rows
