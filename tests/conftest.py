import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURE_MESHES = ["beam_solid", "tube_solid", "beam_fluid", "tube_fluid"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything native is built in-tree once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def synthetic_matrix(l, u, nCells, seed):
    """Same generator as tests/golden/make_fixtures.py (kept in sync by test_oracle::test_golden_*)."""
    rng = np.random.default_rng(seed)
    coef = rng.uniform(0.5, 2.0, size=l.shape[0])
    upper = -coef
    diag = np.bincount(l, weights=coef, minlength=nCells) + np.bincount(u, weights=coef, minlength=nCells)
    diag = diag + rng.uniform(0.0, 0.05, size=nCells)*(1.0 + diag)
    b = rng.standard_normal(nCells)
    x0 = 0.1*rng.standard_normal(nCells)
    return diag, upper, b, x0


def load_fixture_mesh(name):
    from fluidstructureinteraction_b200 import meshes
    return meshes.PolyMesh.load_npz(os.path.join(GOLDEN, name + ".polymesh.npz"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".golden.npz"), allow_pickle=False)
