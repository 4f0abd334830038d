"""Host logic of the multi-GPU path on CPU: decomposePar-style sub-domains, processor-patch ordering, and a
world_size-2 gloo run in which every rank owns one sub-domain and swaps halos the way the GPU ranks do."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from fluidstructureinteraction_b200 import cases, decompose
from conftest import ROOT


@pytest.mark.parametrize("dims,procs", [((8, 6, 4), (2, 2, 2)), ((9, 5, 4), (2, 1, 1)), ((6, 6, 6), (2, 2, 1)), ((7, 5, 3), (2, 2, 2))])
def test_closed_form_subdomain_equals_decomposed_global_system(dims, procs):
    nx, ny, nz = dims
    px, py, pz = procs
    s = cases.cantilever_system(nx, ny, nz)
    cellProc = decompose.simple_cell_proc(nx, ny, nz, px, py, pz)
    doms = decompose.decompose_ldu(s["l"], s["u"], s["diag"], s["upper"], s["b"], np.zeros(s["nCells"]), cellProc, px*py*pz)
    for r, d in enumerate(doms):
        q = cases.cantilever_subdomain(nx, ny, nz, px, py, pz, r)
        assert np.array_equal(q["cells"], d["cells"])
        assert np.array_equal(q["l"], d["l"]) and np.array_equal(q["u"], d["u"])
        assert np.allclose(q["upper"], d["upper"], rtol=0, atol=0)
        assert np.allclose(q["diag"], d["diag"], rtol=1e-15)
        assert np.array_equal(q["b"], d["b"])
        assert q["patchNbrDomain"] == d["patchNbrDomain"]
        for a, b_ in zip(q["patchFaceCells"], d["patchFaceCells"]):
            assert np.array_equal(a, b_)
        for a, b_ in zip(q["patchBouCoeffs"], d["patchBouCoeffs"]):
            assert np.array_equal(a, b_)


def test_patch_faces_pair_up_across_ranks():
    """Face i of my patch towards rank q and face i of q's patch towards me are the two sides of one mesh face."""
    nx, ny, nz, P = 8, 6, 4, (2, 2, 2)
    subs = [cases.cantilever_subdomain(nx, ny, nz, *P, r) for r in range(8)]
    for r, s in enumerate(subs):
        for p, q in enumerate(s["patchNbrDomain"]):
            t = subs[q]
            back = t["patchNbrDomain"].index(r)
            mine = s["cells"][s["patchFaceCells"][p]]
            theirs = t["cells"][t["patchFaceCells"][back]]
            delta = np.abs(mine - theirs)
            assert np.all((delta == 1) | (delta == nx) | (delta == nx*ny)) and len(set(delta)) == 1


def test_two_rank_gloo_halo_swap_reproduces_global_amul():
    script = os.path.join(ROOT, "tests", "gloo_halo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", script],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "HALO-OK" in r.stdout, r.stdout[-3000:]
