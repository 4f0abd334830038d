"""the incremental solid model, unsIncrTotalLagrangianStress (stressModels/unsIncrTotalLagrangianStress/
unsIncrTotalLagrangianStress.C:946-1187 evolve(), :1212-1233 updateTotalFields(); unsIncrTotalLagrangianStressSolve.C:695-716
residual(); fvPatchFields/tractionDisplacementIncrement/tractionDisplacementIncrementFvPatchVectorField.C:147-307) -- its linear
branch, on the case that ships a complete set-up for it: run/fsiFoam/3dTube/solid (0/DD with tractionDisplacementIncrement /
fixedValue / symmetryPlane patches, unsIncrTotalLagrangianStressCoeffs with nonLinear no; the case selects the total model, a
user switches `stressModel`).  The increment DD goes through the same device kernels as D; the loop differs in residual() and in
its exit test (gpupcg_evolve_controls.incremental)."""
import json
import os

import numpy as np
import pytest

from oracle import evolve as oev
from fluidstructureinteraction_b200 import cases, geometry, meshes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
P_WALL = 1333.2


def _tube():
    m = meshes.PolyMesh.load_npz(os.path.join(GOLD, "tube_solid.polymesh.npz"))
    with open(os.path.join(GOLD, "tube_solid.case.json")) as f:
        S = cases.solid_case_settings(json.load(f))
    g = geometry.compute(m)
    mu, lam = cases.lame(S["E"], S["nu"], S["planeStress"])
    tr, pr, fv = cases.boundary_fields(m, S, pressure={"inner-wall": P_WALL})
    return m, g, S, mu, lam, tr, pr, fv


def _ev(c, S, tol=1e-9):
    return c.evolve(S["nCorrectors"], tol, 0.0, S["tolerance"], S["relTol"], d2dt2="steadyState", ddt="steadyState",
                    deltaT=S["deltaT"], limitCoeff=S["limitCoeff"], relaxD=S["relaxD"], maxIter=S["maxIter"])


def test_oracle_incremental_model_superposes_to_the_total_model():
    """Linear elasticity: three pressure increments through the incremental model must add up to the total model's answer at
    the final pressure (both stopped by the shipped nCorrectors 200 at a residual of ~1e-6, hence the bar) -- a check of the
    traction increment DTraction = t - (n & sigmaf), of DSigmaf and of updateTotalFields() against the model that is pinned."""
    m, g, S, mu, lam, tr, pr, fv = _tube()
    kw = dict(fixedValue=fv, pointPatchTypes=S["pointPatchTypes"])
    tot = oev.SolidCase(m, g, S["patchTypes"], mu, lam, S["rho"], traction=tr, pressure=pr, **kw)
    tot.new_time_step()
    rt = _ev(tot, S)
    inc = oev.IncrSolidCase(m, g, S["patchTypes"], mu, lam, S["rho"], **kw)
    for k in range(3):
        inc.pressure = pr*(k + 1)/3.0
        inc.new_time_step()
        r = _ev(inc, S)
        assert r["nOuterIterations"] > 10
        inc.update_total_fields()
        # one third of the load each: the increments are (nearly) equal
        assert abs(np.abs(inc.inc.D).max()/(np.abs(tot.D).max()/3.0) - 1.0) < 1e-2
    assert rt["res"] < 1e-5 and r["res"] < 1e-5
    assert np.abs(inc.D - tot.D).max() < 1e-3*np.abs(tot.D).max()
    assert np.abs(inc.sigma - tot.sigma).max() < 1e-3*np.abs(tot.sigma).max()
    assert np.abs(inc.pointD - tot.pointD).max() < 1e-3*np.abs(tot.pointD).max()
    # the exit test: res > convergenceTolerance alone (no relConvergenceTolerance) -- a loose tolerance stops the loop early
    inc2 = oev.IncrSolidCase(m, g, S["patchTypes"], mu, lam, S["rho"], pressure=pr, **kw)
    inc2.new_time_step()
    r2 = _ev(inc2, S, tol=1e-3)
    assert r2["nOuterIterations"] < 60 and r2["res"] <= 1e-3 and r2["residualHistory"][-2] > 1e-3


@pytest.mark.gpu
def test_gpu_incremental_model_matches_the_oracle_on_the_shipped_tube():
    from fluidstructureinteraction_b200.solid import GpuIncrSolidCase
    m, g, S, mu, lam, tr, pr, fv = _tube()
    kw = dict(fixedValue=fv, pointPatchTypes=S["pointPatchTypes"])
    ref = oev.IncrSolidCase(m, g, S["patchTypes"], mu, lam, S["rho"], **kw)
    gc = GpuIncrSolidCase(m, g, S["patchTypes"], mu, lam, S["rho"], **kw)
    for k, load in enumerate((1.0/6.0, 0.5, 1.0)):          # unequal increments: the previous DD is not the next answer
        ref.pressure = pr*load; gc.pressure = pr*load
        ref.new_time_step(); gc.new_time_step()
        r, rg = _ev(ref, S, tol=1e-5), _ev(gc, S, tol=1e-5)
        # stopped by `res > convergenceTolerance` alone, before the nCorrectors cap
        assert rg["nOuterIterations"] == r["nOuterIterations"] and 10 < r["nOuterIterations"] < S["nCorrectors"] - 1, (k, r["nOuterIterations"])
        assert np.array_equal(rg["pcgIterations"], np.array(r["pcgIterations"]))
        h, hg = r["residualHistory"], rg["residualHistory"]
        assert np.all(np.abs(h - hg) <= 1e-6*np.abs(h) + 1e-10*np.abs(h).max())
        for name in ("D", "pointD", "gradD", "sigma", "epsilon"):
            a = getattr(ref.inc, name)
            b = gc.inc.get(name).reshape(a.shape)
            assert np.abs(a - b).max() <= 1e-9*np.abs(a).max() + 1e-300, (k, name)
        assert np.abs(ref.DSigmafB - gc.DSigmafB).max() <= 1e-9*np.abs(ref.DSigmafB).max()
        ref.update_total_fields(); gc.update_total_fields()
        for name in ("D", "pointD", "gradDf", "sigma", "sigmafB"):
            a, b = getattr(ref, name), getattr(gc, name)
            assert np.abs(a - b).max() <= 1e-9*np.abs(a).max() + 1e-300, (k, name)
    gc.close()
