"""the GPU solid model registered in the stress-model run-time selection table next to the stock one
(reference: stressModels/stressModel/newStressModel.C:40-100 selects by `stressModel ...;` of constant/stressProperties;
unsTotalLagrangianStress.C:53-54 registers the stock model).  host/gpuStressModel.{H,C} is the C++ class over the foam-extend
API stand-in; it reads the SHIPPED dictionaries of config 1 and must reproduce the resident Python harness (same C-ABI calls,
hence the same bits) and the oracle's loop."""
import json
import os

import numpy as np
import pytest

from fluidstructureinteraction_b200 import blockmesh, cases, geometry, hostapi, lsq

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

STRESS = """
stressModel %s;
gpuUnsTotalLagrangianStressCoeffs
{
    nCorrectors 1000;               // run/stressFoam/plateHole/plateHole/constant/stressProperties:19-26
    convergenceTolerance 1e-8;
    nonLinear no;
    debug no;
    moveMesh no;
}
"""
RHEOLOGY = """
planeStress     yes;                // run/stressFoam/plateHole/plateHole/constant/rheologyProperties:17-25
rheology
{
    type            linearElastic;
    rho             rho [1 -3 0 0 0 0 0] 7854;
    E               E [1 -1 -2 0 0 0 0] 2e+11;
    nu              nu [0 0 0 0 0 0 0] 0.3;
}
"""
FVSOLUTION = """
solvers { D { solver %s; preconditioner DIC; tolerance 1e-09; relTol 0.01; } }      // system/fvSolution:18-27
relaxationFactors { }
"""
FVSCHEMES = """
d2dt2Schemes { default %s; }
ddtSchemes { default steadyState; }
laplacianSchemes { default none; laplacian(DD,D) Gauss linear skewCorrected 1; }
snGradSchemes { snGrad(D) %s; }
"""


def _plate_hole():
    with open(os.path.join(GOLD, "plateHole.blockMeshDict.json")) as f:
        m = blockmesh.block_mesh(json.load(f))
    g = geometry.compute(m)
    PH = cases.PLATE_HOLE
    return m, g, PH, cases.plate_hole_tractions(m, g, PH["patchTypes"])


def _model(m, g, PH, tr, model="gpuUnsTotalLagrangianStress", solver="gpuPCG", d2dt2="steadyState", sn="skewCorrected 1", **kw):
    return hostapi.StressModel(m, g, PH["patchTypes"], STRESS % model, RHEOLOGY, FVSOLUTION % solver, FVSCHEMES % (d2dt2, sn),
                               lsq=lsq.build(m, g, PH["patchTypes"]), pointConstraints=lsq.point_constraints(m, g, PH["pointPatchTypes"]),
                               solutionD=(1, 1, 0), traction=tr, **kw)


def test_run_time_selection_and_dictionary_checks():
    """No GPU needed: selection by name and the refusal of settings the resident loop does not implement happen before any
    device call; a supported set-up then fails loudly for want of a GPU (no CPU fallback)."""
    import torch
    m, g, PH, tr = _plate_hole()
    with pytest.raises(hostapi.FoamFatalError) as e:
        _model(m, g, PH, tr, model="unsTotalLagrangianStress")
    assert "Unknown stressModel type unsTotalLagrangianStress" in str(e.value) and "gpuUnsTotalLagrangianStress" in str(e.value)
    for kw, msg in ((dict(solver="GAMG"), "PCG/DIC"), (dict(d2dt2="CrankNicolson"), "CrankNicolson"), (dict(sn="corrected"), "skewCorrected")):
        with pytest.raises(hostapi.FoamFatalError) as e:
            _model(m, g, PH, tr, **kw)
        assert msg in str(e.value), (kw, str(e.value))
    if not torch.cuda.is_available():
        with pytest.raises(hostapi.FoamFatalError) as e:
            _model(m, g, PH, tr)
        assert "gpupcg_mesh_create failed" in str(e.value) and "no CPU fallback" in str(e.value)


@pytest.mark.gpu
def test_gpu_stress_model_runs_config_1_like_the_harness_and_the_oracle():
    from oracle import evolve as oev
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m, g, PH, tr = _plate_hole()
    mu, lam = cases.lame(PH["E"], PH["nu"], True)
    kw = dict(traction=tr, validCmpts=(1, 1, 0), pointPatchTypes=PH["pointPatchTypes"])
    ref = oev.SolidCase(m, g, PH["patchTypes"], mu, lam, PH["rho"], **kw)
    r = ref.evolve(PH["nCorrectors"], PH["convergenceTolerance"], 0.0, PH["tolerance"], PH["relTol"])
    gc = GpuSolidCase(m, g, PH["patchTypes"], mu, lam, PH["rho"], **kw)
    rg = gc.evolve(PH["nCorrectors"], PH["convergenceTolerance"], 0.0, PH["tolerance"], PH["relTol"])
    # the C++ model: zero tractions at construction, then what setPlateHoleBC::setBC does through setTraction()
    sm = _model(m, g, PH, np.zeros_like(tr))
    nIF = m.nInternalFaces
    for p in m.patches:
        if PH["patchTypes"][p["name"]] == "tractionDisplacement":
            sm.set_traction(p["name"], tr[p["start"] - nIF:p["start"] - nIF + p["size"]])
    with pytest.raises(hostapi.FoamFatalError):
        sm.set_traction("left", np.zeros((30, 3)))             # not a tractionDisplacement patch (:528-541)
    rs = sm.evolve()
    assert rs["converged"] and rs["nOuterIterations"] == rg["nOuterIterations"] == r["nOuterIterations"]
    assert np.array_equal(rs["pcgIterations"], rg["pcgIterations"]) and np.array_equal(rs["pcgIterations"], np.array(r["pcgIterations"]))
    assert np.array_equal(rs["residualHistory"], rg["residualHistory"])
    for k in ("D", "pointD", "sigma"):
        a = sm.get(k)
        assert np.array_equal(a, gc.get(k).ravel()), k                                  # same calls, same bits
        b = getattr(ref, k).ravel()
        assert np.abs(a - b).max() <= 1e-9*np.abs(b).max(), k
    e = cases.plate_hole_errors(m, g, sm.get("D").reshape(-1, 3), sm.get("pointD").reshape(-1, 3), sm.get("sigma").reshape(-1, 6))
    assert e["DError"] < 0.004*e["DMax"] and e["sigmaXXErr"] < 0.03*e["sigmaXXMax"]
    sm.close(); gc.close()


@pytest.mark.gpu
def test_gpu_stress_model_runs_the_shipped_tube_case_with_its_backward_schemes():
    """run/fsiFoam/3dTube/solid through the C++ model with the SHIPPED dictionaries as text (tests/golden/tube_solid.case.json
    holds them parsed; written back in foam syntax here): d2dt2(D) / ddt(D) backward -- the model keeps D's four old-time levels
    on the device and works out the schemes' start-up rule itself (host/gpuStressModel.C: touchOldTimes / ensureOldTimes).
    Five time steps = the Python harness bit for bit (same C-ABI calls) and the oracle at the north_star bars."""
    from oracle import evolve as oev
    from fluidstructureinteraction_b200 import meshes
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    m = meshes.PolyMesh.load_npz(os.path.join(GOLD, "tube_solid.polymesh.npz"))
    with open(os.path.join(GOLD, "tube_solid.case.json")) as f:
        d = json.load(f)
    g = geometry.compute(m)
    S = cases.solid_case_settings(d)
    assert (S["d2dt2"], S["ddt"]) == ("backward", "backward")
    tr, pr, fv = cases.boundary_fields(m, S, pressure={"inner-wall": 1333.2})
    mu, lam = cases.lame(S["E"], S["nu"], S["planeStress"])
    stress = """stressModel gpuUnsTotalLagrangianStress;
gpuUnsTotalLagrangianStressCoeffs { nCorrectors %d; convergenceTolerance %r; relConvergenceTolerance %r; nonLinear %s; debug no; moveMesh no; }
""" % (S["nCorrectors"], S["convergenceTolerance"], S["relConvergenceTolerance"], "yes" if S["nonLinear"] else "no")
    rheology = """planeStress %s;
rheology { type linearElastic; rho rho [1 -3 0 0 0 0 0] %r; E E [1 -1 -2 0 0 0 0] %r; nu nu [0 0 0 0 0 0 0] %r; }
""" % ("yes" if S["planeStress"] else "no", S["rho"], S["E"], S["nu"])
    fvSolution = "solvers { D { solver gpuPCG; preconditioner DIC; tolerance %r; relTol %r; } }\nrelaxationFactors { %s }\n" % (
        S["tolerance"], S["relTol"], ("D %r;" % S["relaxD"]) if S["relaxD"] is not None else "")
    fvSchemes = """d2dt2Schemes { default none; d2dt2(D) %s; }
ddtSchemes { default none; ddt(D) %s; }
snGradSchemes { default none; snGrad(D) skewCorrected %r; }
""" % (S["d2dt2"], S["ddt"], S["limitCoeff"])
    L = lsq.build(m, g, S["patchTypes"])
    sm = hostapi.StressModel(m, g, S["patchTypes"], stress, rheology, fvSolution, fvSchemes, deltaT=S["deltaT"], pressure=pr,
                             fixedValue=fv, lsq=L, pointConstraints=lsq.point_constraints(m, g, S["pointPatchTypes"]))
    kw = dict(traction=tr, pressure=pr, fixedValue=fv, pointPatchTypes=S["pointPatchTypes"])
    ref = oev.SolidCase(m, g, S["patchTypes"], mu, lam, S["rho"], **kw)
    gc = GpuSolidCase(m, g, S["patchTypes"], mu, lam, S["rho"], lsq=L, **kw)
    ev = lambda c: c.evolve(S["nCorrectors"], S["convergenceTolerance"], S["relConvergenceTolerance"], S["tolerance"], S["relTol"],
                            d2dt2=S["d2dt2"], ddt=S["ddt"], deltaT=S["deltaT"], nonLinear=S["nonLinear"], limitCoeff=S["limitCoeff"],
                            relaxD=S["relaxD"], maxIter=S["maxIter"])
    for step in range(5):
        sm.new_time_step(); gc.new_time_step(); ref.new_time_step()
        rs, rg, r = sm.evolve(), ev(gc), ev(ref)
        assert rs["nOuterIterations"] == rg["nOuterIterations"] == r["nOuterIterations"], step
        assert np.array_equal(rs["pcgIterations"], rg["pcgIterations"]) and np.array_equal(rs["pcgIterations"], np.array(r["pcgIterations"]))
        for k in ("D", "pointD", "sigma", "U"):
            a = sm.get(k)
            assert np.array_equal(a, gc.get(k).ravel()), (step, k)
            b = getattr(ref, k).ravel()
            assert np.abs(a - b).max() <= 1e-9*np.abs(b).max(), (step, k)
    sm.close(); gc.close()
