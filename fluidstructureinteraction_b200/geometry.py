"""fvMesh geometry as foam-extend-3.1 computes it (INPUT generator for the assembly kernels and their oracle).

In a foam-extend run these fields come from the mesh object (mesh.Sf(), mesh.magSf(), mesh.Cf(), mesh.C(), mesh.V(),
mesh.weights(), mesh.deltaCoeffs()) and are uploaded once per mesh; here they are restated from the published
algorithms (SURVEY.md Appendix A.8 -- [EXT] primitiveMeshFaceCentresAndAreas.C, primitiveMeshCellCentresAndVols.C,
surfaceInterpolation::makeWeights / makeDeltaCoeffs).  Both the oracle and the GPU kernels take them as given, so a
detail that differed from the real 3.1 source would move config-1 numbers but not oracle <-> GPU parity.
"""
import numpy as np

VSMALL = 1.0e-300


class FvGeometry(object):
    """Sf (nF,3), magSf (nF,), Cf (nF,3), C (nC,3), V (nC,), weights (nIF,), deltaCoeffs (nF,) incl. boundary faces."""
    pass


def _face_geometry(points, faceStart, facePoints):
    nF = faceStart.shape[0] - 1
    sizes = np.diff(faceStart)
    Cf = np.zeros((nF, 3))
    Sf = np.zeros((nF, 3))
    for n in np.unique(sizes):
        sel = np.flatnonzero(sizes == n)
        idx = faceStart[sel][:, None] + np.arange(n)[None, :]
        P = points[facePoints[idx]]                                     # (m, n, 3)
        if n == 3:
            Cf[sel] = (1.0/3.0)*(P[:, 0] + P[:, 1] + P[:, 2])
            Sf[sel] = 0.5*np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
        else:
            fCentre = P[:, 0].copy()
            for k in range(1, n):
                fCentre += P[:, k]
            fCentre /= n
            sumN = np.zeros((sel.shape[0], 3))
            sumA = np.zeros(sel.shape[0])
            sumAc = np.zeros((sel.shape[0], 3))
            for k in range(n):
                nxt = P[:, (k + 1) % n]
                c = P[:, k] + nxt + fCentre
                nvec = np.cross(nxt - P[:, k], fCentre - P[:, k])
                a = np.sqrt(np.sum(nvec*nvec, axis=1))
                sumN += nvec
                sumA += a
                sumAc += a[:, None]*c
            Cf[sel] = (1.0/3.0)*sumAc/(sumA + VSMALL)[:, None]
            Sf[sel] = 0.5*sumN
    return Cf, Sf


def compute(mesh):
    g = FvGeometry()
    own, nei = mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64)
    nC, nF, nIF = mesh.nCells, mesh.nFaces, mesh.nInternalFaces
    Cf, Sf = _face_geometry(mesh.points, mesh.faceStart.astype(np.int64), mesh.facePoints.astype(np.int64))
    # cell centres and volumes by pyramid decomposition about the face-centre average
    cEst = np.zeros((nC, 3))
    nCellFaces = np.zeros(nC)
    np.add.at(cEst, own, Cf)
    np.add.at(nCellFaces, own, 1.0)
    np.add.at(cEst, nei, Cf[:nIF])
    np.add.at(nCellFaces, nei, 1.0)
    cEst /= nCellFaces[:, None]
    ctr = np.zeros((nC, 3))
    vol = np.zeros(nC)
    p3 = np.maximum(np.sum(Sf*(Cf - cEst[own]), axis=1), VSMALL)
    pc = 0.75*Cf + 0.25*cEst[own]
    np.add.at(ctr, own, p3[:, None]*pc)
    np.add.at(vol, own, p3)
    p3n = np.maximum(np.sum(Sf[:nIF]*(cEst[nei] - Cf[:nIF]), axis=1), VSMALL)
    pcn = 0.75*Cf[:nIF] + 0.25*cEst[nei]
    np.add.at(ctr, nei, p3n[:, None]*pcn)
    np.add.at(vol, nei, p3n)
    g.C = ctr/vol[:, None]
    g.V = vol*(1.0/3.0)
    g.Sf, g.Cf = Sf, Cf
    g.magSf = np.sqrt(np.sum(Sf*Sf, axis=1))
    # linear interpolation weights and delta coefficients
    SfdOwn = np.abs(np.sum(Sf[:nIF]*(Cf[:nIF] - g.C[own[:nIF]]), axis=1))
    SfdNei = np.abs(np.sum(Sf[:nIF]*(g.C[nei] - Cf[:nIF]), axis=1))
    g.weights = SfdNei/(SfdOwn + SfdNei)
    delta = g.C[nei] - g.C[own[:nIF]]
    magDelta = np.sqrt(np.sum(delta*delta, axis=1))
    nHat = Sf[:nIF]/g.magSf[:nIF, None]
    dc = np.empty(nF)
    dc[:nIF] = 1.0/np.maximum(np.sum(nHat*delta, axis=1), 0.05*magDelta)
    db = Cf[nIF:] - g.C[own[nIF:]]
    dc[nIF:] = 1.0/np.sqrt(np.sum(db*db, axis=1))
    g.deltaCoeffs = dc
    return g
