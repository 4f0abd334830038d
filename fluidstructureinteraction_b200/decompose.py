"""decomposePar-style domain decomposition (one sub-domain per GPU): LDU level and whole polyMesh.

Mirrors what the reference's forked decomposePar produces (src/utilities/decomposePar/decomposeMesh.C:44-843,
domainDecomposition.C:95-620): cells of a processor in ascending global id, its internal faces in ascending
global face id (so every sub-domain keeps upper-triangular order), then the faces of the original patches, then
the cut faces grouped into one processor patch per neighbouring processor -- neighbours in the order in which the
ascending face loop first meets them (decomposeMesh.C:121-230), faces in global face order: the same order on
both sides, which is what the halo swap relies on.  On the box decompositions of the benchmark that order is the
ascending rank order (-z, -y, -x, +x, +y, +z).

The interface coefficient of a cut face is what fvMatrix hands to lduMatrix::Amul for a processor patch:
Ax[faceCells[i]] -= bouCoeffs[i]*xNbr[i], i.e. bouCoeffs = -upper[f] of the undecomposed matrix.
"""
import numpy as np


def simple_cell_proc(nx, ny, nz, px, py, pz):
    """foam-extend `simple` decomposition of the structured box into px x py x pz blocks (ranks x-fastest)."""
    c = np.arange(nx*ny*nz, dtype=np.int64)
    i, j, k = c % nx, (c // nx) % ny, c // (nx*ny)
    bi = np.minimum(i*px // nx, px - 1)
    bj = np.minimum(j*py // ny, py - 1)
    bk = np.minimum(k*pz // nz, pz - 1)
    return (bi + px*(bj + py*bk)).astype(np.int32)


def proc_grid(nRanks):
    """2x1x1, 2x2x1, 2x2x2 ...."""
    grid = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    if nRanks not in grid:
        raise ValueError("supported rank counts: 1, 2, 4, 8")
    return grid[nRanks]


def decompose_addressing(l, u, cellProc, nProcs):
    """Returns per-processor dict(cells, faces, l, u, patchNbr, patchFaces(global ids), patchFaceCells(local),
    patchNbrPatch) -- addressing only."""
    l = np.asarray(l, dtype=np.int64)
    u = np.asarray(u, dtype=np.int64)
    cellProc = np.asarray(cellProc, dtype=np.int64)
    n = cellProc.shape[0]
    local = np.empty(n, dtype=np.int64)
    doms = []
    for p in range(nProcs):
        cells = np.flatnonzero(cellProc == p)
        local[cells] = np.arange(cells.shape[0])
        doms.append(dict(cells=cells))
    pl, pu = cellProc[l], cellProc[u]
    cut = np.flatnonzero(pl != pu)
    for p in range(nProcs):
        d = doms[p]
        f = np.flatnonzero((pl == p) & (pu == p))
        d["faces"] = f
        d["l"] = local[l[f]].astype(np.int32)
        d["u"] = local[u[f]].astype(np.int32)
        # cut faces seen from p, with the neighbour processor and the local cell
        mine_l = cut[pl[cut] == p]
        mine_u = cut[pu[cut] == p]
        fcut = np.concatenate([mine_l, mine_u])
        nbr = np.concatenate([pu[mine_l], pl[mine_u]])
        cellLoc = np.concatenate([local[l[mine_l]], local[u[mine_u]]])
        order = np.lexsort((fcut, nbr))
        fcut, nbr, cellLoc = fcut[order], nbr[order], cellLoc[order]
        d["patchNbr"] = []
        d["patchFaces"] = []
        d["patchFaceCells"] = []
        # neighbours in order of first encounter by the ascending face loop = ascending smallest cut-face id
        uq = np.unique(nbr)
        firstFace = np.array([fcut[nbr == q].min() for q in uq]) if uq.size else np.zeros(0)
        for q in uq[np.argsort(firstFace, kind="stable")]:
            sel = nbr == q
            d["patchNbr"].append(int(q))
            d["patchFaces"].append(fcut[sel])
            d["patchFaceCells"].append(cellLoc[sel].astype(np.int32))
    for p in range(nProcs):
        d = doms[p]
        d["patchNbrPatch"] = [doms[q]["patchNbr"].index(p) for q in d["patchNbr"]]
    return doms


def decompose_ldu(l, u, diag, upper, b, x, cellProc, nProcs):
    """Sub-domain systems in the layout oracle.pcg_solve_multi and GpuPCG(rank) consume."""
    doms = decompose_addressing(l, u, cellProc, nProcs)
    out = []
    for d in doms:
        cells = d["cells"]
        out.append(dict(
            cells=cells, l=d["l"], u=d["u"],
            diag=np.ascontiguousarray(diag[cells]), upper=np.ascontiguousarray(upper[d["faces"]]),
            b=np.ascontiguousarray(b[cells]), x=np.ascontiguousarray(x[cells], dtype=np.float64),
            patchFaceCells=d["patchFaceCells"],
            patchBouCoeffs=[np.ascontiguousarray(-upper[f]) for f in d["patchFaces"]],
            patchNbrDomain=d["patchNbr"], patchNbrPatch=d["patchNbrPatch"]))
    return out


def reconstruct(doms, nCells, key="x"):
    """reconstructPar for one cell field."""
    out = np.empty(nCells, dtype=np.float64)
    for d in doms:
        out[d["cells"]] = d[key]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# [EXT] simpleGeomDecomp (decompositionMethods/simpleGeomDecomp.C): `method simple; simpleCoeffs { n (nx ny nz); delta d; }`
# ---------------------------------------------------------------------------------------------------------------------
def _assign_groups(n, nGroups):
    """simpleGeomDecomp::assignToProcessorGroup: the first (n mod nGroups) groups take one element more."""
    jump = n // nGroups
    first = n - jump*nGroups
    return np.concatenate([np.repeat(np.arange(first), jump + 1), np.repeat(np.arange(first, nGroups), jump)]).astype(np.int64)


def simple_decomp(cellCentres, n, delta=0.001):
    """cellToProc of `method simple`: the points are rotated by a small angle (rotDelta, so that no two share a coordinate),
    sorted along each direction and cut into n[d] groups of (almost) equal size; proc = gx + n.x*(gy + n.y*gz)."""
    C = np.asarray(cellCentres, dtype=np.float64)
    d, d2 = float(delta), 1.0 - 0.5*float(delta)*float(delta)
    R = np.array([[d2, -d, d], [d, d2, -d], [-d, d, d2]])
    P = C@R.T
    out = np.zeros(C.shape[0], dtype=np.int64)
    mult = 1
    for ax in range(3):
        order = np.argsort(P[:, ax], kind="stable")
        g = np.empty(C.shape[0], dtype=np.int64)
        g[order] = _assign_groups(C.shape[0], int(n[ax]))
        out += mult*g
        mult *= int(n[ax])
    return out.astype(np.int32)


# ---------------------------------------------------------------------------------------------------------------------
# the processor meshes decomposePar writes to processorN/constant/polyMesh, with the *ProcAddressing lists
# ---------------------------------------------------------------------------------------------------------------------
def decompose_mesh(mesh, cellProc, nProcs, filterEmptyPatches=False):
    """decomposeMesh.C:44-843 + domainDecomposition.C:95-620 (no cyclics, no global face zones).  Returns per processor
    dict(mesh (PolyMesh; processor patches carry myProcNo / neighbProcNo), cellProcAddressing, faceProcAddressing (global face
    + 1, negative: turned), pointProcAddressing, boundaryProcAddressing (original patch index, -1 processor patch))."""
    from .meshes import PolyMesh
    cellProc = np.asarray(cellProc, dtype=np.int64)
    own = mesh.owner.astype(np.int64)
    nei = mesh.neighbour.astype(np.int64)
    nIF = mesh.nInternalFaces
    fs = mesh.faceStart.astype(np.int64)
    fpts = mesh.facePoints.astype(np.int64)
    doms = decompose_addressing(own[:nIF], nei, cellProc, nProcs)
    out = []
    for p in range(nProcs):
        d = doms[p]
        cells = d["cells"]
        cellLookup = np.full(mesh.nCells, -1, dtype=np.int64)
        cellLookup[cells] = np.arange(cells.shape[0])
        faceAddr = [d["faces"] + 1]
        patches, bAddr = [], []
        start = d["faces"].shape[0]
        for pi, pt in enumerate(mesh.patches):
            f = np.arange(pt["start"], pt["start"] + pt["size"], dtype=np.int64)
            f = f[cellProc[own[f]] == p]
            if filterEmptyPatches and f.size == 0:
                continue
            patches.append(dict(name=pt["name"], type=pt["type"], start=start, size=int(f.size)))
            bAddr.append(pi)
            faceAddr.append(f + 1)
            start += int(f.size)
        for q, f in zip(d["patchNbr"], d["patchFaces"]):
            turned = cellProc[own[f]] != p
            faceAddr.append(np.where(turned, -(f + 1), f + 1))
            patches.append(dict(name="procBoundary%dto%d" % (p, q), type="processor", start=start, size=int(f.size),
                                myProcNo=p, neighbProcNo=int(q)))
            bAddr.append(-1)
            start += int(f.size)
        faceAddr = np.concatenate(faceAddr)
        g = np.abs(faceAddr) - 1
        used = np.zeros(mesh.nPoints, dtype=bool)
        sizes = fs[g + 1] - fs[g]
        flat = np.concatenate([fpts[fs[f]:fs[f + 1]] for f in g]) if g.size else np.zeros(0, dtype=np.int64)
        used[flat] = True
        pointAddr = np.flatnonzero(used)
        pointLookup = np.full(mesh.nPoints, -1, dtype=np.int64)
        pointLookup[pointAddr] = np.arange(pointAddr.shape[0])
        newStart = np.zeros(g.shape[0] + 1, dtype=np.int64)
        np.cumsum(sizes, out=newStart[1:])
        newPts = np.empty(int(newStart[-1]), dtype=np.int64)
        for i, (f, a) in enumerate(zip(g, faceAddr)):
            loop = fpts[fs[f]:fs[f + 1]]
            if a < 0:
                loop = np.concatenate([loop[:1], loop[:0:-1]])         # face::reverseFace(): first point stays
            newPts[newStart[i]:newStart[i + 1]] = pointLookup[loop]
        procOwner = np.where(faceAddr > 0, cellLookup[own[g]], cellLookup[nei[np.minimum(g, nIF - 1)]] if nIF else -1)
        nInt = d["faces"].shape[0]
        procNei = cellLookup[nei[g[:nInt]]]
        m = PolyMesh(mesh.points[pointAddr], newStart, newPts, procOwner, procNei, patches, cells.shape[0])
        out.append(dict(mesh=m, cellProcAddressing=cells.astype(np.int32), faceProcAddressing=faceAddr.astype(np.int32),
                        pointProcAddressing=pointAddr.astype(np.int32), boundaryProcAddressing=np.array(bAddr, dtype=np.int32)))
    return out


def write_case(caseDir, mesh, cellProc, nProcs, filterEmptyPatches=False):
    """processor0 ... processorN-1 / constant/polyMesh/{points, faces, owner, neighbour, boundary, cellProcAddressing,
    faceProcAddressing, pointProcAddressing, boundaryProcAddressing} as decomposePar writes them (domainDecomposition.C:340-620)."""
    import os
    from . import meshes
    doms = decompose_mesh(mesh, cellProc, nProcs, filterEmptyPatches)
    for p, d in enumerate(doms):
        pm = os.path.join(caseDir, "processor%d" % p, "constant", "polyMesh")
        meshes.write_polymesh(pm, d["mesh"])
        for k in ("cellProcAddressing", "faceProcAddressing", "pointProcAddressing", "boundaryProcAddressing"):
            meshes.write_label_list(os.path.join(pm, k), d[k], k)
    return doms


def read_case(caseDir, nProcs):
    import os
    from . import meshes
    out = []
    for p in range(nProcs):
        pm = os.path.join(caseDir, "processor%d" % p, "constant", "polyMesh")
        d = dict(mesh=meshes.read_polymesh(pm))
        for k in ("cellProcAddressing", "faceProcAddressing", "pointProcAddressing", "boundaryProcAddressing"):
            d[k] = meshes.read_label_list(os.path.join(pm, k))
        out.append(d)
    return out
