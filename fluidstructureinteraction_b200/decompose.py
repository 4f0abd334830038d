"""decomposePar-style domain decomposition at the LDU level (one sub-domain per GPU).

Mirrors what the reference's forked decomposePar produces (src/utilities/decomposePar/decomposeMesh.C:67-230,
domainDecomposition.C:293-340): cells of a processor in ascending global id, its internal faces in ascending
global face id (so every sub-domain keeps upper-triangular order), and the cut faces grouped into one
processor patch per neighbouring processor, neighbours ascending, faces in global face order -- the same
order on both sides, which is what the halo swap relies on.

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
        for q in np.unique(nbr):
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
