"""Host-side construction of the reference's least-squares vol->point interpolation data
(numerics/leastSquaresVolPointInterpolation/leastSquaresVolPointInterpolation.C): what a foam-extend run reads off its
`leastSquaresVolPointInterpolation volToPoint_` member (pointCells / pointBndFaces stencils, weights() == 1, origins(),
invLsMatrices(), mirror planes) and hands to gpupcg_mesh_set_lsq once per mesh.  Setup code (once per mesh, not on the
timed path), vectorised over points with numpy; the per-outer-iteration interpolation itself is the CUDA kernel
k_lsq_interpolate.  Serial (one sub-domain) case: cells + non-empty/non-wedge boundary faces (+ mirror images).

  makePointFaces               :54-205    boundary faces of a point: labelHashSet::toc() order [EXT HashTable: bucket =
                                          label & 127 ascending, newest insertion first inside a bucket]
  makeWeights                  :856-1135  all 1 (the distance weighting is commented out)
  makeOrigins                  :1138-1423 mean of the stencil points
  makeInvLsMatrices            :1425-1826 inv(M^T M) M^T with M = stencil points - origin, tensor inverse [EXT TensorI.H]
  makeMirrorPlaneTransformation :1829-1940 points of empty patches: (pointNormal, I)
"""
import numpy as np

SMALL = 1.0e-15
VSMALL = 1.0e-300


def _csr_from_pairs(keys, vals, n):
    """rows of `vals` grouped by key (stable in input order): start (n+1,), values."""
    order = np.argsort(keys, kind="stable")
    cnt = np.bincount(keys, minlength=n)
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=start[1:])
    return start, vals[order]


def build(mesh, geom, patchTypes):
    """Returns dict(start (nP+1,) i4, entry (rows,) i4, inv (rows,3) f8, origin (nP,3) f8, mirror (nP,) u1)."""
    nP, nF, nIF, nC = mesh.nPoints, mesh.nFaces, mesh.nInternalFaces, mesh.nCells
    fs = mesh.faceStart.astype(np.int64)
    fpts = mesh.facePoints.astype(np.int64)
    faceOf = np.repeat(np.arange(nF, dtype=np.int64), np.diff(fs))
    own = mesh.owner.astype(np.int64)
    nei = mesh.neighbour.astype(np.int64)
    # ---- mesh.pointCells(): every (cell, point) pair once, cells ascending per point
    cellA = own[faceOf]
    pairs = [cellA*nP + fpts]
    inner = faceOf < nIF
    pairs.append(nei[faceOf[inner]]*nP + fpts[inner])
    key = np.unique(np.concatenate(pairs))                  # sorted by (cell, point)
    pc_cell, pc_point = key // nP, key % nP
    pcStart, pcCells = _csr_from_pairs(pc_point, pc_cell, nP)  # stable: cells ascending inside a point
    # ---- pointBndFaces: non-empty boundary faces of a point in labelHashSet::toc() order
    bfType = np.zeros(nF - nIF, dtype=np.int64)
    emptyPatches = []
    for p in mesh.patches:
        t = patchTypes[p["name"]]
        bfType[p["start"] - nIF:p["start"] - nIF + p["size"]] = 4 if t == "empty" else 0
        if t == "empty":
            emptyPatches.append(p)
    bsel = (faceOf >= nIF)
    bsel[bsel] = bfType[faceOf[bsel] - nIF] != 4
    bf_face, bf_point = faceOf[bsel], fpts[bsel]
    # order inside a point: bucket (f & 127) ascending, then f descending
    o = np.lexsort((-bf_face, bf_face & 127, bf_point))
    bf_face, bf_point = bf_face[o], bf_point[o]
    pbStart = np.zeros(nP + 1, dtype=np.int64)
    np.cumsum(np.bincount(bf_point, minlength=nP), out=pbStart[1:])
    # ---- mirror planes (empty patches): point normal = sum of unit face normals / (mag + VSMALL)
    mirror = np.zeros(nP, dtype=bool)
    mn = np.zeros((nP, 3))
    for p in emptyPatches:
        faces = np.arange(p["start"], p["start"] + p["size"], dtype=np.int64)
        acc = np.zeros((nP, 3))
        touched = np.zeros(nP, dtype=bool)
        unit = geom.Sf[faces]/geom.magSf[faces][:, None]
        # faces ascending, sequential accumulation per point (np.add.at applies repeated indices in order)
        for j in range(int(np.diff(fs)[faces].max()) if faces.size else 0):
            has = (fs[faces + 1] - fs[faces]) > j
            pts = fpts[fs[faces[has]] + j]
            touched[pts] = True
        # accumulate in (face ascending, position in face) order == the order the oracle's loops visit
        rep = np.repeat(np.arange(faces.size), (fs[faces + 1] - fs[faces]))
        pts_all = np.concatenate([fpts[fs[f]:fs[f + 1]] for f in faces]) if faces.size else np.zeros(0, dtype=np.int64)
        for d in range(3):
            np.add.at(acc[:, d], pts_all, unit[rep, d])
        m = np.sqrt(acc[:, 0]*acc[:, 0] + acc[:, 1]*acc[:, 1] + acc[:, 2]*acc[:, 2]) + VSMALL
        mn[touched] = acc[touched]/m[touched][:, None]
        mirror[touched] = True
    mirror &= np.sqrt(mn[:, 0]*mn[:, 0] + mn[:, 1]*mn[:, 1] + mn[:, 2]*mn[:, 2]) > SMALL
    # ---- rows
    nCellRows = np.diff(pcStart)
    nBndRows = np.diff(pbStart)
    nReal = nCellRows + nBndRows
    nRows = np.where(mirror, 2*nReal, nReal)
    start = np.zeros(nP + 1, dtype=np.int64)
    np.cumsum(nRows, out=start[1:])
    rows = int(start[-1])
    entry = np.zeros(rows, dtype=np.int64)
    X = np.zeros((rows, 3))
    # real rows: cells then boundary faces
    pid_c = np.repeat(np.arange(nP), nCellRows)
    pos_c = np.arange(pcCells.shape[0]) - pcStart[pid_c]
    r_c = start[pid_c] + pos_c
    entry[r_c] = pcCells
    X[r_c] = geom.C[pcCells]
    pid_b = np.repeat(np.arange(nP), nBndRows)
    pos_b = np.arange(bf_face.shape[0]) - pbStart[pid_b]
    r_b = start[pid_b] + nCellRows[pid_b] + pos_b
    entry[r_b] = -1 - (bf_face - nIF)
    X[r_b] = geom.Cf[bf_face]
    # mirrored rows: p + transform(I - 2*n*n, x - p)
    rowPoint = np.repeat(np.arange(nP), nRows)
    posInPoint = np.arange(rows) - start[rowPoint]
    isMir = mirror[rowPoint] & (posInPoint >= nReal[rowPoint])
    if isMir.any():
        src = np.flatnonzero(isMir) - nReal[rowPoint[isMir]]
        pp = mesh.points[rowPoint[isMir]]
        nn = mn[rowPoint[isMir]]
        two = 2*nn
        dv = X[src] - pp
        T = np.empty((src.shape[0], 3, 3))
        for a in range(3):
            for b in range(3):
                T[:, a, b] = (1.0 if a == b else 0.0) - two[:, a]*nn[:, b]
        tv = np.empty_like(dv)
        for a in range(3):
            tv[:, a] = T[:, a, 0]*dv[:, 0] + T[:, a, 1]*dv[:, 1] + T[:, a, 2]*dv[:, 2]
        X[isMir] = pp + tv
        entry[isMir] = entry[src]
    # ---- origins, least-squares matrices: sequential sums over the rows of a point, vectorised across points
    mMax = int(nRows.max()) if nP else 0
    o = np.zeros((nP, 3))
    sw = np.zeros(nP)
    for j in range(mMax):
        has = nRows > j
        r = start[:-1][has] + j
        o[has] += 1.0*X[r]
        sw[has] += 1.0
    o /= sw[:, None]
    Mx = (X - o[rowPoint])*1.0
    ls = np.zeros((nP, 3, 3))
    for a in range(3):
        for b in range(3):
            for j in range(mMax):
                has = nRows > j
                r = start[:-1][has] + j
                ls[has, a, b] += Mx[r, a]*Mx[r, b]
    t = ls.reshape(nP, 9)
    det = (t[:, 0]*t[:, 4]*t[:, 8] + t[:, 1]*t[:, 5]*t[:, 6] + t[:, 2]*t[:, 3]*t[:, 7] - t[:, 0]*t[:, 5]*t[:, 7]
           - t[:, 1]*t[:, 3]*t[:, 8] - t[:, 2]*t[:, 4]*t[:, 6])
    il = np.empty((nP, 9))
    il[:, 0] = (t[:, 4]*t[:, 8] - t[:, 7]*t[:, 5])/det; il[:, 1] = (t[:, 2]*t[:, 7] - t[:, 1]*t[:, 8])/det
    il[:, 2] = (t[:, 1]*t[:, 5] - t[:, 2]*t[:, 4])/det; il[:, 3] = (t[:, 6]*t[:, 5] - t[:, 3]*t[:, 8])/det
    il[:, 4] = (t[:, 0]*t[:, 8] - t[:, 2]*t[:, 6])/det; il[:, 5] = (t[:, 3]*t[:, 2] - t[:, 0]*t[:, 5])/det
    il[:, 6] = (t[:, 3]*t[:, 7] - t[:, 4]*t[:, 6])/det; il[:, 7] = (t[:, 1]*t[:, 6] - t[:, 0]*t[:, 7])/det
    il[:, 8] = (t[:, 0]*t[:, 4] - t[:, 3]*t[:, 1])/det
    ilr = il[rowPoint]
    inv = np.empty((rows, 3))
    for a in range(3):
        s = np.zeros(rows)
        for b in range(3):
            s = s + ilr[:, 3*a + b]*Mx[:, b]*1.0
        inv[:, a] = s
    return dict(start=start.astype(np.int32), entry=entry.astype(np.int32), inv=inv, origin=o, mirror=mirror.astype(np.uint8))


def point_constraints(mesh, geom, pointPatchTypes, pointPatchValues=None):
    """The point patch fields of pointD that do something in pf.correctBoundaryConditions()
    (leastSquaresVolPointInterpolate.C:301), as the 0/pointD dictionaries of the reference declare them
    (run/stressFoam/plateHole/plateHole/0/pointD:19-50, run/fsiFoam/beamInCrossFlow/solid/0/pointD:19-37):
    `symmetryPlane` removes the component along the patch's point normal, `fixedValue` sets the value; `calculated` and
    `empty` do nothing.  Returns dict(point (n,), start (n+1,), kind (rows,), vec (rows,3)) for gpupcg_mesh_set_point_constraints:
    constrained point i owns rows start[i]..start[i+1] in patch order (kind 1 symmetryPlane: vec = nHat, 2 fixedValue: vec = value)."""
    fs = mesh.faceStart.astype(np.int64)
    fpts = mesh.facePoints.astype(np.int64)
    nP = mesh.nPoints
    pts_l, kind_l, vec_l = [], [], []
    for p in mesh.patches:
        t = pointPatchTypes.get(p["name"], "calculated")
        if t not in ("symmetryPlane", "fixedValue") or p["size"] == 0:
            continue
        faces = np.arange(p["start"], p["start"] + p["size"], dtype=np.int64)
        allp = np.concatenate([fpts[fs[f]:fs[f + 1]] for f in faces])
        _, first = np.unique(allp, return_index=True)
        mp = allp[np.sort(first)]                                   # meshPoints(): order of first appearance
        if t == "fixedValue":
            v = np.zeros(3) if not pointPatchValues or p["name"] not in pointPatchValues else np.asarray(pointPatchValues[p["name"]], dtype=np.float64)
            vec = np.broadcast_to(v, (mp.shape[0], 3)).copy()
            kind = np.full(mp.shape[0], 2, dtype=np.int32)
        else:
            unit = geom.Sf[faces]/(geom.magSf[faces] + VSMALL)[:, None]
            rep = np.repeat(np.arange(faces.size), fs[faces + 1] - fs[faces])
            acc = np.zeros((nP, 3))
            for d in range(3):
                np.add.at(acc[:, d], allp, unit[rep, d])            # faces ascending: PrimitivePatch::pointFaces order
            a = acc[mp]
            vec = a/(np.sqrt(a[:, 0]*a[:, 0] + a[:, 1]*a[:, 1] + a[:, 2]*a[:, 2]) + VSMALL)[:, None]
            kind = np.full(mp.shape[0], 1, dtype=np.int32)
        pts_l.append(mp); kind_l.append(kind); vec_l.append(vec)
    if not pts_l:
        return dict(point=np.zeros(0, dtype=np.int32), start=np.zeros(1, dtype=np.int32), kind=np.zeros(0, dtype=np.int32),
                    vec=np.zeros((0, 3)))
    pts = np.concatenate(pts_l); kind = np.concatenate(kind_l); vec = np.concatenate(vec_l)
    o = np.argsort(pts, kind="stable")                              # rows of a point stay in patch order
    pts, kind, vec = pts[o], kind[o], vec[o]
    up, cnt = np.unique(pts, return_counts=True)
    start = np.zeros(up.shape[0] + 1, dtype=np.int32)
    np.cumsum(cnt, out=start[1:])
    return dict(point=up.astype(np.int32), start=start, kind=kind.astype(np.int32), vec=np.ascontiguousarray(vec))
