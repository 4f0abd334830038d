"""Mesh inputs for the D-equation path: synthetic cantilever generators and a polyMesh reader.

A mesh here is what foam-extend's polyMesh stores (constant/polyMesh/{points,faces,owner,neighbour,
boundary}, e.g. /root/reference/run/fsiFoam/beamInCrossFlow/solid/constant/polyMesh/): points,
faces as point loops whose normal points from owner to neighbour, owner/neighbour labels with the
internal faces first and in upper-triangular order, then the boundary patches as contiguous face
ranges.  These are INPUT generators (test/bench data), not part of the timed path.
"""
import os
import re

import numpy as np


class PolyMesh(object):
    """points (P,3) f8; faces: CSR (faceStart (nF+1,), facePoints) i4; owner (nF,), neighbour (nIF,) i4;
    patches: list of dict(name, type, start, size)."""

    def __init__(self, points, faceStart, facePoints, owner, neighbour, patches, nCells=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.faceStart = np.ascontiguousarray(faceStart, dtype=np.int32)
        self.facePoints = np.ascontiguousarray(facePoints, dtype=np.int32)
        self.owner = np.ascontiguousarray(owner, dtype=np.int32)
        self.neighbour = np.ascontiguousarray(neighbour, dtype=np.int32)
        self.patches = patches
        self.nCells = int(nCells if nCells is not None else (self.owner.max() + 1 if self.owner.size else 0))

    @property
    def nFaces(self):
        return int(self.owner.shape[0])

    @property
    def nInternalFaces(self):
        return int(self.neighbour.shape[0])

    @property
    def nPoints(self):
        return int(self.points.shape[0])

    def lower_upper(self):
        """lduAddressing of the fvMesh: lowerAddr = owner[:nIF], upperAddr = neighbour."""
        return self.owner[:self.nInternalFaces].copy(), self.neighbour.copy()

    def patch(self, name):
        for p in self.patches:
            if p["name"] == name:
                return p
        raise KeyError(name)

    def patch_face_cells(self, p):
        if isinstance(p, str):
            p = self.patch(p)
        return self.owner[p["start"]:p["start"] + p["size"]]

    def repatch(self, labels, names):
        """[EXT autoPatch / createPatch-like] a mesh with the boundary faces regrouped: labels[bf] in 0..len(names)-1 per boundary
        face; the faces of a new patch keep their relative order.  Returns (mesh, old boundary-face index of every new one)."""
        nIF, nF = self.nInternalFaces, self.nFaces
        labels = np.asarray(labels, dtype=np.int64)
        order = np.argsort(labels, kind="stable")
        idx = np.concatenate([np.arange(nIF), nIF + order])
        cnt = np.diff(self.faceStart)[idx]
        fs = np.concatenate([[0], np.cumsum(cnt)])
        fp = np.concatenate([self.facePoints[self.faceStart[f]:self.faceStart[f + 1]] for f in idx]) if nF else self.facePoints
        sizes = np.bincount(labels, minlength=len(names))
        patches, start = [], nIF
        for n, sz in zip(names, sizes):
            patches.append(dict(name=n, type="patch", start=int(start), size=int(sz)))
            start += int(sz)
        return PolyMesh(self.points, fs, fp, self.owner[idx], self.neighbour, patches, self.nCells), order

    def is_upper_triangular(self):
        l, u = self.lower_upper()
        if l.size == 0:
            return True
        key = l.astype(np.int64)*(self.nCells + 1) + u
        return bool(np.all(l < u) and np.all(np.diff(key) > 0))

    def save_npz(self, path):
        np.savez_compressed(
            path, points=self.points, faceStart=self.faceStart, facePoints=self.facePoints,
            owner=self.owner, neighbour=self.neighbour, nCells=np.int64(self.nCells),
            patchNames=np.array([p["name"] for p in self.patches]),
            patchTypes=np.array([p["type"] for p in self.patches]),
            patchStart=np.array([p["start"] for p in self.patches], dtype=np.int64),
            patchSize=np.array([p["size"] for p in self.patches], dtype=np.int64))

    @staticmethod
    def load_npz(path):
        z = np.load(path, allow_pickle=False)
        patches = [dict(name=str(n), type=str(t), start=int(s), size=int(k))
                   for n, t, s, k in zip(z["patchNames"], z["patchTypes"], z["patchStart"], z["patchSize"])]
        return PolyMesh(z["points"], z["faceStart"], z["facePoints"], z["owner"], z["neighbour"], patches,
                        int(z["nCells"]))


# ------------------------------------------------------------------------------------------------
# structured hex box ("synthetic 3-D hex cantilever")
# ------------------------------------------------------------------------------------------------
def hex_ldu(nx, ny, nz):
    """lowerAddr/upperAddr of an nx x ny x nz box, cell id = i + nx*(j + ny*k), faces ordered by owner then
    neighbour (per owner: +x, +y, +z).  O(N) and memory-lean: usable for the 64 M-cell case."""
    n = nx*ny*nz
    c = np.arange(n, dtype=np.int64)
    i = c % nx
    j = (c // nx) % ny
    k = c // (nx*ny)
    hx = (i < nx - 1)
    hy = (j < ny - 1)
    hz = (k < nz - 1)
    del i, j, k
    cnt = hx.astype(np.int64) + hy + hz
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=start[1:])
    nF = int(start[-1])
    l = np.empty(nF, dtype=np.int32)
    u = np.empty(nF, dtype=np.int32)
    s = start[:-1]
    px = s[hx]
    l[px] = c[hx]
    u[px] = c[hx] + 1
    py = (s + hx)[hy]
    l[py] = c[hy]
    u[py] = c[hy] + nx
    pz = (s + hx + hy)[hz]
    l[pz] = c[hz]
    u[pz] = c[hz] + nx*ny
    return l, u


def hex_box(nx, ny, nz, L=2.0, W=0.5, H=1.0, split_free=False):
    """Full polyMesh of the cantilever box.  Patches: fixed (x=0), loaded (x=L), free (the four sides); with split_free
    the four sides become `sides` (y = 0, y = W) and `frontAndBack` (z = 0, z = H: the empty patch of a 2-D case, nz = 1)."""
    npx, npy = nx + 1, ny + 1
    xs = np.linspace(0.0, L, nx + 1)
    ys = np.linspace(0.0, W, ny + 1)
    zs = np.linspace(0.0, H, nz + 1)
    K, J, I = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    points = np.stack([xs[I.ravel()], ys[J.ravel()], zs[K.ravel()]], axis=1)

    def P(i, j, k):
        return (i + npx*(j + npy*k)).astype(np.int32)

    n = nx*ny*nz
    c = np.arange(n, dtype=np.int64)
    ci = c % nx
    cj = (c // nx) % ny
    ck = c // (nx*ny)
    l, u = hex_ldu(nx, ny, nz)
    nIF = l.shape[0]
    fp = np.empty((nIF, 4), dtype=np.int32)
    # which direction is each internal face?  u - l is 1, nx or nx*ny
    d = u.astype(np.int64) - l
    i, j, k = ci[l], cj[l], ck[l]
    mx = d == 1
    # when nx == 1 there are no x faces; when nx*ny == nx (ny == 1) y/z deltas coincide: resolve by position
    my = (d == nx) & ~mx & (ny > 1)
    mz = ~mx & ~my
    a, b, e = i[mx], j[mx], k[mx]
    fp[mx] = np.stack([P(a + 1, b, e), P(a + 1, b + 1, e), P(a + 1, b + 1, e + 1), P(a + 1, b, e + 1)], 1)
    a, b, e = i[my], j[my], k[my]
    fp[my] = np.stack([P(a, b + 1, e), P(a, b + 1, e + 1), P(a + 1, b + 1, e + 1), P(a + 1, b + 1, e)], 1)
    a, b, e = i[mz], j[mz], k[mz]
    fp[mz] = np.stack([P(a, b, e + 1), P(a + 1, b, e + 1), P(a + 1, b + 1, e + 1), P(a, b + 1, e + 1)], 1)

    owners = [l]
    faces = [fp]
    patches = []
    start = nIF

    def add_patch(name, ptype, cells, quad):
        nonlocal start
        patches.append(dict(name=name, type=ptype, start=start, size=int(cells.shape[0])))
        owners.append(cells.astype(np.int32))
        faces.append(quad)
        start += int(cells.shape[0])

    z0 = np.zeros
    # fixed: x = 0, outward normal -x
    cells = c[ci == 0]; j, k = cj[cells], ck[cells]; o = z0(cells.shape[0], dtype=np.int64)
    add_patch("fixed", "patch", cells, np.stack([P(o, j, k), P(o, j, k + 1), P(o, j + 1, k + 1), P(o, j + 1, k)], 1))
    # loaded: x = L, +x
    cells = c[ci == nx - 1]; j, k = cj[cells], ck[cells]; o = z0(cells.shape[0], dtype=np.int64) + nx
    add_patch("loaded", "patch", cells, np.stack([P(o, j, k), P(o, j + 1, k), P(o, j + 1, k + 1), P(o, j, k + 1)], 1))
    # free: y = 0, y = W, z = 0, z = H
    fc, fq = [], []
    cells = c[cj == 0]; i, k = ci[cells], ck[cells]; o = z0(cells.shape[0], dtype=np.int64)
    fc.append(cells); fq.append(np.stack([P(i, o, k), P(i + 1, o, k), P(i + 1, o, k + 1), P(i, o, k + 1)], 1))
    cells = c[cj == ny - 1]; i, k = ci[cells], ck[cells]; o = z0(cells.shape[0], dtype=np.int64) + ny
    fc.append(cells); fq.append(np.stack([P(i, o, k), P(i, o, k + 1), P(i + 1, o, k + 1), P(i + 1, o, k)], 1))
    cells = c[ck == 0]; i, j = ci[cells], cj[cells]; o = z0(cells.shape[0], dtype=np.int64)
    fc.append(cells); fq.append(np.stack([P(i, j, o), P(i, j + 1, o), P(i + 1, j + 1, o), P(i + 1, j, o)], 1))
    cells = c[ck == nz - 1]; i, j = ci[cells], cj[cells]; o = z0(cells.shape[0], dtype=np.int64) + nz
    fc.append(cells); fq.append(np.stack([P(i, j, o), P(i + 1, j, o), P(i + 1, j + 1, o), P(i, j + 1, o)], 1))
    if split_free:
        add_patch("sides", "patch", np.concatenate(fc[:2]), np.concatenate(fq[:2]))
        add_patch("frontAndBack", "empty", np.concatenate(fc[2:]), np.concatenate(fq[2:]))
    else:
        add_patch("free", "patch", np.concatenate(fc), np.concatenate(fq))

    owner = np.concatenate(owners)
    quads = np.concatenate(faces)
    faceStart = np.arange(0, 4*quads.shape[0] + 1, 4, dtype=np.int32)
    return PolyMesh(points, faceStart, quads.ravel(), owner, u, patches, n)


# ------------------------------------------------------------------------------------------------
# generic builder: cells given as vertex lists of one shape (tets) -> polyMesh in OpenFOAM order
# ------------------------------------------------------------------------------------------------
_TET_FACES = np.array([[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]])   # outward for a positively oriented tet


def polymesh_from_tets(points, tets, classify):
    """tets (nC,4) vertex ids, positively oriented ((b-a)x(c-a)).(d-a) > 0.  classify(face centres (n,3)) ->
    integer patch id per boundary face; returns PolyMesh with patches named by `classify.names`."""
    points = np.asarray(points, dtype=np.float64)
    tets = np.asarray(tets, dtype=np.int64)
    nC = tets.shape[0]
    fv = tets[:, _TET_FACES].reshape(-1, 3)                 # (4nC,3) outward-oriented from its cell
    cell = np.repeat(np.arange(nC, dtype=np.int64), 4)
    key = np.sort(fv, axis=1)
    order = np.lexsort((cell, key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    first = np.ones(order.shape[0], dtype=bool)
    first[1:] = ~same
    paired = np.zeros(order.shape[0], dtype=bool)
    paired[:-1] = same
    # internal faces: entry `first & paired` is the lower cell (lexsort keeps cell ascending inside a key)
    iA = order[first & paired]
    iB = order[np.flatnonzero(first & paired) + 1]
    own, nei = cell[iA], cell[iB]
    fpts = fv[iA]                                           # outward from the owner = towards the neighbour
    srt = np.lexsort((nei, own))
    own, nei, fpts = own[srt], nei[srt], fpts[srt]
    # boundary faces
    iBnd = order[first & ~paired]
    iBnd.sort()
    bpts = fv[iBnd]
    bown = cell[iBnd]
    ctr = points[bpts].mean(axis=1)
    pid = classify(ctr)
    names = classify.names
    owners = [own]
    faces = [fpts]
    patches = []
    start = own.shape[0]
    for p, name in enumerate(names):
        sel = np.flatnonzero(pid == p)
        sel = sel[np.argsort(bown[sel], kind="stable")]
        patches.append(dict(name=name, type="patch", start=start, size=int(sel.shape[0])))
        owners.append(bown[sel])
        faces.append(bpts[sel])
        start += int(sel.shape[0])
    owner = np.concatenate(owners)
    tri = np.concatenate(faces)
    faceStart = np.arange(0, 3*tri.shape[0] + 1, 3, dtype=np.int32)
    return PolyMesh(points, faceStart, tri.ravel(), owner, nei, patches, nC)


def tet_box(nx, ny, nz, L=2.0, W=0.5, H=1.0, jitter=0.0, seed=20240601):
    """Each hex of the box split into 6 tets around the (0,0,0)-(1,1,1) diagonal (Kuhn/Freudenthal, faces
    conform), cells numbered hex-major.  jitter: interior vertices moved by U(-j*h, j*h) (skewness)."""
    npx, npy = nx + 1, ny + 1
    xs = np.linspace(0.0, L, nx + 1)
    ys = np.linspace(0.0, W, ny + 1)
    zs = np.linspace(0.0, H, nz + 1)
    K, J, I = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    points = np.stack([xs[I], ys[J], zs[K]], axis=1)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        h = np.array([L/nx, W/ny, H/nz])
        interior = (I > 0) & (I < nx) & (J > 0) & (J < ny) & (K > 0) & (K < nz)
        points[interior] += rng.uniform(-jitter, jitter, size=(int(interior.sum()), 3))*h
    c = np.arange(nx*ny*nz, dtype=np.int64)
    i, j, k = c % nx, (c // nx) % ny, c // (nx*ny)

    def P(di, dj, dk):
        return (i + di) + npx*((j + dj) + npy*(k + dk))

    v = {(a, b, e): P(a, b, e) for a in (0, 1) for b in (0, 1) for e in (0, 1)}
    # the 6 monotone lattice paths from (0,0,0) to (1,1,1); orientation fixed so the volume is positive
    paths = [((1, 0, 0), (1, 1, 0)), ((1, 0, 0), (1, 0, 1)), ((0, 1, 0), (1, 1, 0)),
             ((0, 1, 0), (0, 1, 1)), ((0, 0, 1), (1, 0, 1)), ((0, 0, 1), (0, 1, 1))]
    tets = np.empty((c.shape[0], 6, 4), dtype=np.int64)
    for t, (p1, p2) in enumerate(paths):
        a, b, e, d = v[(0, 0, 0)], v[p1], v[p2], v[(1, 1, 1)]
        e1 = np.array(p1, dtype=float)
        e2 = np.array(p2, dtype=float)
        vol = np.dot(np.cross(e1, e2), np.ones(3))
        if vol > 0:
            tets[:, t] = np.stack([a, b, e, d], 1)
        else:
            tets[:, t] = np.stack([a, e, b, d], 1)
    tets = tets.reshape(-1, 4)

    eps = 1e-9*max(L, W, H)

    def classify(ctr):
        pid = np.full(ctr.shape[0], 2, dtype=np.int64)
        pid[ctr[:, 0] < eps] = 0
        pid[ctr[:, 0] > L - eps] = 1
        return pid
    classify.names = ["fixed", "loaded", "free"]
    return polymesh_from_tets(points, tets, classify)


# ------------------------------------------------------------------------------------------------
# OpenFOAM ASCII polyMesh reader (format as in run/fsiFoam/*/constant/polyMesh of the reference)
# ------------------------------------------------------------------------------------------------
def _strip_header(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    m = re.search(r"FoamFile\s*\{.*?\}", text, flags=re.S)
    if m:
        text = text[m.end():]
    return text


def _read_list_body(path):
    with open(path, "r") as f:
        text = _strip_header(f.read())
    m = re.search(r"(\d+)\s*\(", text)
    n = int(m.group(1))
    body = text[m.end():text.rindex(")")]
    return n, body


def read_polymesh(polyMeshDir):
    n, body = _read_list_body(os.path.join(polyMeshDir, "points"))
    pts = np.array(re.findall(r"[-+0-9.eE]+", body), dtype=np.float64).reshape(n, 3)
    n, body = _read_list_body(os.path.join(polyMeshDir, "owner"))
    owner = np.array(body.split(), dtype=np.int64)
    assert owner.shape[0] == n
    n, body = _read_list_body(os.path.join(polyMeshDir, "neighbour"))
    neighbour = np.array(body.split(), dtype=np.int64)
    assert neighbour.shape[0] == n
    nF, body = _read_list_body(os.path.join(polyMeshDir, "faces"))
    faceStart = [0]
    facePoints = []
    for m in re.finditer(r"(\d+)\s*\(([^)]*)\)", body):
        ids = m.group(2).split()
        assert len(ids) == int(m.group(1))
        facePoints.extend(ids)
        faceStart.append(len(facePoints))
    assert len(faceStart) == nF + 1
    from . import foamdict
    with open(os.path.join(polyMeshDir, "boundary"), "r") as f:
        text = _strip_header(f.read())
    # `N ( name { type ...; nFaces ...; startFace ...; } ... )`: patch names may carry '-' or '.' (inner-wall, symmetry-x)
    entries = foamdict.parse("boundary " + text[text.index("("):] + ";")["boundary"]
    patches = []
    for name, d in zip(entries[0::2], entries[1::2]):
        pt = dict(name=str(name), type=str(d.get("type", "patch")), start=int(d["startFace"]), size=int(d["nFaces"]))
        for k in ("myProcNo", "neighbProcNo"):          # processorPolyPatch entries (processorN/constant/polyMesh/boundary)
            if k in d:
                pt[k] = int(d[k])
        patches.append(pt)
    return PolyMesh(pts, np.array(faceStart), np.array(facePoints, dtype=np.int64), owner, neighbour, patches)


# ------------------------------------------------------------------------------------------------
# OpenFOAM ASCII polyMesh writer (what decomposePar / blockMesh write: constant/polyMesh/*)
# ------------------------------------------------------------------------------------------------
_BANNER = """/*--------------------------------*- C++ -*----------------------------------*\\
| =========                 |                                                 |
| \\\\      /  F ield         | foam-extend: Open Source CFD                    |
|  \\\\    /   O peration     | Version:     3.1                                |
|   \\\\  /    A nd           | Web:         http://www.extend-project.de       |
|    \\\\/     M anipulation  |                                                 |
\\*---------------------------------------------------------------------------*/
"""


def _header(cls, obj, location="constant/polyMesh", note=None):
    h = _BANNER + "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       %s;\n" % cls
    if note:
        h += "    note        \"%s\";\n" % note
    h += "    location    \"%s\";\n    object      %s;\n}\n" % (location, obj)
    return h + "// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n"


_FOOTER = "\n\n// ************************************************************************* //\n"


def write_label_list(path, a, obj, cls="labelList", note=None):
    a = np.asarray(a)
    with open(path, "w") as f:
        f.write(_header(cls, obj, note=note))
        f.write("%d\n(\n" % a.shape[0])
        f.write("\n".join(str(int(v)) for v in a))
        f.write("\n)" + _FOOTER)


def read_label_list(path):
    n, body = _read_list_body(path)
    a = np.array(body.split(), dtype=np.int64)
    assert a.shape[0] == n, path
    return a.astype(np.int32)


def write_polymesh(polyMeshDir, mesh):
    os.makedirs(polyMeshDir, exist_ok=True)
    note = "nPoints: %d nCells: %d nFaces: %d nInternalFaces: %d" % (mesh.nPoints, mesh.nCells, mesh.nFaces, mesh.nInternalFaces)
    with open(os.path.join(polyMeshDir, "points"), "w") as f:
        f.write(_header("vectorField", "points"))
        f.write("%d\n(\n" % mesh.nPoints)
        f.write("\n".join("(%s %s %s)" % (repr(float(p[0])), repr(float(p[1])), repr(float(p[2]))) for p in mesh.points))
        f.write("\n)" + _FOOTER)
    with open(os.path.join(polyMeshDir, "faces"), "w") as f:
        f.write(_header("faceList", "faces"))
        f.write("%d\n(\n" % mesh.nFaces)
        fs, fp = mesh.faceStart, mesh.facePoints
        f.write("\n".join("%d(%s)" % (fs[i + 1] - fs[i], " ".join(str(int(v)) for v in fp[fs[i]:fs[i + 1]])) for i in range(mesh.nFaces)))
        f.write("\n)" + _FOOTER)
    write_label_list(os.path.join(polyMeshDir, "owner"), mesh.owner, "owner", note=note)
    write_label_list(os.path.join(polyMeshDir, "neighbour"), mesh.neighbour, "neighbour", note=note)
    with open(os.path.join(polyMeshDir, "boundary"), "w") as f:
        f.write(_header("polyBoundaryMesh", "boundary"))
        f.write("%d\n(\n" % len(mesh.patches))
        for p in mesh.patches:
            f.write("    %s\n    {\n        type            %s;\n        nFaces          %d;\n        startFace       %d;\n"
                    % (p["name"], p["type"], p["size"], p["start"]))
            for k in ("myProcNo", "neighbProcNo"):
                if k in p:
                    f.write("        %-15s %d;\n" % (k, p[k]))
            f.write("    }\n")
        f.write(")" + _FOOTER)
