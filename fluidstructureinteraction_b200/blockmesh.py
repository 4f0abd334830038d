"""blockMesh restated: `constant/polyMesh/blockMeshDict` -> polyMesh in foam-extend's numbering (INPUT generator).

The reference ships its first case (run/stressFoam/plateHole, BASELINE.json configs[0]) only as a blockMeshDict with arc
edges (run/stressFoam/plateHole/plateHole/constant/polyMesh/blockMeshDict:19-106); foam-extend's `blockMesh` utility turns
it into points/faces/owner/neighbour/boundary.  blockMesh is [EXT] (foam-extend-3.1 src/mesh/blockMesh, absent here), so
this follows its published algorithm:

  curvedEdges/arcEdge.C       centre/radius/angle of the circle through (start, mid, end); position(lambda) rotates the start
                              radius by lambda*angle about the circle's axis
  curvedEdges/lineDivide.C    points of an edge at the grading's divisions (uniform or geometric progression)
  block/blockCreate.C         block points: edge-weighted transfinite interpolation of the 12 edges ("importance factors"),
                              cells k-j-i with i fastest, boundary quads per block face
  blockMesh/blockMeshMerge.C  coincident points of neighbouring blocks merged, numbered by first appearance
  polyMesh/polyMeshFromShapeMesh.C   faces from the hex shapes: per cell, the faces shared with higher-numbered cells in
                              ascending neighbour order (upper-triangular), face loop = the owner's hex-model face; then the
                              patches in dictionary order

What could differ from a real blockMesh run is confined to rounding of point coordinates and to the order of faces inside
a boundary patch -- neither moves a discretisation-level check such as the Kirsch errors (oracle/ASSUMPTIONS.md #10).
"""
import math
import re

import numpy as np

from .meshes import PolyMesh, _strip_header

# hex cell model [EXT cellModels]: faces x-min, x-max, y-min, y-max, z-min, z-max, outward
HEX_FACES = np.array([[0, 4, 7, 3], [1, 2, 6, 5], [0, 1, 5, 4], [3, 7, 6, 2], [0, 3, 2, 1], [4, 5, 6, 7]])
# block edges 0-11 as (start, end) block-vertex pairs: x1..x4, y1..y4, z1..z4 [EXT blockDescriptorEdges.C]
BLOCK_EDGES = [(0, 1), (3, 2), (7, 6), (4, 5), (0, 3), (1, 2), (5, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]


# ---------------------------------------------------------------------------------------------------------------------
# dictionary reader (the subset blockMesh dictionaries of the reference use: vertices, blocks hex/simpleGrading, edges arc,
# patches)
# ---------------------------------------------------------------------------------------------------------------------
def _balanced(text, start):
    """text[start] == '(' -> index one past its matching ')'."""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced parentheses")


def _section(text, name):
    m = re.search(r"\b" + name + r"\s*\(", text)
    if not m:
        return ""
    end = _balanced(text, m.end() - 1)
    return text[m.end():end - 1]


def _floats(s):
    return [float(t) for t in s.split()]


def parse_block_mesh_dict(text):
    """-> dict(scale, vertices [[x,y,z]], blocks [dict(v[8], n[3], grading[3])], edges [dict(type, a, b, point)],
    patches [dict(type, name, faces [[4]])])."""
    text = _strip_header(text)
    m = re.search(r"convertToMeters\s+([-+0-9.eE]+)\s*;", text)
    scale = float(m.group(1)) if m else 1.0
    vertices = [_floats(t) for t in re.findall(r"\(([^()]*)\)", _section(text, "vertices"))]
    blocks = []
    for m in re.finditer(r"hex\s*\(([^()]*)\)\s*(?:\w+\s*)?\(([^()]*)\)\s*simpleGrading\s*\(([^()]*)\)", _section(text, "blocks")):
        blocks.append(dict(v=[int(t) for t in m.group(1).split()], n=[int(t) for t in m.group(2).split()],
                           grading=_floats(m.group(3))))
    edges = []
    for m in re.finditer(r"(arc)\s+(\d+)\s+(\d+)\s*\(([^()]*)\)", _section(text, "edges")):
        edges.append(dict(type=m.group(1), a=int(m.group(2)), b=int(m.group(3)), point=_floats(m.group(4))))
    patches = []
    body = _section(text, "patches")
    for m in re.finditer(r"(\w+)\s+(\w+)\s*\(", body):
        end = _balanced(body, m.end() - 1)
        faces = [[int(t) for t in f.split()] for f in re.findall(r"\(([^()]*)\)", body[m.end():end - 1])]
        patches.append(dict(type=m.group(1), name=m.group(2), faces=faces))
    return dict(scale=scale, vertices=vertices, blocks=blocks, edges=edges, patches=patches)


def read_block_mesh_dict(path):
    with open(path, "r") as f:
        return parse_block_mesh_dict(f.read())


# ---------------------------------------------------------------------------------------------------------------------
# curved edges
# ---------------------------------------------------------------------------------------------------------------------
class _LineEdge(object):
    def __init__(self, p1, p2):
        self.p1, self.p2 = p1, p2

    def position(self, lam):
        return self.p1 + lam*(self.p2 - self.p1)


class _ArcEdge(object):
    """[EXT arcEdge.C] calcAngle(): circle through p1, p2 (any point on the arc), p3."""

    def __init__(self, p1, p2, p3):
        a, b = p2 - p1, p3 - p1
        asqr, bsqr, adotb = a@a, b@b, a@b
        denom = asqr*bsqr - adotb*adotb
        if np.linalg.norm(np.cross(a, b)) < 1e-10 or abs(denom) < 1e-300:
            raise ValueError("arc edge: the three points are collinear")
        fact = 0.5*(bsqr - adotb)/denom
        centre = 0.5*a + fact*np.cross(np.cross(a, b), a)
        centre = centre + p1
        r1, r2, r3 = p1 - centre, p2 - centre, p3 - centre
        ang = math.degrees(math.acos((r3@r1)/(np.linalg.norm(r3)*np.linalg.norm(r1))))
        if np.cross(r1, r2)@np.cross(r1, r3) < 0.0:
            ang = 360.0 - ang
        if ang <= 180.0:
            axis = np.cross(r1, r3)
            if np.linalg.norm(axis)/(np.linalg.norm(r1)*np.linalg.norm(r3)) < 0.001:
                axis = np.cross(r1, r2)
        else:
            axis = np.cross(r3, r1)
        self.angle, self.radius, self.centre = ang, np.linalg.norm(r3), centre
        # cylindricalCS(origin = centre, axis, direction = r1), angles in degrees
        e3 = axis/np.linalg.norm(axis)
        e1 = r1 - (r1@e3)*e3
        e1 = e1/np.linalg.norm(e1)
        self.e1, self.e2, self.e3 = e1, np.cross(e3, e1), e3

    def position(self, lam):
        t = math.radians(lam*self.angle)
        return self.centre + (self.radius*math.cos(t))*self.e1 + (self.radius*math.sin(t))*self.e2


def _line_divide(edge, n, expansion):
    """[EXT lineDivide.C] points (n+1,3) and divisions (n+1,) of an edge cut into n cells with last/first cell-size ratio
    `expansion` (simpleGrading)."""
    pts = np.empty((n + 1, 3)); div = np.empty(n + 1)
    if expansion == 1.0 or n == 1:
        y = 1.0/n
        for i in range(n + 1):
            pts[i] = edge.position(i/float(n)); div[i] = y*i
    else:
        r = expansion**(1.0/(n - 1))
        pts[0] = edge.position(0.0); div[0] = 0.0
        for i in range(1, n + 1):
            lam = (1.0 - r**i)/(1.0 - r**n)
            pts[i] = edge.position(lam); div[i] = lam
    return pts, div


# ---------------------------------------------------------------------------------------------------------------------
# blocks
# ---------------------------------------------------------------------------------------------------------------------
def _block_points(corner, edgePts, edgeW, n):
    """[EXT blockCreate.C block::createPoints]: points of one block, index i + (ni+1)*(j + (nj+1)*k)."""
    ni, nj, nk = n
    p000, p100, p110, p010, p001, p101, p111, p011 = corner
    I, J, K = np.meshgrid(np.arange(ni + 1), np.arange(nj + 1), np.arange(nk + 1), indexing="ij")
    # flatten with i fastest
    I = I.transpose(2, 1, 0).ravel(); J = J.transpose(2, 1, 0).ravel(); K = K.transpose(2, 1, 0).ravel()
    idx = (I, I, I, I, J, J, J, J, K, K, K, K)
    w = [edgeW[e][idx[e]] for e in range(12)]
    p = [edgePts[e][idx[e]] for e in range(12)]
    c = lambda a: a[:, None]
    ends = [(p000, p100), (p010, p110), (p011, p111), (p001, p101),
            (p000, p010), (p100, p110), (p101, p111), (p001, p011),
            (p000, p001), (p100, p101), (p110, p111), (p010, p011)]
    edge = [s + (e - s)*c(w[k]) for k, (s, e) in enumerate(ends)]
    imp = [
        (1.0 - w[0])*(1.0 - w[4])*(1.0 - w[8]) + w[0]*(1.0 - w[5])*(1.0 - w[9]),
        (1.0 - w[1])*w[4]*(1.0 - w[11]) + w[1]*w[5]*(1.0 - w[10]),
        (1.0 - w[2])*w[7]*w[11] + w[2]*w[6]*w[10],
        (1.0 - w[3])*(1.0 - w[7])*w[8] + w[3]*(1.0 - w[6])*w[9],
        (1.0 - w[4])*(1.0 - w[0])*(1.0 - w[8]) + w[4]*(1.0 - w[1])*(1.0 - w[11]),
        (1.0 - w[5])*w[0]*(1.0 - w[9]) + w[5]*w[1]*(1.0 - w[10]),
        (1.0 - w[6])*w[3]*w[9] + w[6]*w[2]*w[10],
        (1.0 - w[7])*(1.0 - w[3])*w[8] + w[7]*(1.0 - w[2])*w[11],
        (1.0 - w[8])*(1.0 - w[0])*(1.0 - w[4]) + w[8]*(1.0 - w[3])*(1.0 - w[7]),
        (1.0 - w[9])*w[0]*(1.0 - w[5]) + w[9]*w[3]*(1.0 - w[6]),
        (1.0 - w[10])*w[1]*w[5] + w[10]*w[2]*w[6],
        (1.0 - w[11])*(1.0 - w[1])*w[4] + w[11]*(1.0 - w[2])*w[7],
    ]
    for g in (0, 4, 8):
        tot = imp[g] + imp[g + 1] + imp[g + 2] + imp[g + 3]
        for k in range(g, g + 4):
            imp[k] = imp[k]/tot
    cor = [c(imp[k])*(p[k] - edge[k]) for k in range(12)]
    edge = [edge[k]*c(imp[k]) for k in range(12)]
    v = edge[0]
    for k in range(1, 12):
        v = v + edge[k]
    v = v/3.0
    s = cor[0]
    for k in range(1, 12):
        s = s + cor[k]
    return v + s


def _boundary_quads(n, off):
    """[EXT blockBoundary.C] the quads of the six block faces in terms of global-before-merge point ids; returns for face
    0..5 the array (m,4) in blockMesh's face order."""
    ni, nj, nk = n

    def V(i, j, k):
        return off + i + (ni + 1)*(j + (nj + 1)*k)

    out = []
    k, j = [a.ravel() for a in np.meshgrid(np.arange(nk), np.arange(nj), indexing="ij")]
    z = np.zeros_like(j)
    out.append(np.stack([V(z, j, k), V(z, j, k + 1), V(z, j + 1, k + 1), V(z, j + 1, k)], 1))
    out.append(np.stack([V(z + ni, j, k), V(z + ni, j + 1, k), V(z + ni, j + 1, k + 1), V(z + ni, j, k + 1)], 1))
    i, k = [a.ravel() for a in np.meshgrid(np.arange(ni), np.arange(nk), indexing="ij")]
    z = np.zeros_like(i)
    out.append(np.stack([V(i, z, k), V(i + 1, z, k), V(i + 1, z, k + 1), V(i, z, k + 1)], 1))
    out.append(np.stack([V(i, z + nj, k), V(i, z + nj, k + 1), V(i + 1, z + nj, k + 1), V(i + 1, z + nj, k)], 1))
    i, j = [a.ravel() for a in np.meshgrid(np.arange(ni), np.arange(nj), indexing="ij")]
    z = np.zeros_like(i)
    out.append(np.stack([V(i, j, z), V(i, j + 1, z), V(i + 1, j + 1, z), V(i + 1, j, z)], 1))
    out.append(np.stack([V(i, j, z + nk), V(i + 1, j, z + nk), V(i + 1, j + 1, z + nk), V(i, j + 1, z + nk)], 1))
    return out


# cell models [EXT cellModels]: faces of a shape in terms of its point list, outward
SHAPE_FACES = {
    "hex": [[0, 4, 7, 3], [1, 2, 6, 5], [0, 1, 5, 4], [3, 7, 6, 2], [0, 3, 2, 1], [4, 5, 6, 7]],
    "prism": [[0, 2, 1], [3, 4, 5], [0, 3, 5, 2], [1, 2, 5, 4], [0, 1, 4, 3]],
    "pyr": [[0, 3, 2, 1], [0, 4, 3], [3, 4, 2], [1, 2, 4], [0, 1, 4]],
    "tet": [[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]],
}


def polymesh_from_shapes(points, shapes, patchFaces, patchNames, patchTypes, defaultPatch=None):
    """[EXT polyMeshFromShapeMesh.C] shapes: list of (model name, point ids); patchFaces: per patch a list of faces (point
    ids, any rotation / orientation).  Internal faces upper-triangular (per cell, ascending neighbour), face loop = the
    owner's model face; boundary faces in the given order, loop = the inside cell's model face; boundary faces that belong
    to no patch go to `defaultPatch` = (name, type) in cell order (error if None)."""
    faces, cell = [], []
    for c, (model, pts) in enumerate(shapes):
        for mf in SHAPE_FACES[model]:
            faces.append([pts[i] for i in mf]); cell.append(c)
    nC = len(shapes)
    cell = np.array(cell, dtype=np.int64)
    keys = [tuple(sorted(f)) for f in faces]
    first = {}
    pairs = []                                   # (owner face index, neighbour face index)
    for i, k in enumerate(keys):
        j = first.pop(k, None)
        if j is None:
            first[k] = i
        else:
            pairs.append((j, i))
    pairs.sort(key=lambda ab: (cell[ab[0]], cell[ab[1]]))
    own = [int(cell[a]) for a, b in pairs]; nei = [int(cell[b]) for a, b in pairs]
    faceLoops = [faces[a] for a, b in pairs]
    bnd = dict(first)                            # key -> face index of the free boundary faces
    patches = []
    start = len(own)
    for name, ptype, flist in zip(patchNames, patchTypes, patchFaces):
        n0 = len(own)
        for f in flist:
            i = bnd.pop(tuple(sorted(int(v) for v in f)), None)
            if i is None:
                raise ValueError("patch %s: face %s is not a free boundary face" % (name, list(f)))
            own.append(int(cell[i])); faceLoops.append(faces[i])
        patches.append(dict(name=name, type=ptype, start=start, size=len(own) - n0))
        start = len(own)
    if bnd:
        if defaultPatch is None:
            raise ValueError("%d boundary faces belong to no patch (blockMesh would collect them in defaultFaces)" % len(bnd))
        rest = sorted(bnd.values())
        for i in rest:
            own.append(int(cell[i])); faceLoops.append(faces[i])
        patches.append(dict(name=defaultPatch[0], type=defaultPatch[1], start=start, size=len(rest)))
    faceStart = np.zeros(len(faceLoops) + 1, dtype=np.int64)
    np.cumsum([len(f) for f in faceLoops], out=faceStart[1:])
    facePoints = np.array([v for f in faceLoops for v in f], dtype=np.int64)
    return PolyMesh(points, faceStart, facePoints, np.array(own, dtype=np.int64), np.array(nei, dtype=np.int64), patches, nC)


def polymesh_from_hexes(points, hexes, patchQuads, patchNames, patchTypes):
    """Hex-only front end of polymesh_from_shapes (vectorised: blocks of a blockMesh run into millions of cells)."""
    nC = hexes.shape[0]
    fv = hexes[:, HEX_FACES].reshape(-1, 4)                 # face 6c + m of cell c
    cell = np.repeat(np.arange(nC, dtype=np.int64), 6)
    key = np.sort(fv, axis=1)
    order = np.lexsort((cell, key[:, 3], key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    first = np.ones(order.shape[0], dtype=bool); first[1:] = ~same
    paired = np.zeros(order.shape[0], dtype=bool); paired[:-1] = same
    sel = first & paired
    iA, iB = order[sel], order[np.flatnonzero(sel) + 1]
    own, nei, fpts = cell[iA], cell[iB], fv[iA]
    srt = np.lexsort((nei, own))
    own, nei, fpts = own[srt], nei[srt], fpts[srt]
    # boundary: look each patch quad up among the unpaired cell faces
    iBnd = order[first & ~paired]
    lut = {tuple(k): int(f) for k, f in zip(key[iBnd], iBnd)}
    owners, faces, patches = [own], [fpts], []
    start = own.shape[0]
    used = set()
    for name, ptype, quads in zip(patchNames, patchTypes, patchQuads):
        ids = []
        for q in np.sort(np.asarray(quads, dtype=np.int64), axis=1):
            f = lut.get(tuple(q))
            if f is None or f in used:
                raise ValueError("patch %s: face %s is not a free boundary face" % (name, q))
            used.add(f); ids.append(f)
        ids = np.array(ids, dtype=np.int64)
        patches.append(dict(name=name, type=ptype, start=start, size=int(ids.shape[0])))
        owners.append(cell[ids]); faces.append(fv[ids])
        start += int(ids.shape[0])
    if len(used) != iBnd.shape[0]:
        raise ValueError("%d boundary faces belong to no patch (blockMesh would collect them in defaultFaces)"
                         % (iBnd.shape[0] - len(used)))
    owner = np.concatenate(owners); quads = np.concatenate(faces)
    faceStart = np.arange(0, 4*quads.shape[0] + 1, 4, dtype=np.int32)
    return PolyMesh(points, faceStart, quads.ravel(), owner, nei, patches, nC)


def block_mesh(d, refine=1):
    """The polyMesh blockMesh writes for the parsed dictionary `d`; `refine` multiplies every block's cell counts in the
    directions that have more than one cell (a convergence study keeps the 2-D layer one cell thick)."""
    V = np.array(d["vertices"], dtype=np.float64)
    curved = {}
    for e in d["edges"]:
        curved[(e["a"], e["b"])] = np.array(e["point"], dtype=np.float64)
    allPts, hexes, bquads, offs = [], [], [], []
    off = 0
    for b in d["blocks"]:
        n = [c*refine if c > 1 else c for c in b["n"]]
        ni, nj, nk = n
        corner = [V[v] for v in b["v"]]
        edgePts, edgeW = [], []
        for e, (s, t) in enumerate(BLOCK_EDGES):
            vs, vt = b["v"][s], b["v"][t]
            ne = n[e // 4]; g = b["grading"][e // 4]
            if (vs, vt) in curved:
                edge = _ArcEdge(V[vs], curved[(vs, vt)], V[vt]); pts, w = _line_divide(edge, ne, g)
            elif (vt, vs) in curved:                        # edge defined in the opposite sense: walk it backwards
                edge = _ArcEdge(V[vt], curved[(vt, vs)], V[vs]); pts, w = _line_divide(edge, ne, 1.0/g)
                pts = pts[::-1].copy(); w = (1.0 - w)[::-1].copy()
            else:
                pts, w = _line_divide(_LineEdge(V[vs], V[vt]), ne, g)
            edgePts.append(pts); edgeW.append(w)
        allPts.append(_block_points(corner, edgePts, edgeW, n))
        i, j, k = [a.transpose(2, 1, 0).ravel() for a in np.meshgrid(np.arange(ni), np.arange(nj), np.arange(nk), indexing="ij")]

        def P(a, bb, c, off=off, ni=ni, nj=nj):
            return off + a + (ni + 1)*(bb + (nj + 1)*c)
        hexes.append(np.stack([P(i, j, k), P(i + 1, j, k), P(i + 1, j + 1, k), P(i, j + 1, k),
                               P(i, j, k + 1), P(i + 1, j, k + 1), P(i + 1, j + 1, k + 1), P(i, j + 1, k + 1)], 1))
        bquads.append(_boundary_quads(n, off)); offs.append(off)
        off += (ni + 1)*(nj + 1)*(nk + 1)
    pts = np.concatenate(allPts)*d["scale"]
    hexes = np.concatenate(hexes)
    # merge coincident points of neighbouring blocks: numbered by first appearance, coordinates of the last block win
    span = pts.max(axis=0) - pts.min(axis=0)
    tol = 1e-6*float(span.max())/max(max(max(b["n"]) for b in d["blocks"])*refine, 1)
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    pairs = cKDTree(pts).query_pairs(tol, output_type="ndarray")
    nAll = pts.shape[0]
    g = coo_matrix((np.ones(pairs.shape[0]), (pairs[:, 0], pairs[:, 1])), shape=(nAll, nAll))
    _, inv = connected_components(g, directed=False)
    _, firstIdx = np.unique(inv, return_index=True)
    rank = np.empty(firstIdx.shape[0], dtype=np.int64)
    rank[np.argsort(firstIdx, kind="stable")] = np.arange(firstIdx.shape[0])
    merged = rank[inv]
    points = np.empty((firstIdx.shape[0], 3)); points[merged] = pts       # later (higher block) entries overwrite
    hexes = merged[hexes]
    # patches: each dictionary face (four block-vertex labels) names one face of one block
    names, types, quads = [], [], []
    for p in d["patches"]:
        qs = []
        for f in p["faces"]:
            found = False
            for bi, b in enumerate(d["blocks"]):
                bv = np.array(b["v"])
                for m in range(6):
                    if sorted(bv[HEX_FACES[m]].tolist()) == sorted(f):
                        qs.append(merged[bquads[bi][m]]); found = True
                        break
                if found:
                    break
            if not found:
                raise ValueError("patch %s: face %s matches no block face" % (p["name"], f))
        names.append(p["name"]); types.append(p["type"]); quads.append(np.concatenate(qs))
    return polymesh_from_hexes(points, hexes, quads, names, types)
