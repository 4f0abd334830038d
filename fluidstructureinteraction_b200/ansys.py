"""ANSYS `.ans` (PREP7 N / EN records) -> polyMesh: what [EXT] `ansysToFoam` does for the reference's thermal case
(run/thermalStressFoam/flange/flange.ans, run/thermalStressFoam/flange/Allrun: `ansysToFoam flange.ans -scale 0.001`).

  N,id,x,y,z              node (missing trailing coordinates are zero)
  EN,id,n1,...,n8         8-node brick (ET 45 = SOLID45); degenerate bricks become prisms (n3 == n4, n7 == n8), pyramids
                          (n5 == n6 == n7 == n8) and tetrahedra (n3 == n4, n5 == ... == n8), as ansysToFoam's element rules do

Cells in element order, points in node order (renumbered compactly by ascending node id), faces from the cell shapes in
upper-triangular order (polyMeshFromShapeMesh.C, see blockmesh.polymesh_from_shapes); every boundary face goes to one patch
`defaultFaces` -- the utility knows no boundary sets.  (The shipped constant/polyMesh/boundary splits the 3268 boundary faces
into patch1..patch4 = 2440 / 348 / 96 / 384; which face belongs to which is not recorded in the case.)  INPUT handling."""
import numpy as np

from . import blockmesh


def read_ans(path, scale=1.0):
    nodes, elems = {}, []
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if line.startswith("N,"):
                t = [v.strip() for v in line.split(",")]
                xyz = [float(v) if v else 0.0 for v in t[2:5]]
                while len(xyz) < 3:
                    xyz.append(0.0)
                nodes[int(t[1])] = xyz
            elif line.startswith("EN,"):
                t = [v.strip() for v in line.split(",")]
                elems.append([int(v) for v in t[2:10]])
    ids = np.array(sorted(nodes), dtype=np.int64)
    lookup = {int(n): i for i, n in enumerate(ids)}
    points = np.array([nodes[int(n)] for n in ids], dtype=np.float64)*scale
    shapes = []
    for e in elems:
        n = [lookup[v] for v in e]
        if len(set(n)) == 8:
            shapes.append(("hex", n))
        elif n[2] == n[3] and n[6] == n[7] and len(set(n)) == 6:
            shapes.append(("prism", [n[0], n[1], n[2], n[4], n[5], n[6]]))
        elif n[4] == n[5] == n[6] == n[7] and len(set(n)) == 5:
            shapes.append(("pyr", [n[0], n[1], n[2], n[3], n[4]]))
        elif n[2] == n[3] and n[4] == n[5] == n[6] == n[7] and len(set(n)) == 4:
            shapes.append(("tet", [n[0], n[1], n[2], n[4]]))
        else:
            raise ValueError("unsupported degenerate element %r" % (e,))
    return points, shapes


def ans_to_polymesh(path, scale=1.0):
    points, shapes = read_ans(path, scale)
    return blockmesh.polymesh_from_shapes(points, shapes, [], [], [], defaultPatch=("defaultFaces", "patch"))


def boundary_surfaces(mesh, geom, featureAngleDeg=80.0):
    """[EXT autoPatch-like] labels the boundary faces by connected surface: two faces that share an edge belong together when
    their normals differ by less than featureAngleDeg.  Returns (label per boundary face, sizes).  On the flange this recovers
    two of the four shipped patches by their sizes -- the flat face y = -7.5 mm (348 faces = patch2, fixedValue 273 K) and the
    central bore (384 faces = patch4, 573 K); the shipped split of the remaining faces into patch1 / patch3 is not recorded."""
    import collections
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    nIF = mesh.nInternalFaces
    nBF = mesh.nFaces - nIF
    nrm = geom.Sf[nIF:]/geom.magSf[nIF:, None]
    edges = collections.defaultdict(list)
    for i in range(nBF):
        pts = mesh.facePoints[mesh.faceStart[nIF + i]:mesh.faceStart[nIF + i + 1]]
        for a, b in zip(pts, np.roll(pts, -1)):
            edges[(min(int(a), int(b)), max(int(a), int(b)))].append(i)
    c = np.cos(np.radians(featureAngleDeg))
    I, J = [], []
    for fl in edges.values():
        if len(fl) == 2 and nrm[fl[0]]@nrm[fl[1]] > c:
            I.append(fl[0]); J.append(fl[1])
    n, lab = connected_components(coo_matrix((np.ones(len(I)), (I, J)), shape=(nBF, nBF)), directed=False)
    return lab, np.bincount(lab, minlength=n)
