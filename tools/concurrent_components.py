"""Experiment: the three component systems of one fvMatrix<vector>::solve as three CONCURRENT scalar solves (three handles,
three streams, three host threads) against the component-batched solve on one stream.  The sweeps are latency-bound (SMs
mostly idle), the Amul / update kernels bandwidth-bound: if the streams drift out of phase the two kinds overlap.

  python tools/concurrent_components.py [nx ny nz]
"""
import os, sys, threading, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import cases

LIB_ONLY = "--lib-only" in sys.argv          # only the in-library A/B (one handle): large meshes
if LIB_ONLY:
    sys.argv.remove("--lib-only")
nx, ny, nz = [int(v) for v in sys.argv[1:4]] if len(sys.argv) >= 4 and sys.argv[1].isdigit() else (200, 50, 100)
TOL, RELTOL = 1e-9, float(os.environ.get("RELTOL", "0.01"))
TR = (0.3, -0.2, 1.0)
if "--tet" in sys.argv:                     # unstructured: jittered tet box (6 tets per hex), tile sweeps instead of pencils
    from fluidstructureinteraction_b200 import meshes
    m = meshes.tet_box(nx, ny, nz, jitter=0.15)
    l, u = m.lower_upper()
    up = -(1.0 + 0.3*np.sin(0.1*np.arange(l.shape[0])))
    dg = np.zeros(m.nCells); np.subtract.at(dg, l, up); np.subtract.at(dg, u, up)
    s = dict(l=l, u=u, nCells=m.nCells, upper=up, diag=dg*1.000001, b=np.cos(0.01*np.arange(m.nCells)))
    print("tet box:", m.nCells, "cells", l.shape[0], "internal faces", flush=True)
else:
    s = cases.cantilever_system(nx, ny, nz)
n = s["nCells"]
torch.cuda.set_device(0)
bs = [torch.from_numpy(s["b"]*t).cuda() for t in TR]

# batched (GPUPCG_VECTOR_CONCURRENT toggled per call: the handle holds both the batched matrix and the component systems)
g = pkg.GpuPCG(s["l"], s["u"], n, device=0)
os.environ["GPUPCG_VECTOR_CONCURRENT"] = "1"
g.set_matrix_vector(np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1)), s["upper"])
os.environ["GPUPCG_VECTOR_CONCURRENT"] = "0"
d_bv = torch.stack(bs, dim=1).contiguous()
d_xv = torch.zeros((n, 3), dtype=torch.float64, device="cuda")


def batched():
    d_xv.zero_()
    torch.cuda.synchronize()
    t = time.perf_counter()
    perfs, _ = g.solve_vector_device(3, d_xv.data_ptr(), d_bv.data_ptr(), TOL, RELTOL, 0, 1000)
    torch.cuda.synchronize()
    return time.perf_counter() - t, [p["nIterations"] for p in perfs]


for _ in range(3):
    batched()
tb = [batched() for _ in range(5)]
print("batched  : %.2f ms" % (1e3*np.median([t for t, _ in tb])), tb[0][1], flush=True)
xb = d_xv.cpu().numpy()
os.environ["GPUPCG_VECTOR_CONCURRENT"] = "1"
for _ in range(3):
    batched()
tl = [batched() for _ in range(5)]
print("library concurrent (gpupcg_solve_vector_device, GPUPCG_VECTOR_CONCURRENT=1): %.2f ms" % (1e3*np.median([t for t, _ in tl])), tl[0][1], flush=True)
xl = d_xv.cpu().numpy()
print("library concurrent vs batched, max rel diff:", float(np.abs(xl - xb).max()/np.abs(xb).max()), flush=True)
os.environ["GPUPCG_VECTOR_CONCURRENT"] = "0"
if LIB_ONLY:
    sys.exit(0)

# three handles
hs = []
for k in range(3):
    h = pkg.GpuPCG(s["l"], s["u"], n, device=0)
    h.set_matrix(s["diag"], s["upper"])
    hs.append(h)
xs = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(3)]
its = [0, 0, 0]


def one(k):
    torch.cuda.set_device(0)
    perf, _ = hs[k].solve_device(xs[k].data_ptr(), bs[k].data_ptr(), TOL, RELTOL, 0, 1000)
    its[k] = perf["nIterations"]


def concurrent(stagger_us=0):
    for x in xs:
        x.zero_()
    torch.cuda.synchronize()
    t = time.perf_counter()
    th = []
    for k in range(3):
        th.append(threading.Thread(target=one, args=(k,)))
        th[-1].start()
        if stagger_us:
            time.sleep(stagger_us*1e-6)
    for q in th:
        q.join()
    torch.cuda.synchronize()
    return time.perf_counter() - t


def serial():
    for x in xs:
        x.zero_()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for k in range(3):
        one(k)
    torch.cuda.synchronize()
    return time.perf_counter() - t


for _ in range(2):
    serial()
print("serial   : %.2f ms" % (1e3*np.median([serial() for _ in range(5)])), its, flush=True)
print("serial vs batched, max rel diff per component:", [float(np.abs(xs[k].cpu().numpy() - xb[:, k]).max()/np.abs(xb).max()) for k in range(3)], flush=True)
xser = [x.cpu().numpy().copy() for x in xs]
for st in (0, 100, 200):
    for _ in range(2):
        concurrent(st)
    print("concurrent (stagger %3d us): %.2f ms" % (st, 1e3*np.median([concurrent(st) for _ in range(5)])), its, flush=True)
print("concurrent vs serial bitwise:", [bool(np.array_equal(xs[k].cpu().numpy(), xser[k])) for k in range(3)])
print("library concurrent vs python-threaded scalar handles bitwise:", [bool(np.array_equal(xs[k].cpu().numpy(), xl[:, k])) for k in range(3)])
print("concurrent vs batched, max rel diff:", [float(np.abs(xs[k].cpu().numpy() - xb[:, k]).max()/np.abs(xb).max()) for k in range(3)])
