"""In-situ latency of the cross-tile hand-offs of one forward DIC sweep (trace build of the library).
build: nvcc ... -DGPUPCG_DEEPTRACE -o fluidstructureinteraction_b200/libgpupcg_cuda_trace.so   (tools/build_trace.sh)
For every cross-tile entry: hop = delivery time at the consumer - publish time at the producer; slack = completion of
the consuming row - delivery.  Entries with ~zero slack are the ones the sweep is waiting for."""
import os, sys, ctypes as C
os.environ["GPUPCG_TILE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidstructureinteraction_b200 as pkg
from fluidstructureinteraction_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libgpupcg_cuda_trace.so")
nx, ny, nz = [int(v) for v in os.environ.get("CELLS", "200,50,100").split(",")]
s = pkg.cases.cantilever_system(nx, ny, nz)
g = pkg.GpuPCG(s["l"], s["u"], s["nCells"])
g.set_matrix(s["diag"], s["upper"])
info = g.plan_info()
t = g.time_kernel("dic", reps=2, flushL2=False)
nT, nR, nX = info["nTiles"], info["nRows"], info["extForward"]
buf = np.zeros(3*nT + nR + nX, dtype=np.uint64)
fn = g.lib.gpupcg_debug_deep_trace
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_longlong]
assert fn(g.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), buf.size) == 0
buf = buf.astype(np.int64)
tiles = buf[:3*nT].reshape(nT, 3); pub = buf[3*nT:3*nT + nR]; deliv = buf[3*nT + nR:]
t0 = tiles[:, 0].min()
# host copy of the schedule (same plan: same tile rows / mode as the handle used)
ex = pkg.plan_export(s["l"], s["u"], s["nCells"], tileRows=int(os.environ.get("GPUPCG_TILE_ROWS", "256")), tileMode=2)
assert ex["info"]["nTiles"] == nT and ex["info"]["extForward"] == nX, (ex["info"]["nTiles"], nT)
extCol = ex["extFCol"]; extPtr = ex["extFPtr"]; tileStart = ex["tileStart"]
sl = ex["slices"]; LcolTile = ex["LcolTile"]
# consumer row of every cross-tile entry
cons = np.full(nX, -1, dtype=np.int64)
rows = np.arange(nR)
tileOfRow = np.searchsorted(tileStart, rows, side="right") - 1
for j in range(int(sl[:, 3].max())):
    ok = j < sl[rows >> 5, 3]
    e = sl[rows >> 5, 2] + 32*j + (rows & 31)
    code = np.where(ok, LcolTile[np.where(ok, e, 0)], -1)
    m = code <= -2
    cons[extPtr[tileOfRow[m]] + (-2 - code[m])] = rows[m]
assert (cons >= 0).all()
hop = (deliv - pub[extCol])/1e3
slack = (pub[cons] - deliv)/1e3
print("dic ms %.4f  tiles %d  entries %d" % (t, nT, nX))
print("hop  (deliver - publish) us: p10 %.2f p50 %.2f p90 %.2f max %.2f" % tuple(np.percentile(hop, [10, 50, 90, 100])))
print("slack (consumer row done - deliver) us: p10 %.2f p50 %.2f p90 %.2f" % tuple(np.percentile(slack, [10, 50, 90])))
crit = slack < np.percentile(slack, 10)
print("critical entries (lowest 10%% slack, < %.2f us): hop p10 %.2f p50 %.2f p90 %.2f" % ((np.percentile(slack, 10),) + tuple(np.percentile(hop[crit], [10, 50, 90]))))
neg = hop < 0
print("entries delivered 'before' publish stamp (value was there when first polled):", int(neg.sum()))
# per tile: time from first needed delivery to finish, rounds
fin = tiles[:, 2]; st = tiles[:, 1]
first = np.array([deliv[extPtr[k]:extPtr[k + 1]].max() if extPtr[k + 1] > extPtr[k] else st[k] for k in range(nT)])
print("tile: last delivery -> finish us: p10 %.2f p50 %.2f p90 %.2f" % tuple(np.percentile((fin - first)/1e3, [10, 50, 90])))
firstd = np.array([deliv[extPtr[k]:extPtr[k + 1]].min() if extPtr[k + 1] > extPtr[k] else st[k] for k in range(nT)])
print("tile: first delivery -> finish us: p10 %.2f p50 %.2f p90 %.2f" % tuple(np.percentile((fin - firstd)/1e3, [10, 50, 90])))
# publish pace inside a tile: time between consecutive rounds (rows sorted by round) for a few mid tiles
for k in (nT//2, nT//2 + 7):
    r0, r1 = tileStart[k], tileStart[k + 1]
    pt = np.sort(np.unique(pub[r0:r1][pub[r0:r1] > 0]))
    print("tile %d: %d distinct publish stamps, span %.2f us, median gap %.0f ns, max gap %.0f ns" % (k, pt.size, (pt[-1] - pt[0])/1e3, np.median(np.diff(pt)), np.diff(pt).max()))
