"""Short run for ncu: GPU assembly of the D equation + vertex-based gradients on a hex box.
usage: python tools/profile_assembly.py [nx,ny,nz] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("GPUPCG_LIB"):
    from fluidstructureinteraction_b200 import capi
    capi.LIB_PATH = os.environ["GPUPCG_LIB"]
import bench
nx, ny, nz = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "200,50,100").split(",")]
r = bench.measure_assembly(nx, ny, nz, 6549.8)
for k, v in r.items():
    print(k, "%.4f ms  %.1f GB/s  frac %.3f" % (v["ms"], v["achieved_gbs"], v["frac"]))
