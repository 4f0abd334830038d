"""assemble_D and gradients device time at 1 M and 8 M cells (bench.measure_assembly).  usage: python tools/asm_grad_time.py [1m]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
peak = 6549.8
for shape in ([(200, 50, 100)] if len(sys.argv) > 1 else [(200, 50, 100), (400, 100, 200)]):
    r = bench.measure_assembly(*shape, peak)
    print(shape, "assemble_D %.3f ms frac %.3f | gradients %.3f ms frac %.3f" % (
        r["assemble_D"]["ms"], r["assemble_D"]["frac"], r["gradients"]["ms"], r["gradients"]["frac"]), flush=True)
