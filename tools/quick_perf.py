"""Scratch timing of the sweep on one GPU (not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import common as T
from umt_b200 import mesh as M

d = int(sys.argv[1]) if len(sys.argv) > 1 else 10
G = int(sys.argv[2]) if len(sys.argv) > 2 else 128
t = time.time(); m = M.tiled_mesh((d, d, d)); print("mesh", time.time() - t, m.nzones, flush=True)
t = time.time(); p = T.make_problem_3d(m, 2, 2, G, driver_like=True); print("problem", time.time() - t, "planes", p.sched["nHyperPlanes"][:8], flush=True)
ctx = T.gpu_context_3d(p)
unknowns = m.ncornr * p.NA * G
for i in range(4):
    ctx.sweep(savePsi=False)
    tm = ctx.last_times()
    print(tm, "unknowns/s(sweep kernel) %.3e  total %.3e" % (unknowns / tm["sweep_ms"] * 1e3, unknowns / tm["total_ms"] * 1e3), flush=True)
