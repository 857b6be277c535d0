"""Scratch: sweep-kernel time vs angle-batch size / stagger / zones per item at a large domain (one GPU)."""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from umt_b200 import mesh as M, problem as PR, teton

d = int(sys.argv[1]) if len(sys.argv) > 1 else 20
G = int(sys.argv[2]) if len(sys.argv) > 2 else 128
mesh = M.tiled_mesh((d, d, d))
ctx = teton.SweepContext.from_mesh(mesh, G)
ctx.compute_geometry(mesh.px)
NA = ctx.build_product_quadrature(2, 2, 1)
ctx.build_schedule()
tau = PR.tau()
ctx.upload_state(None, None, np.full((mesh.nzones, G), tau), np.zeros((mesh.ncornr, G)), tau)
ctx.init_teton(np.full(mesh.nzones, PR.TR0), PR.group_bounds(G), PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(3), 0.0)
ctx.init_radiation_field()
unknowns = mesh.ncornr * NA * G
configs = [dict(UMT_ANGLE_BATCH=str(k), UMT_BATCH_STAGGER=str(s)) for k in (1, 2, 4, 8, 16, 32) for s in ((0.5,) if k in (1, 32) else (0.25, 0.5, 1.0))]
for extra in sys.argv[3:]:
    k, v = extra.split("=")
    for c in configs:
        c[k] = v
for cfg in configs:
    os.environ.update(cfg)
    ctx.build_schedule()
    ts = []
    for i in range(4):
        ctx.sweep(False)
        ts.append(ctx.last_times()["sweep_ms"])
    best = min(ts[1:])
    print(cfg, "sweep_ms %.2f  unknowns/s %.3e" % (best, unknowns / best * 1e3), flush=True)
