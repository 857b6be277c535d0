"""Scratch: sweep-kernel time at one domain size, for A/B runs of library builds (UMT_LIB) and env knobs."""
import os, sys
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
ts = []
for i in range(5):
    ctx.sweep(False)
    ts.append(ctx.last_times())
best = min(t["sweep_ms"] for t in ts[1:])
print(os.environ.get("UMT_LIB", "default"), {k: v for k, v in os.environ.items() if k.startswith("UMT_") and k != "UMT_LIB"},
      "d=%d G=%d sweep_ms %.2f phi_ms %.2f unknowns/s(sweep kernel) %.3e" % (d, G, best, ts[-1]["phi_ms"], unknowns / best * 1e3), flush=True)
