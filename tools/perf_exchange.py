"""Scratch: two mesh domains in one process (in-process transport) on one GPU, a few lagged-exchange sweeps: device times of the
exchange pieces and a target for ncu captures of pack_tally_kernel / unpack_kernel.  python tools/perf_exchange.py [d] [G]"""
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from umt_b200 import mesh as M, problem as PR, teton  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 12
G = int(sys.argv[2]) if len(sys.argv) > 2 else 128
N = 2
ctxs = []
for r in range(N):
    mesh = M.tiled_mesh((d, d, d), rank=r, size=N)
    ctx = teton.SweepContext.from_mesh(mesh, G)
    ctx.compute_geometry(mesh.px)
    NA = ctx.build_product_quadrature(2, 2, 1)
    for b in mesh.boundaries:
        if b.bc_type == M.BC_SHARED:
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
    ctx.build_schedule()
    tau = PR.tau()
    ctx.upload_state(None, None, np.full((mesh.nzones, G), tau), np.zeros((mesh.ncornr, G)), tau)
    ctx.init_teton(np.full(mesh.nzones, PR.TR0), PR.group_bounds(G), PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(3), 0.0)
    ctx.init_radiation_field()
    ctxs.append(ctx)
teton.connect_local(ctxs)


def group(fn):
    out = [None] * N
    th = [threading.Thread(target=lambda r=r: out.__setitem__(r, fn(ctxs[r]))) for r in range(N)]
    [t.start() for t in th]
    [t.join() for t in th]
    return out


group(lambda c: c.build_exchange())
for _ in range(4):
    group(lambda c: c.sweep(False, 1))
print("2 domains -d %d -G %d in one process:" % (d, G), [c.last_times() for c in ctxs], flush=True)
for c in ctxs:
    c.close()
