"""Device times of the 3-D sweep against the size of the Psi1 ring (single-psi layout): python tools/perf_ring.py d G P A ring[,ring...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from umt_b200 import mesh as M, problem as PR, teton  # noqa: E402

d, G, P, A = (int(x) for x in sys.argv[1:5])
rings = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
mesh = M.tiled_mesh((d, d, d))
ctx = teton.SweepContext.from_mesh(mesh, G)
ctx.compute_geometry(mesh.px)
NA = ctx.build_product_quadrature(P, A, 1)
t0 = time.perf_counter()
ctx.build_schedule()
t_sched = time.perf_counter() - t0
nz, nc = mesh.nzones, mesh.ncornr
tau = PR.tau()
ctx.upload_state(None, None, np.full((nz, G), tau), np.zeros((nc, G)), tau)
ctx.init_teton(np.full(nz, PR.TR0), PR.group_bounds(G), PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(3), 0.0)
unknowns = nc * NA * G
print(f"-d {d} -G {G} P{P} A{A}: {NA} angles, {unknowns:.3e} unknowns, host schedule build {t_sched:.2f} s", flush=True)
ref = None
for ring in rings:
    ctx.set_psi1_ring(ring)
    t0 = time.perf_counter()
    ctx.init_radiation_field()          # finalises the schedule: plan records, work items, Psi1 workspace
    t_fin = time.perf_counter() - t0
    lay = ctx.psi_layout()
    free, total = torch.cuda.mem_get_info()
    for _ in range(2):
        ctx.sweep(False)
    tm = []
    for _ in range(4):
        ctx.sweep(False)
        tm.append(ctx.last_times())
    phi = ctx.download_phi()
    if ref is None:
        ref = phi
    same = bool(np.array_equal(phi, ref))
    sw = np.mean([t["sweep_ms"] for t in tm]); ph = np.mean([t["phi_ms"] for t in tm]); tot = np.mean([t["total_ms"] for t in tm])
    print(f"ring {ring}: slabs {lay['psi1_slabs']} of {NA}, tallied in kernel {lay['angles_tallied_in_sweep']}, psi+psi1 {lay['bytes'] / 1e9:.1f} GB, device used {(total - free) / 1e9:.1f} GB, "
          f"finalize {t_fin * 1e3:.0f} ms | sweep {sw:.2f} ms, phi tail {ph:.2f} ms, total {tot:.2f} ms = {unknowns / tot / 1e6:.1f}e9 unknowns/s, phi identical to first: {same}", flush=True)
ctx.sweep(True)
t = ctx.last_times()
print(f"savePsi sweep (in place): sweep {t['sweep_ms']:.2f} ms, phi {t['phi_ms']:.2f} ms, total {t['total_ms']:.2f} ms", flush=True)
ctx.close()
