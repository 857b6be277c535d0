"""Scratch: device times of the RZ sweep (BASELINE configs[1] per-domain size) and of the GTA pieces (3-D), one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from umt_b200 import mesh as M, problem as PR, teton

what = sys.argv[1] if len(sys.argv) > 1 else "rz"
if what == "rz":
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    G = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    mesh = M.tiled_mesh((d, d, 0))
    ctx = teton.SweepContext.from_mesh(mesh, G)
    ctx.compute_geometry(mesh.px)
    NA = ctx.build_product_quadrature(2, 2, 1)
    ctx.build_schedule()
    tau = PR.tau()
    ctx.upload_state(None, None, np.full((mesh.nzones, G), tau), np.zeros((mesh.ncornr, G)), tau)
    ctx.init_teton(np.full(mesh.nzones, PR.TR0), PR.group_bounds(G), PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(2), 0.0)
    ctx.init_radiation_field()
    unknowns = mesh.ncornr * NA * G           # the driver's count (test_driver.cc:1887), incl. zero-weight directions
    swept = mesh.ncornr * (NA - NA // 6) * G if NA % 6 == 0 else unknowns
    nh = [ctx.schedule_info(a + 1)[0] for a in range(NA)]
    ts = []
    for i in range(5):
        ctx.sweep(False)
        ts.append(ctx.last_times())
    best = min(t["sweep_ms"] for t in ts[1:])
    print({k: v for k, v in os.environ.items() if k.startswith("UMT_")},
          "RZ d=%d G=%d zones=%d angles=%d planes/angle~%d sweep_ms %.3f phi_ms %.3f  unknowns/s (driver count) %.3e  B_alg(58+128/G) -> %.0f GB/s"
          % (d, G, mesh.nzones, NA, max(nh), best, ts[-1]["phi_ms"], unknowns / best * 1e3, swept * (58 + 128.0 / G) / best * 1e3 / 1e9), flush=True)
else:
    rz = what == "gtarz"
    d = int(sys.argv[2]) if len(sys.argv) > 2 else (40 if rz else 20)
    G = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    mesh = M.tiled_mesh((d, d, 0) if rz else (d, d, d))
    ctx = teton.SweepContext.from_mesh(mesh, G)
    ctx.compute_geometry(mesh.px)
    ctx.build_product_quadrature(1, 1, 1)
    nz, nc = mesh.nzones, mesh.ncornr
    rng = np.random.default_rng(7)
    tau = PR.tau()
    ctx.upload_state(None, None, np.full((nz, G), tau), np.zeros((nc, G)), tau)
    ctx.init_teton(np.full(nz, PR.TR0), PR.group_bounds(G), PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(2 if rz else 3), 0.0)
    ctx.init_phi_total()
    t0 = time.time(); ctx.gta_setup(); t_setup = time.time() - t0
    Siga, Sigs, Eta = 5 * rng.random((nz, G)), 20 * rng.random((nz, G)), 0.5 * rng.random(nc)
    Chi = rng.random((nc, G)); Chi /= Chi.sum(1, keepdims=True)
    ctx.gta_compute_opacity(Siga, Sigs, Eta, Chi)
    ctx.collision_rate(Eta, Siga, Sigs, 0)
    for rep in range(3):   # the first solve pays the lazy loading of every kernel it launches: report the last
        ctx.collision_rate(Eta, Siga, Sigs, 0)
        t0 = time.time(); corr, n, err = ctx.gta_solve(); t_solve = time.time() - t0
        print("  solve %d: %.1f ms" % (rep, t_solve * 1e3), flush=True)
    nsweeps = n  # one grey sweep per unit of nGreyIter (1 + 2 per BiCGSTAB iteration)
    print(("GTA r-z" if rz else "GTA") + " d=%d zones=%d corners=%d: setup %.2f s; solve %.1f ms, nGreyIter %d (= grey sweeps), %.3f ms per grey sweep, %.3e corner-angle solves/s, err %.2e"
          % (d, nz, nc, t_setup, t_solve * 1e3, n, t_solve * 1e3 / nsweeps, nsweeps * nc * 8 / t_solve, err), flush=True)
