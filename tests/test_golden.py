"""The committed fixtures of tests/golden/ (made by tools/make_golden.py from the oracle; see its docstring for what they
do and do not pin).  CPU: the oracle built here reproduces them.  GPU: the CUDA path reproduces them without the oracle's
outputs in the loop."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T
from umt_b200 import mesh as M
from umt_b200 import problem as PR

G_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G_DIR, name))


def test_oracle_reproduces_golden_xyz():
    g = _load("sweep_xyz_tiled112_P1A2_G4_seed1234.npz")
    p = T.make_problem_3d(M.tiled_mesh((1, 1, 2)), 1, 2, 4, seed=1234)
    assert np.array_equal(p.sched["nHyperPlanes"], g["nHyperPlanes"]) and np.array_equal(p.sched["nextZ"][0], g["nextZ0"])
    assert T.relerr(T.oracle_sweep_3d(p, False), g["phi_nonfinal"]) <= 1e-13
    assert T.relerr(T.oracle_sweep_3d(p, True), g["phi_final"]) <= 1e-13
    assert T.mixed_err(p.PsiB, g["psib"], 1e-12) <= 1.0


def test_oracle_reproduces_golden_rz():
    g = _load("sweep_rz_tiled22_P2A2_G4_seed1234.npz")
    p = T.make_problem_rz(M.tiled_mesh((2, 2, 0)), 2, 2, 4, seed=1234)
    assert np.array_equal(p.q["weight"], g["weight"]) and np.array_equal(p.q["angDerivFac"], g["angDerivFac"])
    assert T.relerr(T.oracle_sweep_rz(p, False), g["phi_nonfinal"]) <= 1e-13
    assert T.relerr(T.oracle_sweep_rz(p, True), g["phi_final"]) <= 1e-13


def test_oracle_reproduces_golden_quadrature():
    g = _load("quadrature_product.npz")
    om, w = O.quad_xyz(4, 4)
    q = O.quad_rz(2, 2)
    assert np.array_equal(om, g["omega_xyz_P4A4"]) and np.array_equal(w, g["weight_xyz_P4A4"])
    assert np.array_equal(q["omega"], g["omega_rz_P2A2"]) and np.array_equal(q["quadTauW1"], g["tauW1_rz"])


def test_oracle_reproduces_golden_cycle():
    g = _load("cycle_tiled111_P2A2_G2.npz")
    p = T.make_problem_3d(M.tiled_mesh((1, 1, 1)), 2, 2, 2, driver_like=True)
    p.tau = PR.tau()
    p.Sigt[:] = p.tau
    p.Psi[:] = PR.wtiso(3) * O.planck_groups_ref(PR.TR0, PR.group_bounds(2), 1.0, PR.SPEED_LIGHT * PR.RAD_CONSTANT)
    for row in g["edits"]:
        r = T.oracle_cycle_3d(p, PR.DT, PR.TFLOOR ** 4)
        assert np.allclose([r["EnergyRadiation"], r["TrMax"], r["PowerEscape"]], row[:3], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_gpu_reproduces_golden_xyz_and_rz():
    g = _load("sweep_xyz_tiled112_P1A2_G4_seed1234.npz")
    p = T.make_problem_3d(M.tiled_mesh((1, 1, 2)), 1, 2, 4, seed=1234)
    ctx = T.gpu_context_3d(p, own_schedule=True, own_geometry=True, own_quadrature=(1, 2, 1))
    ctx.sweep(False)
    assert T.relerr(ctx.download_phi(), g["phi_nonfinal"]) <= 1e-12
    ctx.sweep(True)
    assert T.relerr(ctx.download_phi(), g["phi_final"]) <= 1e-12
    assert T.mixed_err(ctx.download_psib(), g["psib"], 1e-12) <= 1.0
    assert T.mixed_err(ctx.download_psi()[:, 0, :], g["psi_corner0"], 1e-12) <= 1.0
    ctx.close()
    g = _load("sweep_rz_tiled22_P2A2_G4_seed1234.npz")
    p = T.make_problem_rz(M.tiled_mesh((2, 2, 0)), 2, 2, 4, seed=1234)
    ctx = T.gpu_context_rz(p, own_schedule=True, own_geometry=True, own_quadrature=(2, 2, 1))
    ctx.sweep(False)
    assert T.relerr(ctx.download_phi(), g["phi_nonfinal"]) <= 1e-12
    ctx.sweep(True)
    assert T.relerr(ctx.download_phi(), g["phi_final"]) <= 1e-12
    ctx.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_cycle():
    from umt_b200.cycle import MiniAppCycle
    from umt_b200.teton import SweepContext
    g = _load("cycle_tiled111_P2A2_G2.npz")
    mesh = M.tiled_mesh((1, 1, 1))
    ctx = SweepContext.from_mesh(mesh, 2)
    ctx.compute_geometry(mesh.px)
    ctx.build_product_quadrature(2, 2, 1)
    ctx.build_schedule()
    cyc = MiniAppCycle(ctx, mesh, 2)
    for row in g["edits"]:
        ed = cyc.step()
        assert np.allclose([ed["EnergyRadiation"], ed["TrMax"], ed["PowerEscape"]], row[:3], rtol=1e-10, atol=0)
    assert T.relerr(ctx.download_phi(), g["phi"]) <= 1e-12
    ctx.close()


def test_oracle_reproduces_golden_gta_and_scheduler():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(G_DIR), "..", "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = _load("gta_solve_G4_seed7.npz")
    c3, n3, e3 = mg.gta_solve(M.tiled_mesh((1, 1, 2)))
    c2, n2, e2 = mg.gta_solve(M.tiled_mesh((2, 2, 0)))
    assert n3 == int(g["iters_xyz"]) and n2 == int(g["iters_rz"])
    assert np.abs(c3 - g["corr_xyz_tiled112"]).max() <= 1e-9 * np.abs(c3).max()      # Krylov: compiler-dependent rounding is amplified
    assert np.abs(c2 - g["corr_rz_tiled22"]).max() <= 1e-9 * np.abs(c2).max()
    q = O.gta_quad_rz()
    assert np.array_equal(q["angDerivFac"], g["rz_angDerivFac"]) and np.array_equal(q["quadTauW1"], g["rz_tauW1"])
    s = _load("scheduler_2domains_4sets_seed9.npz")
    problems = [T.make_problem_3d(M.tiled_mesh((2, 2, 1), rank=r, size=2), 1, 2, 2, seed=200 + r) for r in range(2)]
    rng = np.random.default_rng(9)
    nf = [rng.standard_normal((len(T.shared_boundaries(p.mesh)), problems[0].NA)) for p in problems]
    order, recv = T.oracle_sweep_scheduler(problems, 4, nf)
    assert np.array_equal(np.array(order), s["order"]) and np.array_equal(np.array([r[0] for r in recv]), s["recv"])


@pytest.mark.gpu
def test_gpu_reproduces_golden_gta():
    import importlib.util
    from umt_b200.teton import SweepContext
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(G_DIR), "..", "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = _load("gta_solve_G4_seed7.npz")
    for mesh, key, itk in ((M.tiled_mesh((1, 1, 2)), "corr_xyz_tiled112", "iters_xyz"), (M.tiled_mesh((2, 2, 0)), "corr_rz_tiled22", "iters_rz")):
        G = 4
        Siga, Sigs, Eta, Chi, Phi = mg.gta_inputs(mesh, G, 7)
        ctx = SweepContext.from_mesh(mesh, G)
        # the oracle's geometry, not umt_compute_geometry: the tiled mesh has faces whose normal is perpendicular to S2 ordinates
        # (omega . A = 0 up to rounding), and which closure branch such a face takes (SweepGreyUCBxyz.F90:263-290: opposite face
        # incident or not) then hangs on the last bit of A -- two valid discretisations that differ by ~1 % in the correction
        gg = O.geometry(O.OMesh(mesh))
        if mesh.ndim == 3:
            ctx.set_geometry(gg["Volume"], gg["A_fp"], gg["A_ez"], A_bdy=gg["A_bdy"])
        else:
            ctx.set_geometry(gg["Volume"], gg["A_fp"], gg["A_ez"], gg["Area"], gg["RadiusFP"], gg["RadiusEZ"], gg["A_bdy"])
        ctx.build_product_quadrature(1, 1, 1)
        NA = len(O.quad_rz(1, 1)["weight"]) if mesh.ndim == 2 else 8
        tau = PR.tau(1e-3)
        ctx.upload_state(np.tile(Phi / (4 * np.pi if mesh.ndim == 3 else 2 * np.pi), (NA, 1, 1)), None, np.full((mesh.nzones, G), tau), np.zeros((mesh.ncornr, G)), tau)
        ctx.init_phi_total()
        ctx.gta_setup()
        ctx.gta_compute_opacity(Siga, Sigs, Eta, Chi.copy())
        ctx.collision_rate(Eta, Siga, Sigs, 0)
        corr, n, err = ctx.gta_solve()
        assert n == int(g[itk])
        assert np.abs(corr - g[key]).max() <= 1e-8 * np.abs(g[key]).max()
        ctx.close()
