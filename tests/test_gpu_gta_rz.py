"""GPU parity of the r-z grey-transport-acceleration path against the oracle's restatement of SweepGreyUCBrz (KernelNew),
the r-z branch of GTASweep, InitSweepGreyUCBrz, GreySweepNEW and GTASolver, on the level-symmetric S2 angle set of
quadrz.F90.  Opacities are synthetic (seeded): the mini-app build never runs GTA."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T
from umt_b200 import mesh as M
from umt_b200 import problem as PR
from umt_b200.teton import SweepContext

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(mesh, G=4, seed=7, scat=20.0, own_geometry=False):
    om = O.OMesh(mesh)
    g = O.geometry(om)
    q = O.gta_quad_rz()
    sched = O.schedule(om, g, q["omega"], q["finish"])
    rng = np.random.default_rng(seed)
    nz, nc = mesh.nzones, mesh.ncornr
    tau = PR.tau(1e-3)
    s = dict(om=om, g=g, q=q, sched=sched, tau=tau, mesh=mesh, Siga=5 * rng.random((nz, G)), Sigs=scat * rng.random((nz, G)),
             Eta=0.5 * rng.random(nc), Phi=rng.random((nc, G)))
    chi = rng.random((nc, G))
    s["Chi"] = chi / chi.sum(1, keepdims=True)
    # a Sn context in r-z whose PhiTotal is Phi: 1 x 1 product set, Psi = Phi / 2 pi on every weighted ordinate
    qs = O.quad_rz(1, 1)
    ctx = SweepContext.from_mesh(mesh, G)
    if own_geometry:
        ctx.compute_geometry(mesh.px)
    else:
        ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], g["Area"], g["RadiusFP"], g["RadiusEZ"], g["A_bdy"])
    ctx.build_product_quadrature(1, 1, 1)
    ctx.upload_state(np.tile(s["Phi"] / (2 * np.pi), (len(qs["weight"]), 1, 1)), None, np.full((nz, G), tau), np.zeros((nc, G)), tau)
    ctx.init_phi_total()
    ctx.gta_setup()
    s["ctx"] = ctx
    return s


MESHES = [("tiled", lambda: M.tiled_mesh((3, 3, 0))), ("box", lambda: M.box_mesh((5, 4)))]


@pytest.mark.parametrize("name,mk", MESHES)
def test_gta_rz_pieces_match_oracle(name, mk):
    s = _setup(mk())
    ctx, om, g, q = s["ctx"], s["om"], s["g"], s["q"]
    nc, nb = s["mesh"].ncornr, s["mesh"].nbelem
    om_d, w_d = ctx.gta_quadrature()
    assert np.array_equal(om_d, q["omega"]) and np.array_equal(w_d, q["weight"])
    assert T.relerr(ctx.download_phi(), s["Phi"]) <= 1e-13
    chi_ref, chi_dev = s["Chi"].copy(), s["Chi"].copy()
    op = O.gta_set_opacity(om, g, s["tau"], s["Siga"], s["Sigs"], s["Eta"], chi_ref)
    op_d = ctx.gta_compute_opacity(s["Siga"], s["Sigs"], s["Eta"], chi_dev)
    for k in op:
        assert T.relerr(op_d[k], op[k]) <= TOL, k
    gs = O.collision_rate(om, s["Eta"], s["Siga"], s["Sigs"], s["Phi"], np.zeros(nc), 0)
    gs_d = ctx.collision_rate(s["Eta"], s["Siga"], s["Sigs"], 0)
    assert T.relerr(gs_d, gs) <= TOL
    P = O.GtaProblem(om, g, s["sched"], q["omega"], q["weight"], op, gs, PR.wtiso(2), q=q)
    TT = P.init_tt().copy()
    TT_d = ctx.gta_init_tt()
    assert np.abs(TT_d - TT).max() <= TOL * np.abs(TT).max()
    # one GTASweep over the 6 swept directions with non-zero incident PsiB: PhiInc, exiting boundary fluxes; the rows of the
    # finishing directions pass through untouched
    rng = np.random.default_rng(3)
    Pvec = rng.random(nc)
    PsiB0 = rng.random((8, nb))
    tsa = PR.wtiso(2) * (op["GreySigScat"] * Pvec + gs)
    PhiInc, PsiB_ref = np.zeros(nc), PsiB0.copy()
    tPsiM, tInc = np.zeros(nc), np.zeros(nc)
    for a in range(8):
        if q["finish"][a]:
            continue
        P.sweep_angle_rz(a, tsa, PsiB_ref[a], PhiInc, tPsiM, tInc)
    PhiInc_d, PsiB_d = ctx.gta_sweep(Pvec, gs, PsiB0.copy(), True)
    assert T.mixed_err(PhiInc_d, PhiInc, TOL) <= 1.0
    assert T.mixed_err(PsiB_d, PsiB_ref, TOL) <= 1.0
    assert np.array_equal(PsiB_d[[3, 7]], PsiB0[[3, 7]])
    # GreySweepNEW with and without source
    Pr, Br = np.zeros(nc), np.zeros((8, nb))
    Pd, Bd = np.zeros(nc), np.zeros((8, nb))
    P.grey_sweep(Br, Pr, True)
    ctx.gta_grey_sweep(Pd, Bd, True)
    assert T.mixed_err(Pd, Pr, 1e-11) <= 1.0 and T.mixed_err(Bd, Br, 1e-11) <= 1.0
    P.GreySource[:] = 0
    ctx.gta_set_source(np.zeros(nc))
    P.grey_sweep(Br, Pr, False)
    ctx.gta_grey_sweep(Pd, Bd, False)
    assert T.mixed_err(Pd, Pr, 1e-11) <= 1.0 and T.mixed_err(Bd, Br, 1e-11) <= 1.0
    ctx.close()


@pytest.mark.parametrize("own_geometry", [False, True])
@pytest.mark.parametrize("name,mk", MESHES)
def test_gta_rz_solver_matches_oracle(name, mk, own_geometry):
    s = _setup(mk(), own_geometry=own_geometry)
    ctx, om, g, q = s["ctx"], s["om"], s["g"], s["q"]
    nc = s["mesh"].ncornr
    chi_ref, chi_dev = s["Chi"].copy(), s["Chi"].copy()
    op = O.gta_set_opacity(om, g, s["tau"], s["Siga"], s["Sigs"], s["Eta"], chi_ref)
    ctx.gta_compute_opacity(s["Siga"], s["Sigs"], s["Eta"], chi_dev)
    gs = O.collision_rate(om, s["Eta"], s["Siga"], s["Sigs"], s["Phi"], np.zeros(nc), 0)
    ctx.collision_rate(s["Eta"], s["Siga"], s["Sigs"], 0)
    P = O.GtaProblem(om, g, s["sched"], q["omega"], q["weight"], op, gs, PR.wtiso(2), q=q)
    corr, n, err = P.solve(s["Phi"])
    corr_d, n_d, err_d = ctx.gta_solve()
    assert n_d == n and n > 3
    assert np.abs(corr_d - corr).max() <= 1e-8 * np.abs(corr).max()
    assert abs(err_d - err) <= 1e-5 * max(err, 1e-30) + 1e-12
    phi_ref = O.add_grey_corrections(corr, chi_ref, s["Phi"].copy())
    ctx.add_grey_corrections()
    assert np.abs(ctx.download_phi() - phi_ref).max() <= 1e-8 * np.abs(phi_ref).max()
    ctx.close()


# ---------------------------------------------------------------------------
# reflecting boundaries in the r-z grey sweeps
# ---------------------------------------------------------------------------
def _reflect_info(mesh, g, q):
    class _P:
        pass
    p = _P()
    p.mesh, p.geom, p.omega, p.NA, p.q = mesh, g, q["omega"], 8, q
    mrefs = T.oracle_reflected_angles(p)
    stage = T.rz_reflect_stages(p, mrefs)
    ops = [(stage[a], a, int(mr[a]), b.first_elem - 1, b.n_elem) for mr, b in zip(mrefs, T.reflecting_boundaries(mesh))
           for a in range(8) if mr[a] >= 0 and not q["finish"][a]]
    return np.array(stage, np.int32), np.array(ops, np.int32).reshape(-1, 5)


@pytest.mark.parametrize("sides", [(2,), (3,), (2, 3), (1,), (1, 3)])
def test_gta_rz_reflecting_matches_oracle(sides):
    mesh = M.tiled_mesh((2, 2, 0), reflecting=sides)
    s = _setup(mesh)
    ctx, om, g, q = s["ctx"], s["om"], s["g"], s["q"]
    for b in T.reflecting_boundaries(mesh):
        ctx.add_reflecting_boundary(b.first_elem, b.n_elem)
    ctx.gta_setup()
    nc, nb = mesh.ncornr, mesh.nbelem
    refl = _reflect_info(mesh, g, q)
    chi_ref, chi_dev = s["Chi"].copy(), s["Chi"].copy()
    op = O.gta_set_opacity(om, g, s["tau"], s["Siga"], s["Sigs"], s["Eta"], chi_ref)
    ctx.gta_compute_opacity(s["Siga"], s["Sigs"], s["Eta"], chi_dev)
    gs = O.collision_rate(om, s["Eta"], s["Siga"], s["Sigs"], s["Phi"], np.zeros(nc), 0)
    ctx.collision_rate(s["Eta"], s["Siga"], s["Sigs"], 0)
    P = O.GtaProblem(om, g, s["sched"], q["omega"], q["weight"], op, gs, PR.wtiso(2), q=q, reflect=refl)
    P.init_tt()
    ctx.gta_init_tt()
    rng = np.random.default_rng(11)
    Pr, Br = rng.random(nc), rng.random((8, nb))
    Pd, Bd = Pr.copy(), Br.copy()
    P.grey_sweep(Br, Pr, True)
    ctx.gta_grey_sweep(Pd, Bd, True)
    assert T.mixed_err(Pd, Pr, 1e-11) <= 1.0 and T.mixed_err(Bd, Br, 1e-11) <= 1.0
    P2 = O.GtaProblem(om, g, s["sched"], q["omega"], q["weight"], op, gs, PR.wtiso(2), q=q, reflect=refl)
    corr, n, err = P2.solve(s["Phi"])
    corr_d, n_d, err_d = ctx.gta_solve()
    assert n_d == n and n > 3
    assert np.abs(corr_d - corr).max() <= 1e-8 * np.abs(corr).max()
    ctx.close()
