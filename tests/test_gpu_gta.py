"""GPU parity of the grey-transport-acceleration path (3-D, "new" GTA solver) against the oracle's restatement of
setGTAOpacityNEW, getCollisionRate, InitGreySweepUCBxyz, GTASweep + SweepGreyUCBxyz KernelNew, GreySweepNEW
(ScalarIntensityDecompose/Solve), GTASolver (BiCGSTAB) and addGreyCorrections.  The mini-app build never runs GTA
(LinearSolver.F90:87-121), so opacities are synthetic (seeded)."""
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
    omega, w = O.gta_quad_xyz()
    sched = O.schedule(om, g, omega)
    rng = np.random.default_rng(seed)
    nz, nc = mesh.nzones, mesh.ncornr
    tau = PR.tau(1e-3)
    Siga = 5 * rng.random((nz, G))
    Sigs = scat * rng.random((nz, G))
    Eta = 0.5 * rng.random(nc)
    Chi = rng.random((nc, G))
    Chi /= Chi.sum(1, keepdims=True)
    Phi = rng.random((nc, G))
    ctx = SweepContext.from_mesh(mesh, G)
    if own_geometry:   # umt_compute_geometry (built without FMA contraction) must put every omega.A = 0 tie where the oracle's does
        ctx.compute_geometry(mesh.px)
    else:
        ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], A_bdy=g["A_bdy"])
    ctx.build_product_quadrature(1, 1, 1)
    ctx.upload_state(np.tile(Phi / (4 * np.pi), (8, 1, 1)), None, np.full((nz, G), tau), np.zeros((nc, G)), tau)
    ctx.init_phi_total()          # PhiTotal on the device = sum_a w_a Psi = Phi
    ctx.gta_setup()
    return dict(om=om, g=g, omega=omega, w=w, sched=sched, tau=tau, Siga=Siga, Sigs=Sigs, Eta=Eta, Chi=Chi, Phi=Phi, ctx=ctx, mesh=mesh)


MESHES = [("tiled", lambda: M.tiled_mesh((2, 2, 2))), ("unstruct", lambda: M.unstruct_box_mesh(2)), ("box", lambda: M.box_mesh((3, 4, 5)))]


@pytest.mark.parametrize("name,mk", MESHES)
def test_gta_pieces_match_oracle(name, mk):
    s = _setup(mk())
    ctx, om, g = s["ctx"], s["om"], s["g"]
    nc, nb = s["mesh"].ncornr, s["mesh"].nbelem
    om_d, w_d = ctx.gta_quadrature()
    assert np.array_equal(om_d, s["omega"]) and np.array_equal(w_d, s["w"])
    assert T.relerr(ctx.download_phi(), s["Phi"]) <= 1e-13
    # opacities (setGTAOpacityNEW) and the rescaled Chi
    chi_ref, chi_dev = s["Chi"].copy(), s["Chi"].copy()
    op = O.gta_set_opacity(om, g, s["tau"], s["Siga"], s["Sigs"], s["Eta"], chi_ref)
    op_d = ctx.gta_compute_opacity(s["Siga"], s["Sigs"], s["Eta"], chi_dev)
    for k in op:
        assert T.relerr(op_d[k], op[k]) <= TOL, k
    assert T.relerr(chi_dev, chi_ref) <= TOL
    # collision rate (getCollisionRate, both flags)
    gs = O.collision_rate(om, s["Eta"], s["Siga"], s["Sigs"], s["Phi"], np.zeros(nc), 0)
    gs_d = ctx.collision_rate(s["Eta"], s["Siga"], s["Sigs"], 0)
    assert T.relerr(gs_d, gs) <= TOL
    # transfer matrices (InitGreySweepUCBxyz)
    P = O.GtaProblem(om, g, s["sched"], s["omega"], s["w"], op, gs, PR.wtiso(3))
    TT = P.init_tt().copy()
    TT_d = ctx.gta_init_tt()
    assert np.abs(TT_d - TT).max() <= TOL * np.abs(TT).max()
    # one GTASweep (all 8 angles): PhiInc and the exiting boundary fluxes, with non-zero incident PsiB
    rng = np.random.default_rng(3)
    Pvec = rng.random(nc)
    PsiB0 = rng.random((8, nb))
    tsa = PR.wtiso(3) * (op["GreySigScat"] * Pvec + gs)
    PhiInc = np.zeros(nc)
    PsiB_ref = PsiB0.copy()
    for a in range(8):
        P.sweep_angle(a, tsa, PsiB_ref[a], PhiInc)
    PhiInc_d, PsiB_d = ctx.gta_sweep(Pvec, gs, PsiB0.copy(), True)
    assert T.mixed_err(PhiInc_d, PhiInc, TOL) <= 1.0
    assert T.mixed_err(PsiB_d, PsiB_ref, TOL) <= 1.0
    # GreySweepNEW with and without source (LU decomposition of I - TT sigma_s, then solves)
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
@pytest.mark.parametrize("name,mk", MESHES[:2])
def test_gta_solver_matches_oracle(name, mk, own_geometry):
    s = _setup(mk(), own_geometry=own_geometry)
    ctx, om, g = s["ctx"], s["om"], s["g"]
    nc = s["mesh"].ncornr
    chi_ref, chi_dev = s["Chi"].copy(), s["Chi"].copy()
    op = O.gta_set_opacity(om, g, s["tau"], s["Siga"], s["Sigs"], s["Eta"], chi_ref)
    ctx.gta_compute_opacity(s["Siga"], s["Sigs"], s["Eta"], chi_dev)
    gs = O.collision_rate(om, s["Eta"], s["Siga"], s["Sigs"], s["Phi"], np.zeros(nc), 0)
    ctx.collision_rate(s["Eta"], s["Siga"], s["Sigs"], 0)
    P = O.GtaProblem(om, g, s["sched"], s["omega"], s["w"], op, gs, PR.wtiso(3))
    corr, n, err = P.solve(s["Phi"])
    corr_d, n_d, err_d = ctx.gta_solve()
    assert n_d == n and n > 3
    # Krylov recurrences amplify rounding differences (FMA contraction, reduction order): 1e-8 of the correction's scale
    assert np.abs(corr_d - corr).max() <= 1e-8 * np.abs(corr).max()
    assert abs(err_d - err) <= 1e-6 * max(err, 1e-30) + 1e-12
    # addGreyCorrections: PhiTotal += correction * Chi
    phi_ref = O.add_grey_corrections(corr, chi_ref, s["Phi"].copy())
    ctx.add_grey_corrections()
    assert np.abs(ctx.download_phi() - phi_ref).max() <= 1e-8 * np.abs(phi_ref).max()
    ctx.close()


def test_gta_no_scattering_exits_immediately():
    """GreySigScat = 0 everywhere: scat_prod1 of the residual is 0, the solver returns the first sweep (GTASolver.F90:232-237)."""
    s = _setup(M.box_mesh((3, 3, 3)), scat=0.0)
    ctx, om, g = s["ctx"], s["om"], s["g"]
    nc = s["mesh"].ncornr
    s["Siga"][:] = s["Siga"][:, :1]      # grey absorber, no re-emission: greysigt == greysiga
    s["Eta"][:] = 0.0
    chi_ref, chi_dev = s["Chi"].copy(), s["Chi"].copy()
    op = O.gta_set_opacity(om, g, s["tau"], s["Siga"], s["Sigs"], s["Eta"], chi_ref)
    ctx.gta_compute_opacity(s["Siga"], s["Sigs"], s["Eta"], chi_dev)
    assert (op["GreySigScat"] == 0).all()
    gs = O.collision_rate(om, s["Eta"], s["Siga"], s["Sigs"], s["Phi"], np.zeros(nc), 0)
    ctx.collision_rate(s["Eta"], s["Siga"], s["Sigs"], 0)
    P = O.GtaProblem(om, g, s["sched"], s["omega"], s["w"], op, gs, PR.wtiso(3))
    corr, n, _ = P.solve(s["Phi"])
    corr_d, n_d, _ = ctx.gta_solve()
    assert n == 1 and n_d == 1
    assert T.mixed_err(corr_d, corr, 1e-11) <= 1.0
    ctx.close()


@pytest.mark.parametrize("G", [4, 33, 128])
def test_source_build_extension(G):
    """umt_build_source (extension, parity unpinned — the mini-app reference never fills STotal): numpy restatement of its
    formula; and consistency with getCollisionRate: sum_g STotal / wtiso - sum_g emission = collision rate when sum_g Chi = 1."""
    s = _setup(M.box_mesh((3, 3, 2)), G=G)
    ctx, mesh = s["ctx"], s["mesh"]
    nc = mesh.ncornr
    rng = np.random.default_rng(11)
    emis = rng.random((nc, G))
    c2z = np.repeat(np.arange(mesh.nzones), mesh.numCorner)
    wt = PR.wtiso(3)
    ref = wt * (s["Sigs"][c2z] * s["Phi"] + s["Chi"] * (s["Eta"] * (s["Siga"][c2z] * s["Phi"]).sum(1))[:, None] + emis)
    got = ctx.build_source(s["Siga"], s["Sigs"], s["Eta"], s["Chi"], emis)
    assert T.relerr(got, ref) <= 1e-12
    coll = O.collision_rate(s["om"], s["Eta"], s["Siga"], s["Sigs"], s["Phi"], np.zeros(nc), 0)
    assert np.abs(got.sum(1) / wt - emis.sum(1) - coll).max() <= 1e-11 * np.abs(coll).max()
    # the sweep consumes it: one sweep with the built source runs and stays finite
    ctx.build_schedule()
    ctx.sweep(False)
    assert np.isfinite(ctx.download_phi()).all()
    ctx.close()


# ---------------------------------------------------------------------------
# reflecting boundaries in the grey sweeps (GTASweep.F90:151 snreflect on the GTA angle set)
# ---------------------------------------------------------------------------
def _gta_reflect(mesh, g, omega):
    """mirror angles of the S2 set on every reflecting boundary, sweep stages, and the copy ops (stage, minc, mref, first, n)"""
    class _P:
        pass
    p = _P()
    p.mesh, p.geom, p.omega, p.NA = mesh, g, omega, len(omega)
    mrefs = T.oracle_reflected_angles(p)
    stage = T.reflect_stages(mrefs, p.NA)
    ops = [(stage[a], a, int(mr[a]), b.first_elem - 1, b.n_elem) for mr, b in zip(mrefs, T.reflecting_boundaries(mesh)) for a in range(p.NA) if mr[a] >= 0]
    return np.array(stage, np.int32), np.array(ops, np.int32).reshape(-1, 5)


def _setup_reflecting(mesh, G=4, seed=7):
    s = _setup(mesh, G, seed)
    # _setup called gta_setup before the reflecting boundaries were known: add them and set up again
    for b in T.reflecting_boundaries(mesh):
        s["ctx"].add_reflecting_boundary(b.first_elem, b.n_elem)
    s["ctx"].gta_setup()
    return s


@pytest.mark.parametrize("sides", [(1,), (0, 2), (0, 2, 4), (0, 1)])
def test_gta_reflecting_matches_oracle(sides):
    s = _setup_reflecting(M.tiled_mesh((2, 2, 2), reflecting=sides))
    ctx, om, g = s["ctx"], s["om"], s["g"]
    nc, nb = s["mesh"].ncornr, s["mesh"].nbelem
    refl = _gta_reflect(s["mesh"], g, s["omega"])
    chi_ref, chi_dev = s["Chi"].copy(), s["Chi"].copy()
    op = O.gta_set_opacity(om, g, s["tau"], s["Siga"], s["Sigs"], s["Eta"], chi_ref)
    ctx.gta_compute_opacity(s["Siga"], s["Sigs"], s["Eta"], chi_dev)
    gs = O.collision_rate(om, s["Eta"], s["Siga"], s["Sigs"], s["Phi"], np.zeros(nc), 0)
    ctx.collision_rate(s["Eta"], s["Siga"], s["Sigs"], 0)
    P = O.GtaProblem(om, g, s["sched"], s["omega"], s["w"], op, gs, PR.wtiso(3), reflect=refl)
    P.init_tt()
    ctx.gta_init_tt()
    rng = np.random.default_rng(11)
    Pr, Br = rng.random(nc), rng.random((8, nb))
    Pd, Bd = Pr.copy(), Br.copy()
    P.grey_sweep(Br, Pr, True)
    ctx.gta_grey_sweep(Pd, Bd, True)
    assert T.mixed_err(Pd, Pr, 1e-11) <= 1.0 and T.mixed_err(Bd, Br, 1e-11) <= 1.0
    # the solver on top (fresh transfer matrices)
    P2 = O.GtaProblem(om, g, s["sched"], s["omega"], s["w"], op, gs, PR.wtiso(3), reflect=refl)
    corr, n, err = P2.solve(s["Phi"])
    corr_d, n_d, err_d = ctx.gta_solve()
    assert n_d == n and n > 3
    assert np.abs(corr_d - corr).max() <= 1e-8 * np.abs(corr).max()
    ctx.close()


def test_gta_mirror_plane_reproduces_symmetric_full_domain():
    """uniform data on the box [0,1]^3 is symmetric about x = 1/2: the grey correction of the left half closed by a reflecting
    x+ side equals the full-domain correction (solved tightly on both)."""
    G = 2

    def uniform(mesh):
        s = _setup_reflecting(mesh, G) if T.reflecting_boundaries(mesh) else _setup(mesh, G)
        s["Siga"][:] = 1.0; s["Sigs"][:] = 30.0; s["Eta"][:] = 0.2; s["Chi"][:] = 0.5
        return s
    full = uniform(M.box_mesh((8, 4, 4)))
    op = O.gta_set_opacity(full["om"], full["g"], full["tau"], full["Siga"], full["Sigs"], full["Eta"], full["Chi"].copy())
    phi1 = np.ones_like(full["Phi"])
    gs = O.collision_rate(full["om"], full["Eta"], full["Siga"], full["Sigs"], phi1, np.zeros(full["mesh"].ncornr), 0)
    corr_full, n, _ = O.GtaProblem(full["om"], full["g"], full["sched"], full["omega"], full["w"], op, gs, PR.wtiso(3)).solve(phi1, epsPoint=1e-11, maxIters=200)
    full["ctx"].close()
    half = uniform(M.box_mesh((4, 4, 4), lengths=(0.5, 1.0, 1.0), reflecting=(1,)))
    ctx, hm = half["ctx"], half["mesh"]
    ctx.upload_state(np.tile(np.ones((hm.ncornr, G)) / (4 * np.pi), (8, 1, 1)), None, np.full((hm.nzones, G), half["tau"]), np.zeros((hm.ncornr, G)), half["tau"])
    ctx.init_phi_total()
    ctx.gta_compute_opacity(half["Siga"], half["Sigs"], half["Eta"], half["Chi"].copy())
    ctx.collision_rate(half["Eta"], half["Siga"], half["Sigs"], 0)
    corr_half, n_h, _ = ctx.gta_solve(epsPoint=1e-11, maxIters=200)
    fm = full["mesh"]
    zc_f = np.repeat(np.add.reduceat(fm.px, fm.cOffSet) / 8.0, fm.numCorner, axis=0)
    look = {tuple(np.round(np.r_[fm.px[c], zc_f[c]] * 1e8).astype(np.int64)): c for c in range(fm.ncornr)}
    zc_h = np.repeat(np.add.reduceat(hm.px, hm.cOffSet) / 8.0, hm.numCorner, axis=0)
    idx = np.array([look[tuple(np.round(np.r_[hm.px[c], zc_h[c]] * 1e8).astype(np.int64))] for c in range(hm.ncornr)])
    assert np.abs(corr_half - corr_full[idx]).max() <= 1e-7 * np.abs(corr_full).max()
    ctx.close()
