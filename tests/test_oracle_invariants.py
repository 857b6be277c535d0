"""CPU: the oracle (oracle/umt_oracle.c) against the analytic invariants the
reference's algorithm guarantees.  The reference ships no
golden vectors for this path, so these invariants plus the frozen fixtures in
tests/golden/ are what pins the restatement ("parity unpinned" otherwise)."""
import math
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T
from umt_b200 import mesh as M
from umt_b200 import problem as PR


MESHES_3D = [("tiled", lambda: M.tiled_mesh((2, 2, 2))), ("box", lambda: M.box_mesh((3, 4, 5))),
             ("unstruct", lambda: M.unstruct_box_mesh(2)), ("warped", lambda: M.box_mesh((4, 4, 4), warp=0.35, seed=3))]
MESHES_2D = [("tiled2d", lambda: M.tiled_mesh((3, 2, 0))), ("box2d", lambda: M.box_mesh((5, 4)))]


@pytest.mark.parametrize("name,mk", MESHES_3D)
def test_closed_corner_surfaces_3d(name, mk):
    """geometryUCBxyz.F90:110-179: a corner is a closed polyhedron, sum of its FP and EZ area vectors = 0;
    corner volumes add up to the zone volume; the domain volume is the unit box."""
    m = mk()
    g = O.geometry(O.OMesh(m))
    scale = np.abs(g["A_fp"]).max()
    closed = g["A_fp"].sum(1) + g["A_ez"].sum(1)
    assert np.abs(closed).max() <= 1e-13 * scale
    vz = np.add.reduceat(g["Volume"], m.cOffSet)
    assert np.allclose(vz, g["VolumeZone"], rtol=1e-13, atol=0)
    if name != "warped":   # volumeUCBxyz (teton_getvolume) and geometryUCBxyz agree on planar-faced zones
        assert np.allclose(g["Volume_getvolume"], g["Volume"], rtol=1e-12, atol=0)
    if name != "unstruct":   # tiled / box meshes fill the unit cube
        assert abs(g["VolumeZone"].sum() - 1.0) <= 1e-12
    assert (g["Volume"] > 0).all()


@pytest.mark.parametrize("name,mk", MESHES_3D)
def test_area_antisymmetry_3d(name, mk):
    """A_fp is antisymmetric across a zone face (geometryUCBxyz.F90:114-136) and A_ez across the
    corner pair sharing an EZ face (:167-179)."""
    m = mk()
    g = O.geometry(O.OMesh(m))
    nc = m.ncornr
    scale = np.abs(g["A_fp"]).max()
    c = np.repeat(np.arange(nc), m.maxcf).reshape(nc, m.maxcf)
    other = m.cFP - 1
    interior = other < nc
    # find the face slot of the neighbour that points back
    back = np.zeros_like(other)
    for f in range(m.maxcf):
        for f2 in range(m.maxcf):
            hit = interior[:, f] & (m.cFP[np.minimum(other[:, f], nc - 1), f2] - 1 == c[:, f])
            back[hit, f] = f2
    for f in range(m.maxcf):
        i = np.nonzero(interior[:, f])[0]
        s = g["A_fp"][i, f] + g["A_fp"][other[i, f], back[i, f]]
        assert np.abs(s).max(initial=0.0) <= 1e-13 * scale
    # EZ: corner c face f <-> corner (c0 + cEZ-1) whose cEZ points back to c
    c0 = np.repeat(m.cOffSet, m.numCorner)
    local = np.arange(nc) - c0
    for f in range(m.maxcf):
        mate = c0 + m.cEZ[:, f] - 1
        found = np.zeros(nc, bool)
        for f2 in range(m.maxcf):
            hit = (m.cEZ[mate, f2] - 1 == local) & ~found
            s = g["A_ez"][hit, f] + g["A_ez"][mate[hit], f2]
            assert np.abs(s).max(initial=0.0) <= 1e-13 * scale
            found |= hit
        assert found.all()


@pytest.mark.parametrize("name,mk", MESHES_2D)
def test_geometry_rz(name, mk):
    """geometryUCBrz.F90: corner volumes are 2 pi-less r-weighted areas; the r-z unit square
    has volume integral of r dr dz = 1/2; areas add to 1."""
    m = mk()
    g = O.geometry(O.OMesh(m))
    assert abs(g["Area"].sum() - 1.0) <= 1e-12
    assert abs(g["Volume"].sum() - 0.5) <= 1e-12
    assert (g["Volume"] > 0).all() and (g["Area"] > 0).all()


@pytest.mark.parametrize("P,A", [(1, 1), (2, 2), (3, 3), (4, 4), (2, 5)])
def test_quadrature_xyz(P, A):
    """rtquad.F90:95-105: weights sum to 4 pi (sum w * wtiso = 1); odd moments vanish; unit ordinates;
    quadProduct.F90:122-182: 8 octant images of each base ordinate are consecutive."""
    om, w = O.quad_xyz(P, A)
    assert om.shape == (8 * P * A, 3)
    assert abs(w.sum() - 4 * math.pi) <= 1e-12
    assert np.abs((w[:, None] * om).sum(0)).max() <= 1e-12
    assert np.abs((om ** 2).sum(1) - 1).max() <= 1e-13
    # (the tabulated Spence sets are not level-symmetric: second moments are only approximately 4 pi / 3)
    assert np.abs((w[:, None] * om ** 2).sum() - 4 * math.pi) <= 1e-11
    ab = np.abs(om).reshape(P * A, 8, 3)
    assert np.abs(ab - ab[:, :1]).max() <= 1e-15
    signs = np.sign(om.reshape(P * A, 8, 3))
    assert len({tuple(s) for s in signs[0]}) == 8


@pytest.mark.parametrize("P,A", [(1, 1), (2, 2), (3, 4)])
def test_quadrature_rz(P, A):
    """quadrz.F90 product branch + rtquad.F90:107-127: 4P levels of A+2... angles (A weighted per
    quadrant pair + starting + finishing direction), weights sum to 2 pi, start/finish carry no weight."""
    q = O.quad_rz(P, A)
    NA = 4 * P * (A + 1)
    assert q["omega"].shape == (NA, 2)
    assert abs(q["weight"].sum() - 2 * math.pi) <= 1e-12
    assert (q["weight"][q["start"] > 0] == 0).all()
    assert (q["weight"][q["finish"][:NA] > 0] == 0).all()
    assert q["start"].sum() == 2 * P and q["finish"][:NA].sum() == 2 * P
    # every xi-level begins with its starting direction and ends with its finishing direction
    lev = q["level"]
    for l in np.unique(lev):
        idx = np.nonzero(lev == l)[0]
        assert q["start"][idx[0]] == 1 and q["finish"][idx[-1]] == 1
        assert (np.diff(idx) == 1).all()


@pytest.mark.parametrize("name,mk", MESHES_3D)
def test_schedule_validity_3d(name, mk):
    """snnext.F90: every zone exactly once per angle; a zone's upstream neighbours (faces with
    omega.A < 0) come in earlier hyperplanes unless the zone is on the cycle list."""
    m = mk()
    om = O.OMesh(m)
    g = O.geometry(om)
    omega, w = O.quad_xyz(1, 2)
    s = O.schedule(om, g, omega)
    nz = m.nzones
    c2z = np.repeat(np.arange(nz), m.numCorner)
    for a in range(len(w)):
        nh = s["nHyperPlanes"][a]
        zip_ = s["zonesInPlane"][a][:nh]
        assert zip_.sum() == nz and (zip_ > 0).all()
        order = np.abs(s["nextZ"][a]) - 1
        assert np.array_equal(np.sort(order), np.arange(nz))
        plane_of = np.empty(nz, np.int64)
        plane_of[order] = np.repeat(np.arange(nh), zip_)
        # nextC: a permutation of the zone's local corners
        for z in (0, nz // 2, nz - 1):
            nc_z, c0 = m.numCorner[z], m.cOffSet[z]
            assert sorted(s["nextC"][a][c0:c0 + nc_z]) == list(range(1, nc_z + 1))
        cyc = set()
        off, n = s["cycleOffSet"][a], s["numCycles"][a]
        cyc_corners = s["cycleList"][off:off + n] - 1
        afp = np.einsum("cfd,d->cf", g["A_fp"], omega[a])
        up = (afp < 0) & (m.cFP - 1 < m.ncornr)
        ci, fi = np.nonzero(up)
        src_c = m.cFP[ci, fi] - 1
        bad = plane_of[c2z[src_c]] >= plane_of[c2z[ci]]
        # every violated dependency must be lagged through the cycle list (its upstream corner is listed)
        assert set(src_c[bad]).issubset(set(cyc_corners)), (name, a)
    if name != "warped":
        assert s["totalCycles"] == 0 and (s["nextZ"] > 0).all()


def _uniform_3d(mk, G=3):
    m = mk()
    p = T.make_problem_3d(m, 1, 2, G)
    psi0 = np.linspace(0.7, 1.9, G)
    p.Psi[:] = psi0
    p.PsiB[:] = psi0
    p.STotal[:] = (np.repeat(p.Sigt, m.numCorner, axis=0) - p.tau) * psi0
    p.cyclePsi[:] = psi0
    return m, p, psi0


@pytest.mark.parametrize("name,mk", MESHES_3D)
def test_uniform_solution_preserved_3d(name, mk):
    """Infinite-medium: psi = psi0 on all incoming faces, psi^n = psi0, STotal = (Sigt - tau) psi0
    => the sweep returns psi0 in every corner (each corner equation is a balance, sez vanishes)."""
    m, p, psi0 = _uniform_3d(mk)
    phi = T.oracle_sweep_3d(p, True)
    assert np.abs(p.Psi / psi0 - 1).max() <= 1e-11
    assert np.abs(phi / (4 * math.pi * psi0) - 1).max() <= 1e-11
    assert np.abs(p.PsiB / psi0 - 1).max() <= 1e-11


def test_sweep_linearity_3d():
    """The sweep is linear in (STotal, Psi^n, PsiB): sweep(x + 2y) = sweep(x) + 2 sweep(y)."""
    m = M.tiled_mesh((2, 2, 1))
    px = T.make_problem_3d(m, 1, 1, 2, seed=1)
    py = T.make_problem_3d(m, 1, 1, 2, seed=2)
    pz = T.make_problem_3d(m, 1, 1, 2, seed=1)
    py.Sigt[:] = px.Sigt
    pz.Sigt[:] = px.Sigt
    for k in ("STotal", "Psi", "PsiB"):
        getattr(pz, k)[:] = getattr(px, k) + 2.0 * getattr(py, k)
    fx, fy, fz = (T.oracle_sweep_3d(q, True) for q in (px, py, pz))
    assert np.abs(fz - (fx + 2 * fy)).max() <= 1e-12 * np.abs(fz).max()
    assert np.abs(pz.Psi - (px.Psi + 2 * py.Psi)).max() <= 1e-12 * np.abs(pz.Psi).max()


def test_energy_balance_streaming_3d():
    """rtedit.F90:231: with sigma_a = 0 (Sigt = tau, STotal = 0) and vacuum boundaries,
    sum_c V tau (phi - phi^n) + sum_exit w (omega.A) psib = 0 per group (discrete balance of the UCB scheme)."""
    m = M.tiled_mesh((2, 2, 2))
    p = T.make_problem_3d(m, 2, 2, 3, driver_like=True)
    phi_old = np.einsum("a,acg->cg", p.weight, p.Psi)
    phi = T.oracle_sweep_3d(p, True)
    V = p.geom["Volume"][:, None]
    lhs = (V * p.tau * (phi - phi_old)).sum(0)
    leak = np.zeros(p.G)
    for a in range(p.NA):
        for b, c in p.bdy[a]:
            leak += p.weight[a] * float(p.geom["A_bdy"][b - 1] @ p.omega[a]) * p.PsiB[a, b - 1]
    assert np.abs(lhs + leak).max() <= 1e-11 * np.abs(leak).max()


def test_planck_groups_match_reference_source():
    """oracle/_ref/libnbb_ref.so is the reference's own misc/NormalizedBlackBody.cc compiled where it lies;
    Planck group fractions sum to ~1 over [1e-6, 1e2] at T = 0.05 (test_driver.cc:1173-1183)."""
    if not os.path.exists(os.path.join(os.path.dirname(O.__file__), "_ref", "libnbb_ref.so")):
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    b = PR.group_bounds(16)
    B = O.planck_groups_ref(0.05, b)   # k = Bnorm = 1: the groups add up to T^4
    assert (B >= 0).all()
    assert abs(B.sum() / 0.05 ** 4 - 1.0) <= 1e-6


@pytest.mark.parametrize("name,mk", MESHES_2D)
def test_uniform_solution_preserved_rz(name, mk):
    """Same balance in r-z: the angular-derivative terms (PsiM chain, AngleCoef2D.F90) cancel for an
    isotropic, uniform field, including the starting/finishing-direction bookkeeping."""
    m = mk()
    p = T.make_problem_rz(m, 2, 3, 3)
    psi0 = np.linspace(0.7, 1.9, 3)
    p.Psi[:] = psi0
    p.PsiB[:] = psi0
    p.STotal[:] = (np.repeat(p.Sigt, m.numCorner, axis=0) - p.tau) * psi0
    phi = T.oracle_sweep_rz(p, True)
    assert np.abs(p.Psi / psi0 - 1).max() <= 1e-11
    assert np.abs(phi / (2 * math.pi * psi0) - 1).max() <= 1e-11
    assert np.abs(p.PsiB / psi0 - 1).max() <= 1e-11


def test_schedule_validity_rz():
    m = M.tiled_mesh((3, 3, 0))
    p = T.make_problem_rz(m, 2, 2, 1)
    nz = m.nzones
    for a in range(p.NA):
        nh = p.sched["nHyperPlanes"][a]
        if p.q["finish"][a]:
            assert nh == 0   # finishing directions are not swept (rtorder.F90)
            continue
        assert p.sched["zonesInPlane"][a][:nh].sum() == nz
        assert np.array_equal(np.sort(np.abs(p.sched["nextZ"][a])), np.arange(1, nz + 1))


def _gta_problem(mesh, G=4, seed=7):
    om = O.OMesh(mesh)
    g = O.geometry(om)
    omega, w = O.gta_quad_xyz()
    sched = O.schedule(om, g, omega)
    rng = np.random.default_rng(seed)
    nz, nc = mesh.nzones, mesh.ncornr
    tau = PR.tau(1e-3)
    Siga, Sigs, Eta = 5 * rng.random((nz, G)), 20 * rng.random((nz, G)), 0.5 * rng.random(nc)
    Chi = rng.random((nc, G))
    Chi /= Chi.sum(1, keepdims=True)
    Phi = rng.random((nc, G))
    op = O.gta_set_opacity(om, g, tau, Siga, Sigs, Eta, Chi)
    gs = O.collision_rate(om, Eta, Siga, Sigs, Phi, np.zeros(nc), 0)
    return om, g, sched, omega, w, op, gs, Phi, Chi


def test_gta_quadrature_and_opacity():
    """S2 level-symmetric GTA set: 8 ordinates (+-1/sqrt 3), weights pi/2; setGTAOpacityNEW leaves Chi normalised
    with sum_g Chi = 1 / sum(Chi/sigt) scaling and 0 <= GreySigScat < GreySigTotal."""
    omega, w = O.gta_quad_xyz()
    assert np.abs(np.abs(omega) - 0.577350269189625).max() == 0 and np.abs(w - math.pi / 2).max() <= 1e-15
    om, g, sched, omega, w, op, gs, Phi, Chi = _gta_problem(M.tiled_mesh((2, 2, 1)))
    assert (op["GreySigScat"] >= 0).all() and (op["GreySigScat"] < op["GreySigTotal"]).all()
    assert np.abs(op["GreySigtInv"] * op["GreySigTotal"] - 1).max() <= 1e-15
    assert np.abs(Chi.sum(1) - 1).max() <= 1e-14


@pytest.mark.parametrize("mk", [lambda: M.tiled_mesh((2, 2, 2)), lambda: M.unstruct_box_mesh(2)])
def test_gta_solver_solves_the_grey_system(mk):
    """GTASolver's BiCGSTAB answer x satisfies x - M x = b, where b is the first (withSource) grey sweep and M the
    source-free sweep operator (GTASolver.F90:219-402), checked by applying the operator once more."""
    mesh = mk()
    om, g, sched, omega, w, op, gs, Phi, _ = _gta_problem(mesh)
    nc, nb = mesh.ncornr, mesh.nbelem
    P = O.GtaProblem(om, g, sched, omega, w, op, gs, PR.wtiso(3))
    corr, n, err = P.solve(Phi, epsPoint=1e-10, maxIters=200)
    assert 3 < n < 100 and err < 1e-10
    Q = O.GtaProblem(om, g, sched, omega, w, op, gs, PR.wtiso(3))
    Q.init_tt()
    b, bB = np.zeros(nc), np.zeros((8, nb))
    Q.grey_sweep(bB, b, True)
    Q.GreySource[:] = 0
    Mx, MB = corr.copy(), np.zeros((8, nb))
    Q.grey_sweep(MB, Mx, False)
    assert np.abs(b - (corr - Mx)).max() <= 1e-8 * np.abs(b).max()


def test_gta_transfer_matrix_is_the_sweep_response():
    """InitGreySweepUCBxyz: sum_a w_a Pvv is the within-zone response, so for a zone-local unit source and no incident
    flux the angle-integrated corner fluxes of one GTASweep equal TT applied to wtiso * source."""
    mesh = M.box_mesh((1, 1, 1))
    om, g, sched, omega, w, op, gs, Phi, _ = _gta_problem(mesh)
    P = O.GtaProblem(om, g, sched, omega, w, op, gs, PR.wtiso(3))
    TT = P.init_tt().copy()                      # TT[c1, c] = TT(c+1, c1+1)
    nc = mesh.ncornr
    for k in range(nc):
        tsa = np.zeros(nc)
        tsa[k] = 1.0
        phi = np.zeros(nc)
        for a in range(8):
            tPsi, _ = P.sweep_angle(a, tsa, np.zeros(mesh.nbelem), np.zeros(nc))
            phi += w[a] * tPsi[:nc]
        assert np.abs(phi - TT[:, k]).max() <= 1e-13 * np.abs(TT).max()


# ---------------------------------------------------------------------------
# r-z grey transport acceleration (SweepGreyUCBrz KernelNew, InitSweepGreyUCBrz, level-symmetric S2 set)
# ---------------------------------------------------------------------------
def _gta_problem_rz(mesh, G=4, seed=7):
    om = O.OMesh(mesh)
    g = O.geometry(om)
    q = O.gta_quad_rz()
    sched = O.schedule(om, g, q["omega"], q["finish"])
    rng = np.random.default_rng(seed)
    nz, nc = mesh.nzones, mesh.ncornr
    tau = PR.tau(1e-3)
    Siga, Sigs, Eta = 5 * rng.random((nz, G)), 20 * rng.random((nz, G)), 0.5 * rng.random(nc)
    Chi = rng.random((nc, G))
    Chi /= Chi.sum(1, keepdims=True)
    Phi = rng.random((nc, G))
    op = O.gta_set_opacity(om, g, tau, Siga, Sigs, Eta, Chi)
    gs = O.collision_rate(om, Eta, Siga, Sigs, Phi, np.zeros(nc), 0)
    return om, g, q, sched, op, gs, Phi


def test_gta_quadrature_rz():
    """Level-symmetric S2 in r-z: two xi-levels (xi = -+1/sqrt 3) of start, mu = -1/sqrt 3, mu = +1/sqrt 3, finish, weights pi/2;
    weighted-diamond coefficients from AngleCoef2D with the cell edges at phi = pi, pi/2, 0 (mu = -s, 0, s with s = sqrt(2/3))."""
    q = O.gta_quad_rz()
    mu, s = 1 / math.sqrt(3), math.sqrt(2.0 / 3.0)
    assert q["start"].tolist() == [1, 0, 0, 0, 1, 0, 0, 0] and q["finish"][:8].tolist() == [0, 0, 0, 1, 0, 0, 0, 1]
    assert abs(q["weight"].sum() / (2 * math.pi) - 1) <= 1e-15 and q["level"].tolist() == [1, 1, 1, 1, 2, 2, 2, 2]
    assert np.abs(q["omega"][:4, 0] - [-s, -mu, mu, s]).max() <= 1e-15 and np.abs(np.abs(q["omega"][:, 1]) - mu).max() <= 1e-15
    assert np.abs(q["weight"][[1, 2, 5, 6]] - math.pi / 2).max() <= 1e-15
    tau1, tau2 = (s - mu) / s, mu / s
    assert np.abs(q["quadTauW1"][[1, 2]] - [1 / tau1, 1 / tau2]).max() <= 1e-13
    assert np.abs(q["quadTauW2"][[1, 2]] - [(1 - tau1) / tau1, (1 - tau2) / tau2]).max() <= 1e-13
    # alpha_1 = w mu, alpha_2 = 0 (the level closes): angDerivFac = mu_m + alpha_m / (w tau_m)
    assert np.abs(q["angDerivFac"][[1, 2]] - [-mu + mu / tau1, mu]).max() <= 1e-13
    assert np.array_equal(q["angDerivFac"][4:], q["angDerivFac"][:4])


@pytest.mark.parametrize("mk", [lambda: M.tiled_mesh((2, 2, 0)), lambda: M.box_mesh((4, 3))])
def test_gta_rz_uniform_solution_preserved(mk):
    """TsaSource = sigma * c in the volume and c on every incident boundary element: every swept angle returns tPsi = c,
    including the angular-derivative chain through tPsiM (which must therefore carry c from the starting direction on)."""
    mesh = mk()
    om, g, q, sched, op, gs, Phi = _gta_problem_rz(mesh)
    nc, nb = mesh.ncornr, mesh.nbelem
    c = 0.7
    P = O.GtaProblem(om, g, sched, q["omega"], q["weight"], op, gs, PR.wtiso(2), q=q)
    tsa = op["GreySigTotal"] * c
    tPsiM, tInc, PhiInc = np.zeros(nc), np.zeros(nc), np.zeros(nc)
    for a in range(8):
        if q["finish"][a]:
            continue
        if q["start"][a]:
            tPsiM[:] = 0
            tInc[:] = 0
        tPsi, pInc = P.sweep_angle_rz(a, tsa, np.full(nb, c), PhiInc, tPsiM, tInc)
        assert np.abs(tPsi[:nc] - c).max() <= 1e-12


def test_gta_rz_transfer_matrix_is_the_sweep_response():
    """InitGreySweepUCBrz: for a zone-local unit source and no incident flux the angle-integrated corner fluxes of one r-z
    GTASweep equal TT applied to the source (the starting direction feeding tPsiM exactly as Tvv feeds Pvv)."""
    mesh = M.box_mesh((1, 1))
    om, g, q, sched, op, gs, Phi = _gta_problem_rz(mesh)
    P = O.GtaProblem(om, g, sched, q["omega"], q["weight"], op, gs, PR.wtiso(2), q=q)
    TT = P.init_tt().copy()
    nc = mesh.ncornr
    for k in range(nc):
        tsa = np.zeros(nc)
        tsa[k] = 1.0
        phi, tPsiM, tInc = np.zeros(nc), np.zeros(nc), np.zeros(nc)
        for a in range(8):
            if q["finish"][a]:
                continue
            tPsi, _ = P.sweep_angle_rz(a, tsa, np.zeros(mesh.nbelem), np.zeros(nc), tPsiM, tInc)
            phi += q["weight"][a] * tPsi[:nc]
        assert np.abs(phi - TT[:, k]).max() <= 1e-13 * np.abs(TT).max()


def test_gta_rz_solver_solves_the_grey_system():
    mesh = M.tiled_mesh((2, 2, 0))
    om, g, q, sched, op, gs, Phi = _gta_problem_rz(mesh)
    nc, nb = mesh.ncornr, mesh.nbelem
    P = O.GtaProblem(om, g, sched, q["omega"], q["weight"], op, gs, PR.wtiso(2), q=q)
    corr, n, err = P.solve(Phi, epsPoint=1e-10, maxIters=200)
    assert 3 < n < 100 and err < 1e-10
    Q = O.GtaProblem(om, g, sched, q["omega"], q["weight"], op, gs, PR.wtiso(2), q=q)
    Q.init_tt()
    b, bB = np.zeros(nc), np.zeros((8, nb))
    Q.grey_sweep(bB, b, True)
    Q.GreySource[:] = 0
    Mx, MB = corr.copy(), np.zeros((8, nb))
    Q.grey_sweep(MB, Mx, False)
    assert np.abs(b - (corr - Mx)).max() <= 1e-8 * np.abs(b).max()


# ---------------------------------------------------------------------------
# premise of the r-z record kernel's canonical quad labelling (csrc/sweeprz.cu, rz_rec_build_kernel)
# ---------------------------------------------------------------------------
def _canonical_fraction_rz(mesh, npolar=2, nazimuthal=2):
    """fraction of (angle, zone) pairs whose corner order from snnext is source -> its two neighbours (in either order) -> the
    opposite corner, with the EZ faces of every corner leading to the cyclic neighbours local+1 and local+3"""
    p = T.make_problem_rz(mesh, npolar, nazimuthal, 2)
    cez = np.asarray(mesh.cEZ).reshape(mesh.ncornr, 2) - 1
    good = total = 0
    for a in range(p.NA):
        if p.sched["nHyperPlanes"][a] == 0:
            continue
        nextC = p.sched["nextC"][a] - 1
        for z in range(mesh.nzones):
            total += 1
            if mesh.numCorner[z] != 4:
                continue
            c0 = mesh.cOffSet[z]
            ci = nextC[c0:c0 + 4]
            L0 = ci[0]
            L = [L0, (L0 + 1) & 3, (L0 + 3) & 3, (L0 + 2) & 3]
            ok = ci[3] == L[3] and {int(ci[1]), int(ci[2])} == {L[1], L[2]}
            for c in range(4):
                ok = ok and {int(cez[c0 + c, 0]), int(cez[c0 + c, 1])} == {(c + 1) & 3, (c + 3) & 3}
            good += bool(ok)
    return good / max(total, 1)


def test_rz_tiled_mesh_zones_fit_the_canonical_quad_labelling():
    """every zone of the tiled r-z mesh, for every swept ordinate of the product set, is solved source -> neighbours -> sink:
    the record kernel's compile-time neighbour table applies to the whole mesh (a mesh where it does not takes the by-corner records)"""
    for dims in ((3, 3, 0), (2, 5, 0)):
        assert _canonical_fraction_rz(M.tiled_mesh(dims)) == 1.0
