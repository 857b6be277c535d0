"""GPU: SURVEY section 8(f) N4 -- the linear solve of a temperature iteration with grey acceleration on (rt/LinearSolver.F90:55-121:
collision rate -> sweep -> residual -> GTASolver -> addGreyCorrections) plus the rebuild of the isotropic source between
iterations, driven through the C ABI by umt_b200.cycle.FullPhysicsStep, against the same sequence made of the oracle's pieces.
The mini-app reference compiles this loop out and never fills STotal: opacities are synthetic (seeded) and the source formula is
the library's documented extension, so this parity is unpinned by construction."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T
from umt_b200 import mesh as M
from umt_b200 import problem as PR
from umt_b200.cycle import FullPhysicsStep

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,mk", [("tiled", lambda: M.tiled_mesh((2, 2, 2))), ("unstruct", lambda: M.unstruct_box_mesh(2))])
def test_linear_solver_iterations_match_oracle_pieces(name, mk):
    mesh = mk()
    G = 4
    nz, nc = mesh.nzones, mesh.ncornr
    rng = np.random.default_rng(21)
    Siga = 5 * rng.random((nz, G))
    Sigs = 20 * rng.random((nz, G))
    Eta = 0.5 * rng.random(nc)
    Chi = rng.random((nc, G))
    Chi /= Chi.sum(1, keepdims=True)
    Emis = 0.1 * rng.random((nc, G))
    c2z = np.repeat(np.arange(nz), mesh.numCorner)
    wt = PR.wtiso(3)

    p = T.make_problem_3d(mesh, 2, 2, G, seed=5)
    p.tau = PR.tau()
    p.Sigt = Siga + Sigs + p.tau
    ctx = T.gpu_context_3d(p)
    ctx.init_phi_total()
    phi = np.einsum("a,acg->cg", p.weight, p.Psi)          # initPhiTotal
    assert T.relerr(ctx.download_phi(), phi) <= 1e-13
    step = FullPhysicsStep(ctx, mesh, G, Siga, Sigs, Eta, Chi, Emis, flux_iters=1)

    # the oracle's GTA set-up for the same opacities
    g_om, g_w = O.gta_quad_xyz()
    g_sched = O.schedule(p.om, p.geom, g_om)
    chi_ref = Chi.copy()
    op = O.gta_set_opacity(p.om, p.geom, p.tau, Siga, Sigs, Eta, chi_ref)
    assert T.relerr(step.Chi, chi_ref) <= 1e-12
    grey = np.zeros(nc)
    for it, save in enumerate((False, False, True)):
        # source rebuild from the current PhiTotal (extension formula), on both sides
        p.STotal = wt * (Sigs[c2z] * phi + chi_ref * (Eta * (Siga[c2z] * phi).sum(1))[:, None] + Emis)
        st = step.rebuild_source()
        assert np.abs(st - p.STotal).max() <= 1e-8 * np.abs(p.STotal).max()
        # LinearSolver.F90:55-121 from the oracle's pieces
        O.collision_rate(p.om, Eta, Siga, Sigs, phi, grey, 0)
        phi = T.oracle_sweep_3d(p, save)
        O.collision_rate(p.om, Eta, Siga, Sigs, phi, grey, 1)
        P = O.GtaProblem(p.om, p.geom, g_sched, g_om, g_w, op, grey, wt)
        corr, n, err = P.solve(phi)
        phi = O.add_grey_corrections(corr, chi_ref, phi)
        rec = step.linear_solver(save)
        assert rec["grey_sweeps"] == n and n > 3, (it, rec, n)
        got = ctx.download_phi()
        # the Krylov recurrences amplify last-bit differences (FMA contraction, reduction order): 1e-8 of the field's scale
        assert np.abs(got - phi).max() <= 3e-8 * (it + 1) * np.abs(phi).max(), it
        assert abs(rec["correction_max"] - np.abs(corr).max()) <= 1e-7 * np.abs(corr).max()
    # the savePsi sweep of the last iteration (its source carries the 1e-8 of the grey solves before it)
    assert np.abs(ctx.download_psi() - p.Psi).max() <= 1e-7 * np.abs(p.Psi).max()
    ctx.close()
