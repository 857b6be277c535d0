"""GPU parity of the multi-domain path: psib exchange lagged one flux pass, exit-current tallies and the
incident-flux convergence test (SetSweep.F90:68-207, findexit.F90, SendFlux/RecvFlux, setIncidentFlux,
testFluxConv) against the oracle run in lock step on every domain.  The domains live on one GPU and talk
through the library's in-process communicator; tests/test_nccl_ranks.py covers the NCCL transport."""
import numpy as np
import pytest

from tests import common as T
from umt_b200 import mesh as M
from umt_b200 import teton

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(mk, N, P, A, G, driver_like):
    problems = [T.make_problem_3d(mk(r, N), P, A, G, seed=100 + r, driver_like=driver_like) for r in range(N)]
    ctxs = []
    for r, p in enumerate(problems):
        ctx = T.gpu_context_3d(p, own_schedule=False)
        for b in T.shared_boundaries(p.mesh):
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
        ctxs.append(ctx)
    teton.connect_local(ctxs)
    T.run_local_group(ctxs, lambda r, c: c.build_exchange())
    return problems, ctxs


def _check_lists(problems, ctxs, lists):
    for r, (p, ctx) in enumerate(zip(problems, ctxs)):
        for k, _b in enumerate(T.shared_boundaries(p.mesh)):
            for a in range(p.NA):
                ls, lr = ctx.exchange_lists(k, a + 1)
                assert np.array_equal(ls, lists[r][k][a][0]) and np.array_equal(lr, lists[r][k][a][1])


@pytest.mark.parametrize("N,dims", [(2, (2, 2, 2)), (4, (2, 1, 2)), (8, (1, 1, 2))])
def test_lagged_exchange_matches_oracle(N, dims):
    problems, ctxs = _setup(lambda r, n: M.tiled_mesh(dims, rank=r, size=n), N, 1, 2, 4, driver_like=False)
    lists = T.oracle_exchange_lists(problems)
    _check_lists(problems, ctxs, lists)
    for save, iters in ((False, 1), (False, 3), (True, 5)):
        phis, it_ref, inc_ref = T.oracle_multi_sweep_3d(problems, lists, save, iters, 1e-6)
        its = T.run_local_group(ctxs, lambda r, c: c.sweep(save, iters, 1e-6))
        assert its == [it_ref] * N
        for r, (p, ctx) in enumerate(zip(problems, ctxs)):
            assert T.relerr(ctx.download_phi(), phis[r]) <= TOL
            assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
            inc, _old = ctx.incident_flux()
            assert np.abs(inc - inc_ref[r]).max() <= 1e-12 * max(np.abs(inc_ref[r]).max(), 1e-300)
            if save:
                assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("N,dims", [(2, (3, 3, 0)), (4, (3, 3, 0))])
def test_lagged_exchange_rz(N, dims):
    """BASELINE configs[1] shape: 2-D (r,z) tiled mesh, 2 x 2 domains; angle sets are xi-levels (the lower rank classifies the
    first half of each level), one flux-convergence bin per level."""
    problems = [T.make_problem_rz(M.tiled_mesh(dims, rank=r, size=N), 2, 2, 4, seed=100 + r) for r in range(N)]
    ctxs = []
    for p in problems:
        ctx = T.gpu_context_rz(p)
        for b in T.shared_boundaries(p.mesh):
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
        ctxs.append(ctx)
    teton.connect_local(ctxs)
    T.run_local_group(ctxs, lambda r, c: c.build_exchange())
    lists = T.oracle_exchange_lists(problems)
    _check_lists(problems, ctxs, lists)
    nBins = int(problems[0].q["level"].max())
    for save, iters in ((False, 1), (False, 4), (True, 2)):
        phis, it_ref, inc_ref = T.oracle_multi_sweep(problems, lists, save, iters, 1e-6)
        its = T.run_local_group(ctxs, lambda r, c: c.sweep(save, iters, 1e-6))
        assert its == [it_ref] * N
        for r, (p, ctx) in enumerate(zip(problems, ctxs)):
            assert T.relerr(ctx.download_phi(), phis[r]) <= TOL
            assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
            inc, _old = ctx.incident_flux(nBins)
            assert np.abs(inc - inc_ref[r]).max() <= 1e-12 * max(np.abs(inc_ref[r]).max(), 1e-300)
    for c in ctxs:
        c.close()


def test_flux_iteration_converges_to_single_domain_solution():
    """With enough flux passes the decomposed problem reproduces the single-domain sweep: the exchange
    moves the right rows to the right corners.  box mesh 4x4x8 split 1x1x2 vs the same mesh on one domain."""
    n, G = (4, 4, 4), 3
    N = 2
    problems, ctxs = _setup(lambda r, nn: M.box_mesh(n, rank=r, size=nn), N, 1, 1, G, driver_like=True)
    its = T.run_local_group(ctxs, lambda r, c: c.sweep(False, 6, 0.0))
    assert its[0] == its[1] and 2 <= its[0] <= 4
    whole = M.box_mesh((4, 4, 8))
    pw = T.make_problem_3d(whole, 1, 1, G, driver_like=True)
    phi_w = T.oracle_sweep_3d(pw, False)
    key_w = {tuple(np.round(np.r_[pw.mesh.px[c], pw.geom["Volume"][c]] * 1e9).astype(np.int64)): c for c in range(whole.ncornr)}
    zc_w = np.repeat(np.add.reduceat(whole.px, whole.cOffSet) / 8.0, whole.numCorner, axis=0)
    look = {tuple(np.round(np.r_[whole.px[c], zc_w[c]] * 1e8).astype(np.int64)): c for c in range(whole.ncornr)}
    for r, (p, ctx) in enumerate(zip(problems, ctxs)):
        m = p.mesh
        zc = np.repeat(np.add.reduceat(m.px, m.cOffSet) / 8.0, m.numCorner, axis=0)
        idx = np.array([look[tuple(np.round(np.r_[m.px[c], zc[c]] * 1e8).astype(np.int64))] for c in range(m.ncornr)])
        phi = ctx.download_phi()
        assert T.relerr(phi, phi_w[idx]) <= 1e-11
    for c in ctxs:
        c.close()
    del key_w
