"""SweepScheduler (rt/SweepScheduler.F90) + setNetFlux and the per-step exchange of SetSweep.F90:113-170 for comm sets that
hold several angle bins, against the lock-step oracle of tests/common.py.  Domains live on one GPU and talk through the
in-process communicator."""
import numpy as np
import pytest

from tests import common as T
from umt_b200 import mesh as M
from umt_b200 import teton

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(N, dims, P, A, G, nCommSets, sides=()):
    problems = [T.make_problem_3d(M.tiled_mesh(dims, rank=r, size=N), P, A, G, seed=200 + r) for r in range(N)]
    ctxs = []
    for p in problems:
        ctx = T.gpu_context_3d(p, own_schedule=False)
        for b in T.shared_boundaries(p.mesh):
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
        ctx.set_comm_sets(nCommSets)
        ctxs.append(ctx)
    teton.connect_local(ctxs)
    T.run_local_group(ctxs, lambda r, c: c.build_exchange())
    return problems, ctxs


@pytest.mark.parametrize("N,dims,nCommSets", [(2, (2, 2, 1), 1), (2, (2, 2, 1), 4), (4, (2, 1, 1), 2), (8, (1, 1, 1), 1)])
def test_scheduler_order_matches_oracle(N, dims, nCommSets):
    problems, ctxs = _setup(N, dims, 1, 2, 2, nCommSets)
    NA = problems[0].NA
    lists = T.oracle_exchange_lists(problems)
    # (a) explicit net flux: exact comparison, including the neighbours' choices coming back as RecvOrder
    rng = np.random.default_rng(9)
    nf = [rng.standard_normal((len(T.shared_boundaries(p.mesh)), NA)) for p in problems]
    order, recv = T.oracle_sweep_scheduler(problems, nCommSets, nf)
    T.run_local_group(ctxs, lambda r, c: c.sweep_scheduler(nf[r]))
    for r, ctx in enumerate(ctxs):
        ao, ro = ctx.angle_order(len(nf[r]))
        assert np.array_equal(ao - 1, order[r])
        for k in range(len(nf[r])):
            assert np.array_equal(ro[k] - 1, recv[r][k])
        assert sorted(ao.tolist()) == list(range(1, NA + 1))
    # (b) net flux tallied on the device from the (random) boundary fluxes: setNetFlux
    nf2 = T.oracle_net_flux(problems, lists)
    T.run_local_group(ctxs, lambda r, c: c.sweep_scheduler(None))
    nf_dev = [ctx.net_flux(len(nf2[r])) for r, ctx in enumerate(ctxs)]
    for r in range(N):
        assert np.abs(nf_dev[r] - nf2[r]).max() <= 1e-12 * np.abs(nf2[r]).max()
    # bins whose neighbours are all done compete on rounding residue of depend: the order is checked on the device's own net flux
    order2, recv2 = T.oracle_sweep_scheduler(problems, nCommSets, nf_dev)
    for r, ctx in enumerate(ctxs):
        ao, ro = ctx.angle_order(len(nf2[r]))
        assert np.array_equal(ao - 1, order2[r])
        for k in range(len(nf2[r])):
            assert np.array_equal(ro[k] - 1, recv2[r][k])
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("N,dims,nCommSets", [(2, (2, 2, 1), 1), (2, (2, 2, 1), 4), (4, (2, 1, 1), 2)])
def test_ordered_sweeps_with_stepwise_exchange_match_oracle(N, dims, nCommSets):
    problems, ctxs = _setup(N, dims, 1, 2, 4, nCommSets)
    lists = T.oracle_exchange_lists(problems)
    T.run_local_group(ctxs, lambda r, c: c.sweep_scheduler(None))
    nf = [ctx.net_flux(len(T.shared_boundaries(p.mesh))) for p, ctx in zip(problems, ctxs)]
    order, recv = T.oracle_sweep_scheduler(problems, nCommSets, nf)
    for r, ctx in enumerate(ctxs):
        assert np.array_equal(ctx.angle_order()[0] - 1, order[r])
    for save, iters in ((False, 1), (False, 3), (True, 2)):
        phis, it_ref, inc_ref = T.oracle_multi_sweep_ordered(problems, lists, nCommSets, order, save, iters, 1e-6)
        its = T.run_local_group(ctxs, lambda r, c: c.sweep(save, iters, 1e-6))
        assert its == [it_ref] * N
        for r, (p, ctx) in enumerate(zip(problems, ctxs)):
            assert T.relerr(ctx.download_phi(), phis[r]) <= TOL
            assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
            if save:
                assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
    for c in ctxs:
        c.close()


def test_one_comm_set_needs_fewer_flux_passes_than_the_lagged_default():
    """What the scheduler is for: with every angle in one comm set a neighbour that sweeps an angle later in the pass receives
    this pass's exiting flux, so the incident-flux iteration converges in fewer passes than with the fully lagged exchange."""
    N, dims = 2, (2, 2, 2)
    passes = {}
    for nCommSets in (0, 1):
        problems, ctxs = _setup(N, dims, 1, 1, 2, nCommSets)
        if nCommSets:
            T.run_local_group(ctxs, lambda r, c: c.sweep_scheduler(None))
        its = T.run_local_group(ctxs, lambda r, c: c.sweep(False, 50, 1e-8))
        passes[nCommSets] = its[0]
        for c in ctxs:
            c.close()
    assert passes[1] <= passes[0] and passes[0] > 1


# ---------------------------------------------------------------------------
# r-z: the angle bins are the xi-levels (SweepScheduler.F90:110-117); BASELINE configs[1] is the r-z multi-domain case
# ---------------------------------------------------------------------------
def _setup_rz(N, dims, G, nCommSets):
    problems = [T.make_problem_rz(M.tiled_mesh(dims, rank=r, size=N), 2, 2, G, seed=300 + r) for r in range(N)]
    ctxs = []
    for p in problems:
        ctx = T.gpu_context_rz(p)
        for b in T.shared_boundaries(p.mesh):
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
        ctx.set_comm_sets(nCommSets)
        ctxs.append(ctx)
    teton.connect_local(ctxs)
    T.run_local_group(ctxs, lambda r, c: c.build_exchange())
    return problems, ctxs


@pytest.mark.parametrize("N,dims,nCommSets", [(2, (3, 3, 0), 1), (4, (3, 3, 0), 2), (4, (2, 3, 0), 1)])
def test_rz_scheduler_and_ordered_sweeps_match_oracle(N, dims, nCommSets):
    problems, ctxs = _setup_rz(N, dims, 4, nCommSets)
    NA = problems[0].NA
    lev, angles = T.rz_bins(problems[0])
    nBins = len(angles)
    lists = T.oracle_exchange_lists(problems)
    # net flux per xi-level tallied on the device (setNetFlux), then the order on the device's own numbers
    nf_ref = T.oracle_net_flux_bins(problems, lists)
    T.run_local_group(ctxs, lambda r, c: c.sweep_scheduler(None))
    nf_dev = [ctx.net_flux(len(nf_ref[r]))[:, :nBins] for r, ctx in enumerate(ctxs)]
    for r in range(N):
        assert np.abs(nf_dev[r] - nf_ref[r]).max() <= 1e-12 * np.abs(nf_ref[r]).max()
    binOrder, binRecv = T.oracle_sweep_scheduler(problems, nCommSets, nf_dev, nBins=nBins)
    for r, ctx in enumerate(ctxs):
        ao, ro = ctx.angle_order(len(nf_ref[r]))
        assert np.array_equal(ao - 1, np.concatenate([angles[b] for b in binOrder[r]]))    # CSet%AngleOrder: the levels' angles in level order
        for k in range(len(nf_ref[r])):
            assert np.array_equal(ro[k] - 1, np.concatenate([angles[b] for b in binRecv[r][k]]))
        assert sorted(ao.tolist()) == list(range(1, NA + 1))
    for save, iters in ((False, 1), (False, 3), (True, 2)):
        phis, it_ref, inc_ref = T.oracle_multi_sweep_ordered_rz(problems, lists, nCommSets, binOrder, save, iters, 1e-6)
        its = T.run_local_group(ctxs, lambda r, c: c.sweep(save, iters, 1e-6))
        assert its == [it_ref] * N
        for r, (p, ctx) in enumerate(zip(problems, ctxs)):
            assert T.relerr(ctx.download_phi(), phis[r]) <= TOL
            assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
            inc, _old = ctx.incident_flux(nBins)
            assert np.abs(inc - inc_ref[r]).max() <= 1e-12 * max(np.abs(inc_ref[r]).max(), 1e-300)
            if save:
                assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
    for c in ctxs:
        c.close()
