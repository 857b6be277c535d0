"""GPU parity: 2-D (r,z) UCB sweep through the C ABI vs the oracle (SweepUCBrz.F90, SetSweep.F90 angle loop)."""
import numpy as np
import pytest

from tests import common as T
from umt_b200 import mesh as M

pytestmark = pytest.mark.gpu
TOL = 1e-12   # north-star: per-sweep scalar intensity phi within 1e-12 relative


def _run_case(mesh, P, A, G, driver_like=False, sweeps=(False, True), **ctx_kw):
    p = T.make_problem_rz(mesh, P, A, G, driver_like=driver_like)
    ctx = T.gpu_context_rz(p, **ctx_kw)
    for save in sweeps:
        phi_ref = T.oracle_sweep_rz(p, save)
        ctx.sweep(savePsi=save)
        phi = ctx.download_phi()
        assert T.relerr(phi, phi_ref) <= TOL
        assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0      # incl. PsiB(:,:,finishing) <- PsiM
        if save:
            assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0     # incl. Psi(:,:,finishing) <- PsiM
    ctx.close()
    return p


def test_box_small():
    _run_case(M.box_mesh((4, 5)), 1, 1, 3)


def test_tiled_random_state():
    _run_case(M.tiled_mesh((3, 3, 0)), 2, 2, 16)


def test_tiled_driver_problem():
    """BASELINE configs[1] shape (2-D tiled mesh, G = 64, default P2 A2) at a size the oracle finishes in seconds."""
    _run_case(M.tiled_mesh((6, 6, 0)), 2, 2, 64, driver_like=True, sweeps=(False, False, True))


def test_group_counts_ragged():
    for G in (1, 5, 33, 130):
        _run_case(M.box_mesh((5, 5)), 1, 2, G, sweeps=(True,))


def test_high_order_quadrature():
    _run_case(M.tiled_mesh((2, 2, 0)), 3, 4, 4)


def test_warped_mesh():
    _run_case(M.box_mesh((8, 8), warp=0.3, seed=5), 2, 2, 4, sweeps=(False, True))


def test_library_built_inputs_match_oracle_inputs():
    _run_case(M.tiled_mesh((3, 2, 0)), 2, 2, 8, own_schedule=True, own_geometry=True, own_quadrature=(2, 2, 1))


def test_uniform_solution_preserved():
    m = M.tiled_mesh((3, 3, 0))
    p = T.make_problem_rz(m, 2, 2, 4)
    psi0 = np.linspace(0.7, 1.9, 4)
    p.Psi[:] = psi0
    p.PsiB[:] = psi0
    p.STotal[:] = (np.repeat(p.Sigt, m.numCorner, axis=0) - p.tau) * psi0
    ctx = T.gpu_context_rz(p)
    ctx.sweep(savePsi=True)
    assert np.abs(ctx.download_psi() / psi0 - 1).max() <= 1e-12
    assert np.abs(ctx.download_phi() / (2 * np.pi * psi0) - 1).max() <= 1e-12
    ctx.close()


@pytest.mark.parametrize("kernel", ["pipe", "lc", "recflow", "item"])
def test_optional_rz_kernels_match_oracle(kernel):
    """The r-z kernel variants kept as options (the TMA pipeline, the level-chain kernel, the dataflow
    variant on the records, the plain item kernel) stay parity-checked."""
    import os
    os.environ["UMT_RZ_KERNEL"] = kernel
    try:
        _run_case(M.tiled_mesh((3, 3, 0)), 2, 2, 16)
        _run_case(M.tiled_mesh((4, 2, 0)), 2, 2, 64, driver_like=True, sweeps=(False, True))
    finally:
        del os.environ["UMT_RZ_KERNEL"]


def test_control_sweep_sets_rz_group_sets():
    """umt_control_sweep_sets on r-z contexts (two group sets, pipelined) == the same contexts swept one at a time, bit for bit, and
    the oracle's sweep of all groups."""
    import copy
    from umt_b200 import teton
    m = M.tiled_mesh((4, 4, 0))
    sizes = [32, 16]
    p = T.make_problem_rz(m, 2, 2, sum(sizes))
    bounds = np.cumsum([0] + sizes)

    def subset(k):
        q = copy.copy(p)
        sl = slice(bounds[k], bounds[k + 1])
        q.G = sizes[k]
        q.Psi, q.PsiB = np.ascontiguousarray(p.Psi[:, :, sl]), np.ascontiguousarray(p.PsiB[:, :, sl])
        q.Sigt, q.STotal = np.ascontiguousarray(p.Sigt[:, sl]), np.ascontiguousarray(p.STotal[:, sl])
        return q
    subs = [subset(k) for k in range(2)]
    sets = [T.gpu_context_rz(q) for q in subs]
    ones = [T.gpu_context_rz(q) for q in subs]
    for save in (False, True):
        phi_ref = T.oracle_sweep_rz(p, save)
        out = [np.full((m.ncornr, g), -1.0) for g in sizes]
        assert teton.control_sweep_sets(sets, [q.Sigt for q in subs], [q.STotal for q in subs], p.tau, out, savePsi=save) == 1
        for k, (c, q) in enumerate(zip(ones, subs)):
            ref = np.zeros((m.ncornr, sizes[k]))
            c.control_sweep(q.Sigt, q.STotal, p.tau, ref, savePsi=save)
            assert np.array_equal(out[k], ref), (save, k)
            assert T.relerr(out[k], phi_ref[:, bounds[k]:bounds[k + 1]]) <= TOL, (save, k)
    for k, c in enumerate(sets):
        assert T.mixed_err(c.download_psi(), p.Psi[:, :, bounds[k]:bounds[k + 1]], TOL) <= 1.0
    for c in sets + ones:
        c.close()
