"""GPU parity: 3-D UCB sweep through the C ABI vs the oracle (SweepUCBxyz.F90)."""
import numpy as np
import pytest

from tests import common as T
from umt_b200 import mesh as M

pytestmark = pytest.mark.gpu
TOL = 1e-12   # north-star: per-sweep scalar intensity phi within 1e-12 relative


def _run_case(mesh, P, A, G, driver_like=False, sweeps=(False, True), **ctx_kw):
    p = T.make_problem_3d(mesh, P, A, G, driver_like=driver_like)
    ctx = T.gpu_context_3d(p, **ctx_kw)
    if p.sched["totalCycles"] > 0:
        ctx.init_radiation_field()      # cyclePsi <- Psi (also exit PsiB <- Psi)
        for a in range(p.NA):           # same on the oracle side
            for b, c in p.bdy[a]:
                p.PsiB[a, b - 1] = p.Psi[a, c - 1]
    for save in sweeps:
        phi_ref = T.oracle_sweep_3d(p, save)
        ctx.sweep(savePsi=save)
        phi = ctx.download_phi()
        assert T.relerr(phi, phi_ref) <= TOL
        assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
        if save:
            assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
    ctx.close()
    return p


def test_box_small():
    _run_case(M.box_mesh((3, 4, 5)), 1, 1, 3)


def test_tiled_random_state():
    _run_case(M.tiled_mesh((2, 2, 2)), 2, 2, 16)


def test_tiled_driver_problem():
    _run_case(M.tiled_mesh((2, 2, 3)), 2, 2, 2, driver_like=True, sweeps=(False, False, True))


def test_unstructured_box():
    _run_case(M.unstruct_box_mesh(2), 2, 2, 8)


def test_group_counts_ragged():
    for G in (1, 5, 33, 130):
        _run_case(M.box_mesh((3, 3, 3)), 1, 2, G, sweeps=(True,))


def test_warped_mesh_cycles_and_jacobi():
    m = M.box_mesh((4, 4, 4), warp=0.35, seed=3)
    p = _run_case(m, 2, 2, 4, sweeps=(False, False, True))
    assert p.sched["totalCycles"] > 0, "the warped mesh is meant to exercise the cycle list"


def test_strongly_warped_mesh_jacobi_branch():
    m = M.box_mesh((7, 6, 5), warp=0.9, seed=2)
    p = _run_case(m, 2, 2, 4, sweeps=(False, True))
    assert (p.sched["nextZ"] < 0).sum() > 0, "expected zones with an intra-zone cycle"


def test_library_built_inputs_match_oracle_inputs():
    """geometry, quadrature and schedule built by the library (not handed over by the caller)"""
    _run_case(M.tiled_mesh((2, 2, 2)), 2, 2, 8, own_schedule=True, own_geometry=True, own_quadrature=(2, 2, 1))
    _run_case(M.box_mesh((4, 4, 4), warp=0.35, seed=3), 2, 2, 4, own_schedule=True)


def test_device_geometry_matches_oracle():
    from umt_b200.teton import SweepContext
    from oracle import oracle as O
    for m in (M.tiled_mesh((2, 2, 2)), M.unstruct_box_mesh(2), M.box_mesh((4, 4, 4), warp=0.35, seed=3)):
        ctx = SweepContext.from_mesh(m, 1)
        ctx.compute_geometry(m.px)
        g = ctx.download_geometry()
        ref = O.geometry(O.OMesh(m))
        for k in ("Volume", "A_fp", "A_ez", "A_bdy", "VolumeZone"):
            scale = np.abs(ref[k]).max()
            assert np.abs(g[k] - ref[k]).max() <= 1e-13 * scale, k
        # geometry.cu is compiled without FMA contraction and evaluates the reference's expressions in the reference's order: the
        # area vectors (whose sign against omega decides upstream/downstream, ties included) agree to the last bit
        for k in ("A_fp", "A_ez", "A_bdy"):
            assert np.array_equal(g[k], ref[k]), k
        ctx.close()


def test_control_sweep_with_host_buffers_equals_upload_sweep_download():
    """umt_control_sweep (one ControlSweep with the caller's host arrays, chunked phi reduction overlapped with its download)
    gives bit for bit what umt_upload_state + umt_sweep + umt_download_phi give."""
    m = M.tiled_mesh((5, 5, 5))   # large enough for the chunked path (nc * G >= 2^22)
    p = T.make_problem_3d(m, 1, 1, 192)
    ctx = T.gpu_context_3d(p)
    ctx.sweep(savePsi=False)
    phi_a = ctx.download_phi()
    ctx2 = T.gpu_context_3d(p)
    phi_b = np.zeros_like(phi_a)
    it = ctx2.control_sweep(p.Sigt, p.STotal, p.tau, phi_b)
    assert it == 1 and np.array_equal(phi_a, phi_b)
    phi_c = np.full_like(phi_a, -1.0)
    ctx2.control_sweep(None, 2.0 * p.STotal, p.tau, phi_c)       # new source from the host, Sigt kept
    ctx.upload_state(None, None, None, 2.0 * p.STotal, p.tau)
    ctx.sweep(savePsi=False)
    assert np.array_equal(ctx.download_phi(), phi_c)
    ctx.close(); ctx2.close()


def test_control_sweep_sets_pipelined_group_sets():
    """umt_control_sweep_sets: the groups of a domain in three group sets (one context each, uneven sizes, the first large enough for
    the chunked phi download), uploads / sweeps / downloads pipelined across the sets.  Must equal, bit for bit, the same contexts
    swept one at a time, and the one-context sweep of all groups to rounding (groups do not couple, SweepUCBxyz.F90:119-281)."""
    import copy
    from umt_b200 import teton
    m = M.tiled_mesh((5, 5, 5))
    sizes = [192, 64, 128]
    G = sum(sizes)
    p = T.make_problem_3d(m, 1, 1, G)
    full = T.gpu_context_3d(p)
    bounds = np.cumsum([0] + sizes)

    def subset(k):
        q = copy.copy(p)
        sl = slice(bounds[k], bounds[k + 1])
        q.G = sizes[k]
        q.Psi, q.PsiB = np.ascontiguousarray(p.Psi[:, :, sl]), np.ascontiguousarray(p.PsiB[:, :, sl])
        q.Sigt, q.STotal = np.ascontiguousarray(p.Sigt[:, sl]), np.ascontiguousarray(p.STotal[:, sl])
        return q
    subs = [subset(k) for k in range(3)]
    sets = [T.gpu_context_3d(q) for q in subs]
    ones = [T.gpu_context_3d(q) for q in subs]
    for save in (False, True, False):
        scale = 2.0 if save else 1.0   # a new source from the host every call
        full.upload_state(None, None, None, scale * p.STotal, p.tau)
        full.sweep(savePsi=save)
        phi_full = full.download_phi()
        out = [np.full((m.ncornr, g), -1.0) for g in sizes]
        it = teton.control_sweep_sets(sets, [q.Sigt for q in subs], [scale * q.STotal for q in subs], p.tau, out, savePsi=save)
        assert it == 1
        for k, (c, q) in enumerate(zip(ones, subs)):
            ref = np.zeros((m.ncornr, sizes[k]))
            c.control_sweep(q.Sigt, scale * q.STotal, p.tau, ref, savePsi=save)
            assert np.array_equal(out[k], ref), (save, k)
            assert T.relerr(out[k], phi_full[:, bounds[k]:bounds[k + 1]]) <= 1e-13, (save, k)
    for k, c in enumerate(sets):
        assert T.mixed_err(c.download_psi(), full.download_psi()[:, :, bounds[k]:bounds[k + 1]], 1e-13) <= 1.0
        assert T.mixed_err(c.download_psib(), full.download_psib()[:, :, bounds[k]:bounds[k + 1]], 1e-13) <= 1.0
    for c in sets + ones + [full]:
        c.close()
