"""GPU parity of reflecting boundaries (snac/snreflect.F90): mirror angles, staged angle order, the row copies, and the
physical check that a half domain closed by a mirror plane reproduces the symmetric full-domain sweep."""
import numpy as np
import pytest

from tests import common as T
from umt_b200 import mesh as M

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _ctx(p):
    ctx = T.gpu_context_3d(p)
    for b in T.reflecting_boundaries(p.mesh):
        ctx.add_reflecting_boundary(b.first_elem, b.n_elem)
    return ctx


@pytest.mark.parametrize("sides", [(1,), (0, 2), (0, 2, 4), (0, 1), (0, 1, 2, 3, 4, 5)])
def test_reflecting_matches_oracle(sides):
    m = M.tiled_mesh((2, 2, 2), reflecting=sides)
    p = T.make_problem_3d(m, 2, 2, 6)
    ctx = _ctx(p)
    for save in (False, False, True):
        phi_ref, stage, mrefs = T.oracle_sweep_3d_reflecting(p, save)
        ctx.sweep(savePsi=save)
        assert list(ctx.reflect_stages()) == stage
        for k, mr in enumerate(mrefs):
            assert np.array_equal(ctx.reflected_angles(k), np.where(mr >= 0, mr + 1, -1))
        assert T.relerr(ctx.download_phi(), phi_ref) <= TOL
        assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
        if save:
            assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
    if sides in ((1,), (0, 2), (0, 2, 4)):   # orthogonal mirror planes: one extra stage per plane
        assert max(stage) == len(sides)
    ctx.close()


def test_mirror_plane_reproduces_symmetric_full_domain():
    """box 8x4x4 on [0,1]^3 with uniform data is symmetric about x = 1/2; the left half (4x4x4 on [0,1/2]x[0,1]^2) with a
    reflecting x+ side must give the same phi after ONE sweep (mirror images are swept one stage earlier)."""
    G = 3
    full = M.box_mesh((8, 4, 4))
    pf = T.make_problem_3d(full, 2, 2, G, driver_like=True)
    phi_full = T.oracle_sweep_3d(pf, False)
    half = M.box_mesh((4, 4, 4), lengths=(0.5, 1.0, 1.0), reflecting=(1,))
    ph = T.make_problem_3d(half, 2, 2, G, driver_like=True)
    ctx = _ctx(ph)
    ctx.sweep(savePsi=False)
    phi_half = ctx.download_phi()
    zc_f = np.repeat(np.add.reduceat(full.px, full.cOffSet) / 8.0, full.numCorner, axis=0)
    look = {tuple(np.round(np.r_[full.px[c], zc_f[c]] * 1e8).astype(np.int64)): c for c in range(full.ncornr)}
    zc_h = np.repeat(np.add.reduceat(half.px, half.cOffSet) / 8.0, half.numCorner, axis=0)
    idx = np.array([look[tuple(np.round(np.r_[half.px[c], zc_h[c]] * 1e8).astype(np.int64))] for c in range(half.ncornr)])
    assert T.relerr(phi_half, phi_full[idx]) <= 1e-11
    ctx.close()


# ---------------------------------------------------------------------------
# r-z
# ---------------------------------------------------------------------------
def _ctx_rz(p):
    ctx = T.gpu_context_rz(p)
    for b in T.reflecting_boundaries(p.mesh):
        ctx.add_reflecting_boundary(b.first_elem, b.n_elem)
    return ctx


@pytest.mark.parametrize("sides", [(2,), (3,), (2, 3), (1,), (1, 2)])
def test_reflecting_rz_matches_oracle(sides):
    """z-normal planes couple a xi-level to its partner level, the outer r-normal plane couples mu < 0 to mu > 0 inside a level"""
    m = M.tiled_mesh((2, 2, 0), reflecting=sides)
    p = T.make_problem_rz(m, 2, 2, 4)
    ctx = _ctx_rz(p)
    for save in (False, False, True):
        phi_ref, stage, mrefs = T.oracle_sweep_rz_reflecting(p, save)
        ctx.sweep(savePsi=save)
        assert list(ctx.reflect_stages()) == stage
        assert T.relerr(ctx.download_phi(), phi_ref) <= TOL
        assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
        if save:
            assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
    ctx.close()


def test_mirror_plane_rz_reproduces_symmetric_full_domain():
    """box 4 x 8 on [0,1]^2 (r, z) with uniform data is symmetric about z = 1/2; the lower half with a reflecting z+ side gives the
    same phi once the reflected levels have seen this pass's exiting flux (converged over a few sweeps: the level that is incident
    on the mirror is swept after its partner level)."""
    G = 3
    full = M.box_mesh((4, 8))
    pf = T.make_problem_rz(full, 2, 2, G, driver_like=True)
    half = M.box_mesh((4, 4), lengths=(1.0, 0.5), reflecting=(3,))
    ph = T.make_problem_rz(half, 2, 2, G, driver_like=True)
    ctx = _ctx_rz(ph)
    phi_full = T.oracle_sweep_rz(pf, False)
    ctx.sweep(savePsi=False)
    phi_half = ctx.download_phi()
    zc_f = np.repeat(np.add.reduceat(full.px, full.cOffSet) / 4.0, full.numCorner, axis=0)
    look = {tuple(np.round(np.r_[full.px[c], zc_f[c]] * 1e8).astype(np.int64)): c for c in range(full.ncornr)}
    zc_h = np.repeat(np.add.reduceat(half.px, half.cOffSet) / 4.0, half.numCorner, axis=0)
    idx = np.array([look[tuple(np.round(np.r_[half.px[c], zc_h[c]] * 1e8).astype(np.int64))] for c in range(half.ncornr)])
    assert T.relerr(phi_half, phi_full[idx]) <= 1e-11
    ctx.close()
