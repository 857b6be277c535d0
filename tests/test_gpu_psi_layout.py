"""GPU: the two storage layouts of the angular flux (include/umt_sweep.h, "device memory of the angular flux").

single-psi (3-D, no cycle lists / direct-solve zones / reflecting boundaries / staged comm sets): Psi once, Psi1 in a ring of
angle batches, retiring batches tallied into PhiTotal by phi-tally items inside the sweep kernel, savePsi sweeps in place.
legacy: a full second buffer.  Both must give the oracle's PhiTotal / Psi / PsiB; the ring must not change a bit of PhiTotal
(same angle order of the sum); the footprint must be what umt_get_psi_layout reports."""
import os

import numpy as np
import pytest

from tests import common as T
from umt_b200 import mesh as M

pytestmark = pytest.mark.gpu
TOL = 1e-12


class _env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update({k: str(v) for k, v in self.kw.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _sweeps(p, ctx, sweeps=(False, False, True)):
    out = []
    for save in sweeps:
        phi_ref = T.oracle_sweep_3d(p, save)
        ctx.sweep(savePsi=save)
        phi = ctx.download_phi()
        assert T.relerr(phi, phi_ref) <= TOL
        assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
        out.append(phi)
    assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
    return out


@pytest.mark.parametrize("ring,batch", [(2, 2), (3, 4), (5, 2), (0, 4)])
def test_ring_of_angle_batches_matches_oracle_and_full_ring_bitwise(ring, batch):
    mesh = M.tiled_mesh((3, 2, 2))
    G = 16
    with _env(UMT_ANGLE_BATCH=batch, UMT_PHI_CHUNK=64):
        p = T.make_problem_3d(mesh, 2, 2, G)
        ctx = T.gpu_context_3d(p)
        ctx.set_psi1_ring(ring)
        lay = ctx.psi_layout()
        assert lay["angle_batch"] == batch
        if ring:
            assert lay["single"] and lay["batches"] == 32 // batch
            assert lay["ring_batches"] == ring and lay["psi1_slabs"] == ring * batch
            assert lay["angles_tallied_in_sweep"] == 32 - ring * batch
        else:   # everything fits: the two-buffer layout is the fastest
            assert not lay["single"] and lay["psi1_slabs"] == 32 and lay["angles_tallied_in_sweep"] == 0
        assert lay["bytes"] == 8.0 * G * (mesh.ncornr + mesh.nbelem) * (32 + lay["psi1_slabs"])
        phis = _sweeps(p, ctx)
        ctx.close()
        # the same sweeps with every angle in its own slab: PhiTotal must agree to the last bit
        p2 = T.make_problem_3d(mesh, 2, 2, G)
        ctx2 = T.gpu_context_3d(p2)
        ctx2.set_psi1_ring(10 ** 6)
        assert not ctx2.psi_layout()["single"] and ctx2.psi_layout()["angles_tallied_in_sweep"] == 0
        phis2 = _sweeps(p2, ctx2)
        ctx2.close()
    for a, b in zip(phis, phis2):
        assert np.array_equal(a, b)


def test_legacy_layout_forced_on_a_plain_mesh_still_matches():
    mesh = M.tiled_mesh((2, 2, 2))
    with _env(UMT_PSI_LAYOUT="legacy"):
        p = T.make_problem_3d(mesh, 2, 2, 8)
        ctx = T.gpu_context_3d(p)
        assert not ctx.psi_layout()["single"] and ctx.psi_layout()["psi1_slabs"] == 32
        _sweeps(p, ctx)
        ctx.close()


def test_ring_changes_between_sweeps_keep_psib_and_psi():
    """Switching the ring size (and with it the work-item list) between sweeps must not lose state: PsiB lives in the Psi tails."""
    mesh = M.tiled_mesh((2, 2, 2))
    with _env(UMT_ANGLE_BATCH=4):
        p = T.make_problem_3d(mesh, 2, 2, 8)
        ctx = T.gpu_context_3d(p)
        for ring, save in ((2, False), (0, False), (3, True), (2, False)):
            ctx.set_psi1_ring(ring)
            phi_ref = T.oracle_sweep_3d(p, save)
            ctx.sweep(savePsi=save)
            assert T.relerr(ctx.download_phi(), phi_ref) <= TOL
            assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
        assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
        ctx.close()


def test_mesh_with_cycle_lists_keeps_the_full_workspace():
    mesh = M.box_mesh((4, 4, 4), warp=0.35, seed=3)
    p = T.make_problem_3d(mesh, 2, 2, 4)
    ctx = T.gpu_context_3d(p)
    cyc = p.sched["totalCycles"] > 0 or any(int((p.sched["nextZ"][a] < 0).sum()) > 0 for a in range(p.NA))
    assert cyc, "the warped mesh is meant to have cycle lists or direct-solve zones"
    ctx.set_psi1_ring(1)   # even when a small ring is asked for
    lay = ctx.psi_layout()
    assert not lay["single"] and lay["psi1_slabs"] == p.NA
    ctx.close()


def test_multi_domain_exchange_on_the_ring_layout():
    """lock-step oracle parity of the lagged psib exchange with a 2-batch ring on every domain"""
    from umt_b200 import teton
    N = 2
    with _env(UMT_ANGLE_BATCH=2, UMT_RING_BATCHES=2, UMT_PHI_CHUNK=64):
        problems = [T.make_problem_3d(M.tiled_mesh((2, 2, 2), rank=r, size=N), 1, 2, 4, seed=100 + r) for r in range(N)]
        ctxs = []
        for p in problems:
            ctx = T.gpu_context_3d(p)
            for b in T.shared_boundaries(p.mesh):
                ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
            ctxs.append(ctx)
        teton.connect_local(ctxs)
        T.run_local_group(ctxs, lambda r, c: c.build_exchange())
        lists = T.oracle_exchange_lists(problems)
        for save, iters in ((False, 3), (True, 2)):
            phis, it_ref, _inc = T.oracle_multi_sweep_3d(problems, lists, save, iters, 1e-6)
            its = T.run_local_group(ctxs, lambda r, c: c.sweep(save, iters, 1e-6))
            assert its == [it_ref] * N
            for r, (p, ctx) in enumerate(zip(problems, ctxs)):
                assert ctx.psi_layout()["angles_tallied_in_sweep"] == 16 - 4
                assert T.relerr(ctx.download_phi(), phis[r]) <= TOL
                assert T.mixed_err(ctx.download_psib(), p.PsiB, TOL) <= 1.0
                if save:
                    assert T.mixed_err(ctx.download_psi(), p.Psi, TOL) <= 1.0
        for c in ctxs:
            c.close()


def test_configs2_quadrature_P4_A4_on_the_ring_matches_oracle():
    """BASELINE configs[2]'s quadrature (-P 4 -A 4: 128 angles) at a size the oracle finishes in seconds, psi stored once and Psi1 in
    a ring of 3 of its 32 angle batches (what the -d 16 -G 128 bench run of that config does with 13 of 32)."""
    mesh = M.tiled_mesh((3, 3, 3))
    with _env(UMT_ANGLE_BATCH=4, UMT_PHI_CHUNK=256):
        p = T.make_problem_3d(mesh, 4, 4, 8)
        assert p.NA == 128
        ctx = T.gpu_context_3d(p, own_schedule=True, own_geometry=True, own_quadrature=(4, 4, 1))
        ctx.set_psi1_ring(3)
        lay = ctx.psi_layout()
        assert lay["single"] and lay["psi1_slabs"] == 12 and lay["angles_tallied_in_sweep"] == 116
        _sweeps(p, ctx)
        ctx.close()
