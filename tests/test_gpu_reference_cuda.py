"""Pins the 3-D sweep to REAL reference code: Teton's own CUDA implementation of SweepUCBxyz
(`/root/reference/src/teton/gpu/GPU_SweepUCBxyz.cu`, entry point `gpu_sweepucbxyz` :532-994, kernel :172-516) is compiled
unmodified by `oracle/Makefile` into `oracle/_ref/libgpu_sweepucbxyz_ref.so` and run here, on the same inputs, next to

  (a) the oracle's restatement of SweepUCBxyz.F90 (angle by angle: Psi1 rows, exiting PsiB, Phi contribution, cyclePsi), and
  (b) this library's `umt_sweep` through the C ABI (PhiTotal, PsiB, Psi after savePsi).

The reference kernel is the same arithmetic as `snac/SweepUCBxyz.F90:99-322` phrased per (zone, group) thread; what it does
not cover (schedules, geometry, quadrature are inputs here) stays pinned by invariants only."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T
from umt_b200 import mesh as M

pytestmark = pytest.mark.gpu
TOL = 1e-12
_next_stream = [0]


def _stream_id():
    """the reference shim sizes its static device buffers at the first call of a stream id (:650-758)"""
    s = _next_stream[0]
    _next_stream[0] += 1
    assert s < 80
    return s


def _need_ref():
    assert O.ref_cuda_available(), ("oracle/_ref/libgpu_sweepucbxyz_ref.so is missing: run `make -C oracle` where "
                                    "/root/reference and nvcc exist (build() does); the file travels with the snapshot")


def _reference_pass(p, savePsi, sid):
    """One flux pass angle by angle through the reference's CUDA shim; mutates p.Psi (if savePsi), p.PsiB, p.cyclePsi the way
    SetSweep_CUDA.F90 does and returns (PhiTotal, Psi1 of every angle)."""
    m = p.mesh
    nc, nb, G = m.ncornr, m.nbelem, p.G
    phi_total = np.zeros((nc, G))
    psi1_all = np.zeros((p.NA, nc, G))
    for a in range(p.NA):
        Psi1 = np.zeros((nc + nb, G))
        Phi = np.zeros((nc, G))
        O.ref_cuda_sweep_xyz(p.om, p.geom, p.sched, a, p.omega, p.weight, p.tau, p.STotal, p.Sigt, p.Psi[a], Psi1, p.PsiB[a],
                             Phi, p.cyclePsi, savePsi, stream_id=sid)
        phi_total += Phi
        psi1_all[a] = Psi1[:nc]
    return phi_total, psi1_all


def _oracle_pass(p, savePsi):
    m = p.mesh
    nc, nb, G = m.ncornr, m.nbelem, p.G
    s = p.sched
    phi_total = np.zeros((nc, G))
    psi1_all = np.zeros((p.NA, nc, G))
    for a in range(p.NA):
        Psi1 = np.zeros((nc + nb, G))
        Phi = np.zeros((nc, G))
        off, n = int(s["cycleOffSet"][a]), int(s["numCycles"][a])
        for k in range(off, off + n):                         # initFromCycleList
            Psi1[s["cycleList"][k] - 1] = p.cyclePsi[k]
        O.sweep_xyz(p.om, p.geom, s, a, p.omega, p.weight, p.tau, p.STotal, p.Sigt, p.Psi[a], Psi1, p.PsiB[a], Phi, savePsi)
        for k in range(off, off + n):                         # updateCycleList
            p.cyclePsi[k] = Psi1[s["cycleList"][k] - 1]
        phi_total += Phi
        psi1_all[a] = Psi1[:nc]
    return phi_total, psi1_all


def _clone_state(p):
    q = T.Problem()
    q.__dict__.update(p.__dict__)
    q.Psi, q.PsiB, q.cyclePsi = p.Psi.copy(), p.PsiB.copy(), p.cyclePsi.copy()
    return q


def _seed_cycles(p):
    """initializeRadiationField: cyclePsi <- Psi (make_problem_3d did that), exiting PsiB <- Psi"""
    for a in range(p.NA):
        for b, c in p.bdy[a]:
            p.PsiB[a, b - 1] = p.Psi[a, c - 1]


def _three_way(mesh, P, A, G, passes=(False, True), library=True, driver_like=False):
    _need_ref()
    p_ref = T.make_problem_3d(mesh, P, A, G, driver_like=driver_like)
    if p_ref.sched["totalCycles"] > 0:
        _seed_cycles(p_ref)
    p_orc = _clone_state(p_ref)
    ctx = None
    if library:
        ctx = T.gpu_context_3d(p_ref)
        if p_ref.sched["totalCycles"] > 0:
            ctx.init_radiation_field()
    sid = _stream_id()
    for save in passes:
        phi_ref, psi1_ref = _reference_pass(p_ref, save, sid)
        phi_orc, psi1_orc = _oracle_pass(p_orc, save)
        # (a) the oracle's restatement against the reference's own kernel
        assert T.relerr(phi_orc, phi_ref) <= TOL
        assert T.mixed_err(psi1_orc, psi1_ref, TOL) <= 1.0
        assert T.mixed_err(p_orc.PsiB, p_ref.PsiB, TOL) <= 1.0
        assert T.mixed_err(p_orc.cyclePsi, p_ref.cyclePsi, TOL) <= 1.0
        if save:
            assert T.mixed_err(p_orc.Psi, p_ref.Psi, TOL) <= 1.0
        # (b) the product against the reference's own kernel
        if ctx is not None:
            ctx.sweep(savePsi=save)
            assert T.relerr(ctx.download_phi(), phi_ref) <= TOL
            assert T.mixed_err(ctx.download_psib(), p_ref.PsiB, TOL) <= 1.0
            if save:
                assert T.mixed_err(ctx.download_psi(), p_ref.Psi, TOL) <= 1.0
    if ctx is not None:
        ctx.close()
    return p_ref


def test_tiled_mesh_random_state():
    _three_way(M.tiled_mesh((2, 2, 2)), 2, 2, 16)


def test_tiled_mesh_driver_problem_three_passes():
    """BASELINE configs[0] in small: -P 2 -A 2, mini-app opacities, non-final sweeps then the savePsi sweep"""
    _three_way(M.tiled_mesh((2, 2, 3)), 2, 2, 4, passes=(False, False, True), driver_like=True)


def test_box_mesh_more_angles_and_groups():
    _three_way(M.box_mesh((5, 4, 3)), 3, 2, 32)


def test_unstructured_box():
    _three_way(M.unstruct_box_mesh(2), 2, 2, 8)


def test_warped_mesh_with_cycle_list():
    p = _three_way(M.box_mesh((4, 4, 4), warp=0.35, seed=3), 2, 2, 4, passes=(False, False, True))
    assert p.sched["totalCycles"] > 0


def test_strongly_warped_mesh_direct_solve_branch():
    """zones with an intra-zone cycle (nextZ < 0) take the reference's "direct solve" branch (:479-496), which reads the
    previous content of Psi1: compared angle by angle with identical Psi1 input (oracle against reference only)"""
    p = _three_way(M.box_mesh((7, 6, 5), warp=0.9, seed=2), 2, 2, 4, passes=(False, True), library=False)
    assert (p.sched["nextZ"] < 0).sum() > 0
