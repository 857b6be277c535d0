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
    if not O.ref_cuda_available():
        pytest.skip("oracle/_ref/libgpu_sweepucbxyz_ref.so is missing: run `make -C oracle` where /root/reference and nvcc "
                    "exist (__graft_entry__.build() does); the file travels to the GPU box with the snapshot")


def _reference_pass(p, savePsi, sid, lib=None):
    """One flux pass angle by angle through the reference's CUDA shim (or `lib`'s symbol of the same name); mutates p.Psi
    (if savePsi), p.PsiB, p.cyclePsi the way SetSweep_CUDA.F90 does and returns (PhiTotal, Psi1 of every angle)."""
    m = p.mesh
    nc, nb, G = m.ncornr, m.nbelem, p.G
    phi_total = np.zeros((nc, G))
    psi1_all = np.zeros((p.NA, nc, G))
    for a in range(p.NA):
        Psi1 = np.zeros((nc + nb, G))
        Phi = np.zeros((nc, G))
        O.ref_cuda_sweep_xyz(p.om, p.geom, p.sched, a, p.omega, p.weight, p.tau, p.STotal, p.Sigt, p.Psi[a], Psi1, p.PsiB[a],
                             Phi, p.cyclePsi, savePsi, stream_id=sid, lib=lib)
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


def test_baseline_config0_size():
    """BASELINE configs[0] at full size: tiled mesh -d 10,10,10 (24 000 zones, 192 000 corners), -P 2 -A 2 (32 angles), the driver's
    problem data; G = 4 instead of 2 because the reference kernel handles groups in blocks of 4 without a bound check (:32, :207)"""
    _three_way(M.tiled_mesh((10, 10, 10)), 2, 2, 4, passes=(False, True), driver_like=True)


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


# ---------------------------------------------------------------------------
# the product's back-compat seam: libumtsweep.so exports gpu_sweepucbxyz with the reference's signature
# (include/teton_gpu_compat.h), so the same caller drives both libraries
# ---------------------------------------------------------------------------
def _compat_vs_reference(mesh, P, A, G, passes, driver_like=False):
    _need_ref()
    from umt_b200 import teton
    lib = teton.load_library()
    p_ref = T.make_problem_3d(mesh, P, A, G, driver_like=driver_like)
    if p_ref.sched["totalCycles"] > 0:
        _seed_cycles(p_ref)
    p_lib = _clone_state(p_ref)
    sid_ref, sid_lib = _stream_id(), _stream_id()
    for save in passes:
        phi_ref, psi1_ref = _reference_pass(p_ref, save, sid_ref)
        phi_lib, psi1_lib = _reference_pass(p_lib, save, sid_lib, lib=lib)
        assert T.relerr(phi_lib, phi_ref) <= TOL
        assert T.mixed_err(psi1_lib, psi1_ref, TOL) <= 1.0
        assert T.mixed_err(p_lib.PsiB, p_ref.PsiB, TOL) <= 1.0
        assert T.mixed_err(p_lib.cyclePsi, p_ref.cyclePsi, TOL) <= 1.0
        assert T.mixed_err(p_lib.Psi, p_ref.Psi, TOL) <= 1.0
    return p_ref


def test_compat_symbol_tiled_mesh():
    _compat_vs_reference(M.tiled_mesh((2, 2, 2)), 2, 2, 16, passes=(False, True))


def test_compat_symbol_driver_problem():
    _compat_vs_reference(M.tiled_mesh((2, 2, 3)), 2, 2, 4, passes=(False, False, True), driver_like=True)


def test_compat_symbol_cycle_list_and_direct_solve_zones():
    p = _compat_vs_reference(M.box_mesh((4, 4, 4), warp=0.35, seed=3), 2, 2, 4, passes=(False, True))
    assert p.sched["totalCycles"] > 0
    p = _compat_vs_reference(M.box_mesh((7, 6, 5), warp=0.9, seed=2), 2, 2, 4, passes=(False, True))
    assert (p.sched["nextZ"] < 0).sum() > 0


def test_compat_symbol_reflecting_boundary_rows():
    """the one-reflecting-boundary arguments (b0, nBdyElem, PsiBMref): both shims overwrite the incident rows
    [b0, b0+nBdyElem) of PsiB with the mirror angle's rows before the sweep (GPU_SweepUCBxyz.cu:796-807)"""
    _need_ref()
    import ctypes as C
    from umt_b200 import teton
    m = M.box_mesh((3, 3, 3))
    p = T.make_problem_3d(m, 1, 1, 8)
    nc, nb, G = m.ncornr, m.nbelem, p.G
    # first boundary of the mesh (b0 = 0): for b0 > 0 the reference offsets its copy by b0 doubles instead of b0 rows
    # (`d_PsiB + *b0`, `PsiBMref + *b0`, :802-803), which this library does not imitate
    b0, n = 0, m.boundaries[0].n_elem
    out = []
    for lib in (None, teton.load_library()):
        L = lib if lib is not None else O.ref_cuda_lib()
        a, mref = 0, p.NA - 1
        Psi1 = np.zeros((nc + nb, G)); Phi = np.zeros((nc, G))
        PsiB = p.PsiB[a].copy(); PsiBM = p.PsiB[mref].copy(); Psi = p.Psi[a].copy(); cyc = p.cyclePsi.copy()
        i = lambda v: C.byref(C.c_int(int(v)))
        d = lambda v: C.byref(C.c_double(float(v)))
        s, k = p.sched, p.om.keep
        om_a = np.ascontiguousarray(p.omega[a])
        sid = _stream_id()
        L.gpu_sweepucbxyz.restype = None
        L.gpu_sweepucbxyz(
            i(a + 1), i(s["nHyperPlanes"][a]), O._ip(s["zonesInPlane"][a]), O._ip(s["nextZ"][a]), O._ip(s["nextC"][a]),
            O._dp(p.STotal), d(p.tau), O._dp(Psi), i(G), O._dp(p.geom["Volume"]), O._dp(p.Sigt), O._ip(k["nCFaces"]),
            i(3), i(3), i(nc), O._dp(p.geom["A_fp"]), O._dp(om_a), O._ip(k["cFP"]), O._dp(Psi1), i(nb),
            O._dp(p.geom["A_ez"]), O._ip(k["cEZ"]), i(p.NA), d(p.weight[a]), O._dp(Phi), O._dp(PsiB), i(m.maxCorner),
            i(0), i(sid), i(1), i(0), i(0), i(0), O._dp(cyc), O._ip(s["cycleList"]), i(b0), i(n), O._dp(PsiBM), i(mref + 1),
            O._ip(k["numCorner"]), O._ip(k["cOffSet"]))
        L.gpu_streamsynchronize.restype = None
        L.gpu_streamsynchronize(i(sid))
        out.append((Psi1[:nc].copy(), PsiB, Phi))
    (psi1_r, psib_r, phi_r), (psi1_l, psib_l, phi_l) = out
    assert T.relerr(phi_l, phi_r) <= TOL
    assert T.mixed_err(psi1_l, psi1_r, TOL) <= 1.0
    assert T.mixed_err(psib_l, psib_r, TOL) <= 1.0
