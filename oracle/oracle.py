"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE — see umt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.  It never touches the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_NBB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_bp = C.POINTER(C.c_ubyte)


def build(force: bool = False) -> None:
    so = os.path.join(_HERE, "libumt_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("umt_oracle.c", "umt_oracle_gta.c")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale or not os.path.exists(os.path.join(_HERE, "_ref", "libnbb_ref.so")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


class _Mesh(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("ndim", "nzones", "ncornr", "nbelem", "maxcf", "maxCorner", "maxFaces")] + \
               [(n, c_ip) for n in ("numCorner", "cOffSet", "zoneFaces", "zoneOpp", "faceOpp", "nCFaces", "cFP", "cEZ", "CToFace")] + \
               [("BoundaryZone", c_bp), ("px", c_dp)]


def use_fast_build() -> bool:
    """bench.py's CPU legs: switch to the -O3 -march=native build of the same sources (built here, on the machine that runs
    it).  Returns False (and keeps the strict build) if it cannot be built.  Must be called before the first lib()."""
    global _LIB
    try:
        # -B: always rebuilt on the machine that runs it (-march=native: a binary shipped from another host may not run here)
        subprocess.check_call(["make", "-B", "-C", _HERE, "fast"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _LIB = None
        os.environ["UMT_ORACLE_LIB"] = os.path.join(_HERE, "_ref", "libumt_oracle_fast.so")
        return True
    except Exception:
        return False


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.environ.get("UMT_ORACLE_LIB") or os.path.join(_HERE, "libumt_oracle.so"))
        _LIB.orc_quad_xyz.restype = C.c_int
        _LIB.orc_quad_rz.restype = C.c_int
        _LIB.orc_snnext.restype = C.c_int
        _LIB.orc_bdy_exit.restype = C.c_int
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _bp(a):
    return a.ctypes.data_as(c_bp)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class OMesh:
    """Keeps the numpy arrays alive next to the C struct."""

    def __init__(self, m):
        self.m = m
        self.keep = dict(
            numCorner=_i32(m.numCorner), cOffSet=_i32(m.cOffSet), zoneFaces=_i32(m.zoneFaces),
            zoneOpp=_i32(m.zoneOpp), faceOpp=_i32(m.faceOpp), nCFaces=_i32(m.nCFacesArray),
            cFP=_i32(m.cFP), cEZ=_i32(m.cEZ), CToFace=_i32(m.CToFace),
            BoundaryZone=np.ascontiguousarray(m.BoundaryZone, dtype=np.uint8),
            px=np.ascontiguousarray(m.px, dtype=np.float64))
        s = _Mesh()
        for n in ("ndim", "nzones", "ncornr", "nbelem", "maxcf", "maxCorner", "maxFaces"):
            setattr(s, n, int(getattr(m, n)))
        for n in ("numCorner", "cOffSet", "zoneFaces", "zoneOpp", "faceOpp", "nCFaces", "cFP", "cEZ", "CToFace"):
            setattr(s, n, _ip(self.keep[n]))
        s.BoundaryZone = _bp(self.keep["BoundaryZone"])
        s.px = _dp(self.keep["px"])
        self.s = s

    @property
    def ref(self):
        return C.byref(self.s)


# ---------------------------------------------------------------------------
def quad_xyz(npolar: int, nazimuthal: int, polaraxis: int = 1):
    NA = 8 * npolar * nazimuthal
    omega = np.zeros((NA, 3))
    weight = np.zeros(NA)
    n = lib().orc_quad_xyz(npolar, nazimuthal, polaraxis, _dp(omega), _dp(weight))
    assert n == NA
    return omega, weight


def quad_rz(npolar: int, nazimuthal: int) -> Dict[str, np.ndarray]:
    NA = 4 * npolar * (nazimuthal + 1)
    q = dict(omega=np.zeros((NA, 2)), weight=np.zeros(NA), start=np.zeros(NA, np.uint8), finish=np.zeros(NA + 1, np.uint8),
             level=np.zeros(NA, np.int32), alpha=np.zeros(NA), tau=np.zeros(NA), angDerivFac=np.zeros(NA),
             quadTauW1=np.zeros(NA), quadTauW2=np.zeros(NA))
    n = lib().orc_quad_rz(npolar, nazimuthal, _dp(q["omega"]), _dp(q["weight"]), _bp(q["start"]), _bp(q["finish"]),
                          _ip(q["level"]), _dp(q["alpha"]), _dp(q["tau"]), _dp(q["angDerivFac"]),
                          _dp(q["quadTauW1"]), _dp(q["quadTauW2"]))
    assert n == NA
    return q


def geometry(om: OMesh) -> Dict[str, np.ndarray]:
    m = om.m
    nd, nc, mcf = m.ndim, m.ncornr, m.maxcf
    g = dict(A_fp=np.zeros((nc, mcf, nd)), A_ez=np.zeros((nc, mcf, nd)), Volume=np.zeros(nc),
             VolumeZone=np.zeros(m.nzones), A_bdy=np.zeros((max(m.nbelem, 1), nd)))
    if nd == 3:
        lib().orc_geometry_xyz(om.ref, _dp(g["A_fp"]), _dp(g["A_ez"]), _dp(g["Volume"]))
        vol2 = np.zeros(nc)
        lib().orc_volume_xyz(om.ref, _dp(vol2), _dp(g["VolumeZone"]), _dp(g["A_bdy"]))
        g["Volume_getvolume"] = vol2
    else:
        g.update(Area=np.zeros(nc), RadiusFP=np.zeros((nc, 2)), RadiusEZ=np.zeros((nc, 2)), RadiusB=np.zeros(max(m.nbelem, 1)))
        lib().orc_geometry_rz(om.ref, _dp(g["A_fp"]), _dp(g["A_ez"]), _dp(g["Area"]), _dp(g["Volume"]),
                              _dp(g["RadiusFP"]), _dp(g["RadiusEZ"]), _dp(g["VolumeZone"]), _dp(g["A_bdy"]), _dp(g["RadiusB"]))
    return g


def schedule(om: OMesh, geom, omega: np.ndarray, skip: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """rtorder.F90: snnext for every angle (finishing directions are skipped)."""
    m = om.m
    NA = omega.shape[0]
    nz, nc = m.nzones, m.ncornr
    s = dict(nHyperPlanes=np.zeros(NA, np.int32), zonesInPlane=np.zeros((NA, nz), np.int32),
             nextZ=np.zeros((NA, nz), np.int32), nextC=np.zeros((NA, nc), np.int32),
             numCycles=np.zeros(NA, np.int32), cycleOffSet=np.zeros(NA, np.int32))
    cl: List[np.ndarray] = []
    tmp = np.zeros(nc, np.int32)
    off = 0
    om_c = np.ascontiguousarray(omega)
    for a in range(NA):
        s["cycleOffSet"][a] = off
        if skip is not None and skip[a]:
            continue
        ncyc = C.c_int(0)
        s["nHyperPlanes"][a] = lib().orc_snnext(om.ref, _dp(geom["A_fp"]), _dp(geom["A_ez"]), _dp(om_c[a]),
                                                _ip(s["nextZ"][a]), _ip(s["nextC"][a]), _ip(s["zonesInPlane"][a]),
                                                C.byref(ncyc), _ip(tmp))
        s["numCycles"][a] = ncyc.value
        cl.append(tmp[:ncyc.value].copy())
        off += ncyc.value
    s["cycleList"] = np.concatenate(cl + [np.zeros(1, np.int32)]).astype(np.int32)
    s["totalCycles"] = off
    return s


def bdy_exit(om: OMesh, geom, omega_a: np.ndarray, is_shared=None, is_exit_shared=None) -> np.ndarray:
    m = om.m
    out = np.zeros((max(m.nbelem, 1), 2), np.int32)
    n = lib().orc_bdy_exit(m.ndim, m.nbelem, _dp(geom["A_bdy"]), _ip(_i32(m.BdyToC)),
                           _bp(is_shared) if is_shared is not None else None,
                           _bp(is_exit_shared) if is_exit_shared is not None else None,
                           _dp(np.ascontiguousarray(omega_a)), _ip(out))
    return out[:n]


def sweep_xyz(om, geom, sched, a, omega, weight, tau, STotal, Sigt, PsiA, Psi1, PsiBA, Phi, savePsi):
    G = STotal.shape[-1]
    lib().orc_sweep_xyz(om.ref, G, int(sched["nHyperPlanes"][a]), _ip(sched["zonesInPlane"][a]), _ip(sched["nextZ"][a]),
                        _ip(sched["nextC"][a]), _dp(np.ascontiguousarray(omega[a])), C.c_double(weight[a]), C.c_double(tau),
                        _dp(STotal), _dp(Sigt), _dp(geom["Volume"]), _dp(geom["A_fp"]), _dp(geom["A_ez"]),
                        _dp(PsiA), _dp(Psi1), _dp(PsiBA), _dp(Phi), int(bool(savePsi)))


def sweep_rz(om, geom, sched, a, q, tau, STotal, Sigt, Psi, Psi1, PsiM, PsiB, Phi, bdyList, savePsi):
    """One SweepUCBrz call for angle index a (0-based) of quadrature dict q."""
    G = STotal.shape[-1]
    NA = q["omega"].shape[0]
    setfin = bool(q["finish"][a + 1]) if a + 1 < NA + 1 else False
    nxt = a + 1 if a + 1 < NA else a
    lib().orc_sweep_rz(om.ref, G, int(sched["nHyperPlanes"][a]), _ip(sched["zonesInPlane"][a]), _ip(sched["nextZ"][a]),
                       _ip(sched["nextC"][a]), _dp(np.ascontiguousarray(q["omega"][a])), C.c_double(q["weight"][a]),
                       C.c_double(tau), C.c_double(q["angDerivFac"][a]), C.c_double(q["quadTauW1"][a]),
                       C.c_double(q["quadTauW2"][a]), int(bool(q["start"][a])), int(setfin),
                       _dp(STotal), _dp(Sigt), _dp(geom["Volume"]), _dp(geom["Area"]), _dp(geom["A_fp"]), _dp(geom["A_ez"]),
                       _dp(geom["RadiusFP"]), _dp(geom["RadiusEZ"]), int(len(bdyList)), _ip(np.ascontiguousarray(bdyList, np.int32)),
                       _dp(Psi[a]), _dp(Psi[nxt]), _dp(Psi1), _dp(PsiM), _dp(PsiB[a]), _dp(PsiB[nxt]), _dp(Phi), int(bool(savePsi)))


def setsweep_xyz(om, geom, sched, omega, weight, tau, STotal, Sigt, Psi, PsiB, cyclePsi, savePsi, nthreads=0):
    """SetSweep.F90 single flux pass + getPhiTotal, one angle per set, OpenMP over sets."""
    m = om.m
    NA = omega.shape[0]
    G = STotal.shape[-1]
    Phi = np.zeros((m.ncornr, G))
    if cyclePsi is None:
        cyclePsi = np.zeros((max(int(sched["totalCycles"]), 1), G))
    lib().orc_setsweep_xyz(om.ref, G, NA, _ip(sched["nHyperPlanes"]), _ip(sched["zonesInPlane"]), _ip(sched["nextZ"]),
                           _ip(sched["nextC"]), _ip(sched["numCycles"]), _ip(sched["cycleOffSet"]), _ip(sched["cycleList"]),
                           _dp(cyclePsi), _dp(np.ascontiguousarray(omega)), _dp(np.ascontiguousarray(weight)), C.c_double(tau),
                           _dp(STotal), _dp(Sigt), _dp(geom["Volume"]), _dp(geom["A_fp"]), _dp(geom["A_ez"]),
                           _dp(Psi), _dp(PsiB), _dp(Phi), int(bool(savePsi)), int(nthreads))
    return Phi


def max_threads() -> int:
    return lib().orc_max_threads()


# ---------------------------------------------------------------------------
# reference's own Planck integrator, compiled from /root/reference (oracle/_ref)
# ---------------------------------------------------------------------------
def planck_groups_ref(T: float, bounds: np.ndarray, k: float = 1.0, Bnorm: float = 1.0) -> np.ndarray:
    global _NBB
    if _NBB is None:
        build()
        _NBB = C.CDLL(os.path.join(_HERE, "_ref", "libnbb_ref.so"))
    ng = len(bounds) - 1
    B = np.zeros(ng)
    b = np.ascontiguousarray(bounds, dtype=np.float64)
    _NBB.NBB_integrateBlackBodyGroups(C.c_double(T), C.c_double(k), C.c_double(Bnorm), C.c_int(ng), _dp(b), _dp(B))
    return B


# ---------------------------------------------------------------------------
# grey transport acceleration (umt_oracle_gta.c), 3-D, "new" GTA solver
# ---------------------------------------------------------------------------
class _Gta(C.Structure):
    _fields_ = [("M", C.POINTER(_Mesh)), ("nAng", C.c_int), ("nHyperPlanes", c_ip), ("zonesInPlane", c_ip), ("nextZ", c_ip),
                ("nextC", c_ip), ("omega", c_dp), ("weight", c_dp), ("Volume", c_dp), ("A_fp", c_dp), ("A_ez", c_dp),
                ("GreySigTotal", c_dp), ("GreySigtInv", c_dp), ("GreySigScat", c_dp), ("GreySigScatVol", c_dp),
                ("GreySource", c_dp), ("TT", c_dp), ("wtiso", C.c_double),
                ("Area", c_dp), ("RadiusFP", c_dp), ("RadiusEZ", c_dp), ("angDerivFac", c_dp), ("quadTauW1", c_dp), ("quadTauW2", c_dp),
                ("start", c_bp), ("finish", c_bp), ("nStagesR", C.c_int), ("nReflOps", C.c_int), ("angleStage", c_ip), ("reflOps", c_ip)]


def gta_quad_xyz():
    omega, weight = np.zeros((8, 3)), np.zeros(8)
    lib().orc_gta_quad_xyz(_dp(omega), _dp(weight))
    return omega, weight


def gta_quad_rz() -> Dict[str, np.ndarray]:
    """the GTA angle set in r-z (level-symmetric S2: two xi-levels of start, mu<0, mu>0, finish) with its angular-derivative coefficients"""
    NA = 8
    q = dict(omega=np.zeros((NA, 2)), weight=np.zeros(NA), start=np.zeros(NA, np.uint8), finish=np.zeros(NA + 1, np.uint8),
             level=np.zeros(NA, np.int32), angDerivFac=np.zeros(NA), quadTauW1=np.zeros(NA), quadTauW2=np.zeros(NA))
    n = lib().orc_gta_quad_rz(_dp(q["omega"]), _dp(q["weight"]), _bp(q["start"]), _bp(q["finish"]), _ip(q["level"]),
                              _dp(q["angDerivFac"]), _dp(q["quadTauW1"]), _dp(q["quadTauW2"]))
    assert n == NA
    return q


def gta_set_opacity(om: OMesh, geom, tau, Siga, Sigs, Eta, Chi):
    """setGTAOpacityNEW; Chi (nc,ngr) is rescaled in place.  Returns dict of grey opacities."""
    nc = om.m.ncornr
    ngr = Siga.shape[-1]
    o = dict(GreySigTotal=np.zeros(nc), GreySigScat=np.zeros(nc), GreySigScatVol=np.zeros(nc), GreySigtInv=np.zeros(nc))
    lib().orc_gta_set_opacity(om.ref, ngr, C.c_double(tau), _dp(Siga), _dp(Sigs), _dp(Eta), _dp(Chi), _dp(geom["Volume"]),
                              _dp(o["GreySigTotal"]), _dp(o["GreySigScat"]), _dp(o["GreySigScatVol"]), _dp(o["GreySigtInv"]))
    return o


def collision_rate(om: OMesh, Eta, Siga, Sigs, PhiTotal, GreySource, residualFlag):
    lib().orc_collision_rate(om.ref, Siga.shape[-1], _dp(Eta), _dp(Siga), _dp(Sigs), _dp(PhiTotal), _dp(GreySource), int(residualFlag))
    return GreySource


class GtaProblem:
    """Keeps every array the C struct points at alive."""

    def __init__(self, om: OMesh, geom, sched, omega, weight, opac, GreySource, wtiso, q=None, reflect=None):
        """q: the r-z angle-set dict of gta_quad_rz() (2-D meshes); reflect = (angleStage (nAng), ops (n, 5) int32 rows of
        (stage, minc, mref, first element, count), all 0-based) for reflecting boundaries (3-D)"""
        self.om, self.geom, self.sched, self.q = om, geom, sched, q
        self.omega = np.ascontiguousarray(omega)
        self.weight = np.ascontiguousarray(weight)
        self.opac = {k: np.ascontiguousarray(v) for k, v in opac.items()}
        self.GreySource = np.ascontiguousarray(GreySource, dtype=np.float64).copy()
        self.TT = np.zeros((om.m.ncornr, om.m.maxCorner))
        s = _Gta()
        s.M = C.pointer(om.s)
        s.nAng = len(weight)
        s.nHyperPlanes, s.zonesInPlane, s.nextZ, s.nextC = (_ip(sched[k]) for k in ("nHyperPlanes", "zonesInPlane", "nextZ", "nextC"))
        s.omega, s.weight = _dp(self.omega), _dp(self.weight)
        s.Volume, s.A_fp, s.A_ez = _dp(geom["Volume"]), _dp(geom["A_fp"]), _dp(geom["A_ez"])
        for k in ("GreySigTotal", "GreySigtInv", "GreySigScat", "GreySigScatVol"):
            setattr(s, k, _dp(self.opac[k]))
        s.GreySource, s.TT, s.wtiso = _dp(self.GreySource), _dp(self.TT), wtiso
        if q is not None:
            s.Area, s.RadiusFP, s.RadiusEZ = _dp(geom["Area"]), _dp(geom["RadiusFP"]), _dp(geom["RadiusEZ"])
            s.angDerivFac, s.quadTauW1, s.quadTauW2 = _dp(q["angDerivFac"]), _dp(q["quadTauW1"]), _dp(q["quadTauW2"])
            s.start, s.finish = _bp(q["start"]), _bp(q["finish"])
        if reflect is not None and len(reflect[1]):
            self._stage = np.ascontiguousarray(reflect[0], np.int32)
            self._ops = np.ascontiguousarray(reflect[1], np.int32)
            s.nStagesR, s.nReflOps = int(self._stage.max()) + 1, len(self._ops)
            s.angleStage, s.reflOps = _ip(self._stage), _ip(self._ops)
        self.s = s

    def init_tt(self):
        o = self.opac
        if self.q is not None:
            q, g = self.q, self.geom
            lib().orc_gta_init_tt_rz(self.om.ref, len(self.weight), _ip(self.sched["nextC"]), _dp(self.omega), _dp(self.weight), _bp(q["start"]),
                                     _bp(q["finish"]), _dp(q["angDerivFac"]), _dp(q["quadTauW1"]), _dp(q["quadTauW2"]), _dp(g["Volume"]),
                                     _dp(g["Area"]), _dp(g["A_fp"]), _dp(g["A_ez"]), _dp(g["RadiusFP"]), _dp(g["RadiusEZ"]),
                                     _dp(o["GreySigTotal"]), _dp(self.TT))
            return self.TT
        lib().orc_gta_init_tt(self.om.ref, len(self.weight), _dp(self.omega), _dp(self.weight), _dp(self.geom["Volume"]),
                              _dp(self.geom["A_fp"]), _dp(self.geom["A_ez"]), _dp(o["GreySigTotal"]), _dp(self.TT))
        return self.TT

    def grey_sweep(self, PsiB, P, withSource):
        """GreySweepNEW: P (nc) and PsiB (nAng, nb) in/out."""
        lib().orc_gta_grey_sweep(C.byref(self.s), _dp(PsiB), _dp(P), int(bool(withSource)))

    def sweep_angle(self, a, TsaSource, PsiBa, PhiInc):
        m = self.om.m
        tPsi, pInc = np.zeros(m.ncornr + m.nbelem), np.zeros(m.ncornr)
        s, o = self.sched, self.opac
        lib().orc_gta_sweep_angle(self.om.ref, int(s["nHyperPlanes"][a]), _ip(s["zonesInPlane"][a]), _ip(s["nextZ"][a]), _ip(s["nextC"][a]),
                                  _dp(self.omega[a]), C.c_double(self.weight[a]), _dp(self.geom["Volume"]), _dp(self.geom["A_fp"]),
                                  _dp(self.geom["A_ez"]), _dp(o["GreySigTotal"]), _dp(o["GreySigtInv"]), _dp(TsaSource), _dp(tPsi), _dp(pInc),
                                  _dp(PsiBa), _dp(PhiInc))
        return tPsi, pInc

    def sweep_angle_rz(self, a, TsaSource, PsiBa, PhiInc, tPsiM, tInc):
        """SweepGreyUCBrz for angle a (not a finishing direction); tPsiM / tInc (nc) carry the half-angle values along the xi-level"""
        m = self.om.m
        tPsi, pInc = np.zeros(m.ncornr + m.nbelem), np.zeros(m.ncornr)
        s, o, q, g = self.sched, self.opac, self.q, self.geom
        lib().orc_gta_sweep_angle_rz(self.om.ref, int(s["nHyperPlanes"][a]), _ip(s["zonesInPlane"][a]), _ip(s["nextZ"][a]), _ip(s["nextC"][a]),
                                     _dp(self.omega[a]), C.c_double(self.weight[a]), C.c_double(q["angDerivFac"][a]), C.c_double(q["quadTauW1"][a]),
                                     C.c_double(q["quadTauW2"][a]), int(q["start"][a]), _dp(g["Volume"]), _dp(g["Area"]), _dp(g["A_fp"]),
                                     _dp(g["A_ez"]), _dp(g["RadiusFP"]), _dp(g["RadiusEZ"]), _dp(o["GreySigTotal"]), _dp(o["GreySigtInv"]),
                                     _dp(TsaSource), _dp(tPsi), _dp(pInc), _dp(tPsiM), _dp(tInc), _dp(PsiBa), _dp(PhiInc))
        return tPsi, pInc

    def solve(self, PhiTotal, epsPoint=1e-6, maxIters=21, epsGrey=0.1, enforceHardMax=False):
        """GTASolver; returns (GreyCorrection, nGreyIter, maxRelErrGrey).  Consumes GreySource and TT."""
        nc = self.om.m.ncornr
        corr = np.zeros(nc)
        err = C.c_double(0.0)
        lib().orc_gta_solver.restype = C.c_int
        n = lib().orc_gta_solver(C.byref(self.s), PhiTotal.shape[-1], _dp(PhiTotal), _dp(self.geom["VolumeZone"]), C.c_double(epsPoint),
                                 int(maxIters), C.c_double(epsGrey), int(bool(enforceHardMax)), _dp(corr), C.byref(err))
        return corr, n, err.value


def add_grey_corrections(GreyCorrection, Chi, PhiTotal):
    nc, ngr = PhiTotal.shape
    lib().orc_add_grey_corrections(ngr, nc, _dp(GreyCorrection), _dp(Chi), _dp(PhiTotal))
    return PhiTotal


# ---------------------------------------------------------------------------
# The reference's own CUDA sweep (gpu/GPU_SweepUCBxyz.cu, compiled unmodified into _ref/libgpu_sweepucbxyz_ref.so by the
# Makefile).  Needs a GPU, so only the -m gpu tests and bench.py's reference_cuda leg call it.
_REFCUDA = None


def ref_cuda_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libgpu_sweepucbxyz_ref.so"))


def ref_cuda_lib():
    global _REFCUDA
    if _REFCUDA is None:
        _REFCUDA = C.CDLL(os.path.join(_HERE, "_ref", "libgpu_sweepucbxyz_ref.so"))
        _REFCUDA.gpu_sweepucbxyz.restype = None
        _REFCUDA.gpu_streamsynchronize.restype = None
    return _REFCUDA


def ref_cuda_sweep_xyz(om, geom, sched, a, omega, weight, tau, STotal, Sigt, PsiA, Psi1, PsiBA, Phi, cyclePsi, savePsi,
                       stream_id=0, sync=True, lib=None):
    """One call of the reference's `gpu_sweepucbxyz` (GPU_SweepUCBxyz.cu:532-572, argument order of the Fortran interface
    SweepUCBxyzToGPU.F90:30-91) for angle index a (0-based): every argument by reference, host arrays in Teton's layout.
    Like the Fortran caller (SetSweep_CUDA.F90) the whole cycleList/cyclePsi are passed with this angle's offset and count;
    the shim runs initFromCycleList, Q = STotal + tau*Psi, the sweep, Phi += quadwt*Psi1 and updateCycleList on the device
    and copies PsiB, Phi, Psi1, cyclePsi (and Psi when savePsi) back.  No reflecting boundaries (nBdyElem = 0).
    The static device buffers of a stream id are sized by its first call: use a new stream_id (< 80) for a new problem size.
    `lib`: another library exporting the same three symbols (the product's back-compat seam, include/teton_gpu_compat.h)."""
    m = om.m
    G = STotal.shape[-1]
    assert m.ndim == 3 and m.maxcf == 3, "the reference kernel indexes omega.A with ndim where maxcf is meant (:327, :360)"
    assert G % 4 == 0, "GROUPS_IN_BLOCK = 4 and no group bound check in the reference kernel (:207)"
    assert (m.ncornr * G) % 2 == 0
    for arr in (STotal, Sigt, PsiA, Psi1, PsiBA, Phi, cyclePsi):
        assert arr.flags["C_CONTIGUOUS"] and arr.dtype == np.float64
    i = lambda v: C.byref(C.c_int(int(v)))
    d = lambda v: C.byref(C.c_double(float(v)))
    nhp = int(sched["nHyperPlanes"][a])
    om_a = np.ascontiguousarray(omega[a], dtype=np.float64)
    keep = om.keep
    L = lib if lib is not None else ref_cuda_lib()
    L.gpu_sweepucbxyz.restype = None
    L.gpu_streamsynchronize.restype = None
    L.gpu_sweepucbxyz(
        i(a + 1), i(nhp), _ip(sched["zonesInPlane"][a]), _ip(sched["nextZ"][a]), _ip(sched["nextC"][a]),
        _dp(STotal), d(tau), _dp(PsiA), i(G), _dp(geom["Volume"]), _dp(Sigt), _ip(keep["nCFaces"]),
        i(m.ndim), i(m.maxcf), i(m.ncornr), _dp(geom["A_fp"]), _dp(om_a), _ip(keep["cFP"]), _dp(Psi1), i(m.nbelem),
        _dp(geom["A_ez"]), _ip(keep["cEZ"]), i(omega.shape[0]), d(weight[a]), _dp(Phi), _dp(PsiBA), i(m.maxCorner),
        i(0), i(stream_id), i(1), i(1 if savePsi else 0), i(sched["numCycles"][a]), i(sched["cycleOffSet"][a]),
        _dp(cyclePsi), _ip(sched["cycleList"]), i(0), i(0), _dp(PsiBA), i(0), _ip(keep["numCorner"]), _ip(keep["cOffSet"]))
    if sync:
        L.gpu_streamsynchronize(i(stream_id))
