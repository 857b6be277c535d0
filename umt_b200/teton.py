"""Host-side mirror of the C ABI in include/umt_sweep.h, for tests and bench.

It plays the role of Teton's Fortran caller (rt/ControlSweep.F90:55-73 ->
SetSweep_CUDA): it owns nothing but a context handle and passes flat arrays in
Teton's layout.  There is no CPU fallback here: if libumtsweep.so is missing or no
GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UMT_LIB") or os.path.join(_HERE, "libumtsweep.so")   # UMT_LIB: A/B builds of the same library
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_bp = C.POINTER(C.c_ubyte)


class UmtError(RuntimeError):
    pass


def load_library():
    """dlopen libumtsweep.so (built in-tree by __graft_entry__.build / make)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UmtError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        _lib.umt_last_error.restype = C.c_char_p
        _lib.umt_last_error.argtypes = [C.c_void_p]
        _lib.umt_version.restype = C.c_char_p
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def _bp(a):
    return None if a is None else a.ctypes.data_as(c_bp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


class SweepContext:
    """One mesh domain resident on one B200."""

    def __init__(self, ndim, nzones, ncornr, nbelem, maxcf, maxCorner, ngr, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        self.ndim, self.nz, self.nc, self.nb, self.maxcf, self.maxCorner, self.G = ndim, nzones, ncornr, nbelem, maxcf, maxCorner, ngr
        self.NA = 0
        rc = self.lib.umt_ctx_create(device, ndim, nzones, ncornr, nbelem, maxcf, maxCorner, ngr, C.byref(self.h))
        if rc:
            raise UmtError(f"umt_ctx_create -> {rc}: {self.lib.umt_last_error(None).decode()}")

    # -- plumbing ----------------------------------------------------------
    def _ck(self, rc, what):
        if rc:
            raise UmtError(f"{what} -> {rc}: {self.lib.umt_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            for ptr in getattr(self, "_host_blocks", []):
                self.lib.umt_host_free(self.h, ptr)
            self._host_blocks = []
            self.lib.umt_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_mesh(cls, mesh, ngr, device=0):
        ctx = cls(mesh.ndim, mesh.nzones, mesh.ncornr, mesh.nbelem, mesh.maxcf, mesh.maxCorner, ngr, device)
        ctx.set_connectivity(mesh)
        return ctx

    # -- setters -----------------------------------------------------------
    def set_connectivity(self, m):
        self._keep = [_i32(m.numCorner), _i32(m.cOffSet), _i32(m.nCFacesArray), _i32(m.cFP), _i32(m.cEZ),
                      _i32(m.zoneFaces), _i32(m.zoneOpp), _i32(m.faceOpp), _i32(m.CToFace), _u8(m.BoundaryZone), _i32(m.BdyToC)]
        k = self._keep
        self._ck(self.lib.umt_set_connectivity(self.h, _ip(k[0]), _ip(k[1]), _ip(k[2]), _ip(k[3]), _ip(k[4]), int(m.maxFaces),
                                               _ip(k[5]), _ip(k[6]), _ip(k[7]), _ip(k[8]), _bp(k[9]), _ip(k[10])), "umt_set_connectivity")

    def set_geometry(self, Volume, A_fp, A_ez, Area=None, RadiusFP=None, RadiusEZ=None, A_bdy=None):
        a = [_f64(x) for x in (Volume, A_fp, A_ez, Area, RadiusFP, RadiusEZ, A_bdy)]
        self._ck(self.lib.umt_set_geometry(self.h, *[_dp(x) for x in a]), "umt_set_geometry")

    def compute_geometry(self, px):
        self._ck(self.lib.umt_compute_geometry(self.h, _dp(_f64(px))), "umt_compute_geometry")

    def download_geometry(self):
        nd, nc, mcf, nb, nz = self.ndim, self.nc, self.maxcf, self.nb, self.nz
        g = dict(Volume=np.zeros(nc), A_fp=np.zeros((nc, mcf, nd)), A_ez=np.zeros((nc, mcf, nd)),
                 A_bdy=np.zeros((max(nb, 1), nd)), VolumeZone=np.zeros(nz))
        if nd == 2:
            g.update(Area=np.zeros(nc), RadiusFP=np.zeros((nc, 2)), RadiusEZ=np.zeros((nc, 2)))
        self._ck(self.lib.umt_download_geometry(self.h, _dp(g["Volume"]), _dp(g["A_fp"]), _dp(g["A_ez"]), _dp(g.get("Area")),
                                                _dp(g.get("RadiusFP")), _dp(g.get("RadiusEZ")), _dp(g["A_bdy"]), _dp(g["VolumeZone"])),
                 "umt_download_geometry")
        return g

    def set_quadrature(self, omega, weight, start=None, finish=None, angDerivFac=None, quadTauW1=None, quadTauW2=None):
        omega = _f64(omega)
        self.NA = omega.shape[0]
        self._ck(self.lib.umt_set_quadrature(self.h, self.NA, _dp(omega), _dp(_f64(weight)), _bp(_u8(start)), _bp(_u8(finish)),
                                             _dp(_f64(angDerivFac)), _dp(_f64(quadTauW1)), _dp(_f64(quadTauW2))), "umt_set_quadrature")

    def build_product_quadrature(self, npolar, nazimuthal, polaraxis=1):
        n = C.c_int(0)
        self._ck(self.lib.umt_build_product_quadrature(self.h, npolar, nazimuthal, polaraxis, C.byref(n)), "umt_build_product_quadrature")
        self.NA = n.value
        return self.NA

    def get_quadrature(self):
        om = np.zeros((self.NA, self.ndim))
        w = np.zeros(self.NA)
        self._ck(self.lib.umt_get_quadrature(self.h, _dp(om), _dp(w)), "umt_get_quadrature")
        return om, w

    def set_schedule(self, angle, nHyp, zonesInPlane, nextZ, nextC, cycleList=None, bdyList=None):
        cl = _i32(cycleList if cycleList is not None else np.zeros(0))
        bl = _i32(bdyList if bdyList is not None else np.zeros((0, 2)))
        self._ck(self.lib.umt_set_schedule(self.h, int(angle), int(nHyp), _ip(_i32(zonesInPlane)), _ip(_i32(nextZ)), _ip(_i32(nextC)),
                                           int(len(cl)), _ip(cl), int(bl.size // 2), _ip(bl)), "umt_set_schedule")

    def build_schedule(self):
        self._ck(self.lib.umt_build_schedule(self.h), "umt_build_schedule")

    def schedule_info(self, angle):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.umt_get_schedule_info(self.h, int(angle), C.byref(a), C.byref(b), C.byref(c)), "umt_get_schedule_info")
        return a.value, b.value, c.value

    def get_schedule(self, angle):
        nh, ncyc, _ = self.schedule_info(angle)
        zip_ = np.zeros(max(nh, 1), np.int32)
        nz_ = np.zeros(self.nz, np.int32)
        ncn = np.zeros(self.nc, np.int32)
        cl = np.zeros(max(ncyc, 1), np.int32)
        self._ck(self.lib.umt_get_schedule(self.h, int(angle), _ip(zip_), _ip(nz_), _ip(ncn), _ip(cl)), "umt_get_schedule")
        return dict(nHyperPlanes=nh, zonesInPlane=zip_[:nh], nextZ=nz_, nextC=ncn, cycleList=cl[:ncyc])

    # -- state -------------------------------------------------------------
    def upload_state(self, Psi=None, PsiB=None, Sigt=None, STotal=None, tau=0.0):
        self._ck(self.lib.umt_upload_state(self.h, _dp(_f64(Psi)), _dp(_f64(PsiB)), _dp(_f64(Sigt)), _dp(_f64(STotal)), C.c_double(tau)),
                 "umt_upload_state")

    def upload_set(self, g0, Groups, angle0, NumAngles, Psi=None, PsiB=None):
        self._ck(self.lib.umt_upload_set(self.h, g0, Groups, angle0, NumAngles, _dp(_f64(Psi)), _dp(_f64(PsiB))), "umt_upload_set")

    def download_set(self, g0, Groups, angle0, NumAngles):
        Psi = np.zeros((NumAngles, self.nc, Groups))
        PsiB = np.zeros((NumAngles, max(self.nb, 1), Groups))
        self._ck(self.lib.umt_download_set(self.h, g0, Groups, angle0, NumAngles, _dp(Psi), _dp(PsiB)), "umt_download_set")
        return Psi, PsiB[:, :self.nb]

    def download_psi(self, out=None):
        out = np.zeros((self.NA, self.nc, self.G)) if out is None else out
        self._ck(self.lib.umt_download_psi(self.h, _dp(out)), "umt_download_psi")
        return out

    def download_psib(self):
        out = np.zeros((self.NA, max(self.nb, 1), self.G))
        if self.nb:
            self._ck(self.lib.umt_download_psib(self.h, _dp(out)), "umt_download_psib")
        return out[:, :self.nb]

    def download_phi(self, out=None):
        out = np.zeros((self.nc, self.G)) if out is None else out
        self._ck(self.lib.umt_download_phi(self.h, _dp(out)), "umt_download_phi")
        return out

    def init_teton(self, Trz, groupBounds, speedLight, radConstant, wtiso, efloor=0.0):
        self._ck(self.lib.umt_init_teton(self.h, _dp(_f64(Trz)), _dp(_f64(groupBounds)), C.c_double(speedLight), C.c_double(radConstant),
                                         C.c_double(wtiso), C.c_double(efloor)), "umt_init_teton")

    def set_boundary_sources(self):
        self._ck(self.lib.umt_set_boundary_sources(self.h), "umt_set_boundary_sources")

    def init_phi_total(self, volRatio=None):
        self._ck(self.lib.umt_init_phi_total(self.h, _dp(_f64(volRatio))), "umt_init_phi_total")

    def init_radiation_field(self):
        self._ck(self.lib.umt_init_radiation_field(self.h), "umt_init_radiation_field")

    def init_cycle_psi(self):
        self._ck(self.lib.umt_init_cycle_psi(self.h), "umt_init_cycle_psi")

    def set_psi1_ring(self, nBatches):
        self._ck(self.lib.umt_set_psi1_ring(self.h, int(nBatches)), "umt_set_psi1_ring")

    def psi_layout(self):
        info = np.zeros(6, np.int32)
        b = C.c_double(0.0)
        self._ck(self.lib.umt_get_psi_layout(self.h, _ip(info), C.byref(b)), "umt_get_psi_layout")
        return dict(single=bool(info[0]), psi1_slabs=int(info[1]), angle_batch=int(info[2]), ring_batches=int(info[3]),
                    batches=int(info[4]), angles_tallied_in_sweep=int(info[5]), bytes=b.value)

    def host_array(self, shape):
        """float64 array in page-locked host memory on the NUMA node of this context's GPU (umt_host_alloc); freed with the context"""
        n = int(np.prod(shape))
        ptr, node = C.c_void_p(), C.c_int(-1)
        self._ck(self.lib.umt_host_alloc(self.h, C.c_size_t(max(n, 1) * 8), C.byref(ptr), C.byref(node)), "umt_host_alloc")
        if not hasattr(self, "_host_blocks"):
            self._host_blocks = []
        self._host_blocks.append(ptr)
        self.host_numa_node = node.value
        buf = (C.c_double * max(n, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)

    # -- hot path ------------------------------------------------------------
    def sweep(self, savePsi=False, maxFluxIters=1, fluxTol=1e-6):
        it = C.c_int(0)
        self._ck(self.lib.umt_sweep(self.h, int(bool(savePsi)), int(maxFluxIters), C.c_double(fluxTol), C.byref(it)), "umt_sweep")
        return it.value

    def control_sweep(self, Sigt, STotal, tau, PhiTotal, savePsi=False, maxFluxIters=1, fluxTol=1e-6):
        """ControlSweep with host buffers: Sigt (nz, G), STotal (nc, G) in (None keeps the device copy), PhiTotal (nc, G) out (filled in place)"""
        it = C.c_int(0)
        assert PhiTotal.dtype == np.float64 and PhiTotal.flags.c_contiguous and PhiTotal.shape == (self.nc, self.G)
        self._ck(self.lib.umt_control_sweep(self.h, None if Sigt is None else _dp(_f64(Sigt)), None if STotal is None else _dp(_f64(STotal)),
                                            C.c_double(tau), int(bool(savePsi)), int(maxFluxIters), C.c_double(fluxTol), C.byref(it), _dp(PhiTotal)),
                 "umt_control_sweep")
        return it.value

    def last_times(self):
        t = np.zeros(4)
        self._ck(self.lib.umt_last_sweep_times(self.h, _dp(t)), "umt_last_sweep_times")
        return dict(sweep_ms=t[0], phi_ms=t[1], exchange_ms=t[2], total_ms=t[3])

    def last_launches(self):
        n = C.c_int(0)
        self._ck(self.lib.umt_last_sweep_launches(self.h, C.byref(n)), "umt_last_sweep_launches")
        return n.value

    def synchronize(self):
        self._ck(self.lib.umt_synchronize(self.h), "umt_synchronize")

    def cycle_edits(self, speedLight, radConstant, tr4floor, want_trz=False, want_density=False):
        out = np.zeros(5)
        trz = np.zeros(self.nz) if want_trz else None
        esc = np.zeros(self.G)
        dens = np.zeros((self.G, self.nz)) if want_density else None
        self._ck(self.lib.umt_cycle_edits(self.h, C.c_double(speedLight), C.c_double(radConstant), C.c_double(tr4floor), _dp(out), _dp(trz), _dp(esc), _dp(dens)),
                 "umt_cycle_edits")
        return dict(EnergyRadiation=out[0], TrMax=out[1], PowerEscape=out[2], PowerIncident=out[3], RadPowerEscape=esc, trz=trz,
                    RadEnergyDensity=None if dens is None else dens.T)

    def build_source(self, Siga, Sigs, Eta, Chi, EmissionRate=None):
        out = np.zeros((self.nc, self.G))
        self._ck(self.lib.umt_build_source(self.h, _dp(_f64(Siga)), _dp(_f64(Sigs)), _dp(_f64(Eta)), _dp(_f64(Chi)), _dp(_f64(EmissionRate)), _dp(out)),
                 "umt_build_source")
        return out

    # -- grey transport acceleration ------------------------------------------
    def gta_setup(self):
        self._ck(self.lib.umt_gta_setup(self.h), "umt_gta_setup")

    def gta_quadrature(self):
        om, w = np.zeros((8, self.ndim)), np.zeros(8)
        self._ck(self.lib.umt_gta_get_quadrature(self.h, _dp(om), _dp(w)), "umt_gta_get_quadrature")
        return om, w

    def gta_set_opacity(self, GreySigTotal, GreySigScat, GreySigScatVol):
        self._ck(self.lib.umt_gta_set_opacity(self.h, _dp(_f64(GreySigTotal)), _dp(_f64(GreySigScat)), _dp(_f64(GreySigScatVol))), "umt_gta_set_opacity")

    def gta_compute_opacity(self, Siga, Sigs, Eta, Chi):
        """Chi (nc, ngr) is rescaled in place (must be a contiguous float64 array)."""
        assert Chi.dtype == np.float64 and Chi.flags.c_contiguous
        self._ck(self.lib.umt_gta_compute_opacity(self.h, _dp(_f64(Siga)), _dp(_f64(Sigs)), _dp(_f64(Eta)), _dp(Chi)), "umt_gta_compute_opacity")
        o = dict(GreySigTotal=np.zeros(self.nc), GreySigScat=np.zeros(self.nc), GreySigScatVol=np.zeros(self.nc), GreySigtInv=np.zeros(self.nc))
        self._ck(self.lib.umt_gta_get_opacity(self.h, _dp(o["GreySigTotal"]), _dp(o["GreySigScat"]), _dp(o["GreySigScatVol"]), _dp(o["GreySigtInv"])),
                 "umt_gta_get_opacity")
        return o

    def collision_rate(self, Eta, Siga, Sigs, residualFlag=0):
        out = np.zeros(self.nc)
        self._ck(self.lib.umt_collision_rate(self.h, _dp(_f64(Eta)), _dp(_f64(Siga)), _dp(_f64(Sigs)), int(residualFlag), _dp(out)), "umt_collision_rate")
        return out

    def gta_set_source(self, GreySource):
        self._ck(self.lib.umt_gta_set_source(self.h, _dp(_f64(GreySource))), "umt_gta_set_source")

    def gta_init_tt(self):
        TT = np.zeros((self.nc, self.maxCorner))
        self._ck(self.lib.umt_gta_init_tt(self.h, _dp(TT)), "umt_gta_init_tt")
        return TT

    def gta_sweep(self, P, GreySource=None, PsiB=None, withSource=True):
        PhiInc = np.zeros(self.nc)
        PsiB = np.zeros((8, max(self.nb, 1))) if PsiB is None else PsiB
        self._ck(self.lib.umt_gta_sweep(self.h, _dp(_f64(P)), _dp(_f64(GreySource)), _dp(PsiB), _dp(PhiInc), int(bool(withSource))), "umt_gta_sweep")
        return PhiInc, PsiB

    def gta_grey_sweep(self, P, PsiB, withSource):
        """GreySweepNEW: P (nc) and PsiB (8, nb) are updated in place."""
        self._ck(self.lib.umt_gta_grey_sweep(self.h, _dp(P), _dp(PsiB), int(bool(withSource))), "umt_gta_grey_sweep")

    def gta_solve(self, epsPoint=1e-6, maxIters=21, epsGrey=0.1, enforceHardMax=False):
        n, e = C.c_int(0), C.c_double(0.0)
        self._ck(self.lib.umt_gta_solve(self.h, C.c_double(epsPoint), int(maxIters), C.c_double(epsGrey), int(bool(enforceHardMax)), C.byref(n), C.byref(e)),
                 "umt_gta_solve")
        corr = np.zeros(self.nc)
        self._ck(self.lib.umt_gta_get_correction(self.h, _dp(corr)), "umt_gta_get_correction")
        return corr, n.value, e.value

    def add_grey_corrections(self):
        self._ck(self.lib.umt_add_grey_corrections(self.h), "umt_add_grey_corrections")

    # -- reflecting boundaries ------------------------------------------------
    def add_reflecting_boundary(self, firstBdyElem, nBdyElem):
        self._ck(self.lib.umt_add_reflecting_boundary(self.h, int(firstBdyElem), int(nBdyElem)), "umt_add_reflecting_boundary")

    def reflected_angles(self, reflIndex):
        m = np.zeros(self.NA, np.int32)
        self._ck(self.lib.umt_get_reflected_angles(self.h, int(reflIndex), _ip(m)), "umt_get_reflected_angles")
        return m

    def reflect_stages(self):
        s = np.zeros(self.NA, np.int32)
        self._ck(self.lib.umt_get_reflect_stages(self.h, _ip(s)), "umt_get_reflect_stages")
        return s

    # -- domain decomposition -----------------------------------------------
    def add_shared_boundary(self, neighborRank, firstBdyElem, nBdyElem):
        self._ck(self.lib.umt_add_shared_boundary(self.h, int(neighborRank), int(firstBdyElem), int(nBdyElem)), "umt_add_shared_boundary")

    def set_comm(self, myRank, nRanks, id128: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(id128)
        self._ck(self.lib.umt_set_comm(self.h, int(myRank), int(nRanks), buf), "umt_set_comm")

    def set_rank(self, myRank, nRanks):
        self._ck(self.lib.umt_set_rank(self.h, int(myRank), int(nRanks)), "umt_set_rank")

    def get_incident_test(self, sharedIndex, nBdyElem):
        out = np.zeros((self.NA, nBdyElem), np.int8)
        self._ck(self.lib.umt_get_incident_test(self.h, int(sharedIndex), out.ctypes.data_as(C.c_void_p)), "umt_get_incident_test")
        return out

    def set_incident_test(self, sharedIndex, incTestNeighbor):
        a = np.ascontiguousarray(incTestNeighbor, dtype=np.int8)
        self._ck(self.lib.umt_set_incident_test(self.h, int(sharedIndex), a.ctypes.data_as(C.c_void_p)), "umt_set_incident_test")

    def build_exchange(self):
        self._ck(self.lib.umt_build_exchange(self.h), "umt_build_exchange")

    def exchange_counts(self, sharedIndex):
        ns, nr = np.zeros(self.NA, np.int32), np.zeros(self.NA, np.int32)
        self._ck(self.lib.umt_get_exchange_counts(self.h, int(sharedIndex), _ip(ns), _ip(nr)), "umt_get_exchange_counts")
        return ns, nr

    def exchange_lists(self, sharedIndex, angle):
        ns, nr = self.exchange_counts(sharedIndex)
        ls, lr = np.zeros(max(ns[angle - 1], 1), np.int32), np.zeros(max(nr[angle - 1], 1), np.int32)
        self._ck(self.lib.umt_get_exchange_lists(self.h, int(sharedIndex), int(angle), _ip(ls), _ip(lr)), "umt_get_exchange_lists")
        return ls[:ns[angle - 1]], lr[:nr[angle - 1]]

    def set_comm_sets(self, nCommSets):
        self._ck(self.lib.umt_set_comm_sets(self.h, int(nCommSets)), "umt_set_comm_sets")

    def sweep_scheduler(self, netFlux=None):
        """SweepScheduler for every comm set (collective); netFlux (nShared, NA) or None to tally it from the device PsiB"""
        nf = None if netFlux is None else _dp(_f64(netFlux))
        self._ck(self.lib.umt_sweep_scheduler(self.h, nf), "umt_sweep_scheduler")

    def net_flux(self, nShared):
        nf = np.zeros((nShared, self.NA))
        self._ck(self.lib.umt_get_net_flux(self.h, _dp(nf)), "umt_get_net_flux")
        return nf

    def angle_order(self, nShared=0):
        ao = np.zeros(self.NA, np.int32)
        ro = np.zeros((max(nShared, 1), self.NA), np.int32)
        self._ck(self.lib.umt_get_angle_order(self.h, _ip(ao), _ip(ro) if nShared else None), "umt_get_angle_order")
        return ao, ro[:nShared]

    def incident_flux(self, nBins=None):
        n = nBins if nBins is not None else self.NA
        a, b = np.zeros(n), np.zeros(n)
        self._ck(self.lib.umt_get_incident_flux(self.h, _dp(a), _dp(b)), "umt_get_incident_flux")
        return a, b

    def set_flux_floor(self, v):
        self._ck(self.lib.umt_set_flux_floor(self.h, C.c_double(v)), "umt_set_flux_floor")


def planck_groups(T, bounds, k=1.0, Bnorm=1.0):
    lib = load_library()
    b = _f64(bounds)
    B = np.zeros(len(b) - 1)
    rc = lib.umt_planck_groups(C.c_double(T), C.c_double(k), C.c_double(Bnorm), len(B), _dp(b), _dp(B))
    if rc:
        raise UmtError(f"umt_planck_groups -> {rc}")
    return B


def connect_local(contexts):
    """The contexts become ranks 0..n-1 of an in-process group (drive each from its own thread)."""
    lib = load_library()
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    rc = lib.umt_connect_local(arr, len(contexts))
    if rc:
        raise UmtError(f"umt_connect_local -> {rc}")


def control_sweep_sets(contexts, Sigt, STotal, tau, PhiTotal, savePsi=False, maxFluxIters=1, fluxTol=1e-6):
    """One ControlSweep over the group sets of a domain (umt_control_sweep_sets): contexts[k] holds group set k, Sigt[k] (nz, G_k) and
    STotal[k] (nc, G_k) in (None keeps the device copies), PhiTotal[k] (nc, G_k) out, filled in place.  Sets are pipelined: upload of
    set k+1 and download of set k-1 run under the sweep of set k."""
    lib = load_library()
    n = len(contexts)
    assert len(PhiTotal) == n
    dpp = C.POINTER(C.c_double) * n

    def ptrs(arrs, shape_of):
        if arrs is None:
            return None
        out = []
        for c, a in zip(contexts, arrs):
            if a is None:
                out.append(C.POINTER(C.c_double)())
            else:
                assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == shape_of(c)
                out.append(_dp(a))
        return dpp(*out)
    it = C.c_int(0)
    keep = [None if Sigt is None else [None if a is None else _f64(a) for a in Sigt],
            None if STotal is None else [None if a is None else _f64(a) for a in STotal]]
    ctxs = (C.c_void_p * n)(*[c.h for c in contexts])
    rc = lib.umt_control_sweep_sets(ctxs, n, ptrs(keep[0], lambda c: (c.nz, c.G)), ptrs(keep[1], lambda c: (c.nc, c.G)), C.c_double(tau),
                                    int(bool(savePsi)), int(maxFluxIters), C.c_double(fluxTol), C.byref(it), ptrs(PhiTotal, lambda c: (c.nc, c.G)))
    if rc:
        raise UmtError(f"umt_control_sweep_sets -> {rc}: " + "; ".join(lib.umt_last_error(c.h).decode() for c in contexts))
    return it.value


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = (C.c_ubyte * 128)()
    rc = lib.umt_nccl_unique_id(buf)
    if rc:
        raise UmtError(f"umt_nccl_unique_id -> {rc}")
    return bytes(buf)
