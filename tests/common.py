"""Shared problem builders for the tests.  The oracle supplies the reference-side
inputs exactly as Teton's Fortran would hand them to the library (geometry,
quadrature, schedules) and the expected outputs."""
import numpy as np

from oracle import oracle as O
from umt_b200 import mesh as M

SPEED_LIGHT = 299.792458           # mods/radconstant_mod.F90:27
RAD_CONSTANT = 0.013720169264801055


class Problem:
    pass


def make_problem_3d(mesh, npolar=2, nazimuthal=2, G=4, seed=1234, tau=None, driver_like=False):
    p = Problem()
    p.mesh = mesh
    p.om = O.OMesh(mesh)
    p.geom = O.geometry(p.om)
    p.omega, p.weight = O.quad_xyz(npolar, nazimuthal)
    p.NA = len(p.weight)
    p.sched = O.schedule(p.om, p.geom, p.omega)
    p.G = G
    p.bdy = [O.bdy_exit(p.om, p.geom, p.omega[a]) for a in range(p.NA)]
    rng = np.random.default_rng(seed)
    nc, nb, nz = mesh.ncornr, mesh.nbelem, mesh.nzones
    if driver_like:
        # mini-app problem: Sigt = tau = 1/(c dt), STotal = 0
        p.tau = 1.0 / (SPEED_LIGHT * 1e-3) if tau is None else tau
        p.Sigt = np.full((nz, G), p.tau)
        p.STotal = np.zeros((nc, G))
        p.Psi = np.tile(np.linspace(1.0, 2.0, G), (p.NA, nc, 1)).copy()
        p.PsiB = np.zeros((p.NA, nb, G))
    else:
        p.tau = 3.0 if tau is None else tau
        p.Sigt = p.tau + 20.0 * rng.random((nz, G))
        p.STotal = rng.random((nc, G))
        p.Psi = 0.5 + rng.random((p.NA, nc, G))
        p.PsiB = 0.5 + rng.random((p.NA, nb, G))
    p.cyclePsi = np.zeros((max(int(p.sched["totalCycles"]), 1), G))
    for a in range(p.NA):   # initCyclePsi
        for m in range(p.sched["numCycles"][a]):
            mc = p.sched["cycleOffSet"][a] + m
            p.cyclePsi[mc] = p.Psi[a, p.sched["cycleList"][mc] - 1]
    return p


def oracle_sweep_3d(p, savePsi, nthreads=0):
    """One ControlSweep on the oracle; mutates p.Psi (if savePsi), p.PsiB, p.cyclePsi; returns PhiTotal."""
    return O.setsweep_xyz(p.om, p.geom, p.sched, p.omega, p.weight, p.tau, p.STotal, p.Sigt, p.Psi, p.PsiB,
                          p.cyclePsi, savePsi, nthreads)


def gpu_context_3d(p, device=0, own_schedule=False, own_geometry=False, own_quadrature=None):
    from umt_b200.teton import SweepContext
    m = p.mesh
    ctx = SweepContext.from_mesh(m, p.G, device)
    if own_geometry:
        ctx.compute_geometry(m.px)
    else:
        ctx.set_geometry(p.geom["Volume"], p.geom["A_fp"], p.geom["A_ez"], A_bdy=p.geom["A_bdy"])
    if own_quadrature:
        ctx.build_product_quadrature(*own_quadrature)
    else:
        ctx.set_quadrature(p.omega, p.weight)
    if own_schedule:
        ctx.build_schedule()
    else:
        s = p.sched
        for a in range(p.NA):
            off, n = s["cycleOffSet"][a], s["numCycles"][a]
            ctx.set_schedule(a + 1, s["nHyperPlanes"][a], s["zonesInPlane"][a][:s["nHyperPlanes"][a]], s["nextZ"][a], s["nextC"][a],
                             s["cycleList"][off:off + n], p.bdy[a])
    ctx.upload_state(p.Psi, p.PsiB, p.Sigt, p.STotal, p.tau)
    ctx.init_radiation_field_needed = True
    return ctx


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def mixed_err(a, b, rtol=1e-12, atol_scale=1e-14):
    """max |a-b| / (rtol*|b| + atol_scale*max|b|): <= 1 passes.  Angular fluxes next to
    vacuum boundaries are differences of nearly equal numbers (values ~1e-6 of the field
    maximum, some slightly negative), so a pure element-wise relative test measures the
    compiler's FMA contraction, not the algorithm: the same oracle source built with and
    without -ffp-contract differs by 5e-13 element-wise on PsiB but 6e-16 in max norm."""
    a = np.asarray(a); b = np.asarray(b)
    if a.size == 0:
        return 0.0
    scale = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b) / (rtol * np.abs(b) + atol_scale * scale + 1e-300)))
