"""One Teton time step (teton_radtr, aux/radtr.F90:12-60) of the mini-app build driven through the C ABI:
what Teton's unchanged Fortran would do around the sweep, with every per-corner field staying on the device.

  initializeSets (control/initializeSets.F90:71-91,481-545): tau = 1/(c dt); Sigt = siga + sigs + tau
      (setTotalOpacity.F90:52; mini-app: siga = sigs = 0); setBoundarySources (PsiB = 0); initPhiTotal;
      initializeRadiationField; initializeZones (EnergyRadBOC)
  rtmainsn (rt/rtmainsn.F90:167-249): temperature iterations — the mini-app cannot exit on the first one (:210) and
      its errors are exactly zero on the second (ConvergenceTest.F90:130,156) — each a LinearSolver = one ControlSweep
      (GTA compiled out, LinearSolver.F90:87-121); then the final ControlSweep(savePsi = true)
  finalizeSets + rtedit (aux/rtedit.F90:142-232): EnergyRadiation, TrMax, PowerEscape, EnergyCheck
"""
from __future__ import annotations

import numpy as np

from . import problem as PR


class MiniAppCycle:
    def __init__(self, ctx, mesh, ngr, dt=PR.DT, Tr0=PR.TR0, Te0=PR.TE0, tfloor=PR.TFLOOR, incident_flux_max_it=2, flux_tol=1e-6):
        self.ctx, self.mesh, self.G, self.dt = ctx, mesh, ngr, dt
        self.tau = PR.tau(dt)
        self.tr4floor = tfloor ** 4
        self.flux_iters, self.flux_tol = incident_flux_max_it, flux_tol
        nz, nc = mesh.nzones, mesh.ncornr
        self.tec = np.full(nc, Te0)
        # InitTeton.F90:82-118: psi = wtiso * B_g(Tr0) in every corner and angle
        ctx.upload_state(None, None, np.full((nz, ngr), self.tau), np.zeros((nc, ngr)), self.tau)
        ctx.init_teton(np.full(nz, Tr0), PR.group_bounds(ngr), PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(mesh.ndim), 0.0)
        self.history = []

    def step(self):
        ctx = self.ctx
        ctx.upload_state(None, None, np.full((self.mesh.nzones, self.G), self.tau), None, self.tau)   # setTotalOpacity
        ctx.set_boundary_sources()
        ctx.init_phi_total()
        ctx.init_radiation_field()
        boc = ctx.cycle_edits(PR.SPEED_LIGHT, PR.RAD_CONSTANT, self.tr4floor)                          # initializeZones: EnergyRadBOC
        sweeps = passes = 0
        for _temp_iter in range(2):                                                                   # rtmainsn.F90:167-236
            passes += ctx.sweep(False, self.flux_iters, self.flux_tol)
            sweeps += 1
        passes += ctx.sweep(True, self.flux_iters, self.flux_tol)                                     # rtmainsn.F90:239-249
        sweeps += 1
        ed = ctx.cycle_edits(PR.SPEED_LIGHT, PR.RAD_CONSTANT, self.tr4floor)
        # advanceMaterialProperties.F90:96-123: tez = volume average of tec (nothing heats the material in the mini-app)
        ed["TeMax"] = float(self.tec.max())
        ed["EnergyRadBOC"] = boc["EnergyRadiation"]
        ed["deltaERad"] = ed["EnergyRadiation"] - boc["EnergyRadiation"]
        ed["EnergyCheck"] = self.dt * (ed["PowerIncident"] - ed["PowerEscape"]) - ed["deltaERad"]     # rtedit.F90:231 (deltaEMat = 0)
        ed["sweeps"], ed["flux_passes"] = sweeps, passes
        self.history.append(ed)
        return ed


class FullPhysicsStep:
    """The linear solve of a temperature iteration with grey acceleration switched on (rt/LinearSolver.F90:55-121, what the mini-app
    build compiles out), every per-corner field staying on the device:

      getCollisionRate(residualFlag = 0)            GTA%GreySource = sum_g (Eta siga + sigs) PhiTotal
      ControlSweep(savePsi)                         the multigroup sweep + PhiTotal
      getCollisionRate(residualFlag = 1)            GreySource = new collision rate - old one (the residual the grey solve corrects)
      GTASolver + addGreyCorrections                BiCGSTAB on the grey transport operator; PhiTotal += correction * Chi

    followed, between iterations, by the rebuild of the isotropic source GSet%STotal from the corrected PhiTotal (umt_build_source:
    the mini-app reference never fills STotal, so that formula is this library's documented extension; parity unpinned).  Opacities
    are the caller's: Mat%Siga, Mat%Sigs (nzones, ngr), Mat%Eta (ncornr), GTA%Chi (ncornr, ngr), emission (ncornr, ngr)."""

    def __init__(self, ctx, mesh, ngr, Siga, Sigs, Eta, Chi, EmissionRate, dt=PR.DT, flux_iters=2, flux_tol=1e-6):
        self.ctx, self.mesh, self.G = ctx, mesh, ngr
        self.Siga, self.Sigs, self.Eta, self.Emis = Siga, Sigs, Eta, EmissionRate
        self.Chi = np.ascontiguousarray(Chi, dtype=np.float64).copy()
        self.tau = PR.tau(dt)
        self.flux_iters, self.flux_tol = flux_iters, flux_tol
        # setTotalOpacity.F90:52: Sigt = siga + sigs + tau; then GTA setup (angle set, sweep order, setGTAOpacityNEW, transfer matrices)
        ctx.upload_state(None, None, Siga + Sigs + self.tau, None, self.tau)
        ctx.gta_setup()
        self.opacity = ctx.gta_compute_opacity(Siga, Sigs, Eta, self.Chi)   # Chi comes back rescaled, as setGTAOpacityNEW leaves it
        self.history = []

    def rebuild_source(self):
        """GSet%STotal from the PhiTotal on the device (stays there; returned for inspection)"""
        return self.ctx.build_source(self.Siga, self.Sigs, self.Eta, self.Chi, self.Emis)

    def linear_solver(self, savePsi=False):
        ctx = self.ctx
        ctx.collision_rate(self.Eta, self.Siga, self.Sigs, 0)
        passes = ctx.sweep(savePsi, self.flux_iters, self.flux_tol)
        ctx.collision_rate(self.Eta, self.Siga, self.Sigs, 1)
        corr, n, err = ctx.gta_solve()
        ctx.add_grey_corrections()
        rec = dict(flux_passes=passes, grey_sweeps=n, grey_error=err, correction_max=float(np.abs(corr).max()))
        self.history.append(rec)
        return rec
