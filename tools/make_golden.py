#!/usr/bin/env python
"""Freeze small oracle outputs as fixtures under tests/golden/ (run in the build container: `python tools/make_golden.py`).

The reference ships no golden vectors and cannot be built here (SURVEY.md sections 4, 8c), so these fixtures do NOT pin the
oracle to the reference ("parity unpinned"); they pin the oracle to itself across compilers/machines and give the GPU tests a
second, oracle-library-independent target.  Inputs are regenerated from the seeds recorded in each file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from tests import common as T           # noqa: E402
from umt_b200 import mesh as M          # noqa: E402
from umt_b200 import problem as PR      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def case_xyz():
    m = M.tiled_mesh((1, 1, 2))
    p = T.make_problem_3d(m, 1, 2, 4, seed=1234)
    phi1 = T.oracle_sweep_3d(p, False)
    phi2 = T.oracle_sweep_3d(p, True)
    np.savez_compressed(os.path.join(OUT, "sweep_xyz_tiled112_P1A2_G4_seed1234.npz"), phi_nonfinal=phi1, phi_final=phi2,
                        psib=p.PsiB, psi_corner0=p.Psi[:, 0, :], nHyperPlanes=p.sched["nHyperPlanes"], nextZ0=p.sched["nextZ"][0])


def case_rz():
    m = M.tiled_mesh((2, 2, 0))
    p = T.make_problem_rz(m, 2, 2, 4, seed=1234)
    phi1 = T.oracle_sweep_rz(p, False)
    phi2 = T.oracle_sweep_rz(p, True)
    np.savez_compressed(os.path.join(OUT, "sweep_rz_tiled22_P2A2_G4_seed1234.npz"), phi_nonfinal=phi1, phi_final=phi2, psib=p.PsiB,
                        weight=p.q["weight"], angDerivFac=p.q["angDerivFac"], nHyperPlanes=p.sched["nHyperPlanes"])


def case_quadrature():
    om, w = O.quad_xyz(4, 4)
    q = O.quad_rz(2, 2)
    np.savez_compressed(os.path.join(OUT, "quadrature_product.npz"), omega_xyz_P4A4=om, weight_xyz_P4A4=w, omega_rz_P2A2=q["omega"],
                        weight_rz_P2A2=q["weight"], tauW1_rz=q["quadTauW1"], tauW2_rz=q["quadTauW2"])


def case_cycle():
    m = M.tiled_mesh((1, 1, 1))
    p = T.make_problem_3d(m, 2, 2, 2, driver_like=True)
    p.tau = PR.tau()
    p.Sigt[:] = p.tau
    p.Psi[:] = PR.wtiso(3) * O.planck_groups_ref(PR.TR0, PR.group_bounds(2), 1.0, PR.SPEED_LIGHT * PR.RAD_CONSTANT)
    rows = []
    for _ in range(3):
        r = T.oracle_cycle_3d(p, PR.DT, PR.TFLOOR ** 4)
        rows.append([r["EnergyRadiation"], r["TrMax"], r["PowerEscape"], r["EnergyCheck"]])
    np.savez_compressed(os.path.join(OUT, "cycle_tiled111_P2A2_G2.npz"), edits=np.array(rows), phi=r["phi"])


def gta_inputs(mesh, G, seed):
    rng = np.random.default_rng(seed)
    nz, nc = mesh.nzones, mesh.ncornr
    Siga, Sigs, Eta = 5 * rng.random((nz, G)), 20 * rng.random((nz, G)), 0.5 * rng.random(nc)
    Chi = rng.random((nc, G))
    Chi /= Chi.sum(1, keepdims=True)
    Phi = rng.random((nc, G))
    return Siga, Sigs, Eta, Chi, Phi


def gta_solve(mesh, G=4, seed=7):
    om = O.OMesh(mesh)
    g = O.geometry(om)
    rz = mesh.ndim == 2
    q = O.gta_quad_rz() if rz else None
    omega, w = (q["omega"], q["weight"]) if rz else O.gta_quad_xyz()
    sched = O.schedule(om, g, omega, q["finish"] if rz else None)
    Siga, Sigs, Eta, Chi, Phi = gta_inputs(mesh, G, seed)
    op = O.gta_set_opacity(om, g, PR.tau(1e-3), Siga, Sigs, Eta, Chi)
    gs = O.collision_rate(om, Eta, Siga, Sigs, Phi, np.zeros(mesh.ncornr), 0)
    P = O.GtaProblem(om, g, sched, omega, w, op, gs, PR.wtiso(mesh.ndim), q=q)
    corr, n, err = P.solve(Phi)
    return corr, n, err


def case_gta():
    c3, n3, e3 = gta_solve(M.tiled_mesh((1, 1, 2)))
    c2, n2, e2 = gta_solve(M.tiled_mesh((2, 2, 0)))
    q = O.gta_quad_rz()
    np.savez_compressed(os.path.join(OUT, "gta_solve_G4_seed7.npz"), corr_xyz_tiled112=c3, iters_xyz=n3, err_xyz=e3, corr_rz_tiled22=c2, iters_rz=n2,
                        err_rz=e2, rz_angDerivFac=q["angDerivFac"], rz_tauW1=q["quadTauW1"])


def case_scheduler():
    N, nCommSets = 2, 4
    problems = [T.make_problem_3d(M.tiled_mesh((2, 2, 1), rank=r, size=N), 1, 2, 2, seed=200 + r) for r in range(N)]
    NA = problems[0].NA
    rng = np.random.default_rng(9)
    nf = [rng.standard_normal((len(T.shared_boundaries(p.mesh)), NA)) for p in problems]
    order, recv = T.oracle_sweep_scheduler(problems, nCommSets, nf)
    np.savez_compressed(os.path.join(OUT, "scheduler_2domains_4sets_seed9.npz"), order=np.array(order), recv=np.array([r[0] for r in recv]))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for f in (case_xyz, case_rz, case_quadrature, case_cycle, case_gta, case_scheduler):
        f()
    print(sorted(os.listdir(OUT)), sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")
