"""GPU: whole mini-app time steps through the C ABI (umt_b200/cycle.py) against the oracle's cycle, to the north-star's
end-of-cycle tolerance (radiation energy, temperatures: 1e-10 relative), and the reference driver's own acceptance check
|EnergyCheck / EnergyRadiation| <= 1e-9 (test_driver.cc:1981-2003)."""
import numpy as np
import pytest

from tests import common as T
from umt_b200 import mesh as M
from umt_b200 import problem as PR
from umt_b200.cycle import MiniAppCycle
from umt_b200.teton import SweepContext, planck_groups

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mk,P,A,G", [(lambda: M.tiled_mesh((2, 2, 2)), 2, 2, 2), (lambda: M.unstruct_box_mesh(2), 2, 2, 8)])
def test_cycles_match_oracle(mk, P, A, G):
    mesh = mk()
    p = T.make_problem_3d(mesh, P, A, G, driver_like=True)
    p.tau = PR.tau()
    p.Sigt[:] = p.tau
    B = planck_groups(PR.TR0, PR.group_bounds(G), 1.0, PR.SPEED_LIGHT * PR.RAD_CONSTANT)   # InitTeton.F90:96-101
    p.Psi[:] = PR.wtiso(3) * B
    ctx = SweepContext.from_mesh(mesh, G)
    ctx.set_geometry(p.geom["Volume"], p.geom["A_fp"], p.geom["A_ez"], A_bdy=p.geom["A_bdy"])
    ctx.set_quadrature(p.omega, p.weight)
    ctx.build_schedule()
    cyc = MiniAppCycle(ctx, mesh, G)
    assert T.relerr(ctx.download_psi(), p.Psi) <= 1e-13          # device-built initial psi = host Planck integrals
    for _ in range(3):   # BASELINE configs[0] runs 10 cycles; 3 keep the oracle side short
        ref = T.oracle_cycle_3d(p, PR.DT, PR.TFLOOR ** 4)
        ed = cyc.step()
        assert ed["sweeps"] == 3
        for k in ("EnergyRadiation", "TrMax", "PowerEscape", "EnergyRadBOC"):
            assert abs(ed[k] - ref[k]) <= 1e-10 * abs(ref[k]), k
        assert ed["TeMax"] == PR.TE0
        assert T.relerr(ed["RadPowerEscape"], ref["RadPowerEscape"]) <= 1e-10
        assert T.relerr(ctx.download_phi(), ref["phi"]) <= 1e-12
        assert abs(ed["EnergyCheck"] / ed["EnergyRadiation"]) <= 1e-9      # "RESULT CHECK PASSED"
        assert abs(ref["EnergyCheck"] / ref["EnergyRadiation"]) <= 1e-9
    dens = ctx.cycle_edits(PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.TFLOOR ** 4, want_trz=True, want_density=True)
    assert T.relerr(dens["trz"], ref["trz"]) <= 1e-10
    V = p.geom["Volume"]
    d_ref = np.add.reduceat(V[:, None] * ref["phi"], mesh.cOffSet, axis=0) / p.geom["VolumeZone"][:, None] / PR.SPEED_LIGHT
    assert T.relerr(dens["RadEnergyDensity"], d_ref) <= 1e-12
    ctx.close()


def test_rz_cycles_match_oracle():
    """The r-z mini-app time step (BASELINE configs[1]'s geometry): boundary edit with the 2 pi r factor (BoundaryEdit.F90:123),
    energies with geometryFactor = 2 pi, three cycles to 1e-10."""
    G = 4
    mesh = M.tiled_mesh((3, 3, 0))
    p = T.make_problem_rz(mesh, 2, 2, G, driver_like=True)
    p.tau = PR.tau()
    p.Sigt[:] = p.tau
    B = planck_groups(PR.TR0, PR.group_bounds(G), 1.0, PR.SPEED_LIGHT * PR.RAD_CONSTANT)
    p.Psi[:] = PR.wtiso(2) * B
    ctx = T.gpu_context_rz(p, own_schedule=True)
    cyc = MiniAppCycle(ctx, mesh, G)
    assert T.relerr(ctx.download_psi(), p.Psi) <= 1e-13
    for _ in range(3):
        ref = T.oracle_cycle_rz(p, PR.DT, PR.TFLOOR ** 4)
        ed = cyc.step()
        for k in ("EnergyRadiation", "TrMax", "PowerEscape", "EnergyRadBOC"):
            assert abs(ed[k] - ref[k]) <= 1e-10 * abs(ref[k]), k
        assert T.relerr(ed["RadPowerEscape"], ref["RadPowerEscape"]) <= 1e-10
        assert T.relerr(ctx.download_phi(), ref["phi"]) <= 1e-12
        assert abs(ed["EnergyCheck"] / ed["EnergyRadiation"]) <= 1e-9
        assert abs(ref["EnergyCheck"] / ref["EnergyRadiation"]) <= 1e-9
    ctx.close()
