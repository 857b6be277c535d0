"""GPU: the dataflow kernels (values as their own completion flags: default r-z grey sweep, optional r-z Sn sweeps) must turn a value
that never becomes real into UMT_ERR_STATE, not into a hung GPU.  The trigger here is an input that carries the very bit pattern
the kernels use as their 'not computed yet' mark; UMT_SPIN_LIMIT shortens the polling budget so the test takes milliseconds."""
import os
import struct

import numpy as np
import pytest

from tests import common as T
from umt_b200 import mesh as M
from umt_b200.teton import UmtError

pytestmark = pytest.mark.gpu
MARK = struct.unpack("<d", struct.pack("<Q", 0xFFFFDEADFFFFDEAD))[0]


@pytest.mark.timeout(120)
def test_rz_dataflow_sweep_with_a_marked_input_returns_an_error():
    os.environ["UMT_RZ_KERNEL"] = "recflow"
    os.environ["UMT_SPIN_LIMIT"] = "4096"
    try:
        mesh = M.tiled_mesh((3, 3, 0))
        p = T.make_problem_rz(mesh, 2, 2, 8)
        bad = p.PsiB.copy()
        bad[:] = MARK                      # every incident boundary flux reads as 'not computed yet' for ever
        ctx = T.gpu_context_rz(p)
        ctx.sweep(False)                   # sane inputs: fine
        assert np.isfinite(ctx.download_phi()).all()
        ctx.upload_state(None, bad, None, None, p.tau)
        with pytest.raises(UmtError, match="gave up waiting"):
            ctx.sweep(False)
        ctx.upload_state(None, p.PsiB, None, None, p.tau)   # and the context is usable again
        ctx.sweep(False)
        assert np.isfinite(ctx.download_phi()).all()
        ctx.close()
    finally:
        del os.environ["UMT_RZ_KERNEL"], os.environ["UMT_SPIN_LIMIT"]


@pytest.mark.timeout(120)
def test_rz_grey_dataflow_sweep_with_a_marked_input_returns_an_error():
    """the default r-z grey sweep (gta_sweep_rz_flow_kernel) with incident grey boundary fluxes that read as 'not computed yet'"""
    from tests.test_gpu_gta_rz import _setup
    os.environ["UMT_SPIN_LIMIT"] = "4096"
    try:
        s = _setup(M.tiled_mesh((3, 3, 0)))
        ctx = s["ctx"]
        nc, nb = s["mesh"].ncornr, s["mesh"].nbelem
        ctx.gta_compute_opacity(s["Siga"], s["Sigs"], s["Eta"], s["Chi"].copy())
        ctx.collision_rate(s["Eta"], s["Siga"], s["Sigs"], 0)
        P = np.random.default_rng(1).random(nc)
        phi_inc, _ = ctx.gta_sweep(P, None, np.zeros((8, nb)), True)
        assert np.isfinite(phi_inc).all()
        with pytest.raises(UmtError, match="gave up waiting"):
            ctx.gta_sweep(P, None, np.full((8, nb), MARK), True)
        phi_inc2, _ = ctx.gta_sweep(P, None, np.zeros((8, nb)), True)   # usable again, same answer
        assert np.array_equal(phi_inc, phi_inc2)
        ctx.close()
    finally:
        del os.environ["UMT_SPIN_LIMIT"]
