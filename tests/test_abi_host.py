"""CPU: the C-ABI library loads, exports every symbol include/umt_sweep.h declares, and its
host-side logic (product quadrature, rtorder/snnext sweep schedules, exit lists, argument
checking) agrees with the oracle.  No kernel is launched: contexts here are host-only
(device = -1) and every compute entry point must refuse to run rather than fall back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from umt_b200 import mesh as M
from umt_b200 import teton

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names += re.findall(r"\b((?:umt|gpu)_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = teton.load_library()
    names = _declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"include/*.h declares symbols the library does not export: {missing}"
    assert b"sm_100a" in lib.umt_version()


def test_create_rejects_bad_sizes_and_missing_gpu():
    lib = teton.load_library()
    h = C.c_void_p()
    assert lib.umt_ctx_create(-1, 4, 1, 8, 0, 3, 8, 1, C.byref(h)) == 1      # ndim = 4
    assert lib.umt_ctx_create(-1, 3, 1, 8, 0, 2, 8, 1, C.byref(h)) == 1      # maxcf != ndim
    assert b"bad sizes" in lib.umt_last_error(None)
    import torch
    if not torch.cuda.is_available():
        rc = lib.umt_ctx_create(0, 3, 1, 8, 0, 3, 8, 1, C.byref(h))
        assert rc == 2 and not h.value, "without a GPU a device context must fail (UMT_ERR_CUDA), not fall back"


def _host_ctx(m, G=2):
    return teton.SweepContext.from_mesh(m, G, device=-1)


def test_host_only_context_refuses_compute():
    m = M.box_mesh((2, 2, 2))
    ctx = _host_ctx(m)
    g = O.geometry(O.OMesh(m))
    ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], A_bdy=g["A_bdy"])
    ctx.build_product_quadrature(1, 1, 1)
    ctx.build_schedule()
    for call in (lambda: ctx.sweep(), lambda: ctx.upload_state(np.zeros((8, m.ncornr, 2))), lambda: ctx.compute_geometry(m.px),
                 lambda: ctx.download_phi(), lambda: ctx.init_phi_total()):
        with pytest.raises(teton.UmtError):
            call()
    # the group-set call refuses as well (no CPU fallback), and says which entry point it was
    ctx2 = _host_ctx(m)
    out = [np.zeros((m.ncornr, 2)), np.zeros((m.ncornr, 2))]
    with pytest.raises(teton.UmtError, match="umt_control_sweep"):
        teton.control_sweep_sets([ctx, ctx2], None, None, 1.0, out)
    ctx2.close()
    ctx.close()


@pytest.mark.parametrize("ndim,P,A,axis", [(3, 1, 1, 1), (3, 2, 2, 1), (3, 4, 4, 1), (3, 3, 2, 2), (3, 2, 3, 3), (2, 2, 2, 1), (2, 3, 4, 1)])
def test_product_quadrature_matches_oracle(ndim, P, A, axis):
    m = M.box_mesh((2, 2, 2) if ndim == 3 else (2, 2))
    ctx = _host_ctx(m)
    NA = ctx.build_product_quadrature(P, A, axis)
    om, w = ctx.get_quadrature()
    if ndim == 3:
        om_ref, w_ref = O.quad_xyz(P, A, axis)
    else:
        q = O.quad_rz(P, A)
        om_ref, w_ref = q["omega"], q["weight"]
    assert NA == len(w_ref)
    assert np.array_equal(om, om_ref) and np.array_equal(w, w_ref)   # same tables, same operation order: bit-identical
    ctx.close()


def test_product_quadrature_rejects_bad_orders():
    ctx = _host_ctx(M.box_mesh((2, 2, 2)))
    for bad in ((0, 1, 1), (1, 33, 1), (1, 1, 4)):
        with pytest.raises(teton.UmtError):
            ctx.build_product_quadrature(*bad)
    ctx.close()


MESHES = [("tiled", lambda: M.tiled_mesh((2, 2, 2))), ("unstruct", lambda: M.unstruct_box_mesh(2)),
          ("warped", lambda: M.box_mesh((4, 4, 4), warp=0.35, seed=3)), ("jacobi", lambda: M.box_mesh((7, 6, 5), warp=0.9, seed=2)),
          ("tiled2d", lambda: M.tiled_mesh((3, 3, 0)))]


@pytest.mark.parametrize("name,mk", MESHES)
def test_schedule_matches_oracle(name, mk):
    """umt_build_schedule (csrc/schedule.cpp) against the oracle's restatement of rtorder/snnext/
    snneed/findseeds/getDownStreamData/cyclebreaker/sccsearch and findexit: integer-exact."""
    m = mk()
    om = O.OMesh(m)
    g = O.geometry(om)
    ctx = _host_ctx(m)
    if m.ndim == 3:
        omega, w = O.quad_xyz(2, 2)
        skip = None
        ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], A_bdy=g["A_bdy"])
        ctx.set_quadrature(omega, w)
    else:
        q = O.quad_rz(2, 2)
        omega, w, skip = q["omega"], q["weight"], q["finish"]
        ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], g["Area"], g["RadiusFP"], g["RadiusEZ"], g["A_bdy"])
        ctx.set_quadrature(omega, w, q["start"], q["finish"][:len(w)], q["angDerivFac"], q["quadTauW1"], q["quadTauW2"])
    ctx.build_schedule()
    s = O.schedule(om, g, omega, skip)
    for a in range(len(w)):
        got = ctx.get_schedule(a + 1)
        nh = int(s["nHyperPlanes"][a])
        assert got["nHyperPlanes"] == nh, (name, a)
        if nh == 0:
            continue
        assert np.array_equal(got["zonesInPlane"], s["zonesInPlane"][a][:nh])
        assert np.array_equal(got["nextZ"], s["nextZ"][a])
        assert np.array_equal(got["nextC"], s["nextC"][a])
        off, n = int(s["cycleOffSet"][a]), int(s["numCycles"][a])
        assert np.array_equal(got["cycleList"], s["cycleList"][off:off + n])
        info = ctx.schedule_info(a + 1)
        assert info[1] == n and info[2] == int((s["nextZ"][a] < 0).sum())
    if name in ("warped", "jacobi"):
        assert s["totalCycles"] > 0
    if name == "jacobi":
        assert (s["nextZ"] < 0).any()
    ctx.close()


def test_set_schedule_validates():
    m = M.box_mesh((2, 2, 2))
    ctx = _host_ctx(m)
    omega, w = O.quad_xyz(1, 1)
    ctx.set_quadrature(omega, w)
    nz, nc = m.nzones, m.ncornr
    with pytest.raises(teton.UmtError):   # planes do not hold all zones
        ctx.set_schedule(1, 1, [nz - 1], np.arange(1, nz + 1), np.ones(nc))
    with pytest.raises(teton.UmtError):   # zone id out of range
        ctx.set_schedule(1, 1, [nz], np.full(nz, nz + 1), np.ones(nc))
    with pytest.raises(teton.UmtError):   # angle out of range
        ctx.set_schedule(9, 1, [nz], np.arange(1, nz + 1), np.ones(nc))
    ctx.close()


def test_connectivity_validation():
    m = M.box_mesh((2, 2, 2))
    bad = M.box_mesh((2, 2, 2))
    bad.cFP = bad.cFP.copy()
    bad.cFP[0, 0] = m.ncornr + m.nbelem + 5
    with pytest.raises(teton.UmtError):
        _host_ctx(bad)


def test_planck_groups_match_reference_build():
    """umt_planck_groups (csrc/planck.cu host path) against the reference's own NormalizedBlackBody.cc
    compiled from /root/reference into oracle/_ref (cpu_baseline.kind "reference" for this one function)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libnbb_ref.so")):
        pytest.skip("oracle/_ref not built")
    from umt_b200 import problem as PR
    for G, Tr in ((2, 0.05), (16, 0.05), (128, 0.05), (64, 0.5), (7, 3.0)):
        b = PR.group_bounds(G)
        ours = teton.planck_groups(Tr, b)
        ref = O.planck_groups_ref(Tr, b)
        assert np.abs(ours - ref).max() <= 1e-14 * ref.max()


# ---------------------------------------------------------------------------
# the reference's own C seam (include/teton_gpu_compat.h)
# ---------------------------------------------------------------------------
_GPU_SWEEP_ARGS = ("Angle nHyperPlanes nZonesInPlane nextZ nextC STotal tau Psi Groups Volume Sigt nCFacesArray ndim maxcf ncorner "
                   "A_fp omega cFP Psi1 nbelem A_ez cEZ NumAngles quadwt Phi PsiB maxCorner mem0solve1 streamIdPtr totalStreams savePsi "
                   "numCycles cycleOffSet cyclePsi cycleList b0 nBdyElem PsiBMref Mref Geom_numCorner Geom_cOffSet").split()


def _prototype_args(src, name):
    m = re.search(r"void\s+" + name + r"\s*\((.*?)\)\s*[;{]", re.sub(r"/\*.*?\*/|//[^\n]*", "", src, flags=re.S), flags=re.S)
    assert m, name
    return [(a.split()[0], a.replace("*", " ").split()[-1]) for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]


def test_compat_header_has_the_reference_signature():
    """gpu_sweepucbxyz: 41 by-reference arguments in the order of the Fortran interface block
    (gpu/SweepUCBxyzToGPU.F90:43-91); compared with the reference source itself when the tree is present."""
    hdr = open(os.path.join(ROOT, "include", "teton_gpu_compat.h")).read()
    args = _prototype_args(hdr, "gpu_sweepucbxyz")
    assert [n for _, n in args] == _GPU_SWEEP_ARGS and len(args) == 41
    ref = "/root/reference/src/teton/gpu/GPU_SweepUCBxyz.cu"
    if os.path.exists(ref):
        rargs = _prototype_args(open(ref).read(), "gpu_sweepucbxyz")
        assert rargs == args, "types/names differ from the reference's definition"
        assert _prototype_args(open(ref).read(), "gpu_streamsynchronize") == _prototype_args(hdr, "gpu_streamsynchronize")
    lib = teton.load_library()
    for n in ("gpu_sweepucbxyz", "gpu_streamsynchronize", "gpu_devicesynchronize"):
        assert hasattr(lib, n)


def test_quadrature_tables_match_the_reference_data_module(tmp_path):
    """umt_b200/csrc/quad_tables.inc (used by the library and the oracle alike) is the numeric content of the reference's
    mods/QuadratureData_mod.F90:1144-4528: regenerated here from the reference tree and compared value by value (skipped where
    the tree is absent, e.g. on the GPU box)."""
    ref = "/root/reference/src/teton/mods/QuadratureData_mod.F90"
    if not os.path.exists(ref):
        pytest.skip("reference tree absent")
    import subprocess
    import sys
    out = tmp_path / "quad_tables.inc"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "extract_quadrature_tables.py"), ref, str(out)],
                          stdout=subprocess.DEVNULL)
    num = re.compile(r"[-+]?\d+\.\d+e[-+]\d+")
    a = [float(x) for x in num.findall(open(out).read())]
    b = [float(x) for x in num.findall(open(os.path.join(ROOT, "umt_b200", "csrc", "quad_tables.inc")).read())]
    assert len(a) == len(b) == 6 * 528 and a == b


def test_problem_constants_match_the_reference_sources():
    """umt_b200/problem.py restates the driver's problem; its constants are read back from the reference sources when present."""
    base = "/root/reference/src/teton"
    if not os.path.exists(base):
        pytest.skip("reference tree absent")
    from umt_b200 import problem as PR
    rc = open(os.path.join(base, "mods", "radconstant_mod.F90")).read()
    assert float(re.search(r"speed_light\s*=\s*([0-9.]+)_adqt", rc).group(1)) == PR.SPEED_LIGHT
    assert float(re.search(r"rad_constant\s*=\s*([0-9.]+)_adqt", rc).group(1)) == PR.RAD_CONSTANT
    drv = open(os.path.join(base, "driver", "test_driver.cc")).read()
    for key, val in (("thermo_density", PR.RHO), ("electron_specific_heat", PR.CV), ("radiation_temperature", PR.TR0),
                     ("electron_temperature", PR.TE0)):
        m = re.search(r'material_field_vals\[1\]\["%s"\]\s*=\s*([0-9.eE+-]+);' % key, drv)
        assert m and float(m.group(1)) == val, key


def test_unstructured_box_base_mesh_matches_the_reference_generator():
    """BASELINE configs[3]: the 12-vertex / 6-hex-per-layer base mesh of umt_b200/mesh.py (_UB_XY6, _UB_QUADS) is the one
    driver/makeUnstructuredBox.cc:39-110 builds (its AddVertex / AddHex calls are read back here; the MFEM refinement on top
    of it stays a stand-in)."""
    src_path = "/root/reference/src/teton/driver/makeUnstructuredBox.cc"
    if not os.path.exists(src_path):
        pytest.skip("reference tree absent")
    from umt_b200 import mesh as MM
    src = open(src_path).read()
    verts = re.findall(r"mesh\.AddVertex\(([^;]*?),\s*([^;]*?),\s*i \* width / 3\.0\);", src)
    assert len(verts) == 12
    xy = np.array([[eval(x, {"width": 1.0}), eval(y, {"width": 1.0})] for x, y in verts])
    assert np.allclose(xy, MM._UB_XY6 / 6.0, rtol=0, atol=1e-15)
    letters = {k: i for i, k in enumerate("abcdefghijkl")}
    hexes = re.findall(r"mesh\.AddHex\((\w) \+ B, (\w) \+ B, (\w) \+ B, (\w) \+ B,", src)
    assert len(hexes) == 12 and hexes[:6] == hexes[6:]       # two identical layers
    assert [[letters[c] for c in h] for h in hexes[:6]] == MM._UB_QUADS.tolist()
    assert MM._UB_LAYERS == 2
