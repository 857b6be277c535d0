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


# ---------------------------------------------------------------------------
# RZ
# ---------------------------------------------------------------------------
def make_problem_rz(mesh, npolar=2, nazimuthal=2, G=4, seed=1234, tau=None, driver_like=False):
    p = Problem()
    p.mesh = mesh
    p.om = O.OMesh(mesh)
    p.geom = O.geometry(p.om)
    p.q = O.quad_rz(npolar, nazimuthal)
    p.omega, p.weight = p.q["omega"], p.q["weight"]
    p.NA = len(p.weight)
    p.sched = O.schedule(p.om, p.geom, p.omega, p.q["finish"])
    p.G = G
    p.bdy = [O.bdy_exit(p.om, p.geom, p.omega[a]) for a in range(p.NA)]
    rng = np.random.default_rng(seed)
    nc, nb, nz = mesh.ncornr, mesh.nbelem, mesh.nzones
    if driver_like:
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
    return p


def oracle_sweep_rz(p, savePsi):
    """SetSweep.F90:81-170 for one comm set holding every angle in quadrature order (single domain),
    then getPhiTotal; mutates p.Psi (if savePsi) and p.PsiB; returns PhiTotal."""
    m = p.mesh
    Phi = np.zeros((m.ncornr, p.G))
    PsiM = np.zeros((m.ncornr, p.G))
    Psi1 = np.zeros((m.ncornr + m.nbelem, p.G))
    for a in range(p.NA):
        if p.q["finish"][a]:
            continue
        O.sweep_rz(p.om, p.geom, p.sched, a, p.q, p.tau, p.STotal, p.Sigt, p.Psi, Psi1, PsiM, p.PsiB, Phi, p.bdy[a], savePsi)
    return Phi


def gpu_context_rz(p, device=0, own_schedule=False, own_geometry=False, own_quadrature=None):
    from umt_b200.teton import SweepContext
    m = p.mesh
    ctx = SweepContext.from_mesh(m, p.G, device)
    g, q = p.geom, p.q
    if own_geometry:
        ctx.compute_geometry(m.px)
    else:
        ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], g["Area"], g["RadiusFP"], g["RadiusEZ"], g["A_bdy"])
    if own_quadrature:
        ctx.build_product_quadrature(*own_quadrature)
    else:
        ctx.set_quadrature(q["omega"], q["weight"], q["start"], q["finish"][:p.NA], q["angDerivFac"], q["quadTauW1"], q["quadTauW2"])
    if own_schedule:
        ctx.build_schedule()
    else:
        s = p.sched
        for a in range(p.NA):
            off, n = s["cycleOffSet"][a], s["numCycles"][a]
            nh = s["nHyperPlanes"][a]
            ctx.set_schedule(a + 1, nh, s["zonesInPlane"][a][:nh], s["nextZ"][a], s["nextC"][a], s["cycleList"][off:off + n], p.bdy[a])
    ctx.upload_state(p.Psi, p.PsiB, p.Sigt, p.STotal, p.tau)
    return ctx


# ---------------------------------------------------------------------------
# several spatial domains (one per rank), psib exchange lagged one flux pass
# ---------------------------------------------------------------------------
def shared_boundaries(mesh):
    return [b for b in mesh.boundaries if b.bc_type == M.BC_SHARED]


def angle_set_range(p, a):
    """angle set of angle a as the reference forms it without reflecting boundaries (decomposeAngleSets.F90): 3-D every
    angle its own set, 2-D one set per xi-level."""
    if p.mesh.ndim == 3:
        return a, a + 1
    lev = p.q["level"]
    a0, a1 = a, a + 1
    while a0 > 0 and lev[a0 - 1] == lev[a]:
        a0 -= 1
    while a1 < p.NA and lev[a1] == lev[a]:
        a1 += 1
    return a0, a1


def oracle_exchange_lists(problems):
    """findexit.F90:102-294 for every rank: lists[r][k][a] = (send elements, recv elements), 1-based boundary
    elements of rank r's k-th shared boundary.  Of every angle set the lower rank of a pair classifies the first
    NumAngles/2 angles, the higher rank the rest, and each side negates what it receives (3-D: sets of one angle, so
    the higher rank classifies everything)."""
    out = []
    for r, p in enumerate(problems):
        per_b = []
        for b in shared_boundaries(p.mesh):
            q = problems[b.neighbor]
            bq = [x for x in shared_boundaries(q.mesh) if x.neighbor == r][0]
            assert bq.n_elem == b.n_elem
            mine = p.geom["A_bdy"][b.first_elem - 1:b.first_elem - 1 + b.n_elem] @ p.omega.T      # (n, NA)
            theirs = q.geom["A_bdy"][bq.first_elem - 1:bq.first_elem - 1 + bq.n_elem] @ q.omega.T
            per_a = []
            for a in range(p.NA):
                a0, a1 = angle_set_range(p, a)
                first_half = (a - a0) < (a1 - a0) // 2
                i_decide = first_half if r < b.neighbor else not first_half
                t = np.sign(mine[:, a]) if i_decide else -np.sign(theirs[:, a])
                el = np.arange(b.first_elem, b.first_elem + b.n_elem)
                per_a.append((el[t > 0], el[t < 0]))
            per_b.append(per_a)
        out.append(per_b)
    return out


def oracle_multi_sweep(problems, lists, savePsi, maxFluxIters=1, fluxTol=1e-6):
    """SetSweep.F90 on every rank in lock step (psib lagged one flux pass); returns (PhiTotal per rank, flux passes done,
    IncFlux per rank, one bin per comm set: angle in 3-D, xi-level in 2-D)."""
    N = len(problems)
    NA = problems[0].NA
    nd = problems[0].mesh.ndim
    bin_of = np.arange(NA) if nd == 3 else np.asarray(problems[0].q["level"]) - 1
    nBins = int(bin_of.max()) + 1
    sweep = oracle_sweep_3d if nd == 3 else oracle_sweep_rz

    def exit_currents():
        """setIncidentFlux.F90:84-108 on every rank: ExitFlux[r][k][a]"""
        res = []
        for r, p in enumerate(problems):
            per_b = []
            for k, b in enumerate(shared_boundaries(p.mesh)):
                ex = np.zeros(NA)
                for a in range(NA):
                    el = lists[r][k][a][0]
                    dot = p.geom["A_bdy"][el - 1] @ p.omega[a]
                    ex[a] = p.weight[a] * float((dot * p.PsiB[a, el - 1].sum(axis=1)).sum())
                per_b.append(ex)
            res.append(per_b)
        return res

    def incident(ex):
        inc = []
        for r, p in enumerate(problems):
            v = np.zeros(nBins)
            for b in shared_boundaries(p.mesh):
                q = problems[b.neighbor]
                kq = [i for i, x in enumerate(shared_boundaries(q.mesh)) if x.neighbor == r][0]
                np.add.at(v, bin_of, ex[b.neighbor][kq])
            inc.append(v)
        return inc

    inc = incident(exit_currents())
    it = 0
    while True:
        it += 1
        # SendFlux/RecvFlux: every rank's exiting rows (as they are now) land in the neighbours' incident rows
        snap = [p.PsiB.copy() for p in problems]
        for r, p in enumerate(problems):
            for k, b in enumerate(shared_boundaries(p.mesh)):
                q = problems[b.neighbor]
                kq = [i for i, x in enumerate(shared_boundaries(q.mesh)) if x.neighbor == r][0]
                for a in range(NA):
                    recv = lists[r][k][a][1]
                    send = lists[b.neighbor][kq][a][0]
                    assert len(recv) == len(send)
                    p.PsiB[a, recv - 1] = snap[b.neighbor][a, send - 1]
        phis = [sweep(p, savePsi) for p in problems]
        inc_old, inc = inc, incident(exit_currents())
        if savePsi:
            break
        notconv = 0
        for r in range(N):
            for bn in range(nBins):
                tot = inc[r][bn]
                rel = abs(inc[r][bn] - inc_old[r][bn]) / inc[r][bn] if tot != 0.0 else 0.0
                notconv += not (rel <= fluxTol)
        if notconv == 0 or it >= maxFluxIters:
            break
    return phis, it, inc


oracle_multi_sweep_3d = oracle_multi_sweep


def run_local_group(contexts, fn):
    """Call fn(rank, ctx) from one host thread per context (the in-process communicator needs it)."""
    import threading
    out, err = [None] * len(contexts), [None] * len(contexts)

    def work(r):
        try:
            out[r] = fn(r, contexts[r])
        except BaseException as e:   # noqa: BLE001
            err[r] = e
    th = [threading.Thread(target=work, args=(r,)) for r in range(len(contexts))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


# ---------------------------------------------------------------------------
# reflecting boundaries (snreflect.F90 + findReflectedAngles.F90 / reflectAxis.F90, axis-aligned planes)
# ---------------------------------------------------------------------------
def reflecting_boundaries(mesh):
    return [b for b in mesh.boundaries if b.bc_type == M.BC_REFL]


def oracle_reflected_angles(p):
    """mref[k][a] = 0-based mirror angle of angle a on the k-th reflecting boundary (-1: not incident)."""
    out = []
    for b in reflecting_boundaries(p.mesh):
        A = p.geom["A_bdy"][b.first_elem - 1]
        nmax = int(np.argmax(np.abs(A)))
        mref = np.full(p.NA, -1)
        for a in range(p.NA):
            if p.omega[a] @ A < -1e-15:
                for ia in range(p.NA):   # reflectAxis.F90:101-113, last match wins
                    if abs(p.omega[ia, nmax] + p.omega[a, nmax]) < 1e-6 and all(
                            abs(p.omega[ia, d] - p.omega[a, d]) < 1e-6 for d in range(p.omega.shape[1]) if d != nmax):
                        mref[a] = ia
                assert mref[a] >= 0
        out.append(mref)
    return out


def reflect_stages(mrefs, NA):
    """stage(a) = 1 + max stage of its mirror images; back edges (facing planes) are lagged.  Same DFS as csrc/reflect.cu."""
    deps = [[int(m[a]) for m in mrefs if m[a] >= 0] for a in range(NA)]
    stage, state = [0] * NA, [0] * NA
    for root in range(NA):
        if state[root]:
            continue
        st = [[root, 0]]
        state[root] = 1
        while st:
            a, i = st[-1]
            if i < len(deps[a]):
                st[-1][1] += 1
                m = deps[a][i]
                if state[m] == 0:
                    state[m] = 1
                    st.append([m, 0])
                elif state[m] == 2:
                    stage[a] = max(stage[a], stage[m] + 1)
            else:
                state[a] = 2
                st.pop()
                if st:
                    stage[st[-1][0]] = max(stage[st[-1][0]], stage[a] + 1)
    return stage


def oracle_sweep_3d_reflecting(p, savePsi):
    """SetSweep angle loop with snreflect before each angle; angles ordered by reflection stage, the copies of a
    stage made before its sweeps (what one persistent launch per stage does).  No mesh cycles expected."""
    assert p.sched["totalCycles"] == 0
    m = p.mesh
    mrefs = oracle_reflected_angles(p)
    stage = reflect_stages(mrefs, p.NA)
    refl = reflecting_boundaries(m)
    PhiSets = np.zeros((p.NA, m.ncornr, p.G))
    Psi1 = np.zeros((m.ncornr + m.nbelem, p.G))
    for s in range(max(stage) + 1):
        angles = [a for a in range(p.NA) if stage[a] == s]
        for a in angles:
            for k, b in enumerate(refl):
                if mrefs[k][a] >= 0:
                    sl = slice(b.first_elem - 1, b.first_elem - 1 + b.n_elem)
                    p.PsiB[a, sl] = p.PsiB[mrefs[k][a], sl]
        for a in angles:
            O.sweep_xyz(p.om, p.geom, p.sched, a, p.omega, p.weight, p.tau, p.STotal, p.Sigt, p.Psi[a], Psi1, p.PsiB[a], PhiSets[a], savePsi)
    Phi = np.zeros((m.ncornr, p.G))
    for a in range(p.NA):
        Phi = Phi + PhiSets[a]
    return Phi, stage, mrefs


# ---------------------------------------------------------------------------
# one mini-app time step on the oracle (what umt_b200/cycle.py drives through the C ABI)
# ---------------------------------------------------------------------------
def oracle_cycle_3d(p, dt, tr4floor):
    """radtr.F90 for the mini-app: PsiB = 0, PhiTotal = sum w psi, exit PsiB <- psi, EnergyRadBOC, 2 temperature
    iterations + the savePsi sweep, then rtedit's EnergyRadiation / TrMax / PowerEscape / EnergyCheck."""
    m = p.mesh
    V = p.geom["Volume"]
    p.PsiB[:] = 0.0
    phi0 = np.einsum("a,acg->cg", p.weight, p.Psi)
    for a in range(p.NA):
        for b, c in p.bdy[a]:
            p.PsiB[a, b - 1] = p.Psi[a, c - 1]
    e_boc = float((V[:, None] * phi0).sum()) / SPEED_LIGHT
    for _ in range(2):
        oracle_sweep_3d(p, False)
    phi = oracle_sweep_3d(p, True)
    erad_z = np.add.reduceat((V[:, None] * phi).sum(1), m.cOffSet)
    e_rad = float(erad_z.sum()) / SPEED_LIGHT
    trz = np.sqrt(np.sqrt(np.maximum(erad_z / (p.geom["VolumeZone"] * RAD_CONSTANT * SPEED_LIGHT), tr4floor)))
    esc = np.zeros(p.G)
    for a in range(p.NA):
        for b, c in p.bdy[a]:
            esc += p.weight[a] * float(p.geom["A_bdy"][b - 1] @ p.omega[a]) * p.Psi[a, c - 1]
    d_erad = e_rad - e_boc
    return dict(EnergyRadiation=e_rad, TrMax=float(trz.max()), PowerEscape=float(esc.sum()), RadPowerEscape=esc, trz=trz,
                EnergyRadBOC=e_boc, EnergyCheck=dt * (0.0 - float(esc.sum())) - d_erad, phi=phi)


def oracle_cycle_rz(p, dt, tr4floor):
    """The same time step in r-z: geometryFactor = 2 pi (Size_mod.F90:275), Volume is volume / 2 pi (volumeUCBrz.F90:127-135),
    the boundary edit carries the radius of the boundary element (BoundaryEdit.F90:123) and only weighted angles are tallied."""
    m = p.mesh
    V = p.geom["Volume"]
    gf = 2.0 * np.pi
    p.PsiB[:] = 0.0
    phi0 = np.einsum("a,acg->cg", p.weight, p.Psi)
    for a in range(p.NA):
        for b, c in p.bdy[a]:
            p.PsiB[a, b - 1] = p.Psi[a, c - 1]
    e_boc = gf * float((V[:, None] * phi0).sum()) / SPEED_LIGHT
    for _ in range(2):
        oracle_sweep_rz(p, False)
    phi = oracle_sweep_rz(p, True)
    erad_z = np.add.reduceat((V[:, None] * phi).sum(1), m.cOffSet)
    e_rad = gf * float(erad_z.sum()) / SPEED_LIGHT
    trz = np.sqrt(np.sqrt(np.maximum(erad_z / (p.geom["VolumeZone"] * RAD_CONSTANT * SPEED_LIGHT), tr4floor)))
    esc = np.zeros(p.G)
    for a in range(p.NA):
        if not p.weight[a] > 0.0:
            continue
        for b, c in p.bdy[a]:
            esc += p.weight[a] * gf * p.geom["RadiusB"][b - 1] * float(p.geom["A_bdy"][b - 1] @ p.omega[a]) * p.Psi[a, c - 1]
    d_erad = e_rad - e_boc
    return dict(EnergyRadiation=e_rad, TrMax=float(trz.max()), PowerEscape=float(esc.sum()), RadPowerEscape=esc, trz=trz,
                EnergyRadBOC=e_boc, EnergyCheck=dt * (0.0 - float(esc.sum())) - d_erad, phi=phi)


# ---------------------------------------------------------------------------
# multi-domain GTA (GTASweep.F90:66-76,139-146 exchange + GTASolver.F90 with its MPIAllReduce calls), lock step on every rank
# ---------------------------------------------------------------------------
def oracle_gta_exchange_lists(meshes, geoms, omega):
    """findexit.F90:102-294 on the GTA angle set (3-D, no reflecting boundaries: sets of one angle, so of every pair the
    higher rank classifies): lists[r][k][a] = (send, recv) positions a*nb+b (0-based) in the flattened (8, nbelem) PsiB."""
    out = []
    for r, m in enumerate(meshes):
        per_b = []
        for b in shared_boundaries(m):
            mq = meshes[b.neighbor]
            bq = [x for x in shared_boundaries(mq) if x.neighbor == r][0]
            mine = geoms[r]["A_bdy"][b.first_elem - 1:b.first_elem - 1 + b.n_elem] @ omega.T
            theirs = geoms[b.neighbor]["A_bdy"][bq.first_elem - 1:bq.first_elem - 1 + bq.n_elem] @ omega.T
            per_a = []
            for a in range(len(omega)):
                t = np.sign(mine[:, a]) if r > b.neighbor else -np.sign(theirs[:, a])
                el = a * m.nbelem + np.arange(b.first_elem - 1, b.first_elem - 1 + b.n_elem)
                per_a.append((el[t > 0], el[t < 0]))
            per_b.append(per_a)
        out.append(per_b)
    return out


def oracle_gta_multi_solve(meshes, Ps, lists, Phis, geoms, epsPoint=1e-6, maxIters=21, epsGrey=0.1):
    """GTASolver.F90:42-425 on N domains: Ps[r] is rank r's oracle GtaProblem (consumed), Phis[r] its PhiTotal (nc, ngr).
    Returns (GreyCorrection per rank, nGreyIter, maxRelErrGrey).  With N = 1 this is orc_gta_solver restated in numpy."""
    N = len(meshes)
    nA = 8

    def exchange(B):   # SendFlux/RecvFlux of every angle: exiting elements -> the neighbour's incident elements (lagged)
        snap = [b.copy() for b in B]
        for r, m in enumerate(meshes):
            for k, sb in enumerate(shared_boundaries(m)):
                kq = [i for i, x in enumerate(shared_boundaries(meshes[sb.neighbor])) if x.neighbor == r][0]
                for a in range(nA):
                    recv, send = lists[r][k][a][1], lists[sb.neighbor][kq][a][0]
                    assert len(recv) == len(send)
                    B[r].reshape(-1)[recv] = snap[sb.neighbor].reshape(-1)[send]

    def grey_sweep(B, X, withSource):
        exchange(B)
        for r in range(N):
            Ps[r].grey_sweep(B[r], X[r], withSource)

    def prod1(X):
        return sum(float((X[r] * Ps[r].opac["GreySigScatVol"]).sum()) for r in range(N))

    def prod(X, Y):
        return sum(float((X[r] * Y[r] * Ps[r].opac["GreySigScatVol"]).sum()) for r in range(N))

    zone_of = [np.repeat(np.arange(m.nzones), m.numCorner) for m in meshes]
    vol = [g["Volume"] for g in geoms]
    volz = [g["VolumeZone"] for g in geoms]
    radE = [np.bincount(zone_of[r], weights=vol[r] * Phis[r].sum(axis=1), minlength=meshes[r].nzones) / volz[r] for r in range(N)]
    for r in range(N):
        Ps[r].init_tt()
    corr = [np.zeros(m.ncornr) for m in meshes]
    pzOld = [np.zeros(m.nzones) for m in meshes]
    R = [np.zeros(m.ncornr) for m in meshes]
    RB = [np.zeros((nA, m.nbelem)) for m in meshes]
    n = 1
    grey_sweep(RB, R, True)
    D, DB = [x.copy() for x in R], [x.copy() for x in RB]
    rrOld = prod1(R)
    for r in range(N):
        Ps[r].GreySource[:] = 0.0
    err = 0.0
    while True:
        if abs(rrOld) < 1e-150:
            if n <= 2:
                corr = [x.copy() for x in R]
            break
        n += 2
        A, AB = [x.copy() for x in D], [x.copy() for x in DB]
        grey_sweep(AB, A, False)
        A = [D[r] - A[r] for r in range(N)]
        AB = [DB[r] - AB[r] for r in range(N)]
        dAd = prod1(A)
        if abs(dAd) < 1e-150:
            break
        alpha = rrOld / dAd
        R = [R[r] - alpha * A[r] for r in range(N)]
        RB = [RB[r] - alpha * AB[r] for r in range(N)]
        AS, ASB = [x.copy() for x in R], [x.copy() for x in RB]
        grey_sweep(ASB, AS, False)
        AS = [R[r] - AS[r] for r in range(N)]
        ASB = [RB[r] - ASB[r] for r in range(N)]
        oNum, oDen = prod(AS, R), prod(AS, AS)
        if abs(oDen) < 1e-150 or abs(oNum) < 1e-150:
            corr = [corr[r] + alpha * D[r] for r in range(N)]
            break
        om = oNum / oDen
        corr = [corr[r] + alpha * D[r] + om * R[r] for r in range(N)]
        R = [R[r] - om * AS[r] for r in range(N)]
        RB = [RB[r] - om * ASB[r] for r in range(N)]
        rr = prod1(R)
        beta = (rr * alpha) / (rrOld * om)
        D = [R[r] + beta * (D[r] - om * A[r]) for r in range(N)]
        DB = [RB[r] + beta * (DB[r] - om * AB[r]) for r in range(N)]
        err = 0.0
        for r in range(N):   # local error norms, then MPIAllReduce(max)
            pz = np.bincount(zone_of[r], weights=vol[r] * corr[r], minlength=meshes[r].nzones) / volz[r]
            ez = pz - pzOld[r]
            phiNew = radE[r] + pz
            errL2, phiL2 = float((volz[r] * ez * ez).sum()), float((volz[r] * phiNew * phiNew).sum())
            nzm = phiNew != 0.0
            relPoint = float(np.abs(ez[nzm] / phiNew[nzm]).max()) if nzm.any() else 0.0
            relL2 = np.sqrt(abs(errL2 / phiL2)) if phiL2 != 0.0 else 0.0
            err = max(err, relPoint, relL2)
            pzOld[r] = pz
        if (err < epsPoint or n >= maxIters) and err < epsGrey:
            break
        if n >= 100 * maxIters:
            raise RuntimeError("oracle_gta_multi_solve: not converging")
        rrOld = rr
    return corr, n, err


# ---------------------------------------------------------------------------
# SweepScheduler (rt/SweepScheduler.F90:32-313) and the per-step exchange of SetSweep.F90:113-170, lock step on every rank
# ---------------------------------------------------------------------------
def oracle_net_flux(problems, lists):
    """setNetFlux.F90:61-141: NetFlux[r][k, a] = my exit current of angle a through shared boundary k minus the neighbour's."""
    NA = problems[0].NA
    ex = []
    for r, p in enumerate(problems):
        per_b = []
        for k, b in enumerate(shared_boundaries(p.mesh)):
            e = np.zeros(NA)
            for a in range(NA):
                el = lists[r][k][a][0]
                dot = p.geom["A_bdy"][el - 1] @ p.omega[a]
                e[a] = p.weight[a] * float((dot * p.PsiB[a, el - 1].sum(axis=1)).sum())
            per_b.append(e)
        ex.append(per_b)
    out = []
    for r, p in enumerate(problems):
        nf = []
        for k, b in enumerate(shared_boundaries(p.mesh)):
            kq = [i for i, x in enumerate(shared_boundaries(problems[b.neighbor].mesh)) if x.neighbor == r][0]
            nf.append(ex[r][k] - ex[b.neighbor][kq])
        out.append(np.array(nf).reshape(len(nf), NA))
    return out


def oracle_sweep_scheduler(problems, nCommSets, netflux, mrefs=None, nBins=None):
    """Every rank's CSet%AngleOrder (comm sets concatenated, 0-based angles) and RecvOrder[k] per shared boundary.
    netflux[r] is (nShared_r, NA); mrefs[r][n][a] the mirror angle of a on reflecting boundary n (-1: not incident).
    nBins: schedule that many angle bins instead of the angles themselves (r-z: the xi-levels; no reflecting boundaries then)."""
    N, NA = len(problems), (nBins if nBins is not None else problems[0].NA)
    bps = NA // nCommSets
    depend = [netflux[r].sum(axis=0) if len(netflux[r]) else np.zeros(NA) for r in range(N)]
    depend = [d.copy() for d in depend]
    nRefl = [np.zeros(NA, int) for _ in range(N)]
    depAngle = []
    for r in range(N):
        mr = mrefs[r] if mrefs is not None else []
        da = -np.ones((max(len(mr), 1), NA), int)
        for n, m in enumerate(mr):
            for a in range(NA):
                if m[a] >= 0:
                    da[n, m[a]] = a
                    nRefl[r][a] += 1
        depAngle.append(da)
    notDone = [np.ones(NA, bool) for _ in range(N)]
    order = [np.zeros(NA, int) for _ in range(N)]
    nbrs = [shared_boundaries(p.mesh) for p in problems]
    recv = [[np.zeros(NA, int) for _ in nbrs[r]] for r in range(N)]
    for i in range(bps):
        new = []
        for r in range(N):
            nb = []
            for c in range(nCommSets):
                bins = [b for b in range(c * bps, (c + 1) * bps) if notDone[r][b]]
                imin = min(bins, key=lambda b: (nRefl[r][b], b))
                if nRefl[r][imin] != 0:
                    nRefl[r][imin] = 0
                ready = [b for b in bins if nRefl[r][b] == 0]
                best = ready[0]
                for b in ready[1:]:
                    if depend[r][b] > depend[r][best]:
                        best = b
                nb.append(best)
            new.append(nb)
        for r in range(N):
            for c in range(nCommSets):
                b = new[r][c]
                order[r][c * bps + i] = b
                for n in range(depAngle[r].shape[0]):
                    aRef = depAngle[r][n, b]
                    if aRef >= 0 and notDone[r][aRef]:
                        nRefl[r][aRef] -= 1
                notDone[r][b] = False
        for r in range(N):
            for k, sb in enumerate(nbrs[r]):
                for c in range(nCommSets):
                    b = new[sb.neighbor][c]
                    recv[r][k][c * bps + i] = b
                    if notDone[r][b]:
                        depend[r][b] -= netflux[r][k, b]
    return order, recv


def oracle_multi_sweep_ordered(problems, lists, nCommSets, order, savePsi, maxFluxIters=1, fluxTol=1e-6):
    """SetSweep.F90 with comm sets of several angles (3-D, no mesh cycles): at step i every comm set of every rank first sends
    the neighbours the rows of the angles *they* sweep at step i (as they are now), then receives, then sweeps its own."""
    N, NA = len(problems), problems[0].NA
    bps = NA // nCommSets
    nbrs = [shared_boundaries(p.mesh) for p in problems]

    def exit_currents():
        res = []
        for r, p in enumerate(problems):
            per_b = []
            for k, b in enumerate(nbrs[r]):
                ex = np.zeros(NA)
                for a in range(NA):
                    el = lists[r][k][a][0]
                    dot = p.geom["A_bdy"][el - 1] @ p.omega[a]
                    ex[a] = p.weight[a] * float((dot * p.PsiB[a, el - 1].sum(axis=1)).sum())
                per_b.append(ex)
            res.append(per_b)
        return res

    def incident(ex):
        inc = []
        for r, p in enumerate(problems):
            v = np.zeros(NA)
            for b in nbrs[r]:
                kq = [i for i, x in enumerate(nbrs[b.neighbor]) if x.neighbor == r][0]
                v += ex[b.neighbor][kq]
            inc.append(v)
        return inc

    inc = incident(exit_currents())
    it = 0
    while True:
        it += 1
        PhiSets = [np.zeros((NA, p.mesh.ncornr, p.G)) for p in problems]
        Psi1 = [np.zeros((p.mesh.ncornr + p.mesh.nbelem, p.G)) for p in problems]
        for i in range(bps):
            snap = [p.PsiB.copy() for p in problems]
            for r, p in enumerate(problems):
                for k, b in enumerate(nbrs[r]):
                    kq = [j for j, x in enumerate(nbrs[b.neighbor]) if x.neighbor == r][0]
                    for c in range(nCommSets):
                        a = order[r][c * bps + i]
                        rcv, snd = lists[r][k][a][1], lists[b.neighbor][kq][a][0]
                        assert len(rcv) == len(snd)
                        p.PsiB[a, rcv - 1] = snap[b.neighbor][a, snd - 1]
            for r, p in enumerate(problems):
                assert p.sched["totalCycles"] == 0
                for c in range(nCommSets):
                    a = order[r][c * bps + i]
                    O.sweep_xyz(p.om, p.geom, p.sched, a, p.omega, p.weight, p.tau, p.STotal, p.Sigt, p.Psi[a], Psi1[r], p.PsiB[a], PhiSets[r][a], savePsi)
        phis = []
        for r, p in enumerate(problems):
            Phi = np.zeros((p.mesh.ncornr, p.G))
            for a in range(NA):
                Phi = Phi + PhiSets[r][a]
            phis.append(Phi)
        inc_old, inc = inc, incident(exit_currents())
        if savePsi:
            break
        notconv = 0
        for r in range(N):   # testFluxConv.F90:55-105 per comm set
            for c in range(nCommSets):
                bins = range(c * bps, (c + 1) * bps)
                total = sum(inc[r][b] for b in bins)
                conv = True
                for b in bins:
                    rel = 0.0
                    if total != 0.0 and inc[r][b] / total > 0.001:
                        rel = abs(inc[r][b] - inc_old[r][b]) / inc[r][b]
                    conv = conv and rel <= fluxTol
                notconv += not conv
        if notconv == 0 or it >= maxFluxIters:
            break
    return phis, it, inc


def rz_reflect_stages(p, mrefs):
    """r-z stages as csrc/reflect.cu assigns them: xi-levels ordered by the mirror dependencies between levels (same DFS, back edges
    lagged), every angle of a level its own step."""
    lev = np.asarray(p.q["level"]) - 1
    nL = int(lev.max()) + 1
    ldeps = [[] for _ in range(nL)]
    for m in mrefs:
        for a in range(p.NA):
            if m[a] >= 0 and lev[m[a]] != lev[a]:
                ldeps[lev[a]].append(int(lev[m[a]]))
    lstage, state = [0] * nL, [0] * nL
    for root in range(nL):
        if state[root]:
            continue
        st = [[root, 0]]
        state[root] = 1
        while st:
            a, i = st[-1]
            if i < len(ldeps[a]):
                st[-1][1] += 1
                m = ldeps[a][i]
                if state[m] == 0:
                    state[m] = 1
                    st.append([m, 0])
                elif state[m] == 2:
                    lstage[a] = max(lstage[a], lstage[m] + 1)
            else:
                state[a] = 2
                st.pop()
                if st:
                    lstage[st[-1][0]] = max(lstage[st[-1][0]], lstage[a] + 1)
    cnt = [0] * nL
    pos = []
    for a in range(p.NA):
        pos.append(cnt[lev[a]])
        cnt[lev[a]] += 1
    maxPos = max(cnt)
    return [lstage[lev[a]] * maxPos + pos[a] for a in range(p.NA)]


def oracle_sweep_rz_reflecting(p, savePsi):
    """SetSweep angle loop in r-z with snreflect right before every swept angle (SetSweep.F90:139-143); levels advance together
    in the stage order of rz_reflect_stages, PsiM kept per xi-level."""
    m = p.mesh
    mrefs = oracle_reflected_angles(p)
    stage = rz_reflect_stages(p, mrefs)
    refl = reflecting_boundaries(m)
    lev = np.asarray(p.q["level"]) - 1
    Phi = np.zeros((m.ncornr, p.G))
    PsiM = {int(l): np.zeros((m.ncornr, p.G)) for l in set(lev.tolist())}
    Psi1 = np.zeros((m.ncornr + m.nbelem, p.G))
    for s in range(max(stage) + 1):
        for a in [a for a in range(p.NA) if stage[a] == s]:
            if p.q["finish"][a]:
                continue
            for k, b in enumerate(refl):
                if mrefs[k][a] >= 0:
                    sl = slice(b.first_elem - 1, b.first_elem - 1 + b.n_elem)
                    p.PsiB[a, sl] = p.PsiB[mrefs[k][a], sl]
            O.sweep_rz(p.om, p.geom, p.sched, a, p.q, p.tau, p.STotal, p.Sigt, p.Psi, Psi1, PsiM[int(lev[a])], p.PsiB, Phi, p.bdy[a], savePsi)
    return Phi, stage, mrefs


# ---------------------------------------------------------------------------
# r-z: the scheduler's angle bins are the xi-levels (SweepScheduler.F90:110-117)
# ---------------------------------------------------------------------------
def rz_bins(p):
    lev = np.asarray(p.q["level"]) - 1
    nBins = int(lev.max()) + 1
    return lev, [np.flatnonzero(lev == b) for b in range(nBins)]


def oracle_net_flux_bins(problems, lists):
    """setNetFlux per shared boundary and xi-level: the per-angle net flux summed over the angles of each bin"""
    lev, angles = rz_bins(problems[0])
    per_angle = oracle_net_flux(problems, lists)
    return [np.stack([nf[:, a].sum(axis=1) for a in angles], axis=1) if len(nf) else np.zeros((0, len(angles))) for nf in per_angle]


def oracle_multi_sweep_ordered_rz(problems, lists, nCommSets, binOrder, savePsi, maxFluxIters=1, fluxTol=1e-6):
    """SetSweep.F90 in r-z with comm sets of several xi-levels: at step i every comm set of every rank first sends the neighbours the
    rows of the levels *they* sweep at step i (as they are now), receives its own, then sweeps the angles of its level in order."""
    N, NA = len(problems), problems[0].NA
    lev, angles = rz_bins(problems[0])
    nBins = len(angles)
    bps = nBins // nCommSets
    nbrs = [shared_boundaries(p.mesh) for p in problems]

    def exit_currents():
        res = []
        for r, p in enumerate(problems):
            per_b = []
            for k, b in enumerate(nbrs[r]):
                ex = np.zeros(NA)
                for a in range(NA):
                    el = lists[r][k][a][0]
                    dot = p.geom["A_bdy"][el - 1] @ p.omega[a]
                    ex[a] = p.weight[a] * float((dot * p.PsiB[a, el - 1].sum(axis=1)).sum())
                per_b.append(ex)
            res.append(per_b)
        return res

    def incident(ex):
        inc = []
        for r, p in enumerate(problems):
            v = np.zeros(nBins)
            for b in nbrs[r]:
                kq = [i for i, x in enumerate(nbrs[b.neighbor]) if x.neighbor == r][0]
                np.add.at(v, lev, ex[b.neighbor][kq])
            inc.append(v)
        return inc

    inc = incident(exit_currents())
    it = 0
    while True:
        it += 1
        Phi = [np.zeros((p.mesh.ncornr, p.G)) for p in problems]
        PsiM = [{b: np.zeros((p.mesh.ncornr, p.G)) for b in range(nBins)} for p in problems]
        Psi1 = [np.zeros((p.mesh.ncornr + p.mesh.nbelem, p.G)) for p in problems]
        for i in range(bps):
            snap = [p.PsiB.copy() for p in problems]
            for r, p in enumerate(problems):
                for k, b in enumerate(nbrs[r]):
                    kq = [j for j, x in enumerate(nbrs[b.neighbor]) if x.neighbor == r][0]
                    for c in range(nCommSets):
                        for a in angles[binOrder[r][c * bps + i]]:
                            rcv, snd = lists[r][k][a][1], lists[b.neighbor][kq][a][0]
                            assert len(rcv) == len(snd)
                            p.PsiB[a, rcv - 1] = snap[b.neighbor][a, snd - 1]
            for r, p in enumerate(problems):
                for c in range(nCommSets):
                    bn = binOrder[r][c * bps + i]
                    for a in angles[bn]:
                        if p.q["finish"][a]:
                            continue
                        O.sweep_rz(p.om, p.geom, p.sched, a, p.q, p.tau, p.STotal, p.Sigt, p.Psi, Psi1[r], PsiM[r][bn], p.PsiB, Phi[r], p.bdy[a], savePsi)
        inc_old, inc = inc, incident(exit_currents())
        if savePsi:
            break
        notconv = 0
        for r in range(N):   # testFluxConv.F90:55-105 per comm set
            for c in range(nCommSets):
                bins = range(c * bps, (c + 1) * bps)
                total = sum(inc[r][b] for b in bins)
                conv = True
                for b in bins:
                    rel = 0.0
                    if total != 0.0 and inc[r][b] / total > 0.001:
                        rel = abs(inc[r][b] - inc_old[r][b]) / inc[r][b]
                    conv = conv and rel <= fluxTol
                notconv += not conv
        if notconv == 0 or it >= maxFluxIters:
            break
    return Phi, it, inc
