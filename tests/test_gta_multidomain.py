"""Multi-domain grey transport acceleration: the psib exchange of GTASweep.F90:66-76,139-146 (every angle's exiting grey
boundary fluxes to the neighbour's incident elements, lagged one grey sweep) and the MPIAllReduce calls of GTASolver.F90 /
scat_prod.F90.  The oracle is tests/common.py::oracle_gta_multi_solve: the per-domain C restatement of GreySweepNEW driven in
lock step on every rank."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T
from umt_b200 import mesh as M
from umt_b200 import problem as PR


def _domain(mesh, G, seed, scat=20.0):
    om = O.OMesh(mesh)
    g = O.geometry(om)
    omega, w = O.gta_quad_xyz()
    sched = O.schedule(om, g, omega)
    rng = np.random.default_rng(seed)
    nz, nc = mesh.nzones, mesh.ncornr
    d = dict(mesh=mesh, om=om, g=g, omega=omega, w=w, sched=sched, tau=PR.tau(1e-3), Siga=5 * rng.random((nz, G)), Sigs=scat * rng.random((nz, G)),
             Eta=0.5 * rng.random(nc), Phi=rng.random((nc, G)))
    chi = rng.random((nc, G))
    d["Chi"] = chi / chi.sum(1, keepdims=True)
    return d


def _oracle_problem(d):
    op = O.gta_set_opacity(d["om"], d["g"], d["tau"], d["Siga"], d["Sigs"], d["Eta"], d["Chi"].copy())
    gs = O.collision_rate(d["om"], d["Eta"], d["Siga"], d["Sigs"], d["Phi"], np.zeros(d["mesh"].ncornr), 0)
    return O.GtaProblem(d["om"], d["g"], d["sched"], d["omega"], d["w"], op, gs, PR.wtiso(3))


def test_lockstep_oracle_equals_c_solver_on_one_domain():
    d = _domain(M.tiled_mesh((2, 2, 2)), 4, 7)
    c1, n1, e1 = _oracle_problem(d).solve(d["Phi"])
    c2, n2, e2 = T.oracle_gta_multi_solve([d["mesh"]], [_oracle_problem(d)], [[]], [d["Phi"]], [d["g"]])
    assert n1 == n2 and n1 > 3
    assert np.abs(c1 - c2[0]).max() <= 1e-13 * np.abs(c1).max() and abs(e1 - e2) <= 1e-9 * e1


def test_two_domains_converge_to_the_single_domain_correction():
    """Same physical problem (uniform data) on one domain and split in two: the converged grey corrections agree in their
    volume integral to the solver tolerance, i.e. the boundary unknowns carried by the exchange close the coupling."""
    def uniform(mesh):
        d = _domain(mesh, 2, 1)
        d["Siga"][:] = 1.0; d["Sigs"][:] = 30.0; d["Eta"][:] = 0.2; d["Phi"][:] = 1.0; d["Chi"][:] = 0.5
        return d
    one = uniform(M.tiled_mesh((2, 2, 2)))
    c1, n1, _ = T.oracle_gta_multi_solve([one["mesh"]], [_oracle_problem(one)], [[]], [one["Phi"]], [one["g"]], epsPoint=1e-10, maxIters=200)
    two = [uniform(M.tiled_mesh((2, 2, 1), rank=r, size=2)) for r in range(2)]
    lists = T.oracle_gta_exchange_lists([d["mesh"] for d in two], [d["g"] for d in two], two[0]["omega"])
    c2, n2, _ = T.oracle_gta_multi_solve([d["mesh"] for d in two], [_oracle_problem(d) for d in two], lists, [d["Phi"] for d in two],
                                         [d["g"] for d in two], epsPoint=1e-10, maxIters=200)
    i1 = float((one["g"]["Volume"] * c1[0]).sum())
    i2 = sum(float((d["g"]["Volume"] * c).sum()) for d, c in zip(two, c2))
    assert abs(i1 - i2) <= 1e-6 * abs(i1)


@pytest.mark.gpu
@pytest.mark.parametrize("N,dims", [(2, (2, 2, 1)), (4, (2, 1, 1))])
def test_gta_solve_on_decomposed_mesh_matches_oracle(N, dims):
    from umt_b200 import teton
    G = 4
    doms = [_domain(M.tiled_mesh(dims, rank=r, size=N), G, 40 + r) for r in range(N)]
    ctxs = []
    for d in doms:
        mesh, g = d["mesh"], d["g"]
        ctx = teton.SweepContext.from_mesh(mesh, G)
        ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], A_bdy=g["A_bdy"])
        ctx.build_product_quadrature(1, 1, 1)
        ctx.upload_state(np.tile(d["Phi"] / (4 * np.pi), (8, 1, 1)), None, np.full((mesh.nzones, G), d["tau"]), np.zeros((mesh.ncornr, G)), d["tau"])
        ctx.init_phi_total()
        for b in T.shared_boundaries(mesh):
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
        ctxs.append(ctx)
    teton.connect_local(ctxs)
    T.run_local_group(ctxs, lambda r, c: c.gta_setup())
    for d, ctx in zip(doms, ctxs):
        ctx.gta_compute_opacity(d["Siga"], d["Sigs"], d["Eta"], d["Chi"].copy())
        ctx.collision_rate(d["Eta"], d["Siga"], d["Sigs"], 0)
    lists = T.oracle_gta_exchange_lists([d["mesh"] for d in doms], [d["g"] for d in doms], doms[0]["omega"])
    corr, n, err = T.oracle_gta_multi_solve([d["mesh"] for d in doms], [_oracle_problem(d) for d in doms], lists, [d["Phi"] for d in doms],
                                            [d["g"] for d in doms])
    res = T.run_local_group(ctxs, lambda r, c: c.gta_solve())
    scale = max(np.abs(c).max() for c in corr)
    for r in range(N):
        corr_d, n_d, err_d = res[r]
        assert n_d == n and n > 3
        assert np.abs(corr_d - corr[r]).max() <= 1e-8 * scale      # Krylov amplification of rounding differences, as on one domain
        assert abs(err_d - err) <= 1e-5 * max(err, 1e-30) + 1e-12
    for c in ctxs:
        c.close()
