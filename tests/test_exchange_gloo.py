"""CPU, world_size 2 and 4 over gloo: the host side of the multi-domain path.  Each rank holds one domain in a
host-only context, classifies its half of the shared-boundary angles (findexit.F90:128-160), trades the
incident tests with its neighbours through torch.distributed (the caller-supplied transport of the C ABI),
builds ListSend/ListRecv, and rank 0 checks every rank's lists against the oracle's and that each send list
has a receive list of the same length on the other side."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dims, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from tests import common as T
    from umt_b200 import mesh as M
    from umt_b200 import teton
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        m = M.tiled_mesh(dims, rank=rank, size=world)
        g = O.geometry(O.OMesh(m))
        ctx = teton.SweepContext.from_mesh(m, 2, device=-1)
        ctx.set_rank(rank, world)
        ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], A_bdy=g["A_bdy"])
        ctx.build_product_quadrature(1, 2, 1)
        shared = T.shared_boundaries(m)
        for b in shared:
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
        mine = [torch.from_numpy(ctx.get_incident_test(k, b.n_elem)) for k, b in enumerate(shared)]
        theirs = [torch.zeros_like(t) for t in mine]
        reqs = []
        for k, b in enumerate(shared):
            reqs.append(dist.isend(mine[k], b.neighbor, tag=min(rank, b.neighbor) * world + max(rank, b.neighbor)))
            reqs.append(dist.irecv(theirs[k], b.neighbor, tag=min(rank, b.neighbor) * world + max(rank, b.neighbor)))
        for q in reqs:
            q.wait()
        for k in range(len(shared)):
            ctx.set_incident_test(k, theirs[k].numpy())
        ctx.build_exchange()
        ctx.build_schedule()   # exit lists with shared elements last must still build
        mylists = [[tuple(x.tolist() for x in ctx.exchange_lists(k, a + 1)) for a in range(ctx.NA)] for k in range(len(shared))]
        gathered = [None] * world
        dist.all_gather_object(gathered, mylists)
        ok = True
        if rank == 0:
            problems = []
            for r in range(world):
                p = T.Problem()
                p.mesh = M.tiled_mesh(dims, rank=r, size=world)
                p.geom = O.geometry(O.OMesh(p.mesh))
                p.omega, p.weight = O.quad_xyz(1, 2)
                p.NA = len(p.weight)
                problems.append(p)
            ref = T.oracle_exchange_lists(problems)
            for r in range(world):
                for k, b in enumerate(T.shared_boundaries(problems[r].mesh)):
                    kq = [i for i, x in enumerate(T.shared_boundaries(problems[b.neighbor].mesh)) if x.neighbor == r][0]
                    for a in range(problems[r].NA):
                        ls, lr = gathered[r][k][a]
                        ok &= ls == ref[r][k][a][0].tolist() and lr == ref[r][k][a][1].tolist()
                        ok &= len(ls) == len(gathered[b.neighbor][kq][a][1])
                        ok &= len(ls) + len(lr) == b.n_elem   # planar faces: nothing grazing
        flag = torch.tensor([1 if ok else 0])
        dist.broadcast(flag, 0)
        ret[rank] = int(flag.item())
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims", [(2, (2, 2, 2)), (4, (1, 2, 2))])
def test_exchange_lists_over_gloo(world, dims):
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    mgr = ctxmp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, dims, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert [ret.get(r) for r in range(world)] == [1] * world
