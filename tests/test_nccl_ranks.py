"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the same lock-step parity as test_gpu_exchange.py
but with one process per GPU and ncclSend/ncclRecv moving the psib rows.  Run by hand with
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/test_nccl_ranks.py"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_main():
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from tests import common as T
    from umt_b200 import mesh as M
    from umt_b200 import teton
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = (2, 2, 2)
    problems = [T.make_problem_3d(M.tiled_mesh(dims, rank=r, size=world), 1, 2, 8, seed=100 + r) for r in range(world)]
    p = problems[rank]
    ctx = T.gpu_context_3d(p, device=local)
    for b in T.shared_boundaries(p.mesh):
        ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(teton.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.set_comm(rank, world, idt.cpu().numpy().tobytes())
    ctx.build_exchange()
    lists = T.oracle_exchange_lists(problems)
    worst = 0.0
    for save, iters in ((False, 1), (False, 4), (True, 3)):
        phis, it_ref, inc_ref = T.oracle_multi_sweep_3d(problems, lists, save, iters, 1e-6)
        it = ctx.sweep(save, iters, 1e-6)
        assert it == it_ref, (it, it_ref)
        e = T.relerr(ctx.download_phi(), phis[rank])
        worst = max(worst, e)
        assert e <= 1e-12, e
        assert T.mixed_err(ctx.download_psib(), p.PsiB, 1e-12) <= 1.0
    print(f"rank {rank}: NCCL psib exchange parity ok, worst phi rel err {worst:.2e}", flush=True)
    ctx.close()
    # grey transport acceleration on the decomposed mesh: grey psib exchange + allreduce of the inner products over NCCL
    from tests.test_gta_multidomain import _domain, _oracle_problem
    G = 4
    doms = [_domain(M.tiled_mesh((2, 2, 1), rank=r, size=world), G, 40 + r) for r in range(world)]
    d = doms[rank]
    mesh, g = d["mesh"], d["g"]
    ctx = teton.SweepContext.from_mesh(mesh, G, local)
    ctx.set_geometry(g["Volume"], g["A_fp"], g["A_ez"], A_bdy=g["A_bdy"])
    ctx.build_product_quadrature(1, 1, 1)
    ctx.upload_state(np.tile(d["Phi"] / (4 * np.pi), (8, 1, 1)), None, np.full((mesh.nzones, G), d["tau"]), np.zeros((mesh.ncornr, G)), d["tau"])
    ctx.init_phi_total()
    for b in T.shared_boundaries(mesh):
        ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(teton.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.set_comm(rank, world, idt.cpu().numpy().tobytes())
    ctx.gta_setup()
    ctx.gta_compute_opacity(d["Siga"], d["Sigs"], d["Eta"], d["Chi"].copy())
    ctx.collision_rate(d["Eta"], d["Siga"], d["Sigs"], 0)
    lists = T.oracle_gta_exchange_lists([x["mesh"] for x in doms], [x["g"] for x in doms], doms[0]["omega"])
    corr, n, err = T.oracle_gta_multi_solve([x["mesh"] for x in doms], [_oracle_problem(x) for x in doms], lists, [x["Phi"] for x in doms],
                                            [x["g"] for x in doms])
    corr_d, n_d, err_d = ctx.gta_solve()
    scale = max(np.abs(c).max() for c in corr)
    assert n_d == n and n > 3, (n_d, n)
    assert np.abs(corr_d - corr[rank]).max() <= 1e-8 * scale
    print(f"rank {rank}: NCCL GTA solve ok, {n_d} grey sweeps, err {np.abs(corr_d - corr[rank]).max() / scale:.2e}", flush=True)
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("put", ["0", "1"])
def test_two_ranks_over_nccl(put):
    """put = 0: the psib rows travel by ncclSend/ncclRecv (the default between ranks); put = 1: the pack kernel stores them into the
    neighbour's receive buffer over CUDA IPC peer memory (opt-in, UMT_EXCHANGE_PUT=1)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, UMT_EXCHANGE_PUT=put)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.abspath(__file__)], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("parity ok") == 2 and r.stdout.count("NCCL GTA solve ok") == 2


if __name__ == "__main__":
    _rank_main()
