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
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_ranks_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.abspath(__file__)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("parity ok") == 2


if __name__ == "__main__":
    _rank_main()
