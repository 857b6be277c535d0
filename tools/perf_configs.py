"""Device times of the other BASELINE configs (not the bench line): configs[1] r-z tiled mesh -d 40,40,0 -G 64 on 4 domains 2x2
(run under torchrun with 4 ranks, NCCL psib exchange) and configs[3] unstructured box -R 6 -G 64 -P 2 -A 2 on one GPU.
  python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/perf_configs.py rz
  python tools/perf_configs.py box"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from umt_b200 import mesh as M, problem as PR, teton

what = sys.argv[1] if len(sys.argv) > 1 else "box"
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G = 64
if what == "rz":
    mesh = M.tiled_mesh((40, 40, 0), rank=rank, size=world)
    nd = 2
else:
    mesh = M.unstruct_box_mesh(6)
    nd = 3
ctx = teton.SweepContext.from_mesh(mesh, G, device=local)
ctx.compute_geometry(mesh.px)
NA = ctx.build_product_quadrature(2, 2, 1)
for b in mesh.boundaries:
    if b.bc_type == M.BC_SHARED:
        ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
if world > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(teton.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.set_comm(rank, world, bytes(idt.cpu().numpy().tobytes()))
ctx.build_schedule()
tau = PR.tau()
ctx.upload_state(None, None, np.full((mesh.nzones, G), tau), np.zeros((mesh.ncornr, G)), tau)
ctx.init_teton(np.full(mesh.nzones, PR.TR0), PR.group_bounds(G), PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(nd), 0.0)
ctx.init_radiation_field()
ts = []
for i in range(6):
    if world > 1:
        dist.barrier()
    ctx.sweep(False, 1)
    ts.append(ctx.last_times())
best = min(ts[1:], key=lambda t: t["total_ms"])
t = torch.tensor([best["total_ms"]], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
unknowns = mesh.ncornr * NA * G * world
if rank == 0:
    print("%s: %d domain(s), zones/domain %d, G=%d, angles %d: %.3f ms per sweep (max over ranks; sweep %.3f phi %.3f exchange %.3f on rank 0) -> %.3e unknowns/s"
          % ("configs[1] r-z -d 40,40,0" if what == "rz" else "configs[3] unstructured box -R 6", world, mesh.nzones, G, NA, float(t.item()), best["sweep_ms"], best["phi_ms"],
             best["exchange_ms"], unknowns / float(t.item()) * 1e3), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
