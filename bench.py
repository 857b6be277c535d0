#!/usr/bin/env python
"""bench.py — Sn sweep throughput (unknowns/s = corners x angles x groups per second
of one full ControlSweep: all angles, psi->phi reduction, psib exchange; schedule
construction excluded) on N B200s, one mesh domain per GPU (weak scaling).

  python bench.py --gpus 1 --steps K --warmup W
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the CPU path timed on the host cores

Workload at N=1: BASELINE.json configs[4] per-domain size (3-D tiled mesh -d 20,20,20,
-G 128, default product quadrature P2 A2 = 32 angles, 6.29e9 unknowns).  configs[2]
(-P 4 -A 4, 128 angles) needs 201 GB for Psi alone and does not fit one B200.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from umt_b200 import mesh as M          # noqa: E402
from umt_b200 import problem as PR      # noqa: E402


def algorithmic_bytes_per_unknown(G, ndim=3, final=False):
    """41 + 180/G (3-D, non-final sweep), 49 + 180/G with savePsi; RZ 58 + 128/G."""
    if ndim == 3:
        return (49.0 if final else 41.0) + 180.0 / G
    return 58.0 + 128.0 / G


def sweep_kernel_bytes_per_unknown(G, ndim=3):
    """Bytes the sweep kernel itself must move: read STotal 8 + read Psi^n 8 +
    write Psi1 8 + Sigt 8/cpz + geometry/G.  The phi tally is a separate streaming kernel."""
    return (25.0 + 180.0 / G) if ndim == 3 else (42.0 + 128.0 / G)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            if self._stop.is_set():
                break
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass

    def stop(self):
        self._stop.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def profiled_traffic_per_unknown():
    """dram__bytes_read.sum + dram__bytes_write.sum of the 3-D sweep kernel divided by the unknowns of that launch, from the committed
    ncu capture of this very workload and kernel build (profiles/r02_sweep3d_d20_G128_ncu_summary.txt: `ncu --set full` of
    `python bench.py --steps 1 --warmup 1`, i.e. -d 20,20,20 -G 128 P2 A2, 6.29e9 unknowns per launch; tools/gpu_evidence.sh)."""
    p = os.path.join(ROOT, "profiles", "r02_sweep3d_d20_G128_ncu_summary.txt")
    try:
        rd = wr = None
        for line in open(p):
            if "dram__bytes_read.sum [Gbyte]" in line:
                rd = float(line.split("=")[1])
            if "dram__bytes_write.sum [Gbyte]" in line:
                wr = float(line.split("=")[1])
        return (rd + wr) * 1e9 / (24 * 20 ** 3 * 8 * 32 * 128), os.path.relpath(p, ROOT)
    except Exception:
        return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference's CPU path; the reference itself needs
# Fortran + MPI + Conduit and cannot be built here) on a bounded sample of the workload
# ---------------------------------------------------------------------------
def cpu_sweep_rate(G, npolar, nazim, sample_dims, steps=1, warmup=0):
    from oracle import oracle as O
    from tests import common as T
    fast = O.use_fast_build()   # -O3 -march=native build of the restatement, compiled on this machine
    m = M.tiled_mesh(sample_dims)
    p = T.make_problem_3d(m, npolar, nazim, G, driver_like=True)
    unknowns = m.ncornr * p.NA * G
    cores = O.max_threads()
    for _ in range(warmup):
        T.oracle_sweep_3d(p, False, cores)
    t0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        T.oracle_sweep_3d(p, False, cores)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return unknowns / dt, cores, dt, f"3-D tiled mesh -d {sample_dims[0]},{sample_dims[1]},{sample_dims[2]} -G {G} P{npolar} A{nazim} ({unknowns:.3e} unknowns per sweep), same problem data, one full SetSweep+getPhiTotal, oracle built {'-O3 -march=native -fopenmp' if fast else '-O2 -fopenmp -ffp-contract=off'}"


def reference_cuda_rate(G, npolar, nazim, sample_dims):
    """The reference's own GPU offload of this path -- `gpu_sweepucbxyz` (T/gpu/GPU_SweepUCBxyz.cu, compiled unmodified for sm_100a
    by oracle/Makefile), called angle by angle with host arrays exactly as SetSweep_CUDA.F90 does (it uploads and downloads
    everything on every call) -- timed on a bounded sample of the workload.  A reported baseline next to cpu_baseline."""
    from oracle import oracle as O
    from tests import common as T
    if not O.ref_cuda_available():
        return None
    m = M.tiled_mesh(sample_dims)
    p = T.make_problem_3d(m, npolar, nazim, G, driver_like=True)
    nc, nb = m.ncornr, m.nbelem
    Psi1 = np.zeros((nc + nb, G)); Phi = np.zeros((nc, G))

    def one_pass():
        Phi[:] = 0.0
        for a in range(p.NA):
            O.ref_cuda_sweep_xyz(p.om, p.geom, p.sched, a, p.omega, p.weight, p.tau, p.STotal, p.Sigt, p.Psi[a], Psi1, p.PsiB[a],
                                 Phi, p.cyclePsi, False, stream_id=0)
    one_pass()   # allocates the shim's static buffers, warms up
    t0 = time.perf_counter()
    one_pass()
    dt = time.perf_counter() - t0
    unknowns = nc * p.NA * G
    return {"value": unknowns / dt, "unit": "unknowns/s", "kind": "reference (its CUDA sweep, unmodified, sm_100a build)",
            "ms_per_sweep": dt * 1e3,
            "sample": f"3-D tiled mesh -d {sample_dims[0]},{sample_dims[1]},{sample_dims[2]} -G {G} P{npolar} A{nazim} ({unknowns:.3e} unknowns), "
                      f"{p.NA} gpu_sweepucbxyz calls with host arrays (one stream), wall clock"}


def reference_sample_dims(args):
    """Tiles per side of the bounded CPU sample: the full -d 20 problem needs 50 GB of host memory per domain for Psi alone."""
    if args.cpu_dims:
        return args.cpu_dims
    return 10 if args.gpus == 1 else (8 if args.gpus <= 4 else 6)


def run_reference(args):
    """The CPU path on the host cores of this box: N = --gpus mesh domains (what `mpirun -n N` of the reference runs), swept
    concurrently, the host cores split evenly over the domains (OpenMP over angle sets inside each, SetSweep.F90:113-116).
    Under torchrun rank 0 alone does this (with every core of the box, whatever OMP_NUM_THREADS the launcher exported)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from tests import common as T
    N = max(args.gpus, 1)
    d = reference_sample_dims(args)
    G = args.groups
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_domain = max(1, cores // N)
    fast = O.use_fast_build()
    problems = [T.make_problem_3d(M.tiled_mesh((d, d, d), rank=r, size=N), args.polar, args.azimuthal, G, driver_like=True) for r in range(N)]
    unknowns = sum(p.mesh.ncornr * p.NA * G for p in problems)

    def one_step():
        th = [threading.Thread(target=T.oracle_sweep_3d, args=(p, False, per_domain)) for p in problems]
        for t in th:
            t.start()
        for t in th:
            t.join()
    for _ in range(min(args.warmup, 1)):
        one_step()
    t0 = time.perf_counter()
    for _ in range(max(args.steps, 1)):
        one_step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = unknowns / dt
    sample = (f"{N} domain(s) of a 3-D tiled mesh -d {d},{d},{d} -G {G} P{args.polar} A{args.azimuthal} ({unknowns:.3e} unknowns per step, "
              f"the bench workload's problem data at a reduced domain size: the -d {args.dims} domain needs {24 * args.dims ** 3 * 8 * 8 * args.polar * args.azimuthal * G * 8 / 1e9:.0f} GB of host memory for Psi), "
              f"domains swept concurrently with {per_domain} OpenMP thread(s) each, one full SetSweep+getPhiTotal per domain and step, no psib exchange "
              f"(lagged exchange cost not charged to the CPU arm), oracle built {'-O3 -march=native -fopenmp' if fast else '-O2 -fopenmp -ffp-contract=off'}")
    cfg = workload_config(args)
    cfg["reference_sample"] = sample
    cfg["reference_sample_dims"] = [d, d, d]
    line = {
        "impl": "reference", "metric": "Sn sweep unknowns/sec (corner x angle x group)", "value": val, "unit": "unknowns/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "unknowns/s", "cores": min(cores, per_domain * N), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "unknowns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (Fortran+MPI+Conduit) cannot be built in this image; this is the CPU restatement of SweepUCBxyz/SetSweep (oracle/), OpenMP over angle sets like SetSweep.F90:113-116; throughput is size-normalised (unknowns/s), the sample size is stated in config.reference_sample",
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    d = args.dims
    which = "BASELINE configs[4] per-domain" if (args.polar, args.azimuthal) == (2, 2) else "BASELINE configs[2] (-P 4 -A 4) at the largest domain that fits one B200" if (args.polar, args.azimuthal) == (4, 4) else "3-D tiled mesh"
    return {"workload": f"{which}: Blueprint 3D tiled mesh -B local -d {d},{d},{d} -G {args.groups} -P {args.polar} -A {args.azimuthal}, one domain per GPU, vacuum BCs, mini-app opacities (Sigt = 1/(c dt), STotal = 0), non-final sweep (savePsi = false)",
            "zones_per_domain": 24 * d * d * d, "groups": args.groups, "angles": 8 * args.polar * args.azimuthal,
            "l2_policy": "inputs (Psi, the Psi1 workspace, STotal, Phi: tens of GB) are far larger than the 126 MB L2; no flush needed",
            "parallelism": f"spatial domains x{args.gpus}, psib exchange lagged one flux pass (the transfer for the next pass overlaps the phi tally)"}


def bind_to_gpu_numa_node(torch, local):
    """One rank per GPU: run this process on the cores of the GPU's NUMA node (its staging buffers come from umt_host_alloc, which
    binds their pages to the same node).  Returns the node, or None when the platform gives no NUMA information."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ---------------------------------------------------------------------------
def _parity_domain(teton, r, N, device, G=8):
    """One small mesh domain (2 x 2 x 2 tiles of the N-domain tiled mesh) with seeded random state, built by the library alone."""
    mesh = M.tiled_mesh((2, 2, 2), rank=r, size=N)
    ctx = teton.SweepContext.from_mesh(mesh, G, device=device)
    ctx.compute_geometry(mesh.px)
    NA = ctx.build_product_quadrature(1, 2, 1)
    for b in mesh.boundaries:
        if b.bc_type == M.BC_SHARED:
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
    ctx.build_schedule()
    rng = np.random.default_rng(4242 + r)
    nz, nc, nb = mesh.nzones, mesh.ncornr, mesh.nbelem
    tau = 3.0
    ctx.upload_state(0.5 + rng.random((NA, nc, G)), 0.5 + rng.random((NA, nb, G)), tau + 20.0 * rng.random((nz, G)), rng.random((nc, G)), tau)
    return ctx


def nccl_parity_check(teton, torch, dist, rank, world, local):
    """Correctness of the multi-GPU path, checked on every rank before anything is timed: a small decomposed problem swept over the
    real NCCL communicator (psib rows by ncclSend/ncclRecv, convergence by ncclAllReduce) must give the PhiTotal / PsiB / Psi that the
    same library computes with all `world` domains inside one process (in-process transport, which tests/test_gpu_exchange.py pins to
    the lock-step oracle at 2, 4 and 8 domains).  Returns the worst relative difference over ranks (expected 0: same kernels, same data)."""
    ctx = _parity_domain(teton, rank, world, local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(teton.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.set_comm(rank, world, bytes(idt.cpu().numpy().tobytes()))
    ctx.build_exchange()
    local_ctxs = [_parity_domain(teton, r, world, local) for r in range(world)]
    teton.connect_local(local_ctxs)

    def group(fn):
        out, err = [None] * world, [None] * world

        def work(r):
            try:
                out[r] = fn(local_ctxs[r])
            except BaseException as e:   # noqa: BLE001
                err[r] = e
        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out
    group(lambda c: c.build_exchange())
    worst = 0.0
    for save, iters in ((False, 3), (True, 1)):
        it = ctx.sweep(save, iters, 1e-6)
        its = group(lambda c: c.sweep(save, iters, 1e-6))
        assert its[rank] == it, f"rank {rank}: {it} flux passes over NCCL, {its[rank]} in process"
        ref = local_ctxs[rank]
        for a, b in ((ctx.download_phi(), ref.download_phi()), (ctx.download_psib(), ref.download_psib()), (ctx.download_psi(), ref.download_psi())):
            worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))))
    ctx.close()
    for c in local_ctxs:
        c.close()
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = float(t.item())
    if not worst <= 1e-13:
        raise SystemExit(f"NCCL parity check failed: worst relative difference {worst:.3e} between the NCCL and the in-process run")
    return worst


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dims", type=int, default=20)
    ap.add_argument("--groups", type=int, default=128)
    ap.add_argument("--polar", type=int, default=2)
    ap.add_argument("--azimuthal", type=int, default=2)
    ap.add_argument("--cpu-dims", type=int, default=0, help="tiles per side of the bounded CPU sample (default: 10 for the cpu_baseline leg, 8 per step for --impl reference)")
    ap.add_argument("--ring", type=int, default=-1, help="Psi1 ring size in angle batches (umt_set_psi1_ring; default: as the free HBM allows)")
    ap.add_argument("--group-sets", default="32,64,32", help="group sets of the pipelined end-to-end call (umt_control_sweep_sets): a count of equal sets or their sizes; sizes that do not add up to --groups fall back to 2 equal sets; 1: one set, umt_control_sweep only")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the pre-timing NCCL parity check (N > 1)")
    ap.add_argument("--flux-iters", type=int, default=1, help="incidentFlux max iterations per sweep (driver default 2)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    if os.environ.get("UMT_TRACE"):   # debugging aid: Python stacks of every thread on stderr if the run is still going after 60 s
        import faulthandler
        faulthandler.dump_traceback_later(60, repeat=False)
    import torch
    import torch.distributed as dist
    from umt_b200 import teton

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    parity = None
    if world > 1 and not args.no_parity:
        parity = {"checked": True, "what": "2x2x2-tile domains, 16 angles, 8 groups: Phi/PsiB/Psi over NCCL == in-process transport on every rank (3 flux passes + savePsi sweep)",
                  "max_rel_diff": nccl_parity_check(teton, torch, dist, rank, world, local), "domains": world}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    d, G = args.dims, args.groups
    mesh = M.tiled_mesh((d, d, d), rank=rank, size=world)
    ctx = teton.SweepContext.from_mesh(mesh, G, device=local)
    ctx.compute_geometry(mesh.px)
    NA = ctx.build_product_quadrature(args.polar, args.azimuthal, 1)
    for b in mesh.boundaries:
        if b.bc_type == M.BC_SHARED:
            ctx.add_shared_boundary(b.neighbor, b.first_elem, b.n_elem)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(teton.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.set_comm(rank, world, bytes(idt.cpu().numpy().tobytes()))
    t0 = time.perf_counter()
    ctx.build_schedule()                      # rtorder/snnext/findexit for every angle on the host threads (the reference redoes it every cycle)
    sched_ms = (time.perf_counter() - t0) * 1e3
    nz, nc = mesh.nzones, mesh.ncornr
    tau = PR.tau()
    # page-locked host buffers (on the GPU's NUMA node) of what the Fortran caller hands over / gets back per ControlSweep
    h_sigt = ctx.host_array((nz, G)); h_sigt[:] = tau
    h_stotal = ctx.host_array((nc, G)); h_stotal[:] = 0.0
    h_phi = ctx.host_array((nc, G))
    if args.ring >= 0:
        ctx.set_psi1_ring(args.ring)
    ctx.upload_state(None, None, h_sigt, h_stotal, tau)
    ctx.init_teton(np.full(nz, PR.TR0), PR.group_bounds(G), PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(3), 0.0)
    ctx.init_phi_total()
    t0 = time.perf_counter()
    ctx.init_radiation_field()                # first use of the schedule: plan records, work items, Psi1 workspace
    finalize_ms = (time.perf_counter() - t0) * 1e3
    layout = ctx.psi_layout()
    mem_free, mem_total = torch.cuda.mem_get_info()
    unknowns = nc * NA * G
    total_unknowns = unknowns * world

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident metric -----------------------------------------------------
    for _ in range(args.warmup):
        ctx.sweep(False, args.flux_iters)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    sweep_ms = phi_ms = exch_ms = dev_ms = 0.0
    launches = 0
    iters = 0
    for _ in range(args.steps):
        iters += ctx.sweep(False, args.flux_iters)
        tm = ctx.last_times()
        sweep_ms += tm["sweep_ms"]; phi_ms += tm["phi_ms"]; exch_ms += tm["exchange_ms"]; dev_ms += tm["total_ms"]
        launches += ctx.last_launches()
    barrier()
    wall = time.perf_counter() - t0
    # wall clock between the two barrier+synchronize brackets; umt_sweep itself ends with an event
    # synchronize on the library's stream, whose CUDA-event times are reported in kernel_ms
    step_ms = max_over_ranks(wall * 1e3 / args.steps)
    device_step_ms = max_over_ranks(dev_ms / args.steps)   # CUDA events on the library's stream around each whole umt_sweep
    sweep_kernel_ms = max_over_ranks(sweep_ms / max(iters, 1))
    value = total_unknowns / (step_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers -----------------------------
    h2d = h_sigt.size * 8 + h_stotal.size * 8
    d2h = h_phi.size * 8
    for _ in range(1):
        ctx.control_sweep(h_sigt, h_stotal, tau, h_phi, False, args.flux_iters)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.control_sweep(h_sigt, h_stotal, tau, h_phi, False, args.flux_iters)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    e2e_what = "umt_control_sweep: Sigt, STotal from pinned host -> whole ControlSweep -> PhiTotal to pinned host (phi reduction and its D2H overlapped)"
    e2e_one_set = None
    host_node = getattr(ctx, "host_numa_node", None)
    sizes = [int(x) for x in str(args.group_sets).split(",")]
    if len(sizes) > 1 and sum(sizes) != G:
        sizes = [2]
    if len(sizes) == 1:
        sizes = [G // sizes[0]] * sizes[0] if sizes[0] > 0 and G % sizes[0] == 0 else [G]
    S = len(sizes)
    sets_note = None
    # (only when everything is resident twice over: a domain that already needs the Psi1 ring has no room for per-set workspaces)
    if world == 1 and S > 1 and sum(sizes) == G and all(g > 0 and g % 2 == 0 for g in sizes) and layout["single"]:
        sets_note = "group sets not tried: this domain runs in the single-psi (ring) layout, device memory is the constraint"
    elif world == 1 and S > 1 and sum(sizes) == G and all(g > 0 and g % 2 == 0 for g in sizes):
        # The same domain and groups as `S` group sets, one context each (the reference's phase-space sets split the groups the same
        # way), through umt_control_sweep_sets: upload of set k+1 and download of set k-1 run under the sweep of set k.
        e2e_one_set = {"value": total_unknowns / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "what": e2e_what}
        ctx.close()
        one_ms, one_what, sets = e2e_ms, e2e_what, []
        try:
            bounds = PR.group_bounds(G)
            g0s = [sum(sizes[:k]) for k in range(S)]
            sets, hs, ht, hp = [], [], [], []
            for k in range(S):
                Gs = sizes[k]
                c = teton.SweepContext.from_mesh(mesh, Gs, device=local)
                c.compute_geometry(mesh.px)
                c.build_product_quadrature(args.polar, args.azimuthal, 1)
                c.build_schedule()
                a = c.host_array((nz, Gs)); a[:] = tau
                b = c.host_array((nc, Gs)); b[:] = 0.0
                hs.append(a); ht.append(b); hp.append(c.host_array((nc, Gs)))
                c.upload_state(None, None, a, b, tau)
                c.init_teton(np.full(nz, PR.TR0), bounds[g0s[k]:g0s[k] + Gs + 1], PR.SPEED_LIGHT, PR.RAD_CONSTANT, PR.wtiso(3), 0.0)
                c.init_phi_total()
                c.init_radiation_field()
                sets.append(c)
            for _ in range(2):
                teton.control_sweep_sets(sets, hs, ht, tau, hp, False, args.flux_iters)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                teton.control_sweep_sets(sets, hs, ht, tau, hp, False, args.flux_iters)
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            set_ms = [c.last_times() for c in sets]
            e2e_what = (f"umt_control_sweep_sets: the {G} groups as {S} group sets of {'+'.join(map(str, sizes))} (one context each), Sigt/STotal of every set from pinned host -> "
                        f"sweep -> PhiTotal to pinned host, sets pipelined (upload of set k+1 and download of set k-1 under the sweep of set k); "
                        f"sweep kernels {', '.join('%.1f' % t['sweep_ms'] for t in set_ms)} ms")
            for c in sets:
                c.close()
            if e2e_ms > one_ms:   # tiny problems: the per-set launches cost more than the copies they hide -- a caller would not split
                sets_note = f"group sets {'+'.join(map(str, sizes))} measured slower at this size ({e2e_ms:.3f} ms per step): one-set call reported"
                e2e_ms, e2e_what, e2e_one_set = one_ms, one_what, None
        except teton.UmtError as e:   # e.g. out of device memory: the one-set number stands
            e2e_ms, e2e_what, e2e_one_set = one_ms, one_what, None
            sets_note = "group sets failed, one-set call reported: " + str(e)[:200]
            for c in sets:
                c.close()
    if rank == 0:
        sampler.stop()

    if rank == 0:
        peak, peak_src = measured_peak()
        balg = sweep_kernel_bytes_per_unknown(G)
        achieved = unknowns * balg / (sweep_kernel_ms * 1e-3) / 1e9
        # the committed capture is of the default workload in the two-buffer layout; any other run reports no traffic figure
        same = (d, G, args.polar, args.azimuthal) == (20, 128, 2, 2) and not layout["single"]
        tpu, tsrc = profiled_traffic_per_unknown() if same else (None, None)
        model41 = unknowns * algorithmic_bytes_per_unknown(G) / ((sweep_ms + phi_ms) / max(iters, 1) * 1e-3) / 1e9
        line = {
            "metric": "Sn sweep unknowns/sec (corner x angle x group)", "value": value, "unit": "unknowns/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": total_unknowns / (e2e_ms * 1e-3), "unit": "unknowns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "what": e2e_what, "one_group_set": e2e_one_set, "note": sets_note},
            "gpu_launches": launches,
            "parity_checked": bool(parity), "parity": parity,
            "flux_passes_per_step": iters / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tpu * unknowns / 1e9) if tpu else None, "traffic_unit": "GB per launch",
                         "traffic_source": (f"ncu dram read+write {tpu:.1f} B/unknown measured at -d 20,20,20 -G 128 P2 A2 ({tsrc}), scaled by this launch's unknowns" if tpu else None),
                         "kernel": "sweep3d_plan_kernel (persistent, all angles; warp-group build: 12 consumer warps per SM)", "bytes_per_unknown": balg, "kernel_ms": sweep_kernel_ms,
                         "peak_source": peak_src,
                         "whole_sweep_41B_model": {"bytes_per_unknown": algorithmic_bytes_per_unknown(G), "achieved": model41, "frac": model41 / peak}},
            "kernel_ms": {"sweep": sweep_ms / args.steps, "phi": phi_ms / args.steps, "exchange": exch_ms / args.steps,
                          "note": "exchange = unpack before the sweep + (tally, currents, transfer of the next pass's rows) on a second stream under the phi tally"},
            "psi_layout": dict(layout, device_memory_used_GB=(mem_total - mem_free) / 1e9),
            "setup_ms": {"build_schedule_host": sched_ms, "first_use_plan_records_items_workspace": finalize_ms,
                         "note": "once per mesh/quadrature (the reference rebuilds its schedules every cycle, control/initializeSets.F90:95-105); not in the timed region"},
            "host_buffers": {"rank_bound_to_numa_node": numa, "pages_bound_to_numa_node": host_node, "kind": "umt_host_alloc (page-locked, bound to the GPU's NUMA node)"},
            "device_ms_per_step": device_step_ms,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu:
            cd = args.cpu_dims or 10
            v, cores, dt, sample = cpu_sweep_rate(G, args.polar, args.azimuthal, (cd, cd, cd), steps=3, warmup=0)
            line["cpu_baseline"] = {"value": v, "unit": "unknowns/s", "cores": cores, "kind": "port", "sample": sample + f", 3 sweeps of {dt:.1f} s"}
            ctx.close()
            try:   # in a child process: the reference shim exit()s on any CUDA error and prints its allocations to stdout
                import subprocess
                out = subprocess.run([sys.executable, "-c", "import json, bench; print('REFCUDA ' + json.dumps(bench.reference_cuda_rate("
                                      f"{G}, {args.polar}, {args.azimuthal}, ({cd}, {cd}, {cd}))))"],
                                     cwd=ROOT, capture_output=True, text=True, timeout=600)
                got = [l for l in out.stdout.splitlines() if l.startswith("REFCUDA ")]
                rc = json.loads(got[-1][8:]) if got else {"unavailable": f"child exited {out.returncode}: {out.stderr[-200:]}"}
                if rc:
                    line["reference_cuda"] = rc
            except Exception as e:   # a reported extra, never allowed to cost the bench line
                line["reference_cuda"] = {"unavailable": repr(e)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
