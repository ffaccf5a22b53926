#!/usr/bin/env python
"""Benchmark of the depth-map integration hot path (BASELINE.json: voxel*view updates/sec,
1024^3 cells x 1000 views 1920x1080, z-slab sharded over 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (the reference's own kernel text on host cores)

One "step" = one full integration of all views into a zeroed volume.  Rank 0 prints ONE JSON line.
  value   device-resident: every rank's share of the views already sits in its HBM; the timed region
          covers best-cost filtering, the NCCL view all-gather (N>1), all integration launches and
          the slab gather to rank 0 (N>1).  CUDA events, barrier + synchronize both sides, max over ranks.
  e2e     the same job through the host-pointer C-ABI call (dmi_process_depth_maps at N=1): pinned
          host views -> H2D -> filter -> integrate -> D2H of the volume, all inside the timed region.
Synthetic data: unit sphere, Fibonacci-sphere pinhole cameras, generated on the GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

THRESH = 0.14            # --threshBestCost default, Reconstruction/main.cxx:79
FLOPS_PER_UNIT = 28.0    # algorithmic flops per voxel*view (FMA = 2)

WORKLOADS = {
    # name: (cells per axis, views, W, H)
    "config5": (1024, 1000, 1920, 1080),    # the configuration the metric is quoted on (fits one B200)
    "config4": (512, 500, 1920, 1080),
    "config3": (256, 100, 1920, 1080),
    "config2": (128, 10, 640, 480),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config5", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "exact"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cull", type=int, default=1, help="0: disable the brick culling of the fast kernel (dense worst case)")
    ap.add_argument("--counters", action="store_true", help="print the fast kernel's tier counters to stderr (slower kernel build)")
    ap.add_argument("--group", type=int, default=40, help="views per all-gather group (N>1)")
    return ap.parse_args()


def algorithmic_bytes(N, V, W, H, scalar_bytes=8, best_cost=True):
    """volume read once + written once, each depth (and best-cost) map read once."""
    return 2 * scalar_bytes * N ** 3 + 8 * V * W * H * (2 if best_cost else 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = f"/tmp/dmi_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own kernel text compiled for the host (oracle/_ref), all host threads
# ------------------------------------------------------------------------------------------------

def cpu_reference_rate(N, V, W, H, target_seconds=12.0, steps=1, warmup=0):
    """Times oracle/_ref/libref_tsdf_host.so (kind "reference") or, if absent, oracle/liboracle.so
    (kind "port") on a bounded sample of the workload: all N x N cells of `nz` z-planes in the middle
    of the grid x `nv` views.  Returns (units/s, description dict)."""
    import torch
    from cudadepthmapintegration_b200 import synthetic as syn
    from tests import _oracle
    ref = _oracle.load_ref_host()
    kind = "reference" if ref is not None else "port"
    orc = _oracle.load_oracle()
    cores = os.cpu_count() or 1
    grid = syn.make_grid(N)
    rp = syn.make_ray_potential(grid)
    nv = min(V, 4)
    K, RT = syn.make_cameras(V, W, H)
    K, RT = K[:nv], RT[:nv]
    d, b, _ = syn.render_views(K, RT, W, H, depth_noise=0.25 * float(grid.spacing.max()), want_color=False)
    depths = orc.apply_depth_threshold(d.numpy(), b.numpy(), THRESH).reshape(nv, H, W)
    # sparse volume: only the sampled planes are touched, but the harness indexes the full grid
    vol = np.zeros(grid.n_voxels, dtype=np.float64)

    def run(nz):
        k0 = N // 2 - nz // 2
        t0 = time.perf_counter()
        if ref is not None:
            ref.run(grid, rp, W, H, depths, K, RT, vol, k0, k0 + nz)
        else:
            orc.tsdf_integrate(grid, rp, W, H, depths, None, 0.0, K, RT, vol, k0, k0 + nz)
        return time.perf_counter() - t0, nz * N * N * nv

    t, u = run(1)                                   # calibration (also warms the caches / thread pool)
    rate = u / t
    nz = int(max(1, min(N, target_seconds * rate / (N * N * nv))))
    for _ in range(warmup):
        run(nz)
    times = []
    for _ in range(max(1, steps)):
        t, u = run(nz)
        times.append(t)
    rate = u / (sum(times) / len(times))
    desc = {"kind": kind, "cores": cores,
            "sample": f"{nz} z-planes x {N}x{N} cells x {nv} views of {W}x{H} ({u:.3g} voxel*views per step), "
                      f"OpenMP over (k,j) on all host threads; mean {1e3 * sum(times) / len(times):.0f} ms/step"}
    return rate, desc, sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, V, W, H = WORKLOADS[args.workload]
    # a few seconds per step so that warmup + steps stay within minutes
    rate, desc, sec = cpu_reference_rate(N, V, W, H, target_seconds=6.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "voxel*view updates/sec", "value": rate, "unit": "voxel*views/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TSDF integration {N}^3 cells x {V} views {W}x{H} (bounded sample per step, see cpu_baseline.sample)"},
        "cpu_baseline": dict(desc, value=rate, unit="voxel*views/s"),
        "e2e": {"value": rate, "unit": "voxel*views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from cudadepthmapintegration_b200 import Context, _lib, sharding, synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, V, W, H = WORKLOADS[args.workload]
    npix = W * H
    grid = syn.make_grid(N)
    rp = syn.make_ray_potential(grid)
    K, RT = syn.make_cameras(V, W, H)
    k0, k1 = sharding.slab_range(N, rank, world)

    ctx = Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, _lib.DMI_TSDF_KERNEL_EXACT if args.kernel == "exact" else _lib.DMI_TSDF_KERNEL_AUTO)
    ctx.set_option(_lib.DMI_OPT_CULL, args.cull)
    if args.counters:
        ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 1)
    ctx.initialize(grid.matrix, grid.point_dims, grid.origin, grid.spacing, rp.thick, rp.rho, rp.eta, rp.delta, (W, H))
    ctx.set_slab(k0, k1)

    # ---- view ownership: groups of G views; inside a group rank r owns a contiguous G/world share, so
    # that an in-place all-gather of the group's region of the resident buffer assembles it.
    G = args.group if world > 1 else V
    G = max(world, (G // world) * world)
    groups = [(g0, min(V, g0 + G)) for g0 in range(0, V, G)]

    def owned(g0, g1):
        n = g1 - g0
        per = (n + world - 1) // world
        a = min(g1, g0 + rank * per)
        return a, min(g1, a + per), per

    # ---- generate this rank's views on its GPU (stands for "loaded from the files it owns")
    noise = 0.25 * float(grid.spacing.max())
    all_depths = torch.empty((V, H, W), dtype=torch.float64, device=dev) if world > 1 else None
    my_idx = []
    for (g0, g1) in groups:
        a, b, _ = owned(g0, g1)
        my_idx.extend(range(a, b))
    my_idx = np.array(my_idx, dtype=np.int64)
    nmine = len(my_idx)
    my_depths = torch.empty((nmine, H, W), dtype=torch.float64, device=dev)
    my_cost = torch.empty((nmine, H, W), dtype=torch.float64, device=dev)
    for s0 in range(0, nmine, 8):
        idx = my_idx[s0:s0 + 8]
        # consecutive runs only (render_views hashes on first_view + offset)
        runs = np.split(idx, np.where(np.diff(idx) != 1)[0] + 1)
        off = s0
        for r in runs:
            d, c, _ = syn.render_views(K[r], RT[r], W, H, first_view=int(r[0]), device=dev, depth_noise=noise, want_color=False)
            my_depths[off:off + len(r)] = d
            my_cost[off:off + len(r)] = c
            off += len(r)
    torch.cuda.synchronize()

    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    full_volume = torch.empty(N ** 3, dtype=torch.float64, device=dev) if (world > 1 and rank == 0) else None

    def _as_tensor(ptr, count):
        """Wrap the context's slab (device memory owned by libdmi_b200) as a tensor, without copying."""
        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        return torch.as_tensor(h, device=dev)

    def step_device():
        """One full job with inputs resident in HBM."""
        ctx.volume_begin(None, np.float64)
        if world == 1:
            ctx.volume_integrate_device(V, my_depths.data_ptr(), my_cost.data_ptr(), THRESH, K, RT)
            return
        cur = torch.cuda.current_stream()
        comm_stream.wait_stream(cur)
        events = []
        off = 0
        # filter own views into place, all-gather group by group on the comm stream
        with torch.cuda.stream(comm_stream):
            for (g0, g1) in groups:
                a, b, per = owned(g0, g1)
                n = b - a
                if n > 0:
                    all_depths[a:b].copy_(my_depths[off:off + n])
                    ctx.set_stream(comm_stream.cuda_stream)
                    ctx.apply_depth_threshold_device(n * npix, all_depths[a:b].data_ptr(), my_cost[off:off + n].data_ptr(), THRESH)
                    ctx.set_stream(cur.cuda_stream)
                    off += n
                if (g1 - g0) == per * world:
                    dist.all_gather_into_tensor(all_depths[g0:g1].view(-1), all_depths[a:b].view(-1))
                else:   # ragged last group: plain broadcasts from each owner
                    for r in range(world):
                        ra = min(g1, g0 + r * per); rb = min(g1, ra + per)
                        if rb > ra:
                            dist.broadcast(all_depths[ra:rb], src=r)
                ev = torch.cuda.Event()
                ev.record(comm_stream)
                events.append(ev)
        for (g0, g1), ev in zip(groups, events):
            cur.wait_event(ev)
            ctx.volume_integrate_device(g1 - g0, all_depths[g0:g1].data_ptr(), None, 0.0, K[g0:g1], RT[g0:g1])
        # the finished slabs are gathered once (for contouring on rank 0)
        ptr, nbytes = ctx.volume_device_ptr()
        slab = _as_tensor(ptr, (k1 - k0) * N * N)
        # slabs may differ by one plane: gather with explicit point-to-point transfers
        if rank == 0:
            full_volume[k0 * N * N:k1 * N * N].copy_(slab)
            reqs = []
            for r in range(1, world):
                a, b = sharding.slab_range(N, r, world)
                reqs.append(dist.irecv(full_volume[a * N * N:b * N * N], src=r))
            for q in reqs:
                q.wait()
        else:
            dist.send(slab, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    units = float(N) ** 3 * V

    # ---- FP peaks for the roofline (MEASURED_PEAKS.json has no FP32/FP64 vector peak)
    fp64_peak = ctx.measure_fp_peak(0, 300.0)
    fp32_peak = ctx.measure_fp_peak(1, 300.0)

    # ---- device-resident timing
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.tsdf_kernel_stats()
    launches0 = ctx.launch_counter()
    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.tsdf_kernel_stats()
    launches0 = ctx.launch_counter()
    if sampler:
        sampler.start()
    ms_step = timed(step_device, args.steps, 0)
    clocks = sampler.stop() if sampler else None
    kernel_ms, kernel_launches = ctx.tsdf_kernel_stats()
    launches = ctx.launch_counter() - launches0
    if args.counters and args.kernel != "exact":
        print("tier counters (rank %d): %s" % (rank, ctx.tsdf_tier_counters()), file=sys.stderr)
    value = units / (ms_step * 1e-3)

    # ---- end to end through the host-pointer ABI
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(args, ctx, torch, dist, dev, rank, world, N, V, W, H, K, RT, my_depths, my_cost, my_idx, k0, k1, units, barrier)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # dominant kernel = the integration kernel; per launch: algorithmic flops / bytes over its mean duration
        k_units = units * args.steps / world                         # this rank's units over the timed region
        k_sec = kernel_ms * 1e-3
        ach_tflops = FLOPS_PER_UNIT * k_units / k_sec / 1e12
        alg_bytes = algorithmic_bytes(N, V, W, H) / world * args.steps
        roofline = {
            "bound": "fp64", "achieved": ach_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tflops / fp64_peak,
            "traffic": None,
            "kernel": "tsdf integration kernel (all launches of the timed region, CUDA events on the launch stream)",
            "peak_source": "DFMA issue-rate microbenchmark run in this process (dmi_measure_fp_peak); MEASURED_PEAKS.json has no FP64 vector peak",
            "algorithmic_flops_per_unit": FLOPS_PER_UNIT,
            "kernel_ms_per_step": kernel_ms / args.steps, "kernel_launches_per_step": kernel_launches / args.steps,
            "fp32_basis": {"peak": fp32_peak, "frac": ach_tflops / fp32_peak, "peak_source": "FFMA microbenchmark, this process"},
            "hbm": {"achieved": alg_bytes / k_sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / k_sec / 1e9 / hbm_peak, "peak_source": hbm_src,
                    "note": "algorithmic bytes (volume once in/out + every depth and best-cost map once) over kernel time; not the binding roofline"},
        }
        line = {
            "metric": "voxel*view updates/sec", "value": value, "unit": "voxel*views/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"TSDF integration {N}^3 cells x {V} views {W}x{H}, best-cost threshold {THRESH}, f64 volume",
                       "name": args.workload, "parallelism": f"z-slab x{world}" + (f", views all-gathered in groups of {G}" if world > 1 else ""),
                       "l2": "inputs (%.1f GB per step) exceed L2; no flush needed" % (algorithmic_bytes(N, V, W, H) / 1e9),
                       "kernel": args.kernel},
            "roofline": roofline, "gpu_launches": launches, "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            rate, desc, _ = cpu_reference_rate(N, V, W, H, target_seconds=12.0)
            line["cpu_baseline"] = dict(desc, value=rate, unit="voxel*views/s")
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def measure_e2e(args, ctx, torch, dist, dev, rank, world, N, V, W, H, K, RT, my_depths, my_cost, my_idx, k0, k1, units, barrier):
    """Host buffers in, host volume out, through dmi_process_depth_maps (N=1) or the streaming ABI (N>1)."""
    import psutil
    npix = W * H
    nmine = my_depths.shape[0]
    slab_cells = (k1 - k0) * N * N
    need = 2 * nmine * npix * 8 + slab_cells * 8
    avail = psutil.virtual_memory().available
    if need > 0.6 * avail:
        return {"value": None, "unit": "voxel*views/s", "skipped": f"host buffers need {need / 1e9:.0f} GB, {avail / 1e9:.0f} GB available"}
    h_depths = torch.empty((nmine, H, W), dtype=torch.float64, pin_memory=True)
    h_cost = torch.empty((nmine, H, W), dtype=torch.float64, pin_memory=True)
    h_vol = torch.zeros(slab_cells, dtype=torch.float64, pin_memory=True)
    h_depths.copy_(my_depths)
    h_cost.copy_(my_cost)
    torch.cuda.synchronize()
    if world > 1:
        # multi-GPU e2e is reported by the device-resident number plus each rank's own H2D/D2H; keep it simple:
        # every rank uploads its views, the job runs as in step_device, every rank downloads its slab.
        return {"value": None, "unit": "voxel*views/s", "skipped": "e2e is measured at N=1 in this round"}
    vol_np = h_vol.numpy()
    d_np, c_np = h_depths.numpy(), h_cost.numpy()

    def step():
        vol_np.fill(0.0)          # the filter zero-fills its output before the call (vtkCudaReconstructionFilter.cxx:133)
        ctx.process_depth_maps(d_np, c_np, THRESH, K, RT, vol_np)

    steps = max(1, min(args.steps, 2))
    step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) / steps
    ms = e0.elapsed_time(e1) / steps
    sec = max(wall, ms * 1e-3)
    return {"value": units / sec, "unit": "voxel*views/s", "ms_per_step": sec * 1e3, "steps": steps, "warmup": 1,
            "h2d_bytes_per_step": int(2 * nmine * npix * 8 + slab_cells * 8), "d2h_bytes_per_step": int(slab_cells * 8),
            "api": "dmi_process_depth_maps (host pointers; pinned host buffers; includes the upload of io_scalar the reference also does)"}


if __name__ == "__main__":
    main()
