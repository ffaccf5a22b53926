#!/usr/bin/env python
"""Benchmark of the depth-map integration hot path (BASELINE.json: voxel*view updates/sec,
1024^3 cells x 1000 views 1920x1080, z-slab sharded over 1/2/4/8 B200; colored points/sec).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (the reference's own kernel text on host cores)

One "step" = one full integration of all views into a zeroed volume.  Rank 0 prints ONE JSON line.
  value   device-resident: every rank's share of the views already sits in its HBM; the timed region
          covers view preparation (best-cost filter), the NCCL view all-gather (N>1), all integration
          launches and the slab gather to rank 0 (N>1).  CUDA events, barrier + synchronize both sides,
          max over ranks.
  e2e     the same job from pinned HOST buffers: H2D of the views (and of io_scalar at N=1, which the
          reference's call also uploads), the job, D2H of the volume, all inside the timed region.
          N=1 goes through dmi_process_depth_maps, the drop-in for ProcessDepthMap<double>.
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
COLOR_FLOPS_PER_UNIT = 37.0

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
    ap.add_argument("--no-coloration", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the B1 leg (the reference's CUDA kernel on this GPU)")
    ap.add_argument("--no-variants", action="store_true", help="skip the dense (--cull 0) and all-valid-depth scene lines")
    ap.add_argument("--cull", type=int, default=1, help="0: disable the brick culling of the fast kernel (dense worst case)")
    ap.add_argument("--emulate-rank", type=int, nargs=2, metavar=("RANK", "WORLD"), default=None,
                    help="N=1 diagnostic: integrate only the z-layers that RANK of WORLD would own (value counts those pairs)")
    ap.add_argument("--quota", type=int, default=0, help="bricks per CTA of the integration kernel (0: the library's default, 32 at N=1, 8 at N>1)")
    ap.add_argument("--breakdown", action="store_true", help="N>1: add the per-rank compute / gather spans of the last step to the line")
    ap.add_argument("--cost-model", default="iid", choices=["iid", "coherent"],
                    help="synthetic best-cost maps: independent per pixel (default, the headline workload) or spatially coherent")
    ap.add_argument("--color-points", type=int, default=0,
                    help="0 (default): colour the vertices of the isosurface of the fused volume (GPU contour, value --contour); "
                         "> 0: that many points on the unit sphere instead")
    ap.add_argument("--contour", type=float, default=1.0, help="isosurface value (Reconstruction/main.cxx:80)")
    ap.add_argument("--color-views", type=int, default=1000)
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
                                          "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
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

def host_threads():
    """Host threads this process may use (its CPU affinity mask, else the core count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def cpu_reference_rate(N, V, W, H, target_seconds=12.0, steps=1, warmup=0):
    """Times oracle/_ref/libref_tsdf_host.so (kind "reference") or, if absent, oracle/liboracle.so
    (kind "port") on a bounded sample of the workload: all N x N cells of `nz` z-planes in the middle
    of the grid x `nv` views.  Returns (units/s, description dict, seconds per step).
    The OpenMP thread count is set explicitly (torch.distributed.run exports OMP_NUM_THREADS=1 to its workers)
    and the count the library reports back is what `cores` states."""
    from cudadepthmapintegration_b200 import synthetic as syn
    from tests import _oracle
    ref = _oracle.load_ref_host()
    kind = "reference" if ref is not None else "port"
    orc = _oracle.load_oracle()
    orc.threads(host_threads())
    cores = ref.threads(host_threads()) if ref is not None else orc.threads()
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
                      f"OpenMP over (k,j) on {cores} host threads (omp_get_max_threads after omp_set_num_threads); "
                      f"mean {1e3 * sum(times) / len(times):.0f} ms/step"}
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
        "config": {"workload": f"TSDF integration {N}^3 cells x {V} views {W}x{H} (bounded sample per step, see cpu_baseline.sample)",
                   "name": args.workload},
        "cpu_baseline": dict(desc, value=rate, unit="voxel*views/s"),
        "e2e": {"value": rate, "unit": "voxel*views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# coloration (secondary metric: colored points/sec)
# ------------------------------------------------------------------------------------------------

def mesh_points(P):
    """Stand-in for the isocontour's vertices (VTK is not installed): P points on the unit sphere in
    scanline order (sorted by z, then y, then x cell like a marching-cubes traversal), float32."""
    from cudadepthmapintegration_b200 import synthetic as syn
    pts = syn.fibonacci_sphere_points(P).astype(np.float64)
    cell = np.floor((pts + 1.2) / (2.4 / 1024)).astype(np.int64)
    order = np.lexsort((cell[:, 0], cell[:, 1], cell[:, 2]))
    return np.ascontiguousarray(pts[order].astype(np.float32))


def measure_coloration(args, ctx, torch, dist, dev, rank, world, W, H, fp64_peak, fp32_peak, timed, mesh, steps=3, warmup=2):
    """Secondary metric (BASELINE.json): colored points/sec at config 5's shape (~10 M points x 1000 views).  N>1: points
    sharded by contiguous index range (dmi_shard_range), every rank renders ("loads") one block of the colour images and
    the library all-gathers them with NCCL inside the timed region (dmi_shard_colorize_device)."""
    from cudadepthmapintegration_b200 import engine, synthetic as syn
    from tests import _oracle
    V = args.color_views
    if mesh is not None:
        pts_h = mesh.cpu().numpy()                          # the isosurface's vertices, in the contour's own order
        mesh_text = "isosurface vertices of the fused volume (GPU contour, float32, contour order)"
    else:
        pts_h = mesh_points(args.color_points)
        mesh_text = "points on the unit sphere (scanline order, float32)"
    P = pts_h.shape[0]
    mesh = None
    K, RT = syn.make_cameras(V, W, H)
    v0, nv = engine.shard_range(V, world, rank)
    p0, npts = engine.shard_range(P, world, rank)
    cols = torch.empty((max(nv, 1), H, W, 3), dtype=torch.uint8, device=dev)
    for q in range(0, nv, 8):
        m = min(8, nv - q)
        _, _, c = syn.render_views(K[v0 + q:v0 + q + m], RT[v0 + q:v0 + q + m], W, H, first_view=v0 + q, device=dev, want_best_cost=False)
        cols[q:q + m] = c
    pts = torch.from_numpy(pts_h[p0:p0 + npts]).to(dev)
    mean = torch.zeros((max(npts, 1), 3), dtype=torch.uint8, device=dev)
    med = torch.zeros_like(mean)
    nb = torch.zeros(max(npts, 1), dtype=torch.int32, device=dev)

    def step():
        if world == 1:
            ctx.colorize_device(P, pts.data_ptr(), np.float32, V, cols.data_ptr(), K, RT, W, H, mean.data_ptr(), med.data_ptr(), nb.data_ptr())
        else:
            ctx.shard_colorize_device(npts, pts.data_ptr(), np.float32, V, cols.data_ptr(), K, RT, W, H, mean.data_ptr(), med.data_ptr(),
                                      nb.data_ptr())

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ctx.color_kernel_stats()
    ms, _ = timed(step, steps)
    kms, kn = ctx.color_kernel_stats()
    # order-independent digest of the three output arrays over ALL points: equal across --gpus N
    wgt = torch.arange(p0 + 1, p0 + npts + 1, dtype=torch.int64, device=dev)
    dig = (wgt * (nb[:npts].to(torch.int64) + 7 * mean[:npts].to(torch.int64).sum(1) + 13 * med[:npts].to(torch.int64).sum(1))).sum().reshape(1)
    if world > 1:
        dist.all_reduce(dig)
    digest = "%016x" % (int(dig.item()) & 0xFFFFFFFFFFFFFFFF)
    if rank != 0:
        return None
    ksec = kms / max(kn, 1) * 1e-3
    tflops = COLOR_FLOPS_PER_UNIT * npts * V / ksec / 1e12          # rank 0's kernel: its points x all views
    alg_bytes = 3.0 * V * W * H + 22.0 * npts                       # every colour image once + points in + results out
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {"metric": "colored points/sec", "value": P / (ms * 1e-3), "unit": "points/s", "point_views_per_s": P * V / (ms * 1e-3),
           "n_gpus": world, "ms_per_step": ms, "steps": steps, "warmup": warmup, "kernel_ms_per_step": kms / max(kn, 1),
           "result_digest": digest,
           "config": {"workload": f"mesh coloration {P} points x {V} views {W}x{H}; mesh = {mesh_text}",
                      "parallelism": "one GPU" if world == 1 else f"points sharded by contiguous index range over {world} GPUs; colour images "
                                     "all-gathered with NCCL inside the timed region (dmi_shard_colorize_device)",
                      "l2": "colour images (%.1f GB) exceed L2; no flush needed" % (3.0 * V * W * H / 1e9)},
           "gpu_launches_per_step": kn / max(steps, 1),
           "roofline": {"bound": "fp64", "achieved": tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tflops / fp64_peak,
                        "kernel": "colorize_kernel", "algorithmic_flops_per_unit": COLOR_FLOPS_PER_UNIT,
                        "definition": "37 algorithmic FP64 flops (the reference's uncontracted projection, SURVEY.md 8d) x rank 0's points x views / "
                                      "mean CUDA-event time of its kernel launches of the timed region",
                        "peak_source": "DFMA issue-rate microbenchmark run in this process (dmi_measure_fp_peak)",
                        "fp32_basis": {"peak": fp32_peak, "frac": tflops / fp32_peak},
                        "hbm": {"achieved": alg_bytes / ksec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / ksec / 1e9 / hbm_peak,
                                "note": "algorithmic bytes (each colour image once, points, results) over kernel time; the gathers are "
                                        "sector-granular L2 traffic, not HBM"},
                        "ncu": ncu_summary("coloration", "colorize", 1)}}
    if world > 1:
        return out
    # ---- end to end through dmi_colorize: pinned HOST buffers in, host arrays out, copies inside the timed region
    import psutil
    if 3.0 * V * W * H * 3 < 0.5 * psutil.virtual_memory().available:
        cols_pin = torch.empty((V, H, W, 3), dtype=torch.uint8, pin_memory=True)
        cols_pin.copy_(cols)
        pts_pin = torch.from_numpy(pts_h).pin_memory()
        torch.cuda.synchronize()
        cols_h, pts_np = cols_pin.numpy(), pts_pin.numpy()
        ctx.colorize(pts_np[:1000], cols_h, K, RT, W, H)             # warm-up: allocates the library's device scratch
        hmean, hmed, hnb = ctx.colorize(pts_np, cols_h, K, RT, W, H)
        t0 = time.perf_counter()
        hmean, hmed, hnb = ctx.colorize(pts_np, cols_h, K, RT, W, H)
        dt = time.perf_counter() - t0
        out["e2e"] = {"value": P / dt, "unit": "points/s", "ms_per_step": dt * 1e3, "steps": 1, "warmup": 1,
                      "h2d_bytes_per_step": int(3 * V * W * H + pts_np.nbytes), "d2h_bytes_per_step": int(10 * P),
                      "api": "dmi_colorize (host pointers; pinned host buffers)"}
        same = (np.array_equal(hnb, nb.cpu().numpy()) and np.array_equal(hmed, med.cpu().numpy()) and np.array_equal(hmean, mean.cpu().numpy()))
        out["e2e"]["equals_device_resident_result"] = bool(same)
    else:
        cols_h = cols.cpu().numpy()
        out["e2e"] = {"value": None, "skipped": "not enough host memory for pinned colour images"}
    # ---- parity spot check + CPU baselines on bounded samples
    orc = _oracle.load_oracle()
    used = orc.threads(host_threads())
    ns = min(P, 100000)
    sel = np.linspace(0, P - 1, ns).astype(np.int64)                 # spread over the whole mesh
    t0 = time.perf_counter()
    wmean, wmed, wnb = orc.colorize(pts_h[sel], cols_h, K, RT, W, H)
    dt = time.perf_counter() - t0
    dsel = torch.from_numpy(sel).to(dev)
    ok = (np.array_equal(nb[dsel].cpu().numpy(), wnb) and np.array_equal(med[dsel].cpu().numpy(), wmed)
          and np.array_equal(mean[dsel].cpu().numpy(), wmean))
    out["parity_sample_ok"] = bool(ok)
    out["parity_sample"] = f"{ns} points spread over the mesh x {V} views, bit-exact mean / median / count against oracle/color_oracle.c"
    out["cpu_baseline"] = {"value": ns / dt, "unit": "points/s", "cores": used, "kind": "port",
                           "sample": f"{ns} points x {V} views, oracle/color_oracle.c, OpenMP over points on {used} threads, {dt * 1e3:.0f} ms"}
    refc = _oracle.load_ref_coloration()
    if refc is not None:
        nr, vr = min(P, 20000), min(V, 100)
        sel2 = np.linspace(0, P - 1, nr).astype(np.int64)
        t0 = time.perf_counter()
        rmean, rmed, rnb = refc.colorize(pts_h[sel2], cols_h[:vr], K[:vr], RT[:vr], W, H)
        dt2 = time.perf_counter() - t0
        omean, omed, onb = orc.colorize(pts_h[sel2], cols_h[:vr], K[:vr], RT[:vr], W, H)
        out["cpu_reference"] = {"value": nr * vr / dt2, "unit": "point*views/s", "cores": 1, "kind": "reference",
                                "sample": f"{nr} points x {vr} views through the reference's own MeshColoration.cxx (single thread, VTK stand-in; "
                                          f"includes its constructor's per-view reads), {dt2 * 1e3:.0f} ms",
                                "equals_oracle": bool(np.array_equal(rnb, onb) and np.array_equal(rmed, omed) and np.array_equal(rmean, omean))}
    return out


def measure_reference_cuda(torch, dev, grid, rp, N, V, W, H, K, RT, my_depths, my_cost, host_bufs, ours, units):
    """B1: depthMapKernel<double> of Reconstruction/CudaReconstruction.cu:47-212, compiled
    unmodified for sm_100a (oracle/_ref/libref_tsdf_cuda*.so), driven like ProcessDepthMap's loop (:343-365: per view
    cudaDeviceSynchronize + 3 H2D copies + one launch) on the same workload, one step.  Also the full-size parity
    check of our volume (`ours`, device tensor) against the -fmad=false build (= the shipped -G numerics)."""
    from tests import _oracle
    if N > 1024:
        return {"unavailable": "the reference's launch shape caps the grid at 1024 voxels in x (CudaReconstruction.cu:330)"}
    if host_bufs is None:
        return {"unavailable": "no host buffers (the end-to-end leg was skipped)"}
    ref_o3, ref_nofma = _oracle.load_ref_cuda(nofma=False), _oracle.load_ref_cuda(nofma=True)
    if ref_o3 is None or ref_nofma is None:
        return {"unavailable": "oracle/_ref/libref_tsdf_cuda*.so not present (built where /root/reference exists)"}
    h_depths, h_vol = host_bufs["depths"], host_bufs["vol"]
    # ReconstructionData::ApplyDepthThresholdFilter runs on the host before the copy (:348): the harness gets filtered maps
    for v0 in range(0, V, 32):
        d, c = my_depths[v0:v0 + 32], my_cost[v0:v0 + 32]
        h_depths[v0:v0 + 32].copy_(torch.where(c > THRESH, torch.full_like(d, -1.0), d))
    torch.cuda.synchronize()
    out = {"kernel": "depthMapKernel<double> (reference text, nvcc -O3 -gencode arch=compute_100a,code=sm_100a)",
           "driven": "like ProcessDepthMap :343-365: per view cudaDeviceSynchronize + H2D of the depth map (pinned host buffer here; "
                     "the reference's is pageable) and of K, RT + one launch; best-cost filter applied beforehand (not timed)",
           "steps": 1, "n_gpus": 1}
    h_vol.zero_()
    _, kernel_ms, span_ms = ref_o3.run(grid, rp, W, H, h_depths.numpy(), K, RT, h_vol.numpy(), per_kernel_events=True)
    out["kernel_only"] = {"value": units / (kernel_ms * 1e-3), "unit": "voxel*views/s", "ms_per_step": kernel_ms,
                          "what": "sum of the CUDA-event times of the V launches"}
    out["as_driven"] = {"value": units / (span_ms * 1e-3), "unit": "voxel*views/s", "ms_per_step": span_ms,
                        "what": "first H2D to last kernel completion (volume upload / download not included)"}
    h_vol.zero_()
    ref_nofma.run(grid, rp, W, H, h_depths.numpy(), K, RT, h_vol.numpy(), per_kernel_events=False)
    n_bad = n_support = 0
    max_abs = 0.0
    step = 1 << 26
    for o in range(0, ours.numel(), step):
        r = h_vol[o:o + step].to(dev, non_blocking=True)
        g = ours[o:o + step]
        err = (g - r).abs()
        n_bad += int((err > 1e-6 + 1e-5 * r.abs()).sum().item())
        n_support += int(((g != 0) != (r != 0)).sum().item())
        max_abs = max(max_abs, float(err.max().item()))
    out["parity_full_size"] = {"against": "the same kernel text built with -fmad=false (numerics of the shipped -G build)",
                               "voxels": int(ours.numel()), "out_of_tolerance": n_bad, "support_mismatches": n_support,
                               "max_abs_err": max_abs, "tolerance": "1e-6 abs + 1e-5 rel"}
    return out


def measure_variants(args, ctx, _lib, torch, dev, syn, grid, N, V, W, H, my_depths, my_cost, scene, step_device, timed, units):
    """The integration kernel on inputs other than the headline's: (1) the same scene with the brick culling off (every
    voxel*view pair evaluated one by one: the dense worst case), (2) a scene in which EVERY pixel of every depth map is
    valid (the sphere seen from inside, cameras at radius 0.3, no best-cost maps).  2 steps each after 1 warm-up."""
    out = {}

    def run(label, note):
        step_device()
        ctx.tsdf_kernel_stats()
        ms, _ = timed(step_device, 2)
        kms, kn = ctx.tsdf_kernel_stats()
        ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 1)
        step_device()
        torch.cuda.synchronize()
        t = ctx.tsdf_tier_counters()
        ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 0)
        out[label] = {"value": units / (ms * 1e-3), "unit": "voxel*views/s", "ms_per_step": ms, "steps": 2, "warmup": 1,
                      "kernel_ms_per_step": kms / 2, "evaluated_fraction_of_pairs": t["units"] / units,
                      "free_space_fraction_of_pairs": t["uniform_front"] / units,
                      "evaluated_pairs_per_s": t["units"] / (kms / 2 * 1e-3), "what": note}

    if args.cull:
        ctx.set_option(_lib.DMI_OPT_CULL, 0)
        run("dense_no_culling", "headline scene, DMI_OPT_CULL = 0: all N^3*V pairs evaluated voxel by voxel")
        ctx.set_option(_lib.DMI_OPT_CULL, 1)
    K2, RT2 = syn.make_cameras(V, W, H, radius=0.3)
    noise = 0.25 * float(grid.spacing.max())
    for v0 in range(0, V, 8):
        d, _, _ = syn.render_views(K2[v0:v0 + 8], RT2[v0:v0 + 8], W, H, first_view=v0, device=dev, depth_noise=noise,
                                   want_color=False, want_best_cost=False, scene="room")
        my_depths[v0:v0 + 8] = d
    torch.cuda.synchronize()
    scene.update(K=K2, RT=RT2, use_cost=False)
    run("all_valid_depth_room", "sphere seen from inside (cameras at radius 0.3 looking through the centre), every pixel valid, no best-cost maps")
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def ncu_summary(workload, kernel, world):
    """dram bytes per launch / issue statistics of the dominant kernel from the committed ncu summary
    (profiles/ncu_summary.json: written by profiles/ncu_summary.py from an `ncu --set full` capture of this same
    bench command; it names the commit and the kernel time of the capture).  None when there is no capture for
    this workload -- nothing is pasted into this file."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f).get(f"{workload}/{kernel}/n{world}")
    except Exception:
        return None


_JSON_OUT = None


def emit(text):
    """The ONE JSON line goes to the real stdout; everything else a library prints to fd 1 (NCCL's version
    banner, for instance) has been redirected to stderr by guard_stdout()."""
    if _JSON_OUT is None:
        print(text, flush=True)
    else:
        os.write(_JSON_OUT, (text + "\n").encode())


def guard_stdout():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)


def main():
    guard_stdout()
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from cudadepthmapintegration_b200 import Context, _lib, engine, synthetic as syn

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

    ctx = Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_option(_lib.DMI_OPT_TSDF_KERNEL, _lib.DMI_TSDF_KERNEL_EXACT if args.kernel == "exact" else _lib.DMI_TSDF_KERNEL_AUTO)
    ctx.set_option(_lib.DMI_OPT_CULL, args.cull)
    nccl_version = None
    if world > 1:
        # the library owns the multi-GPU path (dmi_comm_* / dmi_shard_*): torch.distributed only carries the NCCL
        # unique id to the other ranks and the barriers / max-over-ranks of the timing
        uid = [engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        nccl_version = ctx.comm_info()[2]
        exchange_on = "the copy engines (zero-CTA NCCL communicator, symmetric windows)" if ctx.comm_copy_engines() else "NCCL's kernels"
        ctx.shard_initialize(grid.matrix, grid.point_dims, grid.origin, grid.spacing, rp.thick, rp.rho, rp.eta, rp.delta, (W, H))
    else:
        ctx.initialize(grid.matrix, grid.point_dims, grid.origin, grid.spacing, rp.thick, rp.rho, rp.eta, rp.delta, (W, H))
        if args.emulate_rank:
            ctx.set_slab_layers(32, args.emulate_rank[0], args.emulate_rank[1])
    if args.quota > 0:
        ctx.set_option(_lib.DMI_OPT_BRICK_QUOTA, args.quota)
    slab_cells = ctx.slab_cells

    # ---- this rank's views, generated on its GPU (stands for "loaded from the files it owns"): the library says which
    my_idx = engine.shard_view_indices(V, world, rank).astype(np.int64) if world > 1 else np.arange(V, dtype=np.int64)
    nmine = len(my_idx)
    noise = 0.25 * float(grid.spacing.max())
    my_depths = torch.empty((max(nmine, 1), H, W), dtype=torch.float64, device=dev)
    my_cost = torch.empty((max(nmine, 1), H, W), dtype=torch.float64, device=dev)
    for s0 in range(0, nmine, 8):
        idx = my_idx[s0:s0 + 8]
        runs = np.split(idx, np.where(np.diff(idx) != 1)[0] + 1)     # render_views hashes on first_view + offset
        off = s0
        for r in runs:
            d, c, _ = syn.render_views(K[r], RT[r], W, H, first_view=int(r[0]), device=dev, depth_noise=noise, want_color=False,
                                        cost_model=args.cost_model)
            my_depths[off:off + len(r)] = d
            my_cost[off:off + len(r)] = c
            off += len(r)
    torch.cuda.synchronize()

    full_volume = torch.empty(N ** 3, dtype=torch.float64, device=dev) if (world > 1 and rank == 0) else None
    marks = {}

    def slab_tensor():
        """The context's slab (device memory owned by libdmi_b200) as a tensor, without copying."""
        ptr, _ = ctx.volume_device_ptr()

        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (slab_cells,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        return torch.as_tensor(h, device=dev)

    scene = {"K": K, "RT": RT, "use_cost": True}      # N=1 variants swap the cameras / drop the best-cost maps

    def step_device():
        """One full job with this rank's views resident in its HBM: N=1 = dmi_volume_integrate_device; N>1 = the sharded
        entry point (owner-side preparation, NCCL all-gather of the prepared views group by group behind the integration)
        + the gather of the finished layers into rank 0's whole-grid volume."""
        ctx.volume_begin(None, np.float64)
        if world == 1:
            ctx.volume_integrate_device(V, my_depths.data_ptr(), my_cost.data_ptr() if scene["use_cost"] else None, THRESH,
                                        scene["K"], scene["RT"])
            return
        cur = torch.cuda.current_stream()
        if args.breakdown:
            for nm in ("t0", "compute_done", "gather_done"):
                marks[nm] = torch.cuda.Event(enable_timing=True)
            marks["t0"].record(cur)
        ctx.shard_integrate_device(V, my_depths.data_ptr(), my_cost.data_ptr(), THRESH, K, RT)
        if args.breakdown:
            marks["compute_done"].record(cur)
        ctx.shard_gather_volume_device(0, full_volume.data_ptr() if rank == 0 else None)
        if args.breakdown:
            marks["gather_done"].record(cur)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0].item()) / steps, float(ms[1].item()) / steps

    units = float(N) ** 3 * V if not (args.emulate_rank and world == 1) else float(slab_cells) * V

    # ---- FP peaks for the roofline (MEASURED_PEAKS.json has no FP32/FP64 vector peak)
    fp64_peak = ctx.measure_fp_peak(0, 300.0)
    fp32_peak = ctx.measure_fp_peak(1, 300.0)

    # ---- device-resident timing
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.tsdf_kernel_stats()
    launches0 = ctx.launch_counter()
    if sampler:
        sampler.start()
    ms_step, _ = timed(step_device, args.steps)
    clocks = sampler.stop() if sampler else None
    kernel_ms, kernel_launches = ctx.tsdf_kernel_stats()
    launches = ctx.launch_counter() - launches0
    breakdown = None
    if args.breakdown and world > 1:
        t0 = marks["t0"]
        mine = torch.tensor([t0.elapsed_time(marks["compute_done"]), t0.elapsed_time(marks["gather_done"]), kernel_ms / args.steps],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        breakdown = {"per_rank_ms_of_the_last_step": [{"integration_enqueued_work_done": float(x[0]), "layer_gather_done": float(x[1]),
                                                        "integration_kernels": float(x[2])} for x in allr],
                     "note": "ms from the start of the step on each rank's stream; integration_kernels = summed spans of the integration launches"}
    value = units / (ms_step * 1e-3)
    # bit-pattern digest of the finished volume (untimed): equal across --gpus N when the shards assemble exactly
    final = full_volume if world > 1 else slab_tensor()
    digest = None
    if rank == 0:
        bits = final.view(torch.int64)
        w = torch.arange(1, bits.numel() + 1, dtype=torch.int64, device=dev)
        digest = "%016x" % (int((bits * w).sum().item()) & 0xFFFFFFFFFFFFFFFF)
        del w, bits
    final = None

    # ---- how much of the work the fast kernel actually evaluated (diagnostic build of the kernel, untimed)
    tiers = None
    if args.kernel != "exact":
        ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 1)
        step_device()
        barrier()
        tiers = ctx.tsdf_tier_counters()
        ctx.set_option(_lib.DMI_OPT_TIER_COUNTERS, 0)

    # ---- end to end from host buffers
    e2e, host_bufs = None, None
    if not args.no_e2e:
        e2e, host_bufs = measure_e2e(args, ctx, torch, dev, rank, world, N, V, W, H, K, RT, my_depths, my_cost, slab_cells, units, timed)

    # ---- the stage after the path (Reconstruction/main.cxx:151-189): isosurface of the fused volume, on the GPU; its vertices
    # are the mesh the coloration benchmark colours (BASELINE config 5: "coloration of the extracted ~10M-point mesh")
    contour, mesh = None, None
    if not args.no_coloration and not args.emulate_rank and args.color_points == 0:
        step_device()
        barrier()
        if rank == 0:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nv, nt = ctx.contour_device(full_volume.data_ptr() if world > 1 else None, np.float64, args.contour)
            e1.record()
            torch.cuda.synchronize()
            pv, _, _, _ = ctx.contour_device_ptr()

            class _M:
                pass
            hm = _M()
            hm.__cuda_array_interface__ = {"shape": (nv, 3), "typestr": "<f4", "data": (pv, False), "version": 3}
            mesh = torch.as_tensor(hm, device=dev).clone()
            contour = {"ms": e0.elapsed_time(e1), "vertices": nv, "triangles": nt, "value": args.contour,
                       "what": "dmi_contour_device on the fused volume in device memory: cell -> point averaging, surface vertices (float32, "
                               "grid matrix applied), triangles; 5 kernels + 2 scans, counts read back to the host in between",
                       "instead_of": "D2H of the %.1f GB volume + host vtkCellDataToPointData / vtkContourFilter" % (N ** 3 * 8 / 1e9)}
        if world > 1:
            shape = [tuple(mesh.shape) if rank == 0 else None]
            dist.broadcast_object_list(shape, src=0)
            if rank != 0:
                mesh = torch.empty(shape[0], dtype=torch.float32, device=dev)
            dist.broadcast(mesh, src=0)

    # ---- B1: the reference's own CUDA kernel on this GPU, driven like ProcessDepthMap (N=1 only: it has no multi-GPU path)
    reference_cuda = None
    if world == 1 and not args.no_reference_cuda and not args.emulate_rank and args.kernel != "exact":
        step_device()
        torch.cuda.synchronize()
        reference_cuda = measure_reference_cuda(torch, dev, grid, rp, N, V, W, H, K, RT, my_depths, my_cost, host_bufs, slab_tensor(), units)

    # ---- the same kernel on other inputs, so that the scene's share of the headline is visible (N=1 only)
    variants = None
    if world == 1 and not args.no_variants and not args.emulate_rank and args.kernel != "exact":
        variants = measure_variants(args, ctx, _lib, torch, dev, syn, grid, N, V, W, H, my_depths, my_cost, scene, step_device, timed, units)

    coloration = None
    if not args.no_coloration and not args.emulate_rank:
        del my_depths, my_cost
        host_bufs = None
        full_volume = None
        torch.cuda.empty_cache()
        coloration = measure_coloration(args, ctx, torch, dist, dev, rank, world, W, H, fp64_peak, fp32_peak, timed, mesh)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # dominant kernel = the integration kernel; per launch: algorithmic flops / bytes over its mean duration
        k_units = units * args.steps / world                         # this rank's pairs over the timed region
        k_sec = kernel_ms * 1e-3
        all_pairs_tflops = FLOPS_PER_UNIT * k_units / k_sec / 1e12
        evaluated = 1.0
        if tiers and tiers["brick_views"] > 0:
            evaluated = tiers["units"] / (units / world)          # rank 0's share
        ach = all_pairs_tflops * evaluated
        alg_bytes = algorithmic_bytes(N, V, W, H) / world * args.steps
        ncu = ncu_summary(args.workload, args.kernel, world)
        roofline = {
            "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum per launch of the integration kernel: read from the committed
            # summary of an ncu --set full capture of this command (None when this workload has none)
            "traffic": (ncu or {}).get("dram_bytes_per_launch"),
            "traffic_unit": "bytes per launch (algorithmic: %.3g)" % (alg_bytes / max(kernel_launches, 1)),
            "ncu": ncu,
            "kernel": "tsdf_fast_kernel" if args.kernel != "exact" else "tsdf_exact_kernel",
            "definition": "28 algorithmic flops x the voxel*view pairs the kernel EVALUATED one by one (pairs settled by the exact "
                          "brick tests -- culled, or free space in front of the surface: one add -- are excluded) / summed CUDA-event time of the integration launches of the timed region",
            "peak_source": "FFMA issue-rate microbenchmark run in this process (dmi_measure_fp_peak); MEASURED_PEAKS.json has no FP32/FP64 vector peak",
            "binding_resource": "instruction issue (compares, rounding, address arithmetic, one gather per pair): see profiles/ for smsp__issue_active",
            "evaluated_fraction_of_pairs": evaluated,
            "free_space_fraction_of_pairs": (tiers["uniform_front"] / (units / world)) if tiers else None,
            "all_pairs": {"achieved": all_pairs_tflops, "frac_fp32": all_pairs_tflops / fp32_peak, "frac_fp64": all_pairs_tflops / fp64_peak,
                          "note": "all N^3*V pairs counted as SURVEY.md 8d asks; exceeds the FMA roofline because most pairs are proven to contribute nothing without being evaluated"},
            "fp64_basis": {"peak": fp64_peak, "frac": ach / fp64_peak, "peak_source": "DFMA microbenchmark, this process"},
            "algorithmic_flops_per_unit": FLOPS_PER_UNIT,
            "kernel_ms_per_step": kernel_ms / args.steps, "kernel_launches_per_step": kernel_launches / args.steps,
            "hbm": {"achieved": alg_bytes / k_sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / k_sec / 1e9 / hbm_peak,
                    "peak_source": hbm_src, "note": "algorithmic bytes (volume once in/out + every depth and best-cost map once) over kernel time; not the binding roofline"},
        }
        if tiers:
            u = max(tiers["units"], 1)
            roofline["tiers"] = {"fp32_certified": tiers["t1_certified"] / u, "fp64_tier": tiers["t2_entered"] / u,
                                 "exact_tier": tiers["t3_entered"] / u, "band_fp64": tiers["near_band"] / u,
                                 "far_front": tiers["far_front"] / u, "far_behind": tiers["far_behind"] / u,
                                 "invalid_or_rejected": tiers["invalid_or_rejected"] / u,
                                 "validity_only_phase_c": tiers["validity_only"] / u,
                                 "note": "fractions of the evaluated pairs (rank 0)"}
        par = "one GPU"
        if world > 1:
            par = (f"z-layers of 32 cells dealt round-robin over {world} GPUs (dmi_shard_*); prepared views (8 B/pixel lossless split depth; "
                   f"the tile statistics are rebuilt locally) all-gathered with NCCL {nccl_version} on {exchange_on} in ramped groups of up to "
                   f"{max(1, 128 // world) * world} views behind the integration; "
                   "finished layers gathered into rank 0's volume with NCCL send/recv inside the timed region")
        line = {
            "metric": "voxel*view updates/sec", "value": value, "unit": "voxel*views/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"TSDF integration {N}^3 cells x {V} views {W}x{H}, best-cost threshold {THRESH}, f64 volume",
                       "name": args.workload, "parallelism": par,
                       "l2": "inputs (%.1f GB per step) exceed L2; no flush needed" % (algorithmic_bytes(N, V, W, H) / 1e9),
                       "kernel": args.kernel, "cull": args.cull, "cost_model": args.cost_model},
            "roofline": roofline, "gpu_launches": launches, "clocks": clocks,
            "volume_digest": digest,
        }
        if breakdown is not None:
            line["breakdown"] = breakdown
        if e2e is not None:
            line["e2e"] = e2e
        if reference_cuda is not None:
            line["reference_cuda"] = reference_cuda
        if variants is not None:
            line["variants"] = variants
        if contour is not None:
            line["contour"] = contour
        if coloration is not None:
            line["coloration"] = coloration
        if not args.no_cpu_baseline and world == 1:
            rate, desc, _ = cpu_reference_rate(N, V, W, H, target_seconds=12.0)
            line["cpu_baseline"] = dict(desc, value=rate, unit="voxel*views/s")
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def measure_e2e(args, ctx, torch, dev, rank, world, N, V, W, H, K, RT, my_depths, my_cost, slab_cells, units, timed):
    """Host buffers in, host volume out: N=1 through the streaming drop-in of ProcessDepthMap<double>; N>1 through
    dmi_shard_integrate_host (each rank uploads its share of the views group by group, behind the exchange and the
    integration of earlier groups) + dmi_volume_end (each rank reads its finished layers back)."""
    import psutil
    npix = W * H
    nmine = my_depths.shape[0]
    need = 2 * nmine * npix * 8 + slab_cells * 8
    avail = psutil.virtual_memory().available
    if need * world > 0.7 * avail:
        return {"value": None, "unit": "voxel*views/s", "skipped": f"host buffers need {need * world / 1e9:.0f} GB, {avail / 1e9:.0f} GB available"}, None
    h_depths = torch.empty((nmine, H, W), dtype=torch.float64, pin_memory=True)
    h_cost = torch.empty((nmine, H, W), dtype=torch.float64, pin_memory=True)
    h_vol = torch.zeros(slab_cells, dtype=torch.float64, pin_memory=True)
    h_depths.copy_(my_depths)
    h_cost.copy_(my_cost)
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, 2))
    vol_np, d_np, c_np = h_vol.numpy(), h_depths.numpy(), h_cost.numpy()
    if world == 1:
        def step():
            # what the filter's RequestData does (vtkCudaReconstructionFilter.cxx:133-147), as DmiHostClasses.h /
            # reconstruction.py do it: the zero fill of the output happens on the device, the views stream in
            # from host memory, the finished cell scalars are read back into the host array
            ctx.volume_begin(None, np.float64)
            ctx.volume_integrate_host(d_np, c_np, THRESH, K, RT)
            ctx.volume_end(vol_np)
        api = "dmi_volume_begin(NULL) + dmi_volume_integrate_host + dmi_volume_end (host pointers; pinned host buffers)"
    else:
        def step():
            ctx.volume_begin(None, np.float64)
            ctx.shard_integrate_host(V, d_np, c_np, THRESH, K, RT)
            ctx.volume_end(vol_np)
        api = ("per rank: dmi_volume_begin(NULL) + dmi_shard_integrate_host (its share of the views from pinned host memory) + "
               "dmi_volume_end (its z-layers to host memory)")
    h2d = int(2 * nmine * npix * 8)
    step()
    ms, wall = timed(step, steps)
    sec = max(ms, wall) * 1e-3
    extra = {}
    if world == 1:
        # the one-call entry with an io_scalar the caller zero-filled on the host (the reference's calling sequence)
        t0 = time.perf_counter()
        vol_np.fill(0.0)
        t1 = time.perf_counter()
        ctx.process_depth_maps(d_np, c_np, THRESH, K, RT, vol_np)
        t2 = time.perf_counter()
        extra["via_process_depth_maps"] = {"host_zero_fill_ms": (t1 - t0) * 1e3, "call_ms": (t2 - t1) * 1e3,
                                           "note": "io_scalar zero-filled by the caller; the library verifies that on the host instead of uploading it"}
    return ({**extra, "value": units / sec, "unit": "voxel*views/s", "ms_per_step": sec * 1e3, "steps": steps, "warmup": 1,
             "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(slab_cells * 8), "bytes_are": "per rank", "api": api},
            {"depths": h_depths, "cost": h_cost, "vol": h_vol})


if __name__ == "__main__":
    main()
