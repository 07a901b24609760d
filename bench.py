#!/usr/bin/env python
"""bench.py -- headline benchmark of the horizon hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2]

Metric (BASELINE.json): horizon rays/s = unmasked cells x azimuths per second
(a "ray" is one output element, not one cast; casts per unit are reported
beside it, like the reference prints at horizon_comp.cpp:807-810).

One step = one pass of the hot path over the whole workload: horizon search
for every inner-domain cell + the SVF integral ("horizon+SVF", BASELINE.json
configs[1]).  Default workload: cfg2 = 1201 x 1201 synthetic DEM x 360 azimuths
(SURVEY.md 8d); `--workload cfg4p` is the north-star 6000 x 6000 x 360 line.

  value : DEM + BVH + per-cell inputs resident in HBM before the timed region
          (the reference's "Ray tracing time", horizon_comp.cpp:737-805).
  e2e   : the same metric through the reference-shaped public API
          (horayzon_b200.horizon.horizon_gridded + topo_param.sky_view_factor)
          with HOST buffers: H2D, on-device BVH build, kernels and D2H inside the
          timed region ("Total run time", :816-818).
  N > 1 : strong scaling -- the rows of the same workload are split into N
          contiguous blocks (the reference's own partitioning axis, :739-744),
          one process per GPU, replicated DEM/BVH, one NCCL all-gather of the
          horizon blocks per step inside the timed region.
  --impl reference : the reference's CPU path.  Embree/TBB cannot be installed
          here, so this arm times the CPU oracle (reference-algorithm restatement,
          NOT Embree; OpenMP over rows where the reference uses TBB) on a bounded
          row sample of the same workload, on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "horizon rays/sec (cells x azimuths)"
UNIT = "rays/s"
HORI_ACC = 0.25
ALGORITHM = "guess_constant"


def workload_label(c, K):
    return ("%s: %dx%d synthetic sinusoid DEM (spacing %g m), %d azimuths, %s, hori_acc %.2f deg, "
            "dist_search %g km, inner domain %dx%d, horizon+SVF"
            % (c["name"], c["dem_dim_0"], c["dem_dim_1"], c["spacing"], K, ALGORITHM, HORI_ACC,
               c["dist_search"], c["ny"], c["nx"]))


# --------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, gpu_index=0):
        self.samples, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (ValueError, KeyError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except ValueError:
            pass
    return None


# ---------------------------------------------------------------- CPU oracle
def oracle_sample(c, K, target_s=15.0, rows=None):
    """Time the CPU oracle on a centred block of inner rows (T_rt: BVH build
    excluded, like the reference's 'Ray tracing time')."""
    import oracle
    ny, nx = c["ny"], c["nx"]

    def run(nrows):
        b = (ny - nrows) // 2
        sl = slice(b, b + nrows)
        oracle.horizon_gridded(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["vec_norm"][sl], c["vec_north"][sl],
                               c["offset_0"] + b, c["offset_1"], c["dist_search"], azim_num=K, hori_acc=HORI_ACC,
                               ray_algorithm=ALGORITHM)
        build_s, trace_s = oracle.last_timing()
        return b, nrows * nx * K, trace_s, build_s

    if rows is None:
        _, u, t, _ = run(min(2, ny))
        rows = int(max(2, min(ny, round(2 * target_s / max(t, 1e-3)))))
    b, units, trace_s, build_s = run(rows)
    return {"value": units / trace_s, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "rows": rows,
            "sample": "inner rows %d..%d of %d (%d units), ray tracing %.2f s, BVH build %.2f s excluded; CPU oracle "
                      "= reference-algorithm restatement with OpenMP over rows, NOT Embree+TBB (not installable)"
                      % (b, b + rows, ny, units, trace_s, build_s)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import horayzon_b200.synthetic as syn
    c = syn.make_config(args.workload)
    K = c["azim_num"]
    first = oracle_sample(c, K, target_s=args.ref_seconds)       # sizes the sample
    rows = first["rows"]
    vals = []
    for _ in range(args.warmup):
        oracle_sample(c, K, rows=rows)
    for _ in range(args.steps):
        vals.append(oracle_sample(c, K, rows=rows))
    units = rows * c["nx"] * K
    v = statistics.mean(x["value"] for x in vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": units / v * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(c, K), "sample_rows": rows},
            "cpu_baseline": dict(vals[-1], value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import horayzon_b200 as hb
    from horayzon_b200 import resident, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    c = hb.synthetic.make_config(args.workload)
    K = c["azim_num"]
    ny, nx = c["ny"], c["nx"]
    tilt_np = hb.synthetic.tilt_vectors(c["x"], c["y"], c["z"], c["offset_0"])
    azim_np = np.array([(2 * np.pi) / K * i for i in range(K)], np.float32)

    # ---- resident state (outside the timed region): DEM + BVH + per-cell inputs
    scene = resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], device=local_rank)
    vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
    mask = torch.ones((ny, nx), dtype=torch.uint8, device=dev)
    tilt = torch.from_numpy(tilt_np).to(dev); azim = torch.from_numpy(azim_np).to(dev)
    shards = sharding.row_shards(ny, world)
    b, e = shards[rank]
    per = sharding.padded_rows(ny, world)
    # full-size output: this rank fills rows [b, e); the all-gather fills the rest
    hori = torch.empty((world * per, nx, K), dtype=torch.float32, device=dev)
    svf = torch.empty((world * per, nx), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > L2 (126 MB)
    stream = torch.cuda.current_stream()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def step(i_timed=None):
        flush.zero_()  # L2 flush between iterations (inside the bracket; ~0.05 ms)
        if i_timed is not None:
            k_ev[i_timed][0].record(stream)
        scene.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], hori[:ny], b, e,
                              dist_search=c["dist_search"], hori_acc=HORI_ACC, ray_algorithm=ALGORITHM, stream=stream)
        if i_timed is not None:
            k_ev[i_timed][1].record(stream)
        if e > b:
            resident.sky_view_factor_dev(azim, hori[b:e], tilt[b:e], svf[b:e], stream=stream)
        if world > 1:  # single exchange step: all-gather of the row blocks (in place)
            dist.all_gather_into_tensor(hori, hori[rank * per:(rank + 1) * per])
            dist.all_gather_into_tensor(svf, svf[rank * per:(rank + 1) * per])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    before = scene.stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record(stream)
    for i in range(args.steps):
        step(i)
    t1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t0.elapsed_time(t1)
    kern_ms = sum(a.elapsed_time(z) for a, z in k_ev) / args.steps
    after = scene.stats()
    stat = torch.tensor([total_ms, kern_ms, float(after["rays"] - before["rays"]),
                         float(after["node_visits"] - before["node_visits"]),
                         float(after["prim_tests"] - before["prim_tests"]),
                         float(after["warp_node_visits"] - before["warp_node_visits"])], dtype=torch.float64, device=dev)
    smax = stat.clone()
    if world > 1:
        dist.all_reduce(smax, op=dist.ReduceOp.MAX)
        ssum = stat.clone(); dist.all_reduce(ssum, op=dist.ReduceOp.SUM)
    else:
        ssum = stat
    total_ms, kern_ms_max = float(smax[0]), float(smax[1])
    units_step = ny * nx * K
    value = units_step * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the public (reference-shaped) API with host buffers
    sl = slice(b, e)
    e2e_times = []
    h2d = d2h = 0
    h_host = svf_host = None
    for it in range(1 + args.e2e_steps if args.e2e_steps > 0 else 0):
        h_host = svf_host = None   # a loop that consumes each result before asking for the next one
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        if e > b:
            h_host, a_host = hb.horizon.horizon_gridded(
                c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["vec_norm"][sl], c["vec_north"][sl],
                c["offset_0"] + b, c["offset_1"], c["dist_search"], azim_num=K, hori_acc=HORI_ACC,
                ray_algorithm=ALGORITHM)
            svf_host = hb.topo_param.sky_view_factor(a_host, h_host, tilt_np[sl])
            h2d = (c["dem_dim_0"] * c["dem_dim_1"] * 12 + (e - b) * nx * 25) + (h_host.nbytes + tilt_np[sl].nbytes + a_host.nbytes)
            d2h = h_host.nbytes + svf_host.nbytes
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it > 0:  # first call is warm-up (context, allocator)
            e2e_times.append(float(dt[0]))
    e2e_s = statistics.median(e2e_times) if e2e_times else float('nan')

    if rank == 0:
        peak, peak_src = measured_peak()
        algo_bytes_per_unit = 4.0 + 37.0 / K          # compulsory traffic
        # dominant kernel = k_horizon_wq6 (launched by hzb_horizon_gridded_dev); at N > 1 each rank launches it on units/N
        achieved = algo_bytes_per_unit * (units_step / world) / (kern_ms_max * 1e-3) / 1e9
        rays = float(ssum[2])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(c, K),
                       "parallelism": "rows x%d, replicated DEM+BVH, 1 NCCL all-gather/step" % world if world > 1 else "single GPU",
                       "l2": "256 MB flush write before every step; per-step output %.2f GB >> 126 MB L2" % (units_step * 4 / 1e9)},
            "clocks": clocks,
            "e2e": {"value": units_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3,
                    "api": "horayzon_b200.horizon.horizon_gridded + topo_param.sky_view_factor (host ndarray in/out, "
                           "inputs pageable like the reference's; the returned horizon array is page-locked from the second large call "
                           "of a process on; H2D + BVH build + kernels + D2H timed; median of %d calls after one warm-up call, "
                           "each result released before the next call)" % len(e2e_times)},
            "gpu_launches": int(2 * args.steps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(args.workload), "peak_source": peak_src,
                         "kernel": "k_horizon_wq6", "kernel_ms": kern_ms_max,
                         "algorithmic_bytes_per_unit": algo_bytes_per_unit,
                         "note": "compulsory bytes only (output store + per-cell inputs); BVH traversal is "
                                 "latency/L2-bound, see DESIGN.md"},
            "counters": {"casts_per_unit": rays / (units_step * args.steps),
                         "nodes_per_cast": float(ssum[3]) / max(rays, 1.0), "prims_per_cast": float(ssum[4]) / max(rays, 1.0),
                         "warp_nodes_per_cast": float(ssum[5]) / max(rays, 1.0)},
            "bvh": {"prims": int(after["num_prims"]), "bytes": int(after["bvh_bytes"]), "build_s": after["t_build"],
                    "h2d_s": after["t_h2d"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = oracle_sample(c, K, target_s=args.ref_seconds)
        print(json.dumps(line), flush=True)
    scene.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
