"""Exploratory: kernel time of ONE rank's share of a workload (shard 0 of N, block-interleaved) on one GPU -- what each
GPU of an N-GPU strong-scaling run executes -- under several settings of a debug option (default: wwait 2 vs 3).
Not part of the product."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import horayzon_b200 as hb
from horayzon_b200 import resident, sharding

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--shards", default="1,2,4,8")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--option", default="wwait")
ap.add_argument("--values", default="2,3")
ap.add_argument("--set", default="", help="further options for every run: name=value,name=value")
ap.add_argument("--all", action="store_true", help="print every repetition")
a = ap.parse_args()
c = hb.synthetic.make_config(a.cfg)
K, ny, nx = c["azim_num"], c["ny"], c["nx"]
dev = torch.device("cuda:0")
sc = resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
mask = torch.ones((ny, nx), dtype=torch.uint8, device=dev)
ref = None
for n in [int(x) for x in a.shards.split(",")]:
    rows = sharding.shard_block_rows(ny, 0, n)
    out = torch.empty((rows, nx, K), dtype=torch.float32, device=dev)
    res = {}
    modes = [int(v) for v in a.values.split(",")]
    for mode in modes:
        resident.debug_option("reset", 0); resident.debug_option(a.option, mode)
        for kv in [x for x in a.set.split(",") if x]:
            resident.debug_option(kv.split("=")[0], int(kv.split("=")[1]))
        best = None; every = []
        for rep in range(a.reps + 1):
            out.fill_(float("nan"))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            sc.horizon_gridded_sharded(vn, vno, mask, c["offset_0"], c["offset_1"], out, 0, n, K, packed=True, dist_search=c["dist_search"])
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if rep > 0: best = ms if best is None else min(best, ms); every.append(round(ms, 1))
        if a.all: print("   %s=%d: %s" % (a.option, mode, every), flush=True)
        res[mode] = (best, out.clone())
    real = min(rows, ny)     # the last packed row of an odd row count is padding
    same = all(torch.equal(res[modes[0]][1][:real - 3], res[m][1][:real - 3]) for m in modes[1:])
    st = sc.stats()
    print("shard 0 of %d: %6d rows  %s  outputs identical: %s  fallbacks %d  segment tasks %d recomputed %d" % (
        n, rows, "  ".join("%s=%d %8.2f ms" % (a.option, m, res[m][0]) for m in modes), same,
        st["fallback_packets"], st["segment_tasks"], st["segment_redos"]), flush=True)
resident.debug_option("reset", 0)
