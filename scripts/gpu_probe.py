"""Exploratory GPU probe: time the resident-tier horizon kernel on BASELINE
configs (optionally a row subset) and print counters.  Not part of the product."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import horayzon_b200 as hb
from horayzon_b200 import resident

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--rows", type=int, default=None, help="number of inner rows to compute (centred)")
ap.add_argument("--azim", type=int, default=None)
ap.add_argument("--alg", default="guess_constant")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()

t0 = time.time()
c = hb.synthetic.make_config(a.cfg, a.n)
K = a.azim or c["azim_num"]
print("dem built %.1fs" % (time.time() - t0), flush=True)
t0 = time.time()
sc = resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
st = sc.stats()
print("scene: h2d %.3fs build %.3fs prims %d bvh %.1f MB (wall %.2fs)" % (st["t_h2d"], st["t_build"], st["num_prims"], st["bvh_bytes"] / 1e6, time.time() - t0), flush=True)
ny, nx = c["ny"], c["nx"]
dev = torch.device("cuda:0")
vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
mask = torch.ones((ny, nx), dtype=torch.uint8, device=dev)
rows = a.rows or ny
r0 = (ny - rows) // 2
hori = torch.empty((ny, nx, K), dtype=torch.float32, device=dev)
for rep in range(a.reps):
    before = sc.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], hori, r0, r0 + rows, dist_search=c["dist_search"], ray_algorithm=a.alg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    after = sc.stats()
    d = {k: after[k] - before[k] for k in ("rays", "node_visits", "prim_tests", "units", "warp_node_visits")}
    u = max(d["units"], 1)
    print(json.dumps(dict(cfg=a.cfg, n=c["dem_dim_0"], rows=rows, azim=K, ms=round(ms, 2), units_per_s=u / ms * 1e3,
                          rays_per_unit=d["rays"] / u, nodes_per_ray=d["node_visits"] / max(d["rays"], 1),
                          prims_per_ray=d["prim_tests"] / max(d["rays"], 1),
                          warp_nodes_per_ray=d["warp_node_visits"] / max(d["rays"], 1))), flush=True)
print("hori range", float(hori[r0:r0 + rows].min()), float(hori[r0:r0 + rows].max()))
