"""Exploratory: time the resident-tier horizon kernel under several debug-option settings in ONE process
(hzb_debug_option).  Usage: sweep_probe.py --cfg cfg2 --rows 400 "wrefill=20,wwait=3" "horizon_kernel=1" ...
Not part of the product."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import horayzon_b200 as hb
from horayzon_b200 import resident

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--rows", type=int, default=None)
ap.add_argument("--azim", type=int, default=None)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("settings", nargs="*", default=[""])
a = ap.parse_args()
c = hb.synthetic.make_config(a.cfg, a.n)
K = a.azim or c["azim_num"]
sc = resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
ny, nx = c["ny"], c["nx"]
dev = torch.device("cuda:0")
vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
mask = torch.ones((ny, nx), dtype=torch.uint8, device=dev)
rows = a.rows or ny
r0 = (ny - rows) // 2
hori = torch.empty((ny, nx, K), dtype=torch.float32, device=dev)
ref = None
touched = set()
for setting in a.settings:
    resident.debug_option("reset", 0)
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("="); resident.debug_option(k, int(v)); touched.add(k)
    best = None
    for rep in range(a.reps + 1):
        before = sc.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], hori, r0, r0 + rows, dist_search=c["dist_search"])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if rep > 0: best = ms if best is None else min(best, ms)
    after = sc.stats()
    d = {k: after[k] - before[k] for k in ("rays", "node_visits", "prim_tests", "units")}
    chk = hori[r0:r0 + rows].double().sum().item()
    if ref is None: ref = chk
    print("%-44s %8.2f ms  %.4g units/s  nodes/ray %.2f prims/ray %.2f  %s" % (
        setting or "(default)", best, d["units"] / best * 1e3, d["node_visits"] / max(d["rays"], 1),
        d["prim_tests"] / max(d["rays"], 1), "same" if chk == ref else "DIFFERENT OUTPUT"), flush=True)
