"""BASELINE configs[4] (24001 x 24001 stitched SRTM-like DEM, 576 M quads, 180 azimuths, dist_search 50 km) on ONE GPU:
build the scene, compute row slabs, and compare a block of cells against the same cells computed on a cropped DEM
(identical geometry within dist_search, a different BVH and a different quantisation grid: decisions do not depend
on either).  Prints one JSON line.  Usage: python scripts/cfg5_check.py [n] [slab_rows]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import horayzon_b200 as hb
from horayzon_b200 import resident

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24001
slab = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = hb.synthetic.CONFIGS["cfg5"]
K, dist = cfg["azim_num"], cfg["dist"]
t0 = time.time()
x, y, z = hb.synthetic.sinusoid_dem(n, n, cfg["spacing"], cfg["amp"], cfg["wavelength"], cfg["seed"], cfg["octaves"])
vg = hb.synthetic.rearrange_pad_buffer(x, y, z)
t_dem = time.time() - t0
t0 = time.time(); sc = resident.Scene(vg, n, n); st = sc.stats(); t_scene = time.time() - t0
out = {"dem": "%dx%d" % (n, n), "prims": st["num_prims"], "wide_nodes": st["num_nodes"], "bvh_GB": st["bvh_bytes"] / 1e9,
       "build_s": st["t_build"], "h2d_s": st["t_h2d"], "scene_wall_s": t_scene, "dem_gen_s": t_dem,
       "gpu_mem_GB_after_build": torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9}
print(json.dumps(out), flush=True)
dev = torch.device("cuda:0")
nx = n - 2
r0 = n // 2

def rows(nrows, row0):
    vn_np, vno_np = hb.synthetic.planar_frames(nrows, nx)
    vn = torch.from_numpy(vn_np).to(dev); vno = torch.from_numpy(vno_np).to(dev)
    mask = torch.ones((nrows, nx), dtype=torch.uint8, device=dev)
    hori = torch.empty((nrows, nx, K), dtype=torch.float32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = sc.stats()
    e0.record(); sc.horizon_gridded(vn, vno, mask, row0, 1, hori, 0, nrows, dist_search=dist); e1.record(); torch.cuda.synchronize()
    after = sc.stats()
    return hori, e0.elapsed_time(e1), after["rays"] - before["rays"], after["fallback_packets"]

h8, ms8, _, _ = rows(8, r0)
hs, ms, rays, fb = rows(slab, r0 - slab // 2)
units = slab * nx * K
out.update({"slab_rows": slab, "slab_ms": ms, "rays_per_s": units / (ms * 1e-3), "casts_per_unit": rays / units, "fallback_packets": fb})
big = h8[:, 12000:12064].cpu().numpy() if n > 13000 else h8[:, nx // 2:nx // 2 + 64].cpu().numpy()
c_first = 1 + (12000 if n > 13000 else nx // 2)
sc.close(); del h8, hs
torch.cuda.empty_cache()
m = int(dist * 1000 / cfg["spacing"]) + 50      # crop margin: dist_search + 50 cells
lo_r, hi_r = max(0, r0 - m), min(n, r0 + 8 + m)
lo_c, hi_c = max(0, c_first - m), min(n, c_first + 64 + m)
xs, ys, zs = (np.ascontiguousarray(a[lo_r:hi_r, lo_c:hi_c]) for a in (x, y, z))
vn2, vno2 = hb.synthetic.planar_frames(8, 64)
h2, _ = hb.horizon.horizon_gridded(hb.synthetic.rearrange_pad_buffer(xs, ys, zs), xs.shape[0], xs.shape[1], vn2, vno2,
                                   r0 - lo_r, c_first - lo_c, dist, azim_num=K)
out.update({"crop": "%dx%d" % xs.shape, "crop_vs_full_identical": bool(np.array_equal(h2, big)), "crop_max_abs_diff": float(np.abs(h2 - big).max())})
# a few of those cells against the CPU oracle on the cropped DEM
import oracle
oracle.set_num_threads(os.cpu_count() or 1)
osc = oracle.Scene(hb.synthetic.rearrange_pad_buffer(xs, ys, zs), xs.shape[0], xs.shape[1])
vnf, vnof = hb.synthetic.planar_frames(xs.shape[0] - (r0 - lo_r), 64)
ho = osc.horizon_rows([0, 7], vnf, vnof, r0 - lo_r, c_first - lo_c, dist, azim_num=K)
osc.close()
out["oracle_rows_identical"] = bool(np.array_equal(ho, big[[0, 7]]))
print(json.dumps(out), flush=True)
