"""Scale check: build a 12000 x 12000 scene (144 M primitives) and compute a few rows; compare a
sub-block against the same cells computed on a cropped DEM (identical geometry within dist_search)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import horayzon_b200 as hb
from horayzon_b200 import resident
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
x, y, z = hb.synthetic.sinusoid_dem(n, n, 2.0, 400.0, 3000.0, 4, 6)
vg = hb.synthetic.rearrange_pad_buffer(x, y, z)
t0 = time.time(); sc = resident.Scene(vg, n, n); st = sc.stats()
print("scene %d^2: prims %d nodes %d bvh %.2f GB build %.3fs h2d %.3fs (wall %.1fs)" % (n, st["num_prims"], st["num_nodes"], st["bvh_bytes"] / 1e9, st["t_build"], st["t_h2d"], time.time() - t0), flush=True)
dev = torch.device("cuda:0")
rim = 1; ny = nx = n - 2; K = 24
vn_np, vno_np = hb.synthetic.planar_frames(8, nx)
vn = torch.from_numpy(vn_np).to(dev); vno = torch.from_numpy(vno_np).to(dev)
mask = torch.ones((8, nx), dtype=torch.uint8, device=dev)
hori = torch.empty((8, nx, K), dtype=torch.float32, device=dev)
r0 = n // 2
# inner "domain" = 8 rows starting at DEM row r0 (offset_0 = r0), all columns
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sc.horizon_gridded(vn, vno, mask, r0, rim, hori, 0, 8, dist_search=0.5); e1.record(); torch.cuda.synchronize()
print("8 rows x %d cols x %d az: %.1f ms; range %.4f..%.4f" % (nx, K, e0.elapsed_time(e1), float(hori.min()), float(hori.max())), flush=True)
big = hori[:, 5000:5064].cpu().numpy()
sc.close()
# same cells from a cropped DEM (dist_search 0.5 km = 250 cells: crop margin 300 cells)
m = 300; c0 = rim + 5000
xs, ys, zs = (a[r0 - m:r0 + 8 + m, c0 - m:c0 + 64 + m].copy() for a in (x, y, z))
vn2, vno2 = hb.synthetic.planar_frames(8, 64)
h2, _ = hb.horizon.horizon_gridded(hb.synthetic.rearrange_pad_buffer(xs, ys, zs), xs.shape[0], xs.shape[1], vn2, vno2, m, m, 0.5, azim_num=K)
print("crop vs big: identical =", np.array_equal(h2, big), "max|d| = %.3e" % np.abs(h2 - big).max())
