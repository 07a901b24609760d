"""Host-tier timing breakdown (hzb_get_stats) for one workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import horayzon_b200 as hb
from horayzon_b200 import resident
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
c = hb.synthetic.make_config(cfg)
args = (c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["vec_norm"], c["vec_north"], c["offset_0"], c["offset_1"], c["dist_search"])
tilt = hb.synthetic.tilt_vectors(c["x"], c["y"], c["z"], c["offset_0"])
for it in range(3):
    t0 = time.perf_counter()
    h, az = hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"])
    t1 = time.perf_counter()
    svf = hb.topo_param.sky_view_factor(az, h, tilt)
    t2 = time.perf_counter()
    st = resident.last_stats()
    print("iter %d: horizon_gridded %.3fs (h2d %.3f build %.3f trace %.3f d2h %.3f total-native %.3f) svf %.3fs" % (
        it, t1 - t0, st["t_h2d"], st["t_build"], st["t_trace"], st["t_d2h"], st["t_total"], t2 - t1), flush=True)
