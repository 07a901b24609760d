"""Time Terrain.shadow / sw_dir_cor per sun position on a synthetic DEM (default cfg3 3601^2)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import horayzon_b200 as hb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3601
nsun = int(sys.argv[2]) if len(sys.argv) > 2 else 16
refrac = len(sys.argv) > 3 and sys.argv[3] == "refrac"
c = hb.synthetic.CONFIGS["cfg3"]
x, y, z = hb.synthetic.sinusoid_dem(n, n, c["spacing"], c["amp"], c["wavelength"], c["seed"], c["octaves"])
rim = 1
tilt = hb.synthetic.tilt_vectors(x, y, z, rim)
ny = nx = n - 2
norm, _ = hb.synthetic.planar_frames(ny, nx)
enl = (1.0 / (norm * tilt).sum(axis=2)).astype(np.float32)
elev = np.ascontiguousarray(z[1:-1, 1:-1]); mask = np.ones((ny, nx), np.uint8)
vg = hb.synthetic.rearrange_pad_buffer(x, y, z)
t = hb.shadow.Terrain()
t0 = time.perf_counter(); t.initialise(vg, n, n, rim, rim, tilt, norm, enl, elev, mask, refrac_cor=refrac); print("initialise %.3fs" % (time.perf_counter() - t0))
suns = hb.synthetic.sun_positions_diurnal(nsun)
buf = np.empty((ny, nx), np.uint8); fb = np.empty((ny, nx), np.float32)
t.shadow(suns[0], buf)
t0 = time.perf_counter()
for s in suns: t.shadow(s, buf)
dt = time.perf_counter() - t0
print("shadow: %.2f ms/sun, %.3e cells/s (host API, incl. D2H)" % (dt / nsun * 1e3, ny * nx * nsun / dt))
t0 = time.perf_counter(); out = t.shadow_batch(suns); dt = time.perf_counter() - t0
print("shadow_batch: %.2f ms/sun, %.3e cells/s; lit frac %.3f self %.3f terrain %.3f" % (dt / nsun * 1e3, ny * nx * nsun / dt, (out == 0).mean(), (out == 1).mean(), (out == 2).mean()))
t0 = time.perf_counter()
for s in suns: t.sw_dir_cor(s, fb)
dt = time.perf_counter() - t0
print("sw_dir_cor: %.2f ms/sun, %.3e cells/s" % (dt / nsun * 1e3, ny * nx * nsun / dt))
