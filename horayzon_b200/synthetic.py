"""Synthetic DEMs and the input wire format of the hot path.

No DEM files or network exist in the build environment, so every test and the
benchmark use the synthetic planar DEMs specified in SURVEY.md section 8(d).
``rearrange_pad_buffer`` / ``pad_buffer`` reproduce the reference's flat vertex
buffer layout (``horayzon/auxiliary.py:49-133``): float32 ``[y][x][xyz]`` plus
at least 16 trailing zeros.
"""
import numpy as np


def pad_buffer(buffer):
    """Pad a 1-D geometry buffer with >= 16 zeros (auxiliary.py:100-133)."""
    if not isinstance(buffer, np.ndarray):
        raise ValueError("argument 'buffer' has invalid type")
    if buffer.ndim != 1:
        raise ValueError("argument 'buffer' must be one-dimensional")
    extra = 16
    rem = buffer.nbytes % 16
    if rem != 0:
        extra += (16 - rem) // buffer.itemsize
    return np.concatenate([buffer, np.zeros(extra, dtype=buffer.dtype)])


def rearrange_pad_buffer(x, y, z):
    """Interleave x, y, z (2-D float32) into the padded vertex buffer
    (auxiliary.py:49-95)."""
    for a in (x, y, z):
        if not isinstance(a, np.ndarray):
            raise TypeError("One or more input arguments are of invalid type")
    if (x.dtype != np.float32) or (y.dtype != np.float32) or (z.dtype != np.float32):
        raise TypeError("Not all input arguments are 32-bit floats")
    if any(a.ndim != 2 for a in (x, y, z)) or not (x.shape == y.shape == z.shape):
        raise ValueError("Dimensions of input arguments are erroneous/inconsistent")
    buf = np.empty(x.size * 3, dtype=np.float32)
    buf[0::3] = x.ravel()
    buf[1::3] = y.ravel()
    buf[2::3] = z.ravel()
    return pad_buffer(buf)


def sinusoid_dem(ny, nx, spacing, amplitude, wavelength, seed, octaves=6):
    """z = sum_o A 2^-o sin(2 pi 2^o x / L + phi_o) sin(2 pi 2^o y / L + psi_o)
    on a planar grid x = j*spacing, y = i*spacing."""
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0.0, 2.0 * np.pi, octaves)
    psi = rng.uniform(0.0, 2.0 * np.pi, octaves)
    xs = np.arange(nx, dtype=np.float64) * spacing
    ys = np.arange(ny, dtype=np.float64) * spacing
    z = np.zeros((ny, nx), dtype=np.float64)
    for o in range(octaves):
        f = 2.0 * np.pi * (2.0 ** o) / wavelength
        z += (amplitude * 2.0 ** (-o)) * np.outer(np.sin(f * ys + psi[o]), np.sin(f * xs + phi[o]))
    x2, y2 = np.meshgrid(xs.astype(np.float32), ys.astype(np.float32))
    return x2, y2, z.astype(np.float32)


def sinusoid_vert_grid(n, spacing, amplitude, wavelength, seed, octaves=6, block_rows=1024):
    """The padded vertex buffer of ``sinusoid_dem`` + ``rearrange_pad_buffer`` built row block by row block, without
    the full-size x / y / z (and float64) arrays: for the 24001 x 24001 DEM of BASELINE configs[4] that is 7 GB
    per process instead of ~25 GB.  Bit-identical to the two-step route."""
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0.0, 2.0 * np.pi, octaves)
    psi = rng.uniform(0.0, 2.0 * np.pi, octaves)
    xs = np.arange(n, dtype=np.float64) * spacing
    ys = np.arange(n, dtype=np.float64) * spacing
    buf = np.zeros(n * n * 3 + 16, dtype=np.float32)          # 16 trailing zeros (n*n*3*4 bytes is a multiple of 16 when n*n*3 % 4 == 0)
    extra = 16
    rem = (n * n * 3 * 4) % 16
    if rem != 0:
        extra += (16 - rem) // 4
        buf = np.zeros(n * n * 3 + extra, dtype=np.float32)
    v = buf[:n * n * 3].reshape(n, n, 3)
    xf = xs.astype(np.float32)
    for r0 in range(0, n, block_rows):
        r1 = min(n, r0 + block_rows)
        z = np.zeros((r1 - r0, n), dtype=np.float64)
        for o in range(octaves):
            f = 2.0 * np.pi * (2.0 ** o) / wavelength
            z += (amplitude * 2.0 ** (-o)) * np.outer(np.sin(f * ys[r0:r1] + psi[o]), np.sin(f * xs + phi[o]))
        v[r0:r1, :, 0] = xf[None, :]
        v[r0:r1, :, 1] = ys[r0:r1].astype(np.float32)[:, None]
        v[r0:r1, :, 2] = z.astype(np.float32)
    return buf


def planar_frames(ny, nx):
    """vec_norm = (0,0,1), vec_north = (0,1,0) for every inner cell
    (examples/horizon/gridded_planar_DEM.py:71-76)."""
    vec_norm = np.zeros((ny, nx, 3), dtype=np.float32)
    vec_norm[:, :, 2] = 1.0
    vec_north = np.zeros((ny, nx, 3), dtype=np.float32)
    vec_north[:, :, 1] = 1.0
    return vec_norm, vec_north


# name -> DEM recipe of SURVEY.md 8(d).  `rim` = cells removed on every side to
# form the inner domain; `dist` in km.
CONFIGS = {
    "cfg1": dict(n=128, spacing=100.0, amp=300.0, wavelength=3200.0, seed=1, octaves=1,
                 rim=16, azim_num=36, dist=5.0),
    "cfg2": dict(n=1201, spacing=90.0, amp=1200.0, wavelength=40000.0, seed=2, octaves=6,
                 rim=1, azim_num=360, dist=50.0),
    "cfg3": dict(n=3601, spacing=30.0, amp=1500.0, wavelength=30000.0, seed=3, octaves=6,
                 rim=1, azim_num=288, dist=50.0),
    "cfg4": dict(n=6000, spacing=2.0, amp=400.0, wavelength=3000.0, seed=4, octaves=6,
                 rim=1, azim_num=720, dist=12.0),
    "cfg4p": dict(n=6000, spacing=2.0, amp=400.0, wavelength=3000.0, seed=4, octaves=6,
                  rim=1, azim_num=360, dist=12.0),
    "cfg5": dict(n=24001, spacing=90.0, amp=2000.0, wavelength=80000.0, seed=5, octaves=6,
                 rim=1, azim_num=180, dist=50.0),
}


def make_config(name, n=None):
    """Build the synthetic inputs of one BASELINE config (optionally shrunk to
    an n x n grid for parity tests).  Returns a dict with vert_grid, dims,
    offsets, frames and the horizon parameters."""
    c = dict(CONFIGS[name])
    if n is not None:
        c["n"] = int(n)
    n = c["n"]
    x, y, z = sinusoid_dem(n, n, c["spacing"], c["amp"], c["wavelength"], c["seed"], c["octaves"])
    rim = c["rim"]
    ny = nx = n - 2 * rim
    vec_norm, vec_north = planar_frames(ny, nx)
    return dict(name=name, x=x, y=y, z=z, vert_grid=rearrange_pad_buffer(x, y, z),
                dem_dim_0=n, dem_dim_1=n, offset_0=rim, offset_1=rim, ny=ny, nx=nx,
                vec_norm=vec_norm, vec_north=vec_north, azim_num=c["azim_num"],
                dist_search=c["dist"], spacing=c["spacing"])


def tilt_vectors(x, y, z, rim):
    """Unit surface normals of the inner domain from central differences on a
    planar grid (stand-in for the reference's slope_plane_meth, which is input
    preparation outside the hot path: SURVEY.md section 2 row 6)."""
    dzdx = (z[1:-1, 2:].astype(np.float64) - z[1:-1, :-2]) / (x[1:-1, 2:].astype(np.float64) - x[1:-1, :-2])
    dzdy = (z[2:, 1:-1].astype(np.float64) - z[:-2, 1:-1]) / (y[2:, 1:-1].astype(np.float64) - y[:-2, 1:-1])
    v = np.stack([-dzdx, -dzdy, np.ones_like(dzdx)], axis=2)
    v /= np.linalg.norm(v, axis=2, keepdims=True)
    r = rim - 1
    if r > 0:
        v = v[r:-r, r:-r]
    return np.ascontiguousarray(v.astype(np.float32))


def sun_positions_diurnal(num, dist=1.5e11, elev_min_deg=-10.0, elev_max_deg=45.0):
    """`num` sun positions on one diurnal circle (cfg 3): azimuth
    sweeps 360 degrees, elevation follows a cosine arc between the limits."""
    t = np.arange(num, dtype=np.float64) / num
    az = 2.0 * np.pi * t
    mid = 0.5 * (elev_max_deg + elev_min_deg)
    amp = 0.5 * (elev_max_deg - elev_min_deg)
    el = np.deg2rad(mid - amp * np.cos(2.0 * np.pi * t))
    pos = np.stack([dist * np.cos(el) * np.sin(az), dist * np.cos(el) * np.cos(az),
                    dist * np.sin(el)], axis=1)
    return pos.astype(np.float32)
