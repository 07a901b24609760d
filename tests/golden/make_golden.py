"""Generate the committed golden fixtures.  Run in the BUILD container only:

    python oracle/build_ref.py && python tests/golden/make_golden.py

* ``integrals_ref.npz`` -- inputs and outputs of the UNMODIFIED reference
  ``topo_param.sky_view_factor / visible_sky_fraction / topographic_openness``
  (compiled from /root/reference by oracle/build_ref.py).  These pin the oracle's
  restatement of topo_param.pyx:412-603.
* ``slope_ref.npz`` / ``transform_ref.npz`` -- likewise for the UNMODIFIED reference
  ``topo_param.slope_*`` and ``transform`` / ``direction`` (lonlat2ecef, ecef2enu,
  ecef2enu_vector, wgs2swiss, swiss2wgs, rotation_matrix_glob2loc, surf_norm, north_dir).
* ``horizon_oracle_regression.npz`` -- outputs of the CPU ORACLE (not of the
  reference: Embree is unavailable, the ray path is "parity unpinned") for a
  small seeded DEM and all three search algorithms.  A regression anchor only.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def integral_inputs(seed, ny, nx, K):
    rng = np.random.default_rng(seed)
    azim = np.array([2 * np.pi / K * i for i in range(K)], np.float32)
    hori = rng.uniform(-0.1, 0.7, (ny, nx, K)).astype(np.float32)
    sl = np.deg2rad(rng.uniform(0, 55, (ny, nx)))
    asp = rng.uniform(0, 2 * np.pi, (ny, nx))
    tilt = np.stack([np.sin(sl) * np.sin(asp), np.sin(sl) * np.cos(asp), np.cos(sl)], axis=2).astype(np.float32)
    return azim, hori, tilt


def slope_inputs(n=40, seed=9):
    """Seeded DEM in a slightly rotated 'global' frame + per-cell rotation matrices."""
    rng = np.random.default_rng(seed)
    xs = np.arange(n, dtype=np.float64) * 30.0
    X, Y = np.meshgrid(xs, xs)
    Z = 200.0 * np.sin(X / 300.0) * np.cos(Y / 250.0) + rng.normal(0, 1.0, X.shape)
    # rotate everything by small angles so that local up != z (as in global ENU far from the origin)
    a, b = 0.03, -0.02
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    G = Rx @ Ry
    P = np.stack([X, Y, Z], axis=-1) @ G.T
    rot = np.broadcast_to(G.T.astype(np.float32), (n, n, 3, 3)).copy()   # global -> local
    return (np.ascontiguousarray(P[..., 0].astype(np.float32)), np.ascontiguousarray(P[..., 1].astype(np.float32)),
            np.ascontiguousarray(P[..., 2].astype(np.float32)), rot)


def transform_inputs(seed=11, ny=9, nx=13):
    """A small lon/lat grid around the Alps plus elevations (float32)."""
    rng = np.random.default_rng(seed)
    lon = np.linspace(6.0, 11.5, nx) + rng.uniform(-0.01, 0.01, nx)
    lat = np.linspace(48.2, 44.9, ny) + rng.uniform(-0.01, 0.01, ny)      # north -> south like a DEM
    lon2, lat2 = np.meshgrid(lon, lat)
    h = rng.uniform(-5.0, 4500.0, (ny, nx)).astype(np.float32)
    return np.ascontiguousarray(lon2), np.ascontiguousarray(lat2), h


def main():
    br = _load(os.path.join(ROOT, "oracle", "build_ref.py"), "build_ref")
    br.build()
    tp, tr, di = br.load()
    # transform / direction: the chain of examples/horizon/gridded_curved_DEM.py:60-90 for each ellipsoid
    lon, lat, h = transform_inputs()
    tf = {}
    for el in ("sphere", "GRS80", "WGS84"):
        x, y, z = tr.lonlat2ecef(lon, lat, h, ellps=el)
        t = tr.TransformerEcef2enu(lon_or=float(lon.mean()), lat_or=float(lat.mean()), ellps=el)
        xe, ye, ze = tr.ecef2enu(x, y, z, t)
        nrm = di.surf_norm(lon, lat)
        nth = di.north_dir(x, y, z, nrm, ellps=el)
        nrm_e, nth_e = tr.ecef2enu_vector(nrm, t), tr.ecef2enu_vector(nth, t)
        rot = tr.rotation_matrix_glob2loc(nth_e, nrm_e)
        for k, v in dict(x=x, y=y, z=z, xe=xe, ye=ye, ze=ze, nrm=nrm, nth=nth, nrm_e=nrm_e, nth_e=nth_e, rot=rot,
                         orig=np.array([t.x_ecef_or, t.y_ecef_or, t.z_ecef_or, t.lon_or, t.lat_or])).items():
            tf[el + "_" + k] = np.asarray(v)
    e, n, hc = tr.wgs2swiss(lon, lat, h)
    lo2, la2, hw = tr.swiss2wgs(e, n, hc)
    tf.update(swiss_e=np.asarray(e), swiss_n=np.asarray(n), swiss_h=np.asarray(hc), back_lon=np.asarray(lo2),
              back_lat=np.asarray(la2), back_h=np.asarray(hw))
    np.savez_compressed(os.path.join(HERE, "transform_ref.npz"), **tf)

    out = {}
    for tag, (seed, ny, nx, K) in {"a": (0, 12, 16, 72), "b": (1, 6, 8, 360)}.items():
        azim, hori, tilt = integral_inputs(seed, ny, nx, K)
        out[tag + "_shape"] = np.array([seed, ny, nx, K])
        out[tag + "_svf"] = np.asarray(tp.sky_view_factor(azim, hori, tilt))
        out[tag + "_vsf"] = np.asarray(tp.visible_sky_fraction(azim, hori, tilt))
        out[tag + "_top"] = np.asarray(tp.topographic_openness(azim, hori))
    np.savez_compressed(os.path.join(HERE, "integrals_ref.npz"), **out)

    # slope: unmodified reference slope_plane_meth / slope_vector_meth on a seeded DEM,
    # with identity and with per-cell rotation matrices (curved-grid style)
    x, y, z, rot = slope_inputs()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        sl = {"plane_id": np.asarray(tp.slope_plane_meth(x, y, z)),
              "plane_rot": np.asarray(tp.slope_plane_meth(x, y, z, rot_mat=rot, output_rot=False)),
              "plane_rot_out": np.asarray(tp.slope_plane_meth(x, y, z, rot_mat=rot, output_rot=True)),
              "vector_id": np.asarray(tp.slope_vector_meth(x, y, z)),
              "vector_rot_out": np.asarray(tp.slope_vector_meth(x, y, z, rot_mat=rot, output_rot=True))}
    np.savez_compressed(os.path.join(HERE, "slope_ref.npz"), **sl)

    import oracle
    syn = _load(os.path.join(ROOT, "horayzon_b200", "synthetic.py"), "synthetic")
    c = syn.make_config("cfg1", n=64)
    args = (c["vert_grid"], 64, 64, c["vec_norm"], c["vec_north"], 16, 16, 5.0)
    reg = {}
    for alg in ("guess_constant", "binary_search", "discrete_sampling"):
        h, _, rays = oracle.horizon_gridded(*args, azim_num=24, ray_algorithm=alg, return_rays=True)
        reg[alg] = h
        reg[alg + "_rays"] = np.array([rays])
    np.savez_compressed(os.path.join(HERE, "horizon_oracle_regression.npz"), **reg)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
