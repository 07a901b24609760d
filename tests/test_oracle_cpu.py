"""CPU tests of the oracle: golden vectors from the compiled reference, analytic
known-answer cases, cross-algorithm consistency.  No GPU."""
import importlib.util
import os

import numpy as np
import pytest

import oracle
from horayzon_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _integral_inputs(seed, ny, nx, K):  # must match tests/golden/make_golden.py
    rng = np.random.default_rng(seed)
    azim = np.array([2 * np.pi / K * i for i in range(K)], np.float32)
    hori = rng.uniform(-0.1, 0.7, (ny, nx, K)).astype(np.float32)
    sl = np.deg2rad(rng.uniform(0, 55, (ny, nx)))
    asp = rng.uniform(0, 2 * np.pi, (ny, nx))
    tilt = np.stack([np.sin(sl) * np.sin(asp), np.sin(sl) * np.cos(asp), np.cos(sl)], axis=2).astype(np.float32)
    return azim, hori, tilt


@pytest.mark.parametrize("tag", ["a", "b"])
def test_integrals_match_reference_golden(tag):
    """Oracle restatement vs outputs of the reference's own compiled topo_param
    (fixture written by tests/golden/make_golden.py).  The reference is built
    with -ffast-math; its own fp32 self-noise is 1.2e-6."""
    g = np.load(os.path.join(GOLD, "integrals_ref.npz"))
    seed, ny, nx, K = (int(v) for v in g[tag + "_shape"])
    azim, hori, tilt = _integral_inputs(seed, ny, nx, K)
    assert np.abs(oracle.sky_view_factor(azim, hori, tilt) - g[tag + "_svf"]).max() <= 3e-6
    assert np.abs(oracle.visible_sky_fraction(azim, hori, tilt) - g[tag + "_vsf"]).max() <= 3e-6
    assert np.abs(oracle.topographic_openness(azim, hori) - g[tag + "_top"]).max() <= 3e-6


def test_integrals_known_answers():
    """Analytic cases verified against the compiled reference in SURVEY.md 8c."""
    K = 360
    azim = np.array([2 * np.pi / K * i for i in range(K)], np.float32)
    flat = np.zeros((1, 1, 3), np.float32); flat[..., 2] = 1.0
    h0 = np.zeros((1, 1, K), np.float32)
    assert abs(oracle.sky_view_factor(azim, h0, flat)[0, 0] - 1.0) < 2e-6
    assert abs(oracle.visible_sky_fraction(azim, h0, flat)[0, 0] - 1.0) < 2e-6
    assert oracle.topographic_openness(azim, h0)[0, 0] == np.float32(1.5707995)  # the compiled reference's value (float accumulator)
    h30 = np.full((1, 1, K), np.deg2rad(30.0), np.float32)
    assert abs(oracle.sky_view_factor(azim, h30, flat)[0, 0] - 0.75) < 3e-6       # cos^2 h
    assert abs(oracle.visible_sky_fraction(azim, h30, flat)[0, 0] - 0.5) < 3e-6   # 1 - sin h
    a = np.deg2rad(40.0)
    tilt = np.array([[[np.sin(a), 0.0, np.cos(a)]]], np.float32)
    assert abs(oracle.sky_view_factor(azim, h0, tilt)[0, 0] - (1 + np.cos(a)) / 2) < 5e-6


def test_live_reference_if_built():
    """When oracle/_ref holds the compiled reference (build container), compare live."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_ref", os.path.join(os.path.dirname(GOLD), "..", "oracle", "build_ref.py"))
    br = importlib.util.module_from_spec(spec); spec.loader.exec_module(br)
    try:
        tp, _, _ = br.load()
    except ImportError:
        pytest.skip("oracle/_ref not built")
    azim, hori, tilt = _integral_inputs(42, 9, 11, 120)
    assert np.abs(oracle.sky_view_factor(azim, hori, tilt) - np.asarray(tp.sky_view_factor(azim, hori, tilt))).max() <= 3e-6
    assert np.abs(oracle.visible_sky_fraction(azim, hori, tilt) - np.asarray(tp.visible_sky_fraction(azim, hori, tilt))).max() <= 3e-6
    assert np.abs(oracle.topographic_openness(azim, hori) - np.asarray(tp.topographic_openness(azim, hori))).max() <= 3e-6


def test_tables_match_specification():
    """Defaults: 2101 entries, step 0.05 deg, anchored at 89.98 deg, first entry
    -15.02 deg."""
    t = oracle.tables(360, 50.0, 0.25, -15.0)
    ea = t["elev_ang"]
    assert len(ea) == 2101
    assert abs(np.rad2deg(ea[-1]) - 89.98) < 1e-4 and abs(np.rad2deg(ea[0]) + 15.02) < 1e-3
    assert np.allclose(np.diff(ea.astype(np.float64)), np.deg2rad(0.05), atol=2e-7)
    assert np.allclose(t["elev_sin"], np.sin(ea.astype(np.float64)), atol=1e-7)
    assert t["azim_sin"][0] == 0.0 and t["azim_cos"][0] == 1.0


def _flat(n=48, spacing=50.0, z=0.0):
    xs = (np.arange(n) * spacing).astype(np.float32)
    x, y = np.meshgrid(xs, xs)
    return x, y, np.full((n, n), z, np.float32)


@pytest.mark.parametrize("alg", ["guess_constant", "binary_search", "discrete_sampling"])
def test_kat_flat_plane(alg):
    """Horizontal plane: a ray at e >= 0 from 1 cm up never hits, a ray 0.05 deg
    below horizontal lands 11 m away => horizon = 0 within +-hori_acc."""
    x, y, z = _flat()
    vg = syn.rearrange_pad_buffer(x, y, z)
    nrm, nth = syn.planar_frames(8, 8)
    h, az = oracle.horizon_gridded(vg, 48, 48, nrm, nth, 20, 20, 2.0, azim_num=24, ray_algorithm=alg)
    assert h.shape == (8, 8, 24) and az.shape == (24,)
    assert np.abs(h).max() <= np.deg2rad(0.25) + 1e-6


@pytest.mark.parametrize("alg", ["guess_constant", "binary_search"])
def test_kat_plateau_edge(alg):
    """Plateau of height h north of row y0: seen from a low cell at distance d
    the edge has elevation atan(h cos(a) / d) for azimuth a (clockwise from
    north, y = north: horizon_comp.cpp:447-449) within +-hori_acc."""
    n, sp, hgt, row0 = 96, 25.0, 120.0, 70
    x, y, z = _flat(n, sp)
    z[row0:, :] = hgt
    vg = syn.rearrange_pad_buffer(x, y, z)
    nrm, nth = syn.planar_frames(1, 1)
    i0, j0 = 40, 48
    K = 72
    h, az = oracle.horizon_gridded(vg, n, n, nrm, nth, i0, j0, 5.0, azim_num=K, ray_algorithm=alg)
    d = (row0 - i0) * sp
    for k in range(K):
        a = az[k]
        if np.cos(a) > 0.5:  # looking towards the plateau, edge inside the DEM
            want = np.arctan(hgt * np.cos(a) / d)
            assert abs(h[0, 0, k] - want) <= np.deg2rad(0.25) + 2e-4, (k, h[0, 0, k], want)
        if np.cos(a) < -0.2:  # looking away: flat
            assert abs(h[0, 0, k]) <= np.deg2rad(0.25) + 1e-6


def test_algorithms_agree_and_bvh_equals_brute_force():
    c = syn.make_config("cfg1", n=48)
    c_args = (c["vert_grid"], 48, 48, c["vec_norm"], c["vec_north"], 16, 16, 5.0)
    res = {}
    for alg in ("guess_constant", "binary_search", "discrete_sampling"):
        h, _, rays = oracle.horizon_gridded(*c_args, azim_num=16, ray_algorithm=alg, return_rays=True)
        hb, _, raysb = oracle.horizon_gridded(*c_args, azim_num=16, ray_algorithm=alg, brute_force=True, return_rays=True)
        assert np.array_equal(h, hb) and rays == raysb, "BVH culling changed a decision"
        res[alg] = h
    acc = np.deg2rad(0.25)
    assert np.abs(res["guess_constant"] - res["binary_search"]).max() <= 2 * acc + 1e-6
    assert np.abs(res["guess_constant"] - res["discrete_sampling"]).max() <= 2 * acc + 1e-6


def test_oracle_regression_fixture():
    """Oracle self-regression (NOT a reference vector: ray path is parity-unpinned)."""
    g = np.load(os.path.join(GOLD, "horizon_oracle_regression.npz"))
    c = syn.make_config("cfg1", n=64)
    args = (c["vert_grid"], 64, 64, c["vec_norm"], c["vec_north"], 16, 16, 5.0)
    for alg in ("guess_constant", "binary_search", "discrete_sampling"):
        h, _, rays = oracle.horizon_gridded(*args, azim_num=24, ray_algorithm=alg, return_rays=True)
        assert np.array_equal(h, g[alg]) and rays == int(g[alg + "_rays"][0])


def test_masked_cells_and_termination_rule():
    """Masked cells get hori_fill; a cell on a pedestal looking over empty space
    terminates (the reference would loop forever) at table index 0.
    Cell (0, 0) of the inner domain sits on the apex."""
    x, y, z = _flat(24, 100.0)
    z[11, 11] = 5000.0  # 5 km spike: from its apex nothing is visible down to -15 deg within 1 km
    vg = syn.rearrange_pad_buffer(x, y, z)
    nrm, nth = syn.planar_frames(2, 2)
    mask = np.array([[1, 0], [0, 1]], np.uint8)
    h, _ = oracle.horizon_gridded(vg, 24, 24, nrm, nth, 11, 11, 1.0, azim_num=8, mask=mask, hori_fill=7.0)
    assert np.all(h[0, 1] == 7.0) and np.all(h[1, 0] == 7.0)
    assert np.all(np.isfinite(h)) and h[0, 0].min() < np.deg2rad(-14.0)


def _hemisphere(n=200, spacing=125.0, radius=4500.0):
    xs = ((np.arange(n) - (n - 1) / 2.0) * spacing).astype(np.float32)
    x, y = np.meshgrid(xs, xs)
    r2 = radius ** 2 - x.astype(np.float64) ** 2 - y.astype(np.float64) ** 2
    z = np.sqrt(np.clip(r2, 0.0, None)).astype(np.float32)
    return x, y, z


def test_hemisphere_shadow_partition_and_energy_conservation():
    """Synthetic hemispherical mountain of examples/shadow/gridded_planar_DEM_
    artificial.py:45-63: codes partition the domain, the lit side faces the sun,
    and the domain-mean sw_dir_cor stays close to 1 (:190-204 plots 0.85-1.05)."""
    n, rim = 200, 20  # inner domain +-10 km holds the whole shadow (tip at R / sin(30 deg) = 9 km)
    x, y, z = _hemisphere(n)
    vg = syn.rearrange_pad_buffer(x, y, z)
    tilt = syn.tilt_vectors(x, y, z, rim)
    ny = nx = n - 2 * rim
    norm, _ = syn.planar_frames(ny, nx)
    enl = (1.0 / (norm * tilt).sum(axis=2)).astype(np.float32)
    elev = np.ascontiguousarray(z[rim:-rim, rim:-rim])
    mask = np.ones((ny, nx), np.uint8)
    t = oracle.Terrain()
    t.initialise(vg, n, n, rim, rim, tilt, norm, enl, elev, mask, ang_max=89.99)
    el = np.deg2rad(30.0)
    for azd in (0.0, 90.0, 200.0):
        a = np.deg2rad(azd)
        sun = np.array([1e7 * np.cos(el) * np.sin(a), 1e7 * np.cos(el) * np.cos(a), 1e7 * np.sin(el)], np.float32)
        code = np.empty((ny, nx), np.uint8); sw = np.empty((ny, nx), np.float32)
        t.shadow(sun, code); t.sw_dir_cor(sun, sw)
        assert set(np.unique(code)) <= {0, 1, 2}
        assert (code == 1).sum() > 0 and (code == 2).sum() > 0 and (code == 0).sum() > 0
        assert np.all(sw[code != 0] == 0.0) and np.all(sw[code == 0] > 0.0)
        assert abs(float(sw.mean()) - 1.0) < 0.05
        # terrain shadow lies on the far side of the mountain from the sun
        yy, xx = np.nonzero(code == 2)
        cx = x[rim:-rim, rim:-rim][yy, xx].mean(); cy = y[rim:-rim, rim:-rim][yy, xx].mean()
        assert cx * np.sin(a) + cy * np.cos(a) < 0.0


def _slope_inputs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    return mg.slope_inputs()


def test_slope_matches_reference_golden():
    """Oracle restatement of slope_plane_meth / slope_vector_meth vs the unmodified
    compiled reference (tests/golden/slope_ref.npz).  The reference solves with LAPACK
    sgesv and is built -ffast-math: unit-vector components agree to 2e-5."""
    g = np.load(os.path.join(GOLD, "slope_ref.npz"))
    x, y, z, rot = _slope_inputs()
    cases = {"plane_id": oracle.slope_plane_meth(x, y, z),
             "plane_rot": oracle.slope_plane_meth(x, y, z, rot_mat=rot, output_rot=False),
             "plane_rot_out": oracle.slope_plane_meth(x, y, z, rot_mat=rot, output_rot=True),
             "vector_id": oracle.slope_vector_meth(x, y, z),
             "vector_rot_out": oracle.slope_vector_meth(x, y, z, rot_mat=rot, output_rot=True)}
    for name, got in cases.items():
        want = g[name]
        assert got.shape == want.shape
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        assert np.nanmax(np.abs(got - want)) <= 2e-5, (name, float(np.nanmax(np.abs(got - want))))
    # known answer: z = 0.5 x  ->  (-0.4472136, 0, 0.8944272)
    xs = np.arange(5, dtype=np.float32); X, Y = np.meshgrid(xs, xs)
    v = oracle.slope_plane_meth(X, Y, (0.5 * X).astype(np.float32))[2, 2]
    assert np.allclose(v, [-0.4472136, 0.0, 0.8944272], atol=1e-6)


def _ulp32(ref):
    return np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)


def test_transform_matches_reference_golden():
    """Coordinate preparation (scope row 8f-3): the oracle against the outputs of the UNMODIFIED
    compiled reference transform.pyx / direction.pyx (tests/golden/transform_ref.npz).  The
    reference is built with -ffast-math, the oracle without: float64 outputs agree to 2e-9 m
    (2 ulp at Earth-radius magnitude), float32 outputs to 1 ulp."""
    import oracle
    g = np.load(os.path.join(GOLD, "transform_ref.npz"))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    lon, lat, h = mg.transform_inputs()

    class T:  # TransformerEcef2enu attributes as stored by the reference
        pass
    for el in ("sphere", "GRS80", "WGS84"):
        x, y, z = oracle.lonlat2ecef(lon, lat, h, el)
        for a, k in ((x, "x"), (y, "y"), (z, "z")):
            assert np.abs(a - g[el + "_" + k]).max() <= 2e-9, (el, k)
        t = T()
        t.x_ecef_or, t.y_ecef_or, t.z_ecef_or, t.lon_or, t.lat_or = g[el + "_orig"]
        enu = oracle.ecef2enu(g[el + "_x"], g[el + "_y"], g[el + "_z"], t)
        for a, k in zip(enu, ("xe", "ye", "ze")):
            assert (np.abs(a.astype(np.float64) - g[el + "_" + k]) <= _ulp32(g[el + "_" + k])).all(), (el, k)
        nrm = oracle.surf_norm(lon, lat)
        assert (np.abs(nrm.astype(np.float64) - g[el + "_nrm"]) <= _ulp32(g[el + "_nrm"])).all()
        nth = oracle.north_dir(g[el + "_x"], g[el + "_y"], g[el + "_z"], g[el + "_nrm"], el)
        assert np.abs(nth.astype(np.float64) - g[el + "_nth"]).max() <= 1.2e-7, el
        for src, k in ((g[el + "_nrm"], "nrm_e"), (g[el + "_nth"], "nth_e")):
            v = oracle.ecef2enu_vector(src, t)
            assert np.abs(v.astype(np.float64) - g[el + "_" + k]).max() <= 1.2e-7, (el, k)
        rot = oracle.rotation_matrix_glob2loc(g[el + "_nth_e"], g[el + "_nrm_e"])
        assert np.array_equal(np.isnan(rot), np.isnan(g[el + "_rot"]))
        assert np.nanmax(np.abs(rot - g[el + "_rot"])) <= 6e-8
    e, n, hc = oracle.wgs2swiss(lon, lat, h)
    assert np.abs(e - g["swiss_e"]).max() <= 2e-9 and np.abs(n - g["swiss_n"]).max() <= 2e-9
    assert (np.abs(hc.astype(np.float64) - g["swiss_h"]) <= _ulp32(g["swiss_h"])).all()
    lo2, la2, hw = oracle.swiss2wgs(g["swiss_e"], g["swiss_n"], g["swiss_h"])
    assert np.abs(lo2 - g["back_lon"]).max() <= 5e-14 and np.abs(la2 - g["back_lat"]).max() <= 5e-14   # degrees; -ffast-math reassociation
    assert (np.abs(hw.astype(np.float64) - g["back_h"]) <= _ulp32(g["back_h"])).all()
    # known answer: the ENU origin maps to (0, 0, 0), its normal to (0, 0, 1), north to (0, 1, 0)
    t = T()
    t.x_ecef_or, t.y_ecef_or, t.z_ecef_or, t.lon_or, t.lat_or = g["WGS84_orig"]
    lo, la = np.array([[t.lon_or]]), np.array([[t.lat_or]])
    x0, y0, z0 = oracle.lonlat2ecef(lo, la, np.zeros((1, 1), np.float32), "WGS84")
    assert max(abs(float(v[0, 0])) for v in oracle.ecef2enu(x0, y0, z0, t)) <= 1e-6
    n0 = oracle.surf_norm(lo, la)
    assert np.allclose(oracle.ecef2enu_vector(n0, t)[0, 0], [0, 0, 1], atol=1e-7)
    assert np.allclose(oracle.ecef2enu_vector(oracle.north_dir(x0, y0, z0, n0, "WGS84"), t)[0, 0], [0, 1, 0], atol=1e-7)


def test_shared_diagonal_quad_test_equals_two_triangle_tests():
    """The production quad test (csrc/hzb_tri.cuh -- the source the kernels compile, built here for the host)
    evaluates five Pluecker edge functions per quad for two rays: the diagonal's function of the second
    triangle is taken as the NEGATIVE of the first triangle's (exact in round-to-nearest).  On two million
    random and adversarial cases (rays through vertices, along the diagonal, on outer edges, short tfar, a
    companion ray tilted by up to half a degree) it must make the oracle's decisions tri_hit(T1) || tri_hit(T2)
    for both rays, and the product's tri_hit must equal the oracle's in decision and distance."""
    import ctypes
    L = oracle.lib()
    L.orc_selftest_shared_diagonal.restype = ctypes.c_longlong
    hits = ctypes.c_longlong(0)
    bad = L.orc_selftest_shared_diagonal(ctypes.c_ulonglong(12345), ctypes.c_longlong(2_000_000), ctypes.byref(hits))
    assert bad == 0
    assert 200_000 < hits.value < 1_900_000      # the cases are not trivially all-hit or all-miss


def test_folded_bias_box_test_is_conservative():
    """The packet step's box test (csrc/hzb_box.cuh -- the source the kernels compile, built here for the host:
    PRMT decode of the 16-bit planes, 2^23 bias folded into the ray constant, clamped reciprocal, shared x/y
    selectors, no slack on tmax) must never reject a box that a ray of the packet meets in exact arithmetic --
    also for flat boxes, axis-parallel rays, anisotropic grids and coordinates of 3e6 m magnitude.  It relies on
    the extra quantum bvh_wide.cu adds to every plane: without it the same test does reject such boxes."""
    import ctypes
    L = oracle.lib()
    L.orc_selftest_folded_slab.restype = ctypes.c_longlong
    acc = ctypes.c_longlong(0)
    bad = L.orc_selftest_folded_slab(ctypes.c_ulonglong(777), ctypes.c_longlong(3_000_000), 1, ctypes.byref(acc))
    assert bad == 0
    assert acc.value > 300_000
    bad0 = L.orc_selftest_folded_slab(ctypes.c_ulonglong(777), ctypes.c_longlong(3_000_000), 0, ctypes.byref(acc))
    assert bad0 > 0          # the test is sharp: the quantum is what makes the folded form conservative


@pytest.mark.parametrize("alg", ["guess_constant", "binary_search", "discrete_sampling"])
def test_product_state_machine_on_the_cpu(alg):
    """The CUDA kernels' search state machine (csrc/hzb_search.cuh -- the very source the device code is built
    from, compiled here for the host) asks for casts and packet companions; the oracle answers them.  Outputs
    must equal the oracle's own algorithm bit for bit and the cast count must be the reference's (per-cell
    tilted frames; origins and ray directions come from the product's make_frame / ray_dir and must carry the
    oracle's bits): a companion
    result counts only when the search would have cast it.  Also with every third companion refused (what the
    kernel does when two rays cannot share plane selectors), and on cliffs with a high lower limit where the
    stepping searches run into both ends of the elevation table."""
    c = syn.make_config("cfg1", n=56)
    base = (c["vert_grid"], 56, 56, c["offset_0"], c["offset_1"], c["ny"], c["nx"], 24, c["dist_search"])
    for refuse in (0, 3):
        bad, ref, got, used = oracle.selftest_state_machine(*base, ray_algorithm=alg, refuse_every=refuse)
        assert bad == 0 and ref == got and ref > 0
        if alg != "binary_search":
            assert used > 0      # companions were really consumed
    x, y, z = syn.sinusoid_dem(40, 40, 5.0, 300.0, 140.0, 5, 3)        # cliffs: slopes far beyond 60 degrees
    vg = syn.rearrange_pad_buffer(x, y, z)
    for low, acc in ((-2.0, 1.0), (-40.0, 0.1)):
        bad, ref, got, used = oracle.selftest_state_machine(vg, 40, 40, 2, 2, 36, 36, 15, 0.3, hori_acc=acc,
                                                            elev_ang_low_lim=low, ray_algorithm=alg, refuse_every=2)
        assert bad == 0 and ref == got



@pytest.mark.parametrize("alg", ["guess_constant", "binary_search", "discrete_sampling"])
def test_azimuth_segments_on_the_cpu(alg):
    """Azimuth segments (the tail of a launch, csrc/horizon.cu): segment 1.. of every cell runs as a task of its own
    through the product's state machine -- prelude, then its azimuths -- on the oracle's casts, half of the tasks with
    the head's bisection result handed over.  A task whose prelude found the chain's index must reproduce the chain bit
    for bit and write nothing outside its segment; where it did not find it (the chain was clamped at the table's low
    end on its way: cells that look out over the DEM's edge) the product's fix-up pass recomputes, which the GPU suite
    covers.  Away from the edge the assumption must hold for (nearly) every task."""
    c = syn.make_config("cfg1", n=72)
    bad, ref, tasks, missed, pre = oracle.selftest_segments(c["vert_grid"], 72, 72, c["offset_0"], c["offset_1"], c["ny"], c["nx"], 40,
                                                            c["dist_search"], ray_algorithm=alg, segments=4)
    assert bad == 0 and tasks == 3 * c["ny"] * c["nx"]
    if alg == "guess_constant":
        assert 0 < pre < 20 * tasks and missed <= 0.01 * tasks
    else:
        assert pre == 0 and missed == 0          # independent azimuths: no prelude, nothing to assume
    # the rim of a steep DEM, low limit close to the horizontal: chains are clamped all the time, most guesses miss,
    # verified segments must still be exact
    x, y, z = syn.sinusoid_dem(40, 40, 5.0, 300.0, 140.0, 5, 3)
    vg = syn.rearrange_pad_buffer(x, y, z)
    bad, ref, tasks, missed, pre = oracle.selftest_segments(vg, 40, 40, 1, 1, 38, 38, 32, 0.3, hori_acc=0.5, elev_ang_low_lim=-2.0,
                                                            ray_algorithm=alg, segments=4)
    assert bad == 0 and tasks > 0
    if alg == "guess_constant":
        assert missed > 0


def test_fast_traversal_equals_plain():
    """The oracle's fast any-hit traversal (implicit 4-ary hierarchy over the grid quads, four boxes per SSE step; used by
    the CPU TIMING legs of bench.py) against its plain binary-BVH walker (used by every parity check): identical horizon
    arrays and cast counts for all three algorithms and odd grid sizes (partial groups), and identical shadow codes
    (casts with tfar = inf)."""
    try:
        for name, n, K, alg in (("cfg1", None, 36, "guess_constant"), ("cfg1", 77, 20, "binary_search"),
                                ("cfg1", 50, 12, "discrete_sampling"), ("cfg2", 131, 48, "guess_constant"),
                                ("cfg4p", 97, 40, "guess_constant")):
            c = syn.make_config(name, n)
            args = (c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["vec_norm"], c["vec_north"], c["offset_0"], c["offset_1"],
                    c["dist_search"])
            oracle.set_fast_traversal(False)
            h0, _, r0 = oracle.horizon_gridded(*args, azim_num=K, ray_algorithm=alg, return_rays=True)
            oracle.set_fast_traversal(True)
            h1, _, r1 = oracle.horizon_gridded(*args, azim_num=K, ray_algorithm=alg, return_rays=True)
            assert np.array_equal(h0, h1) and r0 == r1, (name, n, alg)
        # shadow codes: one any-hit ray per cell towards the sun
        n, rim = 101, 10
        x, y, z = syn.sinusoid_dem(n, n, 50.0, 400.0, 4000.0, 5, 4)
        tilt = syn.tilt_vectors(x, y, z, rim)
        norm, _ = syn.planar_frames(n - 2 * rim, n - 2 * rim)
        enl = (1.0 / (norm * tilt).sum(axis=2)).astype(np.float32)
        elev = np.ascontiguousarray(z[rim:-rim, rim:-rim]); mask = np.ones(elev.shape, np.uint8)
        vg = syn.rearrange_pad_buffer(x, y, z)
        codes = []
        for fast in (False, True):
            oracle.set_fast_traversal(fast)
            t = oracle.Terrain()
            t.initialise(vg, n, n, rim, rim, tilt, norm, enl, elev, mask)
            out = []
            for sun in syn.sun_positions_diurnal(6):
                b = np.empty(mask.shape, np.uint8); t.shadow(sun, b); out.append(b)
            codes.append(np.stack(out))
        assert np.array_equal(codes[0], codes[1]) and len(np.unique(codes[0])) > 1
    finally:
        oracle.set_fast_traversal(False)


def test_queue_order_covers_every_task_once():
    """The work queue of a horizon launch (csrc/hzb_queue.cuh, the kernels' own source built for the host): band tiles
    first, interior in row order, the last tiles once per azimuth segment -- enumerated for 4000 random geometries
    (tile grids, block sharding, bands, tail sizes incl. none / everything).  Every task exactly once, whole chains
    before segments, and the predicates the lanes, the fix-up kernel and the host tier's row counters use agree."""
    assert oracle.selftest_queue(seed=11, iters=4000) == 0


def test_parity_sensitivity_to_the_rounding_of_the_triangle_test():
    """The ray path is 'parity unpinned' (Embree absent).  What CAN be measured: how many outputs depend on the
    rounding of the triangle test at all.  The oracle with the specified fp32 Pluecker arithmetic against the same
    predicate in double on the exact inputs (no epsilon): a handful of outputs per million may move (by one
    search step of 10 table entries), everything else is decided far from any rounding."""
    import oracle
    from horayzon_b200 import synthetic
    total = diff = 0
    for name, n, nrows in (("cfg1", None, 96), ("cfg2", 401, 12)):
        c = synthetic.make_config(name, n)
        sc = oracle.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
        rows = sorted(set(np.round(np.linspace(0, c["ny"] - 1, nrows)).astype(int).tolist()))
        a = (rows, c["vec_norm"], c["vec_north"], c["offset_0"], c["offset_1"], c["dist_search"])
        h32 = sc.horizon_rows(*a, azim_num=c["azim_num"])
        oracle.set_exact_predicate(True)
        try:
            h64 = sc.horizon_rows(*a, azim_num=c["azim_num"])
        finally:
            oracle.set_exact_predicate(False)
        sc.close()
        total += h32.size; diff += int((h32 != h64).sum())
        assert np.abs(h32 - h64).max() <= 2.1 * np.deg2rad(0.25) + 1e-6      # a flipped decision moves the bracket by one step
    print("parity sensitivity: %d of %d outputs depend on the rounding of the triangle test" % (diff, total))
    assert diff <= 1e-5 * total + 2
