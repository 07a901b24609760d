"""GPU parity: the CUDA product (through its reference-shaped Python/C-ABI
surface) against the CPU oracle on identical seeded inputs.

Bar: horizon outputs are table entries, so parity is
DECISION-exact: arrays must be bit-identical (tolerance 1e-4 rad from the
north-star is asserted as well and is implied).  Shadow codes bit-exact;
sw_dir_cor bit-exact without refraction; SVF/VSF/openness |d| <= 5e-6.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built):
    import horayzon_b200 as hb
    import oracle
    return hb, oracle


@pytest.fixture
def dbg(mods):
    """hzb_debug_option with automatic reset: selects the second implementations / tuning knobs explicitly
    (no environment variable changes the product's kernels)."""
    hb, _ = mods
    yield hb.resident.debug_option
    hb.resident.debug_option("reset", 0)


def _cfg(hb, name, n=None):
    c = hb.synthetic.make_config(name, n)
    args = (c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["vec_norm"], c["vec_north"],
            c["offset_0"], c["offset_1"], c["dist_search"])
    return c, args


def _assert_same(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype
    diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
    nbad = int((diff > 1e-4).sum())  # north-star tolerance [rad]
    assert nbad == 0, f"{what}: {nbad}/{a.size} elements differ by more than 1e-4 rad (max {diff.max():.3e})"
    assert np.array_equal(a, b), f"{what}: not bit-identical (max |d| {diff.max():.3e})"


@pytest.mark.parametrize("alg", ["guess_constant", "binary_search", "discrete_sampling"])
def test_horizon_gridded_cfg1(mods, alg):
    hb, oracle = mods
    c, args = _cfg(hb, "cfg1")
    h_gpu, az_gpu = hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"], ray_algorithm=alg)
    h_cpu, az_cpu, rays = oracle.horizon_gridded(*args, azim_num=c["azim_num"], ray_algorithm=alg, return_rays=True)
    _assert_same(h_gpu, h_cpu, "cfg1 " + alg)
    assert np.array_equal(az_gpu, az_cpu)
    st = hb.resident.last_stats()
    assert st["rays"] == rays, "ray-cast counter differs from the oracle's"
    assert st["units"] == c["ny"] * c["nx"] * c["azim_num"]


@pytest.mark.parametrize("opts", [{"horizon_kernel": 1}, {"no_overlap": 1}, {"wrefill": 32, "wwait": 1},
                                  {"wrefill": 8, "wwait": 32}, {"stack_limit": 2}, {"stack_limit": 5}])
@pytest.mark.parametrize("alg", ["guess_constant", "binary_search"])
def test_horizon_kernel_variants_agree(mods, dbg, opts, alg):
    """The reference-shaped per-lane kernel (binary BVH), the non-overlapped host path, other
    scheduling thresholds and a traversal stack so small that most packets take the full-stack
    fallback must all give the oracle's bits and cast counts: decisions do not depend on
    traversal order or BVH layout."""
    hb, oracle = mods
    for k, v in opts.items():
        dbg(k, v)
    c, args = _cfg(hb, "cfg1")
    h_gpu, _ = hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"], ray_algorithm=alg)
    st = hb.resident.last_stats()
    h_cpu, _, rays = oracle.horizon_gridded(*args, azim_num=c["azim_num"], ray_algorithm=alg, return_rays=True)
    _assert_same(h_gpu, h_cpu, "variant %s" % opts)
    if "stack_limit" in opts:     # cells whose stack was full were recomputed by the fix-up kernel
        assert st["fallback_packets"] > 0, "the lowered stack limit did not exercise the fallback"
    else:
        assert st["fallback_packets"] == 0 and st["rays"] == rays


@pytest.mark.parametrize("alg", ["guess_constant", "binary_search", "discrete_sampling"])
@pytest.mark.parametrize("extra", [{}, {"stack_limit": 3}])
def test_azimuth_segments_everywhere(mods, dbg, alg, extra):
    """Tail splitting forced onto EVERY cell, rim included (no band): segments 1..3 of every cell run as tasks of their
    own.  cfg1's rim cells look out over the DEM's edge, where the chain is clamped at the table's low end and the
    prelude's assumption fails, so the fix-up pass has real work: the outputs must still be the oracle's bits and --
    without the full-stack fallback -- the cast counter the reference's."""
    hb, oracle = mods
    dbg("tail_segments", 4); dbg("tail_band", 0); dbg("tail_tiles", 1 << 30)
    for k, v in extra.items():
        dbg(k, v)
    c, args = _cfg(hb, "cfg1")
    h_gpu, _ = hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"], ray_algorithm=alg)
    st = hb.resident.last_stats()
    h_cpu, _, rays = oracle.horizon_gridded(*args, azim_num=c["azim_num"], ray_algorithm=alg, return_rays=True)
    _assert_same(h_gpu, h_cpu, "segments " + alg)
    print("\nsegments %s %s: tasks %d, recomputed %d, fallback packets %d" % (alg, extra, st["segment_tasks"], st["segment_redos"], st["fallback_packets"]))
    assert st["segment_tasks"] > 0.5 * 3 * c["ny"] * c["nx"]          # (the outermost tile ring is band by construction)
    assert st["units"] == c["ny"] * c["nx"] * c["azim_num"]
    if alg != "guess_constant" and not extra:
        assert st["segment_redos"] == 0
    if not extra:
        assert st["fallback_packets"] == 0 and st["rays"] == rays, "cast counter differs from the oracle's"


def test_azimuth_segments_layouts_and_quantised(mods, dbg):
    """Split cells with a mask, the azimuth-first layout, the 16-bit output and a sharded, packed launch."""
    hb, oracle = mods
    import torch
    c, args = _cfg(hb, "cfg2", n=161)
    K = 72
    mask = np.ones((c["ny"], c["nx"]), np.uint8); mask[::7, ::5] = 0
    dbg("tail_segments", 1)
    ref, _ = hb.horizon.horizon_gridded(*args, azim_num=K, mask=mask, hori_fill=-1.0)
    st_ref = hb.resident.last_stats()
    assert st_ref["segment_tasks"] == 0
    dbg("tail_segments", 4); dbg("tail_band", 0); dbg("tail_tiles", 1 << 30)
    a, _ = hb.horizon.horizon_gridded(*args, azim_num=K, mask=mask, hori_fill=-1.0)
    st = hb.resident.last_stats()
    print("\nsplit cfg2@161: %d segment tasks, %d recomputed by the fix-up pass" % (st["segment_tasks"], st["segment_redos"]))
    assert st["segment_tasks"] > 0
    # this DEM's rim cells look out over its edge: their chains are clamped at the table's low end, the preludes'
    # assumption fails there and the fix-up pass recomputes those segments -- with the wrong tasks' casts taken back
    assert st["segment_redos"] > 0
    assert st["rays"] == st_ref["rays"] and st["units"] == st_ref["units"] and st["fallback_packets"] == 0
    assert np.array_equal(a, ref)
    b, _ = hb.horizon.horizon_gridded(*args, azim_num=K, mask=mask, hori_fill=-1.0, azim_first=True)
    assert np.array_equal(np.moveaxis(b, 0, 2), ref)
    idx, first, table = hb.horizon.horizon_gridded_quantised(*args, azim_num=K, mask=mask, hori_fill=-1.0)[:3]
    assert np.array_equal(hb.horizon.dequantise(idx, first, table), ref)
    dev = torch.device("cuda:0")
    sc = hb.resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
    vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev); mk = torch.from_numpy(mask).to(dev)
    for world in (2, 3):
        per = hb.sharding.padded_block_rows(c["ny"], world)
        gathered = torch.full((world * per, c["nx"], K), float("nan"), device=dev)
        for r in range(world):
            sc.horizon_gridded_sharded(vn, vno, mk, c["offset_0"], c["offset_1"], gathered[r * per:(r + 1) * per], r, world, K,
                                       packed=True, dist_search=c["dist_search"], hori_fill=-1.0)
        torch.cuda.synchronize()
        assert np.array_equal(hb.sharding.unpack_blocks(gathered, c["ny"], world).cpu().numpy(), ref), "sharded x%d" % world
    assert sc.stats()["segment_tasks"] > 0


def test_shadow_kernel_variants_agree(mods, dbg):
    hb, oracle = mods
    vg, n, rim, tilt, norm, enl, elev, mask = _terrain_inputs(hb)
    t = hb.shadow.Terrain()
    t.initialise(vg, n, n, rim, rim, tilt, norm, enl, elev, mask)
    sun = hb.synthetic.sun_positions_diurnal(8)[2]
    a = np.empty(mask.shape, np.uint8); b = np.empty(mask.shape, np.uint8)
    t.shadow(sun, a)
    for variant in ({"shadow_kernel": 1}, {"shadow_kernel": 2}, {"stack_limit": 2}, {"shadow_kernel": 2, "stack_limit": 3}):
        dbg("reset", 0)    # per-lane BVH2 kernel, nearest-first traversal, full-stack fallback
        for k, v in variant.items():
            dbg(k, v)
        b[:] = 77
        t.shadow(sun, b)
        assert np.array_equal(a, b), variant


def test_horizon_gridded_cfg2_shrunk(mods):
    hb, oracle = mods
    c, args = _cfg(hb, "cfg2", n=301)
    h_gpu, _ = hb.horizon.horizon_gridded(*args, azim_num=360)
    h_cpu, _ = oracle.horizon_gridded(*args, azim_num=360)
    _assert_same(h_gpu, h_cpu, "cfg2@301")


def test_horizon_gridded_cfg4_shrunk_fine_accuracy(mods):
    hb, oracle = mods
    c, args = _cfg(hb, "cfg4p", n=256)
    kw = dict(azim_num=120, hori_acc=0.15, elev_ang_low_lim=-25.0, ray_org_elev=0.05)
    h_gpu, _ = hb.horizon.horizon_gridded(*args, **kw)
    h_cpu, _ = oracle.horizon_gridded(*args, **kw)
    _assert_same(h_gpu, h_cpu, "cfg4p@256")


def test_horizon_gridded_mask_fill_and_odd_azimuths(mods):
    hb, oracle = mods
    c, args = _cfg(hb, "cfg1")
    rng = np.random.default_rng(7)
    mask = (rng.random((c["ny"], c["nx"])) < 0.6).astype(np.uint8)
    kw = dict(azim_num=37, mask=mask, hori_fill=-1.25)  # 37: exercises the scalar-store path
    h_gpu, _ = hb.horizon.horizon_gridded(*args, **kw)
    h_cpu, _ = oracle.horizon_gridded(*args, **kw)
    _assert_same(h_gpu, h_cpu, "masked")
    assert np.all(h_gpu[mask == 0] == np.float32(-1.25))


def test_horizon_gridded_curved_frames_and_tin(mods):
    """Non-trivial per-cell frames (tilted normals / rotated north) plus an outer
    TIN ring of coarse triangles (horizon_comp.cpp:199-218)."""
    hb, oracle = mods
    c, args = _cfg(hb, "cfg1")
    ny, nx = c["ny"], c["nx"]
    rng = np.random.default_rng(11)
    nrm = np.zeros((ny, nx, 3)); nrm[..., 2] = 1.0
    nrm[..., 0] = rng.normal(0, 0.02, (ny, nx)); nrm[..., 1] = rng.normal(0, 0.02, (ny, nx))
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    nth = np.zeros((ny, nx, 3)); nth[..., 1] = 1.0; nth[..., 0] = 0.05
    nth -= (nth * nrm).sum(axis=2, keepdims=True) * nrm
    nth /= np.linalg.norm(nth, axis=2, keepdims=True)
    nrm, nth = nrm.astype(np.float32), nth.astype(np.float32)
    # TIN: a ring of 8 big triangles around the DEM, 200..600 m high
    L = 127 * 100.0
    vs = np.array([[-3000, -3000, 200], [L / 2, -4000, 600], [L + 3000, -3000, 300], [L + 4000, L / 2, 500],
                   [L + 3000, L + 3000, 250], [L / 2, L + 4000, 550], [-3000, L + 3000, 350], [-4000, L / 2, 450],
                   [-6000, -6000, 0], [L + 6000, -6000, 0], [L + 6000, L + 6000, 0], [-6000, L + 6000, 0]], np.float32)
    ti = np.array([[0, 1, 8], [1, 2, 9], [2, 3, 9], [3, 4, 10], [4, 5, 10], [5, 6, 11], [6, 7, 11], [7, 0, 8]], np.int32)
    a = (c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], nrm, nth, c["offset_0"], c["offset_1"], 8.0)
    kw = dict(azim_num=48, vert_simp=vs.ravel(), num_vert_simp=len(vs), tri_ind_simp=ti.ravel(), num_tri_simp=len(ti))
    h_gpu, _ = hb.horizon.horizon_gridded(*a, **kw)
    h_cpu, _ = oracle.horizon_gridded(*a, **kw)
    _assert_same(h_gpu, h_cpu, "frames+TIN")


@pytest.mark.parametrize("alg,dist_out", [("binary_search", False), ("binary_search", True),
                                          ("discrete_sampling", True), ("guess_constant", False)])
def test_horizon_locations(mods, alg, dist_out):
    hb, oracle = mods
    c, _ = _cfg(hb, "cfg1")
    rng = np.random.default_rng(3)
    n = 23
    xy = rng.uniform(2000.0, 10000.0, (n, 2))
    coords = np.concatenate([xy, rng.uniform(-500.0, 900.0, (n, 1))], axis=1).astype(np.float32)
    coords[0] = (-5000.0, -5000.0, 0.0)  # off the DEM: stays NaN
    nrm = np.zeros((n, 3), np.float32); nrm[:, 2] = 1.0
    nth = np.zeros((n, 3), np.float32); nth[:, 1] = 1.0
    roe = rng.uniform(0.01, 2.0, n).astype(np.float32)
    kw = dict(azim_num=72, ray_algorithm=alg, ray_org_elev=roe, hori_dist_out=dist_out,
              elev_ang_low_lim=-89.98 if alg != "discrete_sampling" else -30.0)
    r_gpu = hb.horizon.horizon_locations(c["vert_grid"], 128, 128, coords, nrm, nth, 6.0, **kw)
    r_cpu = oracle.horizon_locations(c["vert_grid"], 128, 128, coords, nrm, nth, 6.0, **kw)
    assert len(r_gpu) == len(r_cpu)
    assert np.all(np.isnan(r_gpu[0][0]))
    for g, o in zip(r_gpu, r_cpu):
        assert np.array_equal(g, o, equal_nan=True)
    # the per-lane binary-BVH kernels (second implementation) and a forced full-stack fallback give the same bits
    for opt, val in (("horizon_kernel", 1), ("stack_limit", 2)):
        hb.resident.debug_option(opt, val)
        try:
            r_alt = hb.horizon.horizon_locations(c["vert_grid"], 128, 128, coords, nrm, nth, 6.0, **kw)
        finally:
            hb.resident.debug_option("reset", 0)
        for g, o in zip(r_alt, r_cpu):
            assert np.array_equal(g, o, equal_nan=True), (opt, alg, dist_out)


def test_horizon_locations_reference_use(mods):
    """The reference's own use of horizon_locations (examples/horizon/locations_curved_DEM.py): 1440 azimuths,
    hori_acc 0.1, distance to the horizon, per-location ray_org_elev -- here 150 locations on tilted frames."""
    hb, oracle = mods
    c, _ = _cfg(hb, "cfg2", n=301)
    rng = np.random.default_rng(9)
    n = 150
    ij = rng.integers(20, 280, (n, 2))
    coords = np.stack([c["x"][ij[:, 0], ij[:, 1]] + rng.uniform(-30, 30, n), c["y"][ij[:, 0], ij[:, 1]] + rng.uniform(-30, 30, n),
                       c["z"][ij[:, 0], ij[:, 1]] + rng.uniform(-200, 200, n)], axis=1).astype(np.float32)
    nrm = np.zeros((n, 3)); nrm[:, 2] = 1.0; nrm[:, :2] = rng.normal(0, 0.01, (n, 2)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nth = np.zeros((n, 3)); nth[:, 1] = 1.0; nth -= (nth * nrm).sum(axis=1, keepdims=True) * nrm; nth /= np.linalg.norm(nth, axis=1, keepdims=True)
    roe = rng.uniform(0.5, 3.0, n).astype(np.float32)
    kw = dict(azim_num=1440, hori_acc=0.1, ray_algorithm="binary_search", ray_org_elev=roe, hori_dist_out=True)
    a = (c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], coords, nrm.astype(np.float32), nth.astype(np.float32), 15.0)
    r_gpu = hb.horizon.horizon_locations(*a, **kw)
    st = hb.resident.last_stats()
    r_cpu = oracle.horizon_locations(*a, **kw)
    for g, o in zip(r_gpu, r_cpu):
        assert np.array_equal(g, o, equal_nan=True)
    assert st["fallback_packets"] == 0


def _terrain_inputs(hb, n=200, seed=5):
    x, y, z = hb.synthetic.sinusoid_dem(n, n, 50.0, 400.0, 4000.0, seed, 4)
    rim = 20
    tilt = hb.synthetic.tilt_vectors(x, y, z, rim)
    ny = nx = n - 2 * rim
    norm, _ = hb.synthetic.planar_frames(ny, nx)
    enl = (1.0 / (norm * tilt).sum(axis=2)).astype(np.float32)
    elev = np.ascontiguousarray(z[rim:-rim, rim:-rim])
    rng = np.random.default_rng(seed)
    mask = (rng.random((ny, nx)) < 0.9).astype(np.uint8)
    vg = hb.synthetic.rearrange_pad_buffer(x, y, z)
    return vg, n, rim, tilt, norm, enl, elev, mask


@pytest.mark.parametrize("refrac", [False, True])
def test_terrain_shadow_and_sw_dir_cor(mods, refrac):
    hb, oracle = mods
    vg, n, rim, tilt, norm, enl, elev, mask = _terrain_inputs(hb)
    tg, to = hb.shadow.Terrain(), oracle.Terrain()
    for t in (tg, to):
        t.initialise(vg, n, n, rim, rim, tilt, norm, enl, elev, mask, sw_dir_cor_fill=-9.0, ang_max=89.5,
                     refrac_cor=refrac)
    suns = hb.synthetic.sun_positions_diurnal(12)
    ny, nx = mask.shape
    n_code_diff = n_sw_diff = 0
    for s in suns:
        a = np.empty((ny, nx), np.uint8); b = np.empty((ny, nx), np.uint8)
        tg.shadow(s, a); to.shadow(s, b)
        fa = np.empty((ny, nx), np.float32); fb = np.empty((ny, nx), np.float32)
        tg.sw_dir_cor(s, fa); to.sw_dir_cor(s, fb)
        n_code_diff += int((a != b).sum())
        if refrac:
            n_sw_diff += int((~np.isclose(fa, fb, rtol=2e-6, atol=2e-6)).sum())
        else:
            n_sw_diff += int((fa != fb).sum())
        assert set(np.unique(a)) <= {0, 1, 2, 3}
        assert np.all(a[mask == 0] == 3) and np.all(fa[mask == 0] == np.float32(-9.0))
    if refrac:
        # libm float functions differ in the last ulp between glibc and CUDA: a
        # grazing ray may flip.  Bound the fraction instead of demanding zero.
        print("refraction: %d shadow codes and %d sw_dir_cor values of %d differ from the oracle"
              % (n_code_diff, n_sw_diff, len(suns) * ny * nx))
        assert n_code_diff <= 1e-4 * len(suns) * ny * nx
        assert n_sw_diff <= 1e-4 * len(suns) * ny * nx
    else:
        assert n_code_diff == 0 and n_sw_diff == 0
    batch = tg.shadow_batch(suns)
    one = np.empty((ny, nx), np.uint8); tg.shadow(suns[3], one)
    assert np.array_equal(batch[3], one)


def test_integrals_against_oracle(mods):
    hb, oracle = mods
    rng = np.random.default_rng(0)
    ny, nx, K = 37, 53, 360
    azim = np.array([2 * np.pi / K * i for i in range(K)], np.float32)
    hori = rng.uniform(0.0, 0.6, (ny, nx, K)).astype(np.float32)
    sl = np.deg2rad(rng.uniform(0, 50, (ny, nx))); asp = rng.uniform(0, 2 * np.pi, (ny, nx))
    tilt = np.stack([np.sin(sl) * np.sin(asp), np.sin(sl) * np.cos(asp), np.cos(sl)], axis=2).astype(np.float32)
    for name, args in (("sky_view_factor", (azim, hori, tilt)), ("visible_sky_fraction", (azim, hori, tilt)),
                       ("topographic_openness", (azim, hori))):
        g = getattr(hb.topo_param, name)(*args)
        o = getattr(oracle, name)(*args)
        assert g.shape == o.shape and g.dtype == np.float32
        assert np.abs(g - o).max() <= 5e-6, name  # stated tolerance


def test_slope_against_oracle_and_reference_golden(mods):
    """Scope row 'next 1': slope on the device vs the oracle (<= 2e-6: same float
    operation order, LU with partial pivoting) and vs the compiled reference's golden
    vectors (<= 2e-5)."""
    import os
    hb, oracle = mods
    gold = os.path.join(os.path.dirname(__file__), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gold, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    x, y, z, rot = mg.slope_inputs()
    g = np.load(os.path.join(gold, "slope_ref.npz"))
    runs = {"plane_id": (hb.topo_param.slope_plane_meth, oracle.slope_plane_meth, dict()),
            "plane_rot": (hb.topo_param.slope_plane_meth, oracle.slope_plane_meth, dict(rot_mat=rot, output_rot=False)),
            "plane_rot_out": (hb.topo_param.slope_plane_meth, oracle.slope_plane_meth, dict(rot_mat=rot, output_rot=True)),
            "vector_id": (hb.topo_param.slope_vector_meth, oracle.slope_vector_meth, dict()),
            "vector_rot_out": (hb.topo_param.slope_vector_meth, oracle.slope_vector_meth, dict(rot_mat=rot, output_rot=True))}
    for name, (f_gpu, f_cpu, kw) in runs.items():
        a = f_gpu(x, y, z, **kw); b = f_cpu(x, y, z, **kw)
        assert a.shape == b.shape == g[name].shape and a.dtype == np.float32
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        assert np.nanmax(np.abs(a - b)) <= 2e-6, (name, float(np.nanmax(np.abs(a - b))))
        assert np.nanmax(np.abs(a - g[name])) <= 2e-5, name


def test_transform_direction_against_oracle_and_reference_golden(mods):
    """Scope row 8f-3: coordinate preparation on the device through the reference-shaped
    wrappers, against the CPU oracle (same operation order; sin/cos differ by <= 2 ulp) and the
    golden outputs of the compiled reference.  float64: 4e-9 m; float32: 1 ulp / 1.2e-7 for
    unit vectors.  The fused resident kernel must give the bits of the step-by-step chain."""
    import importlib.util
    import os
    import torch
    hb, oracle = mods
    gold = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(gold, "transform_ref.npz"))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gold, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    lon, lat, h = mg.transform_inputs()
    ulp32 = lambda r: np.spacing(np.abs(r).astype(np.float32)).astype(np.float64)
    for el in ("sphere", "GRS80", "WGS84"):
        t = hb.transform.TransformerEcef2enu(lon_or=float(lon.mean()), lat_or=float(lat.mean()), ellps=el)
        assert np.allclose([t.x_ecef_or, t.y_ecef_or, t.z_ecef_or, t.lon_or, t.lat_or], g[el + "_orig"], rtol=0, atol=2e-9)
        x, y, z = hb.transform.lonlat2ecef(lon, lat, h, ellps=el)
        for a, k in ((x, "x"), (y, "y"), (z, "z")):
            assert a.dtype == np.float64 and a.shape == lon.shape
            assert np.abs(a - g[el + "_" + k]).max() <= 4e-9, (el, k)
        ox, oy, oz = oracle.lonlat2ecef(lon, lat, h, el)
        assert max(np.abs(x - ox).max(), np.abs(y - oy).max(), np.abs(z - oz).max()) <= 4e-9
        enu = hb.transform.ecef2enu(x, y, z, t)
        for a, k in zip(enu, ("xe", "ye", "ze")):
            assert a.dtype == np.float32
            assert (np.abs(a.astype(np.float64) - g[el + "_" + k]) <= ulp32(g[el + "_" + k])).all(), (el, k)
        nrm = hb.direction.surf_norm(lon, lat)
        nth = hb.direction.north_dir(x, y, z, nrm, ellps=el)
        nrm_e, nth_e = hb.transform.ecef2enu_vector(nrm, t), hb.transform.ecef2enu_vector(nth, t)
        for a, k in ((nrm, "nrm"), (nth, "nth"), (nrm_e, "nrm_e"), (nth_e, "nth_e")):
            assert a.shape == lon.shape + (3,) and np.abs(a.astype(np.float64) - g[el + "_" + k]).max() <= 1.2e-7, (el, k)
        rot = hb.transform.rotation_matrix_glob2loc(nth_e, nrm_e)
        assert np.array_equal(rot, oracle.rotation_matrix_glob2loc(nth_e, nrm_e), equal_nan=True)
        assert np.nanmax(np.abs(rot - g[el + "_rot"])) <= 2.4e-7
        # fused resident kernel: lon / lat are 1-D here, the grid is their outer product
        lon1, lat1 = np.ascontiguousarray(lon[0]), np.ascontiguousarray(lat[:, 0])
        lon2, lat2 = np.meshgrid(lon1, lat1)
        xs, ys, zs = hb.transform.lonlat2ecef(lon2, lat2, h, ellps=el)
        chain = np.stack(hb.transform.ecef2enu(xs, ys, zs, t), axis=2)
        n_in = hb.direction.surf_norm(lon2[1:-1, 1:-1], lat2[1:-1, 1:-1])
        t_in = hb.direction.north_dir(xs[1:-1, 1:-1], ys[1:-1, 1:-1], zs[1:-1, 1:-1], n_in, ellps=el)
        dev = torch.device("cuda:0")
        ny, nx = h.shape
        vg = torch.empty((ny, nx, 3), dtype=torch.float32, device=dev)
        vn = torch.empty((ny - 2, nx - 2, 3), dtype=torch.float32, device=dev)
        vt = torch.empty_like(vn)
        hb.resident.prep_enu_dev(torch.from_numpy(lon1).to(dev), torch.from_numpy(lat1).to(dev), torch.from_numpy(h).to(dev),
                                 el, t, 1, 1, ny - 2, nx - 2, vg, vn, vt)
        torch.cuda.synchronize()
        assert np.array_equal(vg.cpu().numpy(), chain)
        assert np.array_equal(vn.cpu().numpy(), hb.transform.ecef2enu_vector(n_in, t))
        assert np.array_equal(vt.cpu().numpy(), hb.transform.ecef2enu_vector(t_in, t))
    e, n, hc = hb.transform.wgs2swiss(lon, lat, h)
    assert np.abs(e - g["swiss_e"]).max() <= 2e-9 and np.abs(n - g["swiss_n"]).max() <= 2e-9
    assert (np.abs(hc.astype(np.float64) - g["swiss_h"]) <= ulp32(g["swiss_h"])).all()
    lo2, la2, hw = hb.transform.swiss2wgs(e, n, hc)
    assert np.abs(lo2 - g["back_lon"]).max() <= 5e-14 and np.abs(la2 - g["back_lat"]).max() <= 5e-14
    assert (np.abs(hw.astype(np.float64) - g["back_h"]) <= ulp32(g["back_h"])).all()
    # the reference's argument checks
    with pytest.raises(ValueError, match="Unknown value for 'ellps'"):
        hb.transform.lonlat2ecef(lon, lat, h, ellps="bessel")
    with pytest.raises(ValueError, match="incorrect data type"):
        hb.transform.lonlat2ecef(lon.astype(np.float32), lat, h, ellps="WGS84")
    with pytest.raises(ValueError, match="TransformerEcef2enu"):
        hb.transform.ecef2enu(lon, lat, lon, object())


def test_full_size_cfg2_two_implementations_and_oracle_rows(mods, dbg):
    """BASELINE configs[1] at FULL size (1201 x 1201 x 360 = 5.2e8 units): the production kernel
    (two-ray packets on the compressed 4-wide BVH) and the reference-shaped per-lane kernel on the
    binary BVH -- different traversal, different culling structure, different cast grouping --
    must agree bit for bit on every unit and on the number of casts; twelve rows spread over the
    domain (incl. both rims) are checked against the CPU oracle."""
    import torch
    hb, oracle = mods
    c, _ = _cfg(hb, "cfg2")
    ny, nx, K = c["ny"], c["nx"], c["azim_num"]
    dev = torch.device("cuda:0")
    sc = hb.resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
    vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
    mask = torch.ones((ny, nx), dtype=torch.uint8, device=dev)
    out = []
    rays = []
    for kern in (0, 1):
        dbg("horizon_kernel", kern)
        h = torch.full((ny, nx, K), float("nan"), dtype=torch.float32, device=dev)
        before = sc.stats()["rays"]
        sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], h, 0, ny, dist_search=c["dist_search"])
        torch.cuda.synchronize()
        rays.append(sc.stats()["rays"] - before)
        out.append(h)
    dbg("reset", 0)
    assert not torch.isnan(out[0]).any()
    assert torch.equal(out[0], out[1]), "production and per-lane kernels differ at full size"
    assert rays[0] == rays[1], "cast counters differ"
    rows = sorted(set(np.linspace(0, ny - 1, 12).astype(int).tolist()))
    host = out[0].cpu().numpy()
    for r in rows:
        h_cpu, _ = oracle.horizon_gridded(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["vec_norm"][r:r + 1],
                                          c["vec_north"][r:r + 1], c["offset_0"] + r, c["offset_1"], c["dist_search"],
                                          azim_num=K)
        _assert_same(host[r:r + 1], h_cpu, "cfg2 full size, row %d" % r)
    sc.close()


def test_azimuth_first_layout_is_the_transposed_reference_layout(mods):
    """Scope row 8f-4 (additive): azim_first=True gives np.moveaxis(hori, 2, 0) bit for bit, through the
    host tier and through the resident tier with row sharding."""
    import torch
    hb, oracle = mods
    c, args = _cfg(hb, "cfg1")
    mask = np.ones((c["ny"], c["nx"]), np.uint8); mask[5:9, 20:31] = 0
    h_ref, az = hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"], mask=mask, hori_fill=-1.0)
    h_t, az_t = hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"], mask=mask, hori_fill=-1.0, azim_first=True)
    assert h_t.shape == (c["azim_num"], c["ny"], c["nx"]) and np.array_equal(az, az_t)
    assert np.array_equal(h_t, np.moveaxis(h_ref, 2, 0))
    dev = torch.device("cuda:0")
    sc = hb.resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
    out = torch.full((c["azim_num"], c["ny"], c["nx"]), float("nan"), dtype=torch.float32, device=dev)
    vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
    for b, e in ((0, 37), (37, c["ny"])):   # two row shards into one azimuth-first array
        sc.horizon_gridded(vn, vno, torch.from_numpy(mask).to(dev), c["offset_0"], c["offset_1"], out, b, e,
                           dist_search=c["dist_search"], hori_fill=-1.0, azim_first=True)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), h_t)
    sc.close()


def test_page_locked_output_arrays(mods, monkeypatch):
    """Large outputs are backed by pooled page-locked memory (DMA instead of staging): same bits as the
    pageable path, ordinary ndarray semantics, blocks return to the pool and are reused."""
    hb, oracle = mods
    c, args = _cfg(hb, "cfg2", n=301)          # 299 x 299 x 360 floats = 129 MB >= the 64 MB threshold
    monkeypatch.setenv("HZB_PINNED_OUTPUT", "1")   # default: only from the second large output of a process
    h1, az = hb.horizon.horizon_gridded(*args, azim_num=360)
    assert type(h1.base).__name__ == "_PinnedBlock" and h1.flags.c_contiguous and h1.flags.writeable
    monkeypatch.setenv("HZB_PINNED_OUTPUT", "0")
    h0, _ = hb.horizon.horizon_gridded(*args, azim_num=360)
    monkeypatch.setenv("HZB_PINNED_OUTPUT", "1")
    assert h0.base is None and np.array_equal(h0, h1)
    tilt = hb.synthetic.tilt_vectors(c["x"], c["y"], c["z"], c["offset_0"])
    assert np.array_equal(hb.topo_param.sky_view_factor(az, h1, tilt), hb.topo_param.sky_view_factor(az, h0, tilt))
    view = h1[10:20]                           # a view keeps the block alive after the array is gone
    keep = view.copy()
    addr = h1.ctypes.data
    del h1
    h2, _ = hb.horizon.horizon_gridded(*args, azim_num=360)      # the block is still referenced: a new one
    assert h2.ctypes.data != addr and np.array_equal(view, keep) and np.array_equal(h2, h0)
    del view, h2
    h3, _ = hb.horizon.horizon_gridded(*args, azim_num=360)      # both blocks are back in the pool: reused
    assert np.array_equal(h3, h0)
    h3[0, 0, 0] = 123.0                        # writable like any ndarray
    assert h3[0, 0, 0] == 123.0


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106])
def test_randomised_configurations(mods, seed):
    """Seeded random small cases across the parameter space the packet kernel has to honour: all three
    algorithms, odd / tiny azimuth counts, coarse and fine accuracy, elevation tables whose ends are
    reached (steep terrain with a high lower limit: the search runs into index 0 / the top index, where the
    companion indices clamp), random masks, non-square inner domains, lifted origins, tilted frames."""
    hb, oracle = mods
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 72))
    spacing = float(rng.choice([5.0, 30.0, 100.0]))
    amp = spacing * n * float(rng.choice([0.05, 0.3, 1.5]))          # gentle ... cliffs (slopes beyond 60 deg)
    x, y, z = hb.synthetic.sinusoid_dem(n, n + 3, spacing, amp, spacing * n * 0.7, seed, int(rng.integers(1, 5)))
    vg = hb.synthetic.rearrange_pad_buffer(x, y, z)
    off0, off1 = int(rng.integers(1, 6)), int(rng.integers(1, 6))
    ny, nx = n - off0 - int(rng.integers(1, 6)), n + 3 - off1 - int(rng.integers(1, 6))
    nrm = np.zeros((ny, nx, 3)); nrm[..., 2] = 1.0
    if seed % 2:
        nrm[..., 0] = rng.normal(0, 0.03, (ny, nx)); nrm[..., 1] = rng.normal(0, 0.03, (ny, nx))
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    nth = np.zeros((ny, nx, 3)); nth[..., 1] = 1.0
    nth -= (nth * nrm).sum(axis=2, keepdims=True) * nrm
    nth /= np.linalg.norm(nth, axis=2, keepdims=True)
    mask = (rng.random((ny, nx)) < 0.85).astype(np.uint8)
    a = (vg, n, n + 3, nrm.astype(np.float32), nth.astype(np.float32), off0, off1, float(spacing * n * 1.5 / 1000.0))
    kw = dict(azim_num=int(rng.choice([3, 8, 25, 60])), hori_acc=float(rng.choice([0.1, 0.25, 1.0, 4.0])),
              elev_ang_low_lim=float(rng.choice([-40.0, -15.0, -2.0])), mask=mask, hori_fill=float(rng.normal()),
              ray_org_elev=float(rng.choice([0.006, 0.05, 2.0])),
              ray_algorithm=str(rng.choice(["guess_constant", "guess_constant", "binary_search", "discrete_sampling"])))
    h_gpu, _ = hb.horizon.horizon_gridded(*a, **kw)
    st = hb.resident.last_stats()
    h_cpu, _, rays = oracle.horizon_gridded(*a, return_rays=True, **kw)
    _assert_same(h_gpu, h_cpu, "random case %d %s" % (seed, kw))
    assert st["rays"] == rays, "cast counter differs from the oracle's in case %d" % seed



# ---------------------------------------------------------------------------------------------
# round 2: fused SVF, concurrent launches, block sharding, the big BASELINE configs at full size
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n", [("cfg1", None), ("cfg2", 301)])
def test_fused_sky_view_factor_equals_the_two_call_sequence(mods, name, n):
    """horizon_gridded(..., svf_vec_tilt=) integrates the SVF on the device-resident horizon: same horizon
    bits and the same SVF bits as horizon_gridded + topo_param.sky_view_factor (which uploads the array
    again), and both within 5e-6 of the oracle's SVF."""
    hb, oracle = mods
    c, args = _cfg(hb, name, n)
    tilt = hb.synthetic.tilt_vectors(c["x"], c["y"], c["z"], c["offset_0"])
    h2, az2 = hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"])
    svf2 = hb.topo_param.sky_view_factor(az2, h2, tilt)
    h1, az1, svf1 = hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"], svf_vec_tilt=tilt)
    assert np.array_equal(h1, h2) and np.array_equal(az1, az2)
    assert np.array_equal(svf1, svf2)
    assert np.abs(svf1 - oracle.sky_view_factor(az1, h1, tilt)).max() <= 5e-6
    with pytest.raises(ValueError):
        hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"], svf_vec_tilt=tilt[:-1])
    with pytest.raises(ValueError):
        hb.horizon.horizon_gridded(*args, azim_num=c["azim_num"], svf_vec_tilt=tilt, azim_first=True)


def test_concurrent_launches_on_one_scene(mods):
    """Two launches on different streams against the same scene -- different row ranges AND different
    table parameters -- must not disturb each other (every launch has its own work-queue counter; the
    device tables are cached per parameter set and never overwritten)."""
    import torch
    hb, oracle = mods
    c, _ = _cfg(hb, "cfg2", n=301)
    ny, nx = c["ny"], c["nx"]
    dev = torch.device("cuda:0")
    sc = hb.resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
    vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
    mask = torch.ones((ny, nx), dtype=torch.uint8, device=dev)
    kw_a = dict(dist_search=c["dist_search"], hori_acc=0.25)
    kw_b = dict(dist_search=20.0, hori_acc=0.5)
    ref_a = torch.full((ny, nx, 360), float("nan"), device=dev); ref_b = torch.full((ny, nx, 90), float("nan"), device=dev)
    sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], ref_a, 0, ny, **kw_a)
    sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], ref_b, 0, ny, **kw_b)
    torch.cuda.synchronize()
    out_a = torch.full((ny, nx, 360), float("nan"), device=dev); out_b = torch.full((ny, nx, 90), float("nan"), device=dev)
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    half = ny // 2
    sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], out_a, 0, half, stream=s1, **kw_a)
    sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], out_b, 0, ny, stream=s2, **kw_b)
    sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], out_a, half, ny, stream=s3, **kw_a)
    torch.cuda.synchronize()
    assert torch.equal(out_a, ref_a) and torch.equal(out_b, ref_b)
    sc.close()


def test_block_sharding_full_size_cfg2(mods):
    """Multi-GPU partition on one GPU, BASELINE configs[1] at full size: two interleaved block shards (packed send
    buffers, joined like the all-gather joins them) and three in-place shards both reproduce the
    unsharded pass bit for bit, with the same number of casts."""
    import torch
    hb, oracle = mods
    from horayzon_b200 import sharding
    c, _ = _cfg(hb, "cfg2")
    ny, nx, K = c["ny"], c["nx"], c["azim_num"]
    dev = torch.device("cuda:0")
    sc = hb.resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
    vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
    mask = torch.ones((ny, nx), dtype=torch.uint8, device=dev)
    full = torch.full((ny, nx, K), float("nan"), device=dev)
    r0 = sc.stats()["rays"]
    sc.horizon_gridded(vn, vno, mask, c["offset_0"], c["offset_1"], full, 0, ny, dist_search=c["dist_search"])
    torch.cuda.synchronize()
    rays_full = sc.stats()["rays"] - r0
    world = 2
    per = sharding.padded_block_rows(ny, world)
    gathered = torch.full((world * per, nx, K), float("nan"), device=dev)
    r0 = sc.stats()["rays"]
    for r in range(world):
        assert hb.resident.shard_rows(ny, r, world) == sharding.shard_block_rows(ny, r, world)
        sc.horizon_gridded_sharded(vn, vno, mask, c["offset_0"], c["offset_1"], gathered[r * per:(r + 1) * per], r, world, K,
                                   packed=True, dist_search=c["dist_search"])
    torch.cuda.synchronize()
    assert sc.stats()["rays"] - r0 == rays_full
    assert torch.equal(sharding.unpack_blocks(gathered, ny, world), full)
    del gathered
    inplace = torch.full((ny, nx, K), float("nan"), device=dev)
    for r in range(3):
        sc.horizon_gridded_sharded(vn, vno, mask, c["offset_0"], c["offset_1"], inplace, r, 3, K, packed=False,
                                   dist_search=c["dist_search"])
    torch.cuda.synchronize()
    assert torch.equal(inplace, full)
    sc.close()


def test_full_size_cfg3_shadow_rows_against_oracle(mods):
    """BASELINE configs[2] at FULL size (3601 x 3601 shadow map, 26 M triangles): shadow codes and sw_dir_cor on
    eight rows spread over the domain (rims included) x four sun positions against the CPU oracle, without and
    with refraction.  Without refraction bit-exact; with refraction the differing cells are COUNTED and
    reported (glibc vs CUDA libm last-ulp differences can flip a grazing ray)."""
    hb, oracle = mods
    syn = hb.synthetic
    c = syn.make_config("cfg3")
    ny, nx = c["ny"], c["nx"]
    tilt = syn.tilt_vectors(c["x"], c["y"], c["z"], c["offset_0"])
    enl = (1.0 / np.maximum(tilt[..., 2], 1e-3)).astype(np.float32)
    elev = np.ascontiguousarray(c["z"][c["offset_0"]:c["offset_0"] + ny, c["offset_1"]:c["offset_1"] + nx])
    rows = sorted(set(np.round(np.linspace(0, ny - 1, 8)).astype(int).tolist()))
    full_mask = np.ones((ny, nx), np.uint8)
    row_mask = np.zeros((ny, nx), np.uint8); row_mask[rows] = 1      # the oracle casts rays for these rows only
    suns = syn.sun_positions_diurnal(288)[[30, 80, 150, 230]]
    for refrac in (False, True):
        tg, to = hb.shadow.Terrain(), oracle.Terrain()
        tg.initialise(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["offset_0"], c["offset_1"], tilt, c["vec_norm"], enl, elev,
                      full_mask, refrac_cor=refrac)
        to.initialise(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"], c["offset_0"], c["offset_1"], tilt, c["vec_norm"], enl, elev,
                      row_mask, refrac_cor=refrac)
        n_code = n_sw = 0
        for s in suns:
            a = np.empty((ny, nx), np.uint8); b = np.empty((ny, nx), np.uint8)
            tg.shadow(s, a); to.shadow(s, b)
            fa = np.empty((ny, nx), np.float32); fb = np.empty((ny, nx), np.float32)
            tg.sw_dir_cor(s, fa); to.sw_dir_cor(s, fb)
            n_code += int((a[rows] != b[rows]).sum())
            if refrac:
                n_sw += int((~np.isclose(fa[rows], fb[rows], rtol=2e-6, atol=2e-6)).sum())
            else:
                n_sw += int((fa[rows] != fb[rows]).sum())
            assert 0 < (a == 2).mean() < 1 or (a == 1).mean() > 0     # a real mix of lit / shaded cells
        total = len(suns) * len(rows) * nx
        print("cfg3 full size, refrac_cor=%s: %d shadow codes, %d sw_dir_cor values of %d differ" % (refrac, n_code, n_sw, total))
        if refrac:
            assert n_code <= 2e-5 * total + 1 and n_sw <= 2e-5 * total + 1
        else:
            assert n_code == 0 and n_sw == 0
        del tg, to


def test_full_size_cfg4p_rows_against_oracle(mods):
    """The north-star workload at FULL size (6000 x 6000 DEM, 72 M triangles, 0.77 GB BVH that no longer fits L2,
    16-bit box quantisation over a 12 km extent): rim, quarter and centre rows against the CPU oracle, bit for
    bit, with equal cast counts and without a single full-stack fallback."""
    import torch
    hb, oracle = mods
    c, _ = _cfg(hb, "cfg4p")
    ny, nx, K = c["ny"], c["nx"], c["azim_num"]
    rows = [0, ny // 4, ny // 2]
    dev = torch.device("cuda:0")
    sc = hb.resident.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
    vn = torch.from_numpy(c["vec_norm"]).to(dev); vno = torch.from_numpy(c["vec_north"]).to(dev)
    mask = torch.ones((1, nx), dtype=torch.uint8, device=dev)
    outs = []
    r0 = sc.stats()["rays"]
    for r in rows:
        o = torch.full((1, nx, K), float("nan"), device=dev)
        # a one-row inner domain whose offset selects the row: only these three rows are computed on the GPU
        sc.horizon_gridded(vn[r:r + 1], vno[r:r + 1], mask, c["offset_0"] + r, c["offset_1"], o, 0, 1,
                           dist_search=c["dist_search"])
        outs.append(o)
    torch.cuda.synchronize()
    st = sc.stats()
    rays_gpu = st["rays"] - r0
    assert st["fallback_packets"] == 0
    osc = oracle.Scene(c["vert_grid"], c["dem_dim_0"], c["dem_dim_1"])
    h_cpu, rays_cpu = osc.horizon_rows(rows, c["vec_norm"], c["vec_north"], c["offset_0"], c["offset_1"], c["dist_search"],
                                       azim_num=K, return_rays=True)
    osc.close()
    for i, r in enumerate(rows):
        _assert_same(outs[i].cpu().numpy()[0], h_cpu[i], "cfg4p full size, row %d" % r)
    assert rays_gpu == rays_cpu
    sc.close()


def test_multi_gpu_host_api(mods):
    """hzb_horizon_gridded_multi behind horizon_gridded(devices=): the inner domain dealt out in 4-row blocks to
    several shards (one host thread each; as many GPUs as the box has, shards round-robin over them) gives
    the single-GPU result bit for bit, incl. the fused SVF, masks and a row count that is no multiple of 4."""
    hb, oracle = mods
    c, args = _cfg(hb, "cfg2", n=203)          # 201 inner rows: the last block is ragged
    tilt = hb.synthetic.tilt_vectors(c["x"], c["y"], c["z"], c["offset_0"])
    rng = np.random.default_rng(11)
    mask = (rng.random((c["ny"], c["nx"])) < 0.9).astype(np.uint8)
    kw = dict(azim_num=90, mask=mask, hori_fill=-2.0)
    h1, az1, s1 = hb.horizon.horizon_gridded(*args, svf_vec_tilt=tilt, **kw)
    rays1 = hb.resident.last_stats()["rays"]
    for shards in (2, 3, 5):
        h, az, s = hb.horizon.horizon_gridded(*args, svf_vec_tilt=tilt, devices=0, _shards=shards, **kw)
        assert np.array_equal(h, h1) and np.array_equal(az, az1) and np.array_equal(s, s1), shards
        assert hb.resident.last_stats()["rays"] == rays1
    h, az = hb.horizon.horizon_gridded(*args, devices=0, **kw)          # one shard per visible GPU
    assert np.array_equal(h, h1)
    import torch
    if torch.cuda.device_count() >= 2:                                    # the NCCL all-gather path needs >= 2 GPUs
        h, az, s = hb.horizon.horizon_gridded(*args, svf_vec_tilt=tilt, devices=0, device_gather=True, **kw)
        assert np.array_equal(h, h1) and np.array_equal(s, s1)


def test_multi_terrain_shards_sun_positions(mods):
    """MultiTerrain: the terrain replicated per GPU (here: as many replicas as GPUs, at least two), the sun
    positions of a batch dealt out to the replicas -- same maps as one Terrain."""
    import torch
    hb, oracle = mods
    vg, n, rim, tilt, norm, enl, elev, mask = _terrain_inputs(hb)
    one = hb.shadow.Terrain()
    one.initialise(vg, n, n, rim, rim, tilt, norm, enl, elev, mask)
    suns = hb.synthetic.sun_positions_diurnal(13)
    ref_s, ref_f = one.shadow_batch(suns), one.sw_dir_cor_batch(suns)
    ndev = torch.cuda.device_count()
    mt = hb.multi.MultiTerrain(devices=list(range(ndev)) if ndev >= 2 else [0, 0, 0])
    mt.initialise(vg, n, n, rim, rim, tilt, norm, enl, elev, mask)
    assert np.array_equal(mt.shadow_batch(suns), ref_s)
    assert np.array_equal(mt.sw_dir_cor_batch(suns), ref_f, equal_nan=True)


def test_quantised_output_is_lossless(mods, dbg):
    """Scope row 8f-4: 16-bit table indices + the first azimuth's float reproduce the float array of
    horizon_gridded bit for bit (masked cells, fill value, a fix-up pass after a forced full stack included)."""
    hb, oracle = mods
    c, args = _cfg(hb, "cfg2", n=301)
    rng = np.random.default_rng(5)
    mask = (rng.random((c["ny"], c["nx"])) < 0.8).astype(np.uint8)
    kw = dict(azim_num=180, mask=mask, hori_fill=-0.5)
    h, az = hb.horizon.horizon_gridded(*args, **kw)
    for lim in (None, 3):
        if lim:
            dbg("stack_limit", lim)
        idx, first, table, az2 = hb.horizon.horizon_gridded_quantised(*args, **kw)
        assert idx.dtype == np.uint16 and idx.shape == h.shape and first.shape == h.shape[:2]
        assert np.array_equal(az, az2) and np.array_equal(table, oracle.tables(180, c["dist_search"], 0.25, -15.0)["elev_ang"])
        assert np.array_equal(hb.horizon.dequantise(idx, first, table), h)
        assert np.all(idx[:, :, 0] == 0xFFFF) and np.all(idx[mask == 0] == 0xFFFF) and np.all(first[mask == 0] == np.float32(-0.5))
        assert idx.nbytes * 2 == h.nbytes
