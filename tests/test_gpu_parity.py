"""GPU parity tests: the CUDA path, called through the C ABI (via the ctypes host mirror), against the
CPU oracle, the reference-generated vectors and the reference's 15 shipped golden images.

Bars (stated here, enforced below):
  * integer / index work and every reference-order kernel (tile construction, points, projected,
    Perlin, texture hooks, exact lattice): BIT-EXACT vs the oracle.
  * the FAST multiband lattice kernel (separable, FMA): max |err| <= 1e-5 * (tile max - tile min)
    per unit of band weight (north_star tolerance).
"""
import numpy as np
import pytest

import experiment_cases as ex
import wnpkg

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape or got.size == want.size, what
    bad = bits(got).ravel() != bits(want).ravel()
    if bad.any():
        i = int(np.flatnonzero(bad)[0])
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} floats differ; first at {i}: "
                             f"{got.ravel()[i]!r} vs {want.ravel()[i]!r}")


@pytest.fixture(scope="module")
def wn():
    return wnpkg.load()


@pytest.fixture(scope="module")
def gpu_tiles(wn):
    t2 = wn.WaveletNoise(128, 12345)
    t2.generateNoiseTile2D()
    t3 = wn.WaveletNoise(128, 12345)
    t3.generateNoiseTile3D()
    return {2: t2, 3: t3}


# ---------------------------------------------------------------------------------------------
# tile construction
# ---------------------------------------------------------------------------------------------
def test_tile_n128_bit_exact(gpu_tiles, tiles128):
    assert_bits(gpu_tiles[3].getNoiseCoefficients(), tiles128[3], "3D tile n=128")
    assert_bits(gpu_tiles[2].getNoiseCoefficients(), tiles128[2], "2D tile n=128")


@pytest.mark.parametrize("n", [2, 4, 8, 16, 30, 31, 34, 62, 100, 256])
@pytest.mark.parametrize("dims", [2, 3])
def test_tile_sizes_bit_exact(wn, oracle, n, dims):
    if dims == 3 and n == 256:
        pytest.skip("n=256 3D covered by test_tile_n256_3d")
    w = wn.WaveletNoise(n, 1000 + n)
    (w.generateNoiseTile3D if dims == 3 else w.generateNoiseTile2D)()
    want = oracle.generate_tile(n, 1000 + n, dims)
    assert w.getTileSize() == oracle.adjust(n)
    assert_bits(w.getNoiseCoefficients(), want, f"tile n={n} dims={dims}")


def test_tile_n256_3d(wn, oracle):
    # BASELINE config 2's second size.  Same Gaussian field on both sides (oracle fill), filters on the GPU.
    g = oracle.rng(12345)
    R = oracle.gaussian_fill(g, 256 ** 3)
    w = wn.WaveletNoise(256, 12345)
    w.generateNoiseTile3D(field=R)
    oracle.set_threads(0)
    want = oracle.tile_from_field(R, 256, 3)
    assert_bits(w.getNoiseCoefficients(), want, "3D tile n=256")
    # and the library's own host fill draws the same field
    w2 = wn.WaveletNoise(256, 12345)
    assert_bits(w2.gaussian_field(1 << 20), R[:1 << 20], "gaussian field")


def test_second_generate_continues_rng_stream(wn, oracle):
    w = wn.WaveletNoise(16, 99)
    w.generateNoiseTile2D()
    w.generateNoiseTile3D()
    g = oracle.rng(99)
    oracle.generate_tile(16, 99, 2, g)
    assert_bits(w.getNoiseCoefficients(), oracle.generate_tile(16, 99, 3, g), "second generate")


def test_seeded_build_and_device_field(wn, oracle, tiles128):
    """wn_tile_build_seeded: MT19937 + libstdc++ polar method + glibc-logf restatement ON THE GPU, bit-exact."""
    import torch
    w = wn.WaveletNoise(128, 12345)
    w.generate_seeded(3)
    assert_bits(w.getNoiseCoefficients(), tiles128[3], "device-seeded 3D tile n=128")
    w.generate_seeded(2)
    assert_bits(w.getNoiseCoefficients(), tiles128[2], "device-seeded 2D tile n=128")
    for n, dims, seed in ((2, 2, 1), (8, 3, 77), (30, 3, 807), (31, 2, 5), (100, 3, 4242)):
        w = wn.WaveletNoise(n, seed)
        w.generate_seeded(dims)
        assert_bits(w.getNoiseCoefficients(), oracle.generate_tile(n, seed, dims), f"device-seeded n={n} dims={dims}")
    # device-resident Gaussian field in, nothing crosses PCIe
    R = oracle.gaussian_fill(oracle.rng(5), 32 ** 3)
    w = wn.WaveletNoise(32, 5)
    w.generateNoiseTile3D(field=torch.from_numpy(R).cuda())
    w.ctx.synchronize()
    assert_bits(w.getNoiseCoefficients(), oracle.tile_from_field(R, 32, 3), "device field")


def test_seeded_build_n256(wn, oracle):
    w = wn.WaveletNoise(256, 12345)
    w.generate_seeded(3)
    oracle.set_threads(0)
    assert_bits(w.getNoiseCoefficients(), oracle.generate_tile(256, 12345, 3), "device-seeded 3D tile n=256")


def test_odd_offset_flag_matches_paper_restatement(wn, oracle):
    R = oracle.gaussian_fill(oracle.rng(7), 20 ** 3)
    w = wn.WaveletNoise(20, 7, flags=wn.WN_TILE_ODD_OFFSET)
    w.generateNoiseTile3D(field=R)
    want = oracle.odd_offset3d(oracle.tile_from_field(R, 20, 3), 20)
    assert_bits(w.getNoiseCoefficients(), want, "odd-offset tile")


def test_tile_cache_round_trip(wn, oracle, tmp_path):
    """Tile cache: a miss generates on the GPU and writes the file, a hit uploads it; both objects hold the oracle's
    tile bit for bit and continue the generator stream identically (the second tile of each equals the oracle's)."""
    g = oracle.rng(4242)
    want1 = oracle.generate_tile(32, 4242, 3, g)
    want2 = oracle.generate_tile(32, 4242, 3, g)
    a = wn.WaveletNoise(32, 4242)
    assert a.generate_cached(3, str(tmp_path)) is False
    b = wn.WaveletNoise(32, 4242)
    assert b.generate_cached(3, str(tmp_path)) is True
    for w in (a, b):
        assert_bits(w.getNoiseCoefficients(), want1, "cached tile")
        w.generateNoiseTile3D()
        assert_bits(w.getNoiseCoefficients(), want2, "tile after the cached one")
    with pytest.raises(wn.WnError):
        b.generate_cached(3, str(tmp_path))


def test_gpu_stats_match_the_viewer_json_ranges(wn, golden_dir):
    """The viewer JSON's original_range (threejs/convert_raw_to_json.py:46-49: float64 min / max / mean / std of the
    float32 image) against wn_stats_compute on the 15 shipped images: min and max exactly, mean and variance to float32
    rounding (calculateStats accumulates in double and narrows to float, WaveletNoise.cpp:268-288)."""
    ctx = wn.default_context()
    for octave in ex.OCTAVES:
        for kind in ("w2d", "w3d", "wproj", "p2d", "p3d"):
            img = ex.load_raw(golden_dir, kind, octave).ravel()
            st = ctx.stats(img)
            a = img.astype(np.float64)
            assert st.min_val == np.float32(a.min()) and st.max_val == np.float32(a.max())
            assert abs(st.avg - a.mean()) <= 1e-6 and abs(st.var - a.std() ** 2) <= 2e-6 * max(1.0, a.std() ** 2)


def test_stats(wn, oracle, gpu_tiles, tiles128):
    st = gpu_tiles[3].ctx.stats(tiles128[3])
    avg, var, mn, mx = oracle.stats(tiles128[3])
    assert st.min_val == mn and st.max_val == mx
    assert abs(st.var - var) <= 2e-7 and abs(st.avg - avg) <= 1e-9
    empty = gpu_tiles[3].ctx.stats(np.empty(0, np.float32))
    assert empty.avg == 0 and empty.var == 0


# ---------------------------------------------------------------------------------------------
# the 15 golden images (config 1) -- through the experiment drop-in
# ---------------------------------------------------------------------------------------------
def test_experiment_driver_reproduces_all_15_goldens_bit_exact(wn, golden_dir, tmp_path):
    exp = wnpkg.load_sub("experiment")
    images = exp.main(str(tmp_path))
    assert len(images) == 15
    for octave in ex.OCTAVES:
        for kind in ("w2d", "w3d", "wproj", "p2d", "p3d"):
            name = ex.raw_name(kind, octave)
            want = ex.load_raw(golden_dir, kind, octave)
            got = np.fromfile(tmp_path / name, dtype="<f4").reshape(256, 256)
            assert_bits(got, want, name)


# ---------------------------------------------------------------------------------------------
# evaluators vs reference-generated vectors and the oracle
# ---------------------------------------------------------------------------------------------
def test_points_match_reference_vectors(wn, gpu_tiles, ref_vectors):
    rv = ref_vectors
    pts = rv["pts3"]
    assert_bits(gpu_tiles[3].evaluate3D_points(pts), rv["eval3d"], "evaluate3D")
    assert_bits(gpu_tiles[2].evaluate2D_points(pts[:, :2].copy()), rv["eval2d"], "evaluate2D")
    sel = slice(0, 2048)
    assert_bits(gpu_tiles[3].evaluate3DProjected_points(pts[sel], rv["normals"][sel]), rv["proj_pernormal"], "proj/normal")
    assert_bits(gpu_tiles[3].evaluate3DProjected_points(pts[sel], rv["nshared"]), rv["proj_shared"], "proj shared")
    for seed in (12345, 5489):
        assert_bits(wn.PerlinNoise(seed).noise_points(pts[:5376]), rv[f"perlin_{seed}"], f"perlin {seed}")


def test_projected_far_from_origin_and_axis_normals(wn, oracle, gpu_tiles, tiles128, ref_vectors):
    """k_proj skips rows of the candidate box that cannot contribute; the skip must never drop a tap the reference
    weighs above its 1e-6 threshold.  |p| ~ 1e3 .. 1e6 (float spacing up to 1/16), normals exactly on an axis, within
    1e-4 / 1e-6 of one, and generic: bit-exact against vectors generated by the unmodified reference, for per-point
    normals, for each normal shared by a batch, and on the non-power-of-two tile."""
    rv = ref_vectors
    hp, hn = rv["proj_huge_pts"], rv["proj_huge_normals"]
    assert_bits(gpu_tiles[3].evaluate3DProjected_points(hp, hn), rv["proj_huge"], "projected, huge |p|")
    for k in range(0, hp.shape[0], 7):                       # shared-normal entry point, one normal at a time
        got = gpu_tiles[3].evaluate3DProjected_points(hp[k:k + 1], hn[k])
        assert_bits(got, rv["proj_huge"][k:k + 1], f"projected shared normal {k}")
    w = wn.WaveletNoise(30, 807)
    w.generateNoiseTile3D()
    assert_bits(w.evaluate3DProjected_points(hp, hn), rv["proj_huge_n30"], "projected n=30, huge |p|")
    # more of the same against the oracle (pinned on the vectors above by tests/test_oracle_cpu.py)
    rs = np.random.RandomState(77)
    for mag in (1e3, 1e4, 1e5, 1e6):
        p = (rs.uniform(-1, 1, (4096, 3)) * mag).astype(np.float32)
        nv = np.eye(3)[rs.randint(0, 3, 4096)] * rs.choice([-1.0, 1.0], (4096, 1)) + rs.normal(size=(4096, 3)) * \
            rs.choice([0.0, 1e-6, 1e-4, 1e-2], (4096, 1))
        nv = (nv / np.linalg.norm(nv, axis=1, keepdims=True)).astype(np.float32)
        want = oracle.eval3d_projected_points(tiles128[3], 128, p, nv)
        assert_bits(gpu_tiles[3].evaluate3DProjected_points(p, nv), want, f"projected |p|~{mag:g}")


def test_non_pow2_tile_evaluation(wn, oracle, ref_vectors):
    rv = ref_vectors
    w = wn.WaveletNoise(30, 807)
    w.generateNoiseTile3D()
    assert_bits(w.evaluate3D_points(rv["pts3"]), rv["eval3d_n30"], "evaluate3D n=30")
    assert_bits(w.evaluate3DProjected_points(rv["pts3"][:512], rv["nshared"]), rv["proj_n30"], "projected n=30")


def test_scalar_api_and_empty_tile(wn, oracle, gpu_tiles, tiles128):
    p = [3.25, -7.5, 100.125]
    assert gpu_tiles[3].evaluate3D(p) == oracle.eval3d(tiles128[3], 128, p)
    assert gpu_tiles[2].evaluate2D(p[:2]) == oracle.eval2d(tiles128[2], 128, p[:2])
    assert gpu_tiles[3].evaluate3DProjected(p, [.6, 0, .8]) == oracle.eval3d_projected(tiles128[3], 128, p, [.6, 0, .8])
    fresh = wn.WaveletNoise(128, 1)
    assert fresh.evaluate3D(p) == 0.0 and fresh.evaluate2D(p[:2]) == 0.0 and fresh.getNoiseCoefficients().size == 0
    assert gpu_tiles[3].evaluate3D_points(np.empty((0, 3), np.float32)).size == 0
    with pytest.raises(wn.WnError):
        gpu_tiles[2].evaluate3D_points(np.zeros((1, 3), np.float32))       # 2D tile, 3D evaluator


def test_texture_hooks(wn, ref_vectors):
    rv = ref_vectors
    tp = rv["tex_pts"]
    wt = wn.wavelet_texture(1.0, 4, True)
    assert_bits(wt.values(tp), rv["tex_wavelet_s1_o4"], "wavelet_texture s1 o4")
    assert wt.value(0, 0, tp[0])[0] == rv["tex_wavelet_s1_o4"][0]
    wt2 = wn.wavelet_texture(0.37, 3, True)
    assert_bits(wt2.values(tp), rv["tex_wavelet_s0.37_o3"], "wavelet_texture s.37 o3")
    assert_bits(wn.wavelet_texture(1.0, 4, False).values(tp), rv["tex_wavelet2d_s1_o4"], "wavelet_texture 2D s1 o4")
    assert_bits(wn.wavelet_texture(0.37, 3, False).values(tp), rv["tex_wavelet2d_s0.37_o3"], "wavelet_texture 2D s.37 o3")
    assert_bits(wn.noise_texture(1.0, 4).values(tp), rv["tex_perlin_s1_o4"], "noise_texture s1 o4")
    assert_bits(wn.noise_texture(0.37, 5).values(tp), rv["tex_perlin_s0.37_o5"], "noise_texture s.37 o5")


def test_device_buffers_equal_host_buffers(wn, gpu_tiles, ref_vectors):
    import torch
    gpu_tiles[3].ctx.use_torch_stream()
    try:
        pts = torch.from_numpy(ref_vectors["pts3"]).cuda()
        out = gpu_tiles[3].evaluate3D_points(pts)
        assert out.is_cuda
        assert_bits(out.cpu().numpy(), ref_vectors["eval3d"], "device-space evaluate3D")
    finally:
        gpu_tiles[3].ctx.set_stream(None)


# ---------------------------------------------------------------------------------------------
# multiband lattice (config 3) and affine grids (config 4)
# ---------------------------------------------------------------------------------------------
BANDS = np.array([2.0 * 2 ** b for b in range(4, 9)], np.float32)          # q_b = 2 p 2^b
WEIGHTS = np.array([2.0 ** -(b - 4) for b in range(4, 9)], np.float32)
POST = np.float32(1.0 / np.sqrt(np.float32((WEIGHTS * WEIGHTS).sum()) * np.float32(0.18402)))


def lattice_axis(idx, size=1024):
    return (np.asarray(idx, np.float32) / np.float32(size)) * np.float32(4.0)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_multiband_lattice_subvolumes(wn, oracle, gpu_tiles, tiles128, mode):
    rs = np.random.RandomState(3)
    rng = float(tiles128[3].max() - tiles128[3].min())
    tol = 1e-5 * rng                    # strict north_star bound (measured 1.4e-6), not inflated by sum(w) * post
    for trial in range(6):
        x0, y0, z0 = rs.randint(0, 1024 - 40, 3)
        nx, ny, nz = (40, 33, 9) if trial % 2 else (37, 16, 12)
        xs, ys, zs = (lattice_axis(np.arange(a, a + m)) for a, m in ((x0, nx), (y0, ny), (z0, nz)))
        want = oracle.multiband3d_lattice(tiles128[3], 128, xs, ys, zs, BANDS, WEIGHTS, POST)
        got = gpu_tiles[3].multiband3D_lattice(xs, ys, zs, BANDS, WEIGHTS, float(POST),
                                               mode=wn.WN_EVAL_EXACT if mode == "exact" else wn.WN_EVAL_FAST)
        if mode == "exact":
            assert_bits(got, want, "exact multiband lattice")
        else:
            err = np.abs(got - want).max()
            assert err <= tol, (err, tol)


def test_multiband_lattice_irregular_axes_fast(wn, oracle, gpu_tiles, tiles128):
    # non-uniform, non-monotone, negative and large coordinates; ragged sizes; single band
    rs = np.random.RandomState(11)
    xs = np.concatenate([rs.uniform(-50, 50, 77), [0.0, 0.5, -0.5, 127.5, 128.0, 1e4]]).astype(np.float32)
    ys = rs.uniform(-3, 3, 19).astype(np.float32)
    zs = np.sort(rs.uniform(-10, 300, 7)).astype(np.float32)
    rng = float(tiles128[3].max() - tiles128[3].min())
    for bs, w in (([1.0], [1.0]), ([1.0, 2.0, 4.0], [1.0, 0.5, 0.25])):
        want = oracle.multiband3d_lattice(tiles128[3], 128, xs, ys, zs, bs, w, 1.0)
        got = gpu_tiles[3].multiband3D_lattice(xs, ys, zs, bs, w, 1.0)
        assert np.abs(got - want).max() <= 1e-5 * rng * sum(w)
    assert gpu_tiles[3].multiband3D_lattice(xs, ys, np.empty(0, np.float32), [1.0], [1.0]).size == 0


def test_multiband_full_size_properties(wn, oracle, gpu_tiles, tiles128):
    """Config 3 at full 1024^2 x 8 slab size (device output): periodicity and oracle spot checks."""
    import torch
    t = gpu_tiles[3]
    ax = lattice_axis(np.arange(1024))
    zs = lattice_axis(np.arange(512, 520))
    out = t.multiband3D_lattice(ax, ax, zs, BANDS, WEIGHTS, float(POST), device_out=True)
    t.ctx.synchronize()
    out = out.cpu().numpy()
    rs = np.random.RandomState(5)
    idx = rs.randint(0, 1024, (4096, 2))
    kz = rs.randint(0, 8, 4096)
    pts = np.stack([ax[idx[:, 0]], ax[idx[:, 1]], zs[kz]], -1)
    want = oracle.multiband3d_points(tiles128[3], 128, pts, BANDS, WEIGHTS, POST)
    rng = float(tiles128[3].max() - tiles128[3].min())
    assert np.abs(out[kz, idx[:, 1], idx[:, 0]] - want).max() <= 1e-5 * rng * float(WEIGHTS.sum()) * float(POST)
    # single band 4 over one full period (1024 samples = 128 cells): mean of N over a period ~ tile mean
    one = t.multiband3D_lattice(ax, ax, zs[:1], BANDS[:1], [1.0], 1.0)
    assert abs(float(one.mean())) < 0.05


def test_projected_and_perlin_affine_grid_config4(wn, oracle, gpu_tiles, tiles128):
    """Config 4 on a 96x80 window of the 8192^2 plane, oblique basis; coordinates are built un-fused on both sides."""
    f = np.float32
    nrm = (np.array([1, 2, 3], np.float64) / np.sqrt(14.0)).astype(f)
    e1 = (np.array([2, -1, 0], np.float64) / np.sqrt(5.0)).astype(f)
    e2 = (np.array([3, 6, -5], np.float64) / np.sqrt(70.0)).astype(f)
    origin = np.array([0, 0, 1], f)
    us = (np.arange(4000, 4096, dtype=f) / f(8192)) * f(4)
    vs = (np.arange(100, 180, dtype=f) / f(8192)) * f(4)
    pre = f(2.0 * 2 ** 4)
    inv = f(1.0) / np.sqrt(f(0.296))
    got = gpu_tiles[3].evaluate3DProjected_grid(origin, e1, us, e2, vs, nrm, float(pre), float(inv))
    # same coordinate formula in numpy float32: ((o + u*e1) + v*e2) * pre
    U, V = np.meshgrid(us, vs)
    P = ((origin[None, None, :] + U[..., None] * e1) + V[..., None] * e2).astype(f) * pre
    want = oracle.eval3d_projected_points(tiles128[3], 128, P.reshape(-1, 3), nrm, 1.0, inv).reshape(80, 96)
    assert_bits(got, want, "projected affine grid")
    got3 = gpu_tiles[3].evaluate3D_grid(origin, e1, us, e2, vs, float(pre), 1.0)
    assert_bits(got3, oracle.eval3d_points(tiles128[3], 128, P.reshape(-1, 3)).reshape(80, 96), "evaluate3D affine grid")
    pn = wn.PerlinNoise(12345)
    gp = pn.noise_grid(origin, e1, us, e2, vs, float(f(2.0 ** 4)))
    P4 = ((origin[None, None, :] + U[..., None] * e1) + V[..., None] * e2).astype(f) * f(2.0 ** 4)
    assert_bits(gp, oracle.perlin_points(oracle.perlin_perm(12345), P4.reshape(-1, 3)).reshape(80, 96), "perlin affine grid")


def test_projected_grid_random_geometry(wn, oracle, gpu_tiles, tiles128):
    """k_proj_grid builds one candidate-cell list per 8 x 4 pixel patch (row intervals of the support region) and evaluates
    it two cells at a time on the packed FP32 pipe; which of its paths runs depends on the geometry.  Random planes with
    pixel spacings from 0.004 to 1.7 tile cells (whole-box patches, patches whose lanes' boxes differ by one layer, by
    more than one layer, and patches too incoherent for a list), normals on an axis, within 1e-6 / 1e-3 of one and generic,
    origins up to 300 cells (and 1e5) from zero, image sizes that are not multiples of the patch, a power-of-two and a
    non-power-of-two tile: every pixel BIT-EXACT against the oracle."""
    f = np.float32
    rs = np.random.RandomState(1234)
    small = wn.WaveletNoise(30, 807)
    small.generateNoiseTile3D()
    small_coeffs = small.getNoiseCoefficients()
    tiles = [(gpu_tiles[3], tiles128[3], 128), (small, small_coeffs, 30)]
    case = 0
    for spacing in (0.004, 0.03, 0.11, 0.3, 0.9, 1.7):
        for kind in ("axis", "near6", "near3", "generic"):
            case += 1
            tile, coeffs, n = tiles[case % 2]
            axis = np.eye(3)[rs.randint(0, 3)] * rs.choice([-1.0, 1.0])
            if kind == "axis":
                nv = axis
            elif kind == "near6":
                nv = axis + rs.normal(size=3) * 1e-6
            elif kind == "near3":
                nv = axis + rs.normal(size=3) * 1e-3
            else:
                nv = rs.normal(size=3)
            nv = (nv / np.linalg.norm(nv)).astype(f)
            a = rs.normal(size=3); a /= np.linalg.norm(a)
            b = np.cross(a, rs.normal(size=3)); b /= np.linalg.norm(b)
            e1, e2 = a.astype(f), b.astype(f)
            mag = 1e5 if case % 7 == 0 else 300.0
            origin = (rs.uniform(-1, 1, 3) * mag).astype(f)
            nu, nv_rows = int(rs.randint(33, 90)), int(rs.randint(5, 40))
            us = (np.arange(nu, dtype=f) * f(spacing)).astype(f)
            vs = (np.arange(nv_rows, dtype=f) * f(spacing * 1.3)).astype(f)
            pre, post = f(1.0), f(1.0) / np.sqrt(f(0.296))
            got = tile.evaluate3DProjected_grid(origin, e1, us, e2, vs, nv, float(pre), float(post))
            U, V = np.meshgrid(us, vs)
            P = ((origin[None, None, :] + U[..., None] * e1) + V[..., None] * e2).astype(f) * pre
            want = oracle.eval3d_projected_points(coeffs, n, P.reshape(-1, 3), nv, 1.0, post).reshape(nv_rows, nu)
            assert_bits(got, want, f"case {case}: spacing {spacing}, normal {kind}, n {n}, |o| ~{mag:g}")


def test_config4_full_size_8192_squared(wn, oracle, gpu_tiles, tiles128):
    """BASELINE config 4 at its full size: WProjectedNoise on the 8192^2 oblique plane and Perlin(12345) octave 4 on the
    same grid, device resident.  65 536 random pixels plus two full rows (first / last) and two full columns against
    the oracle, BIT-EXACT (the affine coordinate formula is evaluated un-fused in a fixed order on both sides)."""
    import torch
    f = np.float32
    S = 8192
    nrm = (np.array([1, 2, 3], np.float64) / np.sqrt(14.0)).astype(f)
    e1 = (np.array([2, -1, 0], np.float64) / np.sqrt(5.0)).astype(f)
    e2 = (np.array([3, 6, -5], np.float64) / np.sqrt(70.0)).astype(f)
    origin = np.array([0, 0, 1], f)
    ax = (np.arange(S, dtype=f) / f(S)) * f(4)
    pre_w, pre_p = f(2.0 * 2 ** 4), f(2.0 ** 4)
    inv = f(1.0) / np.sqrt(f(0.296))
    t = gpu_tiles[3]
    proj = t.evaluate3DProjected_grid(origin, e1, ax, e2, ax, nrm, float(pre_w), float(inv), device_out=True)
    perlin = wn.PerlinNoise(12345, t.ctx)
    perl = perlin.noise_grid(origin, e1, ax, e2, ax, float(pre_p), device_out=True)
    t.ctx.synchronize()
    proj, perl = proj.cpu().numpy(), perl.cpu().numpy()
    assert proj.shape == (S, S) and np.isfinite(proj).all() and np.isfinite(perl).all()
    rs = np.random.RandomState(4)
    ii = np.concatenate([rs.randint(0, S, 1 << 16), np.arange(S), np.arange(S), np.zeros(S, int), np.full(S, S - 1)])
    jj = np.concatenate([rs.randint(0, S, 1 << 16), np.zeros(S, int), np.full(S, S - 1), np.arange(S), np.arange(S)])
    P = ((origin[None, :] + ax[ii][:, None] * e1) + ax[jj][:, None] * e2).astype(f)
    want = oracle.eval3d_projected_points(tiles128[3], 128, P * pre_w, nrm, 1.0, inv)
    assert_bits(proj[jj, ii], want, "config 4 projected, 8192^2")
    wantp = oracle.perlin_points(oracle.perlin_perm(12345), P * pre_p)
    assert_bits(perl[jj, ii], wantp, "config 4 Perlin, 8192^2")
    # FP32 fast mode of the Perlin kernel: opt-in, within 1e-5 * range (range of Perlin noise = 2) of the FP64 kernel
    perlin.set_precision(wn.WN_PERLIN_F32)
    fast = perlin.noise_grid(origin, e1, ax, e2, ax, float(pre_p), device_out=True)
    t.ctx.synchronize()
    assert float(np.abs(fast.cpu().numpy() - perl).max()) <= 1e-5 * 2.0
    perlin.set_precision(wn.WN_PERLIN_F64)
    again = perlin.noise_grid(origin, e1, ax[:64], e2, ax[:64], float(pre_p))
    assert_bits(again, perl[:64, :64], "back to FP64")


def test_perlin_double_precision_entry(wn, oracle):
    """PerlinNoise::noise(double, double, double): coordinates that are NOT float-valued, double result (no narrowing)."""
    pn = wn.PerlinNoise(12345)
    perm = oracle.perlin_perm(12345)
    rs = np.random.RandomState(8)
    pts = rs.uniform(-300, 300, (512, 3))
    got = pn.noise_points_f64(pts)
    want = np.array([oracle.perlin_noise(perm, *p) for p in pts])
    assert (got.view(np.uint64) == want.view(np.uint64)).all()
    assert pn.noise(0.1, 0.2, 0.3) == oracle.perlin_noise(perm, 0.1, 0.2, 0.3)
    assert pn.noise(0.3, 0.7) == oracle.perlin_noise(perm, 0.3, 0.7, 0.0)


def test_axis_table_tap_cells_are_the_reference_integers(wn, oracle, gpu_tiles):
    """north_star: "bit-exact tile indexing".  The tap cells of the FAST lattice kernels (k_axis_tables: first cell +
    0..2, wrapped) are compared AS INTEGERS with the 27 tile indices the reference's evaluate3D gathers
    (orc_eval3d_taps restates cpp:194-209), for coordinates that are negative, huge, on cell and half-cell boundaries,
    and for a power-of-two and a non-power-of-two tile edge; the weights must equal the un-fused reference formula."""
    ctx = gpu_tiles[3].ctx
    rs = np.random.RandomState(21)
    coords = np.concatenate([rs.uniform(-300, 300, 2048), rs.uniform(-2, 2, 512), np.round(rs.uniform(-64, 64, 512) * 2) / 2,
                             rs.uniform(-1, 1, 256) * 1e6, [0.0, 0.5, -0.5, 127.5, 128.0, -128.0, 2.0, 2.5]]).astype(np.float32)
    coords = coords[: coords.size // 3 * 3]
    for scale in (1.0, 32.0, 0.37):
        w, first = ctx.axis_entries(coords, scale)
        q = coords * np.float32(scale)
        # reference weights, un-fused float32 (cpp:194-200)
        a = q - np.float32(0.5)
        mid = np.ceil(a)
        t = (mid - a).astype(np.float32)
        w0 = (t * t * np.float32(0.5)).astype(np.float32)
        s1 = (np.float32(1.0) - t).astype(np.float32)
        w2 = (s1 * s1 * np.float32(0.5)).astype(np.float32)
        w1 = ((np.float32(1.0) - w0).astype(np.float32) - w2).astype(np.float32)
        assert_bits(w[:, 0], w0, "w0"); assert_bits(w[:, 1], w1, "w1"); assert_bits(w[:, 2], w2, "w2")
        for n in (128, 30):
            pts = q.reshape(-1, 3)
            f3 = first.reshape(-1, 3)
            for p, f in zip(pts[::7], f3[::7]):
                idx = oracle.eval3d_taps(n, p)                     # 27 linear indices, fz outer, fx inner
                cx = [(int(f[0]) + k) % n for k in range(3)]
                cy = [(int(f[1]) + k) % n for k in range(3)]
                cz = [(int(f[2]) + k) % n for k in range(3)]
                mine = [cx[fx] + n * cy[fy] + n * n * cz[fz] for fz in range(3) for fy in range(3) for fx in range(3)]
                assert list(idx) == mine, (p, n)


def test_band_limits_by_power_spectrum(wn):
    """What experient/analyze.py:398-527 shows as pictures (the paper's Figures 8 and 9), asserted as numbers from a GPU
    FFT on freshly generated 512^2 octave-4 images: 2D wavelet noise is band-limited (>= 93 % of its power inside the
    band, < 0.3 % below half its lower edge); a 2D slice of 3D noise leaks low frequencies (> 3 %); projecting along the
    normal restores the band limit (< 1 %); Perlin noise is not band-limited (no octave holds more than 60 % of its
    power: about half lies above the octave of its lattice)."""
    expm = wnpkg.load_sub("experiment")
    n2, n3, pn = wn.WaveletNoise(128, 12345), wn.WaveletNoise(128, 12345), wn.PerlinNoise(12345)
    n2.generateNoiseTile2D()
    n3.generateNoiseTile3D()
    size, octave = 512, 4
    cells_per_pixel = 4.0 / size * 2.0 ** octave * 2.0         # u = x/size*4, q = 2 u 2^octave
    in2, lo2 = expm.band_energy(expm.generate2DOctaveBandNoise(size, octave, None, n2), cells_per_pixel)
    in3, lo3 = expm.band_energy(expm.generate3DSlicedOctaveBandNoise(size, octave, None, n3), cells_per_pixel)
    inp, lop = expm.band_energy(expm.generate3DProjectedOctaveBandNoise(size, octave, None, n3), cells_per_pixel)
    inperlin, _ = expm.band_energy(expm.generatePerlinNoise2D(size, octave, None, pn), cells_per_pixel / 2.0)
    assert in2 >= 0.93 and lo2 <= 0.003, (in2, lo2)
    assert lo3 >= 0.03, lo3
    assert inp >= 0.90 and lop <= 0.01, (inp, lop)
    assert inperlin <= 0.6, inperlin


def test_paper_wmultibandnoise_signature(wn, oracle, gpu_tiles, tiles128):
    """Cook & DeRose App. 2 WMultibandNoise(p, s, normal, firstBand, nbands, w): band cut-off by the scale s, optional
    projection normal, variance normalisation over all nbands weights -- BIT-EXACT against the statement-by-statement
    restatement over the oracle's evaluate3D / evaluate3DProjected (parity unpinned by the reference: it has no
    multiband function)."""
    t = gpu_tiles[3]
    rs = np.random.RandomState(31)
    pts = np.concatenate([rs.uniform(-3, 3, (3000, 3)), rs.uniform(-200, 200, (1000, 3))]).astype(np.float32)
    w = np.array([1.0, 0.5, 0.25, 0.125, 0.0625], np.float32)
    nrm = (np.array([1, 2, 3], np.float64) / np.sqrt(14.0)).astype(np.float32)
    for s, first, normal in ((-20.0, -2, None), (-3.5, -6, None), (-1.0, -3, nrm), (-20.0, 0, nrm), (5.0, -2, None)):
        got = t.WMultibandNoise(pts, s, normal, first, w)
        want = oracle.wmultiband_points(tiles128[3], 128, pts, s, normal, first, w)
        assert_bits(got, want, f"WMultibandNoise s={s} first={first} normal={normal is not None}")
    assert not t.WMultibandNoise(pts, 5.0, None, -2, w).any()          # every band cut off


def test_large_host_call_is_chunked_consistently(wn, gpu_tiles):
    """A WN_HOST lattice larger than one staging chunk equals the same lattice computed in one device call."""
    t = gpu_tiles[3]
    ax = lattice_axis(np.arange(1024))
    zs = lattice_axis(np.arange(0, 40))                    # 40 Mi samples > 32 Mi chunk
    host = t.multiband3D_lattice(ax, ax, zs, BANDS[:2], WEIGHTS[:2], 1.0)
    dev = t.multiband3D_lattice(ax, ax, zs, BANDS[:2], WEIGHTS[:2], 1.0, device_out=True)
    t.ctx.synchronize()
    assert_bits(host, dev.cpu().numpy(), "chunked host vs device")


def test_folding_non_pow2_tile_and_periods(wn, oracle, monkeypatch):
    """Periodic folding with a non-power-of-two tile (n=30) and non-power-of-two periods (60 / 30 / 15 samples):
    exercises the `%` paths of the brick kernels, the float4 kernel with Lx % 4 == 0 and the scalar one otherwise.
    (Lattices this small only fold when the cost model's per-level launch cost is set to zero.)"""
    monkeypatch.setenv("WN_FOLD_LEVEL_COST", "0")
    tile = oracle.generate_tile(30, 807, 3)
    w = wn.WaveletNoise(30, 807)
    w.generateNoiseTile3D()
    rng = float(tile.max() - tile.min())
    # (242, 61, 33): nx % 4 != 0 -> one-sample-per-lane kernel; (244, 61, 75) and (128, 40, 130): partial x / y / z
    # bricks of the z-streaming kernel, period blocks that wrap inside a brick (periods 15 / 30 / 60 are not
    # multiples of its 32-sample z extent)
    for nx, ny, nz in ((240, 120, 64), (242, 61, 33), (244, 61, 75), (128, 40, 130)):
        xs = np.arange(nx, dtype=np.float32) * np.float32(0.5)
        ys = np.arange(ny, dtype=np.float32) * np.float32(0.5) + np.float32(3.25)
        zs = np.arange(nz, dtype=np.float32) * np.float32(0.5) - np.float32(7.0)
        bs, wts = [1.0, 2.0, 4.0, 0.25], [1.0, 0.5, 0.25, 2.0]
        want = oracle.multiband3d_lattice(tile, 30, xs, ys, zs, bs, wts, 0.7)
        got = w.multiband3D_lattice(xs, ys, zs, bs, wts, 0.7)
        assert np.abs(got - want).max() <= 1e-5 * rng * sum(wts) * 0.7
        exact = w.multiband3D_lattice(xs, ys, zs, bs, wts, 0.7, mode=wn.WN_EVAL_EXACT)
        assert_bits(exact, want, "exact lattice n=30")


def test_sharded_slabs_reassemble_bit_exactly(wn, gpu_tiles):
    """Config 3 contiguous z-slab sharding: the slabs 1/2/4/8 ranks would compute, each with its own call.  A call folds
    the bands that repeat on ITS lattice (a short slab has fewer whole z periods than the volume), but a sample is
    always summed in the canonical band order, so the slabs are bit-identical to the single call."""
    sh = wnpkg.load_sub("sharding")
    t = gpu_tiles[3]
    ax = lattice_axis(np.arange(1024))
    zs = ax[:512:2]                                                # 256 slices, still commensurate with the tile
    full = t.multiband3D_lattice(ax[:256], ax[:256], zs, BANDS, WEIGHTS, float(POST))
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            b, e = sh.slab_range(zs.size, r, world)
            parts.append(t.multiband3D_lattice(ax[:256], ax[:256], zs[b:e], BANDS, WEIGHTS, float(POST)))
        assert_bits(np.concatenate(parts, 0), full, f"{world} slabs")


def test_block_cyclic_shards_reassemble_bit_exactly(wn, gpu_tiles, tiles128):
    """Config 3 block-cyclic z sharding (what bench.py runs at N > 1): every rank evaluates the 32-slice chunks dealt to
    it as one lattice call on its own, non-uniform z axis.  1-GPU and N-GPU volumes are bit-identical (canonical
    summation), and one chunk is checked against the exact kernel."""
    sh = wnpkg.load_sub("sharding")
    t = gpu_tiles[3]
    ax = lattice_axis(np.arange(1024))
    xs = ax[:256]
    full = t.multiband3D_lattice(xs, xs, ax, BANDS, WEIGHTS, float(POST))
    for world in (2, 8):
        got = np.empty_like(full)
        for r in range(world):
            idx = sh.cyclic_slab_indices(ax.size, r, world)
            got[idx] = t.multiband3D_lattice(xs, xs, ax[idx], BANDS, WEIGHTS, float(POST))
        assert_bits(got, full, f"{world} block-cyclic shards")
    idx = sh.cyclic_slab_indices(ax.size, 5, 8)[:32]
    exact = t.multiband3D_lattice(xs, xs, ax[idx], BANDS, WEIGHTS, float(POST), mode=wn.WN_EVAL_EXACT)
    tol = 1e-5 * float(tiles128[3].max() - tiles128[3].min()) * float(WEIGHTS.sum()) * float(POST)
    assert np.abs(full[idx] - exact).max() <= tol


def test_fast_result_does_not_depend_on_folding_or_kernel(wn, gpu_tiles, monkeypatch):
    """The same lattice evaluated with every band per sample (no folding), with a small fold budget, with the default
    one, with the brick4 main kernel and with unsorted band order: bit-identical (canonical summation order)."""
    t = gpu_tiles[3]
    ax = lattice_axis(np.arange(1024))
    xs, ys, zs = ax[:256], ax[:128], ax[:192]
    base = t.multiband3D_lattice(xs, ys, zs, BANDS, WEIGHTS, float(POST))
    for env in ({"WN_FOLD_BUDGET": "0"}, {"WN_FOLD_BUDGET": str(1 << 20)}, {"WN_COL4": "0"}, {"WN_FOLD_NEST": "0"},
                {"WN_BRICK": "4"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got = t.multiband3D_lattice(xs, ys, zs, BANDS, WEIGHTS, float(POST))
        for k in env:
            monkeypatch.delenv(k)
        assert_bits(got, base, str(env))
    perm = [3, 0, 4, 2, 1]
    got = t.multiband3D_lattice(xs, ys, zs, BANDS[perm], WEIGHTS[perm], float(POST))
    assert_bits(got, base, "band order")


def test_replica_kernel_variants_are_bit_identical(wn, oracle, gpu_tiles, tiles128, monkeypatch):
    """k_mb3d_rep (period-block value shared between x / y replicas, shared bands, TMA or cp.async ring) against the
    round-1 z-streaming kernel (WN_REP=0) and the exact kernel: every variant must produce the SAME bits, on whole
    bricks (TMA eligible), partial bricks, lattices whose halves are not replicas, and a z axis that jumps by more than
    three cells inside a brick (the advance-mask fallback)."""
    t = gpu_tiles[3]
    ax = lattice_axis(np.arange(1024))
    rng = float(tiles128[3].max() - tiles128[3].min())
    zjump = np.concatenate([ax[:20], ax[400:420], ax[900:924]])          # 64 slices, jumps of 47.5 and 60 cells at band 4
    cases = [
        (ax, ax[:512], ax[:64], BANDS, WEIGHTS),            # x and y replicas, shared band 5, whole bricks
        (ax[:512], ax[:512], ax[:96], BANDS[1:], WEIGHTS[1:]),   # replicas at stride 256
        (ax[:1000], ax[:500], ax[:40], BANDS, WEIGHTS),     # halves are not replicas; partial bricks
        (ax[:520], ax[:260], ax[:33], BANDS[:2], WEIGHTS[:2]),   # no fold at all (two direct bands), partial bricks
        (ax[:256], ax[:64], zjump, BANDS[:1], WEIGHTS[:1]),      # jumps inside a brick: advance-mask fallback
        (ax[:512], ax[:256], zjump, BANDS, WEIGHTS),
    ]
    variants = [{"WN_REP": "0"}, {}, {"WN_REP_TMA": "0"}, {"WN_REP": "12"}, {"WN_REP": "11"}, {"WN_REP_SHARE": "0"},
                {"WN_REP": "12", "WN_REP_YPW": "4"}, {"WN_FOLD_BUDGET": str(1 << 27), "WN_REP_TMA": "1"},
                {"WN_FOLD_BUDGET": str(1 << 27), "WN_REP": "22", "WN_REP_TMA": "0"},
                {"WN_REP_BY": "16"}, {"WN_REP_BY": "16", "WN_REP": "12"}, {"WN_REP_BY": "16", "WN_REP_SHARE": "0"}]
    for ci, (xs, ys, zs, bs, ws) in enumerate(cases):
        base = None
        for env in variants:
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got = t.multiband3D_lattice(xs, ys, zs, bs, ws, float(POST))
            for k in env:
                monkeypatch.delenv(k)
            if base is None:
                base = got
                exact = t.multiband3D_lattice(xs, ys, zs[:8], bs, ws, float(POST), mode=wn.WN_EVAL_EXACT)
                assert np.abs(got[:8] - exact).max() <= 1e-5 * rng, ci
            else:
                assert_bits(got, base, f"case {ci} {env}")


def test_band_split_in_place_is_bit_identical(wn, oracle, gpu_tiles, tiles128, monkeypatch):
    """Three or more direct bands: the upper bands go through the brick kernel, the lowest one or two are added on top in
    place by the streaming kernel (brick_pass, WN_SPLIT).  The canonical sum makes that the same operations in the same
    order, so the result must equal the single-pass kernels (WN_SPLIT=0) bit for bit -- on lattices that do not fold at
    all (base range 4.1), with partial bricks in x, y and z, on a lattice that folds only its upper bands, through the
    host-buffer entry, and on a tile whose size is not a power of two -- and the exact kernel within the FAST tolerance."""
    t = gpu_tiles[3]
    rng = float(tiles128[3].max() - tiles128[3].min())
    ax41 = (np.arange(1024, dtype=np.float32) / np.float32(1024)) * np.float32(4.1)
    ax = lattice_axis(np.arange(1024))
    cases = [
        (t, ax41, ax41[:256], ax41[:64], BANDS, WEIGHTS),                # whole bricks, TMA ring on the in-place block
        (t, ax41[:1000], ax41[:516], ax41[:40], BANDS, WEIGHTS),         # partial bricks: cp.async ring
        (t, ax41[:512], ax41[:200], ax41[:70], BANDS[:3], WEIGHTS[:3]),  # three fine bands: nothing for the brick kernel
        (t, ax41[:512], ax41[:128], ax41[:33], BANDS[1:], WEIGHTS[1:]),
        (t, ax, ax[:512], ax41[:48], BANDS, WEIGHTS),                    # x and y periodic, z not: no fold, replicas possible
    ]
    ctx = wn.Context(0)
    odd = wn.WaveletNoise(62, 99, ctx)
    odd.generateNoiseTile3D()
    cases.append((odd, ax41[:768], ax41[:130], ax41[:36], BANDS[:4], WEIGHTS[:4]))
    for ci, (tile, xs, ys, zs, bs, ws) in enumerate(cases):
        monkeypatch.setenv("WN_SPLIT", "0")
        base = tile.multiband3D_lattice(xs, ys, zs, bs, ws, float(POST))
        monkeypatch.delenv("WN_SPLIT")
        got = tile.multiband3D_lattice(xs, ys, zs, bs, ws, float(POST))
        assert_bits(got, base, f"case {ci}")
        if tile is t:
            exact = tile.multiband3D_lattice(xs, ys, zs[:6], bs, ws, float(POST), mode=wn.WN_EVAL_EXACT)
            assert np.abs(got[:6] - exact).max() <= 1e-5 * rng, ci
    # the same through device-resident output (no chunking by the host path)
    import torch
    xs, ys, zs = ax41[:640], ax41[:264], ax41[:96]
    monkeypatch.setenv("WN_SPLIT", "0")
    base = t.multiband3D_lattice(xs, ys, zs, BANDS, WEIGHTS, float(POST), device_out=True)
    monkeypatch.delenv("WN_SPLIT")
    got = t.multiband3D_lattice(xs, ys, zs, BANDS, WEIGHTS, float(POST), device_out=True)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int32), base.view(torch.int32))


def test_back_to_back_device_calls_with_tile_rebuilds(wn, monkeypatch):
    """Consecutive device-resident FAST calls without any synchronisation in between (the period-block chain of call
    i+1 runs on the side stream while call i's main kernel is still in flight), with different lattices and the tile
    rebuilt between some of them.  Every result must equal the exact kernel evaluated afterwards, call by call, on a
    second object that replays the same tile sequence."""
    import torch
    monkeypatch.setenv("WN_FOLD_LEVEL_COST", "0")            # fold even on these small lattices: period-block scratch in flight
    ctx = wn.Context(0)
    ctx.use_torch_stream()
    w = wn.WaveletNoise(32, 77, ctx)
    w.generateNoiseTile3D()
    bs, wts = np.array([1.0, 2.0, 4.0, 8.0], np.float32), np.array([1.0, 0.5, 0.25, 0.125], np.float32)
    shapes = [(256, 64, 96), (128, 128, 64), (256, 64, 96), (192, 40, 70), (128, 128, 64), (256, 64, 96), (64, 64, 32)]
    rebuild_before = {2, 3, 6}
    axes = lambda n, off: (np.arange(n, dtype=np.float32) + np.float32(off)) * np.float32(0.125)      # noqa: E731
    outs = []
    for it, (nx, ny, nz) in enumerate(shapes):
        if it in rebuild_before:
            w.generateNoiseTile3D()                          # continues the member RNG stream: a different tile
        outs.append(w.multiband3D_lattice(axes(nx, 0), axes(ny, 3), axes(nz, it), bs, wts, 0.9, device_out=True))
    torch.cuda.synchronize()
    ref = wn.WaveletNoise(32, 77, ctx)
    ref.generateNoiseTile3D()
    for it, (nx, ny, nz) in enumerate(shapes):
        if it in rebuild_before:
            ref.generateNoiseTile3D()
        tile = np.asarray(ref.getNoiseCoefficients())
        want = ref.multiband3D_lattice(axes(nx, 0), axes(ny, 3), axes(nz, it), bs, wts, 0.9, mode=wn.WN_EVAL_EXACT)
        tol = 1e-5 * float(tile.max() - tile.min()) * float(wts.sum()) * 0.9
        assert np.abs(outs[it].cpu().numpy() - want).max() <= tol, it
    ctx.set_stream(None)


def test_argument_errors(wn, gpu_tiles):
    t3 = gpu_tiles[3]
    ax = lattice_axis(np.arange(8))
    with pytest.raises(wn.WnError):
        t3.multiband3D_lattice(ax, ax, ax, np.ones(17, np.float32), np.ones(17, np.float32))     # > 16 bands
    with pytest.raises(ValueError):
        t3.multiband3D_lattice(ax, ax, ax, [1.0, 2.0], [1.0])
    with pytest.raises(wn.WnError):
        w = wn.WaveletNoise(4, 0)
        w.allocate(3)
        w.evaluate3D_points(np.zeros((1, 3), np.float32))                                       # tile never built


def test_config3_full_size_1024_cubed(wn, oracle, gpu_tiles, tiles128):
    """BASELINE config 3 at its full size (1024^3 samples, 4 GiB, device resident).
    (1) EXACT (reference operation order) vs the CPU oracle, BIT-EXACT, on: 64 random 32^3 bricks; the full planes on
        both sides of every period-block wrap (z = 127/128, 255/256, 511/512), of the 32-slice CTA brick edges next to
        them (z = 31/32, 479/480, 991/992) and the last plane (1023); the x = 511/512 and y = 511/512 planes' rows where
        the in-thread replicas of the main kernel meet.
    (2) FAST vs EXACT over the WHOLE volume on the device: max |diff| <= 1e-5 * (tile max - tile min) -- the strict
        north_star bound, NOT inflated by the band weights or the post scale.  With (1) this bounds FAST against the
        reference everywhere, including every seam of the fold / replica / brick machinery.
    (3) linearity: the 5-band result equals the weighted sum of five single-band evaluations; finite, mean ~ 0,
        variance of order 1."""
    import torch
    t = gpu_tiles[3]
    t.ctx.use_torch_stream()
    try:
        ax = lattice_axis(np.arange(1024))
        rng = float(tiles128[3].max() - tiles128[3].min())
        strict = 1e-5 * rng
        full = t.multiband3D_lattice(ax, ax, ax, BANDS, WEIGHTS, float(POST), device_out=True)
        # (2) whole volume, 64 slices at a time
        exact = torch.empty((64, 1024, 1024), dtype=torch.float32, device=full.device)
        worst = 0.0
        seam_planes = [31, 32, 127, 128, 255, 256, 479, 480, 511, 512, 991, 992, 1023]
        kept = {}
        for z0 in range(0, 1024, 64):
            t.multiband3D_lattice(ax, ax, ax[z0:z0 + 64], BANDS, WEIGHTS, float(POST), mode=wn.WN_EVAL_EXACT, out=exact)
            worst = max(worst, float((full[z0:z0 + 64] - exact).abs().max()))
            for z in seam_planes:
                if z0 <= z < z0 + 64:
                    kept[z] = exact[z - z0].cpu().numpy()
        assert worst <= strict, (worst, strict)
        # (1) exact vs oracle, bit for bit
        for z, plane in kept.items():
            want = oracle.multiband3d_lattice(tiles128[3], 128, ax, ax, ax[z:z + 1], BANDS, WEIGHTS, POST)[0]
            assert_bits(plane, want, f"exact plane z={z}")
        rs = np.random.RandomState(9)
        for trial in range(64):
            x0, y0, z0 = (int(v) for v in rs.randint(0, 1024 - 32, 3))
            if trial < 8:                                          # bricks straddling the replica seams x / y = 512
                x0, y0 = (496, y0) if trial % 2 else (x0, 496)
            got = t.multiband3D_lattice(ax[x0:x0 + 32], ax[y0:y0 + 32], ax[z0:z0 + 32], BANDS, WEIGHTS, float(POST),
                                        mode=wn.WN_EVAL_EXACT)
            want = oracle.multiband3d_lattice(tiles128[3], 128, ax[x0:x0 + 32], ax[y0:y0 + 32], ax[z0:z0 + 32], BANDS,
                                              WEIGHTS, POST)
            assert_bits(got, want, f"exact brick {x0},{y0},{z0}")
            sub = full[z0:z0 + 32, y0:y0 + 32, x0:x0 + 32].cpu().numpy()
            assert np.abs(sub - want).max() <= strict
        # the x = 511/512 and y = 511/512 seams of the FAST volume against the oracle on a few planes
        for z in (0, 300, 777):
            want = oracle.multiband3d_lattice(tiles128[3], 128, ax[504:520], ax, ax[z:z + 1], BANDS, WEIGHTS, POST)[0]
            assert np.abs(full[z, :, 504:520].cpu().numpy() - want).max() <= strict
            want = oracle.multiband3d_lattice(tiles128[3], 128, ax, ax[504:520], ax[z:z + 1], BANDS, WEIGHTS, POST)[0]
            assert np.abs(full[z, 504:520, :].cpu().numpy() - want).max() <= strict
        del exact
        # (3)
        acc = torch.zeros_like(full)
        tmp = torch.empty_like(full)
        for b in range(len(BANDS)):
            t.multiband3D_lattice(ax, ax, ax, BANDS[b:b + 1], WEIGHTS[b:b + 1], float(POST), out=tmp)
            acc += tmp
        torch.cuda.synchronize()
        assert float((full - acc).abs().max()) <= strict
        del acc, tmp
        assert bool(torch.isfinite(full).all())
        mean = float(full.double().mean())
        var = float(full.double().var())
        assert abs(mean) < 0.01 and 0.4 < var < 0.7, (mean, var)
    finally:
        t.ctx.set_stream(None)


# ---------------------------------------------------------------------------------------------
# device groups (single-process multi-GPU over the C ABI); N > 1 needs a multi-GPU box
# ---------------------------------------------------------------------------------------------
def _group_sizes():
    import torch
    n = torch.cuda.device_count()
    return [k for k in (1, 2, 4, 8) if k <= n]


def test_device_group_matches_single_context(wn, oracle, gpu_tiles, tiles128):
    """wn_group: tile built on rank 0 and replicated (NCCL broadcast when N > 1), config 3 sharded along z (block-cyclic
    and contiguous slabs), config 4 by row-band, config 5 points by contiguous runs -- every GPU count available on this
    box must reproduce the single-context results BIT FOR BIT (samples are independent; FAST sums canonically)."""
    t = gpu_tiles[3]
    ax = lattice_axis(np.arange(1024))
    xs, ys, zs = ax[:256], ax[:256], ax[:320]
    want = t.multiband3D_lattice(xs, ys, zs, BANDS, WEIGHTS, float(POST))
    f = np.float32
    nrm = (np.array([1, 2, 3], np.float64) / np.sqrt(14.0)).astype(f)
    e1 = (np.array([2, -1, 0], np.float64) / np.sqrt(5.0)).astype(f)
    e2 = (np.array([3, 6, -5], np.float64) / np.sqrt(70.0)).astype(f)
    origin = np.array([0, 0, 1], f)
    us, vs = ax[:96], ax[100:181]
    want_proj = t.evaluate3DProjected_grid(origin, e1, us, e2, vs, nrm, 32.0, 1.5)
    pts = np.random.RandomState(3).uniform(-10, 10, (5001, 3)).astype(f)
    want_tex = t.texture_values(pts, 1.0, 4)
    for n in _group_sizes():
        g = wn.DeviceGroup(n)
        assert g.size == n
        gt = g.tile(128, 3, 12345)
        for sharding in (wn.WN_SHARD_CYCLIC, wn.WN_SHARD_SLAB):
            got, ms = gt.multiband3D_lattice(xs, ys, zs, BANDS, WEIGHTS, float(POST), sharding=sharding)
            assert_bits(got, want, f"group of {n}, sharding {sharding}")
            assert ms > 0.0
        exact, _ = gt.multiband3D_lattice(xs[:64], ys[:64], zs[:70], BANDS, WEIGHTS, float(POST), mode=wn.WN_EVAL_EXACT,
                                          sharding=wn.WN_SHARD_SLAB)
        assert_bits(exact, oracle.multiband3d_lattice(tiles128[3], 128, xs[:64], ys[:64], zs[:70], BANDS, WEIGHTS, POST),
                    f"group of {n}, exact")
        got, _ = gt.evaluate3DProjected_grid(origin, e1, us, e2, vs, nrm, 32.0, 1.5)
        assert_bits(got, want_proj, f"group of {n}, projected row-bands")
        assert_bits(gt.texture_values(pts, 1.0, 4), want_tex, f"group of {n}, texture points")
        g.close()
