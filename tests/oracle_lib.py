"""ctypes access to the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

`Oracle`  wraps oracle/libwn_oracle.so  (our plain-C restatement, oracle/wn_oracle.c).
`RefLib`  wraps oracle/_ref/libwnref.so (the unmodified reference sources compiled by
          oracle/Makefile; absent if it was never built).
Nothing under the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libwn_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libwnref.so")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


def _fp(a):
    return a.ctypes.data_as(f32p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def build_oracle():
    """Compile oracle/ (and oracle/_ref when /root/reference exists). Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


class OrcRng(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int), ("has_saved", C.c_int),
                ("saved", C.c_float), ("draws", C.c_uint64)]


class OrcStats(C.Structure):
    _fields_ = [("avg", C.c_float), ("var", C.c_float), ("min_val", C.c_float), ("max_val", C.c_float)]


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = self.L = C.CDLL(ORACLE_SO)
        L.orc_rng_u32.restype = C.c_uint32
        L.orc_rng_canonical.restype = C.c_float
        L.orc_rng_normal.restype = C.c_float
        L.orc_eval2d.restype = C.c_float
        L.orc_eval3d.restype = C.c_float
        L.orc_eval3d_projected.restype = C.c_float
        L.orc_perlin_noise.restype = C.c_double
        L.orc_perlin_noise.argtypes = [i32p, C.c_double, C.c_double, C.c_double]
        L.orc_wavelet_texture_value.restype = C.c_double
        L.orc_wavelet_texture_value.argtypes = [f32p, C.c_int, f32p, C.c_double, C.c_int]
        L.orc_wavelet_texture2d_value.restype = C.c_double
        L.orc_wavelet_texture2d_value.argtypes = [f32p, C.c_int, f32p, C.c_double, C.c_int]
        L.orc_perlin_texture_value.restype = C.c_double
        L.orc_perlin_texture_value.argtypes = [i32p, f32p, C.c_double, C.c_int]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_gaussian_fill.argtypes = [C.POINTER(OrcRng), f32p, C.c_size_t]
        for name in ("orc_eval2d_points", "orc_eval3d_points"):
            getattr(L, name).argtypes = [f32p, C.c_int, f32p, C.c_size_t, C.c_float, C.c_float, f32p, C.c_int]
        L.orc_eval3d_projected_points.argtypes = [f32p, C.c_int, f32p, f32p, C.c_int, C.c_size_t,
                                                  C.c_float, C.c_float, f32p, C.c_int]
        L.orc_multiband3d_lattice.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, f32p, C.c_int,
                                              f32p, f32p, C.c_int, C.c_float, f32p, C.c_int]
        L.orc_multiband3d_points.argtypes = [f32p, C.c_int, f32p, C.c_size_t, f32p, f32p, C.c_int,
                                             C.c_float, f32p, C.c_int]
        L.orc_eval2d_lattice.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_int,
                                         C.c_float, C.c_float, f32p, C.c_int]
        L.orc_eval3d_projected_lattice.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, C.c_float,
                                                   f32p, C.c_float, C.c_float, f32p, C.c_int]
        L.orc_perlin_points.argtypes = [i32p, f32p, C.c_size_t, C.c_float, f32p, C.c_int]
        L.orc_perlin_lattice.argtypes = [i32p, f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, f32p, C.c_int]
        L.orc_calculate_stats.argtypes = [f32p, C.c_size_t, C.POINTER(OrcStats)]
        L.orc_eval3d_taps.argtypes = [C.c_int, f32p, i32p]
        L.orc_logf_mismatches.restype = C.c_uint64
        L.orc_logf_mismatches.argtypes = [C.c_float, C.c_float, C.c_uint32]

    # --- rng ---
    def rng(self, seed):
        g = OrcRng()
        self.L.orc_rng_seed(C.byref(g), C.c_uint32(seed))
        return g

    def u32(self, g):
        return self.L.orc_rng_u32(C.byref(g))

    def normal(self, g):
        return self.L.orc_rng_normal(C.byref(g))

    def gaussian_fill(self, g, count):
        out = np.empty(count, np.float32)
        self.L.orc_gaussian_fill(C.byref(g), _fp(out), count)
        return out

    def perlin_perm(self, seed):
        p = np.empty(512, np.int32)
        self.L.orc_perlin_perm(C.c_uint32(seed), p.ctypes.data_as(i32p))
        return p

    # --- tiles ---
    def adjust(self, n):
        return self.L.orc_adjust_tile_size(n)

    def tile_from_field(self, R, n, dims):
        R = _f32(R)
        N = np.empty_like(R)
        fn = self.L.orc_tile3d_from_field if dims == 3 else self.L.orc_tile2d_from_field
        fn(_fp(R), _fp(N), n)
        return N

    def generate_tile(self, n, seed, dims, g=None):
        n = self.adjust(n)
        g = g or self.rng(seed)
        N = np.empty(n ** dims, np.float32)
        fn = self.L.orc_generate_tile3d if dims == 3 else self.L.orc_generate_tile2d
        fn(C.byref(g), _fp(N), n)
        return N

    def odd_offset3d(self, N, n):
        N = _f32(N).copy()
        self.L.orc_odd_offset3d(_fp(N), n)
        return N

    # --- evaluation ---
    def eval2d(self, N, n, p):
        p = _f32(p)
        return self.L.orc_eval2d(_fp(N), n, _fp(p))

    def eval3d(self, N, n, p):
        p = _f32(p)
        return self.L.orc_eval3d(_fp(N), n, _fp(p))

    def eval3d_projected(self, N, n, p, normal):
        p, normal = _f32(p), _f32(normal)
        return self.L.orc_eval3d_projected(_fp(N), n, _fp(p), _fp(normal))

    def eval2d_points(self, N, n, pts, pre=1.0, post=1.0, threads=0):
        pts = _f32(pts)
        out = np.empty(pts.size // 2, np.float32)
        self.L.orc_eval2d_points(_fp(N), n, _fp(pts), out.size, pre, post, _fp(out), threads)
        return out

    def eval3d_points(self, N, n, pts, pre=1.0, post=1.0, threads=0):
        pts = _f32(pts)
        out = np.empty(pts.size // 3, np.float32)
        self.L.orc_eval3d_points(_fp(N), n, _fp(pts), out.size, pre, post, _fp(out), threads)
        return out

    def eval3d_projected_points(self, N, n, pts, normals, pre=1.0, post=1.0, threads=0):
        pts, normals = _f32(pts), _f32(normals)
        shared = 1 if normals.size == 3 else 0
        out = np.empty(pts.size // 3, np.float32)
        self.L.orc_eval3d_projected_points(_fp(N), n, _fp(pts), _fp(normals), shared, out.size,
                                           pre, post, _fp(out), threads)
        return out

    def multiband3d_lattice(self, N, n, xs, ys, zs, band_scale, weights, post=1.0, threads=0):
        xs, ys, zs, bs, w = map(_f32, (xs, ys, zs, band_scale, weights))
        out = np.empty((zs.size, ys.size, xs.size), np.float32)
        self.L.orc_multiband3d_lattice(_fp(N), n, _fp(xs), xs.size, _fp(ys), ys.size, _fp(zs), zs.size,
                                       _fp(bs), _fp(w), bs.size, post, _fp(out), threads)
        return out

    def wmultiband_points(self, N, n, pts, s, normal, first_band, weights, threads=0):
        """Paper App. 2 WMultibandNoise(p, s, normal, firstBand, nbands, w) per point (normal None -> WNoise)."""
        N, pts, w = _f32(N), _f32(pts).reshape(-1, 3), _f32(weights)
        out = np.empty(pts.shape[0], np.float32)
        nv = None if normal is None else _f32(normal)
        self.L.orc_wmultiband_points.argtypes = [f32p, C.c_int, f32p, C.c_size_t, C.c_float, f32p, C.c_int, C.c_int, f32p,
                                                 f32p, C.c_int]
        self.L.orc_wmultiband_points.restype = None
        self.L.orc_wmultiband_points(_fp(N), n, _fp(pts), out.size, float(s), None if nv is None else _fp(nv),
                                     int(first_band), w.size, _fp(w), _fp(out), threads)
        return out

    def multiband3d_points(self, N, n, pts, band_scale, weights, post=1.0, threads=0):
        pts, bs, w = map(_f32, (pts, band_scale, weights))
        out = np.empty(pts.size // 3, np.float32)
        self.L.orc_multiband3d_points(_fp(N), n, _fp(pts), out.size, _fp(bs), _fp(w), bs.size, post,
                                      _fp(out), threads)
        return out

    def eval2d_lattice(self, N, n, xs, ys, pre=1.0, post=1.0, threads=0):
        xs, ys = _f32(xs), _f32(ys)
        out = np.empty((ys.size, xs.size), np.float32)
        self.L.orc_eval2d_lattice(_fp(N), n, _fp(xs), xs.size, _fp(ys), ys.size, pre, post, _fp(out), threads)
        return out

    def eval3d_projected_lattice(self, N, n, xs, ys, z, normal, pre=1.0, post=1.0, threads=0):
        xs, ys, normal = _f32(xs), _f32(ys), _f32(normal)
        out = np.empty((ys.size, xs.size), np.float32)
        self.L.orc_eval3d_projected_lattice(_fp(N), n, _fp(xs), xs.size, _fp(ys), ys.size, z, _fp(normal),
                                            pre, post, _fp(out), threads)
        return out

    def eval3d_taps(self, n, p):
        p = _f32(p)
        idx = np.empty(27, np.int32)
        self.L.orc_eval3d_taps(n, _fp(p), idx.ctypes.data_as(i32p))
        return idx

    # --- perlin ---
    def perlin_noise(self, perm, x, y, z):
        return self.L.orc_perlin_noise(perm.ctypes.data_as(i32p), x, y, z)

    def perlin_points(self, perm, pts, pre=1.0, threads=0):
        pts = _f32(pts)
        out = np.empty(pts.size // 3, np.float32)
        self.L.orc_perlin_points(perm.ctypes.data_as(i32p), _fp(pts), out.size, pre, _fp(out), threads)
        return out

    def perlin_lattice(self, perm, xs, ys, zs, threads=0):
        xs, ys, zs = _f32(xs), _f32(ys), _f32(zs)
        out = np.empty((zs.size, ys.size, xs.size), np.float32)
        self.L.orc_perlin_lattice(perm.ctypes.data_as(i32p), _fp(xs), xs.size, _fp(ys), ys.size,
                                  _fp(zs), zs.size, _fp(out), threads)
        return out

    # --- texture / stats / hash ---
    def wavelet_texture_value(self, N, n, p, scale, octave):
        p = _f32(p)
        return self.L.orc_wavelet_texture_value(_fp(N), n, _fp(p), scale, octave)

    def wavelet_texture2d_value(self, N2, n, p, scale, octave):
        p = _f32(p)
        return self.L.orc_wavelet_texture2d_value(_fp(N2), n, _fp(p), scale, octave)

    def perlin_texture_value(self, perm, p, scale, octave):
        p = _f32(p)
        return self.L.orc_perlin_texture_value(perm.ctypes.data_as(i32p), _fp(p), scale, octave)

    def stats(self, data):
        data = _f32(data)
        s = OrcStats()
        self.L.orc_calculate_stats(_fp(data), data.size, C.byref(s))
        return s.avg, s.var, s.min_val, s.max_val

    def fnv(self, arr):
        arr = np.ascontiguousarray(arr)
        return self.L.orc_fnv1a64(arr.ctypes.data_as(C.c_void_p), arr.nbytes)

    def logf_mismatches(self, lo, hi, step=1):
        return self.L.orc_logf_mismatches(lo, hi, step)

    def max_threads(self):
        return self.L.orc_max_threads()

    def set_threads(self, t):
        self.L.orc_set_threads(t)


class RefLib:
    """The unmodified reference (oracle/_ref/libwnref.so)."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        L = self.L = C.CDLL(REF_SO)
        L.ref_wn_create.restype = C.c_void_p
        L.ref_wn_create.argtypes = [C.c_int, C.c_uint]
        for n in ("ref_wn_destroy", "ref_wn_generate2d", "ref_wn_generate3d"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.ref_wn_tile_size.argtypes = [C.c_void_p]
        L.ref_wn_tile_count.argtypes = [C.c_void_p]
        L.ref_wn_tile_count.restype = C.c_size_t
        L.ref_wn_tile_copy.argtypes = [C.c_void_p, f32p]
        for n in ("ref_wn_eval2d_points", "ref_wn_eval3d_points"):
            getattr(L, n).argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_float, C.c_float, f32p, C.c_int]
        L.ref_wn_eval3d_projected_points.argtypes = [C.c_void_p, f32p, f32p, C.c_int, C.c_size_t,
                                                     C.c_float, C.c_float, f32p, C.c_int]
        L.ref_wn_multiband3d_lattice.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int, f32p, C.c_int,
                                                 f32p, f32p, C.c_int, C.c_float, f32p, C.c_int]
        L.ref_perlin_create.restype = C.c_void_p
        L.ref_perlin_create.argtypes = [C.c_uint]
        L.ref_perlin_destroy.argtypes = [C.c_void_p]
        L.ref_perlin_points.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_float, f32p, C.c_int]
        L.ref_perlin_noise.restype = C.c_double
        L.ref_perlin_noise.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        self.has_texture = hasattr(L, "ref_texture_values")
        if self.has_texture:
            L.ref_wavelet_texture_create.restype = C.c_void_p
            L.ref_wavelet_texture_create.argtypes = [C.c_double, C.c_int]
            if hasattr(L, 'ref_wavelet_texture2d_create'):
                L.ref_wavelet_texture2d_create.restype = C.c_void_p
                L.ref_wavelet_texture2d_create.argtypes = [C.c_double, C.c_int]
            L.ref_perlin_texture_create.restype = C.c_void_p
            L.ref_perlin_texture_create.argtypes = [C.c_double, C.c_int]
            L.ref_wavelet_texture_destroy.argtypes = [C.c_void_p]
            L.ref_perlin_texture_destroy.argtypes = [C.c_void_p]
            L.ref_texture_values.argtypes = [C.c_void_p, f32p, C.c_size_t, f32p, C.c_int]

    class Noise:
        def __init__(self, lib, n, seed):
            self.lib, self.L = lib, lib.L
            self.h = self.L.ref_wn_create(n, seed)

        def __del__(self):
            if getattr(self, "h", None):
                self.L.ref_wn_destroy(self.h)
                self.h = None

        def generate(self, dims):
            (self.L.ref_wn_generate3d if dims == 3 else self.L.ref_wn_generate2d)(self.h)
            return self

        @property
        def n(self):
            return self.L.ref_wn_tile_size(self.h)

        def tile(self):
            out = np.empty(self.L.ref_wn_tile_count(self.h), np.float32)
            self.L.ref_wn_tile_copy(self.h, _fp(out))
            return out

        def eval2d_points(self, pts, pre=1.0, post=1.0, threads=1):
            pts = _f32(pts)
            out = np.empty(pts.size // 2, np.float32)
            self.L.ref_wn_eval2d_points(self.h, _fp(pts), out.size, pre, post, _fp(out), threads)
            return out

        def eval3d_points(self, pts, pre=1.0, post=1.0, threads=1):
            pts = _f32(pts)
            out = np.empty(pts.size // 3, np.float32)
            self.L.ref_wn_eval3d_points(self.h, _fp(pts), out.size, pre, post, _fp(out), threads)
            return out

        def eval3d_projected_points(self, pts, normals, pre=1.0, post=1.0, threads=1):
            pts, normals = _f32(pts), _f32(normals)
            out = np.empty(pts.size // 3, np.float32)
            self.L.ref_wn_eval3d_projected_points(self.h, _fp(pts), _fp(normals), 1 if normals.size == 3 else 0,
                                                  out.size, pre, post, _fp(out), threads)
            return out

        def multiband3d_lattice(self, xs, ys, zs, band_scale, weights, post=1.0, threads=1):
            xs, ys, zs, bs, w = map(_f32, (xs, ys, zs, band_scale, weights))
            out = np.empty((zs.size, ys.size, xs.size), np.float32)
            self.L.ref_wn_multiband3d_lattice(self.h, _fp(xs), xs.size, _fp(ys), ys.size, _fp(zs), zs.size,
                                              _fp(bs), _fp(w), bs.size, post, _fp(out), threads)
            return out

    def noise(self, n, seed):
        return RefLib.Noise(self, n, seed)

    def perlin_points(self, seed, pts, pre=1.0, threads=1):
        h = self.L.ref_perlin_create(seed)
        pts = _f32(pts)
        out = np.empty(pts.size // 3, np.float32)
        self.L.ref_perlin_points(h, _fp(pts), out.size, pre, _fp(out), threads)
        self.L.ref_perlin_destroy(h)
        return out

    def perlin_noise(self, seed, x, y, z):
        h = self.L.ref_perlin_create(seed)
        v = self.L.ref_perlin_noise(h, x, y, z)
        self.L.ref_perlin_destroy(h)
        return v

    def texture_values(self, kind, scale, octave, pts, threads=1):
        pts = _f32(pts)
        out = np.empty(pts.size // 3, np.float32)
        if kind == "wavelet2d":
            h = self.L.ref_wavelet_texture2d_create(scale, octave)
            self.L.ref_texture_values(h, _fp(pts), out.size, _fp(out), threads)
            self.L.ref_wavelet_texture_destroy(h)
        elif kind == "wavelet":
            h = self.L.ref_wavelet_texture_create(scale, octave)
            self.L.ref_texture_values(h, _fp(pts), out.size, _fp(out), threads)
            self.L.ref_wavelet_texture_destroy(h)
        else:
            h = self.L.ref_perlin_texture_create(scale, octave)
            self.L.ref_texture_values(h, _fp(pts), out.size, _fp(out), threads)
            self.L.ref_perlin_texture_destroy(h)
        return out

    def max_threads(self):
        return self.L.ref_max_threads()
