"""B200-native wavelet-noise path: Python host mirror of the reference's C++ interface.

Import with ``importlib.import_module("wavelet-noise-in-ray-tracing_b200")`` (the directory name is
fixed by the build contract and is not a Python identifier).

The classes keep the reference's names and argument meaning:

  WaveletNoise   <- class WaveletNoise            (WaveletNoise.h:20-59)
  PerlinNoise    <- class PerlinNoise / perlin    (experient/PerlinNoise.hpp:9-61, perlin.h:14-91)
  wavelet_texture, noise_texture <- texture.h:33-115  (value() hook + batched values())

and add batch entry points (points / lattice / grid) that the reference's per-pixel loops map onto.
All arithmetic happens in libwn_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/wn_b200.h).  There is no CPU fallback: importing this package without the built library,
or creating a Context without a B200, raises.

Buffers may be numpy arrays (host) or CUDA torch tensors (device); small parameter arrays are
always host-side.
"""
import ctypes as C
import sys

import weakref

import numpy as np

from . import _lib
from ._lib import (WN_DEVICE, WN_EVAL_EXACT, WN_EVAL_FAST, WN_HOST, WN_TILE_DEFAULT, WN_TILE_ODD_OFFSET, WnError,
                   WnStats, check, lib)

from ._lib import WN_PERLIN_F32, WN_PERLIN_F64  # noqa: E402

from ._lib import WN_SHARD_CYCLIC, WN_SHARD_SLAB  # noqa: E402

__all__ = ["WN_PERLIN_F32", "WN_PERLIN_F64", "WN_SHARD_CYCLIC", "WN_SHARD_SLAB", "DeviceGroup", "GroupTile", "Context", "WaveletNoise", "PerlinNoise", "wavelet_texture", "noise_texture", "DataStats", "WnError",
           "default_context", "WN_EVAL_FAST", "WN_EVAL_EXACT", "WN_TILE_ODD_OFFSET", "pinned_empty"]


def _is_torch(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _in(x, dtype=np.float32):
    """-> (address, space, keepalive) for an input buffer."""
    if _is_torch(x):
        import torch
        want = {np.float32: torch.float32, np.int32: torch.int32}[dtype]
        if x.dtype != want or not x.is_contiguous():
            x = x.to(want).contiguous()
        return x.data_ptr(), (WN_DEVICE if x.is_cuda else WN_HOST), x
    a = np.ascontiguousarray(x, dtype=dtype)
    return a.ctypes.data, WN_HOST, a


def _host(x, dtype=np.float32):
    """small parameter arrays: always host."""
    if _is_torch(x):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def _out(out, shape, space, like=None):
    """allocate / validate an output buffer in `space`; returns (address, object)."""
    count = int(np.prod(shape))
    if out is None:
        if space == WN_DEVICE:
            import torch
            out = torch.empty(tuple(shape), dtype=torch.float32, device=like.device if like is not None else "cuda")
        else:
            out = np.empty(shape, np.float32)
    if _is_torch(out):
        import torch
        if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != count:
            raise ValueError("out must be a contiguous float32 tensor with %d elements" % count)
        if (WN_DEVICE if out.is_cuda else WN_HOST) != space:
            raise ValueError("input and output buffers must live in the same space (both host or both device)")
        return out.data_ptr(), out
    if out.dtype != np.float32 or not out.flags.c_contiguous or out.size != count:
        raise ValueError("out must be a C-contiguous float32 array with %d elements" % count)
    if space != WN_HOST:
        raise ValueError("device inputs need a CUDA tensor as out")
    return out.ctypes.data, out


def pinned_empty(shape, dtype=np.float32):
    """Page-locked host array (numpy view over cudaMallocHost memory) for full-speed copies."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check(lib.wn_host_alloc(nbytes, C.byref(p)))
    buf = (C.c_char * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[arr.ctypes.data] = p.value
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        check(lib.wn_host_free(p))


class Context:
    """One GPU + one stream (wn_ctx).  device=-1 uses the current CUDA device."""

    def __init__(self, device=-1):
        h = C.c_void_p()
        check(lib.wn_ctx_create(int(device), C.byref(h)))
        self.h = h
        dev, sms = C.c_int(), C.c_int()
        check(lib.wn_ctx_device(self.h, C.byref(dev), C.byref(sms)))
        self.device, self.sm_count = dev.value, sms.value
        self._children = weakref.WeakSet()      # tiles / Perlin tables created on this context

    def close(self):
        """Destroys the context.  Objects created on it release their device state first (a tile or Perlin handle
        must not outlive its wn_ctx: wn_tile_destroy dereferences it)."""
        if getattr(self, "h", None):
            for child in list(getattr(self, "_children", ())):
                child._release()
            lib.wn_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        """Run on a caller's stream (an int cudaStream_t; 0/None restores the library's own stream)."""
        check(lib.wn_ctx_set_stream(self.h, C.c_void_p(cuda_stream_handle or None)))

    def use_torch_stream(self):
        """Enqueue on torch's current stream so device tensors need no extra synchronisation."""
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        self.set_stream(s if s else 1)          # 1 == cudaStreamLegacy (torch's default stream is handle 0)

    def synchronize(self):
        check(lib.wn_ctx_synchronize(self.h))

    @property
    def kernel_launches(self):
        return int(lib.wn_kernel_launches(self.h))

    @property
    def last_kernel_ms(self):
        return float(lib.wn_timing_last_ms(self.h))

    def time_main_kernel(self, on=True):
        """Record one CUDA-event pair around the main kernel of every device-resident FAST lattice call (roofline)."""
        check(lib.wn_timing_main_kernel_enable(self.h, 1 if on else 0))

    def main_kernel_ms(self):
        """Durations (ms) of the main kernels recorded since the last call; synchronises the stream."""
        n = C.c_int()
        buf = np.zeros(4096, np.float32)
        check(lib.wn_timing_main_kernel_collect(self.h, C.c_void_p(buf.ctypes.data), buf.size, C.byref(n)))
        return buf[:min(n.value, buf.size)].copy()

    def axis_entries(self, coords, band_scale=1.0):
        """(weights (n, 3) float32, first tap cell (n,) int32) of the FAST lattice kernels' axis table (diagnostics)."""
        coords = np.ascontiguousarray(coords, np.float32)
        w = np.empty((coords.size, 3), np.float32)
        first = np.empty(coords.size, np.int32)
        check(lib.wn_debug_axis_entries(self.h, C.c_void_p(coords.ctypes.data), coords.size, float(band_scale),
                                        C.c_void_p(w.ctypes.data), C.c_void_p(first.ctypes.data)))
        return w, first

    def stats(self, data):
        ptr, space, keep = _in(data)
        n = keep.numel() if _is_torch(keep) else keep.size
        s = WnStats()
        check(lib.wn_stats_compute(self.h, C.c_void_p(ptr), n, space, C.byref(s)))
        return DataStats(s.avg, s.var, s.min_val, s.max_val)


_DEFAULT = None


def fold_plan(xs, ys, zs, band_scale, tile_n):
    """Host-only diagnostic (no GPU): which bands a FAST multiband3D_lattice call on these axes folds at its top level.
    Returns (folded flags per band in the given order, (Lx, Ly, Lz))."""
    xs, ys, zs = _host(xs), _host(ys), _host(zs)
    bs = _host(band_scale)
    folded = np.zeros(max(bs.size, 1), np.int32)
    block = np.zeros(3, np.int32)
    n = C.c_int(0)
    _lib.check(_lib.lib.wn_debug_fold_plan(xs.ctypes.data, xs.size, ys.ctypes.data, ys.size, zs.ctypes.data, zs.size,
                                           bs.ctypes.data, bs.size, int(tile_n),
                                           folded.ctypes.data_as(_lib.i32p), block.ctypes.data_as(_lib.i32p), C.byref(n)))
    return folded[: bs.size].astype(bool), tuple(int(v) for v in block)


def default_context():
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Context(-1)
    return _DEFAULT


class DataStats:
    """struct DataStats (WaveletNoise.h:11-18); count_nan_inf / energy are never set by the reference."""

    def __init__(self, avg=0.0, var=0.0, min_val=np.finfo(np.float32).max, max_val=-np.finfo(np.float32).max):
        self.avg, self.var, self.min_val, self.max_val = avg, var, min_val, max_val
        self.count_nan_inf = 0
        self.energy = 0.0

    def __repr__(self):
        return f"DataStats(avg={self.avg}, var={self.var}, min_val={self.min_val}, max_val={self.max_val})"


class WaveletNoise:
    """Drop-in for the reference's class (WaveletNoise.h:20-59) with GPU-resident tile.

    generateNoiseTile2D/3D draw the Gaussian field from the reference's generator sequence
    (std::mt19937(seed) + std::normal_distribution<float>) -- on the GPU for the first tile of an object
    (wn_tile_build_seeded, bit-identical; the host generator is then advanced by the same number of raw draws),
    from the host generator objects when the stream continues into a second tile -- run the filter passes on the GPU and keep the
    coefficients on the device; the host copy returned by getNoiseCoefficients() is downloaded lazily.
    """

    def __init__(self, tileSize, seed=0, ctx=None, flags=WN_TILE_DEFAULT):
        self.ctx = ctx or default_context()
        self.tileSizeN = int(tileSize)
        if self.tileSizeN % 2 != 0:                        # WaveletNoise.cpp:22-25
            self.tileSizeN += 1
            print(f"Warning: Tile size adjusted to {self.tileSizeN} (must be even)", file=sys.stderr)
        self.randomSeed = int(seed) & 0xFFFFFFFF
        self.flags = flags
        r = C.c_void_p()
        check(lib.wn_rng_create(self.randomSeed, C.byref(r)))
        self._rng = r
        self._tile = None
        self._dims = 0
        self._host = None
        self._fresh = True              # the host generator has not been used yet (the first fill may run on the GPU)
        self.ctx._children.add(self)

    def __del__(self):
        try:
            self._release()
            if getattr(self, "_rng", None):
                lib.wn_rng_destroy(self._rng)
                self._rng = None
        except Exception:
            pass

    def _release(self):
        self._drop_tile()

    def _drop_tile(self):
        if getattr(self, "_tile", None):
            if getattr(self.ctx, "h", None):     # a closed context has already released this tile
                lib.wn_tile_destroy(self._tile)
            self._tile = None

    def _new_tile(self, dims):
        if self._tile is not None and self._dims == dims:      # same object, same shape: rebuild in place
            self._host = None
            return
        self._drop_tile()
        t = C.c_void_p()
        check(lib.wn_tile_create(self.ctx.h, self.tileSizeN, dims, self.flags, C.byref(t)))
        self._tile, self._dims, self._host = t, dims, None

    # ---- construction -------------------------------------------------------------------------
    def gaussian_field(self, count):
        """Next `count` variates of this object's generator (what the fill loops cpp:74-77/146-147 draw)."""
        R = np.empty(count, np.float32)
        check(lib.wn_rng_fill_gaussian(self._rng, C.c_void_p(R.ctypes.data), count))
        self._fresh = False
        return R

    def _generate(self, dims, field=None):
        self._new_tile(dims)
        count = self.tileSizeN ** dims
        if field is None and self._fresh:
            # First tile of a fresh object: the whole fill runs on the GPU (wn_tile_build_seeded: MT19937 + polar
            # method + logf on the device, bit-identical to the host objects); the host generator is then advanced
            # by the same number of raw draws, so a later generate* continues the stream exactly like the reference.
            draws = C.c_ulonglong(0)
            check(lib.wn_tile_build_seeded(self._tile, self.randomSeed, C.byref(draws)))
            check(lib.wn_rng_discard(self._rng, draws.value))
            self._fresh = False
            return
        if field is None:
            field = self.gaussian_field(count)
        ptr, space, keep = _in(field)
        if (keep.numel() if _is_torch(keep) else keep.size) != count:
            raise ValueError(f"Gaussian field must have {count} elements")
        check(lib.wn_tile_build_from_gaussian(self._tile, C.c_void_p(ptr), space))

    def generateNoiseTile2D(self, field=None):
        self._generate(2, field)

    def generateNoiseTile3D(self, field=None):
        self._generate(3, field)

    def generate_seeded(self, dims, seed=None):
        """Build from a seed through wn_tile_build_seeded (fresh generator, like a new object)."""
        self._new_tile(dims)
        check(lib.wn_tile_build_seeded(self._tile, self.randomSeed if seed is None else int(seed), None))

    # ---- tile cache (SURVEY.md section 5 "checkpoint / resume": the tile is a pure function of n, dims, seed, flags) ----
    def cache_path(self, dims, cache_dir):
        import os
        return os.path.join(cache_dir, f"wn_tile_n{self.tileSizeN}_d{dims}_seed{self.randomSeed}_f{self.flags}.raw")

    def generate_cached(self, dims, cache_dir):
        """First tile of a fresh object through a file cache: float32 little-endian `.raw` in memory order (the format
        experient/analyze.py:6-21 reads), named after (n, dims, seed, flags).  A hit uploads the coefficients and
        advances the host generator past the fill it skipped -- the object is indistinguishable from one that generated
        the tile; a miss generates on the GPU and writes the file.  Returns True on a hit."""
        import os
        if not self._fresh:
            raise WnError(-5, "generate_cached: the generator of this object has already been used")
        path = self.cache_path(dims, cache_dir)
        count = self.tileSizeN ** dims
        if os.path.exists(path) and os.path.getsize(path) == 4 * count + 8:
            blob = np.fromfile(path, dtype="<u1")
            draws = int(blob[-8:].view("<u8")[0])
            self.upload(blob[:-8].view("<f4"), dims)
            check(lib.wn_rng_discard(self._rng, draws))
            self._fresh = False
            return True
        self._new_tile(dims)
        draws = C.c_ulonglong(0)
        check(lib.wn_tile_build_seeded(self._tile, self.randomSeed, C.byref(draws)))
        check(lib.wn_rng_discard(self._rng, draws.value))
        self._fresh = False
        os.makedirs(cache_dir, exist_ok=True)
        tmp = path + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:                      # coefficients, then the raw draws the fill consumed (uint64)
            f.write(np.ascontiguousarray(self.getNoiseCoefficients(), "<f4").tobytes())
            f.write(np.array([draws.value], "<u8").tobytes())
        os.replace(tmp, path)
        return False

    def upload(self, coefficients, dims):
        """Adopt finished coefficients (n^dims floats)."""
        self._new_tile(dims)
        ptr, space, keep = _in(coefficients)
        check(lib.wn_tile_upload(self._tile, C.c_void_p(ptr), space))

    # ---- accessors ----------------------------------------------------------------------------
    def getTileSize(self):
        return self.tileSizeN

    def getNoiseCoefficients(self):
        if self._tile is None:
            return np.empty(0, np.float32)
        if self._host is None:
            out = np.empty(self.tileSizeN ** self._dims, np.float32)
            check(lib.wn_tile_download(self._tile, C.c_void_p(out.ctypes.data), WN_HOST))
            self._host = out
        return self._host

    def device_ptr(self):
        p = C.c_void_p()
        check(lib.wn_tile_device_ptr(self._tile, C.byref(p)))
        return p.value

    def device_tensor(self):
        """The device tile as a torch tensor view (for torch.distributed.broadcast of the replica)."""
        import torch
        count = self.tileSizeN ** self._dims

        class _View:
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (self.device_ptr(), False),
                                        "version": 3, "strides": None}
        return torch.as_tensor(_View(), device=f"cuda:{self.ctx.device}")

    def allocate(self, dims):
        """Create an un-built tile (a replica that will be filled through device_tensor(), e.g. by a broadcast)."""
        self._new_tile(dims)

    def mark_built(self):
        check(lib.wn_tile_mark_built(self._tile))
        self._host = None

    def calculateStats(self, data, name):
        """WaveletNoise.cpp:268-288 (GPU reduction); prints the same line."""
        st = self.ctx.stats(data)
        print(f"{name} stats: avg={st.avg:g}, var={st.var:g}, stddev={np.sqrt(np.float32(st.var)):g}")
        return st

    # ---- scalar evaluation (reference signatures) -----------------------------------------------
    def evaluate2D(self, p):
        if self._tile is None:
            return 0.0                                       # empty tile -> 0.0f (cpp:112)
        return float(self.evaluate2D_points(np.asarray(p, np.float32).reshape(1, 2))[0])

    def evaluate3D(self, p):
        if self._tile is None:
            return 0.0
        return float(self.evaluate3D_points(np.asarray(p, np.float32).reshape(1, 3))[0])

    def evaluate3DProjected(self, p, normal):
        if self._tile is None:
            return 0.0
        return float(self.evaluate3DProjected_points(np.asarray(p, np.float32).reshape(1, 3), normal)[0])

    # ---- batch evaluation -------------------------------------------------------------------------
    def _need(self):
        if self._tile is None:
            raise WnError(-5, "the tile has not been generated")

    def evaluate2D_points(self, pts, pre_scale=1.0, post_scale=1.0, out=None):
        self._need()
        ptr, space, keep = _in(pts)
        count = (keep.numel() if _is_torch(keep) else keep.size) // 2
        optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
        check(lib.wn_eval2d_points(self._tile, C.c_void_p(ptr), count, pre_scale, post_scale, C.c_void_p(optr), space))
        return out

    def evaluate3D_points(self, pts, pre_scale=1.0, post_scale=1.0, out=None):
        self._need()
        ptr, space, keep = _in(pts)
        count = (keep.numel() if _is_torch(keep) else keep.size) // 3
        optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
        check(lib.wn_eval3d_points(self._tile, C.c_void_p(ptr), count, pre_scale, post_scale, C.c_void_p(optr), space))
        return out

    def evaluate3DProjected_points(self, pts, normals, pre_scale=1.0, post_scale=1.0, out=None):
        self._need()
        ptr, space, keep = _in(pts)
        count = (keep.numel() if _is_torch(keep) else keep.size) // 3
        nsize = normals.numel() if _is_torch(normals) else np.asarray(normals).size
        if nsize == 3:
            nh = _host(normals)
            nptr, shared, nkeep = nh.ctypes.data, 1, nh
        else:
            nptr, nspace, nkeep = _in(normals)
            shared = 0
            if nspace != space or nsize != 3 * count:
                raise ValueError("per-point normals must match the points (same space, 3 floats each)")
        optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
        check(lib.wn_eval3d_projected_points(self._tile, C.c_void_p(ptr), C.c_void_p(nptr), shared, count, pre_scale,
                                             post_scale, C.c_void_p(optr), space))
        return out

    def multiband3D_points(self, pts, band_scale, weights, post_scale=1.0, out=None):
        self._need()
        ptr, space, keep = _in(pts)
        count = (keep.numel() if _is_torch(keep) else keep.size) // 3
        bs, w = _host(band_scale), _host(weights)
        optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
        check(lib.wn_multiband3d_points(self._tile, C.c_void_p(ptr), count, C.c_void_p(bs.ctypes.data),
                                        C.c_void_p(w.ctypes.data), bs.size, post_scale, C.c_void_p(optr), space))
        return out

    def WMultibandNoise(self, pts, s, normal, firstBand, w, out=None):
        """Cook & DeRose App. 2 WMultibandNoise(p, s, normal, firstBand, nbands, w) for a batch of points (normal None or
        three floats shared by the batch; nbands = len(w))."""
        self._need()
        ptr, space, keep = _in(pts)
        count = (keep.numel() if _is_torch(keep) else keep.size) // 3
        optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
        w = _host(w)
        nv = None if normal is None else _host(normal)
        check(lib.wn_wmultiband_points(self._tile, C.c_void_p(ptr), count, float(s),
                                       None if nv is None else C.c_void_p(nv.ctypes.data), int(firstBand), w.size,
                                       C.c_void_p(w.ctypes.data), C.c_void_p(optr), space))
        return out

    def evaluate2D_lattice(self, xs, ys, pre_scale=1.0, post_scale=1.0, out=None, device_out=False):
        self._need()
        xs, ys = _host(xs), _host(ys)
        space = WN_DEVICE if (device_out or (out is not None and _is_torch(out) and out.is_cuda)) else WN_HOST
        optr, out = _out(out, (ys.size, xs.size), space)
        check(lib.wn_eval2d_lattice(self._tile, C.c_void_p(xs.ctypes.data), xs.size, C.c_void_p(ys.ctypes.data), ys.size,
                                    pre_scale, post_scale, C.c_void_p(optr), space))
        return out

    def multiband3D_lattice(self, xs, ys, zs, band_scale, weights, post_scale=1.0, mode=WN_EVAL_FAST, out=None,
                            device_out=False):
        """out[k, j, i] = post * sum_b w[b] * evaluate3D((xs[i], ys[j], zs[k]) * band_scale[b])."""
        self._need()
        xs, ys, zs, bs, w = _host(xs), _host(ys), _host(zs), _host(band_scale), _host(weights)
        if bs.size != w.size:
            raise ValueError("band_scale and weights must have the same length")
        space = WN_DEVICE if (device_out or (out is not None and _is_torch(out) and out.is_cuda)) else WN_HOST
        optr, out = _out(out, (zs.size, ys.size, xs.size), space)
        check(lib.wn_multiband3d_lattice(self._tile, C.c_void_p(xs.ctypes.data), xs.size, C.c_void_p(ys.ctypes.data),
                                         ys.size, C.c_void_p(zs.ctypes.data), zs.size, C.c_void_p(bs.ctypes.data),
                                         C.c_void_p(w.ctypes.data), bs.size, post_scale, mode, C.c_void_p(optr), space))
        return out

    def evaluate3D_grid(self, origin, e1, us, e2, vs, pre_scale=1.0, post_scale=1.0, out=None, device_out=False):
        """out[j, i] = evaluate3D((origin + us[i]*e1 + vs[j]*e2) * pre) * post."""
        self._need()
        o, a, b, us, vs = _host(origin), _host(e1), _host(e2), _host(us), _host(vs)
        space = WN_DEVICE if (device_out or (out is not None and _is_torch(out) and out.is_cuda)) else WN_HOST
        optr, out = _out(out, (vs.size, us.size), space)
        check(lib.wn_eval3d_grid(self._tile, C.c_void_p(o.ctypes.data), C.c_void_p(a.ctypes.data),
                                 C.c_void_p(us.ctypes.data), us.size, C.c_void_p(b.ctypes.data),
                                 C.c_void_p(vs.ctypes.data), vs.size, pre_scale, post_scale, C.c_void_p(optr), space))
        return out

    def evaluate3DProjected_grid(self, origin, e1, us, e2, vs, normal, pre_scale=1.0, post_scale=1.0, out=None,
                                 device_out=False):
        self._need()
        o, a, b, us, vs, nr = _host(origin), _host(e1), _host(e2), _host(us), _host(vs), _host(normal)
        space = WN_DEVICE if (device_out or (out is not None and _is_torch(out) and out.is_cuda)) else WN_HOST
        optr, out = _out(out, (vs.size, us.size), space)
        check(lib.wn_eval3d_projected_grid(self._tile, C.c_void_p(o.ctypes.data), C.c_void_p(a.ctypes.data),
                                           C.c_void_p(us.ctypes.data), us.size, C.c_void_p(b.ctypes.data),
                                           C.c_void_p(vs.ctypes.data), vs.size, C.c_void_p(nr.ctypes.data), pre_scale,
                                           post_scale, C.c_void_p(optr), space))
        return out

    def texture_values(self, pts, scale, octave, out=None):
        """grey value of wavelet_texture::value for each hit point (texture.h:67-107)."""
        self._need()
        ptr, space, keep = _in(pts)
        count = (keep.numel() if _is_torch(keep) else keep.size) // 3
        optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
        check(lib.wn_wavelet_texture_values(self._tile, C.c_void_p(ptr), count, float(scale), int(octave),
                                            C.c_void_p(optr), space))
        return out


def _texture2d_values(self, pts, scale, octave, out=None):
    """grey value of wavelet_texture::value's 2D branch for each hit point (texture.h:86-99)."""
    self._need()
    ptr, space, keep = _in(pts)
    count = (keep.numel() if _is_torch(keep) else keep.size) // 3
    optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
    check(lib.wn_wavelet_texture2d_values(self._tile, C.c_void_p(ptr), count, float(scale), int(octave),
                                          C.c_void_p(optr), space))
    return out


WaveletNoise.texture2d_values = _texture2d_values


class PerlinNoise:
    """Drop-in for PerlinNoise (experient/PerlinNoise.hpp:9-61) == perlin (perlin.h:14-91)."""

    default_seed = 5489            # std::mt19937::default_seed

    def __init__(self, seed=default_seed, ctx=None):
        self.ctx = ctx or default_context()
        self.p = np.empty(512, np.int32)
        check(lib.wn_perlin_make_perm(int(seed) & 0xFFFFFFFF, C.c_void_p(self.p.ctypes.data)))
        h = C.c_void_p()
        check(lib.wn_perlin_create(self.ctx.h, C.c_void_p(self.p.ctypes.data), C.byref(h)))
        self.h = h
        self.ctx._children.add(self)

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _release(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):
                lib.wn_perlin_destroy(self.h)
            self.h = None

    def set_precision(self, precision):
        """WN_PERLIN_F64 (default, bit-identical to the reference) or WN_PERLIN_F32 (fast mode, <= 1e-5 * range) for
        the float batch calls noise_points / noise_lattice / noise_grid."""
        check(lib.wn_perlin_set_precision(self.h, int(precision)))

    def noise(self, x, y=None, z=0.0):
        """noise(x,y,z) / noise(x,y) / noise(point3): double coordinates in, double noise out, like the reference's
        signature (experient/PerlinNoise.hpp:36) -- nothing is narrowed to float."""
        if y is None:
            x, y, z = x
        return float(self.noise_points_f64(np.array([[x, y, z]], np.float64))[0])

    def noise_points_f64(self, pts):
        """Host float64 points (n, 3) -> float64 noise (wn_perlin_points_f64)."""
        pts = np.ascontiguousarray(pts, np.float64)
        out = np.empty(pts.size // 3, np.float64)
        check(lib.wn_perlin_points_f64(self.h, C.c_void_p(pts.ctypes.data), pts.size // 3, C.c_void_p(out.ctypes.data),
                                       WN_HOST))
        return out

    def noise_points(self, pts, pre_scale=1.0, out=None):
        ptr, space, keep = _in(pts)
        count = (keep.numel() if _is_torch(keep) else keep.size) // 3
        optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
        check(lib.wn_perlin_points(self.h, C.c_void_p(ptr), count, pre_scale, C.c_void_p(optr), space))
        return out

    def noise_lattice(self, xs, ys, zs, out=None, device_out=False):
        xs, ys, zs = _host(xs), _host(ys), _host(zs)
        space = WN_DEVICE if (device_out or (out is not None and _is_torch(out) and out.is_cuda)) else WN_HOST
        optr, out = _out(out, (zs.size, ys.size, xs.size), space)
        check(lib.wn_perlin_lattice(self.h, C.c_void_p(xs.ctypes.data), xs.size, C.c_void_p(ys.ctypes.data), ys.size,
                                    C.c_void_p(zs.ctypes.data), zs.size, C.c_void_p(optr), space))
        return out

    def noise_grid(self, origin, e1, us, e2, vs, pre_scale=1.0, out=None, device_out=False):
        o, a, b, us, vs = _host(origin), _host(e1), _host(e2), _host(us), _host(vs)
        space = WN_DEVICE if (device_out or (out is not None and _is_torch(out) and out.is_cuda)) else WN_HOST
        optr, out = _out(out, (vs.size, us.size), space)
        check(lib.wn_perlin_grid(self.h, C.c_void_p(o.ctypes.data), C.c_void_p(a.ctypes.data), C.c_void_p(us.ctypes.data),
                                 us.size, C.c_void_p(b.ctypes.data), C.c_void_p(vs.ctypes.data), vs.size, pre_scale,
                                 C.c_void_p(optr), space))
        return out

    def texture_values(self, pts, scale, octave, out=None):
        ptr, space, keep = _in(pts)
        count = (keep.numel() if _is_torch(keep) else keep.size) // 3
        optr, out = _out(out, (count,), space, keep if _is_torch(keep) else None)
        check(lib.wn_perlin_texture_values(self.h, C.c_void_p(ptr), count, float(scale), int(octave), C.c_void_p(optr),
                                           space))
        return out


class DeviceGroup:
    """Single-process multi-GPU group (wn_group): one context per GPU, tile replicated with one NCCL broadcast, sharded
    evaluation calls that enqueue on every GPU before waiting for any.  ngpus <= 0: every visible GPU."""

    def __init__(self, ngpus=0, devices=None):
        h = C.c_void_p()
        dev = None
        if devices is not None:
            dev = (C.c_int * len(devices))(*devices)
            ngpus = len(devices)
        check(lib.wn_group_create(int(ngpus), dev, C.byref(h)))
        self.h = h
        self.size = lib.wn_group_size(self.h)
        self._tiles = weakref.WeakSet()

    def close(self):
        if getattr(self, "h", None):
            for t in list(self._tiles):
                t.close()
            lib.wn_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(lib.wn_group_synchronize(self.h))

    def tile(self, n, dims=3, seed=0, flags=WN_TILE_DEFAULT, coefficients=None):
        """A tile replicated on every GPU of the group: built from `seed` on rank 0 (GPU fill + filters) or uploaded
        from `coefficients`, then broadcast."""
        return GroupTile(self, n, dims, seed, flags, coefficients)


class GroupTile:
    def __init__(self, group, n, dims, seed, flags, coefficients):
        self.group = group
        h = C.c_void_p()
        check(lib.wn_group_tile_create(group.h, int(n), int(dims), int(flags), C.byref(h)))
        self.h = h
        group._tiles.add(self)
        if coefficients is None:
            check(lib.wn_group_tile_build_seeded(self.h, int(seed) & 0xFFFFFFFF, None))
        else:
            c = np.ascontiguousarray(coefficients, np.float32)
            check(lib.wn_group_tile_upload(self.h, C.c_void_p(c.ctypes.data)))

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.group, "h", None):
                lib.wn_group_tile_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def multiband3D_lattice(self, xs, ys, zs, band_scale, weights, post_scale=1.0, mode=WN_EVAL_FAST,
                            sharding=_lib.WN_SHARD_CYCLIC, gather=True):
        """Returns (volume or None, gpu_ms): the volume gathered on the host when gather is true (the shards stay on the
        devices either way), gpu_ms = max over the GPUs of the CUDA-event time."""
        xs, ys, zs = _host(xs), _host(ys), _host(zs)
        bs, w = _host(band_scale), _host(weights)
        if bs.size != w.size:
            raise ValueError("band_scale and weights differ in length")
        out = np.empty((zs.size, ys.size, xs.size), np.float32) if gather else None
        ms = C.c_float()
        check(lib.wn_group_multiband3d_lattice(self.h, C.c_void_p(xs.ctypes.data), xs.size, C.c_void_p(ys.ctypes.data),
                                               ys.size, C.c_void_p(zs.ctypes.data), zs.size, C.c_void_p(bs.ctypes.data),
                                               C.c_void_p(w.ctypes.data), bs.size, float(post_scale), int(mode),
                                               int(sharding), C.c_void_p(out.ctypes.data) if gather else None,
                                               C.byref(ms)))
        return out, ms.value

    def evaluate3DProjected_grid(self, origin, e1, us, e2, vs, normal, pre_scale=1.0, post_scale=1.0, gather=True):
        o, a, b, us, vs, nv = _host(origin), _host(e1), _host(e2), _host(us), _host(vs), _host(normal)
        out = np.empty((vs.size, us.size), np.float32) if gather else None
        ms = C.c_float()
        check(lib.wn_group_eval3d_projected_grid(self.h, C.c_void_p(o.ctypes.data), C.c_void_p(a.ctypes.data),
                                                 C.c_void_p(us.ctypes.data), us.size, C.c_void_p(b.ctypes.data),
                                                 C.c_void_p(vs.ctypes.data), vs.size, C.c_void_p(nv.ctypes.data),
                                                 float(pre_scale), float(post_scale),
                                                 C.c_void_p(out.ctypes.data) if gather else None, C.byref(ms)))
        return out, ms.value

    def texture_values(self, pts, scale, octave):
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.empty(pts.size // 3, np.float32)
        check(lib.wn_group_wavelet_texture_values(self.h, C.c_void_p(pts.ctypes.data), pts.size // 3, float(scale),
                                                  int(octave), C.c_void_p(out.ctypes.data), 1))
        return out


class noise_texture:
    """texture.h:33-49: Perlin texture, default-seeded perlin, value() = 0.5*(1+noise(p*scale*2^octave))."""

    def __init__(self, scale, octave=4, ctx=None):
        self.noise = PerlinNoise(PerlinNoise.default_seed, ctx)
        self.scale, self.octave_level = float(scale), int(octave)

    def value(self, u, v, p):
        g = float(self.values(np.asarray(p, np.float32).reshape(1, 3))[0])
        return (g, g, g)

    def values(self, pts, out=None):
        return self.noise.texture_values(pts, self.scale, self.octave_level, out)


class wavelet_texture:
    """texture.h:51-115: builds a 2D and (use_3d) a 3D tile, n=128 seed 12345, in its constructor."""

    TILE_SIZE = 128
    SEED = 12345

    def __init__(self, scale=1.0, octave=4, use_3d=True, ctx=None):
        self.scale, self.octave_level, self.use_3d_noise = float(scale), int(octave), bool(use_3d)
        self.noise_2d = WaveletNoise(self.TILE_SIZE, self.SEED, ctx)
        self.noise_2d.generateNoiseTile2D()
        self.noise_3d = None
        if use_3d:
            self.noise_3d = WaveletNoise(self.TILE_SIZE, self.SEED, ctx)
            self.noise_3d.generateNoiseTile3D()

    def value(self, u, v, p):
        g = float(self.values(np.asarray(p, np.float32).reshape(1, 3))[0])
        return (g, g, g)

    def values(self, pts, out=None):
        if self.use_3d_noise and self.noise_3d is not None:
            return self.noise_3d.texture_values(pts, self.scale, self.octave_level, out)
        # 2D branch (texture.h:86-99), never taken by the reference renderer
        return self.noise_2d.texture2d_values(pts, self.scale, self.octave_level, out)
