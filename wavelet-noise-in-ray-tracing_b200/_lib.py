"""ctypes binding of libwn_b200.so (the C ABI declared in include/wn_b200.h).

The product path has no CPU fallback: if the shared library is missing this module raises at
import, and if no sm_100 GPU is present `wn_ctx_create` fails with WN_ENODEVICE.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwn_b200.so")

WN_HOST, WN_DEVICE = 0, 1
WN_TILE_DEFAULT, WN_TILE_ODD_OFFSET = 0, 1
WN_EVAL_FAST, WN_EVAL_EXACT = 0, 1
WN_PERLIN_F64, WN_PERLIN_F32 = 0, 1
WN_SHARD_SLAB, WN_SHARD_CYCLIC = 0, 1
WN_ENODEVICE = -2

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
vp = C.c_void_p


class WnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"wn_b200 error {code}: {msg}")
        self.code = code


class WnStats(C.Structure):
    _fields_ = [("avg", C.c_float), ("var", C.c_float), ("min_val", C.c_float), ("max_val", C.c_float),
                ("count_nan_inf", C.c_longlong), ("energy", C.c_float)]


# name -> (restype, argtypes).  Every symbol include/wn_b200.h declares is listed here; tests check both ways.
SIGNATURES = {
    "wn_last_error": (C.c_char_p, []),
    "wn_version": (C.c_char_p, []),
    "wn_kernel_launches": (C.c_uint64, [vp]),
    "wn_timing_last_ms": (C.c_float, [vp]),
    "wn_timing_main_kernel_enable": (C.c_int, [vp, C.c_int]),
    "wn_timing_main_kernel_collect": (C.c_int, [vp, vp, C.c_int, C.POINTER(C.c_int)]),
    "wn_debug_axis_entries": (C.c_int, [vp, vp, C.c_int, C.c_float, vp, vp]),
    "wn_debug_fold_plan": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, i32p, i32p,
                                     C.POINTER(C.c_int)]),
    "wn_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "wn_ctx_destroy": (C.c_int, [vp]),
    "wn_ctx_set_stream": (C.c_int, [vp, vp]),
    "wn_ctx_synchronize": (C.c_int, [vp]),
    "wn_ctx_device": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "wn_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "wn_host_free": (C.c_int, [vp]),
    "wn_rng_create": (C.c_int, [C.c_uint, C.POINTER(vp)]),
    "wn_rng_destroy": (C.c_int, [vp]),
    "wn_rng_fill_gaussian": (C.c_int, [vp, vp, C.c_size_t]),
    "wn_rng_discard": (C.c_int, [vp, C.c_ulonglong]),
    "wn_perlin_make_perm": (C.c_int, [C.c_uint, vp]),
    "wn_adjust_tile_size": (C.c_int, [C.c_int]),
    "wn_tile_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]),
    "wn_tile_destroy": (C.c_int, [vp]),
    "wn_tile_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "wn_tile_build_from_gaussian": (C.c_int, [vp, vp, C.c_int]),
    "wn_tile_build_seeded": (C.c_int, [vp, C.c_uint, C.POINTER(C.c_ulonglong)]),
    "wn_tile_upload": (C.c_int, [vp, vp, C.c_int]),
    "wn_tile_download": (C.c_int, [vp, vp, C.c_int]),
    "wn_tile_device_ptr": (C.c_int, [vp, C.POINTER(vp)]),
    "wn_tile_mark_built": (C.c_int, [vp]),
    "wn_eval2d_points": (C.c_int, [vp, vp, C.c_size_t, C.c_float, C.c_float, vp, C.c_int]),
    "wn_eval3d_points": (C.c_int, [vp, vp, C.c_size_t, C.c_float, C.c_float, vp, C.c_int]),
    "wn_eval3d_projected_points": (C.c_int, [vp, vp, vp, C.c_int, C.c_size_t, C.c_float, C.c_float, vp, C.c_int]),
    "wn_multiband3d_points": (C.c_int, [vp, vp, C.c_size_t, vp, vp, C.c_int, C.c_float, vp, C.c_int]),
    "wn_wmultiband_points": (C.c_int, [vp, vp, C.c_size_t, C.c_float, vp, C.c_int, C.c_int, vp, vp, C.c_int]),
    "wn_eval2d_lattice": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, C.c_float, C.c_float, vp, C.c_int]),
    "wn_multiband3d_lattice": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_float,
                                         C.c_int, vp, C.c_int]),
    "wn_eval3d_projected_grid": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_float, C.c_float, vp,
                                           C.c_int]),
    "wn_eval3d_grid": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_float, vp, C.c_int]),
    "wn_perlin_create": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "wn_perlin_destroy": (C.c_int, [vp]),
    "wn_perlin_set_precision": (C.c_int, [vp, C.c_int]),
    "wn_perlin_points_f64": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_int]),
    "wn_perlin_points": (C.c_int, [vp, vp, C.c_size_t, C.c_float, vp, C.c_int]),
    "wn_perlin_lattice": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int]),
    "wn_perlin_grid": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, vp, C.c_int]),
    "wn_wavelet_texture_values": (C.c_int, [vp, vp, C.c_size_t, C.c_double, C.c_int, vp, C.c_int]),
    "wn_wavelet_texture2d_values": (C.c_int, [vp, vp, C.c_size_t, C.c_double, C.c_int, vp, C.c_int]),
    "wn_perlin_texture_values": (C.c_int, [vp, vp, C.c_size_t, C.c_double, C.c_int, vp, C.c_int]),
    "wn_stats_compute": (C.c_int, [vp, vp, C.c_size_t, C.c_int, C.POINTER(WnStats)]),
    "wn_debug_shard_indices": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.POINTER(C.c_int)]),
    "wn_group_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "wn_group_destroy": (C.c_int, [vp]),
    "wn_group_size": (C.c_int, [vp]),
    "wn_group_ctx": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "wn_group_synchronize": (C.c_int, [vp]),
    "wn_group_tile_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]),
    "wn_group_tile_destroy": (C.c_int, [vp]),
    "wn_group_tile_build_seeded": (C.c_int, [vp, C.c_uint, C.POINTER(C.c_ulonglong)]),
    "wn_group_tile_upload": (C.c_int, [vp, vp]),
    "wn_group_tile_rank": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "wn_group_multiband3d_lattice": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_float,
                                               C.c_int, C.c_int, vp, C.POINTER(C.c_float)]),
    "wn_group_shard": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "wn_group_eval3d_projected_grid": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_float, C.c_float, vp,
                                                 C.POINTER(C.c_float)]),
    "wn_group_perlin_grid": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, vp,
                                       C.POINTER(C.c_float)]),
    "wn_group_wavelet_texture_values": (C.c_int, [vp, vp, C.c_size_t, C.c_double, C.c_int, vp, C.c_int]),
    "wn_group_perlin_texture_values": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_double, C.c_int, vp, C.c_int]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the wavelet-noise path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


def check(rc):
    if rc != 0:
        raise WnError(rc, lib.wn_last_error().decode("utf-8", "replace"))
