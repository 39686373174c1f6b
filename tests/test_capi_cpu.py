"""CPU-side checks of the C-ABI boundary: the library loads, exports exactly what include/wn_b200.h
declares, and refuses to work without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import wnpkg

ROOT = wnpkg.ROOT
HEADER = os.path.join(ROOT, "include", "wn_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = wnpkg.load_sub("_lib")
    names = header_functions()
    assert len(names) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (wn_[a-z0-9_]+)$", out, flags=re.M))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in wn_b200.h but not exported: {missing}"
    extra = sorted(exported - set(names))
    assert not extra, f"exported but not declared in wn_b200.h: {extra}"
    # the Python binding covers the same set
    assert sorted(lib.SIGNATURES) == names


def test_no_oracle_in_product_path():
    """The product package must never import or link the oracle."""
    pkg_dir = os.path.join(ROOT, wnpkg.PKG)
    for dirpath, _, files in os.walk(pkg_dir):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh", ".sh")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "wn_oracle" not in text and "libwnref" not in text and "oracle_lib" not in text, f
    lib = wnpkg.load_sub("_lib")
    ldd = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "wnref" not in ldd


def test_pure_host_helpers_and_argument_errors():
    lib = wnpkg.load_sub("_lib").lib
    assert lib.wn_adjust_tile_size(127) == 128 and lib.wn_adjust_tile_size(128) == 128
    assert b"sm_100a" in lib.wn_version()
    perm = (C.c_int32 * 512)()
    assert lib.wn_perlin_make_perm(12345, perm) == 0
    assert list(perm[:8]) == [48, 218, 61, 238, 202, 125, 107, 148]
    assert list(perm[256:264]) == list(perm[:8])
    rng = C.c_void_p()
    assert lib.wn_rng_create(12345, C.byref(rng)) == 0
    buf = (C.c_float * 8)()
    assert lib.wn_rng_fill_gaussian(rng, buf, 8) == 0
    got = [C.c_uint32.from_buffer(C.c_float(v)).value for v in buf]
    assert got == [0xbf492c2c, 0xbec80f26, 0x3f0763bd, 0xbef51152, 0x3f96f80f, 0x401f5e2b, 0x3f0491d6, 0x3dde10bb]
    lib.wn_rng_destroy(rng)
    assert lib.wn_ctx_create(0, None) == -1 and b"NULL" in lib.wn_last_error()


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    wn = wnpkg.load()
    with pytest.raises(wn.WnError) as e:
        wn.Context()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_fold_plan_host_logic(monkeypatch):
    """The fold decision of the fast lattice path is pure host code (period detection on the axis tables + cost model):
    config 3, its contiguous and block-cyclic shards, a non-commensurate lattice, a non-power-of-two tile."""
    import numpy as np
    wn = wnpkg.load()
    sh = wnpkg.load_sub("sharding")
    for k in ("WN_FOLD_BUDGET", "WN_FOLD_NEST"):
        monkeypatch.delenv(k, raising=False)
    ax = sh.lattice_axes_config3(1024)
    scale, _, _ = sh.config3_bands(4, 8)
    # whole volume: bands 5..8 repeat (periods 512/256/128/64 samples), band 4's period is the lattice itself.  Bands 6..8
    # fold onto a 256^3 block; band 5 stays direct because the replica kernel evaluates it once per four samples (the x
    # and y halves of the lattice are a whole number of its periods apart) and a 512^3 block would cost an HBM round trip
    folded, block = wn.fold_plan(ax, ax, ax, scale, 128)
    assert list(folded) == [False, False, True, True, True] and block == (256, 256, 256)
    # the decision follows the bands, not their order in the call
    perm = [3, 0, 4, 2, 1]
    folded_p, block_p = wn.fold_plan(ax, ax, ax, scale[perm], 128)
    assert list(folded_p) == [True, False, True, True, False] and block_p == block
    # contiguous 128-slice slab (1/8 of the volume): no whole z period of band 6 -> the block spans the slab in z
    folded, block = wn.fold_plan(ax, ax, ax[:128], scale, 128)
    assert list(folded) == [False, False, True, True, True] and block == (256, 256, 128)
    # block-cyclic shard of rank 3 of 8: slices congruent modulo 256 stay together -> z period 32 for band 6
    zs = ax[sh.cyclic_slab_indices(1024, 3, 8)]
    folded, block = wn.fold_plan(ax, ax, zs, scale, 128)
    assert list(folded) == [False, False, True, True, True] and block == (256, 256, 32)
    # x extent 1536: the halves are 1.5 periods of band 5 apart, so its replicas do not coincide -> band 5 folds too
    ax15 = (np.arange(1536, dtype=np.float32) / np.float32(1024)) * np.float32(4.0)
    folded, block = wn.fold_plan(ax15, ax, ax, scale, 128)
    assert list(folded) == [False, True, True, True, True] and block == (512, 512, 512)
    # a smaller budget (2^21 samples) keeps band 6 per sample as well
    monkeypatch.setenv("WN_FOLD_BUDGET", str(1 << 21))
    folded, block = wn.fold_plan(ax, ax, ax, scale, 128)
    assert list(folded) == [False, False, False, True, True] and block == (128, 128, 128)
    monkeypatch.setenv("WN_FOLD_BUDGET", "0")
    assert not wn.fold_plan(ax, ax, ax, scale, 128)[0].any()
    monkeypatch.delenv("WN_FOLD_BUDGET")
    # coordinates that are not commensurate with the tile never fold
    irr = (np.arange(512, dtype=np.float32) * np.float32(0.0137)).astype(np.float32)
    folded, block = wn.fold_plan(irr, irr, irr, scale, 128)
    assert not folded.any() and block == (1, 1, 1)
    # non-power-of-two tile (n=30), step 1/2 cell: periods 60 / 30 / 15 samples for scales 1 / 2 / 4.  A lattice this
    # small is not worth a dependent launch per level (nothing folds) unless the per-level cost is set to zero; then the
    # suffix of the canonical order that has whole periods on every axis folds: scale 1 has no whole z period in 64 slices
    xs = np.arange(240, dtype=np.float32) * np.float32(0.5)
    assert not wn.fold_plan(xs, xs[:120], xs[:64], [1.0, 2.0, 4.0, 0.25], 30)[0].any()
    monkeypatch.setenv("WN_FOLD_LEVEL_COST", "0")
    folded, block = wn.fold_plan(xs, xs[:120], xs[:64], [1.0, 2.0, 4.0, 0.25], 30)
    assert list(folded) == [False, True, True, False] and block == (30, 30, 30)
    folded, block = wn.fold_plan(xs, xs[:120], xs[:120], [1.0, 2.0, 4.0, 0.25], 30)
    assert list(folded) == [True, True, True, False] and block == (60, 60, 60)


def test_viewer_json_format_matches_reference_converter(golden_dir):
    """The JSON emitter reproduces threejs/convert_raw_to_json.py on the shipped .raw (reference present only)."""
    import json
    import numpy as np
    ref_json = "/root/reference/threejs/result_json/wavelet_noise_3d_sliced_octave4.json"
    if not os.path.exists(ref_json):
        pytest.skip("reference tree not present")
    import torch  # noqa: F401  (the package import below needs the built library, not a GPU)
    exp = wnpkg.load_sub("experiment")
    raw = np.fromfile(os.path.join(golden_dir, "result_raw", "wavelet_noise_3Dsliced_octave_4.raw"), dtype="<f4")
    got = exp.raw_to_json_dict(raw)
    want = json.load(open(ref_json))
    assert got["width"] == want["width"] and got["height"] == want["height"]
    assert got["original_range"] == want["original_range"]
    assert got["data"] == want["data"]


def test_export_json_writes_the_viewer_files(golden_dir, tmp_path):
    """export_json on the 15 shipped images: the reference viewer's file names (threejs/result_json/*.json) and, when the
    reference tree is present, the same JSON content as threejs/convert_raw_to_json.py produced."""
    import json
    import numpy as np
    import experiment_cases as ex
    exp = wnpkg.load_sub("experiment")
    images = {}
    for octave in ex.OCTAVES:
        for kind in ("w2d", "w3d", "wproj", "p2d", "p3d"):
            images[ex.raw_name(kind, octave)[:-4]] = ex.load_raw(golden_dir, kind, octave)
    exp.export_json(images, str(tmp_path))
    names = sorted(os.listdir(tmp_path))
    assert len(names) == 15 and "wavelet_noise_3d_projected_octave5.json" in names and "perlin_noise_2d_octave3.json" in names
    ref_dir = "/root/reference/threejs/result_json"
    for name in names:
        got = json.load(open(tmp_path / name))
        assert got["width"] == 256 and got["height"] == 256 and len(got["data"]) == 65536
        assert 0.0 <= min(got["data"]) and max(got["data"]) <= 1.0
        if os.path.isdir(ref_dir):
            want = json.load(open(os.path.join(ref_dir, name)))
            # min / max (and the data normalised with them) exactly; mean / std to the last-digit differences between
            # numpy versions (the shipped files were written by another one)
            g, w = got["original_range"], want["original_range"]
            assert g["min"] == w["min"] and g["max"] == w["max"], name
            assert abs(g["mean"] - w["mean"]) <= 1e-12 and abs(g["std"] - w["std"]) <= 1e-12, name
            assert got["data"] == want["data"], name


def test_group_sharding_matches_the_python_partition():
    """The C ABI's device groups cut volumes and images exactly like sharding.py (the process-per-GPU path): every
    unit is owned by exactly one rank, block-cyclic = 32-slice chunks dealt round-robin, slabs = contiguous with the
    remainder on the lowest ranks.  Host-only, no GPU."""
    import ctypes as C
    import numpy as np
    wn = wnpkg.load()
    sh = wnpkg.load_sub("sharding")
    lib = wn._lib.lib
    for total in (0, 1, 31, 32, 33, 128, 1000, 1024, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(total, np.int32)
            for rank in range(world):
                for sharding in (wn.WN_SHARD_SLAB, wn.WN_SHARD_CYCLIC):
                    buf = np.empty(max(total, 1), np.int32)
                    n = C.c_int()
                    wn._lib.check(lib.wn_debug_shard_indices(total, rank, world, sharding,
                                                             buf.ctypes.data_as(C.POINTER(C.c_int32)), buf.size, C.byref(n)))
                    got = buf[:n.value]
                    if sharding == wn.WN_SHARD_CYCLIC:
                        want = sh.cyclic_slab_indices(total, rank, world)
                        seen[got] += 1
                    else:
                        b, e = sh.slab_range(total, rank, world)
                        want = np.arange(b, e)
                    assert np.array_equal(got, want), (total, world, rank, sharding)
            assert (seen == 1).all()


def test_bench_helpers():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    peak, src = b.peaks()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
    t = b.profiled_traffic_bytes()
    assert t is None or 4.0e9 < t < 6.0e9
    s = b.ClockSampler(0)
    s.proc = object()
    s.lines = ["0, 1965, 1965, 400.5, Not Active, Not Active, Not Active, Active", "0, 1950, 1965, 410.0, Not Active, Not Active, Not Active, Not Active"]
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    out = s.stop()
    assert out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965 and out["reasons"] == ["sw_power_cap"]


def test_header_is_valid_c_and_links_from_c(tmp_path):
    """include/wn_b200.h is a C header: a C99 translation unit that includes it compiles with -Wall -Werror -pedantic,
    links against libwn_b200.so and can call the host-only entry points (what a C / cgo / JNI host would bind)."""
    import subprocess
    src = tmp_path / "host.c"
    src.write_text('''
#include "wn_b200.h"
#include <stdio.h>
int main(void) {
    float ax[64]; int folded[2], block[3], n = -1, i;
    const float scale[2] = { 8.0f, 64.0f };
    for (i = 0; i < 64; ++i) ax[i] = (float)i / 64.0f * 4.0f;
    if (wn_adjust_tile_size(31) != 32) return 1;
    if (wn_debug_fold_plan(ax, 64, ax, 64, ax, 64, scale, 2, 32, folded, block, &n) != WN_OK) return 2;
    if (wn_debug_fold_plan(ax, 64, ax, 64, ax, 64, scale, 2, 1, folded, block, &n) != WN_EINVAL) return 3;
    printf("%s folded=%d\\n", wn_version(), n);
    return 0;
}
''')
    exe = tmp_path / "host"
    libdir = os.path.join(ROOT, "wavelet-noise-in-ray-tracing_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-lwn_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "folded=" in out


def test_fold_plan_properties_on_random_lattices(monkeypatch):
    """Randomised host-logic check of the fold planner: the folded bands are a suffix of the ascending-scale order, the
    period block fits the lattice and shrinks it at least 4x, and every folded band really repeats with the block's
    periods (axis-table entries recomputed here in numpy float32: same weights, same cell modulo the tile)."""
    import numpy as np
    wn = wnpkg.load()
    monkeypatch.setenv("WN_FOLD_LEVEL_COST", "0")
    rs = np.random.RandomState(11)

    def entries(coords, scale, n):
        a = coords.astype(np.float32) * np.float32(scale) - np.float32(0.5)
        mid = np.ceil(a).astype(np.int64)
        t = mid.astype(np.float32) - a
        w0 = t * t * np.float32(0.5)
        s1 = np.float32(1.0) - t
        w2 = s1 * s1 * np.float32(0.5)
        w1 = np.float32(1.0) - w0 - w2
        return np.stack([w0.view(np.uint32), w1.view(np.uint32), w2.view(np.uint32), ((mid - 1) % n).astype(np.uint32)], -1)

    folded_cases = 0
    for _ in range(60):
        n = int(rs.choice([8, 16, 30, 32, 62, 128]))
        dims = [int(rs.choice([12, 37, 64, 96, 128, 200, 256])) for _ in range(3)]
        step = float(rs.choice([0.125, 0.25, 0.5, 1.0, 0.3, 1.0 / 3.0]))
        offs = [float(rs.choice([0.0, 0.5, 3.25, -7.0])) for _ in range(3)]
        axes = [(np.arange(d, dtype=np.float32) * np.float32(step) + np.float32(o)).astype(np.float32) for d, o in zip(dims, offs)]
        nb = int(rs.randint(1, 7))
        scale = rs.choice([0.25, 0.5, 1.0, 2.0, 4.0, 8.0, 3.0], nb, replace=True).astype(np.float32)
        folded, block = wn.fold_plan(axes[0], axes[1], axes[2], scale, n)
        order = np.argsort(scale, kind="stable")
        flags = folded[order]
        assert not (flags[:-1] & ~flags[1:]).any(), (scale, folded)           # suffix of the canonical order
        if not folded.any():
            assert block == (1, 1, 1)
            continue
        folded_cases += 1
        assert all(1 <= b <= d for b, d in zip(block, dims))
        assert block[0] * block[1] * block[2] * 4 <= dims[0] * dims[1] * dims[2]
        for b in np.flatnonzero(folded):
            for ax, L in zip(axes, block):
                e = entries(ax, scale[b], n)
                if L < len(ax):
                    assert (e[L:] == e[:-L]).all(), (n, step, scale[b], L)
    assert folded_cases >= 10


def test_projected_grid_kernel_has_no_contracted_packed_products():
    """ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false (the scalar .rn forms are
    never contracted).  The projected-noise kernels (k_proj_grid, k_proj<points / grid>, k_wmultiband) must keep the
    reference's separately rounded products (WaveletNoise.cpp:239-256), so the only FFMA2 their SASS may hold are the
    harmless ones: a product added to zero, or a multiplication by 0.5 (exact,
    so x * 0.5 + y rounds like fl(x * 0.5) + y)."""
    import shutil
    lib = wnpkg.load_sub("_lib")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    inside, packed, bad = False, 0, []
    for line in sass.splitlines():
        if "Function :" in line:
            inside = any(k in line for k in ("k_proj", "k_wmultiband"))     # every kernel that evaluates projected noise
            continue
        if not inside:
            continue
        packed += bool(re.search(r"\bF(ADD|MUL)2\b", line))
        if re.search(r"\bFFMA2\b", line):
            ops = line.split("FFMA2", 1)[1].split(";")[0]
            if not (re.search(r",\s*-?0\.5\s*,", ops) or re.search(r",\s*RZ(\.F32)?\s*$", ops.strip())):
                bad.append(line.strip())
    assert packed >= 40, "the projected-noise kernels no longer use the packed FP32 pipe"
    assert not bad, "contracted packed product in a projected-noise kernel:\n" + "\n".join(bad[:5])
