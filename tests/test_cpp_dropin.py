"""The reference-facing C++ layer (wavelet-noise-in-ray-tracing_b200/cpp): GPU tests run the built binaries and
compare with the goldens / the oracle; the CPU test checks the layer builds and that the UNMODIFIED reference
driver compiles and links against it."""
import os
import struct
import subprocess

import numpy as np
import pytest

import experiment_cases as ex
import wnpkg

CPP = os.path.join(wnpkg.ROOT, wnpkg.PKG, "cpp")
BIN = os.path.join(CPP, "bin")


def test_cpp_layer_builds_and_reference_driver_links():
    subprocess.run(["make", "-s", "-C", CPP, "all"], check=True)
    for exe in ("experiment_b200", "dropin_selftest", "sharded_b200"):
        assert os.access(os.path.join(BIN, exe), os.X_OK)
    if os.path.isdir("/root/reference/experient"):
        subprocess.run(["make", "-s", "-C", CPP, "reference_dropin"], check=True)
        syms = subprocess.run(["nm", "-D", os.path.join(BIN, "reference_main_dropin")], capture_output=True, text=True).stdout
        assert "wn_tile_build_from_gaussian" in syms and "wn_perlin_points" in syms and "wn_eval3d_points" in syms
    import torch
    if not torch.cuda.is_available():       # no GPU: the binaries must fail loudly, not fall back
        r = subprocess.run([os.path.join(BIN, "dropin_selftest")], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


def _golden_equal(out_dir, golden_dir):
    for octave in ex.OCTAVES:
        for kind in ("w2d", "w3d", "wproj", "p2d", "p3d"):
            name = ex.raw_name(kind, octave)
            got = np.fromfile(os.path.join(out_dir, name), dtype="<f4")
            want = ex.load_raw(golden_dir, kind, octave).ravel()
            assert got.size == want.size, name
            assert (got.view(np.uint32) == want.view(np.uint32)).all(), name


@pytest.mark.gpu
def test_cpp_experiment_driver_reproduces_goldens(tmp_path, golden_dir):
    out = tmp_path / "result_raw"
    r = subprocess.run([os.path.join(BIN, "experiment_b200"), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Generated Wavelet 3D Projected Octave 5 noise" in r.stdout
    _golden_equal(str(out), golden_dir)


@pytest.mark.gpu
def test_unmodified_reference_driver_on_gpu_classes(tmp_path, golden_dir):
    """experient/main.cpp itself, compiled against our WaveletNoise.h/.cpp and PerlinNoise.hpp (scalar calls)."""
    exe = os.path.join(BIN, "reference_main_dropin")
    if not os.path.exists(exe):
        pytest.skip("reference_main_dropin was not built (needs /root/reference at build time)")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    _golden_equal(str(tmp_path / "result_raw"), golden_dir)


@pytest.mark.gpu
def test_cpp_scalar_surface_matches_oracle(oracle):
    r = subprocess.run([os.path.join(BIN, "dropin_selftest")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Tile size adjusted to 128" in r.stderr
    got = {}
    for ln in r.stdout.splitlines():
        parts = ln.split()
        if len(parts) == 2:
            got[parts[0]] = parts[1]
    def f(name):
        return struct.unpack("<f", struct.pack("<I", int(got[name], 16)))[0]
    p = [3.25, -7.5, 100.125]
    tile = oracle.generate_tile(127, 4242, 3)
    assert got["tile_size"] == "128" and f("empty_eval3d") == 0.0
    assert f("eval3d") == oracle.eval3d(tile, 128, p)
    assert f("copy_eval3d") == f("eval3d")
    assert f("proj") == oracle.eval3d_projected(tile, 128, p, [.6, 0, .8])
    assert f("coeff0") == tile[0] and f("coeff_last") == tile[-1]
    ta, tb = oracle.generate_tile(16, 1, 3), oracle.generate_tile(16, 2, 3)
    assert f("assign_before") == oracle.eval3d(ta, 16, p)
    assert f("assign_after") == f("assign_source") == oracle.eval3d(tb, 16, p) and f("assign_coeff0") == tb[0]
    assert f("assign_before") != f("assign_after")
    g = oracle.rng(7)
    oracle.generate_tile(16, 7, 2, g)
    t2 = oracle.generate_tile(16, 7, 2, g)
    assert f("eval2d_second") == oracle.eval2d(t2, 16, p[:2])
    _, _, mn, mx = oracle.stats(tile)
    assert f("stats_min") == mn and f("stats_max") == mx
    assert "tile3D stats: avg=" in r.stdout
    perm = oracle.perlin_perm(12345)
    assert got["perm0"] == str(perm[0])
    assert f("perlin") == np.float32(oracle.perlin_noise(perm, 3.25, -7.5, 100.125))
    assert f("perlin2d") == np.float32(oracle.perlin_noise(perm, np.float32(0.3), np.float32(0.7), 0.0))
    assert f("tex0") == np.float32(oracle.wavelet_texture_value(tile, 128, [0.5, 1.5, -2.25], 1.0, 4))
    assert f("tex1") == np.float32(oracle.wavelet_texture_value(tile, 128, [9.0, 9.5, 10.25], 1.0, 4))


@pytest.mark.gpu
def test_sharded_cpp_drivers_are_gpu_count_invariant():
    """cpp/sharded_main.cpp (C ABI device groups): config 3 at 256^3 and config 4 at 512^2 produce the same output hash
    for every GPU count on this box and for both z shardings."""
    import json
    import torch
    exe = os.path.join(BIN, "sharded_b200")
    counts = [k for k in (1, 2, 4, 8) if k <= torch.cuda.device_count()]
    vol, plane = set(), set()
    for n in counts:
        for sharding in ("cyclic", "slab"):
            r = subprocess.run([exe, "volume", "--gpus", str(n), "--size", "256", "--reps", "2", "--gather", "--sharding",
                                sharding], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-1000:]
            line = json.loads(r.stdout.strip().splitlines()[-1])
            assert line["n_gpus"] == n and line["gsamples_s"] > 0
            vol.add(line["fnv1a64"])
        r = subprocess.run([exe, "plane", "--gpus", str(n), "--size", "512", "--reps", "1"], capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-1000:]
        line = json.loads(r.stdout.strip().splitlines()[-1])
        plane.add((line["fnv1a64_projected"], line["fnv1a64_perlin"]))
    assert len(vol) == 1 and len(plane) == 1, (vol, plane)


@pytest.mark.gpu
@pytest.mark.parametrize("choice,name", [("1\n4\n", "raytrace_Wavelet3D_octave4.png"), ("0\n4\n", "raytrace_Perlin_octave4.png")])
def test_deferred_renderer_reproduces_shipped_png(tmp_path, golden_dir, choice, name):
    """Config 5: the reference scene at the shipped 1000x500 / 100 spp, noise path batched on the GPU: PNG byte-identical."""
    exe = os.path.join(BIN, "render_deferred")
    if not os.path.exists(exe):
        pytest.skip("render_deferred was not built (needs the reference's RTIOW headers at build time)")
    r = subprocess.run([exe, "--out", str(tmp_path)], input=choice, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    got = open(tmp_path / name, "rb").read()
    want = open(os.path.join(golden_dir, "result_raytracing", name), "rb").read()
    assert got == want, r.stdout[-500:]
    assert "texture lookups" in r.stdout
