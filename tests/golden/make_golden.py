#!/usr/bin/env python
"""Regenerate tests/golden/ (run in the build container, where /root/reference exists).

1. result_raw/*.raw  -- byte copies of the 15 images the reference SHIPS in experient/result_raw
   (float32 LE, 256x256; produced by experient/main.cpp:131-168).  These are the reference's only
   golden vectors.
2. result_raytracing/*.png -- byte copies of the two renders the reference ships (config 5 parity).
3. ref_vectors.npz   -- outputs of the UNMODIFIED reference compiled by oracle/Makefile
   (oracle/_ref/libwnref.so) on small seeded inputs: tiles for n in {8,16,30,31}, evaluate2D/3D/
   3DProjected at random / negative / half-integer / huge points, Perlin (seeds 12345 and 5489),
   the two texture::value hooks.  The GPU box has no /root/reference, so these travel as fixtures.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import RefLib, build_oracle  # noqa: E402

REF = "/root/reference"


def main():
    build_oracle()
    dst = os.path.join(HERE, "result_raw")
    os.makedirs(dst, exist_ok=True)
    src = os.path.join(REF, "experient", "result_raw")
    for f in sorted(os.listdir(src)):
        if f.endswith(".raw"):
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))

    # the two renders the reference ships (result_raytracing/*.png, 1000x500, 100 spp; main.cpp:80-215)
    dstp = os.path.join(HERE, "result_raytracing")
    os.makedirs(dstp, exist_ok=True)
    srcp = os.path.join(REF, "result_raytracing")
    for f in sorted(os.listdir(srcp)):
        if f.endswith(".png"):
            shutil.copyfile(os.path.join(srcp, f), os.path.join(dstp, f))

    ref = RefLib()
    rs = np.random.RandomState(20251018)
    out = {}
    # tiles: small sizes incl. non-power-of-two (30) and odd (31 -> 32)
    for n in (8, 16, 30, 31):
        for dims in (2, 3):
            nz = ref.noise(n, 777 + n).generate(dims)
            out[f"tile{dims}d_n{n}_seed{777 + n}"] = nz.tile()
    # evaluation on the n=128 seed 12345 tiles used everywhere in the reference
    pts3 = np.concatenate([
        rs.uniform(-300, 300, (4096, 3)),
        rs.uniform(-2, 2, (1024, 3)),
        np.round(rs.uniform(-64, 64, (512, 3)) * 2) / 2,          # integers and half-integers
        rs.uniform(-1, 1, (256, 3)) * 1.0e6,                       # huge
        np.array([[3.25, -7.5, 100.125], [0, 0, 0], [127.5, 127.5, 127.5], [128, -128, 0.5]]),
    ]).astype(np.float32)
    out["pts3"] = pts3
    n3 = ref.noise(128, 12345).generate(3)
    n2 = ref.noise(128, 12345).generate(2)
    out["eval3d"] = n3.eval3d_points(pts3)
    out["eval2d"] = n2.eval2d_points(pts3[:, :2].copy())
    normals = rs.normal(size=(pts3.shape[0], 3))
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    normals = normals.astype(np.float32)
    normals[:8] = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0], [0, 0, -1], [.6, 0, .8], [0, .6, .8],
                            [.57735026, .57735026, .57735026], [-1, 0, 0]], np.float32)
    out["normals"] = normals
    sel = slice(0, 2048)
    out["proj_pernormal"] = n3.eval3d_projected_points(pts3[sel], normals[sel])
    nshared = (np.array([1, 2, 3], np.float64) / np.sqrt(14.0)).astype(np.float32)
    out["nshared"] = nshared
    out["proj_shared"] = n3.eval3d_projected_points(pts3[sel], nshared)
    # small-tile evaluation incl. n=30 (non power of two Mod)
    t30 = ref.noise(30, 807).generate(3)
    out["eval3d_n30"] = t30.eval3d_points(pts3)
    out["proj_n30"] = t30.eval3d_projected_points(pts3[:512], nshared)
    # Perlin
    for seed in (12345, 5489):
        out[f"perlin_{seed}"] = ref.perlin_points(seed, pts3[:5376])
    # texture hooks (scale 1, octave 4 = what main.cpp uses; plus another combination)
    tp = rs.uniform(-10, 10, (4096, 3)).astype(np.float32)
    out["tex_pts"] = tp
    out["tex_wavelet_s1_o4"] = ref.texture_values("wavelet", 1.0, 4, tp)
    out["tex_wavelet_s0.37_o3"] = ref.texture_values("wavelet", 0.37, 3, tp)
    out["tex_wavelet2d_s1_o4"] = ref.texture_values("wavelet2d", 1.0, 4, tp)
    out["tex_wavelet2d_s0.37_o3"] = ref.texture_values("wavelet2d", 0.37, 3, tp)
    out["tex_perlin_s1_o4"] = ref.texture_values("perlin", 1.0, 4, tp)
    out["tex_perlin_s0.37_o5"] = ref.texture_values("perlin", 0.37, 5, tp)
    # projected noise far from the origin and with normals on / next to a coordinate axis (the row culling of k_proj
    # must only ever skip candidates the reference weighs with 0): |p| ~ 1e3 .. 1e6, 16 normals per magnitude class
    hp, hn = [], []
    axes = np.eye(3)
    for mag in (1.0e3, 1.0e4, 1.0e5, 1.0e6):
        for k in range(64):
            p = rs.uniform(-1, 1, 3) * mag
            if k % 4 == 0:
                p = np.round(p)                                  # integer coordinates
            elif k % 4 == 1:
                p = np.round(p * 2) / 2                          # half-integers
            a = axes[k % 3] * (1 if (k // 3) % 2 == 0 else -1)
            kind = (k // 6) % 4
            if kind == 0:
                nv = a                                           # exactly on an axis
            elif kind == 1:
                nv = a + rs.normal(size=3) * 1e-4                # within 1e-4 of an axis
            elif kind == 2:
                nv = a + rs.normal(size=3) * 1e-6
            else:
                nv = rs.normal(size=3)                           # generic
            nv = nv / np.linalg.norm(nv)
            hp.append(p)
            hn.append(nv)
    hp, hn = np.array(hp, np.float32), np.array(hn, np.float32)
    out["proj_huge_pts"], out["proj_huge_normals"] = hp, hn
    out["proj_huge"] = n3.eval3d_projected_points(hp, hn)
    out["proj_huge_n30"] = t30.eval3d_projected_points(hp, hn)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
