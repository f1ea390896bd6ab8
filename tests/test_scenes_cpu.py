"""Synthetic scenes (BASELINE configs 3 and 4) and the builder-defined RTOW path on the CPU oracle."""
import numpy as np

import ptb200


def test_sphere_field_is_deterministic_and_complete(ptb, core_lib, oracle):
    a, cam = ptb.scenes.rtow_sphere_field()
    b, _ = ptb.scenes.rtow_sphere_field()
    assert a.to_ptscene_bytes() == b.to_ptscene_bytes()
    assert 470 <= len(a.sph_mat) <= 489 and len(a.mats) == len(a.sph_mat)
    kinds = np.bincount(a.mats["type"], minlength=5)
    assert kinds[ptb.PT_MAT_LAMBERTIAN] > 300 and kinds[ptb.PT_MAT_METAL] > 30 and kinds[ptb.PT_MAT_DIELECTRIC] > 5 and kinds[ptb.PT_MAT_DIFFUSE_LIGHT] == 1
    ok, msg, st = ptb.bvh_selftest(a)
    assert ok, msg
    rgb, _, stats = oracle.render(a, 96, 54, 4, 10, camera=cam)
    assert rgb.mean() > 60 and stats["absorbed_paths"] > 0 and stats["emitter_paths"] > 0.5 * stats["samples"]
    # RNG-stream property: tile renders equal the full frame
    part, _, _ = oracle.render(a, 96, 54, 4, 10, camera=cam, rect=(32, 16, 40, 20))
    rows = slice(54 - 16 - 20, 54 - 16)
    assert np.array_equal(part[rows, 32:72], rgb[rows, 32:72])


def test_displaced_sphere_mesh(ptb, core_lib, duck, oracle):
    sc = ptb.scenes.displaced_sphere_in_cornell(duck, n=48)
    assert len(sc.tri_mat) == 12 + 2 * 48 * 48
    assert np.bincount(sc.tri_mat, minlength=5).tolist()[:4] == [2, 6, 2, 2]
    ok, msg, st = ptb.bvh_selftest(sc)
    assert ok, msg
    rgb, _, stats = oracle.render(sc, 64, 36, 4, 6)
    assert rgb.mean() > 20 and 1.5 < stats["rays"] / stats["samples"] < 4


def test_mixed_scene_touches_every_branch(ptb, core_lib, oracle):
    sc, cam = ptb.scenes.mixed_material_test_scene()
    ok, msg, st = ptb.bvh_selftest(sc)
    assert ok, msg
    rgb, yuv, stats = oracle.render(sc, 80, 45, 8, 8, camera=cam)
    assert stats["emitter_paths"] > 0 and stats["miss_paths"] > 0 and stats["absorbed_paths"] > 0 and stats["depth_paths"] > 0
    assert rgb.max() == 255 and rgb.mean() > 10
