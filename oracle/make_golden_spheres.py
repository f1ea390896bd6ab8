#!/usr/bin/env python3
"""Whole images of sphere scenes from the reference's dead classes (sphere, lambertian, metal, dielectric, diffuse_light: SURVEY 8a D1-D6) compiled
for the HOST (oracle/_ref/ref_cpu_spheres), so that the restatement of those rows is pinned by images and not only by known answers of the pieces.

    python oracle/make_golden_spheres.py     (needs /root/reference for `make -C oracle ref`; writes tests/golden/spheres/)

g++ evaluates the three draws of make_random_float3's argument list right to left (nvcc's device code: left to right), so these images belong to
the z, y, x draw order: the test switches the restatement to it (pto_set_triple_draw_order_zyx).  Scenes: the 486-sphere field of BASELINE
config 3 and random sphere sets (all four materials, fuzz 0..1.3, ior 1.1..2.4, a sky dome or small lamps).  TEST INFRASTRUCTURE."""
import gzip, json, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tests" / "golden" / "spheres"
REF = ROOT / "oracle" / "_ref" / "ref_cpu_spheres"


def random_spheres(ptb200, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 30))
    sph = np.zeros((n, 4), np.float32)
    sph[:, :3] = rng.uniform(-1, 1, (n, 3)) * np.array([3.0, 1.5, 2.0]) + np.array([0, 0, -5.0])
    sph[:, 3] = rng.uniform(0.2, 0.9, n)
    n_mat = int(rng.integers(3, 8))
    mats = np.zeros(n_mat, ptb200.MAT_DTYPE)
    for i in range(n_mat):
        t = [ptb200.PT_MAT_DIFFUSE_LIGHT, ptb200.PT_MAT_LAMBERTIAN, ptb200.PT_MAT_METAL, ptb200.PT_MAT_DIELECTRIC][i] if i < 4 else int(rng.choice([ptb200.PT_MAT_LAMBERTIAN, ptb200.PT_MAT_METAL, ptb200.PT_MAT_DIELECTRIC]))
        mats[i] = (t, tuple(rng.uniform(0.1, 0.95, 3)), tuple(rng.uniform(1, 6, 3)) if t == ptb200.PT_MAT_DIFFUSE_LIGHT else (0, 0, 0), -1, -1, float(rng.uniform(0, 1.3)), float(rng.uniform(1.1, 2.4)))
    sph_mat = rng.integers(0, n_mat, n).astype(np.int32)
    sph_mat[:min(n, n_mat)] = np.arange(min(n, n_mat))
    if rng.random() < 0.5:  # a ground sphere and a sky dome
        sph = np.concatenate([sph, np.array([[0, -1001.5, -5, 1000], [0, 0, 0, 3000]], np.float32)])
        sph_mat = np.concatenate([sph_mat, np.array([1 % n_mat, 0], np.int32)])
    return ptb200.Scene(sph=sph, sph_mat=sph_mat, mats=mats)


def main():
    if not REF.exists():
        raise SystemExit("oracle/_ref/ref_cpu_spheres missing: make -C oracle ref (needs /root/reference)")
    import ptb200
    OUT.mkdir(parents=True, exist_ok=True)
    field, field_cam = ptb200.scenes.rtow_sphere_field()
    cases = [("field_96x54_s4_d10", field, field_cam, 96, 54, 4, 10), ("field_64x36_s16_d5", field, field_cam, 64, 36, 16, 5)]
    for k, seed in enumerate((5, 9, 12, 20, 31, 44)):
        rng = np.random.default_rng(7000 + seed)
        cam = dict(look_from=tuple(float(v) for v in rng.uniform(-0.5, 0.5, 3) + np.array([0, 0, 0.5])), front=tuple(float(v) for v in rng.uniform(-0.2, 0.2, 3) + np.array([0, 0, -1.0])),
                   vfov=float(rng.uniform(35, 70)), hfov=float(rng.uniform(35, 70)))
        cases.append((f"random_{k}", random_spheres(ptb200, seed), cam, 2 * int(rng.integers(10, 30)), 2 * int(rng.integers(8, 20)), int(rng.integers(2, 9)), int(rng.integers(2, 12))))
    meta = {}
    for name, sc, cam, w, h, spp, depth in cases:
        scene_file = "field.ptscene.gz" if name.startswith("field") else f"{name}.ptscene.gz"
        sc.save_ptscene(OUT / scene_file)
        with tempfile.TemporaryDirectory() as td:
            flat, ppm = Path(td) / "s.ptscene", Path(td) / "r.ppm"
            flat.write_bytes(sc.to_ptscene_bytes())
            subprocess.run([str(REF), str(flat), str(w), str(h), str(spp), str(depth), str(ppm), "--cam", *[repr(float(v)) for v in (*cam["look_from"], *cam["front"], cam["vfov"], cam["hfov"])]],
                           check=True, capture_output=True, text=True)
            Image.open(ppm).save(OUT / f"{name}.png", optimize=True)
        img = np.array(Image.open(OUT / f"{name}.png"))
        meta[name] = dict(scene=scene_file, width=w, height=h, spp=spp, depth=depth, camera={k: (list(v) if isinstance(v, tuple) else v) for k, v in cam.items()},
                          spheres=int(len(sc.sph_mat)), materials=int(len(sc.mats)), lit_pixels=int((img.max(axis=2) > 0).sum()))
        print(name, meta[name])
    (OUT / "cases.json").write_text(json.dumps(meta, indent=1) + "\n")


if __name__ == "__main__":
    main()
