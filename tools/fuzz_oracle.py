#!/usr/bin/env python3
"""Random small scenes — triangle soups with random universal materials, emitters, textures and texture coordinates, random cameras, sizes,
spp and depths — rendered by the CPU restatement (oracle/pt_oracle.c) and by the reference's OWN headers compiled for the host
(oracle/_ref/ref_cpu, build container only).  Bit for bit: RGB and I420.  tools/fuzz_oracle.py [cases] [seed]
The committed whole-image fixtures all show cornell_duck; this is the same gate on geometry, materials and textures the fixtures do not have."""
import subprocess, sys, tempfile
from pathlib import Path
import numpy as np
from PIL import Image
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ptb200  # noqa: E402
import _oracle  # noqa: E402
REF = ROOT / "oracle" / "_ref" / "ref_cpu"


def random_scene(rng):
    n = int(rng.integers(1, 40))
    centre = rng.uniform(-1, 1, (n, 1, 3)) * np.array([2.0, 1.5, 1.0]) + np.array([0, 0, -4.0])
    tri = (centre + rng.normal(0, rng.uniform(0.2, 1.5), (n, 3, 3))).astype(np.float32)
    if rng.random() < 0.5:  # a floor and a back wall: paths that bounce
        quad = lambda a, b, c, d: [[a, b, c], [a, c, d]]
        # (tilted a little: a ray that slides INSIDE an axis-aligned plane which is also a face of one of the reference's boxes makes its slab
        # test compute 0 * inf = NaN and drop a hit that exists — a property of the reference's tree, which no other tree can reproduce)
        j = lambda: float(rng.uniform(-0.05, 0.05))
        A, B, C, D = [-4 + j(), -2 + j(), -1 + j()], [4 + j(), -2 + j(), -1 + j()], [4 + j(), -2 + j(), -8 + j()], [-4 + j(), -2 + j(), -8 + j()]
        E, F = [4 + j(), 3 + j(), -8 + j()], [-4 + j(), 3 + j(), -8 + j()]
        tri = np.concatenate([tri, np.array(quad(A, B, C, D) + quad(D, C, E, F), np.float32)])
    n = len(tri)
    n_tex = int(rng.integers(0, 3))
    texs = [rng.integers(0, 256, (int(rng.integers(2, 9)), int(rng.integers(1, 9)), 3)).astype(np.float32) for _ in range(n_tex)]
    n_mat = int(rng.integers(1, 6))
    mats = np.zeros(n_mat, ptb200.MAT_DTYPE)
    for i in range(n_mat):
        emit = i == 0 or rng.random() < 0.25  # material 0 always emits: the reference needs a light
        mats[i] = (ptb200.PT_MAT_UNIVERSAL, tuple(rng.uniform(0, 1, 3)), tuple(rng.uniform(0.5, 20, 3)) if emit else (0, 0, 0),
                   int(rng.integers(0, n_tex)) if n_tex and rng.random() < 0.5 else -1, int(rng.integers(0, n_tex)) if n_tex and emit and rng.random() < 0.3 else -1, 0.0, 1.5)
    tri_mat = rng.integers(0, n_mat, n).astype(np.int32)
    tri_mat[0] = 0
    # texture coordinates the reference can look up without leaving its texel array: u >= 0 (it indexes with (int)(fmod(u, 1) * width),
    # negative for u < 0) and frac(v) >= 1 / height (row `height - (int)(v * height)` is one past the end for the first 1 / height of v:
    # Texture.h:61-70; the restatement DEFINES that row as the last one).  Heights are >= 2 here and frac(v) is kept in [0.55, 1).
    uv = np.zeros((n, 3, 2), np.float32)
    uv[:, :, 0] = rng.integers(0, 3, (n, 1)) + rng.uniform(0.0, 0.999, (n, 3))
    uv[:, :, 1] = rng.integers(0, 3, (n, 1)) + rng.uniform(0.55, 0.999, (n, 3))
    uv = uv.reshape(n, 6)
    return ptb200.Scene(tri_pos=tri.reshape(n, 9), tri_uv=uv, tri_mat=tri_mat, mats=mats, textures=texs)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if not REF.exists():
        raise SystemExit("oracle/_ref/ref_cpu missing: make -C oracle ref (needs /root/reference)")
    orc = _oracle.load()
    bad = 0
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for it in range(cases):
            rng = np.random.default_rng(seed * 100003 + it)
            sc = random_scene(rng)
            w, h = 2 * int(rng.integers(2, 14)), 2 * int(rng.integers(2, 10))
            spp, depth = int(rng.integers(1, 6)), int(rng.integers(1, 8))
            cam = dict(look_from=tuple(float(v) for v in rng.uniform(-0.5, 0.5, 3) + np.array([0, 0, 0.5])), front=tuple(float(v) for v in rng.uniform(-0.3, 0.3, 3) + np.array([0, 0, -1.0])),
                       vfov=float(rng.uniform(25, 80)), hfov=float(rng.uniform(25, 80)))
            flat, ppm, yuvf = td / "s.ptscene", td / "r.ppm", td / "r.yuv"
            flat.write_bytes(sc.to_ptscene_bytes())
            r = subprocess.run([str(REF), str(flat), str(w), str(h), str(spp), str(depth), str(ppm), "--yuv", str(yuvf), "--threads", "4",
                                "--cam", *[repr(v) for v in (*cam["look_from"], *cam["front"], cam["vfov"], cam["hfov"])]], capture_output=True, text=True)
            if r.returncode != 0:
                print("ref_cpu failed:", it, r.stderr[-200:]); bad += 1
                continue
            ref = np.array(Image.open(ppm).convert("RGB"))
            ref_yuv = np.frombuffer(yuvf.read_bytes(), np.uint8)
            rgb, yuv, _ = orc.render(sc, w, h, spp, depth, camera=cam, threads=4)
            if not (np.array_equal(rgb, ref) and np.array_equal(yuv, ref_yuv)):
                bad += 1
                d = np.abs(rgb.astype(int) - ref.astype(int)).max(axis=2)
                print("MISMATCH case", it, "tris", len(sc.tri_mat), "mats", len(sc.mats), "tex", len(sc.textures), (w, h, spp, depth), "pixels", int((d > 0).sum()), "max", int(d.max()), "yuv", int((yuv != ref_yuv).sum()))
    print("cases", cases, "bad", bad)


if __name__ == "__main__":
    main()
