#!/usr/bin/env python3
"""Random sphere sets (all four RTOW materials) rendered by the CPU restatement and by the reference's dead classes compiled for the host
(oracle/_ref/ref_cpu_spheres, build container only), with g++'s draw order.  Bit for bit.  tools/fuzz_oracle_spheres.py [cases]"""
import importlib.util, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
from PIL import Image
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ptb200  # noqa: E402
import _oracle  # noqa: E402
spec = importlib.util.spec_from_file_location("mgs", ROOT / "oracle" / "make_golden_spheres.py")
mgs = importlib.util.module_from_spec(spec); spec.loader.exec_module(mgs)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
orc = _oracle.load()
orc.set_triple_draw_order_zyx(True)
bad = 0
with tempfile.TemporaryDirectory() as td:
    td = Path(td)
    for it in range(cases):
        rng = np.random.default_rng(31337 + it)
        sc = mgs.random_spheres(ptb200, 100 + it)
        cam = dict(look_from=tuple(float(v) for v in rng.uniform(-0.5, 0.5, 3) + np.array([0, 0, 0.5])), front=tuple(float(v) for v in rng.uniform(-0.2, 0.2, 3) + np.array([0, 0, -1.0])),
                   vfov=float(rng.uniform(35, 70)), hfov=float(rng.uniform(35, 70)))
        w, h, spp, depth = 2 * int(rng.integers(6, 24)), 2 * int(rng.integers(5, 16)), int(rng.integers(1, 7)), int(rng.integers(1, 14))
        flat, ppm = td / "s.ptscene", td / "r.ppm"
        flat.write_bytes(sc.to_ptscene_bytes())
        subprocess.run([str(mgs.REF), str(flat), str(w), str(h), str(spp), str(depth), str(ppm), "--cam", *[repr(float(v)) for v in (*cam["look_from"], *cam["front"], cam["vfov"], cam["hfov"])]],
                       check=True, capture_output=True, text=True)
        ref = np.array(Image.open(ppm).convert("RGB"))
        rgb, _, _ = orc.render(sc, w, h, spp, depth, camera=cam)
        if not np.array_equal(rgb, ref):
            bad += 1
            print("MISMATCH", it, len(sc.sph_mat), (w, h, spp, depth), int((np.abs(rgb.astype(int) - ref.astype(int)).max(axis=2) > 0).sum()))
print("cases", cases, "bad", bad)
