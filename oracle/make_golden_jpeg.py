#!/usr/bin/env python3
"""Golden texels for the repo's JPEG reader, produced by the reference's own decoder (oracle/_ref/ref_stb = the vendored stb_image.h).

    python oracle/make_golden_jpeg.py      (needs /root/reference for `make -C oracle _ref/ref_stb`; writes tests/golden/jpeg/)

Small JPEG files written with Pillow (sizes that are not multiples of the MCU, 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 chroma, grey, restart intervals,
optimised Huffman tables, low and high quality, progressive files with and without chroma subsampling) and, beside each, what stb_image decodes
from it: `<name>.raw.gz` = "W H C\\n" + bytes.  TEST INFRASTRUCTURE."""
import gzip, io, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden" / "jpeg"
STB = ROOT / "oracle" / "_ref" / "ref_stb"


def picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(x / 5.0 + seed), 127 + 120 * np.cos(y / 7.0), (x * 9 + y * 5 + seed * 31) % 256], axis=2)
    img += rng.normal(0, 12, img.shape)
    img[h // 3: h // 3 + 3, :, :] = 255  # hard edges: ringing exercises the clamps
    img[:, w // 2: w // 2 + 2, :] = 0
    return np.clip(img, 0, 255).astype(np.uint8)


CASES = [  # name, w, h, mode, kwargs
    ("c444_q90_33x21", 33, 21, "RGB", dict(quality=90, subsampling=0)),
    ("c422_q75_50x31", 50, 31, "RGB", dict(quality=75, subsampling=1)),
    ("c420_q60_67x45", 67, 45, "RGB", dict(quality=60, subsampling=2)),
    ("c420_q95_16x16", 16, 16, "RGB", dict(quality=95, subsampling=2)),
    ("c420_q30_1x1", 1, 1, "RGB", dict(quality=30, subsampling=2)),
    ("c420_q85_17x1", 17, 1, "RGB", dict(quality=85, subsampling=2)),
    ("c420_q85_2x19", 2, 19, "RGB", dict(quality=85, subsampling=2)),
    ("c420_opt_q80_40x40", 40, 40, "RGB", dict(quality=80, subsampling=2, optimize=True)),
    ("c420_rst_q80_70x50", 70, 50, "RGB", dict(quality=80, subsampling=2, restart_marker_blocks=2)),
    ("c444_q100_24x24", 24, 24, "RGB", dict(quality=100, subsampling=0)),
    ("grey_q70_37x29", 37, 29, "L", dict(quality=70)),
    ("c420_q5_48x32", 48, 32, "RGB", dict(quality=5, subsampling=2)),
    ("progressive_q80_32x32", 32, 32, "RGB", dict(quality=80, progressive=True)),
    ("progressive_c444_q92_45x27", 45, 27, "RGB", dict(quality=92, progressive=True, subsampling=0)),
    ("progressive_c420_q40_61x33", 61, 33, "RGB", dict(quality=40, progressive=True, subsampling=2)),
    ("progressive_grey_q85_19x40", 19, 40, "L", dict(quality=85, progressive=True)),
    ("progressive_c422_rst_q75_52x36", 52, 36, "RGB", dict(quality=75, progressive=True, subsampling=1, restart_marker_blocks=3)),
]


def main():
    if not STB.exists():
        raise SystemExit("oracle/_ref/ref_stb missing: make -C oracle _ref/ref_stb (needs /root/reference)")
    OUT.mkdir(parents=True, exist_ok=True)
    for i, (name, w, h, mode, kw) in enumerate(CASES):
        img = picture(w, h, i)
        im = Image.fromarray(img if mode == "RGB" else img[:, :, 0], mode)
        buf = io.BytesIO()
        im.save(buf, "JPEG", **kw)
        (OUT / f"{name}.jpg").write_bytes(buf.getvalue())
        with tempfile.TemporaryDirectory() as td:
            raw = Path(td) / "o.raw"
            r = subprocess.run([str(STB), str(OUT / f"{name}.jpg"), str(raw)], capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit(f"{name}: {r.stderr}")
            (OUT / f"{name}.raw.gz").write_bytes(gzip.compress(raw.read_bytes(), 9, mtime=0))
        print(name, len(buf.getvalue()), "bytes")


if __name__ == "__main__":
    main()
