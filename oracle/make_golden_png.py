#!/usr/bin/env python3
"""Golden texels for the repo's PNG reader, produced by the reference's own decoder (oracle/_ref/ref_stb = the vendored stb_image.h) the way the
reference calls it for an EMBEDDED texture (stbi_load_from_memory, native channel count: src/HostScene.cpp:18-26).

    python oracle/make_golden_png.py       (needs /root/reference for `make -C oracle _ref/ref_stb`; writes tests/golden/png/)

The files are written byte by byte by tools/fuzz_png.py::handmade_png (fixed seeds): colour types 0 / 2 / 3 / 4 / 6, bit depths 1 .. 16, a
random filter type on every row, Adam7 interlacing, tRNS for palettes and as a colour key, the zlib stream split over two IDAT chunks — what an
image library would not write.  Beside each `<name>.png`: `<name>.raw.gz` = "W H C\\n" + stb_image's bytes.  TEST INFRASTRUCTURE."""
import gzip, importlib.util, random, subprocess, tempfile
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden" / "png"
STB = ROOT / "oracle" / "_ref" / "ref_stb"
spec = importlib.util.spec_from_file_location("fuzz_png", ROOT / "tools" / "fuzz_png.py")


def main():
    if not STB.exists():
        raise SystemExit("oracle/_ref/ref_stb missing: make -C oracle _ref/ref_stb (needs /root/reference)")
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    OUT.mkdir(parents=True, exist_ok=True)
    seen, seed = set(), 0
    while len(seen) < 26 and seed < 4000:
        seed += 1
        rnd, rng = random.Random(seed), np.random.default_rng(seed)
        w, h = rnd.randint(1, 40), rnd.randint(1, 30)
        data, desc = fz.handmade_png(rng, rnd, w, h)
        has_trns = b"tRNS" in data
        key = (desc, has_trns)
        if key in seen:
            continue
        seen.add(key)
        name = desc.replace("handmade ", "").replace(" ", "_").lower() + ("_trns" if has_trns else "") + f"_{w}x{h}"
        (OUT / f"{name}.png").write_bytes(data)
        with tempfile.TemporaryDirectory() as td:
            raw = Path(td) / "o.raw"
            r = subprocess.run([str(STB), str(OUT / f"{name}.png"), str(raw)], capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit(f"{name}: {r.stderr}")
            (OUT / f"{name}.raw.gz").write_bytes(gzip.compress(raw.read_bytes(), 9, mtime=0))
        print(name, len(data), "bytes")


if __name__ == "__main__":
    main()
