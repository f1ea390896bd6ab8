#!/usr/bin/env python3
"""Golden texels for the repo's BMP / TGA / GIF / PNM readers, produced by the reference's own decoder (oracle/_ref/ref_stb = the vendored stb_image.h)
the way the reference calls it for a texture FILE: three requested channels (src/HostScene.cpp:29).

    python oracle/make_golden_images.py     (needs /root/reference for `make -C oracle _ref/ref_stb`; writes tests/golden/images/)

The files are written here byte by byte (no image library), so every header variant is under control: BMP with 12 / 40 / 56 / 108 / 124-byte
headers, 1 / 4 / 8-bit palettes, 16-bit 5-5-5 and 5-6-5 bit fields, 24-bit bottom-up and top-down, 32-bit with default and custom masks, a gap
before the pixels, row padding; TGA types 1 / 2 / 3 / 9 / 10 / 11, 8 / 15 / 16 / 24 / 32 bits, palettes of 15 / 16 / 24 / 32 bits with 8- and 16-bit
indices (incl. an index past the palette), top-down and bottom-up, an image-id field, RLE packets that cross rows.  Beside each
`<name>.bmp|tga`: `<name>.raw.gz` = "W H 3\\n" + what stb_image returns.  Files stb_image refuses are listed in `refused.json`
(the repo's reader must refuse them too).  TEST INFRASTRUCTURE."""
import gzip, json, struct, subprocess, tempfile
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden" / "images"
STB = ROOT / "oracle" / "_ref" / "ref_stb"


def picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 37 + y * 11 + seed * 5) % 256, (x * 5 + y * 53 + seed * 17) % 256, (x * y + seed * 29) % 256], axis=2).astype(np.int64)
    img = (img + rng.integers(0, 40, img.shape)) % 256
    return img.astype(np.uint8)  # [h][w][R, G, B], row 0 = top


def rows_bottom_up(rows, top_down):
    return rows if top_down else rows[::-1]


def bmp(w, h, bpp, seed, hsz=40, top_down=False, compress=0, masks=None, palette_n=None, gap=0, os2=False, index_n=None):
    img = picture(w, h, seed)
    a = (img[:, :, 0].astype(np.uint32) * 3 + seed) % 256  # an "alpha" byte for 32-bit files
    pal = b""
    rows = []
    if bpp <= 8:
        n = palette_n or (1 << bpp)
        rng = np.random.default_rng(seed + 100)
        table = rng.integers(0, 256, (n, 3), dtype=np.uint8)
        idx = (img[:, :, 0].astype(np.int64) + img[:, :, 1]) % (index_n or n)
        for e in table:
            pal += bytes([e[2], e[1], e[0]]) + (b"" if os2 else b"\0")
        for r in range(h):
            if bpp == 8:
                row = bytes(idx[r].astype(np.uint8))
            elif bpp == 4:
                v = list(idx[r]) + [0]
                row = bytes((v[i] << 4) | v[i + 1] for i in range(0, w, 2))
            else:
                v = list(idx[r]) + [0] * 7
                row = bytes(sum(v[i + k] << (7 - k) for k in range(8)) for i in range(0, w, 8))
            rows.append(row)
    elif bpp == 24:
        for r in range(h):
            rows.append(bytes(np.stack([img[r, :, 2], img[r, :, 1], img[r, :, 0]], axis=1).reshape(-1)))
    else:
        mr, mg, mb, ma = masks if masks else ((31 << 10, 31 << 5, 31, 0) if bpp == 16 else (0xff << 16, 0xff << 8, 0xff, 0xff << 24))

        def pack(val, mask):
            if mask == 0:
                return np.zeros_like(val, dtype=np.uint64)
            bits = bin(mask).count("1")
            low = (mask & -mask).bit_length() - 1
            return ((val.astype(np.uint64) >> (8 - bits)) << low) & mask
        for r in range(h):
            v = pack(img[r, :, 0], mr) | pack(img[r, :, 1], mg) | pack(img[r, :, 2], mb) | pack(a[r], ma)
            rows.append(v.astype("<u2" if bpp == 16 else "<u4").tobytes())
    body = b""
    for row in rows_bottom_up(rows, top_down):
        body += row + b"\0" * ((-len(row)) & 3)
    if os2:
        info = struct.pack("<IHHHH", 12, w, h, 1, bpp)
    else:
        info = struct.pack("<IiiHHIIiiII", hsz, w, -h if top_down else h, 1, bpp, compress, len(body), 2835, 2835, 0, 0)
        m = masks if masks else (0, 0, 0, 0)
        if hsz == 40 and compress == 3:
            info += struct.pack("<III", *m[:3])          # BI_BITFIELDS masks follow the 40-byte header
        elif hsz == 56:
            info += struct.pack("<IIII", *m)              # "V3" header: the masks are part of it
            if compress == 3:
                info += struct.pack("<III", *m[:3])      # stb_image reads three more words after a 56-byte header
        elif hsz in (108, 124):
            info += struct.pack("<IIII", *m) + struct.pack("<I", 0x73524742) + b"\0" * 48
            if hsz == 124:
                info += struct.pack("<IIII", 4, 0, 0, 0)
    offset = 14 + len(info) + len(pal) + gap
    return b"BM" + struct.pack("<IHHI", offset + len(body), 0, 0, offset) + info + pal + b"\xAA" * gap + body


def tga(w, h, kind, bits, seed, rle=False, top_down=False, pal_bits=0, pal_n=0, idx_bits=8, image_id=b"", right_to_left=False, pal_start=0, bad_index=False):
    img = picture(w, h, seed)
    if rle:  # runs: four equal pixels in a row, and two constant rows (one run that crosses a row end)
        img = np.repeat(picture((w + 3) // 4, h, seed), 4, axis=1)[:, :w].copy()
        img[h // 2:h // 2 + 2] = img[h // 2, 0]
    a = (img[:, :, 1].astype(np.uint32) * 7 + seed) % 256

    def px555(r, g, b, top):
        return struct.pack("<H", ((r >> 3) << 10) | ((g >> 3) << 5) | (b >> 3) | (0x8000 if top else 0))
    pal = b""
    pixels = []
    if kind == "indexed":
        rng = np.random.default_rng(seed + 200)
        table = rng.integers(0, 256, (pal_n, 4), dtype=np.uint8)
        for e in table:
            if pal_bits in (15, 16):
                pal += px555(int(e[0]), int(e[1]), int(e[2]), e[3] & 1)
            elif pal_bits == 8:
                pal += bytes([e[0]])
            else:
                pal += bytes([e[2], e[1], e[0]]) + (bytes([e[3]]) if pal_bits == 32 else b"")
        idx = (img[:, :, 0].astype(np.int64) * 3 + img[:, :, 2]) % pal_n
        if bad_index:
            idx[h // 2, w // 2] = pal_n + 3
        for r in range(h):
            for c in range(w):
                pixels.append(struct.pack("<B" if idx_bits == 8 else "<H", int(idx[r, c])))
    else:
        for r in range(h):
            for c in range(w):
                R, G, B = (int(v) for v in img[r, c])
                if kind == "grey":
                    pixels.append(bytes([R]) if bits == 8 else bytes([R, int(a[r, c])]))
                elif bits in (15, 16):
                    pixels.append(px555(R, G, B, (r + c) & 1))
                else:
                    pixels.append(bytes([B, G, R]) + (bytes([int(a[r, c])]) if bits == 32 else b""))
    rows = [pixels[r * w:(r + 1) * w] for r in range(h)]
    order = [p for row in rows_bottom_up(rows, top_down) for p in row]
    if rle:  # packets ignore row ends: runs of equal pixels (<= 128) and literal packets of varying length
        body, i, k = b"", 0, 0
        while i < len(order):
            run = 1
            while i + run < len(order) and run < 128 and order[i + run] == order[i]:
                run += 1
            if run >= 2:
                body += bytes([0x80 | (run - 1)]) + order[i]
                i += run
            else:
                n = min(len(order) - i, 1 + (k * 7) % 23)
                k += 1
                body += bytes([n - 1]) + b"".join(order[i:i + n])
                i += n
    else:
        body = b"".join(order)
    type_code = {"indexed": 1, "rgb": 2, "grey": 3}[kind] + (8 if rle else 0)
    desc = (0x20 if top_down else 0) | (0x10 if right_to_left else 0) | (8 if bits == 32 else 0)
    head = struct.pack("<BBBHHBHHHHBB", len(image_id), 1 if kind == "indexed" else 0, type_code, pal_start, pal_n, pal_bits, 0, 0, w, h, idx_bits if kind == "indexed" else bits, desc)
    return head + image_id + pal + body


CASES = [
    ("bmp24_hdr40_33x17", "bmp", lambda: bmp(33, 17, 24, 1)),
    ("bmp24_topdown_18x9", "bmp", lambda: bmp(18, 9, 24, 2, top_down=True)),
    ("bmp24_hdr124_7x5", "bmp", lambda: bmp(7, 5, 24, 3, hsz=124)),
    ("bmp24_os2_hdr12_21x6", "bmp", lambda: bmp(21, 6, 24, 4, os2=True)),
    ("bmp8_pal256_37x11", "bmp", lambda: bmp(37, 11, 8, 5)),
    ("bmp8_pal16_os2_10x10", "bmp", lambda: bmp(10, 10, 8, 6, palette_n=16, os2=True, index_n=12)),  # stb_image sizes an OS/2 palette as (offset - 38) / 3: 12 of the 16 entries; the rest would be uninitialised memory there
    ("bmp4_pal16_19x7", "bmp", lambda: bmp(19, 7, 4, 7)),
    ("bmp1_pal2_27x9", "bmp", lambda: bmp(27, 9, 1, 8)),
    ("bmp16_555_default_23x8", "bmp", lambda: bmp(23, 8, 16, 9)),
    ("bmp16_565_bitfields_14x13", "bmp", lambda: bmp(14, 13, 16, 10, compress=3, masks=(0xF800, 0x07E0, 0x001F, 0))),
    ("bmp16_4444_bitfields_v4_9x9", "bmp", lambda: bmp(9, 9, 16, 11, hsz=108, compress=3, masks=(0x0F00, 0x00F0, 0x000F, 0xF000))),
    ("bmp32_default_12x10", "bmp", lambda: bmp(12, 10, 32, 12)),
    ("bmp32_rgba_masks_v4_11x6", "bmp", lambda: bmp(11, 6, 32, 13, hsz=108, compress=3, masks=(0x000000FF, 0x0000FF00, 0x00FF0000, 0xFF000000))),
    ("bmp32_bitfields_hdr40_332_8x8", "bmp", lambda: bmp(8, 8, 32, 14, compress=3, masks=(0xE0000000, 0x001C0000, 0x00000300, 0))),
    ("bmp32_v5_default_5x4", "bmp", lambda: bmp(5, 4, 32, 15, hsz=124)),
    ("bmp24_gap_before_pixels_13x4", "bmp", lambda: bmp(13, 4, 24, 16, gap=8)),
    ("bmp8_pal200_gap_15x5", "bmp", lambda: bmp(15, 5, 8, 17, palette_n=200, gap=12)),
    ("bmp24_hdr56_6x6", "bmp", lambda: bmp(6, 6, 24, 18, hsz=56)),
    ("bmp8_rle_refused_8x8", "bmp", lambda: bmp(8, 8, 8, 19, compress=1)),
    ("tga24_bottomup_31x13", "tga", lambda: tga(31, 13, "rgb", 24, 20)),
    ("tga32_topdown_16x9", "tga", lambda: tga(16, 9, "rgb", 32, 21, top_down=True)),
    ("tga24_rle_29x14", "tga", lambda: tga(29, 14, "rgb", 24, 22, rle=True)),
    ("tga32_rle_topdown_17x17", "tga", lambda: tga(17, 17, "rgb", 32, 23, rle=True, top_down=True)),
    ("tga8_grey_20x11", "tga", lambda: tga(20, 11, "grey", 8, 24)),
    ("tga8_grey_rle_33x5", "tga", lambda: tga(33, 5, "grey", 8, 25, rle=True)),
    ("tga16_greyalpha_9x9", "tga", lambda: tga(9, 9, "grey", 16, 26)),
    ("tga16_555_12x7", "tga", lambda: tga(12, 7, "rgb", 16, 27)),
    ("tga15_555_rle_21x6", "tga", lambda: tga(21, 6, "rgb", 15, 28, rle=True)),
    ("tga_indexed8_pal24_25x10", "tga", lambda: tga(25, 10, "indexed", 0, 29, pal_bits=24, pal_n=64)),
    ("tga_indexed8_pal32_rle_14x14", "tga", lambda: tga(14, 14, "indexed", 0, 30, rle=True, pal_bits=32, pal_n=17)),
    ("tga_indexed16_pal16_10x6", "tga", lambda: tga(10, 6, "indexed", 0, 31, pal_bits=16, pal_n=300, idx_bits=16)),
    ("tga_indexed8_pal15_badindex_8x8", "tga", lambda: tga(8, 8, "indexed", 0, 32, pal_bits=15, pal_n=40, bad_index=True)),
    ("tga_indexed8_pal8_grey_7x7", "tga", lambda: tga(7, 7, "indexed", 0, 33, pal_bits=8, pal_n=50)),
    ("tga24_imageid_righttoleft_11x4", "tga", lambda: tga(11, 4, "rgb", 24, 34, image_id=b"made by make_golden_images", right_to_left=True)),
    ("tga24_1x1", "tga", lambda: tga(1, 1, "rgb", 24, 35)),
]


def more_cases():
    """GIF (first frame) and binary PNM files from the byte-level writers of tools/fuzz_gif_pnm.py, fixed seeds."""
    import importlib.util, random
    spec = importlib.util.spec_from_file_location("fuzz_gif_pnm", ROOT / "tools" / "fuzz_gif_pnm.py")
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    out = []
    for k in range(12):
        rnd, rng = random.Random(500 + k), np.random.default_rng(500 + k)
        out.append((f"gif_handmade_{k}", "gif", (lambda d: (lambda: d))(fz.handmade_gif(rng, rnd))))
    for k in range(8):
        rnd, rng = random.Random(900 + k), np.random.default_rng(900 + k)
        data = fz.pnm(rng, rnd)
        out.append((f"pnm_{k}", "pgm" if data[:2] == b"P5" else "ppm", (lambda d: (lambda: d))(data)))
    return out


def main():
    CASES.extend(more_cases())
    if not STB.exists():
        raise SystemExit("oracle/_ref/ref_stb missing: make -C oracle _ref/ref_stb (needs /root/reference)")
    OUT.mkdir(parents=True, exist_ok=True)
    refused = {}
    for name, ext, make in CASES:
        data = make()
        (OUT / f"{name}.{ext}").write_bytes(data)
        with tempfile.TemporaryDirectory() as td:
            raw = Path(td) / "o.raw"
            r = subprocess.run([str(STB), str(OUT / f"{name}.{ext}"), str(raw), "3"], capture_output=True, text=True)
            if r.returncode != 0:
                refused[f"{name}.{ext}"] = r.stderr.strip()
                print(name, "REFUSED:", r.stderr.strip())
                continue
            (OUT / f"{name}.raw.gz").write_bytes(gzip.compress(raw.read_bytes(), 9, mtime=0))
        print(name, len(data), "bytes")
    (OUT / "refused.json").write_text(json.dumps(refused, indent=1) + "\n")


if __name__ == "__main__":
    main()
