#!/usr/bin/env python3
"""Random PNG files (Pillow: grey / grey+alpha / RGB / RGBA / palette with and without transparency, 1 / 2 / 4 / 8 / 16 bits, random sizes and
compression levels) decoded by the repo's loader and by the reference's own decoder (oracle/_ref/ref_stb, build container only).  The reference
walks the decoded buffer of an embedded texture three bytes per texel whatever its channel count (src/HostScene.cpp:18-46); the expectation is
built the same way from stb_image's output.  tools/fuzz_png.py [cases]"""
import base64, io, json, random, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
from PIL import Image
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ptb200  # noqa: E402
STB = ROOT / "oracle" / "_ref" / "ref_stb"


def gltf(tmp, image_bytes):
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32); uv = np.array([[0, 0], [1, 0], [0, 1]], np.float32); blob = pos.tobytes() + uv.tobytes()
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}], "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "material": 0}]}],
         "materials": [{"name": "photo", "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}], "textures": [{"source": 0}],
         "images": [{"uri": "data:image/png;base64," + base64.b64encode(image_bytes).decode()}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC2"}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 24}],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    p = tmp / "x.gltf"; p.write_text(json.dumps(g)); return p


def handmade_png(rng, rnd, w, h):
    """A PNG written byte by byte: colour types 0 / 2 / 3 / 4 / 6, bit depths 1 .. 16 where the type allows, a random filter type on every
    row and, half of the time, Adam7 interlacing (Pillow writes neither)."""
    import struct, zlib
    ctype = rnd.choice([0, 2, 3, 4, 6])
    depth = rnd.choice({0: [1, 2, 4, 8, 16], 2: [8, 16], 3: [1, 2, 4, 8], 4: [8, 16], 6: [8, 16]}[ctype])
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    interlace = rnd.random() < 0.5
    hi = (1 << depth) if depth < 16 else 65536
    n_pal = rnd.randint(1, 1 << depth) if ctype == 3 else 0
    vals = rng.integers(0, n_pal if ctype == 3 else hi, (h, w, channels))

    def pack_row(px):  # px: [n][channels] -> bytes
        flat = px.reshape(-1)
        if depth == 16:
            return b"".join(struct.pack(">H", int(v)) for v in flat)
        if depth == 8:
            return bytes(int(v) for v in flat)
        out, acc, nb = bytearray(), 0, 0
        for v in flat:
            acc = (acc << depth) | int(v); nb += depth
            if nb == 8:
                out.append(acc); acc = nb = 0
        if nb:
            out.append(acc << (8 - nb))
        return bytes(out)

    bpp = max(1, channels * depth // 8)

    def filtered(rows):
        out, prev = bytearray(), None
        for row in rows:
            ft = rnd.randint(0, 4)
            cur = bytearray(row)
            enc = bytearray(len(cur))
            for i in range(len(cur)):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i] if prev is not None else 0
                c = prev[i - bpp] if (prev is not None and i >= bpp) else 0
                if ft == 0: pred = 0
                elif ft == 1: pred = a
                elif ft == 2: pred = b
                elif ft == 3: pred = (a + b) >> 1
                else:
                    pp = a + b - c; pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                enc[i] = (cur[i] - pred) & 255
            out.append(ft); out += enc
            prev = cur
        return bytes(out)
    if interlace:
        raw = b""
        for (x0, y0, dx, dy) in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]:
            sub = vals[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += filtered([pack_row(r) for r in sub])
    else:
        raw = filtered([pack_row(r) for r in vals])

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xffffffff)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if ctype == 3:
        png += chunk(b"PLTE", bytes(rng.integers(0, 256, n_pal * 3, dtype=np.uint8)))
        if rnd.random() < 0.4:
            png += chunk(b"tRNS", bytes(rng.integers(0, 256, rnd.randint(1, n_pal), dtype=np.uint8)))
    elif ctype in (0, 2) and rnd.random() < 0.3:
        png += chunk(b"tRNS", b"".join(struct.pack(">H", int(v)) for v in vals[h // 2, w // 2]))
    z = zlib.compress(raw, rnd.choice([0, 6, 9]))
    cut = rnd.randint(1, max(1, len(z) - 1)) if rnd.random() < 0.5 else len(z)  # the stream split over two IDAT chunks
    png += chunk(b"IDAT", z[:cut]) + (chunk(b"IDAT", z[cut:]) if cut < len(z) else b"")
    return png + chunk(b"IEND", b""), f"handmade type {ctype} depth {depth} interlace {interlace}"


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rnd = random.Random(13)
    bad = 0
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for it in range(cases):
            w, h = rnd.randint(1, 70), rnd.randint(1, 50)
            rng = np.random.default_rng(it)
            mode = rnd.choice(["L", "LA", "RGB", "RGBA", "P", "Pt", "1", "I;16", "L2", "L4", "hand", "hand", "hand", "hand"])
            kw = dict(compress_level=rnd.choice([0, 1, 6, 9]))
            data = None
            if mode == "hand":
                data, mode = handmade_png(rng, rnd, w, h)
            elif mode in ("L", "LA", "RGB", "RGBA"):
                a = rng.integers(0, 256, (h, w, len(mode)), dtype=np.uint8)
                im = Image.fromarray(np.ascontiguousarray(a[:, :, 0] if mode == "L" else a), mode)
            elif mode in ("P", "Pt"):
                n = rnd.choice([2, 4, 16, 37, 256])
                im = Image.fromarray(rng.integers(0, n, (h, w), dtype=np.uint8), "P")
                im.putpalette(bytes(rng.integers(0, 256, n * 3, dtype=np.uint8)))
                if mode == "Pt":
                    kw["transparency"] = bytes(rng.integers(0, 256, n, dtype=np.uint8))
                kw["bits"] = {2: 1, 4: 2, 16: 4}.get(n, 8)
            elif mode == "1":
                im = Image.fromarray((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8), "L").convert("1")
            elif mode == "I;16":
                im = Image.fromarray(rng.integers(0, 65536, (h, w), dtype=np.uint16), "I;16")
            else:
                bits = int(mode[1])
                im = Image.fromarray(rng.integers(0, 1 << bits, (h, w), dtype=np.uint8), "P")
                im.putpalette(bytes(sum(([v * 255 // ((1 << bits) - 1)] * 3 for v in range(1 << bits)), [])))
                kw["bits"] = bits
            if data is None:
                buf = io.BytesIO()
                im.save(buf, "PNG", **kw)
                data = buf.getvalue()
            f = td / "x.png"; f.write_bytes(data); raw = td / "o.raw"
            r = subprocess.run([str(STB), str(f), str(raw)], capture_output=True, text=True)
            if r.returncode != 0:
                print("stb refused", mode, w, h, r.stderr.strip()); bad += 1
                continue
            head, body = raw.read_bytes().split(b"\n", 1)
            W, H, C = map(int, head.split())
            flat = np.frombuffer(body, np.uint8)
            n_tex = min(W * H, len(flat) // 3)
            expect = np.zeros((W * H, 3), np.float32)
            expect[:n_tex] = flat[:n_tex * 3].reshape(n_tex, 3)
            tex = ptb200.load_scene_file(gltf(td, data)).textures[0]
            got = tex.reshape(-1, 3)
            if tex.shape != (H, W, 3) or not np.array_equal(got[:n_tex], expect[:n_tex]):
                bad += 1
                print("MISMATCH", mode, w, h, C, kw.get("bits"), tex.shape, int((got[:n_tex] != expect[:n_tex]).sum()) if tex.shape == (H, W, 3) else "")
    print("cases", cases, "bad", bad)


if __name__ == "__main__":
    main()
