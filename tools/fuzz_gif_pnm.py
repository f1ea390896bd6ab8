#!/usr/bin/env python3
"""Random GIF files (Pillow: palettes of 2 .. 256 colours, transparency, interlacing; and hand-written ones: a frame smaller than the screen,
background index, local colour tables, graphic-control extensions, comment blocks, LZW with and without early clear codes) and binary PNM
files (P5 / P6, 8- and 16-bit, comments and odd whitespace in the header) decoded by the repo's loader and by the reference's own decoder
(oracle/_ref/ref_stb, three requested channels; build container only).  tools/fuzz_gif_pnm.py [cases]"""
import base64, io, json, random, struct, subprocess, sys, tempfile
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
         "images": [{"uri": "data:image/x;base64," + base64.b64encode(image_bytes).decode()}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC2"}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 24}],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    p = tmp / "x.gltf"; p.write_text(json.dumps(g)); return p


def lzw(indices, min_code, rnd):
    """GIF LZW: variable-width codes, LSB first, a clear code first, optional extra clear codes, 255-byte sub-blocks."""
    clear, end = 1 << min_code, (1 << min_code) + 1
    out_bits, nbits, data = 0, 0, bytearray()

    def put(code, size):
        nonlocal out_bits, nbits
        out_bits |= code << nbits; nbits += size
        while nbits >= 8:
            data.append(out_bits & 255); out_bits >>= 8; nbits -= 8
    table = {(i,): i for i in range(clear)}
    size, nxt = min_code + 1, end + 1
    put(clear, size)
    cur = ()
    for k, sym in enumerate(indices):
        if cur + (sym,) in table:
            cur = cur + (sym,)
            continue
        put(table[cur], size)
        if nxt < 4096:
            table[cur + (sym,)] = nxt
            if nxt == (1 << size) and size < 12:
                size += 1
            nxt += 1
        if nxt >= 4096 or rnd.random() < 0.002:
            put(clear, size)
            table = {(i,): i for i in range(clear)}
            size, nxt = min_code + 1, end + 1
        cur = (sym,)
    if cur:
        put(table[cur], size)
    put(end, size)
    if nbits:
        data.append(out_bits & 255)
    blocks = b""
    for i in range(0, len(data), 255):
        chunk = bytes(data[i:i + 255])
        blocks += bytes([len(chunk)]) + chunk
    return bytes([min_code]) + blocks + b"\0"


def handmade_gif(rng, rnd):
    W, H = rnd.randint(1, 40), rnd.randint(1, 30)
    bits = rnd.randint(1, 8)
    n = 1 << bits
    gpal = bytes(rng.integers(0, 256, n * 3, dtype=np.uint8))
    has_global = rnd.random() < 0.8
    bg = rnd.randrange(n) if rnd.random() < 0.7 else 0
    out = b"GIF89a" if rnd.random() < 0.7 else b"GIF87a"
    out += struct.pack("<HHBBB", W, H, (0x80 | (bits - 1)) if has_global else 0, bg, 0)
    if has_global:
        out += gpal
    if rnd.random() < 0.3:
        out += b"\x21\xFE" + bytes([5]) + b"hello" + b"\0"
    transparent = None
    if rnd.random() < 0.5:
        transparent = rnd.randrange(n)
        out += b"\x21\xF9\x04" + bytes([1 | (rnd.randrange(4) << 2)]) + struct.pack("<H", 7) + bytes([transparent]) + b"\0"
    elif rnd.random() < 0.3:
        out += b"\x21\xF9\x04" + bytes([0]) + struct.pack("<H", 0) + bytes([rnd.randrange(256)]) + b"\0"
    x, y = rnd.randint(0, W - 1), rnd.randint(0, H - 1)
    w, h = rnd.randint(1, W - x), rnd.randint(1, H - y)
    if rnd.random() < 0.5:
        x, y, w, h = 0, 0, W, H
    local = (not has_global) or rnd.random() < 0.3
    lbits = rnd.randint(1, 8) if local else bits
    interlace = rnd.random() < 0.4
    out += b"\x2C" + struct.pack("<HHHHB", x, y, w, h, (0x80 | (lbits - 1) if local else 0) | (0x40 if interlace else 0))
    if local:
        out += bytes(rng.integers(0, 256, (1 << lbits) * 3, dtype=np.uint8))
    ncol = 1 << lbits
    idx = rng.integers(0, ncol, w * h)
    if rnd.random() < 0.5:
        idx = np.repeat(rng.integers(0, ncol, (w * h + 5) // 6), 6)[:w * h]  # runs: longer LZW strings
    out += lzw([int(v) for v in idx], max(2, lbits), rnd)
    out += b"\x3B"
    return out


def pnm(rng, rnd):
    w, h = rnd.randint(1, 40), rnd.randint(1, 30)
    comp = rnd.choice([1, 3])
    maxv = rnd.choice([255, 255, 100, 1, 65535, 1000, 256])
    ws = lambda: rnd.choice([b" ", b"\n", b"\t", b"\r\n", b"  \n", b"\n# a comment\n", b" #c\r"])
    head = (b"P5" if comp == 1 else b"P6") + ws() + str(w).encode() + ws() + str(h).encode() + ws() + str(maxv).encode() + rnd.choice([b"\n", b" ", b"\t"])
    n = w * h * comp * (2 if maxv > 255 else 1)
    return head + bytes(rng.integers(0, 256, n, dtype=np.uint8))


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    rnd = random.Random(21)
    bad = 0
    accepted = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for it in range(cases):
            rng = np.random.default_rng(it)
            kind = it % 3
            if kind == 0:
                data, desc = handmade_gif(rng, rnd), "handmade gif"
            elif kind == 1:
                w, h = rnd.randint(1, 60), rnd.randint(1, 40)
                ncol = rnd.choice([2, 3, 4, 16, 37, 256])
                im = Image.fromarray(rng.integers(0, ncol, (h, w), dtype=np.uint8), "P")
                im.putpalette(bytes(rng.integers(0, 256, ncol * 3, dtype=np.uint8)))
                kw = {}
                if rnd.random() < 0.5:
                    kw["transparency"] = rnd.randrange(ncol)
                if rnd.random() < 0.5:
                    kw["interlace"] = 1
                buf = io.BytesIO(); im.save(buf, "GIF", **kw); data, desc = buf.getvalue(), f"pillow gif {w}x{h} {ncol} {kw}"
            else:
                data, desc = pnm(rng, rnd), "pnm"
            f = td / "x.bin"; f.write_bytes(data); raw = td / "o.raw"
            r = subprocess.run([str(STB), str(f), str(raw), "3"], capture_output=True, text=True)
            tex = ptb200.load_scene_file(gltf(td, data)).textures[0]
            key = desc.split()[0] + " " + desc.split()[1] if " " in desc else desc
            accepted.setdefault(key, [0, 0])[0 if r.returncode == 0 else 1] += 1
            if r.returncode != 0:
                if tex.shape[0] != 0:
                    bad += 1; print("stb refused, we decoded:", it, desc, r.stderr.strip())
                continue
            head, body = raw.read_bytes().split(b"\n", 1)
            W, H, _ = map(int, head.split())
            ref = np.frombuffer(body, np.uint8).reshape(H, W, 3)
            if tex.shape != (H, W, 3) or not np.array_equal(tex, ref.astype(np.float32)):
                bad += 1
                print("MISMATCH", it, desc, tex.shape, (H, W), int((tex != ref).sum()) if tex.shape == (H, W, 3) else "")
    print("cases", cases, "bad", bad, "decoded / refused by stb_image:", accepted)


if __name__ == "__main__":
    main()
