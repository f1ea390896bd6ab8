// BmpTgaDecoder.h — BMP, TGA, GIF and PNM textures -> RGB8, with the texels the reference gets.
//
// The reference reads texture FILES with stbi_load(path, &w, &h, &n, 3) (src/HostScene.cpp:29), i.e. through the vendored
// third-party/stb_image.h (v2.30) with three requested channels.  glTF itself only carries PNG / JPEG, so BMP and TGA reach the
// loader through `map_Kd` of an .mtl or a `uri` of a .gltf.  Both formats are simple containers; what is restated here are stb_image's
// CHOICES, because they decide bytes:
//   BMP  header sizes 12 / 40 / 56 / 108 / 124; 1 / 4 / 8-bit palettes, 16-bit (5-5-5 unless BI_BITFIELDS), 24-bit, 32-bit (masks);
//        a channel of n <= 8 bits is widened by bit replication; RLE and embedded PNG / JPEG are refused; bottom-up unless the
//        height is negative; the gap between header and pixels is skipped the (quirky) way stb_image skips it;
//   TGA  types 1 / 2 / 3 and their RLE forms 9 / 10 / 11; 8-bit grey, 15 / 16-bit as 5-5-5 scaled by (v * 255) / 31 with the top bit
//        ignored, 24 / 32-bit BGR(A); palettes of 8 / 15 / 16 / 24 / 32 bits with 8- or 16-bit indices (an index past the palette
//        reads entry 0); bottom-up unless bit 5 of the descriptor is set, bit 4 (right-to-left) is ignored; no magic number — the
//        same plausibility test decides whether a file is a TGA at all.
// Three channels out: grey is replicated, alpha is dropped (stbi__convert_format).  A read past the end of the data yields zeros, as
// stb_image's memory reader does.  Pinned byte for byte against stb_image itself: oracle/make_golden_images.py -> tests/golden/images/.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace ptimg {

struct Reader {
    const unsigned char *p;
    size_t n, pos = 0;
    Reader(const unsigned char *bytes, size_t len) : p(bytes), n(len) {}
    int u8() { return pos < n ? p[pos++] : (pos++, 0); }
    int u16() { int a = u8(); return a | (u8() << 8); }
    uint32_t u32() { uint32_t a = (uint32_t)u16(); return a | ((uint32_t)u16() << 16); }
    void skip(long k) {
        if (k < 0) { pos = n; return; }  // stb_image: a negative skip parks the reader at the end
        pos += (size_t)k;
    }
};

constexpr int kMaxDimension = 1 << 24;  // STBI_MAX_DIMENSIONS

inline int high_bit(uint32_t z) {
    if (z == 0) return -1;
    int k = 0;
    while (z >>= 1) k++;
    return k;
}
inline int bit_count(uint32_t a) {
    int c = 0;
    for (; a; a &= a - 1) c++;
    return c;
}
// the `bits` top bits of the masked value (moved so that the mask's highest bit is bit 7), widened to 8 bits by repeating the pattern
inline int widen(uint32_t v, int shift, int bits) {
    static const unsigned mul[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
    static const unsigned shr[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
    v = shift < 0 ? v << -shift : v >> shift;
    v >>= (8 - bits);
    return (int)((v * mul[bits]) >> shr[bits]);
}

inline bool is_bmp(const unsigned char *b, size_t n) {
    if (n < 18 || b[0] != 'B' || b[1] != 'M') return false;
    const uint32_t sz = (uint32_t)b[14] | ((uint32_t)b[15] << 8) | ((uint32_t)b[16] << 16) | ((uint32_t)b[17] << 24);
    return sz == 12 || sz == 40 || sz == 56 || sz == 108 || sz == 124;
}

inline bool decode_bmp(const unsigned char *bytes, size_t len, int &w, int &h, std::vector<unsigned char> &rgb, std::string &err) {
    Reader s(bytes, len);
    if (s.u8() != 'B' || s.u8() != 'M') { err = "not BMP"; return false; }
    s.u32(); s.u16(); s.u16();
    const int offset = (int)s.u32();
    const int hsz = (int)s.u32();
    uint32_t mr = 0, mg = 0, mb = 0, ma = 0;
    int extra_read = 14;
    if (offset < 0) { err = "bad BMP"; return false; }
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) { err = "BMP type not supported: unknown"; return false; }
    int img_x, img_y;
    if (hsz == 12) { img_x = s.u16(); img_y = s.u16(); }
    else { img_x = (int)s.u32(); img_y = (int)s.u32(); }
    if (s.u16() != 1) { err = "bad BMP"; return false; }
    const int bpp = s.u16();
    auto default_masks = [&]() {
        if (bpp == 16) { mr = 31u << 10; mg = 31u << 5; mb = 31u; }
        else if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; }
        else mr = mg = mb = ma = 0;
    };
    if (hsz != 12) {
        const int compress = (int)s.u32();
        if (compress == 1 || compress == 2) { err = "BMP type not supported: RLE"; return false; }
        if (compress >= 4 || compress < 0) { err = "BMP type not supported: unsupported compression"; return false; }
        if (compress == 3 && bpp != 16 && bpp != 32) { err = "bad BMP"; return false; }
        for (int i = 0; i < 5; i++) s.u32();
        if (hsz == 40 || hsz == 56) {
            if (hsz == 56) for (int i = 0; i < 4; i++) s.u32();
            if (bpp == 16 || bpp == 32) {
                if (compress == 0) default_masks();
                else if (compress == 3) {
                    mr = s.u32(); mg = s.u32(); mb = s.u32();
                    extra_read += 12;
                    if (mr == mg && mg == mb) { err = "bad BMP"; return false; }
                } else { err = "bad BMP"; return false; }
            }
        } else {
            mr = s.u32(); mg = s.u32(); mb = s.u32(); ma = s.u32();
            if (compress != 3) default_masks();
            s.u32();
            for (int i = 0; i < 12; i++) s.u32();
            if (hsz == 124) for (int i = 0; i < 4; i++) s.u32();
        }
    }
    const bool flip = img_y > 0;
    if (img_y == INT32_MIN) { err = "Very large image (corrupt?)"; return false; }
    img_y = std::abs(img_y);
    if (img_x > kMaxDimension || img_y > kMaxDimension || img_x < 0) { err = "Very large image (corrupt?)"; return false; }
    int psize = 0;
    if (hsz == 12) { if (bpp < 24) psize = (offset - extra_read - 24) / 3; }
    else if (bpp < 16) psize = (offset - extra_read - hsz) >> 2;
    if (psize == 0) {
        const long so_far = (long)s.pos;
        if (so_far <= 0 || so_far > 1024) { err = "Corrupt BMP"; return false; }
        if (offset < so_far || offset - so_far > 1024) { err = "Corrupt BMP"; return false; }
        s.skip(offset - so_far);
    }
    if ((uint64_t)img_x * (uint64_t)img_y * 3u > 0x7fffffffull) { err = "Corrupt BMP"; return false; }
    rgb.assign((size_t)img_x * (size_t)img_y * 3, 0);
    size_t z = 0;
    if (bpp < 16) {
        if (psize == 0 || psize > 256) { err = "Corrupt BMP"; return false; }
        unsigned char pal[256][3] = {};
        for (int i = 0; i < psize; i++) {
            pal[i][2] = (unsigned char)s.u8();
            pal[i][1] = (unsigned char)s.u8();
            pal[i][0] = (unsigned char)s.u8();
            if (hsz != 12) s.u8();
        }
        s.skip((long)offset - extra_read - hsz - (long)psize * (hsz == 12 ? 3 : 4));
        int width;
        if (bpp == 1) width = (img_x + 7) >> 3;
        else if (bpp == 4) width = (img_x + 1) >> 1;
        else if (bpp == 8) width = img_x;
        else { err = "Corrupt BMP"; return false; }
        const int pad = (-width) & 3;
        auto put = [&](int c) { rgb[z++] = pal[c][0]; rgb[z++] = pal[c][1]; rgb[z++] = pal[c][2]; };
        for (int j = 0; j < img_y; j++) {
            if (bpp == 1) {
                int bit = 7, v = s.u8();
                for (int i = 0; i < img_x; i++) {
                    put((v >> bit) & 1);
                    if (i + 1 == img_x) break;
                    if (--bit < 0) { bit = 7; v = s.u8(); }
                }
            } else {
                for (int i = 0; i < img_x; i += 2) {
                    int v = s.u8(), v2 = 0;
                    if (bpp == 4) { v2 = v & 15; v >>= 4; }
                    put(v);
                    if (i + 1 == img_x) break;
                    put(bpp == 8 ? s.u8() : v2);
                }
            }
            s.skip(pad);
        }
    } else {
        s.skip((long)offset - extra_read - hsz);
        int width = bpp == 24 ? 3 * img_x : bpp == 16 ? 2 * img_x : 0;
        const int pad = (-width) & 3;
        int easy = 0;
        if (bpp == 24) easy = 1;
        else if (bpp == 32 && mb == 0xffu && mg == 0xff00u && mr == 0x00ff0000u && ma == 0xff000000u) easy = 2;
        int rs = 0, gs = 0, bs = 0, rc = 0, gc = 0, bc = 0;
        if (!easy) {
            if (!mr || !mg || !mb) { err = "Corrupt BMP"; return false; }
            rs = high_bit(mr) - 7; rc = bit_count(mr);
            gs = high_bit(mg) - 7; gc = bit_count(mg);
            bs = high_bit(mb) - 7; bc = bit_count(mb);
            if (rc > 8 || gc > 8 || bc > 8 || bit_count(ma) > 8) { err = "Corrupt BMP"; return false; }
        }
        for (int j = 0; j < img_y; j++) {
            for (int i = 0; i < img_x; i++) {
                if (easy) {
                    rgb[z + 2] = (unsigned char)s.u8();
                    rgb[z + 1] = (unsigned char)s.u8();
                    rgb[z + 0] = (unsigned char)s.u8();
                    z += 3;
                    if (easy == 2) s.u8();
                } else {
                    const uint32_t v = bpp == 16 ? (uint32_t)s.u16() : s.u32();
                    rgb[z++] = (unsigned char)widen(v & mr, rs, rc);
                    rgb[z++] = (unsigned char)widen(v & mg, gs, gc);
                    rgb[z++] = (unsigned char)widen(v & mb, bs, bc);
                }
            }
            s.skip(pad);
        }
    }
    if (flip) {
        const size_t row = (size_t)img_x * 3;
        for (int j = 0; j < img_y >> 1; j++)
            for (size_t i = 0; i < row; i++) std::swap(rgb[(size_t)j * row + i], rgb[(size_t)(img_y - 1 - j) * row + i]);
    }
    w = img_x;
    h = img_y;
    return true;
}

// stb_image's plausibility test (TGA has no signature): colour-map type 0 / 1, a matching image type, sane sizes and depths
inline bool is_tga(const unsigned char *bytes, size_t len) {
    Reader s(bytes, len);
    s.u8();
    const int color_type = s.u8();
    if (color_type > 1) return false;
    int sz = s.u8();
    if (color_type == 1) {
        if (sz != 1 && sz != 9) return false;
        s.skip(4);
        sz = s.u8();
        if (sz != 8 && sz != 15 && sz != 16 && sz != 24 && sz != 32) return false;
        s.skip(4);
    } else {
        if (sz != 2 && sz != 3 && sz != 10 && sz != 11) return false;
        s.skip(9);
    }
    if (s.u16() < 1) return false;
    if (s.u16() < 1) return false;
    sz = s.u8();
    if (color_type == 1 && sz != 8 && sz != 16) return false;
    if (sz != 8 && sz != 15 && sz != 16 && sz != 24 && sz != 32) return false;
    return true;
}

inline bool decode_tga(const unsigned char *bytes, size_t len, int &w, int &h, std::vector<unsigned char> &rgb, std::string &err) {
    Reader s(bytes, len);
    const int id_len = s.u8();
    const int indexed = s.u8();
    int type = s.u8();
    const int pal_start = s.u16();
    const int pal_len = s.u16();
    const int pal_bits = s.u8();
    s.u16(); s.u16();  // x / y origin: unused
    const int width = s.u16(), height = s.u16();
    const int bpp = s.u8();
    const int descriptor = s.u8();
    bool rle = false;
    if (type >= 8) { type -= 8; rle = true; }
    const bool bottom_up = ((descriptor >> 5) & 1) == 0;
    // channels of a pixel (or palette entry): 8 -> grey, 16 -> grey + alpha when the image is grey, else 5-5-5; 15 -> 5-5-5; 24 / 32 -> BGR(A)
    auto comp_of = [](int bits, bool grey, bool &rgb16) {
        rgb16 = false;
        switch (bits) {
            case 8: return 1;
            case 16: if (grey) return 2;  // fall through
            case 15: rgb16 = true; return 3;
            case 24: return 3;
            case 32: return 4;
            default: return 0;
        }
    };
    bool rgb16 = false;
    const int comp = indexed ? comp_of(pal_bits, false, rgb16) : comp_of(bpp, type == 3, rgb16);
    if (!comp) { err = "Can't find out TGA pixelformat"; return false; }
    if (width > kMaxDimension || height > kMaxDimension) { err = "Very large image (corrupt?)"; return false; }
    if ((uint64_t)width * (uint64_t)height * 4u > 0x7fffffffull) { err = "Corrupt TGA"; return false; }
    std::vector<unsigned char> data((size_t)width * (size_t)height * (size_t)comp, 0);
    s.skip(id_len);
    auto read555 = [&](unsigned char *out) {
        const int px = s.u16();
        out[0] = (unsigned char)((((px >> 10) & 31) * 255) / 31);
        out[1] = (unsigned char)((((px >> 5) & 31) * 255) / 31);
        out[2] = (unsigned char)(((px & 31) * 255) / 31);
    };
    std::vector<unsigned char> palette;
    if (indexed) {
        if (pal_len == 0) { err = "Corrupt TGA"; return false; }
        s.skip(pal_start);
        palette.assign((size_t)pal_len * (size_t)comp, 0);
        if (rgb16) {
            for (int i = 0; i < pal_len; i++) read555(&palette[(size_t)i * 3]);
        } else {
            if (s.pos + palette.size() > s.n) { err = "Corrupt TGA"; return false; }  // a short palette is the one short read stb_image refuses
            for (auto &b : palette) b = (unsigned char)s.u8();
        }
    }
    unsigned char raw[4] = {0, 0, 0, 0};
    int run = 0;
    bool repeating = false, read_next = true;
    const size_t n_px = (size_t)width * (size_t)height;
    for (size_t i = 0; i < n_px; i++) {
        if (rle) {
            if (run == 0) {
                const int cmd = s.u8();
                run = 1 + (cmd & 127);
                repeating = (cmd >> 7) != 0;
                read_next = true;
            } else if (!repeating) {
                read_next = true;
            }
        } else {
            read_next = true;
        }
        if (read_next) {
            if (indexed) {
                int idx = bpp == 8 ? s.u8() : s.u16();
                if (idx >= pal_len) idx = 0;
                for (int j = 0; j < comp; j++) raw[j] = palette[(size_t)idx * (size_t)comp + (size_t)j];
            } else if (rgb16) {
                read555(raw);
            } else {
                for (int j = 0; j < comp; j++) raw[j] = (unsigned char)s.u8();
            }
            read_next = false;
        }
        for (int j = 0; j < comp; j++) data[i * (size_t)comp + (size_t)j] = raw[j];
        --run;
    }
    if (bottom_up) {
        const size_t row = (size_t)width * (size_t)comp;
        for (int j = 0; j * 2 < height; j++)
            for (size_t i = 0; i < row; i++) std::swap(data[(size_t)j * row + i], data[(size_t)(height - 1 - j) * row + i]);
    }
    rgb.assign(n_px * 3, 0);
    for (size_t i = 0; i < n_px; i++) {
        const unsigned char *px = &data[i * (size_t)comp];
        if (comp <= 2) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = px[0];             // grey (+ alpha): replicated
        else if (rgb16) { rgb[3 * i] = px[0]; rgb[3 * i + 1] = px[1]; rgb[3 * i + 2] = px[2]; }  // 5-5-5 was read as R, G, B
        else { rgb[3 * i] = px[2]; rgb[3 * i + 1] = px[1]; rgb[3 * i + 2] = px[0]; }      // BGR(A) in the file
    }
    w = width;
    h = height;
    return true;
}

// ---- GIF (first frame) and binary PNM, the way stb_image returns them for three requested channels ------------------------------------------
//   GIF  87a / 89a, first image only; pixels the frame does not draw stay (0,0,0) unless the screen's background index is > 0, in which case
//        they get that palette entry — copied as stored, i.e. with red and blue exchanged (stb_image's palette is B,G,R,A and the copy is raw);
//        a transparent index (graphic control extension) is "drawn" as nothing; interlaced rows in 8-8-4-2 order; LZW with at most 4096
//        codes of <= 12 bits, deferred clear codes allowed, a stream that does not start with a clear code refused;
//   PNM  P5 / P6 only, no comments inside numbers, one whitespace byte before the samples, maxval NOT used for scaling; 16-bit samples
//        (maxval > 255) lose their HIGH byte: stb_image reads big-endian pairs into native (little-endian) words and keeps `>> 8`.
inline bool is_gif(const unsigned char *b, size_t n) {
    return n >= 6 && b[0] == 'G' && b[1] == 'I' && b[2] == 'F' && b[3] == '8' && (b[4] == '7' || b[4] == '9') && b[5] == 'a';
}

inline bool decode_gif(const unsigned char *bytes, size_t len, int &w, int &h, std::vector<unsigned char> &rgb, std::string &err) {
    Reader s(bytes, len);
    if (!is_gif(bytes, len)) { err = "Corrupt GIF"; return false; }
    s.skip(6);
    const int W = s.u16(), H = s.u16();
    const int flags = s.u8(), bgindex = s.u8();
    s.u8();  // aspect ratio
    if ((uint64_t)W * (uint64_t)H * 4u > 0x7fffffffull) { err = "GIF image is too large"; return false; }
    unsigned char pal[256][4] = {}, lpal[256][4] = {};  // stored B, G, R, A
    auto read_table = [&](unsigned char (*t)[4], int n, int transp) {
        for (int i = 0; i < n; i++) {
            t[i][2] = (unsigned char)s.u8();
            t[i][1] = (unsigned char)s.u8();
            t[i][0] = (unsigned char)s.u8();
            t[i][3] = transp == i ? 0 : 255;
        }
    };
    if (flags & 0x80) read_table(pal, 2 << (flags & 7), -1);
    const size_t pcount = (size_t)W * (size_t)H;
    std::vector<unsigned char> out(pcount * 4, 0), history(pcount, 0);
    int transparent = -1, eflags = 0;
    for (;;) {
        if (s.pos >= s.n + 16) { err = "Corrupt GIF"; return false; }  // reads past the end yield zeros = an unknown block code soon enough
        const int tag = s.u8();
        if (tag == 0x2C) {
            const int x = s.u16(), y = s.u16(), iw = s.u16(), ih = s.u16();
            if (x + iw > W || y + ih > H) { err = "Corrupt GIF"; return false; }
            const long line = (long)W * 4;
            const long start_x = (long)x * 4, start_y = (long)y * line, max_x = start_x + (long)iw * 4, max_y = start_y + (long)ih * line;
            long cur_x = start_x, cur_y = iw == 0 ? max_y : start_y, step;
            int parse;
            const int lflags = s.u8();
            if (lflags & 0x40) { step = 8 * line; parse = 3; } else { step = line; parse = 0; }
            const unsigned char (*table)[4];
            if (lflags & 0x80) { read_table(lpal, 2 << (lflags & 7), (eflags & 1) ? transparent : -1); table = lpal; }
            else if (flags & 0x80) table = pal;
            else { err = "Corrupt GIF"; return false; }
            // LZW raster
            const int lzw_cs = s.u8();
            if (lzw_cs > 12) { err = "Corrupt GIF"; return false; }
            struct Code { int16_t prefix; unsigned char first, suffix; };
            std::vector<Code> codes(8192);
            const int clear = 1 << lzw_cs;
            for (int i = 0; i < clear; i++) codes[(size_t)i] = {(int16_t)-1, (unsigned char)i, (unsigned char)i};
            bool first = true;
            int codesize = lzw_cs + 1, codemask = (1 << codesize) - 1, avail = clear + 2, oldcode = -1, valid_bits = 0, blk = 0;
            int32_t bits = 0;
            std::vector<uint16_t> chain;
            auto emit = [&](int code) {  // the code's string, first symbol first
                chain.clear();
                for (int c = code; c >= 0; c = codes[(size_t)c].prefix) {
                    chain.push_back((uint16_t)c);
                    if (chain.size() > 8192) break;
                }
                for (size_t k = chain.size(); k-- > 0;) {
                    if (cur_y >= max_y) continue;
                    const long idx = cur_x + cur_y;
                    history[(size_t)(idx / 4)] = 1;
                    const unsigned char *c = table[codes[chain[k]].suffix];
                    if (c[3] > 128) {
                        out[(size_t)idx] = c[2]; out[(size_t)idx + 1] = c[1]; out[(size_t)idx + 2] = c[0]; out[(size_t)idx + 3] = c[3];
                    }
                    cur_x += 4;
                    if (cur_x >= max_x) {
                        cur_x = start_x;
                        cur_y += step;
                        while (cur_y >= max_y && parse > 0) {
                            step = (1L << parse) * line;
                            cur_y = start_y + (step >> 1);
                            --parse;
                        }
                    }
                }
            };
            bool done = false;
            while (!done) {
                if (valid_bits < codesize) {
                    if (blk == 0) {
                        blk = s.u8();
                        if (blk == 0) break;  // block terminator: the raster ends here
                    }
                    --blk;
                    bits |= (int32_t)s.u8() << valid_bits;
                    valid_bits += 8;
                    if (s.pos > s.n + 1024) { err = "Corrupt GIF"; return false; }
                } else {
                    const int code = bits & codemask;
                    bits >>= codesize;
                    valid_bits -= codesize;
                    if (code == clear) {
                        codesize = lzw_cs + 1; codemask = (1 << codesize) - 1; avail = clear + 2; oldcode = -1; first = false;
                    } else if (code == clear + 1) {
                        s.skip(blk);
                        while ((blk = s.u8()) > 0) { s.skip(blk); if (s.pos > s.n + 1024) break; }
                        done = true;
                    } else if (code <= avail) {
                        if (first) { err = "Corrupt GIF"; return false; }
                        if (oldcode >= 0) {
                            Code &p = codes[(size_t)avail++];
                            if (avail > 8192) { err = "Corrupt GIF"; return false; }
                            p.prefix = (int16_t)oldcode;
                            p.first = codes[(size_t)oldcode].first;
                            p.suffix = (code == avail) ? p.first : codes[(size_t)code].first;
                        } else if (code == avail) { err = "Corrupt GIF"; return false; }
                        emit(code);
                        if ((avail & codemask) == 0 && avail <= 0x0FFF) { codesize++; codemask = (1 << codesize) - 1; }
                        oldcode = code;
                    } else { err = "Corrupt GIF"; return false; }
                }
            }
            if (bgindex > 0)
                for (size_t pi = 0; pi < pcount; pi++)
                    if (!history[pi]) { out[pi * 4] = pal[bgindex][0]; out[pi * 4 + 1] = pal[bgindex][1]; out[pi * 4 + 2] = pal[bgindex][2]; out[pi * 4 + 3] = 255; }
            break;
        } else if (tag == 0x21) {
            const int ext = s.u8();
            int n;
            if (ext == 0xF9) {
                n = s.u8();
                if (n == 4) {
                    eflags = s.u8();
                    s.u16();
                    if (transparent >= 0) pal[transparent][3] = 255;
                    if (eflags & 1) { transparent = s.u8(); pal[transparent][3] = 0; }
                    else { s.skip(1); transparent = -1; }
                } else { s.skip(n); continue; }
            }
            while ((n = s.u8()) != 0) { s.skip(n); if (s.pos > s.n + 1024) { err = "Corrupt GIF"; return false; } }
        } else if (tag == 0x3B) { err = "GIF without an image"; return false; }
        else { err = "Corrupt GIF"; return false; }
    }
    rgb.resize(pcount * 3);
    for (size_t i = 0; i < pcount; i++) { rgb[3 * i] = out[4 * i]; rgb[3 * i + 1] = out[4 * i + 1]; rgb[3 * i + 2] = out[4 * i + 2]; }
    w = W;
    h = H;
    return true;
}

inline bool is_pnm(const unsigned char *b, size_t n) { return n >= 2 && b[0] == 'P' && (b[1] == '5' || b[1] == '6'); }

inline bool decode_pnm(const unsigned char *bytes, size_t len, int &w, int &h, std::vector<unsigned char> &rgb, std::string &err) {
    if (!is_pnm(bytes, len)) { err = "not PNM"; return false; }
    Reader s(bytes, len);
    s.skip(1);
    const int comp = s.u8() == '6' ? 3 : 1;
    int c = s.u8();
    auto at_eof = [&]() { return s.pos >= s.n; };
    auto is_space = [](int ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\v' || ch == '\f' || ch == '\r'; };
    auto skip_ws = [&]() {
        for (;;) {
            while (!at_eof() && is_space(c)) c = s.u8();
            if (at_eof() || c != '#') break;
            while (!at_eof() && c != '\n' && c != '\r') c = s.u8();
        }
    };
    bool overflow = false;
    auto integer = [&]() {
        int value = 0;
        while (!at_eof() && c >= '0' && c <= '9') {
            value = value * 10 + (c - '0');
            c = s.u8();
            if (value > 214748364 || (value == 214748364 && c > '7')) { overflow = true; return 0; }
        }
        return value;
    };
    skip_ws();
    const int W = integer();
    if (W == 0 || overflow) { err = "PPM image header had zero or overflowing width"; return false; }
    skip_ws();
    const int H = integer();
    if (H == 0 || overflow) { err = "PPM image header had zero or overflowing width"; return false; }
    skip_ws();
    const int maxv = integer();
    if (overflow || maxv > 65535) { err = "PPM image supports only 8-bit and 16-bit images"; return false; }
    const int bytes_per = maxv > 255 ? 2 : 1;
    if (W > kMaxDimension || H > kMaxDimension) { err = "Very large image (corrupt?)"; return false; }
    const uint64_t need = (uint64_t)W * (uint64_t)H * (uint64_t)comp * (uint64_t)bytes_per;
    if (need > 0x7fffffffull) { err = "PNM too large"; return false; }
    if (s.pos > s.n || need > s.n - s.pos) { err = "PNM file truncated"; return false; }
    const unsigned char *px = bytes + s.pos;
    const size_t n_px = (size_t)W * (size_t)H;
    rgb.resize(n_px * 3);
    for (size_t i = 0; i < n_px; i++)
        for (int k = 0; k < 3; k++) {
            const size_t sample = i * (size_t)comp + (size_t)(comp == 3 ? k : 0);
            rgb[3 * i + (size_t)k] = bytes_per == 1 ? px[sample] : px[sample * 2 + 1];  // 16-bit: the second byte of the pair (see above)
        }
    w = W;
    h = H;
    return true;
}

}  // namespace ptimg
