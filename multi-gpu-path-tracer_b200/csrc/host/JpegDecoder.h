// JpegDecoder.h — Huffman JPEG (baseline, extended-sequential, progressive) -> 8-bit pixels, for SceneLoader's textures.
//
// The reference decodes textures with stb_image (third-party/stb_image.h through src/HostScene.cpp:10-51, un-modified public-domain
// code): a JPEG baseColor texture — the usual case of a photographic .glb — must come out with the SAME bytes, or every textured hit
// differs.  A JPEG's entropy decoding is fixed by the standard, but three steps are implementation-defined and are restated here the
// way stb_image does them (the arithmetic only, pinned byte for byte against stb_image itself by oracle/ref_stb.c and
// tests/test_host_logic.py):
//   1. dequantise to 16 bits, then the integer inverse DCT derived from jidctint (12-bit constants, 2 extra bits between the passes,
//      +128 folded into the rounding term)                                                     stb_image.h: stbi__idct_block
//   2. chroma upsampling "jfif-centred" across block borders: 3:1 weights per axis, (x + 2) >> 2 / (x + 8) >> 4 rounding, nearest
//      neighbour for factors other than 1 and 2                                                stbi__resample_row_*
//   3. YCbCr -> RGB in 20-bit fixed point with the constants rounded to 12 bits first and the Cb term of green masked to 16 bits
//                                                                                              stbi__YCbCr_to_RGB_row
// plus its conventions: component planes padded to whole MCUs, colour images come out as 3 channels and grey ones as 1, an Adobe APP14
// transform of 0 without JFIF (or component ids 'R','G','B') means the data are RGB already, CMYK / YCCK through (a * b + 128) / 255.
// Progressive JPEG (SOF2: spectral selection and successive approximation, coefficients kept until the end, dequantised in 16-bit
// arithmetic and then the same inverse DCT) is restated too                                     stbi__jpeg_decode_block_prog_dc / _ac, stbi__jpeg_finish
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace ptjpeg {

struct Huff {
    // canonical code (JPEG Annex C / F.2.2.3): codes of length L are mincode[L] .. maxcode[L], symbol = values[valptr[L] + code - mincode[L]]
    int32_t mincode[17], maxcode[18], valptr[17];
    uint8_t values[256];
    bool ok = false;
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dc_pred = 0;
    int x = 0, y = 0, w2 = 0, h2 = 0;
    std::vector<uint8_t> data;
    std::vector<short> coeff;  // progressive: 64 coefficients per block of the padded plane, w2 / 8 blocks per row
};

class Decoder {
public:
    // returns false with `err` set; on success `out` holds height * width * channels bytes (channels = 3 or 1)
    bool decode(const uint8_t *bytes, size_t n, int &width, int &height, int &channels, std::vector<uint8_t> &out, std::string &err) {
        p_ = bytes;
        end_ = bytes + n;
        if (n < 4 || get8() != 0xFF || get8() != 0xD8) return fail(err, "not a JPEG");
        for (;;) {
            int m = next_marker();
            if (m < 0) return fail(err, "truncated JPEG");
            if (m == 0xD9) break;  // EOI
            if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
                progressive_ = m == 0xC2;
                if (!frame_header(err)) return false;
            } else if (m == 0xDA) {
                if (!have_frame_) return fail(err, "scan before frame header");
                if (!scan_header(err)) return false;
                if (!entropy_coded_data(err)) return false;
            } else if (!table_or_app_marker(m, err)) {
                return false;
            }
        }
        if (!have_frame_) return fail(err, "no frame in JPEG");
        if (progressive_) finish_progressive();
        finish(width, height, channels, out);
        return true;
    }

private:
    static bool fail(std::string &err, const char *why) {
        err = why;
        return false;
    }
    int get8() { return p_ < end_ ? *p_++ : 0; }
    int get16() {
        int a = get8();
        return (a << 8) | get8();
    }
    // stb_image: a marker seen while filling the bit buffer is kept; otherwise the next 0xFF xx with xx != 0xFF
    int next_marker() {
        if (pending_marker_ >= 0) {
            int m = pending_marker_;
            pending_marker_ = -1;
            return m;
        }
        while (p_ < end_) {
            int x = get8();
            if (x != 0xFF) continue;  // junk between segments is skipped
            while (x == 0xFF && p_ < end_) x = get8();
            if (x != 0) return x;
        }
        return -1;
    }

    bool table_or_app_marker(int m, std::string &err) {
        if (m == 0xDD) {  // DRI
            if (get16() != 4) return fail(err, "bad DRI length");
            restart_interval_ = get16();
            return true;
        }
        if (m == 0xDB) {  // DQT
            int L = get16() - 2;
            while (L > 0) {
                const int q = get8(), sixteen = q >> 4, t = q & 15;
                if ((sixteen != 0 && sixteen != 1) || t > 3) return fail(err, "bad DQT");
                for (int i = 0; i < 64; i++) dequant_[t][kZigzag[i]] = (uint16_t)(sixteen ? get16() : get8());
                L -= sixteen ? 129 : 65;
            }
            return L == 0 ? true : fail(err, "bad DQT length");
        }
        if (m == 0xC4) {  // DHT
            int L = get16() - 2;
            while (L > 0) {
                const int q = get8(), tc = q >> 4, th = q & 15;
                if (tc > 1 || th > 3) return fail(err, "bad DHT header");
                int counts[16], total = 0;
                for (int i = 0; i < 16; i++) total += counts[i] = get8();
                if (total > 256) return fail(err, "bad DHT header");
                Huff &h = tc ? ac_[th] : dc_[th];
                for (int i = 0; i < total; i++) h.values[i] = (uint8_t)get8();
                int code = 0, k = 0;
                for (int len = 1; len <= 16; len++) {
                    h.valptr[len] = k;
                    h.mincode[len] = code;
                    code += counts[len - 1];
                    k += counts[len - 1];
                    h.maxcode[len] = counts[len - 1] ? code - 1 : -1;
                    if (counts[len - 1] && code - 1 >= (1 << len)) return fail(err, "bad code lengths");
                    code <<= 1;
                }
                h.maxcode[17] = 0x7fffffff;
                h.ok = true;
                L -= 17 + total;
            }
            return L == 0 ? true : fail(err, "bad DHT length");
        }
        if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) {  // APPn / COM
            int L = get16();
            if (L < 2) return fail(err, "bad segment length");
            L -= 2;
            if (m == 0xE0 && L >= 5) {
                static const char tag[5] = {'J', 'F', 'I', 'F', 0};
                bool ok = true;
                for (int i = 0; i < 5; i++) ok &= get8() == (uint8_t)tag[i];
                L -= 5;
                if (ok) jfif_ = true;
            } else if (m == 0xEE && L >= 12) {
                static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
                bool ok = true;
                for (int i = 0; i < 6; i++) ok &= get8() == (uint8_t)tag[i];
                L -= 6;
                if (ok) {
                    get8();
                    get16();
                    get16();
                    adobe_transform_ = get8();
                    L -= 6;
                }
            }
            p_ = (end_ - p_ < L) ? end_ : p_ + L;
            return true;
        }
        if (m >= 0xD0 && m <= 0xD7) return true;  // stray restart marker
        if (m == 0xDC) {                           // DNL
            get16();
            get16();
            return true;
        }
        return fail(err, "unknown JPEG marker");
    }

    bool frame_header(std::string &err) {
        const int Lf = get16();
        if (Lf < 11) return fail(err, "bad SOF length");
        if (get8() != 8) return fail(err, "only 8-bit JPEG");
        img_y_ = get16();
        img_x_ = get16();
        if (img_x_ == 0 || img_y_ == 0) return fail(err, "empty JPEG");
        n_comp_ = get8();
        if (n_comp_ != 1 && n_comp_ != 3 && n_comp_ != 4) return fail(err, "bad component count");
        if (Lf != 8 + 3 * n_comp_) return fail(err, "bad SOF length");
        if ((uint64_t)img_x_ * (uint64_t)img_y_ * (uint64_t)n_comp_ > 0x7fffffffull) return fail(err, "too large");  // stb_image's 2 GB limit
        rgb_ids_ = 0;
        h_max_ = v_max_ = 1;
        for (int i = 0; i < n_comp_; i++) {
            static const char rgb[3] = {'R', 'G', 'B'};
            Component &c = comp_[i];
            c.id = get8();
            if (n_comp_ == 3 && c.id == rgb[i]) rgb_ids_++;
            const int q = get8();
            c.h = q >> 4;
            c.v = q & 15;
            c.tq = get8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail(err, "bad sampling factors");
            if (c.h > h_max_) h_max_ = c.h;
            if (c.v > v_max_) v_max_ = c.v;
        }
        for (int i = 0; i < n_comp_; i++)
            if (h_max_ % comp_[i].h != 0 || v_max_ % comp_[i].v != 0) return fail(err, "bad sampling factors");
        mcu_w_ = h_max_ * 8;
        mcu_h_ = v_max_ * 8;
        mcu_x_ = (img_x_ + mcu_w_ - 1) / mcu_w_;
        mcu_y_ = (img_y_ + mcu_h_ - 1) / mcu_h_;
        for (int i = 0; i < n_comp_; i++) {
            Component &c = comp_[i];
            c.x = (img_x_ * c.h + h_max_ - 1) / h_max_;  // pixels of this component that the image really has
            c.y = (img_y_ * c.v + v_max_ - 1) / v_max_;
            c.w2 = mcu_x_ * c.h * 8;  // plane padded to whole MCUs
            c.h2 = mcu_y_ * c.v * 8;
            c.data.assign((size_t)c.w2 * (size_t)c.h2, 0);
            if (progressive_) c.coeff.assign((size_t)c.w2 * (size_t)c.h2, 0);
        }
        have_frame_ = true;
        return true;
    }

    bool scan_header(std::string &err) {
        const int Ls = get16();
        scan_n_ = get8();
        if (scan_n_ < 1 || scan_n_ > 4 || scan_n_ > n_comp_ || Ls != 6 + 2 * scan_n_) return fail(err, "bad SOS");
        for (int i = 0; i < scan_n_; i++) {
            const int id = get8(), q = get8();
            int which = 0;
            while (which < n_comp_ && comp_[which].id != id) which++;
            if (which == n_comp_) return fail(err, "bad SOS component");
            comp_[which].hd = q >> 4;
            comp_[which].ha = q & 15;
            if (comp_[which].hd > 3 || comp_[which].ha > 3) return fail(err, "bad SOS tables");
            order_[i] = which;
        }
        spec_start_ = get8();
        spec_end_ = get8();
        const int a = get8();
        succ_high_ = a >> 4;
        succ_low_ = a & 15;
        if (progressive_) {
            if (spec_start_ > 63 || spec_end_ > 63 || spec_start_ > spec_end_ || succ_high_ > 13 || succ_low_ > 13) return fail(err, "bad SOS");
        } else {
            if (spec_start_ != 0 || succ_high_ != 0 || succ_low_ != 0) return fail(err, "bad SOS");
            spec_end_ = 63;
        }
        return true;
    }

    // ---- bit reader: bytes are stuffed (FF 00), a marker ends the data and the reader then supplies zero bits ----
    void reset_entropy() {
        bits_ = 0;
        nbits_ = 0;
        no_more_ = false;
        for (int i = 0; i < 4; i++) comp_[i].dc_pred = 0;
        pending_marker_ = -1;
        todo_ = restart_interval_ ? restart_interval_ : 0x7fffffff;
        eob_run_ = 0;
    }
    void fill() {
        while (nbits_ <= 24) {
            int b = no_more_ ? 0 : get8();
            if (b == 0xFF) {
                int c = get8();
                while (c == 0xFF) c = get8();
                if (c != 0) {
                    pending_marker_ = c;
                    no_more_ = true;
                    b = 0;
                }
            }
            bits_ |= (uint32_t)b << (24 - nbits_);
            nbits_ += 8;
        }
    }
    int take(int n) {  // n <= 16
        if (nbits_ < n) fill();
        const int v = (int)(bits_ >> (32 - n));
        bits_ <<= n;
        nbits_ -= n;
        return v;
    }
    int decode_symbol(const Huff &h) {
        if (nbits_ < 16) fill();
        int code = 0;
        for (int len = 1; len <= 16; len++) {
            code = (int)(bits_ >> (32 - len));
            if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) {
                bits_ <<= len;
                nbits_ -= len;
                return h.values[h.valptr[len] + code - h.mincode[len]];
            }
        }
        return -1;
    }
    int receive_extend(int n) {  // F.2.2.1: n magnitude bits; a leading 0 means a negative value
        const int v = take(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }

    bool decode_block(short data[64], Component &c, std::string &err) {
        const Huff &hd = dc_[c.hd], &ha = ac_[c.ha];
        if (!hd.ok || !ha.ok) return fail(err, "missing Huffman table");
        const uint16_t *dq = dequant_[c.tq];
        memset(data, 0, 64 * sizeof(short));
        const int t = decode_symbol(hd);
        if (t < 0 || t > 15) return fail(err, "bad Huffman code");
        const int diff = t ? receive_extend(t) : 0;
        c.dc_pred += diff;
        data[0] = (short)(c.dc_pred * dq[0]);
        int k = 1;
        do {
            const int rs = decode_symbol(ha);
            if (rs < 0) return fail(err, "bad Huffman code");
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xF0) break;  // end of block
                k += 16;
            } else {
                k += r;
                if (k > 63) return fail(err, "bad AC run");
                const int zig = kZigzag[k++];
                data[zig] = (short)(receive_extend(s) * dq[zig]);
            }
        } while (k < 64);
        return true;
    }

    // ---- progressive scans: one band of one bit plane of the coefficients per scan ----
    bool prog_dc(short *data, Component &c, std::string &err) {
        if (spec_end_ != 0) return fail(err, "can't merge dc and ac");
        if (succ_high_ == 0) {  // first pass of the DC term
            memset(data, 0, 64 * sizeof(short));
            const Huff &hd = dc_[c.hd];
            if (!hd.ok) return fail(err, "missing Huffman table");
            const int t = decode_symbol(hd);
            if (t < 0 || t > 15) return fail(err, "bad Huffman code");
            c.dc_pred += t ? receive_extend(t) : 0;
            data[0] = (short)(c.dc_pred * (1 << succ_low_));
        } else if (take(1)) {  // refinement: one more bit
            data[0] += (short)(1 << succ_low_);
        }
        return true;
    }
    bool prog_ac(short *data, Component &c, std::string &err) {
        if (spec_start_ == 0) return fail(err, "can't merge dc and ac");
        const Huff &ha = ac_[c.ha];
        if (!ha.ok) return fail(err, "missing Huffman table");
        if (succ_high_ == 0) {  // first pass of this band
            if (eob_run_) {
                --eob_run_;
                return true;
            }
            int k = spec_start_;
            do {
                const int rs = decode_symbol(ha);
                if (rs < 0) return fail(err, "bad Huffman code");
                const int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {  // end of band for 2^r + extra blocks
                        eob_run_ = 1 << r;
                        if (r) eob_run_ += take(r);
                        --eob_run_;
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    if (k > 63) return fail(err, "bad AC run");
                    data[kZigzag[k++]] = (short)(receive_extend(s) * (1 << succ_low_));
                }
            } while (k <= spec_end_);
            return true;
        }
        // refinement pass: one more bit for the coefficients that are already non-zero, new +-1 coefficients in between
        const short bit = (short)(1 << succ_low_);
        auto refine = [&](short &p) {
            if (take(1) && (p & bit) == 0) p = (short)(p > 0 ? p + bit : p - bit);
        };
        if (eob_run_) {
            --eob_run_;
            for (int k = spec_start_; k <= spec_end_; k++) {
                short &p = data[kZigzag[k]];
                if (p != 0) refine(p);
            }
            return true;
        }
        int k = spec_start_;
        do {
            const int rs = decode_symbol(ha);
            if (rs < 0) return fail(err, "bad Huffman code");
            int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (r < 15) {
                    eob_run_ = (1 << r) - 1;
                    if (r) eob_run_ += take(r);
                    r = 64;  // to the end of the band
                }
            } else {
                if (s != 1) return fail(err, "bad Huffman code");
                s = take(1) ? bit : -bit;
            }
            while (k <= spec_end_) {  // advance over r zero coefficients, refining the non-zero ones on the way
                short &p = data[kZigzag[k++]];
                if (p != 0) {
                    refine(p);
                } else {
                    if (r == 0) {
                        p = (short)s;
                        break;
                    }
                    --r;
                }
            }
        } while (k <= spec_end_);
        return true;
    }
    bool progressive_scan(std::string &err) {
        bool stop;
        if (scan_n_ == 1) {
            Component &c = comp_[order_[0]];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3, cw = c.w2 / 8;
            for (int j = 0; j < h; j++)
                for (int i = 0; i < w; i++) {
                    short *data = c.coeff.data() + 64 * ((size_t)i + (size_t)j * cw);
                    if (!(spec_start_ == 0 ? prog_dc(data, c, err) : prog_ac(data, c, err))) return false;
                    restart_if_due(stop);
                    if (stop) return true;
                }
            return true;
        }
        for (int j = 0; j < mcu_y_; j++)  // interleaved scans carry DC terms only
            for (int i = 0; i < mcu_x_; i++) {
                for (int k = 0; k < scan_n_; k++) {
                    Component &c = comp_[order_[k]];
                    const int cw = c.w2 / 8;
                    for (int y = 0; y < c.v; y++)
                        for (int x = 0; x < c.h; x++) {
                            short *data = c.coeff.data() + 64 * ((size_t)(i * c.h + x) + (size_t)(j * c.v + y) * cw);
                            if (!prog_dc(data, c, err)) return false;
                        }
                }
                restart_if_due(stop);
                if (stop) return true;
            }
        return true;
    }
    void finish_progressive() {  // dequantise in 16-bit arithmetic, then the same inverse DCT
        for (int n = 0; n < n_comp_; n++) {
            Component &c = comp_[n];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3, cw = c.w2 / 8;
            const uint16_t *dq = dequant_[c.tq];
            for (int j = 0; j < h; j++)
                for (int i = 0; i < w; i++) {
                    short *data = c.coeff.data() + 64 * ((size_t)i + (size_t)j * cw);
                    for (int k = 0; k < 64; k++) data[k] = (short)(data[k] * dq[k]);
                    idct(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, data);
                }
        }
    }

    bool restart_if_due(bool &stop) {
        stop = false;
        if (--todo_ <= 0) {
            if (nbits_ < 24) fill();
            if (!(pending_marker_ >= 0xD0 && pending_marker_ <= 0xD7)) {
                stop = true;  // stb_image: no restart marker here ends the scan quietly
                return true;
            }
            reset_entropy();
        }
        return true;
    }

    bool entropy_coded_data(std::string &err) {
        reset_entropy();
        if (progressive_) return progressive_scan(err);
        short block[64];
        if (scan_n_ == 1) {  // non-interleaved: the component's own blocks, row by row
            Component &c = comp_[order_[0]];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; j++)
                for (int i = 0; i < w; i++) {
                    if (!decode_block(block, c, err)) return false;
                    idct(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, block);
                    bool stop;
                    restart_if_due(stop);
                    if (stop) return true;
                }
            return true;
        }
        for (int j = 0; j < mcu_y_; j++)
            for (int i = 0; i < mcu_x_; i++) {
                for (int k = 0; k < scan_n_; k++) {
                    Component &c = comp_[order_[k]];
                    for (int y = 0; y < c.v; y++)
                        for (int x = 0; x < c.h; x++) {
                            const int x2 = (i * c.h + x) * 8, y2 = (j * c.v + y) * 8;
                            if (!decode_block(block, c, err)) return false;
                            idct(c.data.data() + (size_t)c.w2 * y2 + x2, c.w2, block);
                        }
                }
                bool stop;
                restart_if_due(stop);
                if (stop) return true;
            }
        return true;
    }

    // ---- step 1: stb_image's inverse DCT ----
    static int clamp255(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }
    static constexpr int fx(double c) { return (int)(c * 4096 + 0.5); }
    struct Odd {
        int x0, x1, x2, x3, t0, t1, t2, t3;
    };
    static Odd pass(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
        Odd r;
        int p2 = s2, p3 = s6;
        int p1 = (p2 + p3) * fx(0.5411961f);
        int t2 = p1 + p3 * fx(-1.847759065f);
        int t3 = p1 + p2 * fx(0.765366865f);
        p2 = s0;
        p3 = s4;
        int t0 = (p2 + p3) * 4096, t1 = (p2 - p3) * 4096;
        r.x0 = t0 + t3;
        r.x3 = t0 - t3;
        r.x1 = t1 + t2;
        r.x2 = t1 - t2;
        t0 = s7;
        t1 = s5;
        t2 = s3;
        t3 = s1;
        p3 = t0 + t2;
        int p4 = t1 + t3;
        p1 = t0 + t3;
        p2 = t1 + t2;
        const int p5 = (p3 + p4) * fx(1.175875602f);
        t0 = t0 * fx(0.298631336f);
        t1 = t1 * fx(2.053119869f);
        t2 = t2 * fx(3.072711026f);
        t3 = t3 * fx(1.501321110f);
        p1 = p5 + p1 * fx(-0.899976223f);
        p2 = p5 + p2 * fx(-2.562915447f);
        p3 = p3 * fx(-1.961570560f);
        p4 = p4 * fx(-0.390180644f);
        r.t3 = t3 + p1 + p4;
        r.t2 = t2 + p2 + p3;
        r.t1 = t1 + p2 + p4;
        r.t0 = t0 + p1 + p3;
        return r;
    }
    static void idct(uint8_t *out, int stride, const short d[64]) {
        int v[64];
        for (int i = 0; i < 8; i++) {  // columns; a column with only a DC term is that term times 4
            if (d[i + 8] == 0 && d[i + 16] == 0 && d[i + 24] == 0 && d[i + 32] == 0 && d[i + 40] == 0 && d[i + 48] == 0 && d[i + 56] == 0) {
                const int dc = d[i] * 4;
                for (int k = 0; k < 8; k++) v[i + 8 * k] = dc;
                continue;
            }
            Odd r = pass(d[i], d[i + 8], d[i + 16], d[i + 24], d[i + 32], d[i + 40], d[i + 48], d[i + 56]);
            r.x0 += 512; r.x1 += 512; r.x2 += 512; r.x3 += 512;  // keep 2 extra bits
            v[i + 0] = (r.x0 + r.t3) >> 10;
            v[i + 56] = (r.x0 - r.t3) >> 10;
            v[i + 8] = (r.x1 + r.t2) >> 10;
            v[i + 48] = (r.x1 - r.t2) >> 10;
            v[i + 16] = (r.x2 + r.t1) >> 10;
            v[i + 40] = (r.x2 - r.t1) >> 10;
            v[i + 24] = (r.x3 + r.t0) >> 10;
            v[i + 32] = (r.x3 - r.t0) >> 10;
        }
        for (int i = 0; i < 8; i++) {  // rows: 17 bits to drop, rounding and the +128 level shift folded into one constant
            const int *w = v + 8 * i;
            uint8_t *o = out + (size_t)stride * i;
            Odd r = pass(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
            const int bias = 65536 + (128 << 17);
            r.x0 += bias; r.x1 += bias; r.x2 += bias; r.x3 += bias;
            o[0] = (uint8_t)clamp255((r.x0 + r.t3) >> 17);
            o[7] = (uint8_t)clamp255((r.x0 - r.t3) >> 17);
            o[1] = (uint8_t)clamp255((r.x1 + r.t2) >> 17);
            o[6] = (uint8_t)clamp255((r.x1 - r.t2) >> 17);
            o[2] = (uint8_t)clamp255((r.x2 + r.t1) >> 17);
            o[5] = (uint8_t)clamp255((r.x2 - r.t1) >> 17);
            o[3] = (uint8_t)clamp255((r.x3 + r.t0) >> 17);
            o[4] = (uint8_t)clamp255((r.x3 - r.t0) >> 17);
        }
    }

    // ---- step 2: one output row of a component from its two nearest stored rows ----
    static const uint8_t *upsample_row(uint8_t *out, const uint8_t *near_row, const uint8_t *far_row, int w, int hs, int vs) {
        if (hs == 1 && vs == 1) return near_row;
        if (hs == 1 && vs == 2) {
            for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * near_row[i] + far_row[i] + 2) >> 2);
            return out;
        }
        if (hs == 2 && vs == 1) {
            const uint8_t *in = near_row;
            if (w == 1) {
                out[0] = out[1] = in[0];
                return out;
            }
            out[0] = in[0];
            out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
            int i;
            for (i = 1; i < w - 1; i++) {
                const int n = 3 * in[i] + 2;
                out[i * 2] = (uint8_t)((n + in[i - 1]) >> 2);
                out[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2);
            }
            out[i * 2] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
            out[i * 2 + 1] = in[w - 1];
            return out;
        }
        if (hs == 2 && vs == 2) {
            if (w == 1) {
                out[0] = out[1] = (uint8_t)((3 * near_row[0] + far_row[0] + 2) >> 2);
                return out;
            }
            int t1 = 3 * near_row[0] + far_row[0];
            out[0] = (uint8_t)((t1 + 2) >> 2);
            for (int i = 1; i < w; i++) {
                const int t0 = t1;
                t1 = 3 * near_row[i] + far_row[i];
                out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
                out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
            }
            out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
            return out;
        }
        for (int i = 0; i < w; i++)  // other factors: nearest neighbour horizontally, the near row vertically
            for (int j = 0; j < hs; j++) out[i * hs + j] = near_row[i];
        return out;
    }

    // ---- step 3 ----
    static constexpr int fixed20(double c) { return ((int)(c * 4096.0f + 0.5f)) << 8; }
    static void ycc_to_rgb(uint8_t *out, const uint8_t *y, const uint8_t *cbp, const uint8_t *crp, int count) {
        for (int i = 0; i < count; i++, out += 3) {
            const int yf = (y[i] << 20) + (1 << 19);
            const int cr = crp[i] - 128, cb = cbp[i] - 128;
            int r = yf + cr * fixed20(1.40200f);
            int g = yf + (cr * -fixed20(0.71414f)) + (int)((uint32_t)(cb * -fixed20(0.34414f)) & 0xffff0000u);
            int b = yf + cb * fixed20(1.77200f);
            out[0] = (uint8_t)clamp255(r >> 20);
            out[1] = (uint8_t)clamp255(g >> 20);
            out[2] = (uint8_t)clamp255(b >> 20);
        }
    }
    static uint8_t mul255(uint8_t a, uint8_t b) {  // (a * b) / 255 rounded, without a division
        const unsigned t = (unsigned)a * b + 128;
        return (uint8_t)((t + (t >> 8)) >> 8);
    }

    void finish(int &width, int &height, int &channels, std::vector<uint8_t> &out) {
        width = img_x_;
        height = img_y_;
        const int n = n_comp_ >= 3 ? 3 : 1;
        channels = n;
        const bool is_rgb = n_comp_ == 3 && (rgb_ids_ == 3 || (adobe_transform_ == 0 && !jfif_));
        out.assign((size_t)n * img_x_ * img_y_, 0);
        struct Up {
            int hs, vs, ystep, w_lores, ypos;
            const uint8_t *line0, *line1;
            std::vector<uint8_t> buf;
        } up[4];
        for (int k = 0; k < n_comp_; k++) {
            Up &u = up[k];
            u.hs = h_max_ / comp_[k].h;
            u.vs = v_max_ / comp_[k].v;
            u.ystep = u.vs >> 1;
            u.w_lores = (img_x_ + u.hs - 1) / u.hs;
            u.ypos = 0;
            u.line0 = u.line1 = comp_[k].data.data();
            u.buf.assign((size_t)img_x_ + 8, 0);
        }
        const uint8_t *row[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int j = 0; j < img_y_; j++) {
            uint8_t *o = out.data() + (size_t)n * img_x_ * j;
            for (int k = 0; k < n_comp_; k++) {
                Up &u = up[k];
                const bool bottom = u.ystep >= (u.vs >> 1);  // which of the two stored rows is nearer to this output row
                row[k] = upsample_row(u.buf.data(), bottom ? u.line1 : u.line0, bottom ? u.line0 : u.line1, u.w_lores, u.hs, u.vs);
                if (++u.ystep >= u.vs) {
                    u.ystep = 0;
                    u.line0 = u.line1;
                    if (++u.ypos < comp_[k].y) u.line1 += comp_[k].w2;
                }
            }
            if (n == 1) {
                memcpy(o, row[0], (size_t)img_x_);
            } else if (n_comp_ == 3) {
                if (is_rgb)
                    for (int i = 0; i < img_x_; i++) { o[3 * i] = row[0][i]; o[3 * i + 1] = row[1][i]; o[3 * i + 2] = row[2][i]; }
                else
                    ycc_to_rgb(o, row[0], row[1], row[2], img_x_);
            } else {  // four components
                if (adobe_transform_ == 0) {  // CMYK
                    for (int i = 0; i < img_x_; i++) {
                        const uint8_t m = row[3][i];
                        o[3 * i] = mul255(row[0][i], m); o[3 * i + 1] = mul255(row[1][i], m); o[3 * i + 2] = mul255(row[2][i], m);
                    }
                } else {
                    ycc_to_rgb(o, row[0], row[1], row[2], img_x_);
                    if (adobe_transform_ == 2)  // YCCK
                        for (int i = 0; i < img_x_; i++) {
                            const uint8_t m = row[3][i];
                            o[3 * i] = mul255((uint8_t)(255 - o[3 * i]), m); o[3 * i + 1] = mul255((uint8_t)(255 - o[3 * i + 1]), m); o[3 * i + 2] = mul255((uint8_t)(255 - o[3 * i + 2]), m);
                        }
                }
            }
        }
    }

    static constexpr uint8_t kZigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                                 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                                 39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

    const uint8_t *p_ = nullptr, *end_ = nullptr;
    uint16_t dequant_[4][64] = {};
    Huff dc_[4], ac_[4];
    Component comp_[4];
    int img_x_ = 0, img_y_ = 0, n_comp_ = 0, h_max_ = 1, v_max_ = 1, mcu_w_ = 8, mcu_h_ = 8, mcu_x_ = 0, mcu_y_ = 0;
    int scan_n_ = 0, order_[4] = {0, 0, 0, 0};
    int restart_interval_ = 0, todo_ = 0, rgb_ids_ = 0, adobe_transform_ = -1, pending_marker_ = -1;
    int spec_start_ = 0, spec_end_ = 63, succ_high_ = 0, succ_low_ = 0, eob_run_ = 0;
    bool jfif_ = false, have_frame_ = false, no_more_ = false, progressive_ = false;
    uint32_t bits_ = 0;
    int nbits_ = 0;
};

}  // namespace ptjpeg
