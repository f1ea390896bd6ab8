// SceneLoader.cpp — own scene ingestion (GLB/glTF 2.0, OBJ/MTL, PNG, .ptscene).
//
// Replaces the reference's assimp-backed loader (src/HostScene.cpp:98-278).  The
// behaviours of assimp 5.4.3 (conanfile.txt:2, absent here) that shape the
// HostScene are restated, not linked:
//   * aiProcess_PreTransformVertices: node TRS baked into positions with float
//     4x4 matrices, world = parent * (T * R * S); one output mesh per material in
//     material-index order, primitives inside a material in depth-first node order.
//   * glTF2 importer: V flipped to 1 - v; baseColorFactor default (1,1,1),
//     emissiveFactor default (0,0,0); KHR_materials_emissive_strength ignored.
//   * aiProcess_FindDegenerates + SortByPType(remove POINT|LINE): triangles with
//     two coincident corners are dropped.
//   * texture decode as the reference does it (src/HostScene.cpp:10-51): 8-bit
//     samples copied 3 bytes per texel into float3 0..255.
// Parity for this stage is pinned by SURVEY.md Appendix A (counts, bounds,
// first/last triangles), not by the reference (it has no tests): "parity unpinned".
#include "HostScene.h"
#include "BmpTgaDecoder.h"
#include "JpegDecoder.h"

#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace {

// ---------------------------------------------------------------------------------
// minimal JSON
// ---------------------------------------------------------------------------------
struct JVal;
using JPtr = std::shared_ptr<JVal>;
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<JPtr> arr;
    std::vector<std::pair<std::string, JPtr>> obj;

    const JVal *get(const char *key) const {
        if (kind != Obj) return nullptr;
        for (auto &kv : obj)
            if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    bool has(const char *key) const { return get(key) != nullptr; }
    const JVal &need(const char *key) const {  // a required member: a malformed file raises, it never dereferences null
        const JVal *v = get(key);
        if (!v) throw std::runtime_error(std::string("glTF: missing \"") + key + "\"");
        return *v;
    }
    size_t size() const { return kind == Arr ? arr.size() : 0; }
    const JVal &at(size_t i) const {
        if (kind != Arr || i >= arr.size()) throw std::runtime_error("glTF: array index out of range");
        return *arr[i];
    }
    double number(double dflt) const { return kind == Num ? num : dflt; }
    int integer(int dflt) const { return kind == Num ? (int)num : dflt; }
};

struct JParser {
    const char *p, *end;
    explicit JParser(const char *s, size_t n) : p(s), end(s + n) {}
    [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("JSON parse error: ") + what); }
    void ws() {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    }
    int depth = 0;  // nesting of the value being parsed: bounded, so that a file of 100 000 '[' raises instead of overflowing the stack
    struct Nest {
        int &d;
        explicit Nest(int &dd) : d(dd) { ++d; }
        ~Nest() { --d; }
    };
    JPtr parse() {
        Nest nest(depth);
        if (depth > 256) fail("nesting too deep");
        ws();
        if (p >= end) fail("unexpected end");
        auto v = std::make_shared<JVal>();
        char c = *p;
        if (c == '{') {
            v->kind = JVal::Obj;
            ++p;
            ws();
            if (p < end && *p == '}') { ++p; return v; }
            for (;;) {
                ws();
                std::string k = parseString();
                ws();
                if (p >= end || *p != ':') fail("expected ':'");
                ++p;
                v->obj.emplace_back(std::move(k), parse());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v->kind = JVal::Arr;
            ++p;
            ws();
            if (p < end && *p == ']') { ++p; return v; }
            for (;;) {
                v->arr.push_back(parse());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v->kind = JVal::Str;
            v->str = parseString();
        } else if (c == 't' && end - p >= 4 && !strncmp(p, "true", 4)) {
            v->kind = JVal::Bool; v->b = true; p += 4;
        } else if (c == 'f' && end - p >= 5 && !strncmp(p, "false", 5)) {
            v->kind = JVal::Bool; v->b = false; p += 5;
        } else if (c == 'n' && end - p >= 4 && !strncmp(p, "null", 4)) {
            p += 4;
        } else {
            char *e = nullptr;
            std::string tmp(p, std::min<size_t>(end - p, 64));
            double d = strtod(tmp.c_str(), &e);
            if (e == tmp.c_str()) fail("bad number");
            p += (e - tmp.c_str());
            v->kind = JVal::Num;
            v->num = d;
        }
        return v;
    }
    std::string parseString() {
        if (p >= end || *p != '"') fail("expected string");
        ++p;
        std::string s;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                ++p;
                if (p >= end) fail("bad escape");
                switch (*p) {
                    case 'n': s += '\n'; break;
                    case 't': s += '\t'; break;
                    case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break;
                    case 'u': {
                        if (end - p < 5) fail("bad \\u");
                        unsigned cp = (unsigned)strtoul(std::string(p + 1, 4).c_str(), nullptr, 16);
                        p += 4;
                        if (cp < 0x80) s += (char)cp;
                        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
                        else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: s += *p;
                }
                ++p;
            } else {
                s += *p++;
            }
        }
        if (p >= end) fail("unterminated string");
        ++p;
        return s;
    }
};

std::vector<unsigned char> read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Cannot open file: " + path);
    f.seekg(0, std::ios::end);
    std::streamoff n = f.tellg();
    f.seekg(0);
    std::vector<unsigned char> buf((size_t)n);
    if (n > 0) f.read((char *)buf.data(), n);
    return buf;
}

// ---------------------------------------------------------------------------------
// float 4x4, row-major, the arithmetic order of assimp's aiMatrix4x4t<float>
// ---------------------------------------------------------------------------------
struct Mat4 {
    float m[4][4];
    static Mat4 identity() {
        Mat4 r{};
        for (int i = 0; i < 4; i++) r.m[i][i] = 1.f;
        return r;
    }
};
Mat4 mul(const Mat4 &a, const Mat4 &b) {
    Mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return r;
}
float3 xform(const Mat4 &a, float3 v) {
    return make_float3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z + a.m[0][3],
                       a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z + a.m[1][3],
                       a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z + a.m[2][3]);
}
Mat4 quat_matrix(float x, float y, float z, float w) {
    Mat4 r = Mat4::identity();
    r.m[0][0] = 1.f - 2.f * (y * y + z * z);
    r.m[0][1] = 2.f * (x * y - z * w);
    r.m[0][2] = 2.f * (x * z + y * w);
    r.m[1][0] = 2.f * (x * y + z * w);
    r.m[1][1] = 1.f - 2.f * (x * x + z * z);
    r.m[1][2] = 2.f * (y * z - x * w);
    r.m[2][0] = 2.f * (x * z - y * w);
    r.m[2][1] = 2.f * (y * z + x * w);
    r.m[2][2] = 1.f - 2.f * (x * x + y * y);
    return r;
}

// ---------------------------------------------------------------------------------
// glTF
// ---------------------------------------------------------------------------------
struct Gltf {
    JPtr root;
    std::vector<std::vector<unsigned char>> buffers;
    std::string dir;

    const JVal &top(const char *k) const {
        static JVal empty;
        const JVal *v = root->get(k);
        return v ? *v : empty;
    }

    struct View { const unsigned char *ptr; size_t len; size_t stride; };
    View bufferView(int idx) const {
        const JVal &bv = top("bufferViews").at((size_t)idx);
        int b = bv.get("buffer") ? bv.get("buffer")->integer(0) : 0;
        size_t off = sizeFrom(bv.get("byteOffset"), "byteOffset");
        size_t len = sizeFrom(bv.get("byteLength"), "byteLength");
        size_t stride = sizeFrom(bv.get("byteStride"), "byteStride");
        if (b < 0 || (size_t)b >= buffers.size() || off + len > buffers[(size_t)b].size())
            throw std::runtime_error("glTF: bufferView out of range");
        return {buffers[(size_t)b].data() + off, len, stride};
    }

    static int compSize(int ct) {
        switch (ct) {
            case 5120: case 5121: return 1;
            case 5122: case 5123: return 2;
            case 5125: case 5126: return 4;
        }
        throw std::runtime_error("glTF: unsupported componentType");
    }
    static int typeCount(const std::string &t) {
        if (t == "SCALAR") return 1;
        if (t == "VEC2") return 2;
        if (t == "VEC3") return 3;
        if (t == "VEC4") return 4;
        throw std::runtime_error("glTF: unsupported accessor type " + t);
    }

    // sizes out of the JSON: a negative, absurd or non-numeric value raises instead of being cast
    static size_t sizeFrom(const JVal *v, const char *what) {
        const double d = v ? v->number(-1) : 0.0;
        if (!(d >= 0.0) || d > 1e15) throw std::runtime_error(std::string("glTF: bad ") + what);
        return (size_t)d;
    }
    static size_t countOf(const JVal &acc) { return sizeFrom(&acc.need("count"), "accessor count"); }
    static size_t offsetOf(const JVal &acc) { return sizeFrom(acc.get("byteOffset"), "byteOffset"); }

    // Reads accessor `idx` as floats (normalised integers are scaled) with `want` comps.
    std::vector<float> readFloats(int idx, int want) const {
        const JVal &acc = top("accessors").at((size_t)idx);
        if (acc.has("sparse")) throw std::runtime_error("glTF: sparse accessors unsupported");
        int ct = acc.need("componentType").integer(0);
        int n = typeCount(acc.need("type").str);
        size_t count = countOf(acc);
        bool normalized = acc.get("normalized") && acc.get("normalized")->b;
        if (!acc.has("bufferView")) {
            if (count > (1u << 28)) throw std::runtime_error("glTF: bad accessor count");
            return std::vector<float>(count * (size_t)want, 0.f);
        }
        View v = bufferView(acc.need("bufferView").integer(0));
        size_t aoff = offsetOf(acc);
        size_t cs = (size_t)compSize(ct);
        size_t stride = v.stride ? v.stride : cs * (size_t)n;
        // count, stride and offset are each bounded by the view's length first, so the exact test below cannot wrap around
        if (count > v.len || aoff > v.len || (count > 1 && stride > v.len) || (count && aoff + (count - 1) * stride + cs * (size_t)n > v.len))
            throw std::runtime_error("glTF: accessor overruns bufferView");
        std::vector<float> out(count * (size_t)want, 0.f);
        for (size_t i = 0; i < count; i++) {
            const unsigned char *e = v.ptr + aoff + i * stride;
            for (int c = 0; c < std::min(n, want); c++) {
                const unsigned char *q = e + (size_t)c * cs;
                float f;
                switch (ct) {
                    case 5126: memcpy(&f, q, 4); break;
                    case 5121: f = normalized ? (float)q[0] / 255.f : (float)q[0]; break;
                    case 5120: f = normalized ? std::max((float)(int8_t)q[0] / 127.f, -1.f) : (float)(int8_t)q[0]; break;
                    case 5123: { uint16_t u; memcpy(&u, q, 2); f = normalized ? (float)u / 65535.f : (float)u; break; }
                    case 5122: { int16_t u; memcpy(&u, q, 2); f = normalized ? std::max((float)u / 32767.f, -1.f) : (float)u; break; }
                    case 5125: { uint32_t u; memcpy(&u, q, 4); f = (float)u; break; }
                    default: throw std::runtime_error("glTF: bad componentType");
                }
                out[i * (size_t)want + (size_t)c] = f;
            }
        }
        return out;
    }
    std::vector<uint32_t> readIndices(int idx) const {
        const JVal &acc = top("accessors").at((size_t)idx);
        int ct = acc.need("componentType").integer(0);
        size_t count = countOf(acc);
        View v = bufferView(acc.need("bufferView").integer(0));
        size_t aoff = offsetOf(acc);
        size_t cs = (size_t)compSize(ct);
        size_t stride = v.stride ? v.stride : cs;
        if (count > v.len || aoff > v.len || (count > 1 && stride > v.len) || (count && aoff + (count - 1) * stride + cs > v.len))
            throw std::runtime_error("glTF: index accessor overruns bufferView");
        std::vector<uint32_t> out(count);
        for (size_t i = 0; i < count; i++) {
            const unsigned char *q = v.ptr + aoff + i * stride;
            switch (ct) {
                case 5121: out[i] = q[0]; break;
                case 5123: { uint16_t u; memcpy(&u, q, 2); out[i] = u; break; }
                case 5125: { uint32_t u; memcpy(&u, q, 4); out[i] = u; break; }
                default: throw std::runtime_error("glTF: bad index componentType");
            }
        }
        return out;
    }
};

std::vector<unsigned char> base64_decode(const std::string &s) {
    static int8_t T[256];
    static bool init = false;
    if (!init) {
        memset(T, -1, sizeof T);
        const char *A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) T[(unsigned char)A[i]] = (int8_t)i;
        init = true;
    }
    std::vector<unsigned char> out;
    unsigned acc = 0;
    int bits = 0;
    for (unsigned char c : s) {
        if (T[c] < 0) continue;
        acc = (acc << 6) | (unsigned)T[c];
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((unsigned char)((acc >> bits) & 0xFF));
        }
    }
    return out;
}

bool same_pos(const float3 &a, const float3 &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

material_type type_from_name(const std::string &name) {
    // reference src/obj_loader.h:65-96: name *prefix* selects the class.
    if (name.rfind("lambertian", 0) == 0) return LAMBERTIAN;
    if (name.rfind("metal", 0) == 0) return METAL;
    if (name.rfind("dielectric", 0) == 0) return DIELECTRIC;
    if (name.rfind("diffuse_light", 0) == 0) return DIFFUSE_LIGHT;
    return UNIVERSAL;
}

HostTexture texture_from_bytes(const unsigned char *bytes, size_t n, const std::string &what) {
    int w = 0, h = 0, ch = 0;
    std::vector<unsigned char> px;
    std::string err;
    static const unsigned char png_sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    const bool is_jpeg = n >= 3 && bytes[0] == 0xFF && bytes[1] == 0xD8 && bytes[2] == 0xFF;
    if (is_jpeg) {
        // JPEG (sequential and progressive): JpegDecoder.h restates the three implementation-defined steps of stb_image (IDCT, chroma
        // upsampling, YCbCr -> RGB), so the texels are the reference's byte for byte (pinned against stb_image itself: oracle/ref_stb.c)
        ptjpeg::Decoder dec;
        if (dec.decode(bytes, n, w, h, ch, px, err)) {
            HostTexture t;
            t.width = w;
            t.height = h;
            t.data.resize((size_t)w * (size_t)h);
            // three channels as decoded; a grey JPEG is replicated, which is what stbi_load(path, ..., 3) of the reference's file path does
            // (its embedded-texture path walks a 1-channel buffer three bytes at a time: out of bounds there, src/HostScene.cpp:18-26,37-46)
            for (size_t j = 0; j < t.data.size(); j++)
                t.data[j] = ch == 3 ? make_float3((float)px[3 * j], (float)px[3 * j + 1], (float)px[3 * j + 2]) : make_float3((float)px[j], (float)px[j], (float)px[j]);
            return t;
        }
        fprintf(stderr, "SceneLoader: JPEG texture %s: %s\n", what.c_str(), err.c_str());
    }
    const bool is_png = n >= 8 && memcmp(bytes, png_sig, 8) == 0;
    if (!is_jpeg && !is_png) {
        // BMP and TGA (texture files of an .mtl / .gltf): BmpTgaDecoder.h restates stb_image's choices for stbi_load(path, ..., 3), the call
        // the reference makes for files (src/HostScene.cpp:29); byte-identical texels, tests/golden/images/.  stb_image probes BMP before TGA
        // (TGA has no signature and is its last resort); GIF / PSD / PIC / PNM / HDR, which it probes in between, all fail the TGA test.
        // probing order as in stb_image: BMP, GIF, (PSD, PIC: not read here,) PNM, (HDR,) and TGA last
        const bool bmp = ptimg::is_bmp(bytes, n), gif = !bmp && ptimg::is_gif(bytes, n), pnm = !bmp && !gif && ptimg::is_pnm(bytes, n);
        if (bmp || gif || pnm || ptimg::is_tga(bytes, n)) {
            if (bmp ? ptimg::decode_bmp(bytes, n, w, h, px, err) : gif ? ptimg::decode_gif(bytes, n, w, h, px, err) : pnm ? ptimg::decode_pnm(bytes, n, w, h, px, err)
                                                                                                                        : ptimg::decode_tga(bytes, n, w, h, px, err)) {
                HostTexture t;
                t.width = w;
                t.height = h;
                t.data.resize((size_t)w * (size_t)h);
                for (size_t j = 0; j < t.data.size(); j++) t.data[j] = make_float3((float)px[3 * j], (float)px[3 * j + 1], (float)px[3 * j + 2]);
                return t;
            }
            fprintf(stderr, "SceneLoader: %s texture %s: %s\n", bmp ? "BMP" : gif ? "GIF" : pnm ? "PNM" : "TGA", what.c_str(), err.c_str());
        }
    }
    if (is_jpeg || !is_png) {
        // The reference decodes textures with stb_image (src/HostScene.cpp:10-51), which also reads PSD / PIC / HDR; PNG, JPEG,
        // BMP, TGA, GIF (first frame) and binary PNM are restated here.  Any other image (or a variant stb_image refuses too, e.g. RLE BMP) does not abort the load: the material keeps its slot and gets a
        // texture without texels, which the device shades with the reference's own placeholder colour for a texture without data
        // (242, 45, 27: src/Texture.h:33-35).  README.md / INTEGRATION.md state the restriction.
        const char *kind = is_jpeg ? "JPEG this decoder does not cover" : (n >= 2 && bytes[0] == 'B' && bytes[1] == 'M') ? "BMP variant that is not decoded" : "image format other than PNG / JPEG / BMP / TGA / GIF / PNM";
        fprintf(stderr, "SceneLoader: texture %s is a %s; the placeholder colour (242, 45, 27) is used instead\n", what.c_str(), kind);
        HostTexture t;
        t.width = 1;
        t.height = 1;
        return t;  // no data
    }
    if (!decode_png(bytes, n, w, h, ch, px, err)) throw std::runtime_error("Cannot load texture data, path: " + what + " (" + err + ")");
    // reference src/HostScene.cpp:37-46: walks the decoded buffer 3 bytes per texel
    // whatever the channel count was; identical for 3-channel images (the duck).
    HostTexture t;
    t.width = w;
    t.height = h;
    t.data.resize((size_t)w * (size_t)h);
    size_t total = (size_t)w * (size_t)h * 3;
    for (size_t i = 0, j = 0; i + 2 < total + 0 && i + 2 < px.size(); i += 3, j++)
        t.data[j] = make_float3((float)px[i], (float)px[i + 1], (float)px[i + 2]);
    return t;
}

}  // namespace

// -------------------------------------------------------------------------------------
// PNG (zlib inflate + unfilter + palette/gray expansion); non-interlaced and Adam7.
// -------------------------------------------------------------------------------------
static inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

static bool unfilter_pass(const unsigned char *in, size_t in_len, size_t &consumed, int w, int h, int bits_pp,
                          std::vector<unsigned char> &rows) {
    size_t bpp = (size_t)std::max(1, bits_pp / 8);
    size_t rowbytes = ((size_t)w * (size_t)bits_pp + 7) / 8;
    rows.assign(rowbytes * (size_t)h, 0);
    size_t pos = 0;
    for (int y = 0; y < h; y++) {
        if (pos + 1 + rowbytes > in_len) return false;
        int ft = in[pos++];
        unsigned char *cur = rows.data() + (size_t)y * rowbytes;
        const unsigned char *prev = y ? cur - rowbytes : nullptr;
        for (size_t i = 0; i < rowbytes; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0;
            int b = prev ? prev[i] : 0;
            int c = (prev && i >= bpp) ? prev[i - bpp] : 0;
            int x = in[pos + i];
            switch (ft) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) >> 1; break;
                case 4: x += paeth(a, b, c); break;
                default: return false;
            }
            cur[i] = (unsigned char)x;
        }
        pos += rowbytes;
    }
    consumed = pos;
    return true;
}

bool decode_png(const unsigned char *bytes, size_t n, int &width, int &height, int &channels,
                std::vector<unsigned char> &out, std::string &err) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8 || memcmp(bytes, sig, 8) != 0) { err = "not a PNG (only PNG textures are supported)"; return false; }
    auto be32 = [](const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; };
    size_t pos = 8;
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte, trns;
    bool have_ihdr = false;
    while (pos + 12 <= n) {
        uint32_t len = be32(bytes + pos);
        const unsigned char *tag = bytes + pos + 4;
        const unsigned char *data = bytes + pos + 8;
        if (pos + 12 + (size_t)len > n) { err = "truncated chunk"; return false; }
        if (!memcmp(tag, "IHDR", 4)) {
            if (len < 13) { err = "bad IHDR"; return false; }
            w = (int)be32(data); h = (int)be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            have_ihdr = true;
        } else if (!memcmp(tag, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(tag, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(tag, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w <= 0 || h <= 0) { err = "missing IHDR"; return false; }
    // stb_image's limits (STBI_MAX_DIMENSIONS and a 2 GB image): a corrupted header must not become a 50 GB allocation
    if (w > (1 << 24) || h > (1 << 24) || (uint64_t)w * (uint64_t)h * 8u > 0x7fffffffull) { err = "Very large image (corrupt?)"; return false; }
    int samples;
    switch (ctype) {
        case 0: samples = 1; break;
        case 2: samples = 3; break;
        case 3: samples = 1; break;
        case 4: samples = 2; break;
        case 6: samples = 4; break;
        default: err = "bad colour type"; return false;
    }
    if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) { err = "bad bit depth"; return false; }
    int bits_pp = samples * depth;

    // inflate
    std::vector<unsigned char> raw;
    {
        size_t cap = 0;
        if (!interlace) cap = ((size_t)(((size_t)w * (size_t)bits_pp + 7) / 8) + 1) * (size_t)h;
        else cap = ((size_t)(((size_t)w * (size_t)bits_pp + 7) / 8) + 8) * (size_t)(h + 7) + 64;
        raw.resize(cap);
        z_stream zs{};
        if (inflateInit(&zs) != Z_OK) { err = "zlib init"; return false; }
        zs.next_in = idat.data();
        zs.avail_in = (uInt)idat.size();
        zs.next_out = raw.data();
        zs.avail_out = (uInt)raw.size();
        int rc = inflate(&zs, Z_FINISH);
        size_t got = zs.total_out;
        inflateEnd(&zs);
        if (rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) { err = "zlib inflate failed"; return false; }
        raw.resize(got);
    }

    // unfilter into packed rows of the full image: `samples` per pixel at `depth` bits
    size_t full_rowbytes = ((size_t)w * (size_t)bits_pp + 7) / 8;
    std::vector<unsigned char> img(full_rowbytes * (size_t)h, 0);
    if (!interlace) {
        size_t used = 0;
        if (!unfilter_pass(raw.data(), raw.size(), used, w, h, bits_pp, img)) { err = "bad scanline data"; return false; }
    } else {
        static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1};
        static const int dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
        size_t off = 0;
        for (int p = 0; p < 7; p++) {
            int pw = (w - xs[p] + dx[p] - 1) / dx[p], ph = (h - ys[p] + dy[p] - 1) / dy[p];
            if (pw <= 0 || ph <= 0) continue;
            std::vector<unsigned char> rows;
            size_t used = 0;
            if (!unfilter_pass(raw.data() + off, raw.size() - off, used, pw, ph, bits_pp, rows)) { err = "bad interlaced data"; return false; }
            off += used;
            size_t prow = ((size_t)pw * (size_t)bits_pp + 7) / 8;
            for (int y = 0; y < ph; y++)
                for (int x = 0; x < pw; x++) {
                    int X = xs[p] + x * dx[p], Y = ys[p] + y * dy[p];
                    if (bits_pp >= 8) {
                        size_t b = (size_t)bits_pp / 8;
                        memcpy(&img[(size_t)Y * full_rowbytes + (size_t)X * b], &rows[(size_t)y * prow + (size_t)x * b], b);
                    } else {
                        size_t sb = (size_t)x * (size_t)bits_pp, db = (size_t)X * (size_t)bits_pp;
                        int v = (rows[(size_t)y * prow + sb / 8] >> (8 - bits_pp - (int)(sb % 8))) & ((1 << bits_pp) - 1);
                        img[(size_t)Y * full_rowbytes + db / 8] |= (unsigned char)(v << (8 - bits_pp - (int)(db % 8)));
                    }
                }
        }
    }

    // expand to 8-bit samples
    auto sample = [&](int x, int y, int s) -> int {
        const unsigned char *row = img.data() + (size_t)y * full_rowbytes;
        if (depth == 8) return row[(size_t)x * (size_t)samples + (size_t)s];
        if (depth == 16) return row[((size_t)x * (size_t)samples + (size_t)s) * 2];  // high byte, as stb_image does
        size_t bit = ((size_t)x * (size_t)samples + (size_t)s) * (size_t)depth;
        return (row[bit / 8] >> (8 - depth - (int)(bit % 8))) & ((1 << depth) - 1);
    };
    if (ctype == 3) {
        channels = trns.empty() ? 3 : 4;
        out.resize((size_t)w * (size_t)h * (size_t)channels);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                size_t idx = (size_t)sample(x, y, 0);
                unsigned char *o = &out[((size_t)y * (size_t)w + (size_t)x) * (size_t)channels];
                for (int c = 0; c < 3; c++) o[c] = idx * 3 + (size_t)c < plte.size() ? plte[idx * 3 + (size_t)c] : 0;
                if (channels == 4) o[3] = idx < trns.size() ? trns[idx] : 255;
            }
    } else {
        bool add_alpha = !trns.empty() && (ctype == 0 || ctype == 2);
        channels = samples + (add_alpha ? 1 : 0);
        out.resize((size_t)w * (size_t)h * (size_t)channels);
        int scale = depth < 8 ? 255 / ((1 << depth) - 1) : 1;
        // colour-key transparency (tRNS on a grey or RGB image) as stb_image resolves it: the key's low byte scaled like the samples
        // (16-bit images: the whole 16-bit key against the whole 16-bit samples); alpha 0 where every sample equals the key, else 255.
        // The reference walks an embedded texture's buffer three bytes per texel whatever the channel count, so alpha bytes become texels.
        int key[3] = {0, 0, 0};
        if (add_alpha) {
            if (trns.size() != (size_t)samples * 2) { err = "bad tRNS len"; return false; }
            for (int s = 0; s < samples; s++) {
                const int k16 = (trns[(size_t)s * 2] << 8) | trns[(size_t)s * 2 + 1];
                key[s] = depth == 16 ? k16 : (((k16 & 255) * scale) & 255);
            }
        }
        auto sample16 = [&](int x, int y, int s) -> int {
            const unsigned char *row = img.data() + (size_t)y * full_rowbytes;
            const size_t at = ((size_t)x * (size_t)samples + (size_t)s) * 2;
            return (row[at] << 8) | row[at + 1];
        };
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                unsigned char *o = &out[((size_t)y * (size_t)w + (size_t)x) * (size_t)channels];
                bool is_key = add_alpha;
                for (int s = 0; s < samples; s++) {
                    o[s] = (unsigned char)(sample(x, y, s) * scale);
                    if (add_alpha) is_key = is_key && (depth == 16 ? sample16(x, y, s) : (int)o[s]) == key[s];
                }
                if (add_alpha) o[samples] = is_key ? 0 : 255;
            }
    }
    width = w;
    height = h;
    return true;
}

// -------------------------------------------------------------------------------------
// SceneLoader
// -------------------------------------------------------------------------------------
HostScene SceneLoader::load(std::string &path) {
    size_t typePos = path.find_last_of('.');
    if (typePos == std::string::npos) throw std::runtime_error("Wrong file name privided: " + path);
    std::string type = path.substr(typePos + 1);
    for (auto &c : type) c = (char)tolower(c);
    HostScene scene;
    if (type == "glb") scene = loadGLTF(path, true);
    else if (type == "gltf") scene = loadGLTF(path, false);
    else if (type == "obj") scene = loadOBJ(path);
    else if (type == "ptscene") scene = read_ptscene(path);
    else throw std::runtime_error("Unsupported file type" + path);
    return scene;
}

HostScene SceneLoader::loadGLTF(const std::string &path, bool binary) {
    Gltf g;
    size_t slash = path.find_last_of('/');
    g.dir = slash == std::string::npos ? std::string(".") : path.substr(0, slash);
    std::vector<unsigned char> file = read_file(path);
    std::vector<unsigned char> bin_chunk;
    if (binary) {
        if (file.size() < 20 || memcmp(file.data(), "glTF", 4) != 0) throw std::runtime_error("Not a GLB file: " + path);
        uint32_t total;
        memcpy(&total, file.data() + 8, 4);
        size_t pos = 12;
        bool have_json = false;
        while (pos + 8 <= file.size() && pos + 8 <= total) {
            uint32_t clen, ctype;
            memcpy(&clen, file.data() + pos, 4);
            memcpy(&ctype, file.data() + pos + 4, 4);
            if (pos + 8 + clen > file.size()) throw std::runtime_error("GLB chunk overruns file: " + path);
            if (ctype == 0x4E4F534A) {  // JSON
                JParser jp((const char *)file.data() + pos + 8, clen);
                g.root = jp.parse();
                have_json = true;
            } else if (ctype == 0x004E4942) {  // BIN
                bin_chunk.assign(file.begin() + (long)pos + 8, file.begin() + (long)pos + 8 + clen);
            }
            pos += 8 + (size_t)clen;
        }
        if (!have_json) throw std::runtime_error("GLB without JSON chunk: " + path);
    } else {
        JParser jp((const char *)file.data(), file.size());
        g.root = jp.parse();
    }
    const JVal &buffers = g.top("buffers");
    for (size_t i = 0; i < buffers.size(); i++) {
        const JVal &b = buffers.at(i);
        if (const JVal *uri = b.get("uri")) {
            const std::string &u = uri->str;
            if (u.rfind("data:", 0) == 0) {
                size_t comma = u.find(',');
                g.buffers.push_back(base64_decode(u.substr(comma == std::string::npos ? 0 : comma + 1)));
            } else {
                g.buffers.push_back(read_file(g.dir + "/" + u));
            }
        } else {
            g.buffers.push_back(bin_chunk);
        }
    }

    HostScene scene;

    // --- textures: one entry per glTF image that a material's baseColor/emissive slot uses
    std::map<int, int> imageToTex;
    auto textureForSlot = [&](const JVal *slot) -> std::optional<int> {
        if (!slot || !slot->get("index")) return std::nullopt;
        int ti = slot->need("index").integer(-1);
        const JVal &textures = g.top("textures");
        if (ti < 0 || (size_t)ti >= textures.size()) return std::nullopt;
        const JVal *src = textures.at((size_t)ti).get("source");
        if (!src) return std::nullopt;
        int img = src->integer(-1);
        auto it = imageToTex.find(img);
        if (it != imageToTex.end()) return it->second;
        const JVal &image = g.top("images").at((size_t)img);
        HostTexture tex;
        if (const JVal *bv = image.get("bufferView")) {
            Gltf::View v = g.bufferView(bv->integer(0));
            tex = texture_from_bytes(v.ptr, v.len, "*" + std::to_string(img));
        } else if (const JVal *uri = image.get("uri")) {
            if (uri->str.rfind("data:", 0) == 0) {
                size_t comma = uri->str.find(',');
                auto bytes = base64_decode(uri->str.substr(comma == std::string::npos ? 0 : comma + 1));
                tex = texture_from_bytes(bytes.data(), bytes.size(), "data-uri");
            } else {
                auto bytes = read_file(g.dir + "/" + uri->str);
                tex = texture_from_bytes(bytes.data(), bytes.size(), uri->str);
            }
        } else {
            return std::nullopt;
        }
        int idx = (int)scene.textures.size();
        scene.textures.push_back(std::move(tex));
        imageToTex[img] = idx;
        return idx;
    };

    // --- materials (reference src/HostScene.cpp:145-190 through assimp's glTF2 importer)
    const JVal &materials = g.top("materials");
    for (size_t i = 0; i < materials.size(); i++) {
        const JVal &m = materials.at(i);
        HostMaterial hm;
        if (const JVal *nm = m.get("name")) hm.name = nm->str;
        hm.baseColor = make_float3(1.f, 1.f, 1.f);
        hm.emissiveFactor = make_float3(0.f, 0.f, 0.f);
        if (const JVal *pbr = m.get("pbrMetallicRoughness")) {
            if (const JVal *f = pbr->get("baseColorFactor"))
                if (f->size() >= 3) hm.baseColor = make_float3((float)f->at(0).num, (float)f->at(1).num, (float)f->at(2).num);
            hm.baseColorTextureIdx = textureForSlot(pbr->get("baseColorTexture"));
        }
        if (const JVal *e = m.get("emissiveFactor"))
            if (e->size() >= 3) hm.emissiveFactor = make_float3((float)e->at(0).num, (float)e->at(1).num, (float)e->at(2).num);
        hm.emissiveTextureIdx = textureForSlot(m.get("emissiveTexture"));
        hm.type = UNIVERSAL;  // the glTF path of the reference only ever builds UniversalMaterial
        scene.materials.push_back(std::move(hm));
    }
    bool needDefaultMaterial = false;
    int defaultMaterial = (int)scene.materials.size();

    // --- triangles, bucketed per material (PreTransformVertices output order)
    std::vector<std::vector<Triangle>> perMaterial(scene.materials.size() + 1);
    const JVal &nodes = g.top("nodes");
    const JVal &meshes = g.top("meshes");

    struct Frame { int node; Mat4 parent; };
    std::vector<int> roots;
    {
        int sceneIdx = g.root->get("scene") ? g.root->get("scene")->integer(0) : 0;
        const JVal &scenes = g.top("scenes");
        if (scenes.size() > 0) {
            const JVal *ns = scenes.at((size_t)std::min<int>(sceneIdx, (int)scenes.size() - 1)).get("nodes");
            if (ns)
                for (size_t i = 0; i < ns->size(); i++) roots.push_back(ns->at(i).integer(0));
        } else {
            for (size_t i = 0; i < nodes.size(); i++) roots.push_back((int)i);
        }
    }
    // depth-first, children in order
    std::vector<Frame> stack;
    for (auto it = roots.rbegin(); it != roots.rend(); ++it) stack.push_back({*it, Mat4::identity()});
    size_t guard = 0;
    while (!stack.empty()) {
        // a glTF node graph is a forest (every node has at most one parent): more visits than nodes means a cycle or a shared subtree
        if (++guard > nodes.size() * 4 + 64) throw std::runtime_error("glTF: node graph is cyclic (or shares subtrees)");
        Frame fr = stack.back();
        stack.pop_back();
        const JVal &node = nodes.at((size_t)fr.node);
        Mat4 local = Mat4::identity();
        if (const JVal *mat = node.get("matrix")) {
            if (mat->size() == 16)
                for (int c = 0; c < 4; c++)
                    for (int r = 0; r < 4; r++) local.m[r][c] = (float)mat->at((size_t)(c * 4 + r)).num;  // column-major
        } else {
            if (const JVal *t = node.get("translation")) {
                Mat4 tm = Mat4::identity();
                tm.m[0][3] = (float)t->at(0).num; tm.m[1][3] = (float)t->at(1).num; tm.m[2][3] = (float)t->at(2).num;
                local = mul(local, tm);
            }
            if (const JVal *r = node.get("rotation"))
                local = mul(local, quat_matrix((float)r->at(0).num, (float)r->at(1).num, (float)r->at(2).num, (float)r->at(3).num));
            if (const JVal *s = node.get("scale")) {
                Mat4 sm = Mat4::identity();
                sm.m[0][0] = (float)s->at(0).num; sm.m[1][1] = (float)s->at(1).num; sm.m[2][2] = (float)s->at(2).num;
                local = mul(local, sm);
            }
        }
        Mat4 world = mul(fr.parent, local);
        if (const JVal *mi = node.get("mesh")) {
            const JVal &mesh = meshes.at((size_t)mi->integer(0));
            const JVal *prims = mesh.get("primitives");
            for (size_t p = 0; prims && p < prims->size(); p++) {
                const JVal &prim = prims->at(p);
                int mode = prim.get("mode") ? prim.get("mode")->integer(4) : 4;
                if (mode != 4 && mode != 5 && mode != 6) continue;  // points/lines are removed (SortByPType)
                const JVal *attrs = prim.get("attributes");
                if (!attrs || !attrs->get("POSITION")) continue;
                std::vector<float> pos = g.readFloats(attrs->need("POSITION").integer(0), 3);
                std::vector<float> uv;
                if (attrs->get("TEXCOORD_0")) uv = g.readFloats(attrs->get("TEXCOORD_0")->integer(0), 2);
                size_t nv = pos.size() / 3;
                if (!uv.empty() && uv.size() < nv * 2) throw std::runtime_error("glTF: TEXCOORD_0 has fewer elements than POSITION");
                std::vector<uint32_t> idx;
                if (prim.get("indices")) idx = g.readIndices(prim.get("indices")->integer(0));
                else {
                    idx.resize(nv);
                    for (size_t i = 0; i < nv; i++) idx[i] = (uint32_t)i;
                }
                int matIdx = prim.get("material") ? prim.get("material")->integer(defaultMaterial) : defaultMaterial;
                if (matIdx < 0 || matIdx >= defaultMaterial) { matIdx = defaultMaterial; needDefaultMaterial = true; }
                std::vector<Vertex> verts(nv);
                for (size_t i = 0; i < nv; i++) {
                    verts[i].position = xform(world, make_float3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
                    verts[i].texCoords = uv.empty() ? make_float2(0.f, 0.f) : make_float2(uv[2 * i], 1.f - uv[2 * i + 1]);
                }
                auto emit = [&](uint32_t a, uint32_t b, uint32_t c) {
                    if (a >= nv || b >= nv || c >= nv) throw std::runtime_error("glTF: vertex index out of range");
                    if (a == b || b == c || a == c) return;
                    if (same_pos(verts[a].position, verts[b].position) || same_pos(verts[b].position, verts[c].position) ||
                        same_pos(verts[a].position, verts[c].position))
                        return;
                    Triangle t;
                    t.v0 = verts[a]; t.v1 = verts[b]; t.v2 = verts[c];
                    t.materialIdx = matIdx;
                    t.textureIdx = -1;
                    perMaterial[(size_t)matIdx].push_back(t);
                };
                if (mode == 4) for (size_t i = 0; i + 2 < idx.size(); i += 3) emit(idx[i], idx[i + 1], idx[i + 2]);
                else if (mode == 5) for (size_t i = 0; i + 2 < idx.size(); i++) (i & 1) ? emit(idx[i + 1], idx[i], idx[i + 2]) : emit(idx[i], idx[i + 1], idx[i + 2]);
                else for (size_t i = 1; i + 1 < idx.size(); i++) emit(idx[0], idx[i], idx[i + 1]);
            }
        }
        if (const JVal *ch = node.get("children"))
            for (size_t i = ch->size(); i-- > 0;) stack.push_back({ch->at(i).integer(0), world});
    }
    if (needDefaultMaterial) {
        HostMaterial dm;  // assimp's default material: grey 0.6 diffuse, no emission
        dm.baseColor = make_float3(1.f, 1.f, 1.f);
        dm.name = "DefaultMaterial";
        scene.materials.push_back(dm);
    }
    for (auto &bucket : perMaterial)
        scene.triangles.insert(scene.triangles.end(), bucket.begin(), bucket.end());
    return scene;
}

// -------------------------------------------------------------------------------------
// OBJ / MTL  (routing per reference README.md:60-76 and src/obj_loader.h:65-96:
// lambertian→Ka, metal→Ka+Ns, dielectric→Ni, diffuse_light→Kd; anything else is a
// UniversalMaterial with Kd as base colour and Ke as emission)
// -------------------------------------------------------------------------------------
HostScene SceneLoader::loadOBJ(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Cannot open file: " + path);
    size_t slash = path.find_last_of('/');
    std::string dir = slash == std::string::npos ? std::string(".") : path.substr(0, slash);

    struct Mtl { std::string name; float3 Ka{0, 0, 0}, Kd{0.6f, 0.6f, 0.6f}, Ke{0, 0, 0}; float Ns = 0.f, Ni = 1.5f; std::string map_Kd, map_Ke; };
    std::vector<Mtl> mtls;
    auto loadMtl = [&](const std::string &file) {
        std::ifstream m(dir + "/" + file);
        if (!m) return;
        std::string line;
        while (std::getline(m, line)) {
            std::istringstream ss(line);
            std::string k;
            if (!(ss >> k)) continue;
            if (k == "newmtl") { mtls.emplace_back(); ss >> mtls.back().name; }
            else if (mtls.empty()) continue;
            else if (k == "Ka") ss >> mtls.back().Ka.x >> mtls.back().Ka.y >> mtls.back().Ka.z;
            else if (k == "Kd") ss >> mtls.back().Kd.x >> mtls.back().Kd.y >> mtls.back().Kd.z;
            else if (k == "Ke") ss >> mtls.back().Ke.x >> mtls.back().Ke.y >> mtls.back().Ke.z;
            else if (k == "Ns") ss >> mtls.back().Ns;
            else if (k == "Ni") ss >> mtls.back().Ni;
            else if (k == "map_Kd" || k == "map_Ke") {  // the file name is the last word (options such as -s 1 1 1 may precede it)
                std::string word, file;
                while (ss >> word) file = word;
                (k == "map_Kd" ? mtls.back().map_Kd : mtls.back().map_Ke) = file;
            }
        }
    };

    HostScene scene;
    std::vector<float3> P;
    std::vector<float2> T;
    std::map<std::string, int> matIndex;
    int curMat = -1;
    // Texture files of the .mtl, decoded like the reference's stbi_load(path, ..., 3) (src/HostScene.cpp:29: PNG / JPEG / BMP / TGA here).
    // assimp hands an OBJ's map_Kd over as aiTextureType_DIFFUSE and its map_Ke as aiTextureType_EMISSIVE; the reference loads every
    // texture (loadTextures, :52-72) but binds only BASE_COLOR and EMISSIVE (processMaterial, :174-184) — so map_Ke becomes the emissive
    // texture, map_Kd takes a slot in the texture list and is bound to nothing.  A missing file aborts the load, as there.
    std::map<std::string, int> texIndex;
    auto textureFor = [&](const std::string &file) -> int {
        auto it = texIndex.find(file);
        if (it != texIndex.end()) return it->second;
        std::vector<unsigned char> bytes;
        try {
            bytes = read_file(dir + "/" + file);
        } catch (const std::exception &) {
            throw std::runtime_error("Cannot load texture data, path: " + file);
        }
        const int idx = (int)scene.textures.size();
        scene.textures.push_back(texture_from_bytes(bytes.data(), bytes.size(), file));
        texIndex[file] = idx;
        return idx;
    };
    auto materialFor = [&](const std::string &name) -> int {
        auto it = matIndex.find(name);
        if (it != matIndex.end()) return it->second;
        HostMaterial hm;
        hm.name = name;
        hm.type = type_from_name(name);
        const Mtl *src = nullptr;
        for (auto &m : mtls) if (m.name == name) src = &m;
        Mtl dflt;
        if (!src) src = &dflt;
        switch (hm.type) {
            case LAMBERTIAN: hm.baseColor = src->Ka; break;
            case METAL: hm.baseColor = src->Ka; hm.fuzz = src->Ns; break;
            case DIELECTRIC: hm.baseColor = make_float3(1, 1, 1); hm.ior = src->Ni; break;
            case DIFFUSE_LIGHT: hm.baseColor = make_float3(0, 0, 0); hm.emissiveFactor = src->Kd; break;
            default: hm.baseColor = src->Kd; hm.emissiveFactor = src->Ke; break;
        }
        if (!src->map_Kd.empty()) textureFor(src->map_Kd);
        if (!src->map_Ke.empty()) hm.emissiveTextureIdx = textureFor(src->map_Ke);
        int idx = (int)scene.materials.size();
        scene.materials.push_back(hm);
        matIndex[name] = idx;
        return idx;
    };

    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::string k;
        if (!(ss >> k)) continue;
        if (k == "v") { float3 p; ss >> p.x >> p.y >> p.z; P.push_back(p); }
        else if (k == "vt") { float2 t{0, 0}; ss >> t.x >> t.y; T.push_back(t); }
        else if (k == "mtllib") { std::string file; ss >> file; loadMtl(file); }
        else if (k == "usemtl") { std::string name; ss >> name; curMat = materialFor(name); }
        else if (k == "f") {
            std::vector<Vertex> poly;
            std::string tok;
            while (ss >> tok) {
                int vi = 0, ti = 0;
                size_t s1 = tok.find('/');
                vi = atoi(tok.substr(0, s1).c_str());
                if (s1 != std::string::npos) {
                    size_t s2 = tok.find('/', s1 + 1);
                    std::string t = tok.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
                    if (!t.empty()) ti = atoi(t.c_str());
                }
                if (vi < 0) vi = (int)P.size() + vi + 1;
                if (ti < 0) ti = (int)T.size() + ti + 1;
                if (vi < 1 || vi > (int)P.size()) throw std::runtime_error("OBJ: vertex index out of range in " + path);
                Vertex v;
                v.position = P[(size_t)vi - 1];
                v.texCoords = (ti >= 1 && ti <= (int)T.size()) ? T[(size_t)ti - 1] : make_float2(0, 0);
                poly.push_back(v);
            }
            if (poly.size() < 3) continue;  // points / lines are dropped, as SortByPType does
            if (curMat < 0) curMat = materialFor("DefaultMaterial");
            for (size_t i = 1; i + 1 < poly.size(); i++) {
                if (same_pos(poly[0].position, poly[i].position) || same_pos(poly[i].position, poly[i + 1].position) ||
                    same_pos(poly[0].position, poly[i + 1].position))
                    continue;
                Triangle t;
                t.v0 = poly[0]; t.v1 = poly[i]; t.v2 = poly[i + 1];
                t.materialIdx = curMat;
                scene.triangles.push_back(t);
            }
        }
    }
    return scene;
}

// -------------------------------------------------------------------------------------
// .ptscene
// -------------------------------------------------------------------------------------
namespace {
struct PtsHeader { char magic[4]; uint32_t version, n_tris, n_spheres, n_mats, n_tex; };
struct PtsTri { float pos[9]; float uv[6]; int32_t mat; int32_t tex; };
struct PtsSphere { float c[3]; float r; int32_t mat; };
struct PtsMat { int32_t type; float base[3]; float emis[3]; int32_t base_tex; int32_t emis_tex; float fuzz; float ior; };
}  // namespace

void write_ptscene(const HostScene &s, const std::string &path) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Cannot write file: " + path);
    PtsHeader h{{'P', 'T', 'S', 'C'}, 1, (uint32_t)s.triangles.size(), (uint32_t)s.spheres.size(), (uint32_t)s.materials.size(), (uint32_t)s.textures.size()};
    f.write((const char *)&h, sizeof h);
    for (auto &t : s.triangles) {
        PtsTri r;
        const Vertex *v[3] = {&t.v0, &t.v1, &t.v2};
        for (int k = 0; k < 3; k++) {
            r.pos[3 * k] = v[k]->position.x; r.pos[3 * k + 1] = v[k]->position.y; r.pos[3 * k + 2] = v[k]->position.z;
            r.uv[2 * k] = v[k]->texCoords.x; r.uv[2 * k + 1] = v[k]->texCoords.y;
        }
        r.mat = t.materialIdx;
        r.tex = t.textureIdx;
        f.write((const char *)&r, sizeof r);
    }
    for (auto &sp : s.spheres) {
        PtsSphere r{{sp.center.x, sp.center.y, sp.center.z}, sp.radius, sp.materialIdx};
        f.write((const char *)&r, sizeof r);
    }
    for (auto &m : s.materials) {
        PtsMat r{(int32_t)m.type, {m.baseColor.x, m.baseColor.y, m.baseColor.z}, {m.emissiveFactor.x, m.emissiveFactor.y, m.emissiveFactor.z},
                 m.baseColorTextureIdx.value_or(-1), m.emissiveTextureIdx.value_or(-1), m.fuzz, m.ior};
        f.write((const char *)&r, sizeof r);
    }
    for (auto &t : s.textures) {
        int32_t wh[2] = {t.width, t.height};
        f.write((const char *)wh, sizeof wh);
        std::vector<unsigned char> px((size_t)t.width * (size_t)t.height * 3);
        for (size_t i = 0; i < t.data.size(); i++) {
            px[3 * i] = (unsigned char)t.data[i].x; px[3 * i + 1] = (unsigned char)t.data[i].y; px[3 * i + 2] = (unsigned char)t.data[i].z;
        }
        f.write((const char *)px.data(), (std::streamsize)px.size());
    }
}

HostScene read_ptscene(const std::string &path) {
    std::vector<unsigned char> buf = read_file(path);
    size_t pos = 0;
    auto take = [&](void *dst, size_t n) {
        if (pos + n > buf.size()) throw std::runtime_error("ptscene truncated: " + path);
        memcpy(dst, buf.data() + pos, n);
        pos += n;
    };
    PtsHeader h;
    take(&h, sizeof h);
    if (memcmp(h.magic, "PTSC", 4) != 0 || h.version != 1) throw std::runtime_error("not a ptscene v1 file: " + path);
    // the counts are checked against the bytes that are really there BEFORE anything is allocated (a corrupted header must raise, not reserve
    // 100 GB), and every index against the table it points into
    const uint64_t fixed = (uint64_t)h.n_tris * sizeof(PtsTri) + (uint64_t)h.n_spheres * sizeof(PtsSphere) + (uint64_t)h.n_mats * sizeof(PtsMat) + (uint64_t)h.n_tex * 8u;
    if (fixed > buf.size() - pos) throw std::runtime_error("ptscene truncated (or its header is corrupt): " + path);
    auto checkIndex = [&](int32_t idx, uint32_t n, bool optional, const char *what) {
        if ((optional && idx < 0) || (idx >= 0 && (uint32_t)idx < n)) return;
        throw std::runtime_error(std::string("ptscene: ") + what + " index out of range: " + path);
    };
    HostScene s;
    s.triangles.resize(h.n_tris);
    for (auto &t : s.triangles) {
        PtsTri r;
        take(&r, sizeof r);
        checkIndex(r.mat, h.n_mats, false, "triangle material");
        Vertex *v[3] = {&t.v0, &t.v1, &t.v2};
        for (int k = 0; k < 3; k++) {
            v[k]->position = make_float3(r.pos[3 * k], r.pos[3 * k + 1], r.pos[3 * k + 2]);
            v[k]->texCoords = make_float2(r.uv[2 * k], r.uv[2 * k + 1]);
        }
        t.materialIdx = r.mat;
        t.textureIdx = r.tex;
    }
    s.spheres.resize(h.n_spheres);
    for (auto &sp : s.spheres) {
        PtsSphere r;
        take(&r, sizeof r);
        checkIndex(r.mat, h.n_mats, false, "sphere material");
        sp.center = make_float3(r.c[0], r.c[1], r.c[2]);
        sp.radius = r.r;
        sp.materialIdx = r.mat;
    }
    s.materials.resize(h.n_mats);
    for (auto &m : s.materials) {
        PtsMat r;
        take(&r, sizeof r);
        if (r.type < 0 || r.type > 4) throw std::runtime_error("ptscene: unknown material type: " + path);
        checkIndex(r.base_tex, h.n_tex, true, "base-colour texture");
        checkIndex(r.emis_tex, h.n_tex, true, "emissive texture");
        m.type = (material_type)r.type;
        m.baseColor = make_float3(r.base[0], r.base[1], r.base[2]);
        m.emissiveFactor = make_float3(r.emis[0], r.emis[1], r.emis[2]);
        if (r.base_tex >= 0) m.baseColorTextureIdx = r.base_tex;
        if (r.emis_tex >= 0) m.emissiveTextureIdx = r.emis_tex;
        m.fuzz = r.fuzz;
        m.ior = r.ior;
    }
    s.textures.resize(h.n_tex);
    for (auto &t : s.textures) {
        int32_t wh[2];
        take(wh, sizeof wh);
        if (wh[0] < 0 || wh[1] < 0 || (uint64_t)wh[0] * (uint64_t)wh[1] * 3u > buf.size() - pos) throw std::runtime_error("ptscene: bad texture size: " + path);
        t.width = wh[0];
        t.height = wh[1];
        std::vector<unsigned char> px((size_t)t.width * (size_t)t.height * 3);
        take(px.data(), px.size());
        t.data.resize((size_t)t.width * (size_t)t.height);
        for (size_t i = 0; i < t.data.size(); i++) t.data[i] = make_float3((float)px[3 * i], (float)px[3 * i + 1], (float)px[3 * i + 2]);
    }
    return s;
}
