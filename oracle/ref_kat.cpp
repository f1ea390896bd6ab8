// ref_kat.cpp — known-answer dumper over the reference's own headers (host-compiled).
//
// TEST INFRASTRUCTURE (oracle/).  Prints one JSON document with inputs and the values the
// UNMODIFIED reference functions return for them; oracle/make_golden.py stores it as
// tests/golden/ref_kats.json, and tests/test_oracle_kat.py replays the inputs through
// oracle/pt_oracle.c (bit-exact) — that is what pins the restatement.  The reference has no
// tests or golden vectors of its own, so these are manufactured from its code.
#include <float.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "ray.h"
#include "helper_math.h"
#include "hitable_list.h"
#include "sphere.h"
#include "hitable.h"
#include "camera.h"
#include "material.h"
#include "triangle.h"
#include "bvh.h"
#include "pdf.h"

#include <cstdio>
#include <string>
#include <vector>

bool refract(const float3 &v, const float3 &n, float ni_over_nt, float3 &refracted);  // helper_math.cu

static uint32_t lcg_state = 12345u;
static float frand() {  // deterministic inputs in [0,1)
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return (float)(lcg_state >> 8) * (1.0f / 16777216.0f);
}
static float frange(float a, float b) { return a + (b - a) * frand(); }

static void pf(float f) {
    if (f != f) printf("\"nan\"");
    else if (f > FLT_MAX) printf("\"inf\"");
    else if (f < -FLT_MAX) printf("\"-inf\"");
    else printf("%.9g", f);
}
static void p3(float3 v) { printf("["); pf(v.x); printf(", "); pf(v.y); printf(", "); pf(v.z); printf("]"); }
static void key(const char *k) { printf("\"%s\": ", k); }

int main() {
    printf("{\n");

    // ---------------- RNG ----------------
    key("rng"); printf("[\n");
    unsigned long long seeds[] = {1984ULL, 1985ULL, 1984ULL + 2073599ULL, 0ULL, (1ULL << 32) + 5ULL, 1984ULL + 8294399ULL};
    for (size_t i = 0; i < sizeof seeds / sizeof *seeds; i++) {
        curandState s;
        curand_init(seeds[i], 0, 0, &s);
        printf("  {\"seed\": %llu, \"d\": %u, \"v\": [%u, %u, %u, %u, %u], ", seeds[i], s.d, s.v[0], s.v[1], s.v[2], s.v[3], s.v[4]);
        curandState a = s;
        printf("\"raw\": [");
        for (int k = 0; k < 8; k++) printf("%s%u", k ? ", " : "", curand(&a));
        printf("], \"uniform\": [");
        a = s;
        for (int k = 0; k < 8; k++) { if (k) printf(", "); pf(curand_uniform(&a)); }
        printf("]}%s\n", i + 1 < sizeof seeds / sizeof *seeds ? "," : "");
    }
    printf("],\n");

    // ---------------- samplers ----------------
    key("random_cosine_direction"); printf("[\n");
    for (int i = 0; i < 4; i++) {
        curandState s;
        curand_init(1984 + 1000 * i, 0, 0, &s);
        printf("  {\"seed\": %d, \"out\": [", 1984 + 1000 * i);
        for (int k = 0; k < 4; k++) { if (k) printf(", "); p3(random_cosine_direction(&s)); }
        printf("]}%s\n", i < 3 ? "," : "");
    }
    printf("],\n");
    key("random_in_unit_sphere"); printf("[\n");
    for (int i = 0; i < 4; i++) {
        curandState s;
        curand_init(77 + i, 0, 0, &s);
        printf("  {\"seed\": %d, \"out\": [", 77 + i);
        for (int k = 0; k < 4; k++) { if (k) printf(", "); p3(random_in_unit_sphere(&s)); }
        // how many draws did that take (to tell the argument evaluation order of make_random_float3 apart)
        curandState t;
        curand_init(77 + i, 0, 0, &t);
        float first3[3] = {curand_uniform(&t), curand_uniform(&t), curand_uniform(&t)};
        printf("], \"first3_uniforms\": ["); pf(first3[0]); printf(", "); pf(first3[1]); printf(", "); pf(first3[2]);
        printf("]}%s\n", i < 3 ? "," : "");
    }
    printf("],\n");

    // ---------------- onb ----------------
    key("onb"); printf("[\n");
    {
        std::vector<float3> ns = {make_float3(0, 1, 0), make_float3(1, 0, 0), make_float3(0, 0, -1), make_float3(0.95f, 0.1f, 0.2f),
                                  make_float3(-3.f, 4.f, 12.f), make_float3(0.89f, 0.3f, -0.3f)};
        for (int i = 0; i < 10; i++) ns.push_back(make_float3(frange(-2, 2), frange(-2, 2), frange(-2, 2)));
        for (size_t i = 0; i < ns.size(); i++) {
            onb o(ns[i]);
            float3 l = o.local(make_float3(0.3f, -0.7f, 0.5f));
            printf("  {\"n\": "); p3(ns[i]); printf(", \"u\": "); p3(o.u()); printf(", \"v\": "); p3(o.v()); printf(", \"w\": "); p3(o.w());
            printf(", \"local\": "); p3(l); printf("}%s\n", i + 1 < ns.size() ? "," : "");
        }
    }
    printf("],\n");

    // ---------------- camera ----------------
    key("camera"); printf("[\n");
    {
        struct C { float lf[3], fr[3], vf, hf; } cams[] = {
            {{0, 0, 0.5f}, {0, 0, -0.5f}, 45.f, 45.f},
            {{1.5f, 2.f, -3.f}, {0.3f, -0.2f, 0.9f}, 60.f, 90.f},
            {{-20.f, 100.f, 400.f}, {0.1f, -0.1f, -1.f}, 30.f, 50.f}};
        float uvs[][2] = {{0.25f, 0.75f}, {0.f, 0.f}, {1.f, 1.f}, {0.5f, 0.5f}, {0.123456f, 0.987654f}};
        int n = 0;
        for (auto &c : cams)
            for (auto &uv : uvs) {
                camera cam;
                CameraConfig cfg(make_float3(c.lf[0], c.lf[1], c.lf[2]), make_float3(c.fr[0], c.fr[1], c.fr[2]), c.vf, c.hf);
                cam.recalculate_camera_params(cfg);
                ray r = cam.get_ray(uv[0], uv[1]);
                printf("%s  {\"look_from\": [%.9g, %.9g, %.9g], \"front\": [%.9g, %.9g, %.9g], \"vfov\": %.9g, \"hfov\": %.9g, \"u\": %.9g, \"v\": %.9g, \"o\": ",
                       n++ ? ",\n" : "", c.lf[0], c.lf[1], c.lf[2], c.fr[0], c.fr[1], c.fr[2], c.vf, c.hf, uv[0], uv[1]);
                p3(r.origin()); printf(", \"d\": "); p3(r.direction()); printf("}");
            }
        printf("\n");
    }
    printf("],\n");

    // ---------------- triangle ----------------
    key("triangle"); printf("[\n");
    {
        UniversalMaterial mat(make_float3(1, 1, 1), nullptr, make_float3(0, 0, 0), nullptr);
        int n = 0;
        auto emit = [&](Vertex a, Vertex b, Vertex c, float3 o, float3 d, unsigned long long seed) {
            triangle t(a, b, c, &mat);
            hit_record rec;
            rec.t = 0; rec.p = rec.normal = make_float3(0, 0, 0); rec.texCoord = make_float2(0, 0);
            bool h = t.hit(ray(o, d), interval(0.001f, FLT_MAX), rec);
            float pv = t.pdf_value(o, d);
            curandState s;
            curand_init(seed, 0, 0, &s);
            float3 rnd = t.random(o, &s);
            printf("%s  {\"pos\": [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g], \"uv\": [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g], \"o\": ",
                   n++ ? ",\n" : "", a.position.x, a.position.y, a.position.z, b.position.x, b.position.y, b.position.z, c.position.x, c.position.y, c.position.z,
                   a.texCoords.x, a.texCoords.y, b.texCoords.x, b.texCoords.y, c.texCoords.x, c.texCoords.y);
            p3(o); printf(", \"d\": "); p3(d);
            printf(", \"hit\": %d, \"t\": ", h ? 1 : 0); pf(h ? rec.t : 0.f);
            printf(", \"p\": "); p3(h ? rec.p : make_float3(0, 0, 0));
            printf(", \"normal\": "); p3(h ? rec.normal : make_float3(0, 0, 0));
            printf(", \"tex\": ["); pf(h ? rec.texCoord.x : 0.f); printf(", "); pf(h ? rec.texCoord.y : 0.f);
            printf("], \"area\": "); pf(t.area); printf(", \"pdf_value\": "); pf(pv);
            printf(", \"seed\": %llu, \"random\": ", seed); p3(rnd); printf("}");
        };
        Vertex a{make_float3(0, 0, -2), make_float2(0, 0)}, b{make_float3(1, 0, -2), make_float2(1, 0)}, c{make_float3(0, 1, -2), make_float2(0, 1)};
        emit(a, b, c, make_float3(.25f, .25f, 0), make_float3(0, 0, -.5f), 1984);
        emit(a, b, c, make_float3(.25f, .25f, 0), make_float3(0, 0, .5f), 1985);
        emit(a, b, c, make_float3(2.f, 2.f, 0), make_float3(0, 0, -1.f), 1986);
        emit(a, b, c, make_float3(0.f, 0.f, 0), make_float3(0, 0, -1.f), 1987);      // through a vertex
        emit(a, b, c, make_float3(0.5f, 0.5f, 0), make_float3(0, 0, -1.f), 1988);    // on the hypotenuse
        emit(a, b, c, make_float3(.25f, .25f, -2.0005f), make_float3(0, 0, 1.f), 1989);  // inside t_min
        for (int i = 0; i < 40; i++) {
            Vertex v[3];
            for (auto &x : v) { x.position = make_float3(frange(-300, 300), frange(-300, 300), frange(-1200, -600)); x.texCoords = make_float2(frand(), frand()); }
            float3 o = make_float3(frange(-50, 50), frange(-50, 50), frange(0, 1));
            float w0 = frand(), w1 = frand() * (1 - w0);
            float3 target = v[0].position * (1 - w0 - w1) + v[1].position * w0 + v[2].position * w1;
            float3 d = (i % 4 == 3) ? make_float3(frange(-1, 1), frange(-1, 1), -1.f) : (target - o) * frange(0.001f, 2.f);
            emit(v[0], v[1], v[2], o, d, 5000 + (unsigned long long)i);
        }
        printf("\n");
    }
    printf("],\n");

    // ---------------- sphere ----------------
    key("sphere"); printf("[\n");
    {
        int n = 0;
        for (int i = 0; i < 24; i++) {
            float3 cen = make_float3(frange(-5, 5), frange(-5, 5), frange(-15, -5));
            float rad = (i == 0) ? 1000.f : frange(0.2f, 3.f);
            float3 o = (i % 5 == 4) ? cen + make_float3(0.1f, 0.05f, -0.02f) * rad : make_float3(frange(-1, 1), frange(-1, 1), frange(0, 2));
            float3 tgt = cen + make_float3(frange(-1, 1), frange(-1, 1), frange(-1, 1)) * rad * ((i % 3 == 2) ? 1.5f : 0.6f);
            float3 d = (tgt - o) * frange(0.1f, 1.5f);
            sphere s(cen, rad, nullptr);
            hit_record rec;
            rec.t = 0; rec.p = rec.normal = make_float3(0, 0, 0);
            bool h = s.hit(ray(o, d), interval(0.001f, FLT_MAX), rec);
            printf("%s  {\"sph\": [%.9g, %.9g, %.9g, %.9g], \"o\": ", n++ ? ",\n" : "", cen.x, cen.y, cen.z, rad);
            p3(o); printf(", \"d\": "); p3(d); printf(", \"hit\": %d, \"t\": ", h ? 1 : 0); pf(h ? rec.t : 0.f);
            printf(", \"p\": "); p3(h ? rec.p : make_float3(0, 0, 0)); printf(", \"normal\": "); p3(h ? rec.normal : make_float3(0, 0, 0)); printf("}");
        }
        printf("\n");
    }
    printf("],\n");

    // ---------------- texture ----------------
    key("texture"); printf("{");
    {
        const int TW = 5, TH = 4;
        std::vector<float3> data(TW * TH);
        for (int i = 0; i < TW * TH; i++) data[i] = make_float3((float)((i * 37) % 256), (float)((i * 91 + 13) % 256), (float)((i * 53 + 101) % 256));
        BaseColorTexture tex(TW, TH, data.data());
        printf("\"width\": %d, \"height\": %d, \"data\": [", TW, TH);
        for (int i = 0; i < TW * TH; i++) printf("%s%.9g, %.9g, %.9g", i ? ", " : "", data[i].x, data[i].y, data[i].z);
        printf("], \"cases\": [\n");
        // v chosen so that the flipped row stays inside the image (row == height is the reference's out-of-bounds read)
        float uvs[][2] = {{0.1f, 0.3f}, {0.5f, 0.5f}, {0.99f, 0.99f}, {1.25f, 0.75f}, {2.5f, 3.3f}, {-0.3f, 0.6f}, {0.f, 0.26f}, {0.999999f, 0.25f}, {0.2f, 0.999f}};
        for (size_t i = 0; i < sizeof uvs / sizeof *uvs; i++) {
            float3 val = tex.value(make_float2(uvs[i][0], uvs[i][1]), make_float3(0, 0, 0));
            printf("  {\"u\": %.9g, \"v\": %.9g, \"value\": ", uvs[i][0], uvs[i][1]); p3(val);
            printf("}%s\n", i + 1 < sizeof uvs / sizeof *uvs ? "," : "");
        }
        printf("]");
    }
    printf("},\n");

    // ---------------- pdfs ----------------
    key("pdf"); printf("[\n");
    {
        int n = 0;
        for (int i = 0; i < 24; i++) {
            float3 nrm = normalize(make_float3(frange(-1, 1), frange(-1, 1), frange(-1, 1)));
            float3 dir = make_float3(frange(-3, 3), frange(-3, 3), frange(-3, 3));
            cosine_pdf cp(nrm);
            UniversalMaterial mat(make_float3(1, 1, 1), nullptr, make_float3(0, 0, 0), nullptr);
            hit_record rec;
            rec.normal = nrm;
            float sp = mat.scattering_pdf(ray(), rec, ray(make_float3(0, 0, 0), dir));
            printf("%s  {\"normal\": ", n++ ? ",\n" : ""); p3(nrm); printf(", \"dir\": "); p3(dir);
            printf(", \"cosine_pdf\": "); pf(cp.value(dir)); printf(", \"scattering_pdf\": "); pf(sp); printf("}");
        }
        printf("\n");
    }
    printf("],\n");

    // ---------------- light list + mixture (two-triangle emitter like cornell_duck's) ----------------
    key("mixture"); printf("{");
    {
        UniversalMaterial lm(make_float3(0, 0, 0), nullptr, make_float3(1, 1, 1), nullptr);
        Vertex q0{make_float3(109.791084f, 338.58173f, -861.65204f), make_float2(0, 0)};
        Vertex q1{make_float3(-150.20908f, 338.5817f, -861.65204f), make_float2(0, 0)};
        Vertex q2{make_float3(-150.20908f, 338.58167f, -1071.6521f), make_float2(0, 0)};
        Vertex q3{make_float3(109.79106f, 338.5817f, -1071.6521f), make_float2(0, 0)};
        std::vector<triangle> tris;
        tris.reserve(2);
        tris.emplace_back(q0, q1, q2, &lm);
        tris.emplace_back(q0, q2, q3, &lm);
        triangle *ptrs[2] = {&tris[0], &tris[1]};
        hitable_list lights(ptrs, 2);
        printf("\"lights\": [[%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g], [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g]], \"cases\": [\n",
               q0.position.x, q0.position.y, q0.position.z, q1.position.x, q1.position.y, q1.position.z, q2.position.x, q2.position.y, q2.position.z,
               q0.position.x, q0.position.y, q0.position.z, q2.position.x, q2.position.y, q2.position.z, q3.position.x, q3.position.y, q3.position.z);
        for (int i = 0; i < 32; i++) {
            float3 p = make_float3(frange(-250, 250), frange(-200, 300), frange(-1200, -700));
            float3 nrm = normalize(make_float3(frange(-1, 1), frange(-1, 1), frange(-1, 1)));
            unsigned long long seed = 9000 + (unsigned long long)i;
            curandState s;
            curand_init(seed, 0, 0, &s);
            hitable_list_pdf lp(&lights, p);
            cosine_pdf sp(nrm);
            mixture_pdf mp(&lp, &sp);
            float3 dir = mp.generate(&s);
            float val = mp.value(dir);
            float lval = lp.value(dir);
            float next_u = curand_uniform(&s);  // pins the number of draws consumed
            printf("  {\"p\": "); p3(p); printf(", \"normal\": "); p3(nrm); printf(", \"seed\": %llu, \"dir\": ", seed); p3(dir);
            printf(", \"value\": "); pf(val); printf(", \"light_value\": "); pf(lval); printf(", \"next_uniform\": "); pf(next_u);
            printf("}%s\n", i < 31 ? "," : "");
        }
        printf("]");
    }
    printf("},\n");

    // ---------------- RTOW materials (dead code in the reference) ----------------
    key("rtow"); printf("[\n");
    {
        int n = 0;
        for (int i = 0; i < 30; i++) {
            float3 nrm = normalize(make_float3(frange(-1, 1), frange(-1, 1), frange(-1, 1)));
            float3 din = make_float3(frange(-2, 2), frange(-2, 2), frange(-2, 2));
            float3 p = make_float3(frange(-1, 1), frange(-1, 1), frange(-1, 1));
            hit_record rec;
            rec.p = p; rec.normal = nrm; rec.t = 1.f;
            unsigned long long seed = 300 + (unsigned long long)i;
            int kind = i % 3;
            float fuzz = frange(0.f, 1.3f), ior = frange(1.1f, 2.4f);
            float3 albedo = make_float3(frand(), frand(), frand());
            curandState s;
            curand_init(seed, 0, 0, &s);
            float3 att = make_float3(0, 0, 0);
            ray sc(make_float3(0, 0, 0), make_float3(0, 0, 0));
            bool ok;
            if (kind == 0) { lambertian m(albedo); ok = m.scatter(ray(p - din, din), rec, att, sc, &s); }
            else if (kind == 1) { metal m(albedo, fuzz); ok = m.scatter(ray(p - din, din), rec, att, sc, &s); }
            else { dielectric m(ior); ok = m.scatter(ray(p - din, din), rec, att, sc, &s); }
            float next_u = curand_uniform(&s);
            float3 refr = make_float3(0, 0, 0);
            bool rok = refract(din, nrm, 1.0f / ior, refr);
            printf("%s  {\"kind\": %d, \"normal\": ", n++ ? ",\n" : "", kind); p3(nrm); printf(", \"d_in\": "); p3(din); printf(", \"p\": "); p3(p);
            printf(", \"albedo\": "); p3(albedo); printf(", \"fuzz\": %.9g, \"ior\": %.9g, \"seed\": %llu, \"scattered\": %d, \"attenuation\": ", fuzz, ior, seed, ok ? 1 : 0);
            p3(att); printf(", \"dir\": "); p3(sc.direction()); printf(", \"next_uniform\": "); pf(next_u);
            printf(", \"refract_ok\": %d, \"refracted\": ", rok ? 1 : 0); p3(rok ? refr : make_float3(0, 0, 0));
            printf(", \"schlick\": "); pf(schlick_approx(fabsf(dot(normalize(din), nrm)), ior)); printf("}");
        }
        printf("\n");
    }
    printf("],\n");

    // ---------------- quantiser + I420 (expressions of src/DevicePathTracer.h:98-119) ----------------
    key("quantise"); printf("[\n");
    {
        float cols[][3] = {{0, 0, 0}, {1, 1, 1}, {0.5f, 0.25f, 0.125f}, {63.9f, 64.f, 64.1f}, {1e9f, -5.f, 3.3f}, {12.34f, 56.78f, 0.001f}, {640.f, 639.9f, 0.0039f}};
        int spps[] = {1, 64, 1, 64, 8, 64, 640};
        for (size_t i = 0; i < sizeof spps / sizeof *spps; i++) {
            float3 col = make_float3(cols[i][0], cols[i][1], cols[i][2]);
            int sample_per_pixel = spps[i];
            float3 color_modifier = make_float3(1, 1, 1);
            int3 color = make_int3(255.99 * col / float(sample_per_pixel) * color_modifier);
            color.x = min(255, color.x); color.y = min(255, color.y); color.z = min(255, color.z);
            uint8_t r = (uint8_t)color.x, g = (uint8_t)color.y, b = (uint8_t)color.z;
            uint8_t Y = (uint8_t)(((66 * color.x + 129 * color.y + 25 * color.z + 128) >> 8) + 16);
            uint8_t Uc = (uint8_t)(((-38 * color.x - 74 * color.y + 112 * color.z + 128) >> 8) + 128);
            uint8_t Vc = (uint8_t)(((112 * color.x - 94 * color.y - 18 * color.z + 128) >> 8) + 128);
            printf("  {\"col\": [%.9g, %.9g, %.9g], \"spp\": %d, \"rgb\": [%u, %u, %u], \"yuv\": [%u, %u, %u]}%s\n", col.x, col.y, col.z, sample_per_pixel, r, g, b, Y, Uc, Vc,
                   i + 1 < sizeof spps / sizeof *spps ? "," : "");
        }
    }
    printf("]\n}\n");
    return 0;
}
