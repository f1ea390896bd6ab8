/* pt_oracle.c — CPU restatement (plain C) of the reference's path-tracing hot path.
 *
 * TEST INFRASTRUCTURE: see pt_oracle.h.  Compile with -ffp-contract=off (oracle/Makefile) so
 * every float expression below is evaluated exactly as written, which is also how g++ -O2
 * evaluates the reference's headers on x86-64; that is what makes this file bit-identical
 * to oracle/_ref/ref_cpu (checked by tests/test_oracle_parity.py).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * Expressions keep the reference's operand order AND its literal types: an `int` literal
 * next to a float stays float, a `double` literal (M_PI, 1.0, 0.0001, 0.999999, 1e-8, 2.0)
 * promotes the operation to double exactly where the reference's does.
 *
 * Third-party arithmetic on the path: cuRAND XORWOW (CUDA 12.9 curand_kernel.h:772-797,
 * 863-874; curand_uniform.h:69-72), restated in pto_rng_*.
 *
 * Builder-defined pieces (no reference renderer exists for them):
 *   - the integrator glue for LAMBERTIAN / METAL / DIELECTRIC / DIFFUSE_LIGHT hits,
 *   - random_in_unit_sphere draws x, y, z in that order,
 *   - schlick's pow(1-c,5) is evaluated as x2=x*x, x4=x2*x2, x5=x4*x,
 *   - an empty light list (the reference dereferences list[0]: SURVEY §0.3) degrades to
 *     the cosine branch with light pdf 0, consuming the same number of draws,
 *   - texture row `height` (one past the end, Texture.h:68) reads as row height-1.
 */
#include "pt_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef pto_vec3 v3;

/* ---- helper_math.h vector ops, same expression shapes ---- */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float b) { return V(a.x * b, a.y * b, a.z * b); }   /* operator*(float3,float) and (float,float3) */
static inline v3 vdiv(v3 a, float b) { return V(a.x / b, a.y / b, a.z / b); }     /* helper_math.h:1015-1018 */
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* :1266-1269 */
static inline v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } /* :1461-1464 */
static inline float vlength(v3 a) { return sqrtf(vdot(a, a)); }                   /* :1309-1312 */
static inline float host_rsqrtf(float x) { return 1.0f / sqrtf(x); }              /* host fallback, :80-83 */
static inline v3 vnormalize(v3 a) { float inv = host_rsqrtf(vdot(a, a)); return vscale(a, inv); } /* :1327-1331 */

/* ---- cuRAND XORWOW ---- */
void pto_rng_init(pto_rng *s, uint64_t seed) {
    /* curand_kernel.h:772-797 with subsequence = offset = 0 (both skip-aheads are identities) */
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49U;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddU;
    uint32_t t0 = 1099087573U * s0;
    uint32_t t1 = 2591861531U * s1;
    s->d = 6615241U + t1 + t0;
    s->v[0] = 123456789U + t0;
    s->v[1] = 362436069U ^ t0;
    s->v[2] = 521288629U + t1;
    s->v[3] = 88675123U ^ t1;
    s->v[4] = 5783321U + t0;
}

uint32_t pto_rng_next(pto_rng *s) {
    /* curand_kernel.h:863-874 */
    uint32_t t = s->v[0] ^ (s->v[0] >> 2);
    s->v[0] = s->v[1];
    s->v[1] = s->v[2];
    s->v[2] = s->v[3];
    s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437U;
    return s->v[4] + s->d;
}

float pto_uniform(pto_rng *s) {
    /* curand_uniform.h:69-72: x * 2^-32 + 2^-33, in (0, 1] */
    uint32_t x = pto_rng_next(s);
    return (float)x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

typedef struct { pto_rng *s; pto_stats *st; float *events; int max_events, n_events; float cur_sample; } rng_ctx;
static inline float U(rng_ctx *c) {
    if (c->st) c->st->draws++;
    return pto_uniform(c->s);
}

/* ---- samplers ---- */
static v3 random_cosine_direction(rng_ctx *c) {
    /* helper_math.h:1519-1528 — the factor 2 on x and y is the reference's, kept on purpose */
    float r1 = U(c);
    float r2 = U(c);
    float z = sqrtf(1 - r2);
    float phi = (float)(2 * M_PI * r1);
    float x = cosf(phi) * 2 * sqrtf(r2);
    float y = sinf(phi) * 2 * sqrtf(r2);
    return V(x, y, z);
}
pto_vec3 pto_random_cosine_direction(pto_rng *s) { rng_ctx c = {s, NULL, NULL, 0, 0, 0.f}; return random_cosine_direction(&c); }

/* make_random_float3 draws its three components inside ONE argument list (helper_math.h:1505-1507): the order is unspecified in C++.  nvcc's
 * device code evaluates left to right (x, y, z: what the CUDA core and this restatement do), g++ on x86-64 right to left.  The switch exists
 * so that the HOST build of the reference's dead classes (oracle/_ref/ref_cpu_spheres) can pin everything else of rows D1-D6 bit for bit. */
static int g_triple_zyx = 0;
void pto_set_triple_draw_order_zyx(int on) { g_triple_zyx = on; }
static v3 random_in_unit_sphere(rng_ctx *c) {
    /* helper_math.h:1504-1518; draw order x,y,z (builder-defined, see header comment) */
    v3 p;
    do {
        float a, b, d;
        if (g_triple_zyx) { d = U(c); b = U(c); a = U(c); }
        else { a = U(c); b = U(c); d = U(c); }
        p = vsub(vscale(V(a, b, d), 2.0f), V(1.0f, 1.0f, 1.0f));
    } while (vdot(p, p) >= 1.0f);
    return p;
}
pto_vec3 pto_random_in_unit_sphere(pto_rng *s) { rng_ctx c = {s, NULL, NULL, 0, 0, 0.f}; return random_in_unit_sphere(&c); }

void pto_onb(pto_vec3 n, pto_vec3 axis[3]) {
    /* onb.h:8-13 */
    axis[2] = vnormalize(n);
    v3 a = fabsf(axis[2].x) > 0.9f ? V(0.0f, 1.0f, 0.0f) : V(1.0f, 0.0f, 0.0f);
    axis[1] = vnormalize(vcross(axis[2], a));
    axis[0] = vcross(axis[2], axis[1]);
}
static inline v3 onb_local(const v3 axis[3], v3 v) {
    /* onb.h:19-21 */
    return vadd(vadd(vscale(axis[0], v.x), vscale(axis[1], v.y)), vscale(axis[2], v.z));
}

/* ---- camera ---- */
typedef struct { v3 origin, lower_left_corner, horizontal, vertical; } cam_params;
static cam_params camera_params(const PtCamera *cc) {
    /* camera.h:21-36 */
    cam_params c;
    v3 lookFrom = V(cc->look_from[0], cc->look_from[1], cc->look_from[2]);
    v3 front = V(cc->front[0], cc->front[1], cc->front[2]);
    v3 vup = V(0.0f, 1.0f, 0.0f);
    v3 lookAt = vadd(lookFrom, front);
    float theta_v = (float)(cc->vfov * M_PI / 180);
    float half_height = tanf(theta_v / 2);
    float theta_h = (float)(cc->hfov * M_PI / 180);
    float half_width = tanf(theta_h / 2);
    c.origin = lookFrom;
    v3 w = vnormalize(vsub(lookFrom, lookAt));
    v3 u = vnormalize(vcross(vup, w));
    v3 v = vcross(w, u);
    c.lower_left_corner = vsub(vsub(vsub(c.origin, vscale(u, half_width)), vscale(v, half_height)), w);
    c.horizontal = vscale(u, 2 * half_width);
    c.vertical = vscale(v, 2 * half_height);
    return c;
}
static inline void camera_get_ray(const cam_params *c, float u, float v, v3 *o, v3 *d) {
    /* camera.h:95-97 */
    *o = c->origin;
    *d = vsub(vadd(vadd(c->lower_left_corner, vscale(c->horizontal, u)), vscale(c->vertical, v)), c->origin);
}
void pto_camera_ray(const PtCamera *cam, float u, float v, pto_vec3 *o, pto_vec3 *d) {
    cam_params c = camera_params(cam);
    camera_get_ray(&c, u, v, o, d);
}

/* ---- triangle ---- */
int pto_triangle_hit(const float pos[9], const float uv[6], pto_vec3 o, pto_vec3 d, float tmin, float tmax, pto_hit *rec) {
    /* triangle.h:63-113 (Moeller-Trumbore; inv_det is a double divide rounded to float) */
    v3 v0 = V(pos[0], pos[1], pos[2]), v1 = V(pos[3], pos[4], pos[5]), v2 = V(pos[6], pos[7], pos[8]);
    v3 e1 = vsub(v1, v0);
    v3 e2 = vsub(v2, v0);
    v3 pvec = vcross(d, e2);
    float det = vdot(e1, pvec);
    if (det < 1e-8 && det > -1e-8) return 0;
    float inv_det = (float)(1.0 / det);
    v3 tvec = vsub(o, v0);
    float u = vdot(tvec, pvec) * inv_det;
    if (u < 0 || u > 1) return 0;
    v3 qvec = vcross(tvec, e1);
    float v = vdot(d, qvec) * inv_det;
    if (v < 0 || u + v > 1) return 0;
    float t = vdot(e2, qvec) * inv_det;
    if (t < tmax && t > tmin) {
        rec->hit = 1;
        rec->t = t;
        rec->bu = u;
        rec->bv = v;
        rec->p = vadd(o, vscale(d, t));               /* ray.h:19: A + t*B */
        rec->normal = vnormalize(vcross(e1, e2));     /* geometric, never flipped towards the ray */
        if (uv) {
            rec->u = (1 - u - v) * uv[0] + u * uv[2] + v * uv[4];
            rec->v = (1 - u - v) * uv[1] + u * uv[3] + v * uv[5];
        } else {
            rec->u = rec->v = 0.f;
        }
        return 1;
    }
    return 0;
}

float pto_triangle_area(const float pos[9]) {
    /* triangle.h:28 (host code in the reference) */
    v3 v0 = V(pos[0], pos[1], pos[2]), v1 = V(pos[3], pos[4], pos[5]), v2 = V(pos[6], pos[7], pos[8]);
    return vlength(vcross(vsub(v1, v0), vsub(v2, v0))) * 0.5f;
}

float pto_triangle_pdf_value(const float pos[9], pto_vec3 o, pto_vec3 v) {
    /* triangle.h:32-40 */
    pto_hit rec;
    if (!pto_triangle_hit(pos, NULL, o, v, 0.001f, FLT_MAX, &rec)) return 0;
    float distance_squared = rec.t * rec.t * vdot(v, v);
    float cosine = fabsf(vdot(v, rec.normal) / vlength(v));
    return distance_squared / (cosine * pto_triangle_area(pos));
}

static v3 triangle_random(const float pos[9], v3 o, rng_ctx *c) {
    /* triangle.h:41-47 */
    v3 v0 = V(pos[0], pos[1], pos[2]), v1 = V(pos[3], pos[4], pos[5]), v2 = V(pos[6], pos[7], pos[8]);
    float r1 = U(c);
    float r2 = U(c);
    float sqrt_r1 = sqrtf(r1);
    v3 random_point = vadd(vadd(vscale(v0, 1 - sqrt_r1), vscale(v1, sqrt_r1 * (1 - r2))), vscale(v2, sqrt_r1 * r2));
    return vsub(random_point, o);
}
pto_vec3 pto_triangle_random(const float pos[9], pto_vec3 o, pto_rng *s) { rng_ctx c = {s, NULL, NULL, 0, 0, 0.f}; return triangle_random(pos, o, &c); }

/* ---- sphere (dead code in the reference, live here: SURVEY §8a D1) ---- */
int pto_sphere_hit(const float sph[4], pto_vec3 o, pto_vec3 d, float tmin, float tmax, pto_hit *rec) {
    /* sphere.h:21-50 */
    v3 center = V(sph[0], sph[1], sph[2]);
    float radius = sph[3];
    v3 oc = vsub(o, center);
    float a = vdot(d, d);
    float b = (float)(2.0 * vdot(oc, d));
    float c = vdot(oc, oc) - radius * radius;
    float discriminant = b * b - 4 * a * c;
    if (discriminant > 0) {
        float temp = (float)((-b - sqrtf(discriminant)) / (2.0 * a));
        if (temp < tmax && temp > tmin) {
            rec->hit = 1; rec->t = temp; rec->bu = rec->bv = 0.f;
            rec->p = vadd(o, vscale(d, temp));
            rec->normal = vdiv(vsub(rec->p, center), radius);
            rec->u = rec->v = 0.f;
            return 1;
        }
        temp = (float)((-b + sqrtf(discriminant)) / (2.0 * a));
        if (temp < tmax && temp > tmin) {
            rec->hit = 1; rec->t = temp; rec->bu = rec->bv = 0.f;
            rec->p = vadd(o, vscale(d, temp));
            rec->normal = vdiv(vsub(rec->p, center), radius);
            rec->u = rec->v = 0.f;
            return 1;
        }
    }
    return 0;
}

/* ---- texture ---- */
static inline int tex_clamp(int x, int low, int high) {
    /* Texture.h:49-53 */
    if (x < low) return low;
    if (x < high) return x;
    return high - 1;
}
pto_vec3 pto_texture_value(const PtTexture *t, float u, float v) {
    /* Texture.h:30-47,61-70 */
    if (t->height <= 0) return V((float)242 / 255, (float)45 / 255, (float)27 / 255);
    u = fmodf(u, 1.0f);
    v = fmodf(v, 1.0f);
    int i = (int)(u * t->width);
    int j = (int)(v * t->height);
    v3 pixel;
    if (t->rgb == NULL) {
        pixel = V(52, 27, 242);
    } else {
        int x = tex_clamp(i, 0, t->width);
        int y = tex_clamp(j, 0, t->height);
        y = t->height - y;
        if (y >= t->height) y = t->height - 1; /* reference reads one row past the end here (UB); defined as the last row */
        const float *p = t->rgb + ((size_t)y * (size_t)t->width + (size_t)x) * 3;
        pixel = V(p[0], p[1], p[2]);
    }
    double color_scale = 1.0 / 255.0;
    return V((float)(color_scale * pixel.x), (float)(color_scale * pixel.y), (float)(color_scale * pixel.z));
}

/* ---- pdfs ---- */
float pto_cosine_pdf_value(pto_vec3 normal, pto_vec3 dir) {
    /* pdf.h:16-22: uvw = onb(normal); cosine = dot(normalize(direction), uvw.w()) */
    v3 w = vnormalize(normal);
    float cosine = vdot(vnormalize(dir), w);
    return (cosine <= 0) ? 0 : (float)(cosine / M_PI);
}
float pto_scattering_pdf(pto_vec3 normal, pto_vec3 dir) {
    /* material.h:88-91 */
    float cosine = vdot(normal, vnormalize(dir));
    return cosine < 0 ? 0 : (float)(cosine / M_PI);
}

/* ---- RTOW materials (dead code in the reference: SURVEY §8a D2-D5) ---- */
static inline v3 reflect3(v3 i, v3 n) {
    /* helper_math.h:1429-1432: i - 2.0f * n * dot(n,i) */
    return vsub(i, vscale(vscale(n, 2.0f), vdot(n, i)));
}
int pto_refract(pto_vec3 v, pto_vec3 n, float ni_over_nt, pto_vec3 *out) {
    /* helper_math.cu:7-17 */
    v3 uv = vnormalize(v);
    float dt = vdot(uv, n);
    float discriminant = 1.0f - ni_over_nt * ni_over_nt * (1 - dt * dt);
    if (discriminant > 0) {
        *out = vsub(vscale(vsub(uv, vscale(n, dt)), ni_over_nt), vscale(n, sqrtf(discriminant)));
        return 1;
    }
    return 0;
}
float pto_schlick(float cosine, float ref_idx) {
    /* material.h:10-14; pow(x,5) by repeated multiplication (builder-defined) */
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    float x = 1 - cosine;
    float x2 = x * x, x4 = x2 * x2, x5 = x4 * x;
    return r0 + (1 - r0) * x5;
}
static inline int near_zero(v3 a) {
    /* helper_math.h:1531-1535 */
    const double s = 1e-8;
    return (fabsf(a.x) < s) && (fabsf(a.y) < s) && (fabsf(a.z) < s);
}

/* ---- quantiser / colour conversion ---- */
static inline int f2i_trunc(float f) {
    /* CUDA's float->int conversion: NaN -> 0, saturating (cvt.rzi.s32.f32) */
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
void pto_quantise(const float col[3], uint32_t spp, uint8_t rgb[3]) {
    /* DevicePathTracer.h:98-105: make_int3(255.99 * col / float(spp) * (1,1,1)); min(255,.); byte store */
    for (int k = 0; k < 3; k++) {
        float c = (255.99f * col[k]) / (float)spp * 1.0f;
        int q = f2i_trunc(c);
        if (q > 255) q = 255;
        rgb[k] = (uint8_t)q;
    }
}
void pto_yuv(const uint8_t rgb[3], uint8_t *y, uint8_t *u, uint8_t *v) {
    /* DevicePathTracer.h:107-119 */
    int r = rgb[0], g = rgb[1], b = rgb[2];
    *y = (uint8_t)(((66 * r + 129 * g + 25 * b + 128) >> 8) + 16);
    *u = (uint8_t)(((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128);
    *v = (uint8_t)(((112 * r - 94 * g - 18 * b + 128) >> 8) + 128);
}

/* =====================================================================================
 * world: primitives + a plain median-split BVH.  bvh.h:178-246 returns the closest hit over
 * ALL triangles (its boxes ignore ray_t); any exact closest-hit search is equivalent up to
 * the order of exactly-equal t, so the oracle keeps its own simple tree.
 * ===================================================================================== */
typedef struct { float lo[3], hi[3]; int32_t left, right, first, count; } onode;

struct pto_world {
    PtSceneDesc sc;
    float *tri_pos, *tri_uv, *sph;
    int32_t *tri_mat, *sph_mat;
    PtMaterial *mats;
    PtTexture *tex;
    int32_t n_prims;
    int32_t *prim;      /* permutation of primitive ids, leaves index into it */
    onode *nodes;
    int32_t n_nodes;
    int32_t *lights;    /* triangle ids, in scene order */
    float *light_area;
    int32_t n_lights;
};

static void prim_bounds(const pto_world *w, int32_t id, float lo[3], float hi[3]) {
    if (id < w->sc.n_tris) {
        const float *p = w->tri_pos + (size_t)id * 9;
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(p[k], fminf(p[3 + k], p[6 + k]));
            hi[k] = fmaxf(p[k], fmaxf(p[3 + k], p[6 + k]));
        }
    } else {
        const float *s = w->sph + (size_t)(id - w->sc.n_tris) * 4;
        float r = fabsf(s[3]);
        for (int k = 0; k < 3; k++) { lo[k] = s[k] - r; hi[k] = s[k] + r; }
    }
    for (int k = 0; k < 3; k++) {  /* conservative padding: the slab test below must never cull a hit the primitive test accepts */
        float m = fmaxf(fabsf(lo[k]), fabsf(hi[k]));
        float e = m * 1e-5f + 1e-6f;
        lo[k] -= e; hi[k] += e;
    }
}

static int build_axis;
static const pto_world *build_world;
static int cmp_centroid(const void *a, const void *b) {
    float la[3], ha[3], lb[3], hb[3];
    prim_bounds(build_world, *(const int32_t *)a, la, ha);
    prim_bounds(build_world, *(const int32_t *)b, lb, hb);
    float ca = la[build_axis] + ha[build_axis], cb = lb[build_axis] + hb[build_axis];
    return (ca > cb) - (ca < cb);
}

static int32_t build_node(pto_world *w, int32_t first, int32_t count) {
    int32_t idx = w->n_nodes++;
    onode *n = &w->nodes[idx];
    for (int k = 0; k < 3; k++) { n->lo[k] = FLT_MAX; n->hi[k] = -FLT_MAX; }
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int32_t i = 0; i < count; i++) {
        float lo[3], hi[3];
        prim_bounds(w, w->prim[first + i], lo, hi);
        for (int k = 0; k < 3; k++) {
            n->lo[k] = fminf(n->lo[k], lo[k]); n->hi[k] = fmaxf(n->hi[k], hi[k]);
            float c = lo[k] + hi[k];
            clo[k] = fminf(clo[k], c); chi[k] = fmaxf(chi[k], c);
        }
    }
    n->first = first; n->count = count; n->left = n->right = -1;
    if (count <= 4) return idx;
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    if (!(chi[axis] > clo[axis])) return idx;
    build_axis = axis; build_world = w;
    qsort(w->prim + first, (size_t)count, sizeof(int32_t), cmp_centroid);
    int32_t half = count / 2;
    int32_t l = build_node(w, first, half);
    int32_t r = build_node(w, first + half, count - half);
    n = &w->nodes[idx];
    n->left = l; n->right = r; n->count = 0;
    return idx;
}

pto_world *pto_world_create(const PtSceneDesc *sc) {
    pto_world *w = (pto_world *)calloc(1, sizeof *w);
    w->sc = *sc;
    size_t nt = (size_t)sc->n_tris, ns = (size_t)sc->n_spheres;
    w->tri_pos = (float *)malloc(sizeof(float) * 9 * (nt ? nt : 1));
    w->tri_uv = (float *)calloc(6 * (nt ? nt : 1), sizeof(float));
    w->tri_mat = (int32_t *)calloc(nt ? nt : 1, sizeof(int32_t));
    w->sph = (float *)malloc(sizeof(float) * 4 * (ns ? ns : 1));
    w->sph_mat = (int32_t *)calloc(ns ? ns : 1, sizeof(int32_t));
    if (nt) { memcpy(w->tri_pos, sc->tri_pos, sizeof(float) * 9 * nt); memcpy(w->tri_mat, sc->tri_mat, sizeof(int32_t) * nt); }
    if (nt && sc->tri_uv) memcpy(w->tri_uv, sc->tri_uv, sizeof(float) * 6 * nt);
    if (ns) { memcpy(w->sph, sc->sph, sizeof(float) * 4 * ns); memcpy(w->sph_mat, sc->sph_mat, sizeof(int32_t) * ns); }
    w->mats = (PtMaterial *)malloc(sizeof(PtMaterial) * (size_t)(sc->n_mats ? sc->n_mats : 1));
    if (sc->n_mats) memcpy(w->mats, sc->mats, sizeof(PtMaterial) * (size_t)sc->n_mats);
    {   /* DevicePathTracer.h:269-279 (loadMaterials): the texture pointers are declared OUTSIDE the loop over the materials, so a
         * UniversalMaterial without a texture of its own inherits the last one seen ("sticky" pointers). */
        int sticky_base = -1, sticky_emis = -1;
        for (int i = 0; i < sc->n_mats; i++) {
            if (w->mats[i].type != PT_MAT_UNIVERSAL) continue;
            if (w->mats[i].base_tex >= 0) sticky_base = w->mats[i].base_tex;
            if (w->mats[i].emis_tex >= 0) sticky_emis = w->mats[i].emis_tex;
            w->mats[i].base_tex = sticky_base;
            w->mats[i].emis_tex = sticky_emis;
        }
    }
    w->tex = (PtTexture *)calloc((size_t)(sc->n_tex ? sc->n_tex : 1), sizeof(PtTexture));
    for (int i = 0; i < sc->n_tex; i++) {
        w->tex[i] = sc->tex[i];
        size_t n = (size_t)sc->tex[i].width * (size_t)sc->tex[i].height * 3;
        float *cp = (float *)malloc(sizeof(float) * (n ? n : 1));
        if (n && sc->tex[i].rgb) memcpy(cp, sc->tex[i].rgb, sizeof(float) * n);
        w->tex[i].rgb = sc->tex[i].rgb ? cp : NULL;
        if (!sc->tex[i].rgb) free(cp);
    }
    w->n_prims = sc->n_tris + sc->n_spheres;
    w->prim = (int32_t *)malloc(sizeof(int32_t) * (size_t)(w->n_prims ? w->n_prims : 1));
    for (int32_t i = 0; i < w->n_prims; i++) w->prim[i] = i;
    w->nodes = (onode *)malloc(sizeof(onode) * (size_t)(2 * (w->n_prims ? w->n_prims : 1)));
    w->n_nodes = 0;
    if (w->n_prims) build_node(w, 0, w->n_prims);
    /* light list: DevicePathTracer.h:302-307 — triangles whose material's emissiveFactor has a channel > 0.0001, in order */
    w->lights = (int32_t *)malloc(sizeof(int32_t) * (nt ? nt : 1));
    w->light_area = (float *)malloc(sizeof(float) * (nt ? nt : 1));
    for (int32_t i = 0; i < sc->n_tris; i++) {
        const PtMaterial *m = &w->mats[w->tri_mat[i]];
        if (m->type == PT_MAT_UNIVERSAL && (m->emis[0] > 0.0001 || m->emis[1] > 0.0001 || m->emis[2] > 0.0001)) {
            w->lights[w->n_lights] = i;
            w->light_area[w->n_lights] = pto_triangle_area(w->tri_pos + (size_t)i * 9);
            w->n_lights++;
        }
    }
    return w;
}

void pto_world_destroy(pto_world *w) {
    if (!w) return;
    for (int i = 0; i < w->sc.n_tex; i++) free((void *)w->tex[i].rgb);
    free(w->tri_pos); free(w->tri_uv); free(w->tri_mat); free(w->sph); free(w->sph_mat);
    free(w->mats); free(w->tex); free(w->prim); free(w->nodes); free(w->lights); free(w->light_area);
    free(w);
}
int pto_world_n_lights(const pto_world *w) { return w->n_lights; }

static inline int slab(const onode *n, v3 o, v3 inv, float tmax) {
    float t1 = (n->lo[0] - o.x) * inv.x, t2 = (n->hi[0] - o.x) * inv.x;
    float t3 = (n->lo[1] - o.y) * inv.y, t4 = (n->hi[1] - o.y) * inv.y;
    float t5 = (n->lo[2] - o.z) * inv.z, t6 = (n->hi[2] - o.z) * inv.z;
    float tn = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    float tf = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    if (tn != tn || tf != tf) return 1; /* NaN (0*inf): be conservative */
    return tf >= 0.f && tn <= tf * 1.00001f && tn <= tmax;
}

int pto_world_hit(const pto_world *w, pto_vec3 o, pto_vec3 d, float tmin, float tmax, pto_hit *rec) {
    rec->hit = 0;
    if (!w->n_prims) return 0;
    v3 inv = V(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    float closest = tmax;
    while (sp) {
        const onode *n = &w->nodes[stack[--sp]];
        if (!slab(n, o, inv, closest)) continue;
        if (n->left < 0) {
            for (int32_t i = 0; i < n->count; i++) {
                int32_t id = w->prim[n->first + i];
                pto_hit tmp;
                int h;
                if (id < w->sc.n_tris) {
                    h = pto_triangle_hit(w->tri_pos + (size_t)id * 9, w->tri_uv + (size_t)id * 6, o, d, tmin, closest, &tmp);
                    if (h) tmp.mat = w->tri_mat[id];
                } else {
                    h = pto_sphere_hit(w->sph + (size_t)(id - w->sc.n_tris) * 4, o, d, tmin, closest, &tmp);
                    if (h) tmp.mat = w->sph_mat[id - w->sc.n_tris];
                }
                if (h) { tmp.prim = id; *rec = tmp; closest = tmp.t; }
            }
        } else {
            if (sp + 2 > 128) return rec->hit; /* cannot happen for a median-split tree */
            stack[sp++] = n->right;
            stack[sp++] = n->left;
        }
    }
    return rec->hit;
}

/* ---- light list (hitable_list.h:16-26) ---- */
static float lights_pdf_value(const pto_world *w, v3 o, v3 v) {
    if (w->n_lights == 0) return 0.f;
    float weight = 1.0f / w->n_lights;
    float sum = 0.0f;
    for (int i = 0; i < w->n_lights; i++) sum += weight * pto_triangle_pdf_value(w->tri_pos + (size_t)w->lights[i] * 9, o, v);
    return sum;
}
static v3 lights_random(const pto_world *w, v3 o, rng_ctx *c) {
    int index = (int)truncf((float)(U(c) * ((w->n_lights - 1) + 0.999999)));
    return triangle_random(w->tri_pos + (size_t)w->lights[index] * 9, o, c);
}

/* ---- the integrator ---- */
static v3 ray_color(const pto_world *w, v3 ro, v3 rd, uint32_t depth, rng_ctx *c) {
    /* camera.h:49-83 */
    pto_stats *st = c->st;
    v3 cur_o = ro, cur_d = rd;
    v3 cur_attenuation = V(1.0f, 1.0f, 1.0f);
    for (uint32_t i = 0; i < depth; i++) {
        pto_hit rec;
        if (st) st->rays++;
        int got = pto_world_hit(w, cur_o, cur_d, 0.001f, FLT_MAX, &rec);
        if (c->events) {
            if (c->n_events < c->max_events) {
                float *e = c->events + (size_t)c->n_events * 16;
                e[0] = c->cur_sample; e[1] = (float)i; e[2] = got ? (float)rec.prim : -1.f; e[3] = got ? rec.t : FLT_MAX;
                e[4] = got ? rec.bu : 0.f; e[5] = got ? rec.bv : 0.f;
                e[6] = cur_o.x; e[7] = cur_o.y; e[8] = cur_o.z; e[9] = cur_d.x; e[10] = cur_d.y; e[11] = cur_d.z;
                e[12] = cur_attenuation.x; e[13] = cur_attenuation.y; e[14] = cur_attenuation.z; e[15] = 0.f;
            }
            c->n_events++;
        }
        if (got) {
            const PtMaterial *m = &w->mats[rec.mat];
            if (m->type == PT_MAT_UNIVERSAL) {
                /* UniversalMaterial::scatter, material.h:52-78 */
                v3 emis = V(m->emis[0], m->emis[1], m->emis[2]);
                v3 emitted;
                if (m->emis_tex >= 0) emitted = vscale(vmul(pto_texture_value(&w->tex[m->emis_tex], rec.u, rec.v), emis), 50);
                else emitted = vscale(emis, 50);
                if (emitted.x > 0.0001 || emitted.y > 0.0001 || emitted.z > 0.0001) {
                    cur_attenuation = vmul(cur_attenuation, emitted);  /* camera.h:72-75 */
                    if (st) st->emitter_paths++;
                    return cur_attenuation;
                }
                (void)random_cosine_direction(c);  /* material.h:67-69: two draws whose result camera.h:65 overwrites */
                v3 attenuation = V(m->base[0], m->base[1], m->base[2]);
                if (m->base_tex >= 0) attenuation = vmul(attenuation, pto_texture_value(&w->tex[m->base_tex], rec.u, rec.v));

                /* camera.h:62-69: mixture of light sampling and cosine sampling */
                v3 axis[3];
                pto_onb(rec.normal, axis);
                v3 dir;
                if (w->n_lights > 0) {
                    if (U(c) < 0.5f) dir = lights_random(w, rec.p, c);                 /* pdf.h:66-68 */
                    else dir = onb_local(axis, random_cosine_direction(c));            /* pdf.h:69-71 */
                } else {
                    (void)U(c);
                    dir = onb_local(axis, random_cosine_direction(c));
                }
                float pdf_value = 0.5f * lights_pdf_value(w, rec.p, dir) + 0.5f * pto_cosine_pdf_value(rec.normal, dir); /* pdf.h:63-65 */
                float scattering_pdf = pto_scattering_pdf(rec.normal, dir);
                cur_attenuation = vmul(cur_attenuation, vdiv(vscale(attenuation, scattering_pdf), pdf_value));
                cur_o = rec.p;
                cur_d = dir;
            } else if (m->type == PT_MAT_DIFFUSE_LIGHT) {
                /* material.h:210-217 + builder-defined glue */
                if (st) st->emitter_paths++;
                return vmul(cur_attenuation, V(m->emis[0], m->emis[1], m->emis[2]));
            } else {
                v3 attenuation, sdir;
                int scattered = 1;
                if (m->type == PT_MAT_LAMBERTIAN) {
                    /* material.h:113-124 */
                    sdir = vadd(rec.normal, random_in_unit_sphere(c));
                    if (near_zero(sdir)) sdir = rec.normal;
                    attenuation = V(m->base[0], m->base[1], m->base[2]);
                } else if (m->type == PT_MAT_METAL) {
                    /* material.h:133-140 */
                    float fuzz = m->fuzz < 1 ? m->fuzz : 1;
                    v3 reflected = reflect3(vnormalize(cur_d), rec.normal);
                    sdir = vadd(reflected, vscale(random_in_unit_sphere(c), fuzz));
                    attenuation = V(m->base[0], m->base[1], m->base[2]);
                    scattered = vdot(sdir, rec.normal) > 0;
                } else {
                    /* dielectric, material.h:149-179 */
                    float ir = m->ior;
                    v3 outward_normal;
                    v3 reflected = reflect3(cur_d, rec.normal);
                    float ni_over_nt;
                    attenuation = V(1.0f, 1.0f, 1.0f);
                    v3 refracted = V(0.f, 0.f, 0.f);
                    float reflect_prob;
                    float cosine;
                    if (vdot(cur_d, rec.normal) > 0.0f) {
                        outward_normal = vneg(rec.normal);
                        ni_over_nt = ir;
                        cosine = vdot(cur_d, rec.normal) / vlength(cur_d);
                        cosine = sqrtf(1.0f - ir * ir * (1 - cosine * cosine));
                    } else {
                        outward_normal = rec.normal;
                        ni_over_nt = 1.0f / ir;
                        cosine = -vdot(cur_d, rec.normal) / vlength(cur_d);
                    }
                    if (pto_refract(cur_d, outward_normal, ni_over_nt, &refracted)) reflect_prob = pto_schlick(cosine, ir);
                    else reflect_prob = 1.0f;
                    if (U(c) < reflect_prob) sdir = reflected;
                    else sdir = refracted;
                }
                if (!scattered) {
                    if (st) st->absorbed_paths++;
                    return V(0.f, 0.f, 0.f);
                }
                cur_attenuation = vmul(cur_attenuation, attenuation);
                cur_o = rec.p;
                cur_d = sdir;
            }
        } else {
            if (st) st->miss_paths++;
            return vmul(V(0.0f, 0.0f, 0.0f), cur_attenuation); /* camera.h:79,109: background (0,0,0) * attenuation */
        }
    }
    if (st) st->depth_paths++;
    return V(0.0f, 0.0f, 0.0f); /* camera.h:82 */
}

pto_vec3 pto_ray_color(const pto_world *w, pto_vec3 o, pto_vec3 d, uint32_t depth, pto_rng *s, pto_stats *st) {
    rng_ctx c = {s, st, NULL, 0, 0, 0.f};
    return ray_color(w, o, d, depth, &c);
}

/* Builder-defined mode with no counterpart in the reference (ptcore.h: PT_RNG_SAMPLE_KEYED): n_chunks > 0 keys the XORWOW stream by
 * (pixel, sample) — sample s of pixel p draws from curand_init(splitmix64(1984 + p + s * W * H), 0, 0) — and adds the samples chunk by chunk
 * (ceil(spp / n_chunks) samples each), then the chunk sums in order, as the CUDA core does.  0 = the reference's streams (default). */
static uint32_t g_keyed_chunks = 0;
static int g_keyed_hash = 1;
void pto_set_keyed_hash(int on) { g_keyed_hash = on; }
/* seed of sample s of pixel p: the counter 1984 + p + s * W * H, passed through the splitmix64 finaliser so that the 64 bits XORWOW's
 * curand_init(seed, 0, 0) derives its whole state from (curand_kernel.h:772-797) are well mixed even though consecutive counters differ
 * in a few low bits only */
uint64_t pto_keyed_seed(uint32_t pixel_index, uint32_t sample, uint32_t npix) {
    uint64_t z = 1984ull + (uint64_t)pixel_index + (uint64_t)sample * (uint64_t)npix;
    if (!g_keyed_hash) return z;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void pto_set_keyed_chunks(uint32_t n_chunks) { g_keyed_chunks = n_chunks; }

int pto_render(const pto_world *w, const PtCamera *cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t depth,
               int32_t ox, int32_t oy, int32_t tw, int32_t th, uint8_t *rgb, uint8_t *yuv, float *accum, pto_stats *st, int threads) {
    if (!w || !cam || !rgb || width == 0 || height == 0) return -1;
    cam_params cp = camera_params(cam);
    pto_stats total;
    memset(&total, 0, sizeof total);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel
    {
        pto_stats local;
        memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 1)
        for (int32_t j = 0; j < th; j++) {
            for (int32_t i = 0; i < tw; i++) {
                /* DevicePathTracer.h:74-105 */
                int32_t x = ox + i, y = oy + j;
                if (x < 0 || y < 0 || x >= (int32_t)width || y >= (int32_t)height) continue;
                int32_t pixel_index = ((int32_t)height - y - 1) * (int32_t)width + x;
                pto_rng rs;
                pto_rng_init(&rs, (uint64_t)(int64_t)(1984 + pixel_index)); /* render_init, :54 */
                rng_ctx c = {&rs, &local, NULL, 0, 0, 0.f};
                v3 col = V(0, 0, 0);
                const uint32_t chunk_spp = g_keyed_chunks ? (spp + g_keyed_chunks - 1) / g_keyed_chunks : 0;
                v3 part = V(0, 0, 0);
                for (uint32_t s = 0; s < spp; s++) {
                    if (chunk_spp) pto_rng_init(&rs, pto_keyed_seed((uint32_t)pixel_index, s, width * height));
                    float u = (float)(x + U(&c)) / (float)width;
                    float v = (float)(y + U(&c)) / (float)height;
                    v3 ro, rd;
                    camera_get_ray(&cp, u, v, &ro, &rd);
                    local.samples++;
                    if (chunk_spp) {
                        part = vadd(part, ray_color(w, ro, rd, depth, &c));
                        if ((s + 1) % chunk_spp == 0 || s + 1 == spp) { col = vadd(col, part); part = V(0, 0, 0); }
                    } else {
                        col = vadd(col, ray_color(w, ro, rd, depth, &c));
                    }
                }
                float cc[3] = {col.x, col.y, col.z};
                uint8_t q[3];
                pto_quantise(cc, spp, q);
                size_t pi = (size_t)pixel_index;
                rgb[3 * pi] = q[0]; rgb[3 * pi + 1] = q[1]; rgb[3 * pi + 2] = q[2];
                if (accum) { accum[3 * pi] = col.x; accum[3 * pi + 1] = col.y; accum[3 * pi + 2] = col.z; }
                if (yuv) {
                    uint8_t Y, Uc, Vc;
                    pto_yuv(q, &Y, &Uc, &Vc);
                    yuv[pi] = Y;
                    int blockRow = pixel_index / (int32_t)width, blockCol = pixel_index % (int32_t)width;
                    if (blockRow % 2 == 0 && blockCol % 2 == 0) {
                        int totalPixels = (int)(width * height);
                        int uvSize = totalPixels / 4;
                        int uvIndex = (blockRow / 2) * ((int)width / 2) + (blockCol / 2);
                        /* odd width/height: the reference's index leaves its own W*H + 2*(W*H/4) buffer (UB); dropped */
                        int limit = totalPixels + 2 * uvSize;
                        if (totalPixels + uvIndex < limit) yuv[totalPixels + uvIndex] = Uc;
                        if (totalPixels + uvSize + uvIndex < limit) yuv[totalPixels + uvSize + uvIndex] = Vc;
                    }
                }
            }
        }
#pragma omp critical
        {
            total.samples += local.samples; total.rays += local.rays; total.draws += local.draws;
            total.emitter_paths += local.emitter_paths; total.miss_paths += local.miss_paths;
            total.depth_paths += local.depth_paths; total.absorbed_paths += local.absorbed_paths;
        }
    }
    if (st) *st = total;
    return 0;
}

int pto_trace_pixel(const pto_world *w, const PtCamera *cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t depth, int32_t x, int32_t y,
                    float *events, int32_t max_events, float *col_out) {
    cam_params cp = camera_params(cam);
    int32_t pixel_index = ((int32_t)height - y - 1) * (int32_t)width + x;
    pto_rng rs;
    pto_rng_init(&rs, (uint64_t)(int64_t)(1984 + pixel_index));
    rng_ctx c = {&rs, NULL, events, max_events, 0, 0.f};
    v3 col = V(0, 0, 0);
    for (uint32_t s = 0; s < spp; s++) {
        float u = (float)(x + U(&c)) / (float)width;
        float v = (float)(y + U(&c)) / (float)height;
        v3 ro, rd;
        camera_get_ray(&cp, u, v, &ro, &rd);
        c.cur_sample = (float)s;
        col = vadd(col, ray_color(w, ro, rd, depth, &c));
    }
    col_out[0] = col.x; col_out[1] = col.y; col_out[2] = col.z;
    return c.n_events;
}
