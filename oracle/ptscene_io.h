/* ptscene_io.h — header-only C reader for the flat ".ptscene" v1 interchange file and a P6 writer.
 * TEST INFRASTRUCTURE (oracle/): used only by the oracle programs, never by the product.
 * Layout (little-endian), written by multi-gpu-path-tracer_b200/csrc/host/SceneLoader.cpp:
 *   header  { char magic[4]="PTSC"; u32 version=1, n_tris, n_spheres, n_mats, n_tex; }
 *   tris    n_tris    x { f32 pos[3][3]; f32 uv[3][2]; i32 mat; i32 tex; }           68 B
 *   spheres n_spheres x { f32 c[3]; f32 r; i32 mat; }                                 20 B
 *   mats    n_mats    x { i32 type; f32 base[3]; f32 emis[3]; i32 base_tex, emis_tex; f32 fuzz, ior; }  44 B
 *   tex     n_tex     x { i32 w, h; u8 rgb[h][w][3]; }   (row 0 = top row of the image file)
 */
#ifndef PTSCENE_IO_H
#define PTSCENE_IO_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float pos[9]; float uv[6]; int32_t mat; int32_t tex; } pts_tri;
typedef struct { float c[3]; float r; int32_t mat; } pts_sphere;
typedef struct { int32_t type; float base[3]; float emis[3]; int32_t base_tex; int32_t emis_tex; float fuzz; float ior; } pts_mat;
typedef struct { int32_t w, h; float *rgb; /* w*h*3 floats, 0..255 */ } pts_tex;
typedef struct {
    uint32_t n_tris, n_spheres, n_mats, n_tex;
    pts_tri *tris;
    pts_sphere *spheres;
    pts_mat *mats;
    pts_tex *tex;
} pts_scene;

static int pts_load(const char *path, pts_scene *s) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    struct { char magic[4]; uint32_t version, n_tris, n_spheres, n_mats, n_tex; } h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "PTSC", 4) != 0 || h.version != 1) { fclose(f); return -2; }
    memset(s, 0, sizeof *s);
    s->n_tris = h.n_tris; s->n_spheres = h.n_spheres; s->n_mats = h.n_mats; s->n_tex = h.n_tex;
    s->tris = (pts_tri *)malloc(sizeof(pts_tri) * (h.n_tris ? h.n_tris : 1));
    s->spheres = (pts_sphere *)malloc(sizeof(pts_sphere) * (h.n_spheres ? h.n_spheres : 1));
    s->mats = (pts_mat *)malloc(sizeof(pts_mat) * (h.n_mats ? h.n_mats : 1));
    s->tex = (pts_tex *)calloc(h.n_tex ? h.n_tex : 1, sizeof(pts_tex));
    if (h.n_tris && fread(s->tris, sizeof(pts_tri), h.n_tris, f) != h.n_tris) { fclose(f); return -3; }
    if (h.n_spheres && fread(s->spheres, sizeof(pts_sphere), h.n_spheres, f) != h.n_spheres) { fclose(f); return -3; }
    if (h.n_mats && fread(s->mats, sizeof(pts_mat), h.n_mats, f) != h.n_mats) { fclose(f); return -3; }
    for (uint32_t i = 0; i < h.n_tex; i++) {
        int32_t wh[2];
        if (fread(wh, sizeof wh, 1, f) != 1) { fclose(f); return -3; }
        size_t n = (size_t)wh[0] * (size_t)wh[1] * 3;
        unsigned char *px = (unsigned char *)malloc(n ? n : 1);
        if (n && fread(px, 1, n, f) != n) { free(px); fclose(f); return -3; }
        s->tex[i].w = wh[0]; s->tex[i].h = wh[1];
        s->tex[i].rgb = (float *)malloc(sizeof(float) * (n ? n : 1));
        for (size_t k = 0; k < n; k++) s->tex[i].rgb[k] = (float)px[k];
        free(px);
    }
    fclose(f);
    return 0;
}

static void pts_free(pts_scene *s) {
    for (uint32_t i = 0; i < s->n_tex; i++) free(s->tex[i].rgb);
    free(s->tris); free(s->spheres); free(s->mats); free(s->tex);
    memset(s, 0, sizeof *s);
}

static int pts_write_ppm(const char *path, const uint8_t *rgb, int w, int h) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    fwrite(rgb, 1, (size_t)w * (size_t)h * 3, f);
    fclose(f);
    return 0;
}

#endif
