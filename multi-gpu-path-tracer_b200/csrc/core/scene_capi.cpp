// scene_capi.cpp — C ABI over the scene loader (include/ptcore.h: ptscene_*).
#include "../../../include/ptcore.h"
#include "../host/HostScene.h"

#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

struct ptscene {
    HostScene scene;
    std::vector<float> tri_pos, tri_uv, sph;
    std::vector<int32_t> tri_mat, sph_mat;
    std::vector<PtMaterial> mats;
    std::vector<PtTexture> tex;
    PtSceneDesc desc{};
};

// HostScene -> flat arrays.  Texture indices are passed through as they are: the reference's
// "sticky" texture pointer (src/DevicePathTracer.h:269-279) is applied by ptcore_upload_scene itself,
// the one place every front door goes through.
void ptscene_flatten(ptscene *s) {
    const HostScene &h = s->scene;
    size_t n = h.triangles.size();
    s->tri_pos.resize(n * 9);
    s->tri_uv.resize(n * 6);
    s->tri_mat.resize(n);
    for (size_t i = 0; i < n; i++) {
        const Triangle &t = h.triangles[i];
        const Vertex *v[3] = {&t.v0, &t.v1, &t.v2};
        for (int k = 0; k < 3; k++) {
            s->tri_pos[i * 9 + 3 * k + 0] = v[k]->position.x;
            s->tri_pos[i * 9 + 3 * k + 1] = v[k]->position.y;
            s->tri_pos[i * 9 + 3 * k + 2] = v[k]->position.z;
            s->tri_uv[i * 6 + 2 * k + 0] = v[k]->texCoords.x;
            s->tri_uv[i * 6 + 2 * k + 1] = v[k]->texCoords.y;
        }
        s->tri_mat[i] = t.materialIdx;
    }
    s->sph.resize(h.spheres.size() * 4);
    s->sph_mat.resize(h.spheres.size());
    for (size_t i = 0; i < h.spheres.size(); i++) {
        s->sph[i * 4 + 0] = h.spheres[i].center.x;
        s->sph[i * 4 + 1] = h.spheres[i].center.y;
        s->sph[i * 4 + 2] = h.spheres[i].center.z;
        s->sph[i * 4 + 3] = h.spheres[i].radius;
        s->sph_mat[i] = h.spheres[i].materialIdx;
    }
    s->mats.resize(h.materials.size());
    for (size_t i = 0; i < h.materials.size(); i++) {
        const HostMaterial &m = h.materials[i];
        PtMaterial &o = s->mats[i];
        o.type = (int32_t)m.type;
        o.base[0] = m.baseColor.x; o.base[1] = m.baseColor.y; o.base[2] = m.baseColor.z;
        o.emis[0] = m.emissiveFactor.x; o.emis[1] = m.emissiveFactor.y; o.emis[2] = m.emissiveFactor.z;
        o.base_tex = m.baseColorTextureIdx.value_or(-1);
        o.emis_tex = m.emissiveTextureIdx.value_or(-1);
        o.fuzz = m.fuzz;
        o.ior = m.ior;
    }
    s->tex.resize(h.textures.size());
    for (size_t i = 0; i < h.textures.size(); i++) {
        s->tex[i].width = h.textures[i].width;
        s->tex[i].height = h.textures[i].height;
        s->tex[i].rgb = h.textures[i].data.empty() ? nullptr : &h.textures[i].data[0].x;
    }
    PtSceneDesc &d = s->desc;
    d.n_tris = (int32_t)n;
    d.tri_pos = s->tri_pos.data();
    d.tri_uv = s->tri_uv.data();
    d.tri_mat = s->tri_mat.data();
    d.n_spheres = (int32_t)h.spheres.size();
    d.sph = s->sph.data();
    d.sph_mat = s->sph_mat.data();
    d.n_mats = (int32_t)s->mats.size();
    d.mats = s->mats.data();
    d.n_tex = (int32_t)s->tex.size();
    d.tex = s->tex.data();
}

extern "C" {

int ptscene_load(const char *path, ptscene_t **out, char *err, size_t err_len) {
    if (!path || !out) return PT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    try {
        ptscene *s = new ptscene();
        SceneLoader loader;
        s->scene = loader.load(std::string(path));
        ptscene_flatten(s);
        *out = s;
        return PT_OK;
    } catch (const std::exception &e) {
        if (err && err_len) snprintf(err, err_len, "%s", e.what());
        return PT_ERR_SYSTEM;
    }
}

const PtSceneDesc *ptscene_desc(const ptscene_t *s) { return s ? &s->desc : nullptr; }

int ptscene_save(const ptscene_t *s, const char *path) {
    if (!s || !path) return PT_ERR_INVALID_ARGUMENT;
    try {
        write_ptscene(s->scene, path);
        return PT_OK;
    } catch (const std::exception &) {
        return PT_ERR_SYSTEM;
    }
}

void ptscene_free(ptscene_t *s) { delete s; }

}  // extern "C"
