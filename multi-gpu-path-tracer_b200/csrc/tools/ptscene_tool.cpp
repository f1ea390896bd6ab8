// ptscene_tool — convert any supported scene file to .ptscene and print a summary.
//   ptscene_tool convert <in.{glb,gltf,obj,ptscene}> <out.ptscene>
//   ptscene_tool info <file>
#include "../host/HostScene.h"

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>

static void info(const HostScene &s) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    std::map<int, int> per;
    for (auto &t : s.triangles) {
        per[t.materialIdx]++;
        for (const Vertex *v : {&t.v0, &t.v1, &t.v2}) {
            const float p[3] = {v->position.x, v->position.y, v->position.z};
            for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
        }
    }
    printf("{\"triangles\": %zu, \"spheres\": %zu, \"materials\": %zu, \"textures\": %zu,\n", s.triangles.size(), s.spheres.size(), s.materials.size(), s.textures.size());
    printf(" \"bounds_min\": [%.9g, %.9g, %.9g], \"bounds_max\": [%.9g, %.9g, %.9g],\n", lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    printf(" \"per_material\": [");
    for (size_t i = 0; i < s.materials.size(); i++) printf("%s%d", i ? ", " : "", per.count((int)i) ? per[(int)i] : 0);
    printf("],\n \"materials_detail\": [");
    for (size_t i = 0; i < s.materials.size(); i++) {
        auto &m = s.materials[i];
        printf("%s{\"type\": %d, \"base\": [%.9g, %.9g, %.9g], \"emis\": [%.9g, %.9g, %.9g], \"base_tex\": %d, \"emis_tex\": %d}", i ? ", " : "", (int)m.type,
               m.baseColor.x, m.baseColor.y, m.baseColor.z, m.emissiveFactor.x, m.emissiveFactor.y, m.emissiveFactor.z, m.baseColorTextureIdx.value_or(-1), m.emissiveTextureIdx.value_or(-1));
    }
    printf("],\n \"texture_sizes\": [");
    for (size_t i = 0; i < s.textures.size(); i++) printf("%s[%d, %d]", i ? ", " : "", s.textures[i].width, s.textures[i].height);
    printf("]}\n");
}

int main(int argc, char **argv) {
    try {
        SceneLoader loader;
        if (argc == 4 && !strcmp(argv[1], "convert")) {
            HostScene s = loader.load(std::string(argv[2]));
            write_ptscene(s, argv[3]);
            info(s);
            return 0;
        }
        if (argc == 3 && !strcmp(argv[1], "info")) {
            info(loader.load(std::string(argv[2])));
            return 0;
        }
        fprintf(stderr, "usage: ptscene_tool convert <in> <out.ptscene> | info <file>\n");
        return 2;
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
