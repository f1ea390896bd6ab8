// HostScene.h — host-side scene container and loader.
//
// Mirrors the public types of the reference's src/HostScene.h:19-59 (material_type,
// Vertex, HostMaterial, Triangle, HostTexture, HostScene) and the SceneLoader::load
// entry point (src/HostScene.h:61-63, src/HostScene.cpp:98-139) so a caller of the
// reference compiles against this header unchanged.  The implementation is new:
// the reference delegates to assimp 5.4.3 + stb_image, neither of which exists
// here, so SceneLoader carries its own GLB/glTF reader, PNG decoder (zlib) and a
// small Wavefront OBJ/MTL reader (material-name routing per reference
// README.md:60-76 / src/obj_loader.h:65-96).
//
// Additions over the reference (all defaulted, so reference callers are unaffected):
//   HostMaterial::type/fuzz/ior   – routing for lambertian/metal/dielectric/diffuse_light
//   HostScene::spheres            – analytic spheres (reference src/sphere.h is dead code)
//   read_ptscene/write_ptscene    – flat binary interchange used by oracle/ and tests/
#pragma once

#include <vector_types.h>
#include <vector_functions.h>

#include <optional>
#include <string>
#include <vector>

// Same enumerators and order as reference src/HostScene.h:20-26.
enum material_type {
    LAMBERTIAN,
    METAL,
    DIELECTRIC,
    DIFFUSE_LIGHT,
    UNIVERSAL
};

struct Vertex {
    float3 position;
    float2 texCoords;
};

struct HostMaterial {
    float3 baseColor{1.f, 1.f, 1.f};
    std::optional<int> baseColorTextureIdx{};
    float3 emissiveFactor{0.f, 0.f, 0.f};
    std::optional<int> emissiveTextureIdx{};
    // --- additions ---
    material_type type = UNIVERSAL;
    float fuzz = 0.f;  // metal
    float ior = 1.5f;  // dielectric
    std::string name{};
};

struct Triangle {
    Vertex v0;
    Vertex v1;
    Vertex v2;
    int textureIdx = -1;
    int materialIdx = 0;
};

struct HostTexture {
    int width = 0;
    int height = 0;
    std::vector<float3> data;  // RGB in 0..255, row 0 = top row of the image file
};

struct HostSphere {
    float3 center;
    float radius;
    int materialIdx;
};

struct HostScene {
    std::vector<Triangle> triangles{};
    std::vector<HostTexture> textures{};
    std::vector<HostMaterial> materials{};
    std::vector<HostSphere> spheres{};  // addition
};

class SceneLoader {
public:
    // .glb / .gltf / .obj / .ptscene by extension; throws std::runtime_error otherwise
    // (reference src/HostScene.cpp:110-135 throws for unknown types the same way).
    HostScene load(std::string &path);
    HostScene load(const std::string &path) {
        std::string p = path;
        return load(p);
    }

private:
    HostScene loadGLTF(const std::string &path, bool binary);
    HostScene loadOBJ(const std::string &path);
};

// Flat little-endian interchange format ("PTSC" v1); see DESIGN.md §data formats.
void write_ptscene(const HostScene &scene, const std::string &path);
HostScene read_ptscene(const std::string &path);

// PNG → 8-bit samples.  channels is what stb_image would report with req_comp=0.
bool decode_png(const unsigned char *bytes, size_t n, int &width, int &height, int &channels,
                std::vector<unsigned char> &out, std::string &err);
