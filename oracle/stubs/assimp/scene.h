// Opaque stand-ins for the assimp types the reference's HostScene.h names in declarations.
#pragma once
struct aiVector3D { float x, y, z; };
struct aiScene; struct aiMaterial; struct aiMesh; struct aiNode;
