// Empty stand-in so the reference's hot-path headers compile without assimp (absent here).
// TEST INFRASTRUCTURE ONLY (oracle/): never included by the product.
#pragma once
namespace Assimp { class Importer {}; }
