// Link-only stand-ins for the Vulkan-backed AssetManager (src/engine/src/sc_assets.cpp is a Windows/Vulkan
// translation unit and is NOT on the hot path; sc_world_partition.cpp merely references these methods).
// Test infrastructure only (oracle build of the reference). The oracle always passes assets == nullptr.
#include "sc_assets.h"

namespace sc
{
  void AssetManager::beginFrame(uint64_t) {}
  void AssetManager::evictIfNeeded() {}
  void AssetManager::touchMaterial(MaterialHandle) {}
  void AssetManager::pumpTextureLoads(uint32_t) {}
  MaterialHandle AssetManager::createMaterial(const MaterialDesc&) { return 0; }
  MeshHandle AssetManager::loadMesh(const std::string&) { return 0; }
  TextureHandle AssetManager::loadTexture2D(const std::string&, bool) { return 0; }
}
