// scref_renderer.cpp — SURVEY.md §8(f) N1 oracle: the reference's OWN per-frame draw submission loop
// (src/engine/src/sc_vk.cpp:1841-1912: drop draws with an out-of-range mesh or an unknown material, std::sort by
// (pipelineId, materialId, meshId), bind pipeline / material / mesh on change, one vkCmdDrawIndexed per item), compiled
// from the reference source where it lies. oracle/Makefile cuts exactly that block out of sc_vk.cpp into
// oracle/_ref/sc_vk_draw_loop.inc at build time (git-ignored; nothing of the reference is stored in this repository)
// and this file includes it as the body of a member function of a stand-in for VkRenderer that owns only what the
// block touches. The Vulkan entry points it calls are recorders. TEST INFRASTRUCTURE ONLY.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "sc_assets.h"
#include "sc_ecs.h"

namespace
{
  enum { VK_PIPELINE_BIND_POINT_GRAPHICS = 0, VK_INDEX_TYPE_UINT32 = 1, VK_SHADER_STAGE_VERTEX_BIT = 1 };

  struct Recorder
  {
    const sc::DrawItem* first = nullptr;
    uint32_t pending = 0;  // binds since the last draw: 1 pipeline, 2 material, 4 mesh
    const void* model = nullptr;
    std::vector<uint32_t> order;
    std::vector<uint8_t> binds;
  };
  Recorder* g_rec = nullptr;

  void vkCmdBindPipeline(VkCommandBuffer, int, VkPipeline) { g_rec->pending |= 1u; }
  void vkCmdBindDescriptorSets(VkCommandBuffer, int, VkPipelineLayout, uint32_t firstSet, uint32_t, const VkDescriptorSet*, uint32_t,
                               const uint32_t*)
  {
    if (firstSet == 1u) g_rec->pending |= 2u;  // set 0 is the per-frame global set that follows every pipeline bind
  }
  void vkCmdBindVertexBuffers(VkCommandBuffer, uint32_t, uint32_t, const VkBuffer*, const VkDeviceSize*) { g_rec->pending |= 4u; }
  void vkCmdBindIndexBuffer(VkCommandBuffer, VkBuffer, VkDeviceSize, int) {}
  void vkCmdPushConstants(VkCommandBuffer, VkPipelineLayout, int, uint32_t, uint32_t, const void* data) { g_rec->model = data; }
  void vkCmdDrawIndexed(VkCommandBuffer, uint32_t, uint32_t, uint32_t, int32_t, uint32_t)
  {
    const char* item = (const char*)g_rec->model - offsetof(sc::DrawItem, model);
    g_rec->order.push_back((uint32_t)((const sc::DrawItem*)item - g_rec->first));
    g_rec->binds.push_back((uint8_t)g_rec->pending);
    g_rec->pending = 0;
  }
}

namespace sc
{
  struct GpuMesh  // the three fields of VkRenderer's GpuMesh the block reads
  {
    VkBuffer vertexBuffer = VK_NULL_HANDLE;
    VkBuffer indexBuffer = VK_NULL_HANDLE;
    uint32_t indexCount = 0;
  };

  struct RendererLoopHarness
  {
    struct DebugUi { bool isTrianglePaused() const { return false; } } m_debugUI;
    struct Assets
    {
      std::vector<Material> materials;
      std::vector<uint8_t> known;
      const Material* getMaterial(MaterialHandle h) const { return (h < materials.size() && known[h]) ? &materials[h] : nullptr; }
    } m_assets;
    const RenderFrameData* m_renderFrame = nullptr;
    std::vector<GpuMesh> m_meshes;
    VkPipeline m_unlitPipeline = (VkPipeline)0x10, m_texturedPipeline = (VkPipeline)0x20;
    VkPipelineLayout m_pipelineLayout = VK_NULL_HANDLE;
    VkDescriptorSet m_globalSets[2] = { VK_NULL_HANDLE, VK_NULL_HANDLE };
    uint32_t m_frameIndex = 0;

    void submit(VkCommandBuffer cmd)
    {
#include "sc_vk_draw_loop.inc"
    }
  };
}

extern "C" __attribute__((visibility("default")))
// draws80: sc::DrawItem records (RenderFrameData::draws). materialPipeline[m]: PipelineId of material m (0 = UnlitColor,
// 1 = Textured) or 0xFFFFFFFF when getMaterial(m) is null. outOrder[k] = index of the k-th submitted draw, outBinds[k] =
// what was bound before it (1 pipeline | 2 material | 4 mesh). Returns the number of draws submitted.
uint32_t screfRendererSubmit(const void* draws80, uint32_t nDraws, const uint32_t* materialPipeline, uint32_t nMaterials, uint32_t meshCount,
                             uint32_t* outOrder, uint8_t* outBinds)
{
  static_assert(sizeof(sc::DrawItem) == 80, "DrawItem layout");
  sc::RenderFrameData frame;
  frame.draws.resize(nDraws);
  if (nDraws) std::memcpy(frame.draws.data(), draws80, (size_t)nDraws * sizeof(sc::DrawItem));
  sc::RendererLoopHarness h;
  h.m_renderFrame = &frame;
  h.m_meshes.resize(meshCount);
  for (sc::GpuMesh& m : h.m_meshes) m.indexCount = 3;
  h.m_assets.materials.resize(nMaterials);
  h.m_assets.known.resize(nMaterials);
  for (uint32_t m = 0; m < nMaterials; ++m)
  {
    h.m_assets.known[m] = materialPipeline[m] != 0xFFFFFFFFu;
    h.m_assets.materials[m].pipelineId = (sc::PipelineId)(materialPipeline[m] == 0xFFFFFFFFu ? 0u : materialPipeline[m]);
  }
  Recorder rec;
  rec.first = frame.draws.data();
  g_rec = &rec;
  h.submit(VK_NULL_HANDLE);
  g_rec = nullptr;
  if (!rec.order.empty())
  {
    std::memcpy(outOrder, rec.order.data(), rec.order.size() * 4);
    std::memcpy(outBinds, rec.binds.data(), rec.binds.size());
  }
  return (uint32_t)rec.order.size();
}
