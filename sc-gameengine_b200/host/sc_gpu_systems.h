// sc_gpu_systems.h — drop-in replacements for the reference's RenderPrep chain, same plugin signature
// void(World&, float dt, void* user) (src/core/include/sc_scheduler.h:38), same user-state types for the culling
// and render-prep stages, same observable outputs. They are registered exactly where the reference registers its
// own (src/sandbox/src/main.cpp:256-259):
//
//   scheduler.addSystem("Transform",  RenderPrep, sc::gpu::TransformSystem,           &gpuScene,    {...});
//   scheduler.addSystem("Camera",     RenderPrep, sc::CameraSystem,                   &cameraState, {"Transform"});
//   scheduler.addSystem("Culling",    RenderPrep, sc::gpu::CullingSystem,             &gpuCulling,  {"Camera"});
//   scheduler.addSystem("RenderPrep", RenderPrep, sc::gpu::RenderPrepStreamingSystem, &gpuPrep,     {"Culling"});
//
// Compiled against the engine's own headers (sc_ecs.h, sc_world_partition.h); calls only include/scgpu.h.
#pragma once

#include "sc_ecs.h"
#include "sc_world_partition.h"

#include "scgpu.h"

#include <cstdint>
#include <vector>

namespace sc::gpu
{
  // What the GPU side was last told about the owner of one entity INDEX (Entity::index(), 24 bits): a dense table, no
  // hashing anywhere in the per-frame path.
  struct EntityShadow
  {
    uint32_t handle = kInvalidEntity.value;  // the Transform owner uploaded for this index, or invalid
    uint32_t parent = kInvalidEntity.value;  // parent handle last uploaded
    uint32_t pos = 0;                        // position in GpuSceneState::dense
    uint32_t seen = 0;                       // frame stamp of the last Transform pass that met it
    uint32_t meshId = 0, materialId = 0;     // RenderMesh last uploaded
    float aabb[6] = { -0.5f, -0.5f, -0.5f, 0.5f, 0.5f, 0.5f };  // Bounds last uploaded
    uint32_t flags = 0;                      // SCGPU_HAS_* last uploaded
  };

  // One per World. Owns the scgpu context and the host-side shadow needed to turn the engine's "write the
  // component, set dirty" convention (sc_ecs.h:73-96) into delta batches.
  struct GpuSceneState
  {
    ScGpuScene* ctx = nullptr;
    uint32_t maxInstances = 1u << 20;
    uint32_t maxViews = 1;
    int device = 0;

    // test hook: keep Transform::dirty set after the upload so the CPU systems can run on the same World
    bool leaveDirtyFlags = false;
    // write every world matrix back into Transform::worldMatrix (cameras are always written back)
    bool readBackAllWorldMatrices = false;
    // follow World::add / remove of RenderMesh and Bounds and edits of their fields on entities that already own a
    // Transform (sc_traffic_lod.cpp:47-70, sc_imgui.cpp:720): two more component look-ups per entity and frame
    bool trackRenderComponents = true;

    // shadow of the Transform pool as the GPU knows it
    std::vector<uint32_t> dense;          // entity handles in pool order
    std::vector<EntityShadow> shadow;     // by Entity::index()
    uint32_t stamp = 0;                   // frame counter of the passes
    bool transformPassDone = false;       // a transform-only update already ran this frame
    uint64_t resyncs = 0;                 // full rebuilds (pool order could not be replayed)
    double lastHostMs = 0.0;              // wall time of the last TransformSystem call (host bookkeeping + enqueues)
    char lastError[256] = {};

    // per-frame work arrays (kept: no allocation in steady state)
    std::vector<uint32_t> cur, removed, fresh, dirtyE, reparentE, reparentP, renderE, renderMM, renderF, needWorld;
    std::vector<float> dirtyTrs, renderBB, worldOut;
  };

  struct GpuCullingState
  {
    GpuSceneState* scene = nullptr;
    CullingState* culling = nullptr;   // the reference's own state struct, filled identically
    bool fillCulledList = true;        // CullingState::culled (only DebugDraw reads it, sc_debug_draw_system.cpp:136-137)
    bool fillCandidates = true;        // CullingState::candidates (.cpp:1206-1210; nobody reads it outside CullingSystem)
  };

  struct GpuRenderPrepState
  {
    GpuSceneState* scene = nullptr;
    RenderPrepStreamingState* prep = nullptr;  // the reference's own state struct
  };

  bool init(GpuSceneState& s);
  void shutdown(GpuSceneState& s);

  void TransformSystem(World& world, float dt, void* user);            // user: GpuSceneState*
  void CullingSystem(World& world, float dt, void* user);              // user: GpuCullingState*
  void RenderPrepStreamingSystem(World& world, float dt, void* user);  // user: GpuRenderPrepState*
}
