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
#include <unordered_map>
#include <vector>

namespace sc::gpu
{
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

    // shadow of the Transform pool as the GPU knows it
    std::vector<Entity> dense;                       // pool order
    std::unordered_map<uint32_t, uint32_t> parentOf; // entity -> parent handle last uploaded
    bool transformPassDone = false;                  // a transform-only update already ran this frame
    uint64_t resyncs = 0;                            // full rebuilds (pool order could not be replayed)
    char lastError[256] = {};
  };

  struct GpuCullingState
  {
    GpuSceneState* scene = nullptr;
    CullingState* culling = nullptr;   // the reference's own state struct, filled identically
    bool fillCulledList = true;        // CullingState::culled (only DebugDraw reads it, sc_debug_draw_system.cpp:136-137)
    bool fillCandidates = false;       // CullingState::candidates (nobody reads it outside CullingSystem)
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
