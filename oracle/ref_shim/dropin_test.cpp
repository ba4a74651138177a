// dropin_test — in-situ check of the drop-in boundary (test infrastructure; built into oracle/_ref/ because it
// links the reference's own sources). ONE sc::World, built by the reference's own streaming + spawner code (config 1,
// src/sandbox/src/main.cpp:66-99). Every frame the reference's CPU systems and the GPU adapter systems
// (sc-gameengine_b200/host/sc_gpu_systems.cpp, same void(World&,float,void*) signature) both run on it and their
// observable outputs are compared: CullingState::{visible, culled, stats, frustum}, RenderFrameData::draws,
// RenderPrepStats and every Transform::worldMatrix.
#include "scref_api.h"
#include "scref_internal.h"

#include "sc_gpu_systems.h"
#include "sc_jobs.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace
{
  int g_fail = 0;

  bool sameValue(float a, float b)
  {
    if (a != a && b != b) return true;
    return a == b;  // +0 == -0: DESIGN.md "Parity definition"
  }

  void check(bool ok, const char* what, uint32_t frame)
  {
    if (!ok) { std::printf("MISMATCH frame %u: %s\n", frame, what); ++g_fail; }
  }

  uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
}

int main()
{
  ScRefWorld* w = screfCreate(0);
  if (!w) { std::printf("screfCreate failed\n"); return 2; }
  const uint32_t sectors = screfBuildDefaultScene(w, 60);
  sc::World& world = w->world;
  std::printf("scene: %u sectors, %u transforms\n", sectors, world.componentCount<sc::Transform>());

  sc::gpu::GpuSceneState gs{};
  gs.maxInstances = 1u << 16;
  gs.leaveDirtyFlags = true;  // the CPU systems run on the same World right after
  if (!sc::gpu::init(gs)) { std::printf("gpu init failed: %s\n", gs.lastError); return 2; }

  // second set of state structs for the GPU side, same types as the reference's
  sc::RenderFrameData gpuFrame{};
  sc::CullingState gpuCull{};
  gpuCull.frame = &gpuFrame;
  sc::RenderPrepStreamingState gpuPrep{};
  gpuPrep.frame = &gpuFrame;
  gpuPrep.culling = &gpuCull;
  gpuPrep.streaming = w->streaming;
  gpuPrep.assets = nullptr;
  sc::gpu::GpuCullingState gc{ &gs, &gpuCull, true, false };
  sc::gpu::GpuRenderPrepState gp{ &gs, &gpuPrep };

  uint32_t rng = 12345u;
  std::vector<sc::Entity> dense;
  const uint32_t kFrames = 8;
  for (uint32_t frame = 0; frame < kFrames; ++frame)
  {
    dense.clear();
    world.ForEach<sc::Transform>([&](sc::Entity e, sc::Transform&) { dense.push_back(e); });

    // ---- scripted game-logic edits through the engine's own API ----
    if (frame == 1 || frame == 5)
      for (int k = 0; k < 60; ++k)
      {
        sc::Transform* t = world.get<sc::Transform>(dense[lcg(rng) % dense.size()]);
        const float p[3] = { t->localPos[0] + 3.0f, t->localPos[1], t->localPos[2] - 2.0f };
        const float r[3] = { 0.1f * k, t->localRot[1] + 0.5f, 0.0f };
        sc::setLocal(*t, p, r, t->localScale);
      }
    if (frame == 2)
    {
      sc::setParent(*world.get<sc::Transform>(w->spawner.cube), w->spawner.triangle);       // re-parent
      const float p[3] = { 1.0f, 2.0f, 3.0f }, r[3] = { 0.3f, 0.2f, 0.1f }, s[3] = { 1.0f, 1.0f, 1.0f };
      sc::setLocal(*world.get<sc::Transform>(w->spawner.root), p, r, s);                     // dirty root of a chain
    }
    if (frame == 3)
    {
      for (int k = 0; k < 40; ++k) world.destroy(dense[(7 + 13 * k) % dense.size()]);       // ascending pool order
      for (int k = 0; k < 12; ++k)
      {
        const sc::Entity e = world.create();
        sc::Transform& t = world.add<sc::Transform>(e);
        const float p[3] = { 20.0f + k, 1.0f, 30.0f - k }, r[3] = { 0.0f, 0.3f * k, 0.0f }, s[3] = { 1.0f, 2.0f, 1.0f };
        sc::setLocal(t, p, r, s);
        if (k % 3 == 0) sc::setParent(t, w->spawner.root);
        sc::RenderMesh& rm = world.add<sc::RenderMesh>(e);
        rm.meshId = 5; rm.materialId = 7 + k;
        if (k % 4 != 0) { sc::Bounds& b = world.add<sc::Bounds>(e); b.localAabb = { { -1, -1, -1 }, { 1, 2, 1 } }; }
      }
    }
    if (frame == 4)
      for (int k = 0; k < 25; ++k) world.destroy(dense[dense.size() - 1 - (3 * k) % dense.size()]);  // arbitrary order
    w->culling.freezeCulling = gpuCull.freezeCulling = (frame == 6);
    w->streaming->budgets.maxDrawsBudget = (frame == 7) ? 50u : 6000u;

    // ---- GPU adapter: delta upload (reads dirty flags), then the reference CPU chain, then GPU cull/prep ----
    sc::gpu::TransformSystem(world, 1.0f / 60.0f, &gs);
    sc::jobs().beginFrame();
    sc::TransformSystem(world, 1.0f / 60.0f, nullptr);
    sc::CameraSystem(world, 1.0f / 60.0f, &w->camera);
    sc::CullingSystem(world, 1.0f / 60.0f, &w->culling);
    sc::RenderPrepStreamingSystem(world, 1.0f / 60.0f, &w->renderPrep);
    sc::jobs().publishFrameTelemetry();

    gpuFrame.viewProj = world.renderFrame().viewProj;
    sc::gpu::CullingSystem(world, 1.0f / 60.0f, &gc);
    sc::gpu::RenderPrepStreamingSystem(world, 1.0f / 60.0f, &gp);

    // ---- compare ----
    const sc::CullingState& a = w->culling;
    check(a.stats.renderablesTotal == gpuCull.stats.renderablesTotal && a.stats.visible == gpuCull.stats.visible &&
          a.stats.culled == gpuCull.stats.culled, "CullingStats", frame);
    check(a.visible.size() == gpuCull.visible.size() &&
          std::memcmp(a.visible.data(), gpuCull.visible.data(), a.visible.size() * 4) == 0, "CullingState::visible (ordered)", frame);
    check(a.culled.size() == gpuCull.culled.size() &&
          std::memcmp(a.culled.data(), gpuCull.culled.data(), a.culled.size() * 4) == 0, "CullingState::culled (ordered)", frame);
    if (!a.freezeCulling)
      check(std::memcmp(a.frustum.planes, gpuCull.frustum.planes, sizeof(a.frustum.planes)) == 0, "CullingState::frustum", frame);
    const std::vector<sc::DrawItem>& da = world.renderFrame().draws;
    const std::vector<sc::DrawItem>& db = gpuFrame.draws;
    bool drawsOk = da.size() == db.size();
    for (size_t i = 0; drawsOk && i < da.size(); ++i)
    {
      drawsOk = da[i].entity == db[i].entity && da[i].meshId == db[i].meshId && da[i].materialId == db[i].materialId;
      for (int k = 0; drawsOk && k < 16; ++k) drawsOk = sameValue(da[i].model.m[k], db[i].model.m[k]);
    }
    check(drawsOk, "RenderFrameData::draws", frame);
    check(w->renderPrep.stats.drawsEmitted == gpuPrep.stats.drawsEmitted &&
          w->renderPrep.stats.drawsDroppedByBudget == gpuPrep.stats.drawsDroppedByBudget, "RenderPrepStats", frame);

    dense.clear();
    world.ForEach<sc::Transform>([&](sc::Entity e, sc::Transform&) { dense.push_back(e); });
    std::vector<float> gw(dense.size() * 16);
    check(scgpuReadWorld(gs.ctx, (uint32_t)dense.size(), reinterpret_cast<const uint32_t*>(dense.data()), gw.data()) == 1,
          "scgpuReadWorld", frame);
    bool worldOk = true;
    for (size_t i = 0; worldOk && i < dense.size(); ++i)
    {
      const sc::Transform* t = world.get<sc::Transform>(dense[i]);
      for (int k = 0; worldOk && k < 16; ++k) worldOk = sameValue(t->worldMatrix.m[k], gw[i * 16 + k]);
    }
    check(worldOk, "Transform::worldMatrix of every entity", frame);
    std::printf("frame %u: %zu transforms, visible %u culled %u draws %zu dropped %u resyncs %llu %s\n", frame, dense.size(),
                gpuCull.stats.visible, gpuCull.stats.culled, db.size(), gpuPrep.stats.drawsDroppedByBudget,
                (unsigned long long)gs.resyncs, g_fail ? "FAIL" : "ok");
  }
  sc::gpu::shutdown(gs);
  screfDestroy(w);
  std::printf(g_fail ? "DROPIN FAILED (%d mismatches)\n" : "DROPIN OK\n", g_fail);
  return g_fail ? 1 : 0;
}
