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

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

namespace
{
  double nowMs()
  {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }

  // --bench N: the three adapter systems against the reference's own three on ONE World of N entities in groups of
  // five (root + four parented children, all with RenderMesh + Bounds), 10 % of the Transforms dirtied per frame
  // through sc::setLocal. Prints one JSON object: host milliseconds per frame of each system.
  int bench(uint32_t n, uint32_t frames)
  {
    ScRefWorld* w = screfCreate(0);
    sc::World& world = w->world;
    world.reserveEntities(n);
    std::vector<sc::Entity> all;
    all.reserve(n);
    uint32_t rng = 99u;
    sc::Entity root{};
    for (uint32_t i = 0; i < n; ++i)
    {
      const sc::Entity e = world.create();
      sc::Transform& t = world.add<sc::Transform>(e);
      const bool isRoot = (i % 5u) == 0u;
      const float p[3] = { isRoot ? (float)(lcg(rng) % 8192u) - 4096.0f : 0.5f, isRoot ? 0.0f : 0.25f, isRoot ? (float)(lcg(rng) % 8192u) - 4096.0f : -0.5f };
      const float r[3] = { 0.0f, (float)(lcg(rng) % 628u) * 0.01f, 0.0f }, sc3[3] = { 1.0f, 1.0f, 1.0f };
      sc::setLocal(t, p, r, sc3);
      if (isRoot) root = e; else sc::setParent(t, root);
      sc::RenderMesh& rm = world.add<sc::RenderMesh>(e);
      rm.meshId = 1 + i % 3; rm.materialId = 1 + i % 4;
      world.add<sc::Bounds>(e).localAabb = { { -0.5f, -0.5f, -0.5f }, { 0.5f, 0.5f, 0.5f } };
      all.push_back(e);
    }
    const sc::Entity cam = world.create();
    {
      sc::Transform& t = world.add<sc::Transform>(cam);
      const float p[3] = { 32.0f, 6.0f, 44.0f }, r[3] = { 0.0f, 3.14159265f, 0.0f }, s3[3] = { 1.0f, 1.0f, 1.0f };
      sc::setLocal(t, p, r, s3);
      sc::Camera& c = world.add<sc::Camera>(cam);
      c.active = true;
    }
    sc::gpu::GpuSceneState gs{};
    gs.maxInstances = n + 16;
    gs.leaveDirtyFlags = true;
    if (!sc::gpu::init(gs)) { std::printf("gpu init failed: %s\n", gs.lastError); return 2; }
    sc::RenderFrameData gpuFrame{};
    sc::CullingState gpuCull{};
    gpuCull.frame = &gpuFrame;
    sc::RenderPrepStreamingState gpuPrep{};
    gpuPrep.frame = &gpuFrame;
    gpuPrep.culling = &gpuCull;
    gpuPrep.streaming = w->streaming;
    sc::gpu::GpuCullingState gc{ &gs, &gpuCull, false, true };
    sc::gpu::GpuRenderPrepState gp{ &gs, &gpuPrep };
    double tGpuT = 0, tGpuC = 0, tGpuP = 0, tRefT = 0, tRefC = 0, tRefP = 0;
    uint32_t timed = 0, visible = 0;
    bool same = true;
    for (uint32_t f = 0; f < frames + 2; ++f)
    {
      if (f > 0)
        for (uint32_t k = 0; k < n / 10u; ++k)
        {
          sc::Transform* t = world.get<sc::Transform>(all[lcg(rng) % n]);
          const float p[3] = { t->localPos[0] + 0.25f, t->localPos[1], t->localPos[2] }, r[3] = { 0.0f, t->localRot[1] + 0.1f, 0.0f };
          sc::setLocal(*t, p, r, t->localScale);
        }
      const double a0 = nowMs();
      sc::gpu::TransformSystem(world, 1.0f / 60.0f, &gs);
      const double a1 = nowMs();
      sc::jobs().beginFrame();
      sc::TransformSystem(world, 1.0f / 60.0f, nullptr);
      const double a2 = nowMs();
      sc::CameraSystem(world, 1.0f / 60.0f, &w->camera);
      const double a3 = nowMs();
      sc::CullingSystem(world, 1.0f / 60.0f, &w->culling);
      const double a4 = nowMs();
      sc::RenderPrepStreamingSystem(world, 1.0f / 60.0f, &w->renderPrep);
      const double a5 = nowMs();
      sc::jobs().publishFrameTelemetry();
      gpuFrame.viewProj = world.renderFrame().viewProj;
      const double b0 = nowMs();
      sc::gpu::CullingSystem(world, 1.0f / 60.0f, &gc);
      const double b1 = nowMs();
      sc::gpu::RenderPrepStreamingSystem(world, 1.0f / 60.0f, &gp);
      const double b2 = nowMs();
      same = same && w->culling.visible.size() == gpuCull.visible.size() &&
             std::memcmp(w->culling.visible.data(), gpuCull.visible.data(), gpuCull.visible.size() * 4) == 0 &&
             w->culling.candidates.size() == gpuCull.candidates.size() && world.renderFrame().draws.size() == gpuFrame.draws.size();
      visible = gpuCull.stats.visible;
      if (f >= 2)
      {
        tGpuT += a1 - a0; tRefT += a2 - a1; tRefC += a4 - a3; tRefP += a5 - a4; tGpuC += b1 - b0; tGpuP += b2 - b1;
        ++timed;
      }
    }
    const double k = 1.0 / timed;
    std::printf("{\"entities\": %u, \"frames\": %u, \"dirty_fraction\": 0.1, \"visible\": %u, \"outputs_equal\": %s, \"resyncs\": %llu, "
                "\"adapter_ms\": {\"TransformSystem\": %.3f, \"CullingSystem\": %.3f, \"RenderPrepStreamingSystem\": %.3f}, "
                "\"reference_ms\": {\"TransformSystem\": %.3f, \"CullingSystem\": %.3f, \"RenderPrepStreamingSystem\": %.3f}, "
                "\"reference_threads\": %u}\n",
                n, timed, visible, same ? "true" : "false", (unsigned long long)gs.resyncs, tGpuT * k, tGpuC * k, tGpuP * k, tRefT * k,
                tRefC * k, tRefP * k, screfJobWorkers() + 1u);
    sc::gpu::shutdown(gs);
    screfDestroy(w);
    return same ? 0 : 1;
  }
}

int main(int argc, char** argv)
{
  if (argc >= 3 && std::strcmp(argv[1], "--bench") == 0)
    return bench((uint32_t)std::strtoul(argv[2], nullptr, 10), argc >= 4 ? (uint32_t)std::strtoul(argv[3], nullptr, 10) : 8u);
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
  sc::gpu::GpuCullingState gc{ &gs, &gpuCull, true, true };
  sc::gpu::GpuRenderPrepState gp{ &gs, &gpuPrep };

  uint32_t rng = 12345u;
  std::vector<sc::Entity> dense;
  std::vector<sc::Entity> lateOnes;  // Transform-only entities that get RenderMesh / Bounds later (sc_traffic_lod.cpp:47-70)
  const uint32_t kFrames = 12;
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
    if (frame == 8)
      for (int k = 0; k < 9; ++k)   // Transform now, render components later
      {
        const sc::Entity e = world.create();
        sc::Transform& t = world.add<sc::Transform>(e);
        const float p[3] = { 30.0f + k, 1.0f, 40.0f - k }, r[3] = { 0.0f, 0.2f * k, 0.0f }, s[3] = { 1.0f, 1.0f, 1.0f };
        sc::setLocal(t, p, r, s);
        lateOnes.push_back(e);
      }
    if (frame == 9)
      for (size_t k = 0; k < lateOnes.size(); ++k)   // late World::add<RenderMesh / Bounds> on a live Transform
      {
        sc::RenderMesh& rm = world.add<sc::RenderMesh>(lateOnes[k]);
        rm.meshId = 2; rm.materialId = 3 + (uint32_t)k;
        if (k % 2 == 0) world.add<sc::Bounds>(lateOnes[k]).localAabb = { { -2, -1, -2 }, { 2, 1, 2 } };
      }
    if (frame == 10)
    {
      for (size_t k = 0; k < lateOnes.size(); ++k)   // LOD swap: mesh / material ids change, a Bounds box grows
      {
        sc::RenderMesh* rm = world.get<sc::RenderMesh>(lateOnes[k]);
        rm->meshId = 4; rm->materialId += 10;
        if (sc::Bounds* b = world.get<sc::Bounds>(lateOnes[k])) b->localAabb.max.y = 3.0f;
      }
      world.get<sc::RenderMesh>(w->spawner.cube)->materialId = 9;   // sc_imgui.cpp:720
    }
    if (frame == 11)
    {
      world.remove<sc::RenderMesh>(lateOnes[1]);   // no longer a candidate
      world.remove<sc::Bounds>(lateOnes[0]);       // always visible from now on
    }
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
    check(a.candidates.size() == gpuCull.candidates.size() &&
          std::memcmp(a.candidates.data(), gpuCull.candidates.data(), a.candidates.size() * 4) == 0, "CullingState::candidates (ordered)", frame);
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
  {
    // RenderPrepStreamingSystem without a culling stage (state->culling == nullptr, .cpp:1330-1346): every Transform +
    // RenderMesh entity in pool order, budget applied
    sc::RenderFrameData refFrame{};
    sc::RenderPrepStreamingState refPrep = w->renderPrep;
    refPrep.frame = &refFrame;
    refPrep.culling = nullptr;
    refPrep.assets = nullptr;
    gpuPrep.culling = nullptr;
    for (const uint32_t budget : { 0u, 100u })
    {
      w->streaming->budgets.maxDrawsBudget = budget;
      sc::RenderPrepStreamingSystem(world, 1.0f / 60.0f, &refPrep);
      sc::gpu::RenderPrepStreamingSystem(world, 1.0f / 60.0f, &gp);
      bool ok = refFrame.draws.size() == gpuFrame.draws.size() && refPrep.stats.drawsEmitted == gpuPrep.stats.drawsEmitted &&
                refPrep.stats.drawsDroppedByBudget == gpuPrep.stats.drawsDroppedByBudget;
      for (size_t i = 0; ok && i < refFrame.draws.size(); ++i)
      {
        ok = refFrame.draws[i].entity == gpuFrame.draws[i].entity && refFrame.draws[i].meshId == gpuFrame.draws[i].meshId &&
             refFrame.draws[i].materialId == gpuFrame.draws[i].materialId;
        for (int k = 0; ok && k < 16; ++k) ok = sameValue(refFrame.draws[i].model.m[k], gpuFrame.draws[i].model.m[k]);
      }
      check(ok, "RenderPrepStreamingSystem without a culling stage", 100u + budget);
      std::printf("no culling stage, budget %u: %zu draws, dropped %u %s\n", budget, gpuFrame.draws.size(), gpuPrep.stats.drawsDroppedByBudget,
                  g_fail ? "FAIL" : "ok");
    }
  }
  sc::gpu::shutdown(gs);
  screfDestroy(w);
  std::printf(g_fail ? "DROPIN FAILED (%d mismatches)\n" : "DROPIN OK\n", g_fail);
  return g_fail ? 1 : 0;
}
