// scref — C driver around the UNMODIFIED reference sources. Test infrastructure only: see scref_api.h.
// Compiled only by oracle/Makefile, against headers under /root/reference (never copied into this repo).
#include "scref_api.h"

#include "sc_ecs.h"
#include "sc_jobs.h"
#include "sc_math.h"
#include "sc_scheduler.h"
#include "sc_time.h"
#include "sc_world_partition.h"
#include "world_format.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>

#include "scref_internal.h"

namespace
{
  bool g_jobsInit = false;
  uint32_t g_workers = 0;

  sc::Entity ent(uint32_t v) { sc::Entity e{}; e.value = v; return e; }
}

extern "C" {

ScRefWorld* screfCreate(uint32_t jobWorkers)
{
  if (!g_jobsInit)
  {
    const uint32_t hw = std::thread::hardware_concurrency();
    g_workers = jobWorkers ? jobWorkers : (hw > 1 ? hw - 1 : 1);
    if (!sc::jobs().init(g_workers))
      return nullptr;
    g_jobsInit = true;
    // the reference's static JobSystem never joins its workers on its own (main.cpp calls shutdown explicitly)
    std::atexit([]() { screfShutdown(); });
  }
  ScRefWorld* w = new ScRefWorld();
  w->streaming = new sc::WorldStreamingState();
  w->culling.frame = &w->world.renderFrame();
  w->renderPrep.frame = &w->world.renderFrame();
  w->renderPrep.culling = &w->culling;
  w->renderPrep.streaming = w->streaming;
  w->renderPrep.assets = nullptr;
  w->camera.frame = &w->world.renderFrame();
  w->streaming->budgets.maxDrawsBudget = 0u;
  return w;
}

void screfDestroy(ScRefWorld* w)
{
  if (!w) return;
  if (w->streaming)
  {
    w->streaming->partition.shutdownStreaming();
    delete w->streaming;
  }
  delete w;
}

uint32_t screfJobWorkers(void) { return g_workers; }

void screfShutdown(void)
{
  if (g_jobsInit)
  {
    sc::jobs().shutdown();
    g_jobsInit = false;
  }
}

void screfCreateEntities(ScRefWorld* w, uint32_t n, uint32_t* outEntity)
{
  for (uint32_t i = 0; i < n; ++i)
    outEntity[i] = w->world.create().value;
}

void screfAddInstances(ScRefWorld* w, uint32_t n, const uint32_t* entity, const uint32_t* parent,
                       const float* trs9, const float* aabb6, const uint32_t* meshMat2, const uint32_t* flags)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    const sc::Entity e = ent(entity[i]);
    sc::Transform& t = w->world.add<sc::Transform>(e);
    sc::setLocal(t, trs9 + i * 9, trs9 + i * 9 + 3, trs9 + i * 9 + 6);
    if (parent && parent[i] != SCREF_NO_PARENT)
      sc::setParent(t, ent(parent[i]));
    const uint32_t f = flags ? flags[i] : (SCREF_HAS_BOUNDS | SCREF_HAS_MESH);
    if (f & SCREF_HAS_MESH)
    {
      sc::RenderMesh& rm = w->world.add<sc::RenderMesh>(e);
      rm.meshId = meshMat2 ? meshMat2[i * 2 + 0] : 0u;
      rm.materialId = meshMat2 ? meshMat2[i * 2 + 1] : 0u;
    }
    if ((f & SCREF_HAS_BOUNDS) && aabb6)
    {
      sc::Bounds& b = w->world.add<sc::Bounds>(e);
      b.localAabb.min = { aabb6[i * 6 + 0], aabb6[i * 6 + 1], aabb6[i * 6 + 2] };
      b.localAabb.max = { aabb6[i * 6 + 3], aabb6[i * 6 + 4], aabb6[i * 6 + 5] };
    }
  }
}

void screfDestroyEntities(ScRefWorld* w, uint32_t n, const uint32_t* entity)
{
  for (uint32_t i = 0; i < n; ++i)
    w->world.destroy(ent(entity[i]));
}

void screfSetLocal(ScRefWorld* w, uint32_t n, const uint32_t* entity, const float* trs9)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    sc::Transform* t = w->world.get<sc::Transform>(ent(entity[i]));
    if (t) sc::setLocal(*t, trs9 + i * 9, trs9 + i * 9 + 3, trs9 + i * 9 + 6);
  }
}

void screfSetParent(ScRefWorld* w, uint32_t n, const uint32_t* entity, const uint32_t* parent)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    sc::Transform* t = w->world.get<sc::Transform>(ent(entity[i]));
    if (t) sc::setParent(*t, ent(parent[i]));
  }
}

void screfMarkDirty(ScRefWorld* w, uint32_t n, const uint32_t* entity)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    sc::Transform* t = w->world.get<sc::Transform>(ent(entity[i]));
    if (t) sc::markDirty(*t);
  }
}

void screfRunTransform(ScRefWorld* w) { sc::TransformSystem(w->world, 0.0f, nullptr); }

void screfSetViewProj(ScRefWorld* w, const float* vp)
{
  std::memcpy(w->world.renderFrame().viewProj.m, vp, sizeof(float) * 16);
}

void screfSetFreezeCulling(ScRefWorld* w, int freeze) { w->culling.freezeCulling = freeze != 0; }

void screfRunCulling(ScRefWorld* w)
{
  sc::jobs().beginFrame();
  sc::CullingSystem(w->world, 0.0f, &w->culling);
  sc::jobs().publishFrameTelemetry();
}

void screfRunRenderPrep(ScRefWorld* w, uint32_t maxDraws)
{
  w->streaming->budgets.maxDrawsBudget = maxDraws;
  sc::RenderPrepStreamingSystem(w->world, 0.0f, &w->renderPrep);
}

uint32_t screfAddCamera(ScRefWorld* w, const float* trs9, float fovY, float nearZ, float farZ, float aspect)
{
  const sc::Entity e = w->world.create();
  sc::Transform& t = w->world.add<sc::Transform>(e);
  sc::setLocal(t, trs9, trs9 + 3, trs9 + 6);
  sc::Camera& c = w->world.add<sc::Camera>(e);
  c.fovY = fovY; c.nearZ = nearZ; c.farZ = farZ; c.aspect = aspect; c.active = true;
  return e.value;
}

void screfRunCamera(ScRefWorld* w, float aspect)
{
  w->camera.aspect = aspect;
  sc::CameraSystem(w->world, 0.0f, &w->camera);
}

uint32_t screfTransformCount(ScRefWorld* w) { return w->world.componentCount<sc::Transform>(); }

uint32_t screfDenseEntities(ScRefWorld* w, uint32_t cap, uint32_t* out)
{
  uint32_t n = 0;
  w->world.ForEach<sc::Transform>([&](sc::Entity e, sc::Transform&)
  {
    if (n < cap) out[n] = e.value;
    ++n;
  });
  return n;
}

void screfReadWorld(ScRefWorld* w, uint32_t n, const uint32_t* entity, float* out16)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    const sc::Transform* t = w->world.get<sc::Transform>(ent(entity[i]));
    if (t) std::memcpy(out16 + i * 16, t->worldMatrix.m, 64);
    else std::memset(out16 + i * 16, 0, 64);
  }
}

void screfReadTransform(ScRefWorld* w, uint32_t n, const uint32_t* entity, uint32_t* outParent, float* outTrs9,
                        uint8_t* outDirty)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    const sc::Transform* t = w->world.get<sc::Transform>(ent(entity[i]));
    if (!t) continue;
    if (outParent) outParent[i] = t->parent.value;
    if (outTrs9)
    {
      std::memcpy(outTrs9 + i * 9, t->localPos, 12);
      std::memcpy(outTrs9 + i * 9 + 3, t->localRot, 12);
      std::memcpy(outTrs9 + i * 9 + 6, t->localScale, 12);
    }
    if (outDirty) outDirty[i] = t->dirty ? 1 : 0;
  }
}

void screfReadComponents(ScRefWorld* w, uint32_t n, const uint32_t* entity, uint32_t* outFlags, float* outAabb6,
                         uint32_t* outMeshMat2)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    const sc::Entity e = ent(entity[i]);
    uint32_t f = 0;
    if (const sc::Bounds* b = w->world.get<sc::Bounds>(e))
    {
      f |= SCREF_HAS_BOUNDS;
      outAabb6[i * 6 + 0] = b->localAabb.min.x; outAabb6[i * 6 + 1] = b->localAabb.min.y; outAabb6[i * 6 + 2] = b->localAabb.min.z;
      outAabb6[i * 6 + 3] = b->localAabb.max.x; outAabb6[i * 6 + 4] = b->localAabb.max.y; outAabb6[i * 6 + 5] = b->localAabb.max.z;
    }
    if (const sc::RenderMesh* rm = w->world.get<sc::RenderMesh>(e))
    {
      f |= SCREF_HAS_MESH;
      outMeshMat2[i * 2 + 0] = rm->meshId; outMeshMat2[i * 2 + 1] = rm->materialId;
    }
    outFlags[i] = f;
  }
}

void screfGetViewProj(ScRefWorld* w, float* out16) { std::memcpy(out16, w->world.renderFrame().viewProj.m, 64); }

void screfGetPlanes(ScRefWorld* w, float* out24)
{
  for (int p = 0; p < 6; ++p)
  {
    out24[p * 4 + 0] = w->culling.frustum.planes[p].n[0];
    out24[p * 4 + 1] = w->culling.frustum.planes[p].n[1];
    out24[p * 4 + 2] = w->culling.frustum.planes[p].n[2];
    out24[p * 4 + 3] = w->culling.frustum.planes[p].d;
  }
}

void screfGetCullStats(ScRefWorld* w, uint32_t* total, uint32_t* visible, uint32_t* culled)
{
  *total = w->culling.stats.renderablesTotal;
  *visible = w->culling.stats.visible;
  *culled = w->culling.stats.culled;
}

static uint32_t copyEntities(const std::vector<sc::Entity>& v, uint32_t cap, uint32_t* out)
{
  const uint32_t n = (uint32_t)v.size();
  for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = v[i].value;
  return n;
}

uint32_t screfReadVisible(ScRefWorld* w, uint32_t cap, uint32_t* out) { return copyEntities(w->culling.visible, cap, out); }
uint32_t screfReadCulled(ScRefWorld* w, uint32_t cap, uint32_t* out) { return copyEntities(w->culling.culled, cap, out); }
uint32_t screfReadCandidates(ScRefWorld* w, uint32_t cap, uint32_t* out) { return copyEntities(w->culling.candidates, cap, out); }

void screfGetRenderPrepStats(ScRefWorld* w, uint32_t* emitted, uint32_t* dropped)
{
  *emitted = w->renderPrep.stats.drawsEmitted;
  *dropped = w->renderPrep.stats.drawsDroppedByBudget;
}

uint32_t screfReadDraws(ScRefWorld* w, uint32_t cap, void* out80)
{
  const std::vector<sc::DrawItem>& d = w->world.renderFrame().draws;
  const uint32_t n = (uint32_t)d.size();
  const uint32_t c = n < cap ? n : cap;
  if (c) std::memcpy(out80, d.data(), (size_t)c * sizeof(sc::DrawItem));
  return n;
}

uint32_t screfSizeofTransform(void) { return (uint32_t)sizeof(sc::Transform); }
uint32_t screfSizeofDrawItem(void) { return (uint32_t)sizeof(sc::DrawItem); }

uint32_t screfBuildDefaultScene(ScRefWorld* w, uint32_t frames)
{
  // Same constants as src/sandbox/src/main.cpp:66-99; same RenderPrep chain as main.cpp:256-259.
  w->spawner.spawnCount = 0;
  w->spawner.churnEvery = 0;
  w->spawner.churnCount = 0;

  sc::WorldPartitionConfig cfg{};
  cfg.sectorSizeMeters = 64.0f;
  cfg.seed = 424242u;
  cfg.propsPerSectorMin = 18u;
  cfg.propsPerSectorMax = 34u;
  cfg.includeGroundPlane = true;
  w->streaming->partition.configure(cfg);
  w->streaming->partition.setAssetManager(nullptr);

  const float sectorSize = w->streaming->partition.config().sectorSizeMeters;
  w->spawner.overrideCamera = true;
  w->spawner.cameraPos[0] = sectorSize * 0.5f;
  w->spawner.cameraPos[1] = 6.0f;
  w->spawner.cameraPos[2] = sectorSize * 0.5f + 12.0f;
  w->spawner.cameraRot[0] = 0.0f;
  w->spawner.cameraRot[1] = 3.14159265f;
  w->spawner.cameraRot[2] = 0.0f;
  sc::WorldStreamingBudgets& b = w->streaming->budgets;
  b.loadRadiusSectors = 2u;
  b.unloadRadiusSectors = 3u;
  b.maxActiveSectors = 25u;
  b.maxEntitiesBudget = 5000u;
  b.maxDrawsBudget = 6000u;
  b.maxConcurrentLoads = 4u;
  b.maxActivationsPerFrame = 2u;
  b.maxDespawnsPerFrame = 128u;

  w->camera.aspect = 1280.0f / 720.0f;

  sc::Scheduler scheduler;
  scheduler.addSystem("Spawner", sc::SystemPhase::Simulation, sc::SpawnerSystem, &w->spawner);
  scheduler.addSystem("WorldStreaming", sc::SystemPhase::Simulation, sc::WorldStreamingSystem, w->streaming, { "Spawner" });
  scheduler.addSystem("Transform", sc::SystemPhase::RenderPrep, sc::TransformSystem, nullptr);
  scheduler.addSystem("Camera", sc::SystemPhase::RenderPrep, sc::CameraSystem, &w->camera, { "Transform" });
  scheduler.addSystem("Culling", sc::SystemPhase::RenderPrep, sc::CullingSystem, &w->culling, { "Camera" });
  scheduler.addSystem("RenderPrep", sc::SystemPhase::RenderPrep, sc::RenderPrepStreamingSystem, &w->renderPrep, { "Culling" });
  scheduler.finalize();

  for (uint32_t f = 0; f < frames; ++f)
  {
    sc::jobs().beginFrame();
    scheduler.tick(w->world, 1.0f / 60.0f, 0, 1.0f / 60.0f);
    sc::jobs().publishFrameTelemetry();
    // async sector loads complete on worker threads; give them a moment like a real frame would
    std::this_thread::sleep_for(std::chrono::milliseconds(2));
  }
  return w->streaming->partition.loadedSectorCount();
}

void screfTimeFrame(ScRefWorld* w, uint32_t iters, uint32_t nViews, const float* vps,
                    uint32_t nDirty, const uint32_t* dirtyEntity, uint32_t maxDraws,
                    double* outTransformS, double* outCullS, double* outPrepS)
{
  double tT = 0.0, tC = 0.0, tP = 0.0;
  w->streaming->budgets.maxDrawsBudget = maxDraws;
  for (uint32_t it = 0; it < iters; ++it)
  {
    screfMarkDirty(w, nDirty, dirtyEntity);
    sc::jobs().beginFrame();
    const sc::Tick t0 = sc::nowTicks();
    sc::TransformSystem(w->world, 0.0f, nullptr);
    const sc::Tick t1 = sc::nowTicks();
    for (uint32_t v = 0; v < nViews; ++v)
    {
      std::memcpy(w->world.renderFrame().viewProj.m, vps + v * 16, 64);
      sc::CullingSystem(w->world, 0.0f, &w->culling);
    }
    const sc::Tick t2 = sc::nowTicks();
    sc::RenderPrepStreamingSystem(w->world, 0.0f, &w->renderPrep);
    const sc::Tick t3 = sc::nowTicks();
    sc::jobs().publishFrameTelemetry();
    tT += sc::ticksToSeconds(t1 - t0);
    tC += sc::ticksToSeconds(t2 - t1);
    tP += sc::ticksToSeconds(t3 - t2);
  }
  *outTransformS = tT; *outCullS = tC; *outPrepS = tP;
}

void screfMat4Trs(const float* p, const float* r, const float* s, float* out16)
{
  const sc::Mat4 m = sc::mat4_trs(p, r, s);
  std::memcpy(out16, m.m, 64);
}

void screfMat4Mul(const float* a16, const float* b16, float* out16)
{
  sc::Mat4 a{}, b{};
  std::memcpy(a.m, a16, 64); std::memcpy(b.m, b16, 64);
  const sc::Mat4 r = sc::mat4_mul(a, b);
  std::memcpy(out16, r.m, 64);
}

void screfMat4Inverse(const float* a16, float* out16)
{
  sc::Mat4 a{};
  std::memcpy(a.m, a16, 64);
  const sc::Mat4 r = sc::mat4_inverse(a);
  std::memcpy(out16, r.m, 64);
}

void screfMat4Perspective(float fovYRad, float aspect, float zn, float zf, int flipY, float* out16)
{
  const sc::Mat4 r = sc::mat4_perspective_rh_zo(fovYRad, aspect, zn, zf, flipY != 0);
  std::memcpy(out16, r.m, 64);
}

void screfFrustumFromViewProj(const float* vp16, float* out24)
{
  sc::Mat4 vp{};
  std::memcpy(vp.m, vp16, 64);
  const sc::Frustum f = sc::frustumFromViewProj(vp);
  for (int p = 0; p < 6; ++p)
  {
    out24[p * 4 + 0] = f.planes[p].n[0];
    out24[p * 4 + 1] = f.planes[p].n[1];
    out24[p * 4 + 2] = f.planes[p].n[2];
    out24[p * 4 + 3] = f.planes[p].d;
  }
}

int screfSphereInFrustum(const float* planes24, const float* center3, float radius)
{
  sc::Frustum f{};
  for (int p = 0; p < 6; ++p)
  {
    f.planes[p].n[0] = planes24[p * 4 + 0];
    f.planes[p].n[1] = planes24[p * 4 + 1];
    f.planes[p].n[2] = planes24[p * 4 + 2];
    f.planes[p].d = planes24[p * 4 + 3];
  }
  f.valid = true;
  return sc::sphereInFrustum(f, center3, radius) ? 1 : 0;
}

void screfWorldBoundsSphere(const float* world16, const float* aabb6, float* outCenter3, float* outRadius)
{
  sc::Transform t{};
  std::memcpy(t.worldMatrix.m, world16, 64);
  sc::Bounds b{};
  b.localAabb.min = { aabb6[0], aabb6[1], aabb6[2] };
  b.localAabb.max = { aabb6[3], aabb6[4], aabb6[5] };
  sc::computeWorldBoundsSphere(t, b, outCenter3, *outRadius);
}

// The exact libm entry points sc_math.cpp:102-107 resolves to (std::sin/std::cos on float).
float screfSinf(float x) { return std::sin(x); }
float screfCosf(float x) { return std::cos(x); }

void screfSinCosSweep(uint32_t first, uint64_t count, uint32_t stride, uint64_t* outSinHash, uint64_t* outCosHash)
{
  uint64_t hs = 0, hc = 0;
  uint32_t bits = first;
  for (uint64_t i = 0; i < count; ++i, bits += stride)
  {
    float x; std::memcpy(&x, &bits, 4);
    const float s = std::sin(x), c = std::cos(x);
    uint32_t sb, cb;
    std::memcpy(&sb, &s, 4); std::memcpy(&cb, &c, 4);
    if (s != s) sb = 0x7fc00000u;  // canonical NaN
    if (c != c) cb = 0x7fc00000u;
    hs = (hs ^ sb) * 0x100000001b3ull + bits;
    hc = (hc ^ cb) * 0x100000001b3ull + bits;
  }
  *outSinHash = hs; *outCosHash = hc;
}

// ---- .scsector files through the reference's own writer / reader (tools/shared/world_format.cpp:76-334) ----------
int screfWriteSectorFile(const char* path, uint32_t version, int32_t x, int32_t z, uint32_t n, const uint64_t* id,
                         const uint64_t* modelId, const uint64_t* meshId, const uint64_t* materialId, const float* trs9,
                         const uint32_t* tags, uint32_t nExtraChunkBytes)
{
  sc_world::SectorFile f{};
  f.version = version;
  f.sector = { x, z };
  f.instances.resize(n);
  for (uint32_t i = 0; i < n; ++i)
  {
    sc_world::Instance& in = f.instances[i];
    in.id = id[i]; in.model_id = modelId[i]; in.mesh_id = meshId[i]; in.material_id = materialId[i];
    std::memcpy(in.transform.position, trs9 + 9 * i, 12);
    std::memcpy(in.transform.rotation, trs9 + 9 * i + 3, 12);
    std::memcpy(in.transform.scale, trs9 + 9 * i + 6, 12);
    std::snprintf(in.name, sc_world::kInstanceNameMax, "Inst_%u", i);
    in.tags = tags[i];
    in.albedo_texture_id = 77 + i; in.material_flags = i & 1u;
  }
  // other chunk kinds before/after do not matter to the INST reader; add a lane and a spawner so that the walk is tested
  if (nExtraChunkBytes)
  {
    sc_world::Lane lane{}; lane.id = 5; lane.flags = 1; lane.points = { {0, 0, 0}, {1, 0, 2}, {3, 0, 4} };
    f.lanes.push_back(lane);
    sc_world::Spawner sp{}; sp.id = 9; sp.type = 2; sp.rate = 0.5f;
    f.spawners.push_back(sp);
  }
  return sc_world::WriteSectorFile(path, f) ? 1 : 0;
}

// returns the instance count (or -1); fills up to cap entries
int screfReadSectorInstances(const char* path, uint32_t cap, int32_t* outXZ, uint64_t* outId, uint64_t* outMeshId,
                             uint64_t* outMaterialId, float* outTrs9)
{
  sc_world::SectorFile f{};
  if (!sc_world::ReadSectorFile(path, &f)) return -1;
  outXZ[0] = f.sector.x; outXZ[1] = f.sector.z;
  for (uint32_t i = 0; i < f.instances.size() && i < cap; ++i)
  {
    const sc_world::Instance& in = f.instances[i];
    outId[i] = in.id; outMeshId[i] = in.mesh_id; outMaterialId[i] = in.material_id;
    std::memcpy(outTrs9 + 9 * i, in.transform.position, 12);
    std::memcpy(outTrs9 + 9 * i + 3, in.transform.rotation, 12);
    std::memcpy(outTrs9 + 9 * i + 6, in.transform.scale, 12);
  }
  return (int)f.instances.size();
}

uint64_t screfHashAssetPath(const char* path) { return sc_world::HashAssetPath(path); }

}  // extern "C"
