// sc_gpu_systems.cpp — adapter systems: the engine's World stays the source of truth, the GPU holds an SoA mirror.
// See sc_gpu_systems.h. Reference behaviour being mirrored (relative to /root/reference):
//   TransformSystem            src/core/src/sc_ecs.cpp:118-211
//   CullingSystem              src/engine/world/sc_world_partition.cpp:1199-1284
//   RenderPrepStreamingSystem  src/engine/world/sc_world_partition.cpp:1286-1359
#include "sc_gpu_systems.h"

#include "sc_assets.h"
#include "sc_log.h"
#include "sc_math.h"

#include <cstdio>
#include <cstring>
#include <unordered_set>

namespace sc::gpu
{
  static_assert(sizeof(Entity) == sizeof(uint32_t), "Entity must be a 32-bit handle (sc_ecs.h:14-37)");
  static_assert(sizeof(DrawItem) == sizeof(ScGpuDrawItem) && sizeof(DrawItem) == 80, "DrawItem layout (sc_ecs.h:159-165)");

  namespace
  {
    void noteError(GpuSceneState& s, const char* where)
    {
      std::snprintf(s.lastError, sizeof(s.lastError), "%s: %s", where, scgpuLastError(s.ctx));
      sc::log(sc::LogLevel::Error, "scgpu %s", s.lastError);
    }

    struct SpawnBatch
    {
      std::vector<uint32_t> entity, parent, meshMat, flags;
      std::vector<float> trs, aabb;
      void push(World& world, Entity e, Transform& t)
      {
        entity.push_back(e.value);
        parent.push_back(t.parent.value);
        trs.insert(trs.end(), t.localPos, t.localPos + 3);
        trs.insert(trs.end(), t.localRot, t.localRot + 3);
        trs.insert(trs.end(), t.localScale, t.localScale + 3);
        uint32_t f = 0;
        float bb[6] = { -0.5f, -0.5f, -0.5f, 0.5f, 0.5f, 0.5f };
        if (const Bounds* b = world.get<Bounds>(e))
        {
          f |= SCGPU_HAS_BOUNDS;
          bb[0] = b->localAabb.min.x; bb[1] = b->localAabb.min.y; bb[2] = b->localAabb.min.z;
          bb[3] = b->localAabb.max.x; bb[4] = b->localAabb.max.y; bb[5] = b->localAabb.max.z;
        }
        aabb.insert(aabb.end(), bb, bb + 6);
        uint32_t mm[2] = { 0u, 0u };
        if (const RenderMesh* rm = world.get<RenderMesh>(e))
        {
          f |= SCGPU_HAS_MESH;
          mm[0] = rm->meshId; mm[1] = rm->materialId;
        }
        meshMat.insert(meshMat.end(), mm, mm + 2);
        flags.push_back(f);
      }
      bool submit(GpuSceneState& s)
      {
        if (entity.empty()) return true;
        if (!scgpuSpawn(s.ctx, (uint32_t)entity.size(), entity.data(), parent.data(), trs.data(), aabb.data(),
                        meshMat.data(), flags.data()))
        {
          noteError(s, "scgpuSpawn");
          return false;
        }
        return true;
      }
    };

    // TransformSystem's per-entity fix-ups (sc_ecs.cpp:143-164), applied to the host component so that the
    // engine observes the same Transform fields as with the CPU system
    void fixUp(World& world, Entity e, Transform& t)
    {
      if (t.localScale[0] == 0.0f && t.localScale[1] == 0.0f && t.localScale[2] == 0.0f)
      {
        t.localScale[0] = t.localScale[1] = t.localScale[2] = 1.0f;
        t.dirty = true;
      }
      if (isValidEntity(t.parent))
      {
        const bool valid = t.parent != e && world.isAlive(t.parent) && world.has<Transform>(t.parent);
        if (!valid)
        {
          t.dirty = true;
          t.parent = kInvalidEntity;
        }
      }
    }
  }

  bool init(GpuSceneState& s)
  {
    ScGpuSceneDesc d{};
    d.struct_size = sizeof(d);
    d.device = s.device;
    d.max_instances = s.maxInstances;
    d.max_entity_index = 0;
    d.max_views = s.maxViews;
    s.ctx = scgpuCreate(&d);
    if (!s.ctx)
    {
      std::snprintf(s.lastError, sizeof(s.lastError), "scgpuCreate: %s", scgpuLastError(nullptr));
      sc::log(sc::LogLevel::Error, "%s", s.lastError);
      return false;
    }
    return true;
  }

  void shutdown(GpuSceneState& s)
  {
    scgpuDestroy(s.ctx);
    s.ctx = nullptr;
    s.dense.clear();
    s.parentOf.clear();
  }

  void TransformSystem(World& world, float dt, void* user)
  {
    (void)dt;
    GpuSceneState* s = static_cast<GpuSceneState*>(user);
    if (!s || !s->ctx)
      return;
    s->transformPassDone = false;

    // ---- 1. pool membership: what appeared / disappeared since the last frame ----
    std::vector<Entity> cur;
    cur.reserve(s->dense.size() + 64);
    world.ForEach<Transform>([&](Entity e, Transform&) { cur.push_back(e); });

    std::unordered_set<uint32_t> curSet;
    curSet.reserve(cur.size() * 2);
    for (const Entity e : cur) curSet.insert(e.value);

    std::vector<uint32_t> removed;
    for (const Entity e : s->dense)
      if (!curSet.count(e.value)) removed.push_back(e.value);
    if (!removed.empty())
    {
      if (!scgpuDespawn(s->ctx, (uint32_t)removed.size(), removed.data())) noteError(*s, "scgpuDespawn");
      // replay ComponentPool::remove (sc_ecs.h:240-262) on the shadow to predict the GPU's pool order
      std::unordered_map<uint32_t, uint32_t> slotOf;
      slotOf.reserve(s->dense.size() * 2);
      for (uint32_t i = 0; i < s->dense.size(); ++i) slotOf[s->dense[i].value] = i;
      for (const uint32_t h : removed)
      {
        const uint32_t slot = slotOf[h], last = (uint32_t)s->dense.size() - 1u;
        if (slot != last)
        {
          s->dense[slot] = s->dense[last];
          slotOf[s->dense[slot].value] = slot;
        }
        s->dense.pop_back();
        slotOf.erase(h);
        s->parentOf.erase(h);
      }
    }

    std::unordered_set<uint32_t> known;
    known.reserve(s->dense.size() * 2);
    for (const Entity e : s->dense) known.insert(e.value);

    SpawnBatch spawn;
    for (const Entity e : cur)
    {
      if (known.count(e.value)) continue;
      Transform& t = *world.get<Transform>(e);
      fixUp(world, e, t);
      spawn.push(world, e, t);
      s->dense.push_back(e);
      s->parentOf[e.value] = t.parent.value;
      if (!s->leaveDirtyFlags) t.dirty = false;
    }
    spawn.submit(*s);

    // The engine destroyed entities in an order we cannot see; if replaying them in pool order did not end in
    // the engine's pool order, rebuild the mirror (rare: orders agree when a batch is destroyed oldest-first).
    bool same = s->dense.size() == cur.size();
    for (size_t i = 0; same && i < cur.size(); ++i) same = s->dense[i] == cur[i];
    const size_t firstNew = s->dense.size() - spawn.entity.size();
    if (!same)
    {
      ++s->resyncs;
      std::vector<uint32_t> all;
      for (const Entity e : s->dense) all.push_back(e.value);
      if (!all.empty() && !scgpuDespawn(s->ctx, (uint32_t)all.size(), all.data())) noteError(*s, "scgpuDespawn(resync)");
      s->dense = cur;
      s->parentOf.clear();
      SpawnBatch rebuild;
      for (const Entity e : cur)
      {
        Transform& t = *world.get<Transform>(e);
        fixUp(world, e, t);
        rebuild.push(world, e, t);
        s->parentOf[e.value] = t.parent.value;
        if (!s->leaveDirtyFlags) t.dirty = false;
      }
      rebuild.submit(*s);
      // world matrices of clean instances are lost by a rebuild: recompute everything once
      scgpuMarkAllDirty(s->ctx);
    }
    else
    {
      // ---- 2. component edits: the engine writes Transform fields and sets dirty (sc_ecs.h:73-96) ----
      std::vector<uint32_t> dirtyE, reparentE, reparentP;
      std::vector<float> dirtyTrs;
      for (size_t i = 0; i < firstNew; ++i)
      {
        const Entity e = s->dense[i];
        Transform& t = *world.get<Transform>(e);
        fixUp(world, e, t);
        uint32_t& shadowParent = s->parentOf[e.value];
        if (shadowParent != t.parent.value)
        {
          reparentE.push_back(e.value);
          reparentP.push_back(t.parent.value);
          shadowParent = t.parent.value;
        }
        if (t.dirty)
        {
          dirtyE.push_back(e.value);
          dirtyTrs.insert(dirtyTrs.end(), t.localPos, t.localPos + 3);
          dirtyTrs.insert(dirtyTrs.end(), t.localRot, t.localRot + 3);
          dirtyTrs.insert(dirtyTrs.end(), t.localScale, t.localScale + 3);
          if (!s->leaveDirtyFlags) t.dirty = false;
        }
      }
      if (!reparentE.empty() && !scgpuSetParent(s->ctx, (uint32_t)reparentE.size(), reparentE.data(), reparentP.data()))
        noteError(*s, "scgpuSetParent");
      if (!dirtyE.empty() && !scgpuSetLocal(s->ctx, (uint32_t)dirtyE.size(), dirtyE.data(), dirtyTrs.data()))
        noteError(*s, "scgpuSetLocal");
    }

    // ---- 3. matrices the host itself consumes before culling: cameras (CameraSystem, sc_ecs.cpp:268) ----
    std::vector<uint32_t> needWorld;
    world.ForEach<Camera, Transform>([&](Entity e, Camera&, Transform& t)
    {
      if (!isValidEntity(t.parent))
        t.worldMatrix = mat4_trs(t.localPos, t.localRot, t.localScale);  // a root: world == local, O(1) on the host
      else
        needWorld.push_back(e.value);
    });
    if (s->readBackAllWorldMatrices)
    {
      needWorld.clear();
      for (const Entity e : s->dense) needWorld.push_back(e.value);
    }
    if (!needWorld.empty())
    {
      // parented camera (or full read-back requested): run the transform pass now, cull again once the view is known
      const float identity[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
      scgpuSetViews(s->ctx, 1, identity);
      if (!scgpuUpdate(s->ctx, 0)) noteError(*s, "scgpuUpdate(transform)");
      s->transformPassDone = true;
      std::vector<float> m(needWorld.size() * 16);
      if (!scgpuReadWorld(s->ctx, (uint32_t)needWorld.size(), needWorld.data(), m.data())) noteError(*s, "scgpuReadWorld");
      for (size_t i = 0; i < needWorld.size(); ++i)
      {
        Entity e{};
        e.value = needWorld[i];
        if (Transform* t = world.get<Transform>(e)) std::memcpy(t->worldMatrix.m, m.data() + i * 16, 64);
      }
    }
  }

  void CullingSystem(World& world, float dt, void* user)
  {
    (void)world; (void)dt;
    GpuCullingState* st = static_cast<GpuCullingState*>(user);
    if (!st || !st->scene || !st->scene->ctx || !st->culling || !st->culling->frame)
      return;
    GpuSceneState& s = *st->scene;
    CullingState& cs = *st->culling;

    if (!scgpuSetViews(s.ctx, 1, cs.frame->viewProj.m)) noteError(s, "scgpuSetViews");
    uint32_t flags = 0;
    if (cs.freezeCulling) flags |= SCGPU_UPDATE_FREEZE_CULLING;
    if (s.transformPassDone) flags |= SCGPU_UPDATE_SKIP_TRANSFORM;
    if (st->fillCulledList) flags |= SCGPU_UPDATE_CULLED_LISTS;
    if (!scgpuUpdate(s.ctx, flags)) noteError(s, "scgpuUpdate");
    s.transformPassDone = false;

    ScGpuCounts c{};
    if (!scgpuGetCounts(s.ctx, &c)) noteError(s, "scgpuGetCounts");
    cs.stats.renderablesTotal = c.renderablesTotal;
    cs.stats.visible = c.visible[0];
    cs.stats.culled = c.culled[0];

    cs.candidates.clear();
    cs.visible.resize(c.visible[0]);
    cs.culled.clear();
    uint32_t n = 0;
    if (c.visible[0] && !scgpuReadVisible(s.ctx, 0, reinterpret_cast<uint32_t*>(cs.visible.data()), c.visible[0], &n))
      noteError(s, "scgpuReadVisible");
    if (st->fillCulledList && c.culled[0])
    {
      cs.culled.resize(c.culled[0]);
      if (!scgpuReadCulled(s.ctx, 0, reinterpret_cast<uint32_t*>(cs.culled.data()), c.culled[0], &n))
        noteError(s, "scgpuReadCulled");
    }
    if (st->fillCandidates)
    {
      for (const Entity e : s.dense)
        if (world.has<RenderMesh>(e)) cs.candidates.push_back(e);
    }
    if (c.renderablesTotal != 0 && !cs.freezeCulling)
    {
      float planes[24];
      if (scgpuGetViewPlanes(s.ctx, 0, planes))
      {
        for (int p = 0; p < 6; ++p)
        {
          cs.frustum.planes[p].n[0] = planes[p * 4 + 0];
          cs.frustum.planes[p].n[1] = planes[p * 4 + 1];
          cs.frustum.planes[p].n[2] = planes[p * 4 + 2];
          cs.frustum.planes[p].d = planes[p * 4 + 3];
        }
        cs.frustum.valid = true;
      }
    }
  }

  void RenderPrepStreamingSystem(World& world, float dt, void* user)
  {
    (void)world; (void)dt;
    GpuRenderPrepState* st = static_cast<GpuRenderPrepState*>(user);
    if (!st || !st->scene || !st->scene->ctx || !st->prep || !st->prep->frame)
      return;
    RenderPrepStreamingState& rp = *st->prep;
    RenderFrameData& frame = *rp.frame;
    frame.clear();

    const uint32_t maxDraws = rp.streaming ? rp.streaming->budgets.maxDrawsBudget : 0u;
    if (rp.assets && rp.streaming)
    {
      rp.assets->beginFrame(rp.streaming->frameIndex);
      rp.assets->setFreezeEviction(rp.streaming->freezeEviction);
    }

    uint32_t emitted = 0, dropped = 0;
    // sizes first, then straight into RenderFrameData::draws (same 80-byte records)
    if (!scgpuReadDrawItems(st->scene->ctx, 0, maxDraws, nullptr, 0, &emitted, &dropped)) noteError(*st->scene, "scgpuReadDrawItems");
    frame.draws.resize(emitted);
    if (emitted &&
        !scgpuReadDrawItems(st->scene->ctx, 0, maxDraws, reinterpret_cast<ScGpuDrawItem*>(frame.draws.data()), emitted, &emitted, &dropped))
      noteError(*st->scene, "scgpuReadDrawItems");

    if (rp.assets)
    {
      for (const DrawItem& d : frame.draws)
      {
        rp.assets->touchMaterial(d.materialId);
        rp.assets->touchMesh(d.meshId);
      }
      const uint32_t loadLimit = rp.assets->residencyConfig().maxTextureLoadsPerFrame;
      rp.assets->pumpTextureLoads(loadLimit);
      if (!rp.assets->residencyConfig().freezeEviction)
        rp.assets->evictIfNeeded();
    }
    rp.stats.drawsEmitted = emitted;
    rp.stats.drawsDroppedByBudget = dropped;
  }
}
