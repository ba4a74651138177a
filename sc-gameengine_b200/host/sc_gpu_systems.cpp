// sc_gpu_systems.cpp — adapter systems: the engine's World stays the source of truth, the GPU holds an SoA mirror.
// See sc_gpu_systems.h. Reference behaviour being mirrored (relative to /root/reference):
//   TransformSystem            src/core/src/sc_ecs.cpp:118-211
//   CullingSystem              src/engine/world/sc_world_partition.cpp:1199-1284
//   RenderPrepStreamingSystem  src/engine/world/sc_world_partition.cpp:1286-1359
#include "sc_gpu_systems.h"

#include "sc_assets.h"
#include "sc_log.h"
#include "sc_math.h"

#include <chrono>
#include <cstdio>
#include <cstring>

namespace sc::gpu
{
  static_assert(sizeof(Entity) == sizeof(uint32_t), "Entity must be a 32-bit handle (sc_ecs.h:14-37)");
  static_assert(sizeof(DrawItem) == sizeof(ScGpuDrawItem) && sizeof(DrawItem) == 80, "DrawItem layout (sc_ecs.h:159-165)");

  namespace
  {
    constexpr uint32_t kNoEntity = 0xFFFFFFFFu;

    void noteError(GpuSceneState& s, const char* where)
    {
      std::snprintf(s.lastError, sizeof(s.lastError), "%s: %s", where, scgpuLastError(s.ctx));
      sc::log(sc::LogLevel::Error, "scgpu %s", s.lastError);
    }

    EntityShadow& shadowOf(GpuSceneState& s, uint32_t handle)
    {
      const uint32_t idx = handle & Entity::INDEX_MASK;
      if (idx >= s.shadow.size()) s.shadow.resize((size_t)idx + 1u + s.shadow.size() / 2u);
      return s.shadow[idx];
    }

    struct SpawnBatch
    {
      std::vector<uint32_t> entity, parent, meshMat, flags;
      std::vector<float> trs, aabb;
      // reads the components the entity owns right now and records them in the shadow
      void push(GpuSceneState& s, World& world, Entity e, Transform& t, uint32_t pos)
      {
        EntityShadow& sh = shadowOf(s, e.value);
        sh = EntityShadow{};
        sh.handle = e.value;
        sh.parent = t.parent.value;
        sh.pos = pos;
        sh.seen = s.stamp;
        entity.push_back(e.value);
        parent.push_back(t.parent.value);
        trs.insert(trs.end(), t.localPos, t.localPos + 3);
        trs.insert(trs.end(), t.localRot, t.localRot + 3);
        trs.insert(trs.end(), t.localScale, t.localScale + 3);
        if (const Bounds* b = world.get<Bounds>(e))
        {
          sh.flags |= SCGPU_HAS_BOUNDS;
          sh.aabb[0] = b->localAabb.min.x; sh.aabb[1] = b->localAabb.min.y; sh.aabb[2] = b->localAabb.min.z;
          sh.aabb[3] = b->localAabb.max.x; sh.aabb[4] = b->localAabb.max.y; sh.aabb[5] = b->localAabb.max.z;
        }
        aabb.insert(aabb.end(), sh.aabb, sh.aabb + 6);
        if (const RenderMesh* rm = world.get<RenderMesh>(e))
        {
          sh.flags |= SCGPU_HAS_MESH;
          sh.meshId = rm->meshId; sh.materialId = rm->materialId;
        }
        meshMat.push_back(sh.meshId);
        meshMat.push_back(sh.materialId);
        flags.push_back(sh.flags);
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

    // ComponentPool::remove (sc_ecs.h:240-262) on the shadow of the pool
    void shadowRemove(GpuSceneState& s, uint32_t handle)
    {
      EntityShadow& sh = shadowOf(s, handle);
      const uint32_t pos = sh.pos, last = (uint32_t)s.dense.size() - 1u;
      if (pos != last)
      {
        const uint32_t moved = s.dense[last];
        s.dense[pos] = moved;
        shadowOf(s, moved).pos = pos;
      }
      s.dense.pop_back();
      sh = EntityShadow{};
    }

    // RenderMesh / Bounds added to, removed from or edited on an entity whose Transform is already on the GPU: compared
    // with what was last uploaded, queued for this frame's scgpuSetRender batch when different
    void checkRenderComponents(GpuSceneState& s, World& world, Entity e, EntityShadow& sh)
    {
      const RenderMesh* rm = world.get<RenderMesh>(e);
      const Bounds* b = world.get<Bounds>(e);
      uint32_t flags = 0;
      bool changed = false;
      if (rm)
      {
        flags |= SCGPU_HAS_MESH;
        if (sh.meshId != rm->meshId || sh.materialId != rm->materialId) { sh.meshId = rm->meshId; sh.materialId = rm->materialId; changed = true; }
      }
      if (b)
      {
        flags |= SCGPU_HAS_BOUNDS;
        const float bb[6] = { b->localAabb.min.x, b->localAabb.min.y, b->localAabb.min.z, b->localAabb.max.x, b->localAabb.max.y, b->localAabb.max.z };
        if (std::memcmp(bb, sh.aabb, sizeof(bb)) != 0) { std::memcpy(sh.aabb, bb, sizeof(bb)); changed = true; }
      }
      if (flags == sh.flags && !changed) return;
      sh.flags = flags;
      s.renderE.push_back(sh.handle);
      s.renderMM.push_back(sh.meshId); s.renderMM.push_back(sh.materialId);
      s.renderBB.insert(s.renderBB.end(), sh.aabb, sh.aabb + 6);
      s.renderF.push_back(flags);
    }

    // A parent is valid iff it is another entity that is alive and owns a Transform (sc_ecs.cpp:151-164) - i.e. iff the
    // mirror of the Transform pool holds exactly that handle. No World look-up.
    bool parentInMirror(const GpuSceneState& s, uint32_t self, uint32_t parent)
    {
      if (parent == self) return false;
      const uint32_t idx = parent & Entity::INDEX_MASK;
      return idx < s.shadow.size() && s.shadow[idx].handle == parent;
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
    s.shadow.clear();
  }

  // One sequential walk of the Transform pool per frame, like the reference's own TransformSystem (sc_ecs.cpp:118-211
  // walks it twice), against a dense shadow table indexed by Entity::index(): no hash set, no map, no allocation in
  // steady state. What it finds goes to the GPU as delta batches.
  void TransformSystem(World& world, float dt, void* user)
  {
    (void)dt;
    GpuSceneState* s = static_cast<GpuSceneState*>(user);
    if (!s || !s->ctx)
      return;
    const auto t0 = std::chrono::steady_clock::now();
    s->transformPassDone = false;
    if (++s->stamp == 0u) ++s->stamp;

    // ---- 1. the pool as it stands: membership (new / known), component edits of the known ones ----
    s->cur.clear(); s->fresh.clear(); s->dirtyE.clear(); s->dirtyTrs.clear(); s->reparentE.clear(); s->reparentP.clear();
    s->renderE.clear(); s->renderMM.clear(); s->renderF.clear(); s->renderBB.clear();
    uint32_t known = 0;
    world.ForEach<Transform>([&](Entity e, Transform& t)
    {
      const uint32_t i = (uint32_t)s->cur.size();
      s->cur.push_back(e.value);
      EntityShadow& sh = shadowOf(*s, e.value);
      if (sh.handle != e.value) { s->fresh.push_back(i); return; }  // appeared since the last frame
      sh.seen = s->stamp;
      ++known;
      // TransformSystem's fix-ups (sc_ecs.cpp:143-164). The parent is re-validated against the World only when the
      // handle changed; a parent that merely DIED shows up as a removal below and is dealt with there.
      if (t.localScale[0] == 0.0f && t.localScale[1] == 0.0f && t.localScale[2] == 0.0f)
      {
        t.localScale[0] = t.localScale[1] = t.localScale[2] = 1.0f;
        t.dirty = true;
      }
      if (sh.parent != t.parent.value)
      {
        fixUp(world, e, t);
        s->reparentE.push_back(e.value);
        s->reparentP.push_back(t.parent.value);
        sh.parent = t.parent.value;
      }
      if (t.dirty)
      {
        s->dirtyE.push_back(e.value);
        s->dirtyTrs.insert(s->dirtyTrs.end(), t.localPos, t.localPos + 3);
        s->dirtyTrs.insert(s->dirtyTrs.end(), t.localRot, t.localRot + 3);
        s->dirtyTrs.insert(s->dirtyTrs.end(), t.localScale, t.localScale + 3);
        if (!s->leaveDirtyFlags) t.dirty = false;
      }
      if (s->trackRenderComponents) checkRenderComponents(*s, world, e, sh);
    });

    // ---- 2. what disappeared: shadow entries the walk did not meet, in (old) pool order; then the pool shadow
    //         replays ComponentPool::remove for them, which is the order scgpuDespawn produces on the GPU ----
    if (known != s->dense.size())
    {
      s->removed.clear();
      for (const uint32_t h : s->dense)
        if (s->shadow[h & Entity::INDEX_MASK].seen != s->stamp) s->removed.push_back(h);
      if (!scgpuDespawn(s->ctx, (uint32_t)s->removed.size(), s->removed.data())) noteError(*s, "scgpuDespawn");
      for (const uint32_t h : s->removed) shadowRemove(*s, h);
      // children of what just died: the reference detaches them and marks them dirty (sc_ecs.cpp:151-164); the device
      // does the same on its own copy (k_resolve_parents), the host component is patched here
      for (const uint32_t h : s->dense)
      {
        EntityShadow& sh = s->shadow[h & Entity::INDEX_MASK];
        if (sh.parent == kNoEntity || parentInMirror(*s, h, sh.parent)) continue;
        // (a parent spawned THIS frame is not in the mirror yet: ask the World before cutting the link)
        Entity e{};
        e.value = h;
        Transform* t = world.get<Transform>(e);
        if (!t) continue;
        fixUp(world, e, *t);
        sh.parent = t->parent.value;
      }
    }

    // ---- 3. new Transforms, appended in pool order ----
    SpawnBatch spawn;
    for (const uint32_t i : s->fresh)
    {
      Entity e{};
      e.value = s->cur[i];
      Transform& t = *world.get<Transform>(e);
      fixUp(world, e, t);
      spawn.push(*s, world, e, t, (uint32_t)s->dense.size());
      s->dense.push_back(e.value);
      if (!s->leaveDirtyFlags) t.dirty = false;
    }
    spawn.submit(*s);

    // The engine destroyed entities in an order we cannot see; if replaying them in pool order did not end in
    // the engine's pool order, rebuild the mirror (rare: orders agree when a batch is destroyed oldest-first).
    const bool same = s->dense.size() == s->cur.size() &&
                      (s->cur.empty() || std::memcmp(s->dense.data(), s->cur.data(), s->cur.size() * sizeof(uint32_t)) == 0);
    if (!same)
    {
      ++s->resyncs;
      if (!s->dense.empty() && !scgpuDespawn(s->ctx, (uint32_t)s->dense.size(), s->dense.data())) noteError(*s, "scgpuDespawn(resync)");
      for (const uint32_t h : s->dense) s->shadow[h & Entity::INDEX_MASK] = EntityShadow{};
      s->dense.clear();
      SpawnBatch rebuild;
      for (const uint32_t h : s->cur)
      {
        Entity e{};
        e.value = h;
        Transform& t = *world.get<Transform>(e);
        fixUp(world, e, t);
        rebuild.push(*s, world, e, t, (uint32_t)s->dense.size());
        s->dense.push_back(h);
        if (!s->leaveDirtyFlags) t.dirty = false;
      }
      rebuild.submit(*s);
      // world matrices of clean instances are lost by a rebuild: recompute everything once
      scgpuMarkAllDirty(s->ctx);
    }
    else
    {
      // ---- 4. component edits: the engine writes Transform fields and sets dirty (sc_ecs.h:73-96) ----
      if (!s->reparentE.empty() &&
          !scgpuSetParent(s->ctx, (uint32_t)s->reparentE.size(), s->reparentE.data(), s->reparentP.data()))
        noteError(*s, "scgpuSetParent");
      if (!s->dirtyE.empty() && !scgpuSetLocal(s->ctx, (uint32_t)s->dirtyE.size(), s->dirtyE.data(), s->dirtyTrs.data()))
        noteError(*s, "scgpuSetLocal");
      if (!s->renderE.empty() &&
          !scgpuSetRender(s->ctx, (uint32_t)s->renderE.size(), s->renderE.data(), s->renderMM.data(), s->renderBB.data(), s->renderF.data()))
        noteError(*s, "scgpuSetRender");
    }

    // ---- 5. matrices the host itself consumes before culling: cameras (CameraSystem, sc_ecs.cpp:268) ----
    s->needWorld.clear();
    world.ForEach<Camera, Transform>([&](Entity e, Camera&, Transform& t)
    {
      if (!isValidEntity(t.parent))
        t.worldMatrix = mat4_trs(t.localPos, t.localRot, t.localScale);  // a root: world == local, O(1) on the host
      else
        s->needWorld.push_back(e.value);
    });
    if (s->readBackAllWorldMatrices) s->needWorld = s->dense;
    if (!s->needWorld.empty())
    {
      // parented camera (or full read-back requested): run the transform pass now, cull again once the view is known
      const float identity[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
      scgpuSetViews(s->ctx, 1, identity);
      if (!scgpuUpdate(s->ctx, 0)) noteError(*s, "scgpuUpdate(transform)");
      s->transformPassDone = true;
      s->worldOut.resize(s->needWorld.size() * 16);
      if (!scgpuReadWorld(s->ctx, (uint32_t)s->needWorld.size(), s->needWorld.data(), s->worldOut.data())) noteError(*s, "scgpuReadWorld");
      for (size_t i = 0; i < s->needWorld.size(); ++i)
      {
        Entity e{};
        e.value = s->needWorld[i];
        if (Transform* t = world.get<Transform>(e)) std::memcpy(t->worldMatrix.m, s->worldOut.data() + i * 16, 64);
      }
    }
    s->lastHostMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
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
      // ForEach<Transform, RenderMesh> (.cpp:1206-1210): Transform-pool order, filtered by RenderMesh ownership
      cs.candidates.reserve(c.renderablesTotal);
      for (const uint32_t h : s.dense)
        if (s.shadow[h & Entity::INDEX_MASK].flags & SCGPU_HAS_MESH)
        {
          Entity e{};
          e.value = h;
          cs.candidates.push_back(e);
        }
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
    if (!rp.culling)
    {
      // no culling stage registered (.cpp:1330-1346): every Transform + RenderMesh entity is drawn, in Transform-pool
      // order — on the GPU that is the list a frozen culling pass yields; nobody has run the frame's update yet
      GpuSceneState& gs = *st->scene;
      const float identity[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
      if (!scgpuSetViews(gs.ctx, 1, identity)) noteError(gs, "scgpuSetViews");
      if (!scgpuUpdate(gs.ctx, SCGPU_UPDATE_FREEZE_CULLING | (gs.transformPassDone ? SCGPU_UPDATE_SKIP_TRANSFORM : 0u)))
        noteError(gs, "scgpuUpdate");
      gs.transformPassDone = false;
    }

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
