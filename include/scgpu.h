/* scgpu.h — C ABI of libscgpu.so: the B200-native scene-update hot path for SandboxCityEngine.
 *
 * One call per frame, scgpuUpdate(), replaces the reference's RenderPrep chain
 *     sc::TransformSystem            /root/reference/src/core/src/sc_ecs.cpp:118-211
 *     sc::CullingSystem              /root/reference/src/engine/world/sc_world_partition.cpp:1199-1284
 *     sc::RenderPrepStreamingSystem  /root/reference/src/engine/world/sc_world_partition.cpp:1286-1359
 * (registered at src/sandbox/src/main.cpp:256-259) on an HBM-resident SoA mirror of the Transform pool.
 * The adapter systems in sc-gameengine_b200/host/ keep the plugin signature void(World&, float, void*)
 * (src/core/include/sc_scheduler.h:38) and call only the functions below.
 *
 * Conventions (style of src/engine/include/sc_engine_render.h:22-63,130-163): opaque context, POD structs,
 * int 1 = ok / 0 = failure, null-tolerant, no exceptions or aborts across the boundary; the failure text is
 * available from scgpuLastError(). The context owns every device buffer and its stream; the caller owns every
 * host pointer it passes. One context is not re-entrant; calls may come from any thread (the device is selected on
 * entry).
 *
 * Host pointers may be pageable or pinned. PAGEABLE memory is consumed before the call returns and may be reused at
 * once. PINNED or registered memory (cudaHostAlloc / cudaHostRegister) is read by the copy engine asynchronously and
 * without an intermediate copy: it must stay untouched until the next call that waits for the stream — any result
 * call, or scgpuSynchronize — has returned. Large pageable arrays are copied through a pinned ring inside the library
 * by the context's helper threads — SCGPU_HOST_THREADS - 1 of them (environment, default min(8, cores / 2)), created
 * with the context and asleep between calls — which also share the pool bookkeeping of large scgpuSpawn / scgpuDespawn
 * batches; SCGPU_HOST_THREADS=1 keeps every call on the calling thread. Results never depend on the thread count.
 *
 * A device failure AFTER a call has already updated the host mirror of the Transform pool (an upload or launch error
 * inside scgpuSpawn* / scgpuDespawn) leaves host and device out of step for good: the context is poisoned, every later
 * call fails and scgpuLastError() keeps the message of the failure that did it.
 *
 * There is no CPU fallback: without a CUDA device scgpuCreate() fails.
 */
#ifndef SCGPU_H
#define SCGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SCGPU_API __declspec(dllexport)
#else
#define SCGPU_API __attribute__((visibility("default")))
#endif

#define SCGPU_API_VERSION 2u
#define SCGPU_MAX_VIEWS 8u
#define SCGPU_INVALID_ENTITY 0xFFFFFFFFu /* sc::kInvalidEntity, src/core/include/sc_ecs.h:36 */

/* per-instance component flags (which ECS components the entity owns besides Transform) */
#define SCGPU_HAS_BOUNDS 1u /* sc::Bounds     (sc_world_partition.h:298-301); absent => always visible (.cpp:1252-1256) */
#define SCGPU_HAS_MESH   2u /* sc::RenderMesh (sc_ecs.h:113-117); absent => not a culling candidate (.cpp:1206-1210) */

/* scgpuUpdate flags */
#define SCGPU_UPDATE_FREEZE_CULLING 1u /* CullingState::freezeCulling (.cpp:1227-1233): every candidate visible */
#define SCGPU_UPDATE_SKIP_TRANSFORM 2u /* cull/compact only (views changed, transforms did not). Consumes no dirty state:
                                        * instances dirtied before it are recomputed by the next update without the flag */
#define SCGPU_UPDATE_CULLED_LISTS   4u /* also build CullingState::culled (.cpp:1273-1280) for every view */

typedef struct ScGpuScene ScGpuScene;

typedef struct ScGpuSceneDesc
{
  uint32_t struct_size;      /* sizeof(ScGpuSceneDesc) */
  int32_t device;            /* CUDA device ordinal */
  uint32_t max_instances;    /* Transform-pool capacity (slots) */
  uint32_t max_entity_index; /* largest Entity::index()+1 that will be seen (<= 1<<24, sc_ecs.h:18-20); 0 => 1<<24 */
  uint32_t max_views;        /* 1..SCGPU_MAX_VIEWS */
  uint32_t flags;            /* reserved, 0 */
  void* stream;              /* cudaStream_t to run on, or NULL for a context-owned stream */
} ScGpuSceneDesc;

/* sc::DrawItem, src/core/include/sc_ecs.h:159-165: entity@0 meshId@4 materialId@8 model@16, 80 bytes */
typedef struct ScGpuDrawItem
{
  uint32_t entity;
  uint32_t meshId;
  uint32_t materialId;
  uint32_t _pad;
  float model[16];
} ScGpuDrawItem;

/* sc::CullingStats (sc_world_partition.h:334-339) per view + sc::RenderPrepStats (:353-357) */
typedef struct ScGpuCounts
{
  uint32_t transforms;                   /* live Transform slots */
  uint32_t renderablesTotal;             /* Transform && RenderMesh */
  uint32_t visible[SCGPU_MAX_VIEWS];
  uint32_t culled[SCGPU_MAX_VIEWS];
  uint32_t recomputed;                   /* world matrices rewritten by the last update */
  uint32_t slowWindows;                  /* diagnostics: hierarchy windows (<= 32 slots) the last update resolved through the
                                          * generic path (parent outside the window, cycle, non-affine or non-finite matrices) */
  uint32_t extent;                       /* diagnostics: device slots the last update walked (transforms + holes) */
} ScGpuCounts;

/* ---- lifetime ------------------------------------------------------------------------------------- */
SCGPU_API uint32_t scgpuGetApiVersion(void);
SCGPU_API ScGpuScene* scgpuCreate(const ScGpuSceneDesc* desc);
SCGPU_API void scgpuDestroy(ScGpuScene* ctx);
/* never NULL; "" when the last call succeeded. ctx may be NULL (reports the last scgpuCreate failure). */
SCGPU_API const char* scgpuLastError(const ScGpuScene* ctx);

/* ---- ECS deltas (replace World::add / World::destroy / setLocal / setParent / markDirty,
 *      src/core/include/sc_ecs.h:73-96,282-330, src/core/src/sc_ecs.cpp:80-90) ------------------------- */
/* Appends n Transforms in pool order. parent: entity handles or SCGPU_INVALID_ENTITY (NULL => none).
 * trs9: localPos, localRot (radians, XYZ Euler), localScale. aabb6: min.xyz,max.xyz (NULL => unit cube
 * [-0.5,0.5]^3, sc_world_partition.cpp:27). meshMat2: meshId, materialId (NULL => 0,0). flags: SCGPU_HAS_*
 * (NULL => BOUNDS|MESH). New instances are dirty with an identity world matrix (sc_ecs.h:63-71). */
SCGPU_API int scgpuSpawn(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, const uint32_t* parent,
                         const float* trs9, const float* aabb6, const uint32_t* meshMat2, const uint32_t* flags);
/* World::destroy for each handle in order, with ComponentPool::remove's swap-with-last (sc_ecs.h:240-262), so
 * the pool order — and therefore every output order — stays identical to the reference's. Unknown or stale
 * handles are skipped like the reference does. */
SCGPU_API int scgpuDespawn(ScGpuScene* ctx, uint32_t n, const uint32_t* entity);
SCGPU_API int scgpuSetLocal(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, const float* trs9);
/* Delta-sized forms of setLocal. The engine's per-frame writers change position and rotation only — the physics sync
 * (src/engine/physics/sc_physics.cpp:1178-1184) and the traffic tiers (src/engine/traffic/sc_traffic_ai.cpp:449-457) —
 * or the position alone (setLocalPosition, sc_ecs.h:92-96); the scale stays as it is. 28 / 16 bytes per instance over
 * PCIe instead of 40. posRot6: localPos.xyz, localRot.xyz. Both mark the instance dirty. */
SCGPU_API int scgpuSetLocalPosRot(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, const float* posRot6);
SCGPU_API int scgpuSetLocalPosition(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, const float* pos3);
/* The same for a RANGE of the Transform pool in its dense order (ComponentPool::denseEntities, what ForEach<Transform>
 * walks): element j is the Transform at dense index firstDense + j. No handles cross PCIe and nothing is looked up.
 * floatsPerInstance selects the fields: SCGPU_LOCAL_POS, SCGPU_LOCAL_POS_ROT or SCGPU_LOCAL_TRS. */
#define SCGPU_LOCAL_POS 3u
#define SCGPU_LOCAL_POS_ROT 6u
#define SCGPU_LOCAL_TRS 9u
SCGPU_API int scgpuSetLocalRange(ScGpuScene* ctx, uint32_t firstDense, uint32_t n, uint32_t floatsPerInstance, const float* data);
/* RenderMesh / Bounds of entities that already own a Transform: World::add<RenderMesh / Bounds> after the fact and
 * edits of meshId, materialId or the AABB (src/engine/traffic/sc_traffic_lod.cpp:47-70, src/engine/src/sc_imgui.cpp:720).
 * Each array may be NULL (that component field is left alone); flags = the new SCGPU_HAS_* bits. Does not dirty the
 * Transform. Unknown handles are skipped. */
SCGPU_API int scgpuSetRender(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, const uint32_t* meshMat2, const float* aabb6,
                             const uint32_t* flags);
SCGPU_API int scgpuSetParent(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, const uint32_t* parent);
SCGPU_API int scgpuMarkDirty(ScGpuScene* ctx, uint32_t n, const uint32_t* entity);
/* Same as scgpuSetLocal but entity/trs9 are DEVICE pointers (producers that already live in HBM). */
SCGPU_API int scgpuSetLocalDevice(ScGpuScene* ctx, uint32_t n, const uint32_t* d_entity, const float* d_trs9);
/* Marks every live instance dirty (first frame / teleport). */
SCGPU_API int scgpuMarkAllDirty(ScGpuScene* ctx);

/* ---- views (replace RenderFrameData::viewProj + frustumFromViewProj, sc_world_partition.cpp:1071-1103) - */
SCGPU_API int scgpuSetViews(ScGpuScene* ctx, uint32_t nViews, const float* viewProj16);
/* planes24 per view: 6 x (nx, ny, nz, d) in the order left,right,bottom,top,near,far */
SCGPU_API int scgpuSetViewPlanes(ScGpuScene* ctx, uint32_t nViews, const float* planes24);
SCGPU_API int scgpuGetViewPlanes(ScGpuScene* ctx, uint32_t view, float* outPlanes24);

/* ---- the frame ----------------------------------------------------------------------------------------- */
/* Enqueues apply-deltas -> transform -> cull (all views, one pass) -> compact on the context stream and
 * returns without waiting. */
SCGPU_API int scgpuUpdate(ScGpuScene* ctx, uint32_t flags);
SCGPU_API int scgpuSynchronize(ScGpuScene* ctx);

/* ---- results (each waits for the last scgpuUpdate) ------------------------------------------------------ */
SCGPU_API int scgpuGetCounts(ScGpuScene* ctx, ScGpuCounts* out);
/* CullingState::visible / ::culled for one view: entity handles in Transform-pool order. cap is the capacity
 * of out in entries; *outCount receives the full count even when it exceeds cap. */
SCGPU_API int scgpuReadVisible(ScGpuScene* ctx, uint32_t view, uint32_t* outEntity, uint32_t cap, uint32_t* outCount);
SCGPU_API int scgpuReadCulled(ScGpuScene* ctx, uint32_t view, uint32_t* outEntity, uint32_t cap, uint32_t* outCount);
/* RenderFrameData::draws for one view; maxDraws = WorldStreamingBudgets::maxDrawsBudget (0 = unlimited). */
SCGPU_API int scgpuReadDrawItems(ScGpuScene* ctx, uint32_t view, uint32_t maxDraws, ScGpuDrawItem* out, uint32_t cap,
                                 uint32_t* outEmitted, uint32_t* outDropped);
/* Transform::worldMatrix (column-major) of the given entities; unknown handles yield zeros and return 0. */
SCGPU_API int scgpuReadWorld(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, float* out16);
/* Transform-pool dense order (ComponentPool::denseEntities) */
SCGPU_API int scgpuReadDenseEntities(ScGpuScene* ctx, uint32_t* outEntity, uint32_t cap, uint32_t* outCount);
SCGPU_API int scgpuReadParents(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, uint32_t* outParent);

/* ---- zero-copy access for device-side consumers ---------------------------------------------------------- */
typedef struct ScGpuDeviceViews
{
  const uint32_t* visibleEntity[SCGPU_MAX_VIEWS]; /* device, compacted entity handles per view */
  const uint32_t* visibleSlot[SCGPU_MAX_VIEWS];   /* device, the matching Transform-pool slots */
  const uint32_t* visibleCount;                   /* device, [nViews] visible per view, [nViews] = renderablesTotal */
  const float* worldCol[4];                       /* device, world matrix column planes (float4 per slot) */
  const uint32_t* entity;                         /* device, slot -> entity handle (SCGPU_INVALID_ENTITY: a hole) */
  uint32_t count;                                 /* live Transforms (size of the reference's pool) */
  uint32_t extent;                                /* device slots in use: live Transforms + holes left by despawns */
  const uint32_t* rank;                           /* device, slot -> dense index in the reference's Transform pool */
  const uint32_t* perm;                           /* device, dense index -> slot */
} ScGpuDeviceViews;
SCGPU_API int scgpuGetDeviceViews(ScGpuScene* ctx, ScGpuDeviceViews* out);
SCGPU_API void* scgpuGetStream(ScGpuScene* ctx);
/* Materialises draw items for one view into a context-owned device buffer and returns it. */
SCGPU_API int scgpuBuildDrawItemsDevice(ScGpuScene* ctx, uint32_t view, uint32_t maxDraws,
                                        const ScGpuDrawItem** outDevice, uint32_t* outEmitted, uint32_t* outDropped);

/* ---- what the renderer does with RenderFrameData::draws, on the device ---------------------
 * Replaces, for one view, the per-frame CPU work of src/engine/src/sc_vk.cpp:1843-1905: drop draws with an
 * out-of-range meshId or an unknown material (:1847-1850), sort by (pipelineId of the material, materialId, meshId)
 * (:1854-1864), and derive the bind-on-change batches of the submission loop (:1866-1905) as RUNS of consecutive
 * sorted items that share pipeline, material and mesh. std::sort is not stable; this sort is (ties keep
 * CullingState::visible order), i.e. it yields one of the orders the reference may produce. */
typedef struct ScGpuDrawRun
{
  uint32_t pipelineId; /* sc::PipelineId of the material (src/engine/include/sc_assets.h) */
  uint32_t materialId;
  uint32_t meshId;
  uint32_t first;      /* index of the run's first item in the sorted DrawItem array */
  uint32_t count;      /* instances in the run */
} ScGpuDrawRun;
/* materialPipeline[m] = pipeline id (< 63) of material handle m, or 0xFFFFFFFF when AssetManager::getMaterial(m)
 * would return null; nMaterials, meshCount <= 2^29. maxDraws as in scgpuReadDrawItems (applied BEFORE the filter,
 * like RenderPrepStreamingSystem applies its budget before the renderer filters). Results stay valid until the next
 * scgpuUpdate; the device pointers are context-owned. */
SCGPU_API int scgpuBuildSortedDraws(ScGpuScene* ctx, uint32_t view, uint32_t maxDraws, const uint32_t* materialPipeline,
                                    uint32_t nMaterials, uint32_t meshCount, const ScGpuDrawItem** outDevice,
                                    uint32_t* outKept, const ScGpuDrawRun** outRunsDevice, uint32_t* outRuns);
SCGPU_API int scgpuReadSortedDraws(ScGpuScene* ctx, ScGpuDrawItem* outItems, uint32_t cap, ScGpuDrawRun* outRuns,
                                   uint32_t runCap);

/* ---- procedural sectors spawned on the device ------------------------------------------------
 * Replaces, for procedurally generated sectors, generateSectorSpawnsStatic (src/engine/world/sc_world_partition.cpp:
 * 105-169) and the per-record World::add<Transform/RenderMesh/Bounds> + setLocal loop of pumpCompletedLoads (:923-954):
 * the SoA records are generated in HBM from (seed, sector coordinate) with the reference's hash and float expressions;
 * the host only creates the entity handles (World::create) and passes them in spawn order (ground plane first, then
 * the props). No TRS, bounds or mesh/material data crosses PCIe. */
typedef struct ScGpuSectorGen
{
  uint32_t struct_size;         /* sizeof(ScGpuSectorGen) */
  float sectorSizeMeters;       /* WorldPartitionConfig (sc_world_partition.h) */
  uint32_t seed;
  uint32_t propsPerSectorMin, propsPerSectorMax;
  uint32_t includeGroundPlane;
  uint32_t meshCube, meshTriangle;          /* resolveMeshHandle("meshes/cube" | "meshes/triangle") */
  uint32_t matUnlit, matChecker, matTest;   /* resolveMaterialHandle("materials/unlit" | "checker" | "test") */
} ScGpuSectorGen;
/* number of SpawnRecords the sector yields (= entities the caller must create for it); pure host arithmetic */
SCGPU_API uint32_t scgpuSectorSpawnCount(const ScGpuSectorGen* gen, int32_t x, int32_t z);
/* coordXZ: nSectors x (x, z); entity: the handles of all sectors back to back, nEntities = sum of the spawn counts */
SCGPU_API int scgpuSpawnSectors(ScGpuScene* ctx, const ScGpuSectorGen* gen, uint32_t nSectors, const int32_t* coordXZ,
                                const uint32_t* entity, uint32_t nEntities);

/* ---- authored sectors, .scsector INST chunk -> SoA ----------------------------------------------
 * Replaces, for sectors stored on disk, sc_world::ReadSectorFile's INST branch (tools/shared/world_format.cpp:207-281),
 * WorldPartition::readSectorFile (src/engine/world/sc_world_partition.cpp:695-732) and the World::add loop of
 * pumpCompletedLoads (:923-954): the caller hands over the file image as it lies on disk; the chunk table is walked on
 * the host exactly like the reference's reader does (format versions 1..4: optional model id, name, overrides, padding),
 * the raw INST payload is uploaded and unpacked into the record planes by one kernel. Asset ids resolve like
 * resolveMeshHandle / resolveMaterialHandle (:746-800): id 0 -> handle 0, listed id -> its handle, else the default. */
typedef struct ScGpuAssetBinding
{
  uint64_t assetId; /* sc_world::AssetId (FNV-1a of the normalised path, world_format.cpp:63-74) */
  uint32_t handle;  /* MeshHandle / MaterialHandle the engine resolved it to */
  uint32_t _pad;
} ScGpuAssetBinding;
typedef struct ScGpuAssetTable
{
  const ScGpuAssetBinding* meshes;
  uint32_t nMeshes, defaultMesh;        /* default: the handle of "meshes/cube" */
  const ScGpuAssetBinding* materials;
  uint32_t nMaterials, defaultMaterial; /* default: the handle of "materials/unlit" */
} ScGpuAssetTable;
/* host only: sector coordinate, format version and instance count of a file image (0 on a malformed image) */
SCGPU_API int scgpuSectorFileInfo(const void* bytes, size_t nBytes, int32_t* outXZ, uint32_t* outVersion, uint32_t* outInstances);
/* entity: one handle per instance in file order (the caller created them with World::create) */
SCGPU_API int scgpuSpawnSectorFile(ScGpuScene* ctx, const void* bytes, size_t nBytes, const uint32_t* entity, uint32_t nEntities,
                                   const ScGpuAssetTable* assets);

/* ---- SURVEY.md 8(f) N4 (the second mat4_trs caller): the world editor's draw list ------------------------------
 * Replaces BuildDrawItems (tools/world_editor/editor_core/editor_core.cpp:242-264): one ScRenderDrawItem
 * (src/engine/include/sc_engine_render.h:24,51-57: 64-bit mesh and material handles, model[16], flags; 88 bytes) per
 * entity whose mesh and
 * material handles are non-zero, model = mat4_trs(position, rotation, scale), document order kept. Stateless with
 * respect to the scene: the context only lends its device and stream. *outCount = kept entities (may exceed cap). */
typedef struct ScGpuEditorDrawItem
{
  uint64_t mesh;     /* ScRenderHandle */
  uint64_t material; /* ScRenderHandle */
  float model[16];
  uint32_t flags;
} ScGpuEditorDrawItem;
SCGPU_API int scgpuBuildEditorDraws(ScGpuScene* ctx, uint32_t n, const float* trs9, const uint64_t* meshHandle,
                                    const uint64_t* materialHandle, ScGpuEditorDrawItem* out, uint32_t cap, uint32_t* outCount);

/* ---- traffic on rails, a device-resident dirty producer -------------------------------------
 * Replaces, for the agents whose TrafficVehicle::mode is OnRails, the per-agent body of sc::TrafficAISystem
 * (src/engine/traffic/sc_traffic_ai.cpp:165-487, on-rails branch :434-458): speed smoothing (smoothExp :58-62),
 * TrafficLaneGraph::advanceAlongLane / chooseNextSegment / queryNearestLane / laneSpeedLimit
 * (src/engine/traffic/sc_traffic_lanes.cpp:150-169, 239-345, 392-400) and yawFromDir (:72-75). localPos / localRot
 * of the agents' Transforms are written in HBM and stamped dirty for the next scgpuUpdate: nothing crosses PCIe
 * per frame. Bit-exact with the reference (glibc 2.39 expf / atan2f restated on the device).
 * The physics and kinematic tiers (at most 24 + 64 vehicles, sc_traffic_common.h:70-73) need the Bullet world and
 * stay on the host; tier changes (sc_traffic_lod.cpp) are the host calling scgpuTrafficSetAgents again. */
typedef struct ScGpuLaneGraph
{
  uint32_t struct_size;           /* sizeof(ScGpuLaneGraph) */
  uint32_t nNodes, nSegments, nConnections;
  const float* nodePos;           /* [nNodes*3]   LaneNode::pos        (sc_traffic_lanes.h:14-20) */
  const float* nodeSpeedLimit;    /* [nNodes]     LaneNode::speedLimit */
  const uint32_t* nodeConnOffset; /* [nNodes+1]   CSR offsets of LaneNode::connections */
  const uint32_t* nodeConn;       /* [nConnections] segment ids (ids >= nSegments are skipped like :156-157) */
  const uint32_t* segNodes;       /* [nSegments*2] LaneSegment::startNode, endNode (sc_traffic_lanes.h:22-31) */
  const float* segDir;            /* [nSegments*3] LaneSegment::dir */
  const float* segLength;         /* [nSegments]   LaneSegment::length */
  const uint8_t* segActive;       /* [nSegments]   LaneSegment::active; NULL = all active */
  float defaultSpeedLimit;        /* TrafficLaneGraph::speedLimit() */
} ScGpuLaneGraph;
/* uploads (replaces) the lane graph; node indices of segments must be < nNodes */
SCGPU_API int scgpuTrafficSetLanes(ScGpuScene* ctx, const ScGpuLaneGraph* graph);
/* TrafficLaneGraph::removeSector / re-activation (sc_traffic_lanes.cpp:171-183, 224-236) for n segments */
SCGPU_API int scgpuTrafficSetLaneActive(ScGpuScene* ctx, uint32_t n, const uint32_t* segment, const uint8_t* active);
/* replaces the on-rails agent set: TrafficAgent::{laneId, laneS, targetSpeed, lookAheadDist} per entity
 * (sc_traffic_common.h:27-37). Entities without a Transform in the context are skipped by the step. */
SCGPU_API int scgpuTrafficSetAgents(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, const uint32_t* laneId,
                                    const float* laneS, const float* targetSpeed, const float* lookAheadDist);
typedef struct ScGpuTrafficStep
{
  uint32_t struct_size;       /* sizeof(ScGpuTrafficStep) */
  float dt;
  uint32_t hasDebug;          /* TrafficAIState::debug != nullptr: the two fields below apply (:237-238, :296-297) */
  float lookAheadDist;        /* TrafficDebugState::lookAheadDist */
  float speedMultiplier;      /* TrafficDebugState::speedMultiplier */
  const float* obstacleBrake; /* [nAgents] host raycast result (:300-347) or NULL = 0 (TrafficAIState::physics null) */
  const uint8_t* skip;        /* [nAgents] non-zero = agent's sector is not Active (:218-226) or NULL */
} ScGpuTrafficStep;
/* one TrafficAISystem pass over the on-rails agents, asynchronous on the context stream; *outMoved (optional,
 * forces a synchronisation) = number of Transforms written */
SCGPU_API int scgpuTrafficAdvance(ScGpuScene* ctx, const ScGpuTrafficStep* step, uint32_t* outMoved);
/* agent state in scgpuTrafficSetAgents order */
SCGPU_API int scgpuTrafficReadAgents(ScGpuScene* ctx, uint32_t cap, uint32_t* outLaneId, float* outLaneS,
                                     float* outTargetSpeed, float* outLookAheadDist, uint32_t* outCount);
/* Transform::localPos / localRot / localScale of n entities (9 floats each) */
SCGPU_API int scgpuReadLocal(ScGpuScene* ctx, uint32_t n, const uint32_t* entity, float* outTrs9);

/* ---- multi-GPU: one context per process per GPU, instance set sharded by world cell ------------------------
 * The only exchange is the gather of the compacted per-view lists and counts to the submitting rank. */
#define SCGPU_COMM_ID_BYTES 128
SCGPU_API int scgpuCommGetUniqueId(void* outId128);
SCGPU_API int scgpuCommInit(ScGpuScene* ctx, uint32_t nRanks, uint32_t rank, const void* id128);
/* After scgpuUpdate: gathers every rank's counts to all ranks and every rank's visible lists to `root`,
 * concatenated in rank order (shard-major stable order). Enqueued on the context stream. */
SCGPU_API int scgpuGatherVisible(ScGpuScene* ctx, uint32_t root);
/* Optional, collective (every rank calls it with the same arguments, once): from now on the lists travel through
 * NVLink PEER MEMORY instead of NCCL. Every scgpuUpdate's last kernel stores the entity handles it resolves straight
 * into a mailbox in the root's HBM as well (cudaIpc mapping, handle carried by one ncclBroadcast here) and raises a
 * flag — compute and peer store in one kernel; scgpuGatherVisible(root) is then one small kernel on the root that waits
 * for the flags, and nothing at all on the other ranks. No host synchronisation and no NCCL call per frame. All ranks
 * must issue the same number of scgpuUpdate calls per frame (a producer may run one update ahead of the root, not two).
 * capEntries = mailbox capacity per rank in list entries (0 => min(max_instances * max_views, 4 Mi)); a frame whose
 * lists exceed it fails in the read-back calls — that frame only. With the peer gather the counts and lists exist on
 * the root only. */
SCGPU_API int scgpuCommEnablePeerGather(ScGpuScene* ctx, uint32_t root, uint32_t capEntries);
/* counts[rank][view] for all ranks (NCCL gather: valid on every rank; peer gather: on the root) */
SCGPU_API int scgpuGetGatheredCounts(ScGpuScene* ctx, uint32_t* outCounts, uint32_t capRanks);
/* root only: concatenated list of one view; *outCount = sum over ranks */
SCGPU_API int scgpuReadGatheredVisible(ScGpuScene* ctx, uint32_t view, uint32_t* outEntity, uint32_t cap,
                                       uint32_t* outCount);

/* ---- introspection used by bench.py ------------------------------------------------------------------------ */
/* number of kernel launches this context has issued so far */
SCGPU_API uint64_t scgpuKernelLaunchCount(ScGpuScene* ctx);
/* device time of the main fused kernel and of the whole last update, from CUDA events on the context stream */
SCGPU_API int scgpuLastUpdateTimings(ScGpuScene* ctx, float* outFusedKernelMs, float* outUpdateMs);
/* enable: 0 = off, 1 = every update, n > 1 = every n-th update (the event pairs around the fused kernel keep the
 * compaction from overlapping its tail, so a benchmark may sample them) */
SCGPU_API int scgpuEnableTimings(ScGpuScene* ctx, int enable);
/* per-launch device times (ms) of up to the last 256 updates since timings were enabled, oldest first */
SCGPU_API int scgpuReadUpdateTimings(ScGpuScene* ctx, float* outFusedKernelMs, float* outUpdateMs, uint32_t cap,
                                     uint32_t* outCount);

#ifdef __cplusplus
}
#endif
#endif /* SCGPU_H */
