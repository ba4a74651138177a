/* scref — C driver around the UNMODIFIED reference sources (test infrastructure only).
 *
 * Built by oracle/Makefile into oracle/_ref/libscref.so from the sources where they lie under
 * /root/reference (nothing is copied). Every entry point calls the reference's own symbols:
 *   sc::TransformSystem            src/core/src/sc_ecs.cpp:118-211
 *   sc::CameraSystem               src/core/src/sc_ecs.cpp:213-272
 *   sc::CullingSystem              src/engine/world/sc_world_partition.cpp:1199-1284
 *   sc::RenderPrepStreamingSystem  src/engine/world/sc_world_partition.cpp:1286-1359
 *   sc::mat4_* / frustumFromViewProj / sphereInFrustum / computeWorldBoundsSphere
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#ifndef SCREF_API_H
#define SCREF_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ScRefWorld ScRefWorld;

#define SCREF_HAS_BOUNDS 1u
#define SCREF_HAS_MESH   2u
#define SCREF_NO_PARENT  0xFFFFFFFFu

/* jobWorkers==0 -> hardware_concurrency-1 like src/sandbox/src/main.cpp:52-54 */
ScRefWorld* screfCreate(uint32_t jobWorkers);
void screfDestroy(ScRefWorld* w);
uint32_t screfJobWorkers(void);
/* joins the reference job workers (src/core/src/sc_jobs.cpp:152-171); also registered with atexit */
void screfShutdown(void);

/* ECS mutation through the reference's World (sc_ecs.h:282-418) */
void screfCreateEntities(ScRefWorld* w, uint32_t n, uint32_t* outEntity);
void screfAddInstances(ScRefWorld* w, uint32_t n, const uint32_t* entity, const uint32_t* parent,
                       const float* trs9, const float* aabb6, const uint32_t* meshMat2, const uint32_t* flags);
void screfDestroyEntities(ScRefWorld* w, uint32_t n, const uint32_t* entity);
void screfSetLocal(ScRefWorld* w, uint32_t n, const uint32_t* entity, const float* trs9);
void screfSetParent(ScRefWorld* w, uint32_t n, const uint32_t* entity, const uint32_t* parent);
void screfMarkDirty(ScRefWorld* w, uint32_t n, const uint32_t* entity);

/* Systems (the plugin signature void(World&,float,void*), sc_scheduler.h:38) */
void screfRunTransform(ScRefWorld* w);
void screfSetViewProj(ScRefWorld* w, const float* viewProj16);
void screfSetFreezeCulling(ScRefWorld* w, int freeze);
void screfRunCulling(ScRefWorld* w);
/* maxDraws: WorldStreamingBudgets::maxDrawsBudget (0 = unlimited) */
void screfRunRenderPrep(ScRefWorld* w, uint32_t maxDraws);
/* Camera: adds a camera entity (Transform+Camera active) and runs sc::CameraSystem to fill viewProj */
uint32_t screfAddCamera(ScRefWorld* w, const float* trs9, float fovY, float nearZ, float farZ, float aspect);
void screfRunCamera(ScRefWorld* w, float aspect);

/* Read-back */
uint32_t screfTransformCount(ScRefWorld* w);
/* Transform-pool dense order (defines every output order, sc_ecs.h:199-277) */
uint32_t screfDenseEntities(ScRefWorld* w, uint32_t cap, uint32_t* outEntity);
void screfReadWorld(ScRefWorld* w, uint32_t n, const uint32_t* entity, float* out16);
void screfReadTransform(ScRefWorld* w, uint32_t n, const uint32_t* entity, uint32_t* outParent, float* outTrs9,
                        uint8_t* outDirty);
/* which of Bounds / RenderMesh each entity owns, with their contents */
void screfReadComponents(ScRefWorld* w, uint32_t n, const uint32_t* entity, uint32_t* outFlags, float* outAabb6,
                         uint32_t* outMeshMat2);
void screfGetViewProj(ScRefWorld* w, float* out16);
void screfGetPlanes(ScRefWorld* w, float* out24);
void screfGetCullStats(ScRefWorld* w, uint32_t* total, uint32_t* visible, uint32_t* culled);
uint32_t screfReadVisible(ScRefWorld* w, uint32_t cap, uint32_t* out);
uint32_t screfReadCulled(ScRefWorld* w, uint32_t cap, uint32_t* out);
uint32_t screfReadCandidates(ScRefWorld* w, uint32_t cap, uint32_t* out);
void screfGetRenderPrepStats(ScRefWorld* w, uint32_t* emitted, uint32_t* dropped);
/* 80-byte DrawItem records (sc_ecs.h:159-165) copied verbatim */
uint32_t screfReadDraws(ScRefWorld* w, uint32_t cap, void* out80);
uint32_t screfSizeofTransform(void);
uint32_t screfSizeofDrawItem(void);

/* Config 1: the sandbox's default streamed scene (src/sandbox/src/main.cpp:66-99,241-263) run through the
 * reference Scheduler for `frames` frames. Returns number of active sectors. */
uint32_t screfBuildDefaultScene(ScRefWorld* w, uint32_t frames);

/* Frame timing of the reference path: runs Transform -> (Culling per view) -> RenderPrep `iters` times.
 * Before every iteration all entities in dirtyEntity[0..nDirty) get markDirty. Returns seconds per stage. */
void screfTimeFrame(ScRefWorld* w, uint32_t iters, uint32_t nViews, const float* viewProj16xV,
                    uint32_t nDirty, const uint32_t* dirtyEntity, uint32_t maxDraws,
                    double* outTransformS, double* outCullS, double* outPrepS);

/* Raw math entry points for known-answer tests */
void screfMat4Trs(const float* pos3, const float* rot3, const float* scale3, float* out16);
void screfMat4Mul(const float* a16, const float* b16, float* out16);
void screfMat4Inverse(const float* a16, float* out16);
void screfMat4Perspective(float fovYRad, float aspect, float zn, float zf, int flipY, float* out16);
void screfFrustumFromViewProj(const float* vp16, float* out24);
int  screfSphereInFrustum(const float* planes24, const float* center3, float radius);
void screfWorldBoundsSphere(const float* world16, const float* aabb6, float* outCenter3, float* outRadius);
float screfSinf(float x);
float screfCosf(float x);
/* XOR-fold hash of sinf/cosf bits over float bit patterns [first, first+count) with stride */
void screfSinCosSweep(uint32_t first, uint64_t count, uint32_t stride, uint64_t* outSinHash, uint64_t* outCosHash);

/* .scsector files through the reference's own writer / reader (tools/shared/world_format.cpp:76-334) */
int screfWriteSectorFile(const char* path, uint32_t version, int32_t x, int32_t z, uint32_t n, const uint64_t* id,
                         const uint64_t* modelId, const uint64_t* meshId, const uint64_t* materialId, const float* trs9,
                         const uint32_t* tags, uint32_t nExtraChunkBytes);
int screfReadSectorInstances(const char* path, uint32_t cap, int32_t* outXZ, uint64_t* outId, uint64_t* outMeshId,
                             uint64_t* outMaterialId, float* outTrs9);
uint64_t screfHashAssetPath(const char* path);

/* Traffic: the reference's TrafficLaneGraph and TrafficAISystem, scref_traffic.cpp */
typedef struct ScRefLanes ScRefLanes;
ScRefLanes* screfLanesCreate(float laneWidth, float speedLimit);
void screfLanesDestroy(ScRefLanes* l);
uint32_t screfLanesAddNode(ScRefLanes* l, const float* pos3, const float* dir3, float speedLimit);
uint32_t screfLanesAddSegment(ScRefLanes* l, uint32_t a, uint32_t b, const float* dir3, int32_t ownerX, int32_t ownerZ);
/* buildProceduralForSector / removeSector (sc_traffic_lanes.cpp:171-237) with the sector's ground square as bounds */
void screfLanesBuildSector(ScRefLanes* l, int32_t x, int32_t z, float sectorSize);
void screfLanesRemoveSector(ScRefLanes* l, int32_t x, int32_t z);
void screfLanesSetActive(ScRefLanes* l, uint32_t segment, int active);
void screfLanesCounts(ScRefLanes* l, uint32_t* nNodes, uint32_t* nSegs, uint32_t* nConn);
void screfLanesExport(ScRefLanes* l, float* nodePos3, float* nodeSpeed, uint32_t* connOffset, uint32_t* conn,
                      uint32_t* segNodes2, float* segDir3, float* segLen, uint8_t* segActive, float* defaultSpeed);
int screfLaneAdvance(ScRefLanes* l, uint32_t* laneId, float* s, float distance, float* outPos3, float* outDir3);
uint32_t screfLaneQueryNearest(ScRefLanes* l, const float* pos3, float* outS);
void screfTrafficAddAgents(ScRefWorld* w, uint32_t n, const uint32_t* entity, const uint32_t* laneId, const float* laneS,
                           const float* targetSpeed, const float* lookAhead);
void screfTrafficSetPlayer(ScRefWorld* w, uint32_t entity);
void screfRunTrafficAI(ScRefWorld* w, ScRefLanes* l, float dt, int useDebug, float lookAheadDist, float speedMultiplier);
void screfTrafficReadAgents(ScRefWorld* w, uint32_t n, const uint32_t* entity, uint32_t* laneId, float* laneS,
                            float* targetSpeed, float* lookAhead);
float screfExpf(float x);
float screfAtanf(float x);
float screfAtan2f(float y, float x);

#ifdef __cplusplus
}
#endif
#endif
