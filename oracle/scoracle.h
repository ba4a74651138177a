/* scoracle — plain-C CPU restatement of the reference's scene-update hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libscgpu.so) never links or calls it.
 *
 * Parity status: PINNED. Every function here is checked in tests/ against (a) the golden bit patterns in
 * SURVEY.md §8c / tests/golden/, which were produced by the reference's own compiled code, and (b) when
 * oracle/_ref/libscref.so is present, against the reference itself on seeded scenes.
 *
 * Each function cites the reference file:line (relative to /root/reference) it restates.
 */
#ifndef SCORACLE_H
#define SCORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SCO_HAS_BOUNDS 1u
#define SCO_HAS_MESH   2u
#define SCO_INVALID_ENTITY 0xFFFFFFFFu /* kInvalidEntity, src/core/include/sc_ecs.h:36 */

/* libm: glibc 2.39 sinf/cosf, generic (non-FMA) variant — the libm std::sin/std::cos(float) resolve to in
 * src/core/src/sc_math.cpp:102-107 on the oracle platform (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c,
 * sincosf.h, s_sincosf_data.c; glibc is a system dependency absent from /root/reference). */
float sco_sinf(float x);
float sco_cosf(float x);

/* src/core/src/sc_math.cpp:52-68 (SSE path: ((a0*b0 + a1*b1) + a2*b2) + a3*b3, no FMA) */
void sco_mat4_mul(const float* a, const float* b, float* out);
/* src/core/src/sc_math.cpp:100-128 */
void sco_mat4_rotation_xyz(float rx, float ry, float rz, float* out);
/* src/core/src/sc_math.cpp:130-142 */
void sco_mat4_trs(const float* pos, const float* rot, const float* scale, float* out);
/* src/core/src/sc_math.cpp:144-207 */
void sco_mat4_inverse(const float* a, float* out);
/* src/core/src/sc_math.cpp:209-232 (tanf from host libm; camera-only, O(1) per view) */
void sco_mat4_perspective_rh_zo(float fovYRadians, float aspect, float zNear, float zFar, int flipY, float* out);

/* src/engine/world/sc_world_partition.cpp:1071-1103 -> 6 planes x (nx,ny,nz,d) */
void sco_frustum_from_viewproj(const float* vp, float* planes24);
/* src/engine/world/sc_world_partition.cpp:1105-1117 */
int sco_sphere_in_frustum(const float* planes24, const float* center, float radius);
/* src/engine/world/sc_world_partition.cpp:1119-1144 */
void sco_world_bounds_sphere(const float* world16, const float* aabb6, float* center, float* radius);

/* TransformSystem, src/core/src/sc_ecs.cpp:118-211, on a dense-order SoA snapshot of the Transform pool.
 *   entity[n]  handles in Transform-pool dense order        parent[n]  parent handles (in/out: fix-ups)
 *   trs[n*9]   localPos, localRot, localScale (in/out: zero-scale patch)
 *   world[n*16] cached world matrices (in/out)              dirty[n]   in/out
 * Returns the number of nodes recomputed, or -1 on allocation failure. */
int64_t sco_transform_system(uint32_t n, const uint32_t* entity, uint32_t* parent, float* trs, float* world,
                             uint8_t* dirty);

/* CullingSystem, src/engine/world/sc_world_partition.cpp:1199-1284, for one view.
 * Candidates = slots with SCO_HAS_MESH, in dense order. No-bounds => visible. freeze => all visible.
 * outVisible/outCulled receive entity handles (stable order); either may be NULL. */
void sco_culling_system(uint32_t n, const uint32_t* entity, const uint32_t* flags, const float* world,
                        const float* aabb, const float* viewProj, int freezeCulling, uint32_t* outVisible,
                        uint32_t* outVisibleCount, uint32_t* outCulled, uint32_t* outCulledCount,
                        uint32_t* outVisibleSlot);

/* RenderPrepStreamingSystem, src/engine/world/sc_world_partition.cpp:1286-1359: 80-byte DrawItem records
 * (src/core/include/sc_ecs.h:159-165: entity@0, meshId@4, materialId@8, model@16). */
void sco_render_prep(uint32_t nVisible, const uint32_t* visibleSlot, const uint32_t* entity,
                     const uint32_t* meshMat2, const float* world, uint32_t maxDraws, void* outDraws80,
                     uint32_t* outEmitted, uint32_t* outDropped);

/* One whole frame: transform, then V views culled, then draw items for view 0. Used by the cpu_baseline. */
int64_t sco_frame(uint32_t n, const uint32_t* entity, uint32_t* parent, float* trs, float* world, uint8_t* dirty,
                  const uint32_t* flags, const float* aabb, uint32_t nViews, const float* viewProjs,
                  uint32_t* visCounts /*[V]*/, uint32_t* visScratch /*[n]*/);

void sco_sincos_sweep(uint32_t first, uint64_t count, uint32_t stride, uint64_t* outSinHash, uint64_t* outCosHash);

/* ---- traffic on rails ----------------------------------------------------------------------
 * libm: glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c, e_exp2f_data.c), atanf (s_atanf.c) and atan2f
 * (e_atan2f.c), generic variants: what std::exp / std::atan2 on float resolve to in
 * src/engine/traffic/sc_traffic_ai.cpp:58-62, 72-75 on the oracle platform. */
float sco_expf(float x);
float sco_atanf(float x);
float sco_atan2f(float y, float x);
/* counts bit mismatches (NaN == NaN) of sco_expf / sco_atanf against the given functions over float bit patterns
 * first, first+stride, ... (count of them) */
uint64_t sco_unary_sweep(int which /*0 expf, 1 atanf*/, uint32_t first, uint64_t count, uint32_t stride, float (*ref)(float));
/* same for sco_atan2f over count pseudo-random (y, x) pairs: raw bit patterns, unit directions and near-equal
 * exponents in turn */
uint64_t sco_atan2_sweep(uint32_t seed, uint64_t count, float (*ref)(float, float));

/* Lane graph in flat arrays: LaneNode / LaneSegment, src/engine/traffic/sc_traffic_lanes.h:14-32 */
typedef struct ScoLaneGraph
{
  uint32_t nNodes, nSegments;
  const float* nodePos;           /* [nNodes*3] */
  const float* nodeSpeedLimit;    /* [nNodes] */
  const uint32_t* nodeConnOffset; /* [nNodes+1] CSR of LaneNode::connections */
  const uint32_t* nodeConn;
  const uint32_t* segNodes;       /* [nSegments*2] startNode, endNode */
  const float* segDir;            /* [nSegments*3] */
  const float* segLength;         /* [nSegments] */
  const uint8_t* segActive;       /* [nSegments] */
  float defaultSpeedLimit;
} ScoLaneGraph;
/* TrafficLaneGraph::advanceAlongLane, src/engine/traffic/sc_traffic_lanes.cpp:291-345 */
int sco_lane_advance(const ScoLaneGraph* g, uint32_t* laneId, float* s, float distance, float* outPos3, float* outDir3);
/* TrafficLaneGraph::queryNearestLane, src/engine/traffic/sc_traffic_lanes.cpp:239-278; returns the lane id */
uint32_t sco_lane_query_nearest(const ScoLaneGraph* g, const float* pos3, float* outS);
/* TrafficAISystem, src/engine/traffic/sc_traffic_ai.cpp:165-487, for n agents in OnRails mode without a physics
 * world: per agent laneId / laneS / targetSpeed / lookAheadDist (TrafficAgent, sc_traffic_common.h:27-37) and the
 * Transform's local TRS (9 floats; localPos.xz and localRot rewritten when the agent moves). obstacleBrake / skip
 * may be NULL. outMoved[i] = 1 when the Transform was written (tr.dirty = true). */
void sco_traffic_ai_on_rails(const ScoLaneGraph* g, uint32_t n, uint32_t* laneId, float* laneS, float* targetSpeed,
                             float* lookAheadDist, float* trs9, const float* obstacleBrake, const uint8_t* skip, float dt,
                             int hasDebug, float dbgLookAheadDist, float dbgSpeedMultiplier, uint8_t* outMoved);

#ifdef __cplusplus
}
#endif
#endif
