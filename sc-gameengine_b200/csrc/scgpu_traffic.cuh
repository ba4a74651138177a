// scgpu_traffic.cuh — SURVEY §8(f) N4: the on-rails tier of the traffic AI as a device-resident dirty producer.
//
// Replaces, for agents whose TrafficVehicle::mode is OnRails (the bulk tier: up to maxTrafficVehiclesTotal minus
// the physics / kinematic caps, sc_traffic_common.h:70-73), the per-agent body of
//   sc::TrafficAISystem                 src/engine/traffic/sc_traffic_ai.cpp:165-487 (branch :434-458)
//   TrafficLaneGraph::advanceAlongLane  src/engine/traffic/sc_traffic_lanes.cpp:291-345
//   TrafficLaneGraph::chooseNextSegment src/engine/traffic/sc_traffic_lanes.cpp:150-169
//   TrafficLaneGraph::queryNearestLane  src/engine/traffic/sc_traffic_lanes.cpp:239-278
//   TrafficLaneGraph::laneSpeedLimit    src/engine/traffic/sc_traffic_lanes.cpp:392-400
// It writes localPos / localRot of the agent's Transform in HBM and stamps it dirty, so the frame's largest dirty
// source never crosses PCIe. Everything is FP32 in the reference's operation order (no contraction); expf and
// atan2f are the glibc restatements of scgpu_math.cuh.
//
// The per-agent logic is plain __device__ code over a LaneGraphView so that tests/hostsim can compile it for the
// host and check it against the oracle without a GPU.
#pragma once
#include "scgpu_math.cuh"

namespace scgpu
{

constexpr uint32_t kInvalidLane = 0xFFFFFFFFu;  // sc::kInvalidLaneId, sc_traffic_common.h:9

// Lane graph in HBM (sc_traffic_lanes.h:14-32): 16-byte records, one load each
struct LaneGraphView
{
  const float4* nodePos;    // LaneNode::pos xyz, speedLimit in w
  const uint2* nodeConn;    // LaneNode::connections as (first, count) into conn[]
  const uint32_t* conn;     // segment ids
  const float4* segDirLen;  // LaneSegment::dir xyz, length in w
  const uint4* segNodes;    // startNode, endNode, active, unused
  uint32_t nNodes, nSegs;
  float defaultSpeedLimit;  // TrafficLaneGraph::m_speedLimit
};

struct TrafficStepParams
{
  float dt;
  float speedMultiplier;  // TrafficDebugState::speedMultiplier, applied when hasDebug
  float lookAheadDist;    // TrafficDebugState::lookAheadDist, copied into every agent when hasDebug (:237-238)
  uint32_t hasDebug;      // TrafficAIState::debug != nullptr
};

// dot3, sc_traffic_lanes.cpp:29-32: (a0*b0 + a1*b1) + a2*b2
__device__ __forceinline__ float lane_dot3(float ax, float ay, float az, float bx, float by, float bz)
{
  return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// chooseNextSegment, sc_traffic_lanes.cpp:150-169: the active outgoing segment with the largest dot product
// against the incoming direction (first one wins ties; nothing qualifies unless dot > -1)
__device__ __forceinline__ uint32_t lane_choose_next(const LaneGraphView& g, float dx, float dy, float dz, uint32_t node)
{
  uint32_t best = kInvalidLane;
  float bestDot = -1.0f;
  const uint2 c = g.nodeConn[node];
  for (uint32_t i = 0; i < c.y; ++i)
  {
    const uint32_t segId = g.conn[c.x + i];
    if (segId >= g.nSegs) continue;
    if (g.segNodes[segId].z == 0u) continue;
    const float4 sd = g.segDirLen[segId];
    const float d = lane_dot3(dx, dy, dz, sd.x, sd.y, sd.z);
    if (d > bestDot) { bestDot = d; best = segId; }
  }
  return best;
}

// advanceAlongLane, sc_traffic_lanes.cpp:291-345 (at most 8 segment hops)
__device__ __forceinline__ bool lane_advance(const LaneGraphView& g, uint32_t& laneId, float& s, float distance,
                                             float outPos[3], float outDir[3])
{
  if (laneId == kInvalidLane || laneId >= g.nSegs) return false;
  float remaining = distance;
  uint32_t current = laneId;
  float currentS = s;
  for (uint32_t guard = 0; guard < 8; ++guard)
  {
    const uint4 sn = g.segNodes[current];
    if (sn.z == 0u) return false;
    const float4 sd = g.segDirLen[current];
    const float len = sd.w;
    if (len <= 1e-5f) return false;
    const float available = __fsub_rn(len, currentS);
    if (remaining <= available)
    {
      currentS = __fadd_rn(currentS, remaining);
      const float4 a = g.nodePos[sn.x];
      outPos[0] = __fadd_rn(a.x, __fmul_rn(sd.x, currentS));
      outPos[1] = __fadd_rn(a.y, __fmul_rn(sd.y, currentS));
      outPos[2] = __fadd_rn(a.z, __fmul_rn(sd.z, currentS));
      outDir[0] = sd.x; outDir[1] = sd.y; outDir[2] = sd.z;
      laneId = current;
      s = currentS;
      return true;
    }
    remaining = __fsub_rn(remaining, available);
    currentS = 0.0f;
    const uint32_t next = lane_choose_next(g, sd.x, sd.y, sd.z, sn.y);
    if (next == kInvalidLane)
    {
      const float4 e = g.nodePos[sn.y];
      outPos[0] = e.x; outPos[1] = e.y; outPos[2] = e.z;
      outDir[0] = sd.x; outDir[1] = sd.y; outDir[2] = sd.z;
      laneId = current;
      s = len;
      return true;
    }
    current = next;
  }
  return false;
}

// queryNearestLane, sc_traffic_lanes.cpp:239-278: linear scan over every active segment (only agents without a
// lane take it; the spawner always assigns one, sc_traffic_spawner.cpp:283-316)
__device__ __forceinline__ uint32_t lane_query_nearest(const LaneGraphView& g, float px, float py, float pz, float& outS)
{
  uint32_t bestLane = kInvalidLane;
  float bestDist = 0.0f;
  bool hasBest = false;
  for (uint32_t i = 0; i < g.nSegs; ++i)
  {
    const uint4 sn = g.segNodes[i];
    const float4 sd = g.segDirLen[i];
    if (sn.z == 0u || sd.w <= 1e-5f) continue;
    const float4 a = g.nodePos[sn.x];
    const float tx = __fsub_rn(px, a.x), ty = __fsub_rn(py, a.y), tz = __fsub_rn(pz, a.z);
    const float proj = lane_dot3(tx, ty, tz, sd.x, sd.y, sd.z);
    const float mn = (proj < sd.w) ? proj : sd.w;  // std::min(seg.length, proj)
    const float s = (0.0f < mn) ? mn : 0.0f;       // std::max(0.0f, .)
    const float dx = __fsub_rn(px, __fadd_rn(a.x, __fmul_rn(sd.x, s)));
    const float dy = __fsub_rn(py, __fadd_rn(a.y, __fmul_rn(sd.y, s)));
    const float dz = __fsub_rn(pz, __fadd_rn(a.z, __fmul_rn(sd.z, s)));
    const float distSq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (!hasBest || distSq < bestDist)
    {
      hasBest = true;
      bestDist = distSq;
      bestLane = i;
      outS = s;
    }
  }
  return bestLane;
}

// laneSpeedLimit, sc_traffic_lanes.cpp:392-400
__device__ __forceinline__ float lane_speed_limit(const LaneGraphView& g, uint32_t laneId)
{
  if (laneId == kInvalidLane || laneId >= g.nSegs) return g.defaultSpeedLimit;
  const uint32_t a = g.segNodes[laneId].x;
  if (a >= g.nNodes) return g.defaultSpeedLimit;
  return g.nodePos[a].w;
}

// One on-rails agent, sc_traffic_ai.cpp:228-458 with tv.mode == OnRails. `obstacleBrake` is the result of the
// host's front-sensor raycast (:300-347; 0 when TrafficAIState::physics is null). Returns true when the Transform
// was moved (pos[0], pos[2] and yaw written; pos[1] kept, :447).
__device__ __forceinline__ bool traffic_agent_on_rails(const LaneGraphView& g, const TrafficStepParams& st,
                                                       float obstacleBrake, uint32_t& laneId, float& laneS,
                                                       float& targetSpeed, float& lookAhead, float pos[3], float& yaw)
{
  if (st.hasDebug) lookAhead = st.lookAheadDist;
  if (laneId == kInvalidLane)
  {
    float qs = 0.0f;
    const uint32_t q = lane_query_nearest(g, pos[0], pos[1], pos[2], qs);
    if (q != kInvalidLane) { laneId = q; laneS = qs; }
  }
  // getLane: null or inactive => nothing happens this frame (:274-276)
  if (laneId == kInvalidLane || laneId >= g.nSegs) return false;
  if (g.segNodes[laneId].z == 0u) return false;

  // getLookAheadPoint (:278-280) only decides whether the agent is processed at all
  float target[3], dir[3];
  {
    uint32_t id = laneId;
    float ss = laneS;
    if (!lane_advance(g, id, ss, lookAhead, target, dir)) return false;
  }
  const float tx = __fsub_rn(target[0], pos[0]), tz = __fsub_rn(target[2], pos[2]);
  const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(0.0f, 0.0f)), __fmul_rn(tz, tz)));
  if (len < 1e-4f) return false;

  float desiredSpeed = lane_speed_limit(g, laneId);
  if (st.hasDebug) desiredSpeed = __fmul_rn(desiredSpeed, st.speedMultiplier);
  desiredSpeed = (0.0f < desiredSpeed) ? desiredSpeed : 0.0f;  // std::max(0.0f, desiredSpeed)

  // on-rails branch (:434-458)
  const float desired = __fmul_rn(desiredSpeed, __fsub_rn(1.0f, obstacleBrake));
  const float t = __fsub_rn(1.0f, expf_glibc(__fmul_rn(-2.5f, st.dt)));  // smoothExp(.., 2.5f, dt)
  targetSpeed = __fadd_rn(targetSpeed, __fmul_rn(__fsub_rn(desired, targetSpeed), t));
  const float travel = __fmul_rn(targetSpeed, st.dt);
  uint32_t id = laneId;
  float ss = laneS;
  float p[3];
  if (!lane_advance(g, id, ss, travel, p, dir)) return false;
  laneId = id;
  laneS = ss;
  pos[0] = p[0];
  pos[2] = p[2];
  yaw = atan2f_glibc(dir[0], dir[2]);  // yawFromDir
  return true;
}

}  // namespace scgpu

#ifdef __CUDACC__
#include "scgpu_kernels.cuh"

namespace scgpu
{

// One thread per on-rails agent. agent[j] = { entity, laneId, laneS bits, targetSpeed bits }, look[j] =
// TrafficAgent::lookAheadDist. Reads the agent's localPos from the Transform records, writes localPos.xz and
// localRot = (0, yaw, 0) back (sc_traffic_ai.cpp:447-457) and stamps the instance dirty for the next update.
__global__ void __launch_bounds__(kBlock) k_traffic_advance(SceneArrays a, LaneGraphView g, TrafficStepParams st, uint32_t n,
                                                            uint4* __restrict__ agent, float* __restrict__ look,
                                                            const float* __restrict__ obstacleBrake,
                                                            const uint8_t* __restrict__ skip, uint32_t stamp,
                                                            uint32_t* __restrict__ movedCount)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  bool moved = false;
  if (j < n && !(skip && skip[j]))
  {
    const uint4 ag = agent[j];
    const uint32_t s = find_slot(a, ag.x);  // ForEach<TrafficAgent, TrafficVehicle, Transform>: needs a Transform
    if (s != kNone)
    {
      const float4 r0 = a.rec[0][s];
      float pos[3] = { r0.x, r0.y, r0.z };
      float yaw = 0.0f;
      uint32_t lane = ag.y;
      float laneS = __uint_as_float(ag.z), speed = __uint_as_float(ag.w), la = look[j];
      moved = traffic_agent_on_rails(g, st, obstacleBrake ? obstacleBrake[j] : 0.0f, lane, laneS, speed, la, pos, yaw);
      agent[j] = make_uint4(ag.x, lane, __float_as_uint(laneS), __float_as_uint(speed));
      look[j] = la;
      if (moved)
      {
        a.rec[0][s] = make_float4(pos[0], pos[1], pos[2], 0.0f);
        *reinterpret_cast<float2*>(a.rec[1] + s) = make_float2(yaw, 0.0f);
        uint32_t* fw = reinterpret_cast<uint32_t*>(a.rec[3] + s) + 3;
        *fw = (*fw & 0xFFu) | (stamp << kStampShift);
      }
    }
  }
  const int m = __syncthreads_count(moved ? 1 : 0);
  if (threadIdx.x == 0 && m) atomicAdd(movedCount, (uint32_t)m);
}

// LaneSegment::active for n segments (removeSector / re-activation, sc_traffic_lanes.cpp:171-183, 224-236)
__global__ void __launch_bounds__(kBlock) k_lane_set_active(uint4* __restrict__ segNodes, uint32_t n,
                                                            const uint32_t* __restrict__ segment, const uint8_t* __restrict__ active)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  segNodes[segment[j]].z = active[j] ? 1u : 0u;
}

// local TRS of n entities (Transform::localPos/localRot/localScale), zeros for unknown handles
__global__ void __launch_bounds__(kBlock) k_gather_local(SceneArrays a, uint32_t n, const uint32_t* __restrict__ entity,
                                                         float* __restrict__ out9, uint32_t* __restrict__ missing)
{
  const uint32_t j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  float* o = out9 + (size_t)j * 9;
  const uint32_t s = find_slot(a, entity[j]);
  if (s == kNone)
  {
    for (int k = 0; k < 9; ++k) o[k] = 0.0f;
    atomicAdd(missing, 1u);
    return;
  }
  const float4 r0 = a.rec[0][s], r1 = a.rec[1][s];
  o[0] = r0.x; o[1] = r0.y; o[2] = r0.z; o[3] = r0.w; o[4] = r1.x; o[5] = r1.y; o[6] = r1.z; o[7] = r1.w;
  o[8] = reinterpret_cast<const float*>(a.rec[2] + s)[0];
}

}  // namespace scgpu
#endif  // __CUDACC__
