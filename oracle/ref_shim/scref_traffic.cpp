// scref traffic — C driver around the reference's UNMODIFIED traffic sources (test infrastructure only):
//   sc::TrafficAISystem    src/engine/traffic/sc_traffic_ai.cpp:165-487
//   sc::TrafficLaneGraph   src/engine/traffic/sc_traffic_lanes.cpp
// compiled where they lie by oracle/Makefile. TrafficLaneGraph::addNode / addSegment are private (the engine only
// builds the procedural two-road cross per sector); the tests also need arbitrary graphs, so this one translation
// unit sees the class with its private section opened. Layout and code of the class are untouched.
#define private public
#include "sc_traffic_lanes.h"
#undef private
#include "sc_traffic_ai.h"
#include "sc_debug_draw.h"

#include "scref_api.h"
#include "scref_internal.h"

#include <cmath>
#include <cstring>

// Link-only stand-ins: sc_traffic_ai.cpp references these, the oracle always runs with TrafficAIState::physics ==
// nullptr and without a DebugDraw, so none of them is ever called.
namespace sc
{
  bool PhysicsWorld::getBodyTransform(PhysicsBodyHandle, float*, float*) const { return false; }
  bool PhysicsWorld::isBodyActive(PhysicsBodyHandle) const { return false; }
  void PhysicsWorld::activateBody(PhysicsBodyHandle) {}
  RaycastHit PhysicsWorld::raycast(const float*, const float*, float, uint32_t) const { return RaycastHit{}; }
  void DebugDraw::addLine(const float*, const float*, const float*) {}
}

struct ScRefLanes
{
  sc::TrafficLaneGraph graph;
};

namespace
{
  sc::Entity ent(uint32_t v) { sc::Entity e{}; e.value = v; return e; }
}

extern "C" {

ScRefLanes* screfLanesCreate(float laneWidth, float speedLimit)
{
  ScRefLanes* l = new ScRefLanes();
  l->graph.setLaneWidth(laneWidth);
  l->graph.setSpeedLimit(speedLimit);
  return l;
}

void screfLanesDestroy(ScRefLanes* l) { delete l; }

uint32_t screfLanesAddNode(ScRefLanes* l, const float* pos3, const float* dir3, float speedLimit)
{
  return l->graph.addNode(pos3, dir3, speedLimit);
}

uint32_t screfLanesAddSegment(ScRefLanes* l, uint32_t a, uint32_t b, const float* dir3, int32_t ownerX, int32_t ownerZ)
{
  sc::SectorCoord c{};
  c.x = ownerX;
  c.z = ownerZ;
  return l->graph.addSegment(a, b, dir3, c);
}

void screfLanesBuildSector(ScRefLanes* l, int32_t x, int32_t z, float sectorSize)
{
  sc::SectorCoord c{};
  c.x = x;
  c.z = z;
  // WorldPartition::sectorBounds: the sector's square on the ground plane
  sc::AABB b{};
  b.min = sc::Vec3{ (float)x * sectorSize, 0.0f, (float)z * sectorSize };
  b.max = sc::Vec3{ (float)(x + 1) * sectorSize, 0.0f, (float)(z + 1) * sectorSize };
  l->graph.buildProceduralForSector(c, b, 0u);
}

void screfLanesRemoveSector(ScRefLanes* l, int32_t x, int32_t z)
{
  sc::SectorCoord c{};
  c.x = x;
  c.z = z;
  l->graph.removeSector(c);
}

void screfLanesSetActive(ScRefLanes* l, uint32_t segment, int active)
{
  if (segment < l->graph.m_segments.size()) l->graph.m_segments[segment].active = active != 0;
}

void screfLanesCounts(ScRefLanes* l, uint32_t* nNodes, uint32_t* nSegs, uint32_t* nConn)
{
  uint32_t c = 0;
  for (const sc::LaneNode& n : l->graph.m_nodes) c += (uint32_t)n.connections.size();
  *nNodes = (uint32_t)l->graph.m_nodes.size();
  *nSegs = (uint32_t)l->graph.m_segments.size();
  *nConn = c;
}

void screfLanesExport(ScRefLanes* l, float* nodePos3, float* nodeSpeed, uint32_t* connOffset, uint32_t* conn,
                      uint32_t* segNodes2, float* segDir3, float* segLen, uint8_t* segActive, float* defaultSpeed)
{
  uint32_t c = 0, i = 0;
  for (const sc::LaneNode& n : l->graph.m_nodes)
  {
    std::memcpy(nodePos3 + 3 * i, n.pos, 12);
    nodeSpeed[i] = n.speedLimit;
    connOffset[i] = c;
    for (uint32_t s : n.connections) conn[c++] = s;
    ++i;
  }
  connOffset[i] = c;
  i = 0;
  for (const sc::LaneSegment& s : l->graph.m_segments)
  {
    segNodes2[2 * i] = s.startNode;
    segNodes2[2 * i + 1] = s.endNode;
    std::memcpy(segDir3 + 3 * i, s.dir, 12);
    segLen[i] = s.length;
    segActive[i] = s.active ? 1 : 0;
    ++i;
  }
  *defaultSpeed = l->graph.speedLimit();
}

int screfLaneAdvance(ScRefLanes* l, uint32_t* laneId, float* s, float distance, float* outPos3, float* outDir3)
{
  return l->graph.advanceAlongLane(*laneId, *s, distance, outPos3, outDir3) ? 1 : 0;
}

uint32_t screfLaneQueryNearest(ScRefLanes* l, const float* pos3, float* outS)
{
  const sc::LaneQuery q = l->graph.queryNearestLane(pos3);
  *outS = q.s;
  return q.laneId;
}

// TrafficAgent + TrafficVehicle{mode = OnRails} on existing entities (what sc_traffic_spawner.cpp:283-316 adds)
void screfTrafficAddAgents(ScRefWorld* w, uint32_t n, const uint32_t* entity, const uint32_t* laneId, const float* laneS,
                           const float* targetSpeed, const float* lookAhead)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    const sc::Entity e = ent(entity[i]);
    sc::TrafficAgent& a = w->world.add<sc::TrafficAgent>(e);
    a.laneId = laneId[i];
    a.laneS = laneS[i];
    a.targetSpeed = targetSpeed[i];
    a.lookAheadDist = lookAhead[i];
    sc::TrafficVehicle& v = w->world.add<sc::TrafficVehicle>(e);
    v.mode = sc::TrafficSimMode::OnRails;
  }
}

// pickPlayerVehicle (sc_traffic_ai.cpp:77-103) needs one: without it the system returns before touching any agent
void screfTrafficSetPlayer(ScRefWorld* w, uint32_t entity) { w->world.add<sc::PlayerVehicle>(ent(entity)); }

void screfRunTrafficAI(ScRefWorld* w, ScRefLanes* l, float dt, int useDebug, float lookAheadDist, float speedMultiplier)
{
  sc::TrafficDebugState dbg{};
  dbg.lookAheadDist = lookAheadDist;
  dbg.speedMultiplier = speedMultiplier;
  sc::TrafficAIState st{};
  st.lanes = &l->graph;
  st.debug = useDebug ? &dbg : nullptr;
  st.physics = nullptr;
  st.streaming = nullptr;
  sc::TrafficAISystem(w->world, dt, &st);
}

void screfTrafficReadAgents(ScRefWorld* w, uint32_t n, const uint32_t* entity, uint32_t* laneId, float* laneS,
                            float* targetSpeed, float* lookAhead)
{
  for (uint32_t i = 0; i < n; ++i)
  {
    const sc::TrafficAgent* a = w->world.get<sc::TrafficAgent>(ent(entity[i]));
    laneId[i] = a ? a->laneId : 0xFFFFFFFFu;
    laneS[i] = a ? a->laneS : 0.0f;
    targetSpeed[i] = a ? a->targetSpeed : 0.0f;
    lookAhead[i] = a ? a->lookAheadDist : 0.0f;
  }
}

float screfExpf(float x) { return std::exp(x); }
float screfAtanf(float x) { return std::atan(x); }
float screfAtan2f(float y, float x) { return std::atan2(y, x); }

}  // extern "C"
