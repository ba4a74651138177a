// State behind the opaque ScRefWorld handle (test infrastructure only; shared by scref_api.cpp and dropin_test.cpp).
#pragma once
#include "sc_ecs.h"
#include "sc_world_partition.h"

struct ScRefWorld
{
  sc::World world;
  sc::CullingState culling{};
  sc::RenderPrepStreamingState renderPrep{};
  sc::WorldStreamingState* streaming = nullptr;  // heap: holds a WorldPartition
  sc::CameraSystemState camera{};
  sc::SpawnerState spawner{};
};
