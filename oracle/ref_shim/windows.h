// Minimal <windows.h> stand-in for src/core/src/sc_time.cpp:7-8,20-26,41-43 (QueryPerformance*).
// Test infrastructure only (oracle build of the reference).
#pragma once
#include <cstdint>
#include <time.h>

typedef union _LARGE_INTEGER { long long QuadPart; } LARGE_INTEGER;

inline int QueryPerformanceFrequency(LARGE_INTEGER* li) { li->QuadPart = 1000000000LL; return 1; }
inline int QueryPerformanceCounter(LARGE_INTEGER* li)
{
  timespec ts{};
  clock_gettime(CLOCK_MONOTONIC, &ts);
  li->QuadPart = (long long)ts.tv_sec * 1000000000LL + (long long)ts.tv_nsec;
  return 1;
}
