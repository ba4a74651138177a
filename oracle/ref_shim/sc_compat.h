// Forced-include portability header for building the UNMODIFIED reference sources
// (/root/reference, Windows/MSVC code) headless on Linux/glibc with g++.
// Test infrastructure only: nothing under oracle/ is linked into the product library.
//
// Provides the three MSVC CRT calls the hot-path translation units use:
//   strncpy_s            (src/core/src/sc_ecs.cpp:64)
//   _aligned_malloc/free (src/core/src/sc_memory.cpp:11-98)
#pragma once
#ifdef __cplusplus
#include <cstddef>
#include <cstdlib>
#include <cstring>

inline int strncpy_s(char* dst, std::size_t dstSize, const char* src, std::size_t count)
{
  if (!dst || dstSize == 0) return 22;
  if (!src) { dst[0] = '\0'; return 22; }
  std::size_t n = 0;
  while (n < count && n + 1 < dstSize && src[n] != '\0') { dst[n] = src[n]; ++n; }
  dst[n] = '\0';
  return 0;
}

inline void* _aligned_malloc(std::size_t size, std::size_t align)
{
  void* p = nullptr;
  if (align < sizeof(void*)) align = sizeof(void*);
  if (posix_memalign(&p, align, size ? size : align) != 0) return nullptr;
  return p;
}

inline void _aligned_free(void* p) { std::free(p); }
#endif
