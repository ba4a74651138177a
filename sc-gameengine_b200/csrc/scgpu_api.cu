// scgpu_api.cu — the C ABI of include/scgpu.h over the kernels in scgpu_kernels.cuh.
//
// Host-side state is deliberately small: an entity mirror (dense handles + sparse index, the two arrays of the
// reference's ComponentPool<Transform>, src/core/include/sc_ecs.h:199-277) so that despawns replay the
// reference's swap-with-last ORDER exactly, the slot layout (scgpu_layout.h: which device slot a Transform lives in,
// independent of that order), plus device buffer bookkeeping. All arithmetic runs on the GPU.
#include "../../include/scgpu.h"
#include "scgpu_kernels.cuh"
#include "scgpu_draws.cuh"
#include "scgpu_peer.cuh"
#include "scgpu_traffic.cuh"
#include "scgpu_pool.h"
#include "scgpu_layout.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

using namespace scgpu;

namespace
{

thread_local std::string g_createError;

struct DeviceBuffer
{
  void* ptr = nullptr;
  size_t bytes = 0;
};

// ---- NCCL, loaded lazily so that libscgpu.so has no link-time dependency on it -------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint32 = 3 };  // ncclDataType_t: ncclInt8=0, ncclUint8=1, ncclInt32=2, ncclUint32=3
struct NcclApi
{
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool loadNccl(std::string& err)
{
  if (g_nccl.lib) return true;
  // prefer a copy already in the process (torch's bundled NCCL), then the system library
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define SC_SYM(field, name) \
  *(void**)(&g_nccl.field) = dlsym(lib, name); \
  if (!g_nccl.field) { err = std::string("libnccl lacks ") + name; return false; }
  SC_SYM(GetUniqueId, "ncclGetUniqueId")
  SC_SYM(CommInitRank, "ncclCommInitRank")
  SC_SYM(CommDestroy, "ncclCommDestroy")
  SC_SYM(AllGather, "ncclAllGather")
  SC_SYM(Broadcast, "ncclBroadcast")
  SC_SYM(Send, "ncclSend")
  SC_SYM(Recv, "ncclRecv")
  SC_SYM(GroupStart, "ncclGroupStart")
  SC_SYM(GroupEnd, "ncclGroupEnd")
  SC_SYM(GetErrorString, "ncclGetErrorString")
#undef SC_SYM
  g_nccl.lib = lib;
  return true;
}

}  // namespace

struct ScGpuScene
{
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  uint32_t capacity = 0;     // slots
  uint32_t capacityPad = 0;  // rounded up to kTile
  uint32_t sparseSize = 0;
  uint32_t maxViews = 1;
  uint32_t nViews = 0;
  uint32_t count = 0;        // live Transforms (size of the reference's pool)
  uint32_t frame = 1;        // id of the NEXT update; instances dirtied now carry this stamp
  bool topologyDirty = false;
  uint32_t builtExtent = 0;  // the extent k_build_windows last saw
  bool anyParentEver = false;
  bool forceAllDirty = false;
  bool anyDirty = false;     // some call since the last transforming update may have dirtied a Transform
  bool updatedOnce = false;
  bool poisoned = false;     // a device step failed after the host mirror was committed: the two no longer agree
  uint32_t lastUpdateFlags = 0;
  uint32_t lastNumTiles = 0;
  uint32_t lastExtent = 0;
  bool culledListsValid = false;
  SlotLayout layout;         // device slots: extent, holes (scgpu_layout.h)

  SceneArrays a{};
  // hierarchy windows (k_build_windows -> k_scan_tiles -> k_flatten_windows, on topology changes only)
  uint32_t* slotInfo = nullptr;      // per slot: depth + parent lane inside its window, external / unreachable flags
  uint16_t* winLocal = nullptr;      // [tiles][kMaxWin+1] window starts relative to the tile
  uint32_t* tileWinCount = nullptr;  // [tiles]
  uint32_t* tileWinBase = nullptr;   // [tiles+1] exclusive scan of tileWinCount, total at the end
  uint32_t* winList = nullptr;       // [total+1] absolute start slot of every window (bit 31: generic path)
  uint32_t* slowList = nullptr;      // windows k_update_win hands to k_update_win_slow (start | len << 24)
  uint32_t numSMs = 148;
  uint32_t* visBits = nullptr;   // [maxViews + 1][bitWords]: visible per (view, pool rank); last plane: culling candidates
  uint32_t bitWords = 0;
  uint32_t* acc = nullptr;       // kAccWords frame accumulators + the window queue, zero between frames
  uint32_t* totals = nullptr;    // frame totals written by k_compact: [0..nViews) visible, [nViews] candidates, [kMaxViews+1] recomputed
  uint32_t* chunkCounts = nullptr;  // [2 parities][kMaxViews + 1][chunkStride]: set bits per chunk of kChunkRanks ranks
  uint32_t chunkStride = 0;
  uint32_t chunkParity = 0;         // parity the NEXT update's kernels count into (the other one is all zero by then)
  uint32_t* visEntity[kMaxViews] = {};
  uint32_t* visSlot[kMaxViews] = {};
  uint32_t* culledEntity[kMaxViews] = {};
  uint32_t maxTiles = 0;

  ViewPlanes planes{};
  uint32_t* hTotals = nullptr;  // pinned [kTotalsWords]
  cudaEvent_t evDone = nullptr;
  // ring of CUDA event pairs around the fused kernel / the whole update, so that a benchmark can read the
  // per-launch device times of many asynchronous updates after a single synchronise
  static constexpr uint32_t kTimingRing = 256;
  cudaEvent_t evK0[kTimingRing] = {}, evK1[kTimingRing] = {}, evU0[kTimingRing] = {}, evU1[kTimingRing] = {};
  uint32_t timedUpdates = 0;
  bool timings = false;
  uint32_t timingEvery = 1, updateSerial = 0;

  DeviceBuffer staging;   // uploads
  DeviceBuffer scratch;   // read-back gathers / draw items
  DeviceBuffer drawItems;
  DeviceBuffer sortWork;     // sorted draws: keys / positions / flags / scan / cub temp
  DeviceBuffer sortedDraws;  // ScGpuDrawItem[kept] in (pipeline, material, mesh) order
  DeviceBuffer drawRuns;     // ScGpuDrawRun[runs]
  DeviceBuffer matPipe;      // material -> pipeline table of the last scgpuBuildSortedDraws
  uint32_t* dSortCounters = nullptr;  // kept, nRuns
  uint32_t hSortCounters[2] = {0, 0};
  bool sortedValid = false;

  // traffic on rails (scgpu_traffic.cuh): lane graph records and the agent set
  DeviceBuffer laneNodePos, laneNodeConn, laneConn, laneSegDirLen, laneSegNodes;
  DeviceBuffer trafficAgents, trafficLook, trafficInputs;
  uint32_t nLaneNodes = 0, nLaneSegs = 0, nTrafficAgents = 0;
  float laneDefaultSpeed = 12.0f;
  bool lanesSet = false;
  uint32_t* dTrafficMoved = nullptr;

  std::vector<uint32_t> hEntity;  // dense handles (ComponentPool::m_denseEntities)
  std::vector<uint32_t> hSparse;  // index -> dense index + 1 (ComponentPool::m_sparse)
  std::vector<uint32_t> hSlotOf;  // index -> device slot of the Transform it owns
  std::vector<uint32_t> hSlotScratch;  // slots of the current spawn batch / of the victims of the current despawn batch
  std::vector<PoolMove> hMoves;   // results and work arrays of the last despawn batch (kept: no fresh pages per frame)
  std::vector<uint32_t> hRemoved;
  PoolScratch hPoolScratch;
  // pinned staging ring for large uploads from pageable memory (uploadSegs): two chunks per host thread
  static constexpr uint32_t kUpSlots = 8;
  static constexpr size_t kUpChunk = 2u << 20, kUpStagedMin = 1u << 20;
  struct UpPiece { char* dst; const char* src; size_t len; };
  std::vector<UpPiece> upPieces;
  void* upBuf[kUpSlots] = {};
  cudaEvent_t upEv[kUpSlots] = {};
  bool upPending[kUpSlots] = {};
  uint32_t hostThreads = 1;       // host threads for the pool replay and for staged uploads (SCGPU_HOST_THREADS)
  HostWorkers* workers = nullptr; // hostThreads - 1 helper threads that live as long as the context

  uint64_t launches = 0;
  std::string err;

  // multi-GPU
  ncclComm_t comm = nullptr;
  uint32_t nRanks = 1, rank = 0;
  uint32_t* dAllCounts = nullptr;   // [nRanks][maxViews+1]
  uint32_t* hAllCounts = nullptr;   // pinned
  uint32_t* gathered[kMaxViews] = {};
  // peer-memory gather (scgpu_peer.cuh)
  PeerBox peerBox{};            // mailbox in the peer root's memory
  bool peerEnabled = false;
  bool peerMapped = false;      // base is a cudaIpc mapping (not the root)
  uint32_t peerRoot = 0, peerSeq = 0;
  uint32_t* dPeerState = nullptr;  // [0] ticket counter, [1], [2] error word of the gathers with even / odd sequence number
  bool lastGatherPeer = false;
  bool peerPublished = false;   // the last update stored its lists into the root's mailbox (peer gather enabled before it)
  size_t gatheredCap = 0;
  bool gatheredValid = false;
};

namespace
{

bool fail(ScGpuScene* c, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_createError = buf;
  return false;
}

#define SC_CUDA(c, call)                                                                         \
  do                                                                                             \
  {                                                                                              \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return (int)fail((c), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// for device steps that follow a committed change of the host mirror: a failure leaves the two out of step for good
#define SC_CUDA_P(c, call)                                                                       \
  do                                                                                             \
  {                                                                                              \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
    {                                                                                            \
      fail((c), "%s failed: %s (%s:%d); the context is unusable from here on", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      (c)->poisoned = true;                                                                      \
      return 0;                                                                                  \
    }                                                                                            \
  } while (0)

#define SC_NCCL(c, call)                                                                          \
  do                                                                                              \
  {                                                                                               \
    int e__ = (call);                                                                             \
    if (e__ != ncclSuccess)                                                                       \
      return (int)fail((c), "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(e__) : "?"); \
  } while (0)

bool enter(ScGpuScene* c)
{
  if (!c) return false;
  if (c->poisoned) return false;  // the message of the failure that poisoned the context stays in place
  c->err.clear();
  if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, "cudaSetDevice(%d) failed", c->device);
  return true;
}

int ensure(ScGpuScene* c, DeviceBuffer& b, size_t bytes)
{
  if (b.bytes >= bytes) return 1;
  if (b.ptr)
  {
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
    SC_CUDA(c, cudaFree(b.ptr));
    b.ptr = nullptr;
    b.bytes = 0;
  }
  size_t want = std::max(bytes, (size_t)1 << 20);
  want = (want + 255) & ~(size_t)255;
  SC_CUDA(c, cudaMalloc(&b.ptr, want));
  b.bytes = want;
  return 1;
}

inline uint32_t blocksFor(uint64_t n) { return (uint32_t)((n + kBlock - 1) / kBlock); }

// Launch with programmatic stream serialisation: the kernel may be scheduled while its predecessor in the stream still
// runs; it calls pdl_wait() before touching anything the predecessor writes (scgpu_kernels.cuh).
int g_pdlLevel = 2;  // experiment knob (SCGPU_PDL): 0 = plain launches, 1 = only k_update_win_slow chained, 2 = whole frame chained
template <typename... KArgs, typename... Args>
cudaError_t launchPdl(void (*kern)(KArgs...), uint32_t grid, uint32_t block, cudaStream_t st, Args&&... args)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  const bool isSlow = block == kWinBlock;
  cfg.numAttrs = (g_pdlLevel >= 2 || (g_pdlLevel == 1 && isSlow)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

template <typename T>
int devAlloc(ScGpuScene* c, T** p, size_t n, bool zero)
{
  SC_CUDA(c, cudaMalloc((void**)p, std::max(n, (size_t)1) * sizeof(T)));
  if (zero) SC_CUDA(c, cudaMemsetAsync(*p, 0, std::max(n, (size_t)1) * sizeof(T), c->stream));
  return 1;
}

// frustumFromViewProj, src/engine/world/sc_world_partition.cpp:1071-1103. O(1) per view, computed on the host
// in plain IEEE float (this file is compiled with -ffp-contract=off and without -mfma): rows of the column-major
// matrix, planes r3 +- r0, r3 +- r1, r3 +- r2, each scaled by 1/sqrt(a^2+b^2+c^2) when that exceeds 1e-8.
void planesFromViewProj(const float* m, float4* out)
{
  const float r0[4] = { m[0], m[4], m[8], m[12] };
  const float r1[4] = { m[1], m[5], m[9], m[13] };
  const float r2[4] = { m[2], m[6], m[10], m[14] };
  const float r3[4] = { m[3], m[7], m[11], m[15] };
  const float* rows[3] = { r0, r1, r2 };
  for (int p = 0; p < 6; ++p)
  {
    const float* r = rows[p >> 1];
    const bool plus = (p & 1) == 0;
    volatile float a = plus ? r3[0] + r[0] : r3[0] - r[0];
    volatile float b = plus ? r3[1] + r[1] : r3[1] - r[1];
    volatile float cc = plus ? r3[2] + r[2] : r3[2] - r[2];
    volatile float d = plus ? r3[3] + r[3] : r3[3] - r[3];
    float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
    volatile float aa = a * a, bb = b * b, c2 = cc * cc;
    volatile float s1 = aa + bb;
    volatile float lenSq = s1 + c2;
    if (lenSq > 1e-8f)
    {
      volatile float invLen = 1.0f / sqrtf(lenSq);
      pl.x = a * invLen; pl.y = b * invLen; pl.z = cc * invLen; pl.w = d * invLen;
    }
    out[p] = pl;
  }
}

// ViewPlanes::slackK: 2^-19 x the largest magnitude of any normal component in use. The window kernel's fused
// pre-test needs an upper bound of |n . o| from |o|_1 alone; +Inf / NaN planes give +Inf / NaN, which switches the
// pre-test off (every comparison fails) and leaves the exact tests.
void refreshPlaneSlack(ViewPlanes& vp, uint32_t nViews)
{
  float m = 0.f;
  for (uint32_t v = 0; v < nViews; ++v)
    for (int p = 0; p < 6; ++p)
    {
      const float4 pl = vp.planes[v][p];
      const float a = fabsf(pl.x), b = fabsf(pl.y), cc = fabsf(pl.z);
      if (!(a <= m)) m = a;  // written so that NaN sticks
      if (!(b <= m)) m = b;
      if (!(cc <= m)) m = cc;
    }
  vp.slackK = m * 0x1p-19f;
}

void freeAll(ScGpuScene* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (int k = 0; k < 4; ++k) { cudaFree(c->a.rec[k]); cudaFree(c->a.world[k]); }
  cudaFree(c->a.parent); cudaFree(c->a.parentSlot); cudaFree(c->a.entity); cudaFree(c->a.meshMat); cudaFree(c->a.sparse);
  cudaFree(c->slotInfo); cudaFree(c->winLocal); cudaFree(c->tileWinCount); cudaFree(c->tileWinBase); cudaFree(c->winList); cudaFree(c->slowList); cudaFree(c->visBits); cudaFree(c->acc); cudaFree(c->chunkCounts); cudaFree(c->totals);
  cudaFree(c->a.rank); cudaFree(c->a.perm); cudaFree(c->a.tileDirty);
  for (uint32_t v = 0; v < kMaxViews; ++v)
  {
    cudaFree(c->visEntity[v]); cudaFree(c->visSlot[v]); cudaFree(c->culledEntity[v]); cudaFree(c->gathered[v]);
  }
  cudaFree(c->staging.ptr); cudaFree(c->scratch.ptr); cudaFree(c->drawItems.ptr);
  cudaFree(c->sortWork.ptr); cudaFree(c->sortedDraws.ptr); cudaFree(c->drawRuns.ptr); cudaFree(c->matPipe.ptr); cudaFree(c->dSortCounters);
  cudaFree(c->dAllCounts);
  cudaFree(c->laneNodePos.ptr); cudaFree(c->laneNodeConn.ptr); cudaFree(c->laneConn.ptr); cudaFree(c->laneSegDirLen.ptr);
  cudaFree(c->laneSegNodes.ptr); cudaFree(c->trafficAgents.ptr); cudaFree(c->trafficLook.ptr); cudaFree(c->trafficInputs.ptr);
  cudaFree(c->dTrafficMoved);
  if (c->peerBox.base) { if (c->peerMapped) cudaIpcCloseMemHandle(c->peerBox.base); else cudaFree(c->peerBox.base); }
  cudaFree(c->dPeerState);
  if (c->hTotals) cudaFreeHost(c->hTotals);
  for (uint32_t i = 0; i < ScGpuScene::kUpSlots; ++i)
  {
    if (c->upEv[i]) cudaEventDestroy(c->upEv[i]);
  }
  if (c->upBuf[0]) cudaFreeHost(c->upBuf[0]);
  if (c->hAllCounts) cudaFreeHost(c->hAllCounts);
  if (c->evDone) cudaEventDestroy(c->evDone);
  for (uint32_t i = 0; i < ScGpuScene::kTimingRing; ++i)
  {
    if (c->evK0[i]) cudaEventDestroy(c->evK0[i]);
    if (c->evK1[i]) cudaEventDestroy(c->evK1[i]);
    if (c->evU0[i]) cudaEventDestroy(c->evU0[i]);
    if (c->evU1[i]) cudaEventDestroy(c->evU1[i]);
  }
  if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
  delete c->workers;
  delete c;
}

// k_update stages its planes in more than the default 48 KB of shared memory: opt in once per instantiation
template <int V>
cudaError_t optInSmem()
{
  cudaError_t e = cudaFuncSetAttribute(k_update_flat<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdateSmemFlat);
  static_assert(kUpdateSmemWin <= 48 * 1024, "k_update_win keeps its shared memory static");
  return e;
}

int createImpl(ScGpuScene* c, const ScGpuSceneDesc* d)
{
  int nDev = 0;
  cudaError_t e = cudaGetDeviceCount(&nDev);
  if (e != cudaSuccess || nDev == 0)
    return (int)fail(c, "no CUDA device available (%s); scgpu has no CPU fallback", cudaGetErrorString(e));
  if (d->device < 0 || d->device >= nDev) return (int)fail(c, "device ordinal %d out of range (0..%d)", d->device, nDev - 1);
  c->device = d->device;
  SC_CUDA(c, cudaSetDevice(c->device));
  cudaDeviceProp prop{};
  SC_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
  c->numSMs = (uint32_t)prop.multiProcessorCount;
  if (prop.major < 10)
    return (int)fail(c, "device %d is sm_%d%d; libscgpu is built for sm_100a (B200) only", c->device, prop.major, prop.minor);
  SC_CUDA(c, optInSmem<1>()); SC_CUDA(c, optInSmem<2>()); SC_CUDA(c, optInSmem<3>()); SC_CUDA(c, optInSmem<4>());
  SC_CUDA(c, optInSmem<5>()); SC_CUDA(c, optInSmem<6>()); SC_CUDA(c, optInSmem<7>()); SC_CUDA(c, optInSmem<8>());
  if (d->stream) { c->stream = (cudaStream_t)d->stream; c->ownStream = false; }
  else { SC_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->ownStream = true; }

  c->capacity = d->max_instances;
  c->capacityPad = ((d->max_instances + kTile - 1) / kTile) * kTile;
  if (c->capacityPad == 0) c->capacityPad = kTile;
  c->sparseSize = d->max_entity_index ? d->max_entity_index : (1u << 24);
  if (c->sparseSize > (1u << 24)) c->sparseSize = 1u << 24;
  c->maxViews = d->max_views;
  c->maxTiles = c->capacityPad / kTile;

  const size_t n = c->capacityPad;
  for (int k = 0; k < 4; ++k)
  {
    if (!devAlloc(c, &c->a.rec[k], n, true)) return 0;
    if (!devAlloc(c, &c->a.world[k], n, true)) return 0;
  }
  if (!devAlloc(c, &c->a.parent, n, false)) return 0;
  if (!devAlloc(c, &c->a.parentSlot, n, false)) return 0;
  if (!devAlloc(c, &c->a.entity, n, true)) return 0;
  if (!devAlloc(c, &c->a.meshMat, n, true)) return 0;
  if (!devAlloc(c, &c->a.sparse, (size_t)c->sparseSize, true)) return 0;
  c->a.sparseSize = c->sparseSize;
  if (!devAlloc(c, &c->a.rank, n, true)) return 0;
  if (!devAlloc(c, &c->a.perm, n, true)) return 0;
  c->bitWords = (uint32_t)(n / 32);  // n is a multiple of kTile: planes are 16-byte multiples
  if (!devAlloc(c, &c->visBits, (size_t)(c->maxViews + 1) * c->bitWords, true)) return 0;
  if (!devAlloc(c, &c->acc, (size_t)kAccWords, true)) return 0;
  c->chunkStride = c->bitWords / kCompactChunkWords + 1u;
  if (!devAlloc(c, &c->chunkCounts, (size_t)2 * (kMaxViews + 1) * c->chunkStride, true)) return 0;
  if (!devAlloc(c, &c->slotInfo, n, true)) return 0;
  // every tile starts out "to be cut" (k_build_windows is incremental): byte t + 1 belongs to tile t
  if (!devAlloc(c, &c->a.tileDirty, ((size_t)c->maxTiles + 2) * 4, false)) return 0;
  SC_CUDA(c, cudaMemsetAsync(c->a.tileDirty, 1, ((size_t)c->maxTiles + 2) * 4, c->stream));
  if (!devAlloc(c, &c->winLocal, (size_t)c->maxTiles * (kMaxWin + 1), true)) return 0;
  if (!devAlloc(c, &c->tileWinCount, (size_t)c->maxTiles, true)) return 0;
  if (!devAlloc(c, &c->tileWinBase, (size_t)c->maxTiles + 1, true)) return 0;
  if (!devAlloc(c, &c->winList, (size_t)c->maxTiles * kMaxWin + 1, true)) return 0;
  if (!devAlloc(c, &c->slowList, (size_t)c->maxTiles * kMaxWin + 1, true)) return 0;
  if (!devAlloc(c, &c->totals, (size_t)kTotalsWords, true)) return 0;
  for (uint32_t v = 0; v < c->maxViews; ++v)
  {
    if (!devAlloc(c, &c->visEntity[v], n, false)) return 0;
    if (!devAlloc(c, &c->visSlot[v], n, false)) return 0;
  }
  SC_CUDA(c, cudaHostAlloc((void**)&c->hTotals, sizeof(uint32_t) * kTotalsWords, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(c->hTotals, 0, sizeof(uint32_t) * kTotalsWords);
  SC_CUDA(c, cudaEventCreateWithFlags(&c->evDone, cudaEventDisableTiming));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  c->hEntity.reserve(c->capacity);
  c->layout.reset(c->capacity);
  return 1;
}

// Large uploads from PAGEABLE caller memory go through a ring of pinned chunks filled by several host threads: every
// thread owns two chunks, copies its share of the sources into them and queues the DMA of each on the context stream
// (the driver's own pageable path is one thread deep: 7-18 GB/s measured on the box against ~50 GB/s for the link).
// A call with several arrays (scgpuSpawn: six) hands them over together so that the threads start once. The sources are
// consumed when the call returns, as for every host pointer of the ABI.
struct UpSeg
{
  void* dst;
  const void* src;
  size_t bytes;
};

int ensureUploadRing(ScGpuScene* c)
{
  if (c->upBuf[0]) return 1;
  // one pinned block for the whole ring (pinning costs ~1-2 ms per call whatever the size)
  void* block = nullptr;
  SC_CUDA(c, cudaHostAlloc(&block, ScGpuScene::kUpChunk * ScGpuScene::kUpSlots, cudaHostAllocDefault));
  for (uint32_t i = 0; i < ScGpuScene::kUpSlots; ++i)
  {
    c->upBuf[i] = (char*)block + ScGpuScene::kUpChunk * i;
    SC_CUDA(c, cudaEventCreateWithFlags(&c->upEv[i], cudaEventDisableTiming));
  }
  return 1;
}

int uploadSegs(ScGpuScene* c, const UpSeg* segs, uint32_t nSegs)
{
  typedef ScGpuScene::UpPiece Piece;
  std::vector<Piece>& pieces = c->upPieces;
  pieces.clear();
  const size_t chunk = ScGpuScene::kUpChunk;
  size_t pageable = 0;
  bool isPageable[16] = {};
  for (uint32_t i = 0; i < nSegs && i < 16u && c->hostThreads > 1u; ++i)
  {
    if (!segs[i].src || segs[i].bytes < 65536u) continue;  // small ones are cheaper through the driver
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, segs[i].src) != cudaSuccess) { (void)cudaGetLastError(); continue; }
    if (at.type == cudaMemoryTypeUnregistered) { isPageable[i] = true; pageable += segs[i].bytes; }
  }
  const bool staged = pageable >= ScGpuScene::kUpStagedMin;
  for (uint32_t i = 0; i < nSegs; ++i)
  {
    if (!segs[i].src || segs[i].bytes == 0) continue;
    if (staged && i < 16u && isPageable[i])
    {
      for (size_t off = 0; off < segs[i].bytes; off += chunk)
        pieces.push_back(Piece{(char*)segs[i].dst + off, (const char*)segs[i].src + off, std::min(chunk, segs[i].bytes - off)});
    }
    else
      SC_CUDA(c, cudaMemcpyAsync(segs[i].dst, segs[i].src, segs[i].bytes, cudaMemcpyHostToDevice, c->stream));
  }
  if (pieces.empty()) return 1;

  const uint32_t nPieces = (uint32_t)pieces.size();
  const uint32_t T = std::min(std::min(c->hostThreads, ScGpuScene::kUpSlots / 2u), nPieces);  // one share (two ring chunks) per thread
  if (!ensureUploadRing(c)) return 0;
  std::atomic<int> err{(int)cudaSuccess};
  const Piece* const pc = pieces.data();
  auto work = [&](uint32_t t) {
    cudaError_t e = cudaSetDevice(c->device);
    for (uint32_t i = t, round = 0; i < nPieces && e == cudaSuccess; i += T, ++round)
    {
      const uint32_t slot = 2u * t + (round & 1u);
      if (c->upPending[slot]) e = cudaEventSynchronize(c->upEv[slot]);  // the DMA that last read this chunk
      if (e != cudaSuccess) break;
      memcpy(c->upBuf[slot], pc[i].src, pc[i].len);
      e = cudaMemcpyAsync(pc[i].dst, c->upBuf[slot], pc[i].len, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) e = cudaEventRecord(c->upEv[slot], c->stream);
      c->upPending[slot] = true;
    }
    if (e != cudaSuccess) err.store((int)e);
  };
  if (c->workers) c->workers->run(T, work);
  else
    for (uint32_t t = 0; t < T; ++t) work(t);
  if (err.load() != (int)cudaSuccess)
    return (int)fail(c, "staged upload failed: %s", cudaGetErrorString((cudaError_t)err.load()));
  return 1;
}

int uploadTo(ScGpuScene* c, void* dst, const void* src, size_t bytes)
{
  const UpSeg seg{dst, src, bytes};
  return uploadSegs(c, &seg, 1);
}

// stamps are 24 bits wide and 0 means "never"
uint32_t stampOf(uint32_t frame) { return frame & 0xFFFFFFu; }

int waitDone(ScGpuScene* c)
{
  if (!c->updatedOnce) return (int)fail(c, "no scgpuUpdate has been issued yet");
  SC_CUDA(c, cudaEventSynchronize(c->evDone));
  return 1;
}

}  // namespace

extern "C" {

uint32_t scgpuGetApiVersion(void) { return SCGPU_API_VERSION; }

const char* scgpuLastError(const ScGpuScene* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

ScGpuScene* scgpuCreate(const ScGpuSceneDesc* desc)
{
  g_createError.clear();
  if (!desc) { fail(nullptr, "scgpuCreate: desc is NULL"); return nullptr; }
  if (desc->struct_size != sizeof(ScGpuSceneDesc)) { fail(nullptr, "scgpuCreate: struct_size %u != %zu", desc->struct_size, sizeof(ScGpuSceneDesc)); return nullptr; }
  if (desc->max_views == 0 || desc->max_views > SCGPU_MAX_VIEWS) { fail(nullptr, "scgpuCreate: max_views must be 1..%u", SCGPU_MAX_VIEWS); return nullptr; }
  if (desc->max_instances == 0 || desc->max_instances > (1u << 24)) { fail(nullptr, "scgpuCreate: max_instances must be 1..%u (24-bit entity index, sc_ecs.h:18-20)", 1u << 24); return nullptr; }
  ScGpuScene* c = new (std::nothrow) ScGpuScene();
  if (!c) { fail(nullptr, "out of host memory"); return nullptr; }
  if (const char* pl = getenv("SCGPU_PDL")) g_pdlLevel = atoi(pl);
  // test hook: id of the first update, to reach the wrap of the 24-bit dirty stamp without 16.7 M frames
  if (const char* ff = getenv("SCGPU_TEST_FIRST_FRAME")) c->frame = std::max(1ul, strtoul(ff, nullptr, 0));
  {
    // host threads for the pool bookkeeping of large despawn batches and staged uploads: SCGPU_HOST_THREADS, else half
    // the cores up to 8 (the sequential middle pass of the pool replay bounds the gain). The helpers are created once.
    const char* ht = getenv("SCGPU_HOST_THREADS");
    const unsigned hw = std::thread::hardware_concurrency();
    long want = ht ? strtol(ht, nullptr, 10) : (long)std::min(8u, std::max(1u, hw / 2u));
    c->hostThreads = (uint32_t)std::min(64l, std::max(1l, want));
    if (c->hostThreads > 1u)
    {
      c->workers = new (std::nothrow) HostWorkers(c->hostThreads - 1u);
      c->hostThreads = c->workers ? c->workers->helpers() + 1u : 1u;
    }
  }
  if (!createImpl(c, desc) || (c->hostThreads > 1u && !ensureUploadRing(c)))  // ring up front: no first-upload hiccup
  {
    g_createError = c->err;
    freeAll(c);
    return nullptr;
  }
  return c;
}

void scgpuDestroy(ScGpuScene* ctx)
{
#ifdef SCGPU_COMPACT_TIMING
  if (ctx)
  {
    // diagnostics build only: per-CTA phase timestamps of the last k_compact launch
    static unsigned long long h[1024 * 8];
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (cudaMemcpyFromSymbol(h, g_compactStamps, sizeof(h)) == cudaSuccess)
    {
      unsigned long long t0 = ~0ull;
      for (int b = 0; b < 1024; ++b) if (h[b * 8] && h[b * 8] < t0) t0 = h[b * 8];
      for (int b = 0; b < 1024; ++b)
        if (h[b * 8])
          fprintf(stderr, "compact cta %4d: start %6llu wait %6llu prefix %6llu emit %6llu end %6llu ns\n", b, h[b * 8] - t0,
                  h[b * 8 + 1] - t0, h[b * 8 + 2] - t0, h[b * 8 + 3] - t0, h[b * 8 + 4] - t0);
    }
  }
#endif
  freeAll(ctx);
}

void* scgpuGetStream(ScGpuScene* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

uint64_t scgpuKernelLaunchCount(ScGpuScene* ctx) { return ctx ? ctx->launches : 0; }

int scgpuEnableTimings(ScGpuScene* ctx, int enable)
{
  if (!enter(ctx)) return 0;
  if (enable && !ctx->evK0[0])
  {
    for (uint32_t i = 0; i < ScGpuScene::kTimingRing; ++i)
    {
      SC_CUDA(ctx, cudaEventCreate(&ctx->evK0[i]));
      SC_CUDA(ctx, cudaEventCreate(&ctx->evK1[i]));
      SC_CUDA(ctx, cudaEventCreate(&ctx->evU0[i]));
      SC_CUDA(ctx, cudaEventCreate(&ctx->evU1[i]));
    }
  }
  ctx->timings = enable != 0;
  ctx->timingEvery = enable > 1 ? (uint32_t)enable : 1u;
  ctx->updateSerial = 0;
  ctx->timedUpdates = 0;
  return 1;
}

int scgpuSynchronize(ScGpuScene* ctx)
{
  if (!enter(ctx)) return 0;
  SC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->peerEnabled && ctx->lastGatherPeer && ctx->gatheredValid && ctx->rank != ctx->peerRoot)
  {
    // a producer that could not deliver learns it here (the root learns it from its read-back calls)
    uint32_t err = 0;
    SC_CUDA(ctx, cudaMemcpy(&err, ctx->dPeerState + 1 + (ctx->peerSeq & 1u), 4, cudaMemcpyDeviceToHost));
    if (err & 1u) return (int)fail(ctx, "peer gather: the root did not release this rank's mailbox within the time limit; the last frame's lists were not delivered");
  }
  return 1;
}

// ---- deltas -------------------------------------------------------------------------------------------

static int poison(ScGpuScene* c)
{
  c->poisoned = true;
  c->err += " (host mirror already updated: the context is unusable from here on)";
  return 0;
}

// Host half of a spawn batch: World::create + add<Transform> on the pool mirror (every handle is validated first, so a
// refused batch leaves the scene untouched), then a device slot for every element (scgpu_layout.h). *slotOf == nullptr:
// the batch got the run of fresh slots starting at *slot0; else slotOf[j] (host array, valid until the next call).
// Call it after every fallible allocation: a device failure behind it poisons the context.
static int registerSpawn(ScGpuScene* c, uint32_t n, const uint32_t* entity, const uint32_t* parent, const char* who,
                         uint32_t* slot0, const uint32_t** slotOf)
{
  if ((uint64_t)c->count + n > c->capacity) return (int)fail(c, "%s: %u + %u instances exceed max_instances %u", who, c->count, n, c->capacity);
  if (c->hSparse.size() < c->sparseSize) c->hSparse.resize(c->sparseSize, 0u);
  if (c->hSlotOf.size() < c->sparseSize) c->hSlotOf.resize(c->sparseSize, 0u);
  uint32_t at = 0;
  int why = 0;
  // slots are assigned one by one when there are holes to fill or hierarchy groups to keep together and to pack into
  // windows (scgpu_layout.h); a flat batch into a pool without holes simply takes the next n slots
  const bool holes = c->layout.hasHoles() || SlotLayout::batchHasHierarchy(n, parent);
  uint32_t* sl = nullptr;
  if (holes)
  {
    if (c->hSlotScratch.size() < n) c->hSlotScratch.resize(n + n / 4);
    sl = c->hSlotScratch.data();
  }
  if (holes && c->workers && n >= 65536u)
  {
    // pool registration (random writes into the sparse table) and slot placement (a walk of the free-slot bitmap) touch
    // disjoint state: side by side on two threads. Placement cannot fail (holes + tail room >= capacity - count >= n);
    // if the registration refuses the batch the slots are simply handed back.
    c->workers->run(2u, [&](uint32_t part) {
      if (part == 0u) why = poolRegisterSpawn(c->hEntity, c->hSparse, c->count, n, entity, &at);
      else c->layout.placeBatch(n, entity, parent, sl);
    });
    if (why) c->layout.release(n, sl);
  }
  else
  {
    why = poolRegisterSpawn(c->hEntity, c->hSparse, c->count, n, entity, &at);
    if (!why && holes) c->layout.placeBatch(n, entity, parent, sl);
  }
  switch (why)
  {
    case 0: break;
    case 1: return (int)fail(c, "%s: entity[%u] is the invalid handle", who, at);
    case 2: return (int)fail(c, "%s: entity index %u >= max_entity_index %u", who, entity[at] & 0xFFFFFFu, c->sparseSize);
    default: return (int)fail(c, "%s: entity index %u already owns a Transform", who, entity[at] & 0xFFFFFFu);
  }
  c->anyDirty = true;  // new Transforms are dirty
  uint32_t* const so = c->hSlotOf.data();
  if (!holes)
  {
    const uint32_t s0 = c->layout.appendRun(n);  // cannot fail: without holes the extent is the pool size
    for (uint32_t j = 0; j < n; ++j) so[entity[j] & 0xFFFFFFu] = s0 + j;
    *slot0 = s0;
    *slotOf = nullptr;
  }
  else
  {
    poolParallelFor(c->hostThreads, n, [=](uint32_t, uint32_t b, uint32_t e) {
      for (uint32_t j = b; j < e; ++j)
      {
        if (j + 16u < e) __builtin_prefetch(so + (entity[j + 16u] & 0xFFFFFFu), 1);
        so[entity[j] & 0xFFFFFFu] = sl[j];
      }
    }, c->workers);
    *slot0 = 0;
    *slotOf = sl;
  }
  return 1;
}

int scgpuSpawn(ScGpuScene* c, uint32_t n, const uint32_t* entity, const uint32_t* parent, const float* trs9,
               const float* aabb6, const uint32_t* meshMat2, const uint32_t* flags)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!entity || !trs9) return (int)fail(c, "scgpuSpawn: entity and trs9 are required");
  const uint32_t chunkMax = 1u << 20;
  {
    const size_t m = std::min(chunkMax, n);
    if (!ensure(c, c->staging, m * (4 + 4 + 4 + 36 + 24 + 8 + 4))) return 0;  // before the mirror changes
  }
  uint32_t slot0 = 0;
  const uint32_t* slotOf = nullptr;
  if (!registerSpawn(c, n, entity, parent, "scgpuSpawn", &slot0, &slotOf)) return 0;

  for (uint32_t off = 0; off < n; off += chunkMax)
  {
    const uint32_t m = std::min(chunkMax, n - off);
    size_t bytes = 0;
    const size_t oEntity = bytes; bytes += (size_t)m * 4;
    const size_t oParent = bytes; bytes += parent ? (size_t)m * 4 : 0;
    const size_t oSlot = bytes; bytes += slotOf ? (size_t)m * 4 : 0;
    const size_t oTrs = bytes; bytes += (size_t)m * 36;
    const size_t oAabb = bytes; bytes += aabb6 ? (size_t)m * 24 : 0;
    const size_t oMm = bytes; bytes += meshMat2 ? (size_t)m * 8 : 0;
    const size_t oFlags = bytes; bytes += flags ? (size_t)m * 4 : 0;
    char* s = (char*)c->staging.ptr;
    const UpSeg segs[7] = {
      {s + oEntity, entity + off, (size_t)m * 4},
      {s + oParent, parent ? parent + off : nullptr, (size_t)m * 4},
      {s + oSlot, slotOf ? slotOf + off : nullptr, (size_t)m * 4},
      {s + oTrs, trs9 + (size_t)off * 9, (size_t)m * 36},
      {s + oAabb, aabb6 ? aabb6 + (size_t)off * 6 : nullptr, (size_t)m * 24},
      {s + oMm, meshMat2 ? meshMat2 + (size_t)off * 2 : nullptr, (size_t)m * 8},
      {s + oFlags, flags ? flags + off : nullptr, (size_t)m * 4}};
    if (!uploadSegs(c, segs, 7)) return poison(c);
    k_spawn<<<blocksFor(m), kBlock, 0, c->stream>>>(
      c->a, slot0 + off, slotOf ? (const uint32_t*)(s + oSlot) : nullptr, c->count + off, m, (const uint32_t*)(s + oEntity),
      parent ? (const uint32_t*)(s + oParent) : nullptr, (const float*)(s + oTrs), aabb6 ? (const float*)(s + oAabb) : nullptr,
      meshMat2 ? (const uint32_t*)(s + oMm) : nullptr, flags ? (const uint32_t*)(s + oFlags) : nullptr, stampOf(c->frame));
    ++c->launches;
    SC_CUDA_P(c, cudaGetLastError());
  }
  if (parent)
  {
    bool any = c->anyParentEver;
    for (uint32_t j = 0; j < n && !any; ++j) any = parent[j] != SCGPU_INVALID_ENTITY;
    if (any) { c->anyParentEver = true; c->topologyDirty = true; }
  }
  // a new Transform can turn a dangling parent handle valid, and a reused hole changes the windows around it:
  // re-resolve if any hierarchy exists
  if (c->anyParentEver) c->topologyDirty = true;
  c->count += n;
  return 1;
}

// ---- procedural sectors spawned on the device -------------------------------------------------
static SectorGen toSectorGen(const ScGpuSectorGen* g)
{
  SectorGen o{};
  o.sectorSizeMeters = g->sectorSizeMeters; o.seed = g->seed; o.propsMin = g->propsPerSectorMin; o.propsMax = g->propsPerSectorMax;
  o.includeGround = g->includeGroundPlane ? 1u : 0u;
  o.meshCube = g->meshCube; o.meshTriangle = g->meshTriangle;
  o.matUnlit = g->matUnlit; o.matChecker = g->matChecker; o.matTest = g->matTest;
  return o;
}

uint32_t scgpuSectorSpawnCount(const ScGpuSectorGen* gen, int32_t x, int32_t z)
{
  if (!gen || gen->propsPerSectorMax < gen->propsPerSectorMin) return 0;
  const SectorGen g = toSectorGen(gen);
  return sg_prop_count(g, x, z) + g.includeGround;
}

int scgpuSpawnSectors(ScGpuScene* c, const ScGpuSectorGen* gen, uint32_t nSectors, const int32_t* coordXZ, const uint32_t* entity,
                      uint32_t nEntities)
{
  if (!enter(c)) return 0;
  if (nSectors == 0) return 1;
  if (!gen || !coordXZ || !entity) return (int)fail(c, "scgpuSpawnSectors: NULL argument");
  if (gen->propsPerSectorMax < gen->propsPerSectorMin) return (int)fail(c, "scgpuSpawnSectors: propsPerSectorMax < propsPerSectorMin");
  const SectorGen g = toSectorGen(gen);
  std::vector<uint32_t> first(nSectors + 1, 0u);
  for (uint32_t k = 0; k < nSectors; ++k)
    first[k + 1] = first[k] + sg_prop_count(g, coordXZ[2 * k], coordXZ[2 * k + 1]) + g.includeGround;
  const uint32_t n = first[nSectors];
  if (n != nEntities)
    return (int)fail(c, "scgpuSpawnSectors: the sectors yield %u spawn records but %u entity handles were passed", n, nEntities);
  const size_t oCoord = 0, oFirst = (size_t)nSectors * 8, oEntity = oFirst + ((size_t)nSectors + 1) * 4;
  const size_t oSlot = oEntity + (size_t)n * 4;
  const size_t bytes = oSlot + (size_t)n * 4;
  if (!ensure(c, c->staging, bytes)) return 0;
  uint32_t slot0 = 0;
  const uint32_t* slotOf = nullptr;
  if (!registerSpawn(c, n, entity, nullptr, "scgpuSpawnSectors", &slot0, &slotOf)) return 0;
  char* s = (char*)c->staging.ptr;
  if (!uploadTo(c, s + oCoord, coordXZ, (size_t)nSectors * 8)) return poison(c);
  if (!uploadTo(c, s + oFirst, first.data(), ((size_t)nSectors + 1) * 4)) return poison(c);
  if (!uploadTo(c, s + oEntity, entity, (size_t)n * 4)) return poison(c);
  if (slotOf && !uploadTo(c, s + oSlot, slotOf, (size_t)n * 4)) return poison(c);
  k_spawn_sectors<<<nSectors, 64, 0, c->stream>>>(c->a, g, slot0, slotOf ? (const uint32_t*)(s + oSlot) : nullptr, c->count,
                                                  (const int32_t*)(s + oCoord), (const uint32_t*)(s + oFirst),
                                                  (const uint32_t*)(s + oEntity), stampOf(c->frame));
  ++c->launches;
  SC_CUDA_P(c, cudaGetLastError());
  SC_CUDA_P(c, cudaStreamSynchronize(c->stream));  // `first` is a local: the pageable upload must have left it
  if (c->anyParentEver) c->topologyDirty = true;
  c->count += n;
  return 1;
}

// ---- .scsector files ----------------------------------------------------------------------------
namespace
{
struct SectorFileInfo
{
  uint32_t version = 0;
  int32_t x = 0, z = 0;
  uint32_t count = 0;          // instances of the (last) INST chunk
  size_t payload = 0;          // byte offset of its first record
  uint32_t recordSize = 0;     // bytes per record as the reader derives it
  uint32_t meshOffset = 0;     // byte offset of mesh_id inside a record
};

// The chunk walk of sc_world::ReadSectorFile (tools/shared/world_format.cpp:178-334) over a memory image of the file.
// Returns an error text or nullptr.
const char* parseSectorFile(const void* bytes, size_t n, SectorFileInfo& o)
{
  const unsigned char* b = (const unsigned char*)bytes;
  auto u32 = [&](size_t at) { uint32_t v; memcpy(&v, b + at, 4); return v; };
  if (!bytes || n < 16) return "file shorter than its header";
  if (u32(0) != 0x54434553u) return "not a sector file (magic != 'SECT')";
  o.version = u32(4);
  o.x = (int32_t)u32(8);
  o.z = (int32_t)u32(12);
  const uint32_t kInst = (uint32_t)'I' | ((uint32_t)'N' << 8) | ((uint32_t)'S' << 16) | ((uint32_t)'T' << 24);
  size_t pos = 16;
  while (pos + 8 <= n)
  {
    const uint32_t id = u32(pos), size = u32(pos + 4);
    pos += 8;
    if (size == 0) continue;
    if (id != kInst)
    {
      pos += size;  // LANE / SPWN / COLL are consumed by content (== size for a well-formed file), the rest is skipped
      continue;
    }
    if (pos + 4 > n) return "INST chunk cut short";
    const uint32_t count = u32(pos);
    const uint32_t baseV3 = 8 + 8 + 8 + 36 + 4, baseV4 = baseV3 + 8;
    uint32_t recordSize = baseV3;
    if (count > 0 && size >= 4) recordSize = (size - 4) / count;
    const bool hasModel = o.version >= 4;
    if (count > 0 && recordSize < (hasModel ? baseV4 : baseV3)) return "INST record smaller than its fixed fields";
    if ((uint64_t)pos + 4 + (uint64_t)count * recordSize > n) return "INST chunk cut short";
    o.count = count;
    o.payload = pos + 4;
    o.recordSize = recordSize;
    o.meshOffset = hasModel ? 16u : 8u;
    pos += 4 + (size_t)count * recordSize;
  }
  return nullptr;
}
}  // namespace

int scgpuSectorFileInfo(const void* bytes, size_t nBytes, int32_t* outXZ, uint32_t* outVersion, uint32_t* outInstances)
{
  SectorFileInfo f;
  if (parseSectorFile(bytes, nBytes, f)) return 0;
  if (outXZ) { outXZ[0] = f.x; outXZ[1] = f.z; }
  if (outVersion) *outVersion = f.version;
  if (outInstances) *outInstances = f.count;
  return 1;
}

int scgpuSpawnSectorFile(ScGpuScene* c, const void* bytes, size_t nBytes, const uint32_t* entity, uint32_t nEntities,
                         const ScGpuAssetTable* assets)
{
  static_assert(sizeof(ScGpuAssetBinding) == sizeof(AssetBinding), "ScGpuAssetBinding layout");
  if (!enter(c)) return 0;
  SectorFileInfo f;
  if (const char* err = parseSectorFile(bytes, nBytes, f)) return (int)fail(c, "scgpuSpawnSectorFile: %s", err);
  if (f.count != nEntities)
    return (int)fail(c, "scgpuSpawnSectorFile: the file holds %u instances but %u entity handles were passed", f.count, nEntities);
  if (f.count == 0) return 1;
  if (!entity || !assets) return (int)fail(c, "scgpuSpawnSectorFile: NULL argument");
  if ((f.payload | f.recordSize | f.meshOffset) & 3u) return (int)fail(c, "scgpuSpawnSectorFile: INST records are not 4-byte aligned");
  if ((assets->nMeshes && !assets->meshes) || (assets->nMaterials && !assets->materials))
    return (int)fail(c, "scgpuSpawnSectorFile: asset table pointer is NULL");
  const size_t payloadBytes = (size_t)f.count * f.recordSize;
  const size_t oPay = 0, oEnt = (payloadBytes + 255) & ~(size_t)255, oMesh = oEnt + (((size_t)f.count * 4 + 255) & ~(size_t)255);
  const size_t oMat = oMesh + (((size_t)assets->nMeshes * 16 + 255) & ~(size_t)255);
  const size_t oSlot = (oMat + (size_t)assets->nMaterials * 16 + 16 + 255) & ~(size_t)255;
  const size_t total = oSlot + (size_t)f.count * 4;
  if (!ensure(c, c->staging, total)) return 0;
  uint32_t slot0 = 0;
  const uint32_t* slotOf = nullptr;
  if (!registerSpawn(c, f.count, entity, nullptr, "scgpuSpawnSectorFile", &slot0, &slotOf)) return 0;
  char* s = (char*)c->staging.ptr;
  if (!uploadTo(c, s + oPay, (const char*)bytes + f.payload, payloadBytes)) return poison(c);
  if (!uploadTo(c, s + oEnt, entity, (size_t)f.count * 4)) return poison(c);
  if (assets->nMeshes && !uploadTo(c, s + oMesh, assets->meshes, (size_t)assets->nMeshes * 16)) return poison(c);
  if (assets->nMaterials && !uploadTo(c, s + oMat, assets->materials, (size_t)assets->nMaterials * 16)) return poison(c);
  if (slotOf && !uploadTo(c, s + oSlot, slotOf, (size_t)f.count * 4)) return poison(c);
  k_spawn_sector_file<<<blocksFor(f.count), kBlock, 0, c->stream>>>(
    c->a, slot0, slotOf ? (const uint32_t*)(s + oSlot) : nullptr, c->count, f.count, (const uint32_t*)(s + oPay), f.recordSize / 4u,
    f.meshOffset / 4u, (const uint32_t*)(s + oEnt), (const AssetBinding*)(s + oMesh), assets->nMeshes, assets->defaultMesh,
    (const AssetBinding*)(s + oMat), assets->nMaterials, assets->defaultMaterial, stampOf(c->frame));
  ++c->launches;
  SC_CUDA_P(c, cudaGetLastError());
  if (c->anyParentEver) c->topologyDirty = true;
  c->count += f.count;
  return 1;
}

// ---- the editor's BuildDrawItems (editor_core.cpp:242-264) -------------------------------------
int scgpuBuildEditorDraws(ScGpuScene* c, uint32_t n, const float* trs9, const uint64_t* meshHandle, const uint64_t* materialHandle,
                          ScGpuEditorDrawItem* out, uint32_t cap, uint32_t* outCount)
{
  static_assert(sizeof(ScGpuEditorDrawItem) == 88, "ScRenderDrawItem is 88 bytes");
  if (!enter(c)) return 0;
  if (outCount) *outCount = 0;
  if (n == 0) return 1;
  if (!trs9 || !meshHandle || !materialHandle) return (int)fail(c, "scgpuBuildEditorDraws: NULL argument");
  const uint32_t blocks = blocksFor(n);
  const size_t oTrs = 0, oMesh = ((size_t)n * 36 + 255) & ~(size_t)255, oMat = oMesh + (((size_t)n * 8 + 255) & ~(size_t)255);
  const size_t oCnt = oMat + (((size_t)n * 8 + 255) & ~(size_t)255), oOff = oCnt + (((size_t)blocks * 4 + 255) & ~(size_t)255);
  const size_t oTot = oOff + (((size_t)blocks * 4 + 255) & ~(size_t)255), oOut = oTot + 256;
  if (!ensure(c, c->scratch, oOut + (size_t)n * 88)) return 0;
  char* s = (char*)c->scratch.ptr;
  if (!uploadTo(c, s + oTrs, trs9, (size_t)n * 36)) return 0;
  if (!uploadTo(c, s + oMesh, meshHandle, (size_t)n * 8)) return 0;
  if (!uploadTo(c, s + oMat, materialHandle, (size_t)n * 8)) return 0;
  k_editor_count<<<blocks, kBlock, 0, c->stream>>>((const uint64_t*)(s + oMesh), (const uint64_t*)(s + oMat), n, (uint32_t*)(s + oCnt));
  k_scan_tiles<<<1, 1024, 0, c->stream>>>((const uint32_t*)(s + oCnt), (uint32_t*)(s + oOff), (uint32_t*)(s + oTot), blocks);
  k_editor_write<<<blocks, kBlock, 0, c->stream>>>((const float*)(s + oTrs), (const uint64_t*)(s + oMesh), (const uint64_t*)(s + oMat), n,
                                                   (const uint32_t*)(s + oOff), (uint32_t*)(s + oOut));
  c->launches += 3;
  SC_CUDA(c, cudaGetLastError());
  uint32_t total = 0;
  SC_CUDA(c, cudaMemcpyAsync(&total, s + oTot, 4, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (outCount) *outCount = total;
  const uint32_t m = std::min(total, cap);
  if (m && out)
  {
    SC_CUDA(c, cudaMemcpyAsync(out, s + oOut, (size_t)m * 88, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return 1;
}

// ---- traffic on rails (scgpu_traffic.cuh) --------------------------------------------------------
int scgpuTrafficSetLanes(ScGpuScene* c, const ScGpuLaneGraph* gph)
{
  if (!enter(c)) return 0;
  if (!gph) return (int)fail(c, "scgpuTrafficSetLanes: graph is NULL");
  if (gph->struct_size != sizeof(ScGpuLaneGraph)) return (int)fail(c, "scgpuTrafficSetLanes: struct_size %u != %zu", gph->struct_size, sizeof(ScGpuLaneGraph));
  const uint32_t nN = gph->nNodes, nS = gph->nSegments, nC = gph->nConnections;
  if ((nN && (!gph->nodePos || !gph->nodeSpeedLimit || !gph->nodeConnOffset)) || (nC && !gph->nodeConn) ||
      (nS && (!gph->segNodes || !gph->segDir || !gph->segLength)))
    return (int)fail(c, "scgpuTrafficSetLanes: NULL array");
  std::vector<float4> nodePos(nN), segDirLen(nS);
  std::vector<uint2> nodeConn(nN);
  std::vector<uint4> segNodes(nS);
  for (uint32_t i = 0; i < nN; ++i)
  {
    const uint32_t o0 = gph->nodeConnOffset[i], o1 = gph->nodeConnOffset[i + 1];
    if (o1 < o0 || o1 > nC) return (int)fail(c, "scgpuTrafficSetLanes: nodeConnOffset[%u..] = %u, %u is not a CSR of %u connections", i, o0, o1, nC);
    nodePos[i] = make_float4(gph->nodePos[3 * i], gph->nodePos[3 * i + 1], gph->nodePos[3 * i + 2], gph->nodeSpeedLimit[i]);
    nodeConn[i] = make_uint2(o0, o1 - o0);
  }
  for (uint32_t i = 0; i < nS; ++i)
  {
    const uint32_t a = gph->segNodes[2 * i], b = gph->segNodes[2 * i + 1];
    if (a >= nN || b >= nN) return (int)fail(c, "scgpuTrafficSetLanes: segment %u references node %u / %u of %u", i, a, b, nN);
    segDirLen[i] = make_float4(gph->segDir[3 * i], gph->segDir[3 * i + 1], gph->segDir[3 * i + 2], gph->segLength[i]);
    segNodes[i] = make_uint4(a, b, (!gph->segActive || gph->segActive[i]) ? 1u : 0u, 0u);
  }
  if (!ensure(c, c->laneNodePos, (size_t)nN * 16) || !ensure(c, c->laneNodeConn, (size_t)nN * 8) || !ensure(c, c->laneConn, (size_t)nC * 4) ||
      !ensure(c, c->laneSegDirLen, (size_t)nS * 16) || !ensure(c, c->laneSegNodes, (size_t)nS * 16))
    return 0;
  // pageable sources: cudaMemcpyAsync returns after staging them, the vectors may die at scope exit
  if (nN && (!uploadTo(c, c->laneNodePos.ptr, nodePos.data(), (size_t)nN * 16) || !uploadTo(c, c->laneNodeConn.ptr, nodeConn.data(), (size_t)nN * 8))) return 0;
  if (nC && !uploadTo(c, c->laneConn.ptr, gph->nodeConn, (size_t)nC * 4)) return 0;
  if (nS && (!uploadTo(c, c->laneSegDirLen.ptr, segDirLen.data(), (size_t)nS * 16) || !uploadTo(c, c->laneSegNodes.ptr, segNodes.data(), (size_t)nS * 16))) return 0;
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  c->nLaneNodes = nN;
  c->nLaneSegs = nS;
  c->laneDefaultSpeed = gph->defaultSpeedLimit;
  c->lanesSet = true;
  return 1;
}

int scgpuTrafficSetLaneActive(ScGpuScene* c, uint32_t n, const uint32_t* segment, const uint8_t* active)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!segment || !active) return (int)fail(c, "scgpuTrafficSetLaneActive: NULL argument");
  if (!c->lanesSet) return (int)fail(c, "scgpuTrafficSetLaneActive: no lane graph (scgpuTrafficSetLanes)");
  for (uint32_t i = 0; i < n; ++i)
    if (segment[i] >= c->nLaneSegs) return (int)fail(c, "scgpuTrafficSetLaneActive: segment %u of %u", segment[i], c->nLaneSegs);
  const size_t oAct = ((size_t)n * 4 + 255) & ~(size_t)255;
  if (!ensure(c, c->staging, oAct + n)) return 0;
  char* s = (char*)c->staging.ptr;
  if (!uploadTo(c, s, segment, (size_t)n * 4) || !uploadTo(c, s + oAct, active, n)) return 0;
  k_lane_set_active<<<blocksFor(n), kBlock, 0, c->stream>>>((uint4*)c->laneSegNodes.ptr, n, (const uint32_t*)s, (const uint8_t*)(s + oAct));
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  return 1;
}

int scgpuTrafficSetAgents(ScGpuScene* c, uint32_t n, const uint32_t* entity, const uint32_t* laneId, const float* laneS,
                          const float* targetSpeed, const float* lookAheadDist)
{
  if (!enter(c)) return 0;
  if (n && (!entity || !laneId || !laneS || !targetSpeed || !lookAheadDist)) return (int)fail(c, "scgpuTrafficSetAgents: NULL argument");
  if (!c->dTrafficMoved && !devAlloc(c, &c->dTrafficMoved, 1, true)) return 0;
  std::vector<uint4> rec(n);
  for (uint32_t i = 0; i < n; ++i)
  {
    uint32_t sBits, vBits;
    std::memcpy(&sBits, laneS + i, 4);
    std::memcpy(&vBits, targetSpeed + i, 4);
    rec[i] = make_uint4(entity[i], laneId[i], sBits, vBits);
  }
  if (!ensure(c, c->trafficAgents, (size_t)n * 16) || !ensure(c, c->trafficLook, (size_t)n * 4)) return 0;
  if (n && (!uploadTo(c, c->trafficAgents.ptr, rec.data(), (size_t)n * 16) || !uploadTo(c, c->trafficLook.ptr, lookAheadDist, (size_t)n * 4))) return 0;
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  c->nTrafficAgents = n;
  return 1;
}

int scgpuTrafficAdvance(ScGpuScene* c, const ScGpuTrafficStep* st, uint32_t* outMoved)
{
  if (!enter(c)) return 0;
  if (outMoved) *outMoved = 0;
  if (!st) return (int)fail(c, "scgpuTrafficAdvance: step is NULL");
  if (st->struct_size != sizeof(ScGpuTrafficStep)) return (int)fail(c, "scgpuTrafficAdvance: struct_size %u != %zu", st->struct_size, sizeof(ScGpuTrafficStep));
  if (!c->lanesSet) return (int)fail(c, "scgpuTrafficAdvance: no lane graph (scgpuTrafficSetLanes)");  // TrafficAIState::lanes == nullptr
  const uint32_t n = c->nTrafficAgents;
  if (n == 0) return 1;
  const float* dBrake = nullptr;
  const uint8_t* dSkip = nullptr;
  if (st->obstacleBrake || st->skip)
  {
    const size_t oSkip = ((size_t)n * 4 + 255) & ~(size_t)255;
    if (!ensure(c, c->trafficInputs, oSkip + n)) return 0;
    char* s = (char*)c->trafficInputs.ptr;
    if (st->obstacleBrake) { if (!uploadTo(c, s, st->obstacleBrake, (size_t)n * 4)) return 0; dBrake = (const float*)s; }
    if (st->skip) { if (!uploadTo(c, s + oSkip, st->skip, n)) return 0; dSkip = (const uint8_t*)(s + oSkip); }
  }
  LaneGraphView g{};
  g.nodePos = (const float4*)c->laneNodePos.ptr;
  g.nodeConn = (const uint2*)c->laneNodeConn.ptr;
  g.conn = (const uint32_t*)c->laneConn.ptr;
  g.segDirLen = (const float4*)c->laneSegDirLen.ptr;
  g.segNodes = (const uint4*)c->laneSegNodes.ptr;
  g.nNodes = c->nLaneNodes;
  g.nSegs = c->nLaneSegs;
  g.defaultSpeedLimit = c->laneDefaultSpeed;
  TrafficStepParams p{};
  p.dt = st->dt;
  p.speedMultiplier = st->speedMultiplier;
  p.lookAheadDist = st->lookAheadDist;
  p.hasDebug = st->hasDebug ? 1u : 0u;
  SC_CUDA(c, cudaMemsetAsync(c->dTrafficMoved, 0, 4, c->stream));
  k_traffic_advance<<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, g, p, n, (uint4*)c->trafficAgents.ptr, (float*)c->trafficLook.ptr, dBrake,
                                                            dSkip, stampOf(c->frame), c->dTrafficMoved);
  c->anyDirty = true;
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  if (outMoved)
  {
    SC_CUDA(c, cudaMemcpyAsync(outMoved, c->dTrafficMoved, 4, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return 1;
}

int scgpuTrafficReadAgents(ScGpuScene* c, uint32_t cap, uint32_t* outLaneId, float* outLaneS, float* outTargetSpeed,
                           float* outLookAheadDist, uint32_t* outCount)
{
  if (!enter(c)) return 0;
  const uint32_t n = c->nTrafficAgents;
  if (outCount) *outCount = n;
  const uint32_t m = std::min(n, cap);
  if (m == 0) return 1;
  std::vector<uint4> rec(m);
  std::vector<float> look(m);
  SC_CUDA(c, cudaMemcpyAsync(rec.data(), c->trafficAgents.ptr, (size_t)m * 16, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaMemcpyAsync(look.data(), c->trafficLook.ptr, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  for (uint32_t i = 0; i < m; ++i)
  {
    if (outLaneId) outLaneId[i] = rec[i].y;
    if (outLaneS) std::memcpy(outLaneS + i, &rec[i].z, 4);
    if (outTargetSpeed) std::memcpy(outTargetSpeed + i, &rec[i].w, 4);
    if (outLookAheadDist) outLookAheadDist[i] = look[i];
  }
  return 1;
}

int scgpuReadLocal(ScGpuScene* c, uint32_t n, const uint32_t* entity, float* outTrs9)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!entity || !outTrs9) return (int)fail(c, "scgpuReadLocal: NULL argument");
  const size_t oOut = (((size_t)n * 4 + 255) & ~(size_t)255);
  if (!ensure(c, c->scratch, oOut + (size_t)n * 36 + 32)) return 0;
  char* s = (char*)c->scratch.ptr;
  uint32_t* dMissing = (uint32_t*)(s + oOut + (((size_t)n * 36 + 15) & ~(size_t)15));
  SC_CUDA(c, cudaMemcpyAsync(s, entity, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  SC_CUDA(c, cudaMemsetAsync(dMissing, 0, 4, c->stream));
  k_gather_local<<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, n, (const uint32_t*)s, (float*)(s + oOut), dMissing);
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  uint32_t missing = 0;
  SC_CUDA(c, cudaMemcpyAsync(outTrs9, s + oOut, (size_t)n * 36, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaMemcpyAsync(&missing, dMissing, 4, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (missing) return (int)fail(c, "scgpuReadLocal: %u of %u handles own no Transform (zeros returned for them)", missing, n);
  return 1;
}

int scgpuDespawn(ScGpuScene* c, uint32_t n, const uint32_t* entity)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!entity) return (int)fail(c, "scgpuDespawn: entity is NULL");
  if (!ensure(c, c->staging, (size_t)n * 12)) return 0;  // before the mirror changes: at most n moves + n victims
  // replay of ComponentPool::remove on the host mirror (scgpu_pool.h) -> net moves in rank space + the victims
  static_assert(sizeof(PoolMove) == sizeof(uint2), "k_despawn_apply reads the moves as uint2");
  std::vector<PoolMove>& moves = c->hMoves;
  std::vector<uint32_t>& removedIdx = c->hRemoved;
  poolReplayDespawn(c->hEntity, c->hSparse, c->count, n, entity, moves, removedIdx, c->hPoolScratch, c->hostThreads, c->workers);

  const uint32_t nMoves = (uint32_t)moves.size(), nRem = (uint32_t)removedIdx.size();
  if (nRem > 0)
  {
    // the victims' device slots become holes (in batch order: a group destroyed as a whole is one run)
    if (c->hSlotScratch.size() < nRem) c->hSlotScratch.resize(nRem + nRem / 4);
    uint32_t* const sl = c->hSlotScratch.data();
    const uint32_t* const so = c->hSlotOf.data();
    const uint32_t* const ri = removedIdx.data();
    poolParallelFor(c->hostThreads, nRem, [=](uint32_t, uint32_t b, uint32_t e) {
      for (uint32_t v = b; v < e; ++v)
      {
        if (v + 16u < e) __builtin_prefetch(so + ri[v + 16u], 0);
        sl[v] = so[ri[v]];
      }
    }, c->workers);
    c->layout.release(nRem, sl);
    char* s = (char*)c->staging.ptr;
    const UpSeg segs[2] = {{s, moves.data(), (size_t)nMoves * 8}, {s + (size_t)nMoves * 8, sl, (size_t)nRem * 4}};
    if (!uploadSegs(c, segs, 2)) return poison(c);
    k_despawn_apply<<<blocksFor((uint64_t)nMoves + nRem), kBlock, 0, c->stream>>>(
      c->a, nMoves, (const uint2*)s, nRem, (const uint32_t*)(s + (size_t)nMoves * 8));
    ++c->launches;
    SC_CUDA_P(c, cudaGetLastError());
    if (c->anyParentEver) c->topologyDirty = true;
    c->anyDirty = true;  // children of a destroyed parent become dirty roots (k_resolve_parents)
  }
  return 1;
}

static int uploadEntityBatch(ScGpuScene* c, uint32_t n, const uint32_t* entity, const void* payload, size_t payloadBytesPer,
                             const uint32_t** dEntity, const void** dPayload)
{
  const size_t bytes = (size_t)n * 4 + (size_t)n * payloadBytesPer;
  if (!ensure(c, c->staging, bytes)) return 0;
  char* s = (char*)c->staging.ptr;
  const UpSeg segs[2] = {{s, entity, (size_t)n * 4}, {s + (size_t)n * 4, payload, (size_t)n * payloadBytesPer}};
  if (!uploadSegs(c, segs, 2)) return 0;
  *dEntity = (const uint32_t*)s;
  if (dPayload) *dPayload = s + (size_t)n * 4;
  return 1;
}

}  // extern "C" (a template cannot have C linkage)

// the setLocal family: kFloats floats per instance, addressed by handle or (entity == nullptr) by dense index
template <int kFloats>
static int setLocalImpl(ScGpuScene* c, const char* who, uint32_t n, const uint32_t* entity, uint32_t firstDense, const float* data)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!data) return (int)fail(c, "%s: NULL argument", who);
  if (!entity && ((uint64_t)firstDense + n > c->count))
    return (int)fail(c, "%s: dense range [%u, %u + %u) exceeds the %u live Transforms", who, firstDense, firstDense, n, c->count);
  const uint32_t* dE = nullptr;
  const void* dT = nullptr;
  if (entity)
  {
    if (!uploadEntityBatch(c, n, entity, data, kFloats * 4, &dE, &dT)) return 0;
  }
  else
  {
    if (!ensure(c, c->staging, (size_t)n * kFloats * 4)) return 0;
    if (!uploadTo(c, c->staging.ptr, data, (size_t)n * kFloats * 4)) return 0;
    dT = c->staging.ptr;
  }
  k_set_local<kFloats><<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, n, dE, firstDense, (const float*)dT, stampOf(c->frame));
  c->anyDirty = true;
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  return 1;
}

extern "C" {

int scgpuSetLocal(ScGpuScene* c, uint32_t n, const uint32_t* entity, const float* trs9)
{
  if (c && n && !entity) { if (enter(c)) fail(c, "scgpuSetLocal: NULL argument"); return 0; }
  return setLocalImpl<9>(c, "scgpuSetLocal", n, entity, 0, trs9);
}

int scgpuSetLocalPosRot(ScGpuScene* c, uint32_t n, const uint32_t* entity, const float* posRot6)
{
  if (c && n && !entity) { if (enter(c)) fail(c, "scgpuSetLocalPosRot: NULL argument"); return 0; }
  return setLocalImpl<6>(c, "scgpuSetLocalPosRot", n, entity, 0, posRot6);
}

int scgpuSetLocalPosition(ScGpuScene* c, uint32_t n, const uint32_t* entity, const float* pos3)
{
  if (c && n && !entity) { if (enter(c)) fail(c, "scgpuSetLocalPosition: NULL argument"); return 0; }
  return setLocalImpl<3>(c, "scgpuSetLocalPosition", n, entity, 0, pos3);
}

int scgpuSetLocalRange(ScGpuScene* c, uint32_t firstDense, uint32_t n, uint32_t floatsPerInstance, const float* data)
{
  switch (floatsPerInstance)
  {
    case SCGPU_LOCAL_POS: return setLocalImpl<3>(c, "scgpuSetLocalRange", n, nullptr, firstDense, data);
    case SCGPU_LOCAL_POS_ROT: return setLocalImpl<6>(c, "scgpuSetLocalRange", n, nullptr, firstDense, data);
    case SCGPU_LOCAL_TRS: return setLocalImpl<9>(c, "scgpuSetLocalRange", n, nullptr, firstDense, data);
    default: break;
  }
  if (enter(c)) fail(c, "scgpuSetLocalRange: floatsPerInstance must be 3 (position), 6 (position, rotation) or 9 (position, rotation, scale), not %u", floatsPerInstance);
  return 0;
}

int scgpuSetRender(ScGpuScene* c, uint32_t n, const uint32_t* entity, const uint32_t* meshMat2, const float* aabb6, const uint32_t* flags)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!entity) return (int)fail(c, "scgpuSetRender: entity is NULL");
  if (!meshMat2 && !aabb6 && !flags) return 1;
  size_t bytes = 0;
  const size_t oE = bytes; bytes += (size_t)n * 4;
  const size_t oM = bytes; bytes += meshMat2 ? (size_t)n * 8 : 0;
  const size_t oB = bytes; bytes += aabb6 ? (size_t)n * 24 : 0;
  const size_t oF = bytes; bytes += flags ? (size_t)n * 4 : 0;
  if (!ensure(c, c->staging, bytes)) return 0;
  char* s = (char*)c->staging.ptr;
  const UpSeg segs[4] = {{s + oE, entity, (size_t)n * 4}, {s + oM, meshMat2, (size_t)n * 8}, {s + oB, aabb6, (size_t)n * 24},
                         {s + oF, flags, (size_t)n * 4}};
  if (!uploadSegs(c, segs, 4)) return 0;
  k_set_render<<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, n, (const uint32_t*)(s + oE), meshMat2 ? (const uint32_t*)(s + oM) : nullptr,
                                                       aabb6 ? (const float*)(s + oB) : nullptr, flags ? (const uint32_t*)(s + oF) : nullptr);
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  return 1;
}

int scgpuSetLocalDevice(ScGpuScene* c, uint32_t n, const uint32_t* d_entity, const float* d_trs9)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!d_entity || !d_trs9) return (int)fail(c, "scgpuSetLocalDevice: NULL argument");
  k_set_local<9><<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, n, d_entity, 0u, d_trs9, stampOf(c->frame));
  c->anyDirty = true;
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  return 1;
}

int scgpuSetParent(ScGpuScene* c, uint32_t n, const uint32_t* entity, const uint32_t* parent)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!entity || !parent) return (int)fail(c, "scgpuSetParent: NULL argument");
  const uint32_t* dE; const void* dP;
  if (!uploadEntityBatch(c, n, entity, parent, 4, &dE, &dP)) return 0;
  k_set_parent<<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, n, dE, (const uint32_t*)dP, stampOf(c->frame));
  c->anyDirty = true;
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  c->anyParentEver = true;
  c->topologyDirty = true;
  return 1;
}

int scgpuMarkDirty(ScGpuScene* c, uint32_t n, const uint32_t* entity)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!entity) return (int)fail(c, "scgpuMarkDirty: NULL argument");
  const uint32_t* dE;
  if (!uploadEntityBatch(c, n, entity, nullptr, 0, &dE, nullptr)) return 0;
  k_mark_dirty<<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, n, dE, stampOf(c->frame));
  c->anyDirty = true;
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  return 1;
}

int scgpuMarkAllDirty(ScGpuScene* c)
{
  if (!enter(c)) return 0;
  c->forceAllDirty = true;
  return 1;
}

// ---- views --------------------------------------------------------------------------------------------

int scgpuSetViews(ScGpuScene* c, uint32_t nViews, const float* viewProj16)
{
  if (!enter(c)) return 0;
  if (nViews == 0 || nViews > c->maxViews) return (int)fail(c, "scgpuSetViews: nViews %u not in 1..%u", nViews, c->maxViews);
  if (!viewProj16) return (int)fail(c, "scgpuSetViews: NULL matrix");
  for (uint32_t v = 0; v < nViews; ++v) planesFromViewProj(viewProj16 + (size_t)v * 16, c->planes.planes[v]);
  refreshPlaneSlack(c->planes, nViews);
  c->nViews = nViews;
  return 1;
}

int scgpuSetViewPlanes(ScGpuScene* c, uint32_t nViews, const float* planes24)
{
  if (!enter(c)) return 0;
  if (nViews == 0 || nViews > c->maxViews) return (int)fail(c, "scgpuSetViewPlanes: nViews %u not in 1..%u", nViews, c->maxViews);
  if (!planes24) return (int)fail(c, "scgpuSetViewPlanes: NULL planes");
  for (uint32_t v = 0; v < nViews; ++v)
    for (int p = 0; p < 6; ++p)
    {
      const float* s = planes24 + (size_t)v * 24 + p * 4;
      c->planes.planes[v][p] = make_float4(s[0], s[1], s[2], s[3]);
    }
  refreshPlaneSlack(c->planes, nViews);
  c->nViews = nViews;
  return 1;
}

int scgpuGetViewPlanes(ScGpuScene* c, uint32_t view, float* out)
{
  if (!enter(c)) return 0;
  if (view >= c->nViews || !out) return (int)fail(c, "scgpuGetViewPlanes: bad view %u", view);
  memcpy(out, c->planes.planes[view], sizeof(float) * 24);
  return 1;
}

// ---- the frame ----------------------------------------------------------------------------------------

int scgpuUpdate(ScGpuScene* c, uint32_t flags)
{
  if (!enter(c)) return 0;
  if (c->nViews == 0) return (int)fail(c, "scgpuUpdate: no views set (scgpuSetViews)");
  const uint32_t stamp = stampOf(c->frame);
  const uint32_t extent = c->layout.extent();  // slots the frame kernels walk: live Transforms + holes
  const uint32_t numTiles = (extent + kTile - 1) / kTile;
  const uint32_t tslot = c->timedUpdates % ScGpuScene::kTimingRing;
  const bool wantCulled = (flags & SCGPU_UPDATE_CULLED_LISTS) != 0;
  const bool skipTransform = (flags & SCGPU_UPDATE_SKIP_TRANSFORM) != 0;
  if (wantCulled)
    for (uint32_t v = 0; v < c->nViews; ++v)
      if (!c->culledEntity[v] && !devAlloc(c, &c->culledEntity[v], (size_t)c->capacityPad, false)) return 0;
  // event pairs around the fused kernel cut the programmatic-dependent-launch chain: timings may be sampled (every
  // n-th update) so that the other frames run undisturbed
  const bool timed = c->timings && (c->updateSerial++ % c->timingEvery) == 0u;
  if (timed) SC_CUDA(c, cudaEventRecord(c->evU0[tslot], c->stream));

  if (c->topologyDirty && extent)
  {
    k_resolve_parents<<<blocksFor((extent + 3u) / 4u), kBlock, 0, c->stream>>>(c->a, extent, stamp);
    ++c->launches;
    SC_CUDA(c, cudaGetLastError());
    if (c->anyParentEver)
    {
      if (extent != c->builtExtent)
      {
        // the cut depends on the extent near its end: the tiles between the old and the new end (and the one before:
        // kHalo) are cut again; tiles that fall off the end are marked for the day the pool grows back
        const uint32_t lo = std::min(extent, c->builtExtent), hi = std::max(extent, c->builtExtent);
        const uint32_t t0 = lo / kTile > 0u ? lo / kTile - 1u : 0u, t1 = std::min(c->maxTiles, hi / kTile + 1u);
        SC_CUDA(c, cudaMemsetAsync(c->a.tileDirty + ((size_t)1 + t0) * 4, 1, (size_t)(t1 - t0) * 4, c->stream));
        c->builtExtent = extent;
      }
      k_build_windows<<<numTiles, kBlock, 0, c->stream>>>(c->a.parentSlot, c->slotInfo, c->winLocal, c->tileWinCount, extent,
                                                          c->a.tileDirty);
      k_scan_tiles<<<1, 1024, 0, c->stream>>>(c->tileWinCount, c->tileWinBase, c->tileWinBase + numTiles, numTiles);
      k_flatten_windows<<<numTiles, 128, 0, c->stream>>>(c->winLocal, c->tileWinCount, c->tileWinBase, c->winList, numTiles, extent,
                                                          c->a.tileDirty);
      c->launches += 3;
      SC_CUDA(c, cudaGetLastError());
    }
  }
  c->topologyDirty = false;

  if (numTiles)
  {
    UpdateParams p{};
    p.rec0 = c->a.rec[0]; p.rec1 = c->a.rec[1]; p.rec2 = c->a.rec[2]; p.rec3 = c->a.rec[3];
    p.w0 = c->a.world[0]; p.w1 = c->a.world[1]; p.w2 = c->a.world[2]; p.w3 = c->a.world[3];
    p.parentSlot = c->a.parentSlot;
    p.rank = c->a.rank;
    p.visBits = c->visBits;
    p.chunkCounts = c->chunkCounts + (size_t)c->chunkParity * (kMaxViews + 1) * c->chunkStride;
    p.chunkStride = c->chunkStride;
    p.acc = c->acc;
    p.count = extent;
    p.live = c->count;
    p.bitWords = c->bitWords;
    p.stamp = stamp;
    p.nViews = c->nViews;
    p.flags = (c->forceAllDirty ? kUpdForceDirty : 0u) | ((flags & SCGPU_UPDATE_FREEZE_CULLING) ? kUpdFreeze : 0u) |
              (skipTransform ? kUpdSkipTransform : 0u) | (wantCulled ? kUpdCandBits : 0u);
    if (timed) SC_CUDA(c, cudaEventRecord(c->evK0[tslot], c->stream));
    // nothing to transform (a cull-only update, or no delta call since the last transforming update): the stored world
    // matrices are valid as they are and neither records nor hierarchy are needed - pure streaming cull
    const bool cullOnly = skipTransform || (!c->forceAllDirty && !c->anyDirty);
#define SC_LAUNCH_UPDATE(V)                                                                                              \
  case V:                                                                                                                \
    if (cullOnly)                                                                                                        \
      k_cull_only<V><<<(numTiles * kSubTiles + kCullSubTiles - 1u) / kCullSubTiles, kBlock, 0, c->stream>>>(p, c->planes); \
    else if (c->anyParentEver)                                                                                           \
    {                                                                                                                    \
      k_update_win<V><<<c->numSMs * SCGPU_WIN_MINBLOCKS, kWinBlock, 0, c->stream>>>(                                     \
        p, c->planes, c->slotInfo, c->winList, c->tileWinBase + numTiles, c->acc + kAccQueueNext, c->slowList);          \
      SC_CUDA(c, launchPdl(k_update_win_slow<V>, c->numSMs * 4u, kWinBlock, c->stream, p, c->planes,                     \
                           (const uint32_t*)c->slotInfo, (const uint32_t*)(c->acc + kAccQueueSlow),                      \
                           (const uint32_t*)c->slowList));                                                               \
      ++c->launches;                                                                                                     \
    }                                                                                                                    \
    else                                                                                                                 \
      k_update_flat<V><<<numTiles, kBlock, kUpdateSmemFlat, c->stream>>>(p, c->planes);                                   \
    break;
    switch (c->nViews)
    {
      SC_LAUNCH_UPDATE(1) SC_LAUNCH_UPDATE(2) SC_LAUNCH_UPDATE(3) SC_LAUNCH_UPDATE(4)
      SC_LAUNCH_UPDATE(5) SC_LAUNCH_UPDATE(6) SC_LAUNCH_UPDATE(7) SC_LAUNCH_UPDATE(8)
      default: return (int)fail(c, "scgpuUpdate: unsupported view count %u", c->nViews);
    }
#undef SC_LAUNCH_UPDATE
    ++c->launches;
    SC_CUDA(c, cudaGetLastError());
    if (timed) SC_CUDA(c, cudaEventRecord(c->evK1[tslot], c->stream));
  }

  // compaction in pool order, totals, and clean bitmaps / counters / queue for the next frame: one launch
  {
    CompactParams q{};
    q.bits = c->visBits;
    for (uint32_t v = 0; v < kMaxViews; ++v) { q.outSlot[v] = c->visSlot[v]; q.culledEntity[v] = c->culledEntity[v]; }
    q.acc = c->acc; q.totals = c->totals; q.totalsHost = c->hTotals;
    q.chunkCounts = c->chunkCounts + (size_t)c->chunkParity * (kMaxViews + 1) * c->chunkStride;
    q.chunkCountsNext = c->chunkCounts + (size_t)(c->chunkParity ^ 1u) * (kMaxViews + 1) * c->chunkStride;
    q.chunkStride = c->chunkStride;
    c->chunkParity ^= 1u;
    q.nWords = std::min(c->bitWords, (((c->count + 31u) / 32u) + 3u) & ~3u);
    const uint32_t grid = std::max(1u, (q.nWords + kCompactChunkWords - 1u) / kCompactChunkWords);  // one CTA per chunk
    q.bitWords = c->bitWords;
    q.nViews = c->nViews;
    q.culled = wantCulled ? 1u : 0u;
    q.listCap = c->capacityPad;
    switch (c->nViews)
    {
#define SC_LAUNCH_COMPACT(V) case V: SC_CUDA(c, launchPdl(k_compact<V>, grid, kCompactThreads, c->stream, q)); break;
      SC_LAUNCH_COMPACT(1) SC_LAUNCH_COMPACT(2) SC_LAUNCH_COMPACT(3) SC_LAUNCH_COMPACT(4)
      SC_LAUNCH_COMPACT(5) SC_LAUNCH_COMPACT(6) SC_LAUNCH_COMPACT(7) SC_LAUNCH_COMPACT(8)
#undef SC_LAUNCH_COMPACT
      default: return (int)fail(c, "scgpuUpdate: unsupported view count %u", c->nViews);
    }
    ResolveParams rp{};
    rp.perm = c->a.perm; rp.entity = c->a.entity; rp.totals = c->totals;
    for (uint32_t v = 0; v < kMaxViews; ++v) { rp.outEntity[v] = c->visEntity[v]; rp.outSlot[v] = c->visSlot[v]; rp.culledEntity[v] = c->culledEntity[v]; }
    rp.nViews = c->nViews;
    rp.culled = q.culled;
    rp.live = c->count;
    rp.extent = extent;
    if (c->peerEnabled)
    {
      // every update is also the producer side of the gather: lists into the root's mailbox, flag by the last CTA
      rp.box = c->peerBox;
      rp.peer = 1u;
      rp.seq = ++c->peerSeq;
      rp.rank = c->rank;
      rp.isRoot = c->rank == c->peerRoot ? 1u : 0u;
      rp.done = c->dPeerState;
      rp.error = c->dPeerState + 1;
      c->peerPublished = true;
    }
    SC_CUDA(c, launchPdl(k_resolve_lists, c->numSMs * 2u, kBlock, c->stream, rp));
    c->launches += 2;
  }
  c->culledListsValid = wantCulled;
  // (the frame totals reach the pinned host copy from k_compact itself: no copy-engine operation in the chain)
  if (timed) { SC_CUDA(c, cudaEventRecord(c->evU1[tslot], c->stream)); ++c->timedUpdates; }
  SC_CUDA(c, cudaEventRecord(c->evDone, c->stream));

  c->lastUpdateFlags = flags;
  c->lastNumTiles = numTiles;
  c->lastExtent = extent;
  c->updatedOnce = true;
  c->gatheredValid = false;
  c->sortedValid = false;
  if (!skipTransform)
  {
    // A cull-only update consumes no dirty stamp: whatever was dirtied before it is still recomputed by the next
    // transforming update, like the reference's t.dirty, which stays set until TransformSystem has seen the node.
    c->forceAllDirty = false;
    c->anyDirty = false;
    ++c->frame;
    if (stampOf(c->frame) == 0u)
    {
      // the 24-bit stamp wraps: forget every stored stamp (all older than this update) so that none aliases a future id
      if (extent)
      {
        k_clear_stamps<<<blocksFor(extent), kBlock, 0, c->stream>>>(c->a, extent);
        ++c->launches;
        SC_CUDA(c, cudaGetLastError());
      }
      c->frame += 1u;  // stamp 0 is reserved for "never dirty"
    }
  }
  return 1;
}

int scgpuLastUpdateTimings(ScGpuScene* c, float* outFusedKernelMs, float* outUpdateMs)
{
  uint32_t n = 0;
  return scgpuReadUpdateTimings(c, outFusedKernelMs, outUpdateMs, 1, &n);
}

int scgpuReadUpdateTimings(ScGpuScene* c, float* outFusedKernelMs, float* outUpdateMs, uint32_t cap, uint32_t* outCount)
{
  if (!enter(c)) return 0;
  if (!c->timings) return (int)fail(c, "timings are disabled (scgpuEnableTimings)");
  if (!waitDone(c)) return 0;
  const uint32_t have = std::min(c->timedUpdates, ScGpuScene::kTimingRing);
  const uint32_t n = std::min(have, cap);
  // newest last: entries [timedUpdates-n, timedUpdates)
  for (uint32_t i = 0; i < n; ++i)
  {
    const uint32_t slot = (c->timedUpdates - n + i) % ScGpuScene::kTimingRing;
    float k = 0.f, u = 0.f;
    if (c->lastNumTiles) SC_CUDA(c, cudaEventElapsedTime(&k, c->evK0[slot], c->evK1[slot]));
    SC_CUDA(c, cudaEventElapsedTime(&u, c->evU0[slot], c->evU1[slot]));
    if (outFusedKernelMs) outFusedKernelMs[i] = k;
    if (outUpdateMs) outUpdateMs[i] = u;
  }
  if (outCount) *outCount = n;
  return 1;
}

// ---- results ------------------------------------------------------------------------------------------

int scgpuGetCounts(ScGpuScene* c, ScGpuCounts* out)
{
  if (!enter(c)) return 0;
  if (!out) return (int)fail(c, "scgpuGetCounts: NULL out");
  if (!waitDone(c)) return 0;
  memset(out, 0, sizeof(*out));
  out->transforms = c->count;
  out->renderablesTotal = c->hTotals[c->nViews];
  for (uint32_t v = 0; v < c->nViews; ++v)
  {
    out->visible[v] = c->hTotals[v];
    out->culled[v] = out->renderablesTotal - c->hTotals[v];
  }
  out->recomputed = c->hTotals[kMaxViews + 1];
  out->slowWindows = c->hTotals[kMaxViews + 2];
  out->extent = c->lastExtent;
  return 1;
}

int scgpuReadVisible(ScGpuScene* c, uint32_t view, uint32_t* outEntity, uint32_t cap, uint32_t* outCount)
{
  if (!enter(c)) return 0;
  if (view >= c->nViews) return (int)fail(c, "scgpuReadVisible: view %u >= %u", view, c->nViews);
  if (!waitDone(c)) return 0;
  const uint32_t n = c->hTotals[view];
  if (outCount) *outCount = n;
  const uint32_t m = std::min(n, cap);
  if (m && outEntity)
  {
    SC_CUDA(c, cudaMemcpyAsync(outEntity, c->visEntity[view], (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return 1;
}

int scgpuReadCulled(ScGpuScene* c, uint32_t view, uint32_t* outEntity, uint32_t cap, uint32_t* outCount)
{
  if (!enter(c)) return 0;
  if (view >= c->nViews) return (int)fail(c, "scgpuReadCulled: view %u >= %u", view, c->nViews);
  if (!waitDone(c)) return 0;
  if (!c->culledListsValid) return (int)fail(c, "scgpuReadCulled: the last update did not pass SCGPU_UPDATE_CULLED_LISTS");
  const uint32_t n = c->hTotals[c->nViews] - c->hTotals[view];
  if (outCount) *outCount = n;
  const uint32_t m = std::min(n, cap);
  if (m && outEntity)
  {
    SC_CUDA(c, cudaMemcpyAsync(outEntity, c->culledEntity[view], (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return 1;
}

int scgpuBuildDrawItemsDevice(ScGpuScene* c, uint32_t view, uint32_t maxDraws, const ScGpuDrawItem** outDevice,
                              uint32_t* outEmitted, uint32_t* outDropped)
{
  if (!enter(c)) return 0;
  if (view >= c->nViews) return (int)fail(c, "scgpuBuildDrawItemsDevice: view %u >= %u", view, c->nViews);
  if (!waitDone(c)) return 0;
  // RenderPrepStreamingSystem (.cpp:1296,1315-1319): maxDraws == 0 means unlimited
  const uint32_t vis = c->hTotals[view];
  const uint32_t emitted = (maxDraws > 0 && vis > maxDraws) ? maxDraws : vis;
  if (!ensure(c, c->drawItems, (size_t)emitted * sizeof(ScGpuDrawItem))) return 0;
  if (emitted)
  {
    k_build_draw_items<<<blocksFor((uint64_t)emitted * 5ull), kBlock, 0, c->stream>>>(
      c->visSlot[view], c->a.entity, c->a.meshMat, c->a.world[0], c->a.world[1], c->a.world[2], c->a.world[3], emitted,
      (float4*)c->drawItems.ptr);
    ++c->launches;
    SC_CUDA(c, cudaGetLastError());
  }
  if (outDevice) *outDevice = (const ScGpuDrawItem*)c->drawItems.ptr;
  if (outEmitted) *outEmitted = emitted;
  if (outDropped) *outDropped = vis - emitted;
  return 1;
}

int scgpuReadDrawItems(ScGpuScene* c, uint32_t view, uint32_t maxDraws, ScGpuDrawItem* out, uint32_t cap,
                       uint32_t* outEmitted, uint32_t* outDropped)
{
  const ScGpuDrawItem* d = nullptr;
  uint32_t emitted = 0, dropped = 0;
  if (!scgpuBuildDrawItemsDevice(c, view, maxDraws, &d, &emitted, &dropped)) return 0;
  if (outEmitted) *outEmitted = emitted;
  if (outDropped) *outDropped = dropped;
  const uint32_t m = std::min(emitted, cap);
  if (m && out)
  {
    SC_CUDA(c, cudaMemcpyAsync(out, d, (size_t)m * sizeof(ScGpuDrawItem), cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return 1;
}

// ---- sorted draws + instanced runs (sc_vk.cpp:1843-1905 on the device) ----------------------
int scgpuBuildSortedDraws(ScGpuScene* c, uint32_t view, uint32_t maxDraws, const uint32_t* materialPipeline, uint32_t nMaterials,
                          uint32_t meshCount, const ScGpuDrawItem** outDevice, uint32_t* outKept,
                          const ScGpuDrawRun** outRunsDevice, uint32_t* outRuns)
{
  static_assert(sizeof(ScGpuDrawRun) == sizeof(DrawRun), "ScGpuDrawRun layout");
  if (!enter(c)) return 0;
  if (view >= c->nViews) return (int)fail(c, "scgpuBuildSortedDraws: view %u >= %u", view, c->nViews);
  if (nMaterials && !materialPipeline) return (int)fail(c, "scgpuBuildSortedDraws: materialPipeline is NULL");
  if (nMaterials > (1u << kDrawIdBits) || meshCount > (1u << kDrawIdBits))
    return (int)fail(c, "scgpuBuildSortedDraws: more than 2^%u materials or meshes", kDrawIdBits);
  if (!waitDone(c)) return 0;
  c->sortedValid = false;
  const uint32_t vis = c->hTotals[view];
  const uint32_t emitted = (maxDraws > 0 && vis > maxDraws) ? maxDraws : vis;  // RenderPrepStreamingSystem's budget first
  if (!c->dSortCounters && !devAlloc(c, &c->dSortCounters, 2, true)) return 0;
  SC_CUDA(c, cudaMemsetAsync(c->dSortCounters, 0, 8, c->stream));
  c->hSortCounters[0] = c->hSortCounters[1] = 0;
  if (emitted)
  {
    // the key holds exactly the bits the asset tables need: mesh | material | pipeline (+ 1 bit: dropped)
    uint32_t maxPipe = 0;
    for (uint32_t m = 0; m < nMaterials; ++m)
    {
      const uint32_t pid = materialPipeline[m];
      if (pid == 0xFFFFFFFFu) continue;
      if (pid >= 63u) return (int)fail(c, "scgpuBuildSortedDraws: pipeline id %u of material %u is >= 63", pid, m);
      maxPipe = std::max(maxPipe, pid);
    }
    auto bitsFor = [](uint32_t count) { uint32_t b = 1; while (b < 31u && (1u << b) < count) ++b; return b; };  // bits for ids 0 .. count-1, at least one
    DrawKeyLayout lay{};
    lay.meshBits = bitsFor(meshCount);        // ids 0 .. meshCount-1
    lay.matBits = bitsFor(nMaterials);
    lay.pipeBits = bitsFor(maxPipe + 1u);
    const uint32_t keyBits = lay.total() + 1u;
    if (!ensure(c, c->matPipe, std::max<size_t>((size_t)nMaterials * 4, 4))) return 0;
    if (nMaterials) SC_CUDA(c, cudaMemcpyAsync(c->matPipe.ptr, materialPipeline, (size_t)nMaterials * 4, cudaMemcpyHostToDevice, c->stream));
    const uint32_t nTiles = (emitted + kSortTile - 1u) / kSortTile;
    const size_t e8 = ((size_t)emitted * 8 + 255) & ~(size_t)255, e4 = ((size_t)emitted * 4 + 255) & ~(size_t)255;
    const size_t hBytes = ((size_t)kRadix * nTiles * 4 + 255) & ~(size_t)255, tBytes = ((size_t)(nTiles + 1) * 4 + 255) & ~(size_t)255;
    if (!ensure(c, c->sortWork, 2 * e8 + 2 * e4 + 2 * hBytes + 2 * tBytes + 256)) return 0;
    char* w = (char*)c->sortWork.ptr;
    uint64_t* keysA = (uint64_t*)w; uint64_t* keysB = (uint64_t*)(w + e8);
    uint32_t* posA = (uint32_t*)(w + 2 * e8); uint32_t* posB = (uint32_t*)(w + 2 * e8 + e4);
    uint32_t* hist = (uint32_t*)(w + 2 * e8 + 2 * e4); uint32_t* histScan = (uint32_t*)(w + 2 * e8 + 2 * e4 + hBytes);
    uint32_t* tileHeads = (uint32_t*)(w + 2 * e8 + 2 * e4 + 2 * hBytes); uint32_t* tileBase = (uint32_t*)(w + 2 * e8 + 2 * e4 + 2 * hBytes + tBytes);
    uint32_t* dummyTotal = (uint32_t*)(w + 2 * e8 + 2 * e4 + 2 * hBytes + 2 * tBytes);
    if (!ensure(c, c->sortedDraws, (size_t)emitted * sizeof(ScGpuDrawItem))) return 0;
    if (!ensure(c, c->drawRuns, (size_t)emitted * sizeof(ScGpuDrawRun))) return 0;
    k_draw_keys<<<blocksFor(emitted), kBlock, 0, c->stream>>>(c->visSlot[view], c->a.meshMat, (const uint32_t*)c->matPipe.ptr,
                                                                nMaterials, meshCount, emitted, lay, keysA, posA, c->dSortCounters);
    ++c->launches;
    // stable LSD passes, 8 bits each, over the bits in use only
    for (uint32_t shift = 0; shift < keyBits; shift += kRadixBits)
    {
      k_radix_hist<<<nTiles, kBlock, 0, c->stream>>>(keysA, emitted, shift, hist, nTiles);
      k_scan_tiles<<<1, 1024, 0, c->stream>>>(hist, histScan, dummyTotal, kRadix * nTiles);
      k_radix_scatter<<<nTiles, kBlock, 0, c->stream>>>(keysA, posA, keysB, posB, emitted, shift, histScan, nTiles);
      std::swap(keysA, keysB);
      std::swap(posA, posB);
      c->launches += 3;
    }
    k_draw_heads<<<nTiles, kBlock, 0, c->stream>>>(keysA, c->dSortCounters, tileHeads);
    k_scan_tiles<<<1, 1024, 0, c->stream>>>(tileHeads, tileBase, dummyTotal, nTiles);
    k_draw_runs<<<nTiles, kBlock, 0, c->stream>>>(keysA, c->dSortCounters, tileBase, lay, (DrawRun*)c->drawRuns.ptr, c->dSortCounters + 1);
    k_draw_run_counts<<<blocksFor(emitted), kBlock, 0, c->stream>>>((DrawRun*)c->drawRuns.ptr, c->dSortCounters + 1);
    k_gather_sorted_draws<<<blocksFor((uint64_t)emitted * 5ull), kBlock, 0, c->stream>>>(
      c->visSlot[view], posA, c->dSortCounters, c->a.entity, c->a.meshMat, c->a.world[0], c->a.world[1], c->a.world[2],
      c->a.world[3], (float4*)c->sortedDraws.ptr);
    c->launches += 5;
    SC_CUDA(c, cudaGetLastError());
    SC_CUDA(c, cudaMemcpyAsync(c->hSortCounters, c->dSortCounters, 8, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  c->sortedValid = true;
  if (outDevice) *outDevice = (const ScGpuDrawItem*)c->sortedDraws.ptr;
  if (outKept) *outKept = c->hSortCounters[0];
  if (outRunsDevice) *outRunsDevice = (const ScGpuDrawRun*)c->drawRuns.ptr;
  if (outRuns) *outRuns = c->hSortCounters[1];
  return 1;
}

int scgpuReadSortedDraws(ScGpuScene* c, ScGpuDrawItem* outItems, uint32_t cap, ScGpuDrawRun* outRuns, uint32_t runCap)
{
  if (!enter(c)) return 0;
  if (!c->sortedValid) return (int)fail(c, "scgpuReadSortedDraws: no scgpuBuildSortedDraws since the last update");
  const uint32_t n = std::min(cap, c->hSortCounters[0]), r = std::min(runCap, c->hSortCounters[1]);
  if (n && outItems) SC_CUDA(c, cudaMemcpyAsync(outItems, c->sortedDraws.ptr, (size_t)n * sizeof(ScGpuDrawItem), cudaMemcpyDeviceToHost, c->stream));
  if (r && outRuns) SC_CUDA(c, cudaMemcpyAsync(outRuns, c->drawRuns.ptr, (size_t)r * sizeof(ScGpuDrawRun), cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  return 1;
}

int scgpuReadWorld(ScGpuScene* c, uint32_t n, const uint32_t* entity, float* out16)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!entity || !out16) return (int)fail(c, "scgpuReadWorld: NULL argument");
  const size_t bytes = (size_t)n * 4 + 256 + (size_t)n * 64 + 16;
  if (!ensure(c, c->scratch, bytes)) return 0;
  char* s = (char*)c->scratch.ptr;
  uint32_t* dE = (uint32_t*)s;
  const size_t oOut = (((size_t)n * 4 + 255) & ~(size_t)255);
  float4* dOut = (float4*)(s + oOut);
  uint32_t* dMissing = (uint32_t*)(s + oOut + (size_t)n * 64);
  SC_CUDA(c, cudaMemcpyAsync(dE, entity, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  SC_CUDA(c, cudaMemsetAsync(dMissing, 0, 4, c->stream));
  k_gather_world<<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, n, dE, dOut, dMissing);
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  uint32_t missing = 0;
  SC_CUDA(c, cudaMemcpyAsync(out16, dOut, (size_t)n * 64, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaMemcpyAsync(&missing, dMissing, 4, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (missing) return (int)fail(c, "scgpuReadWorld: %u of %u handles own no Transform (zeros returned for them)", missing, n);
  return 1;
}

int scgpuReadParents(ScGpuScene* c, uint32_t n, const uint32_t* entity, uint32_t* outParent)
{
  if (!enter(c)) return 0;
  if (n == 0) return 1;
  if (!entity || !outParent) return (int)fail(c, "scgpuReadParents: NULL argument");
  if (!ensure(c, c->scratch, (size_t)n * 8)) return 0;
  uint32_t* dE = (uint32_t*)c->scratch.ptr;
  uint32_t* dO = dE + n;
  SC_CUDA(c, cudaMemcpyAsync(dE, entity, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  k_gather_parent<<<blocksFor(n), kBlock, 0, c->stream>>>(c->a, n, dE, dO);
  ++c->launches;
  SC_CUDA(c, cudaGetLastError());
  SC_CUDA(c, cudaMemcpyAsync(outParent, dO, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  return 1;
}

int scgpuReadDenseEntities(ScGpuScene* c, uint32_t* outEntity, uint32_t cap, uint32_t* outCount)
{
  if (!enter(c)) return 0;
  if (outCount) *outCount = c->count;
  const uint32_t m = std::min(c->count, cap);
  if (m && outEntity)
  {
    // gathered from the device arrays (entity[perm[rank]]), not copied from the host mirror, so that tests can check
    // that the two agree
    if (!ensure(c, c->scratch, (size_t)m * 4)) return 0;
    k_gather_dense<<<blocksFor(m), kBlock, 0, c->stream>>>(c->a, m, (uint32_t*)c->scratch.ptr);
    ++c->launches;
    SC_CUDA(c, cudaGetLastError());
    SC_CUDA(c, cudaMemcpyAsync(outEntity, c->scratch.ptr, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return 1;
}

int scgpuGetDeviceViews(ScGpuScene* c, ScGpuDeviceViews* out)
{
  if (!enter(c)) return 0;
  if (!out) return (int)fail(c, "scgpuGetDeviceViews: NULL out");
  memset(out, 0, sizeof(*out));
  for (uint32_t v = 0; v < c->maxViews; ++v) { out->visibleEntity[v] = c->visEntity[v]; out->visibleSlot[v] = c->visSlot[v]; }
  out->visibleCount = c->totals;
  for (int k = 0; k < 4; ++k) out->worldCol[k] = (const float*)c->a.world[k];
  out->entity = c->a.entity;
  out->count = c->count;
  out->extent = c->layout.extent();
  out->rank = c->a.rank;
  out->perm = c->a.perm;
  return 1;
}

// ---- multi-GPU ----------------------------------------------------------------------------------------

int scgpuCommGetUniqueId(void* outId128)
{
  std::string err;
  if (!outId128 || !loadNccl(err)) { g_createError = err.empty() ? "scgpuCommGetUniqueId: NULL out" : err; return 0; }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_createError = "ncclGetUniqueId failed"; return 0; }
  memcpy(outId128, &id, SCGPU_COMM_ID_BYTES);
  return 1;
}

int scgpuCommInit(ScGpuScene* c, uint32_t nRanks, uint32_t rank, const void* id128)
{
  if (!enter(c)) return 0;
  if (nRanks == 0 || rank >= nRanks || !id128) return (int)fail(c, "scgpuCommInit: bad arguments");
  std::string err;
  if (!loadNccl(err)) return (int)fail(c, "%s", err.c_str());
  ncclUniqueId id;
  memcpy(&id, id128, SCGPU_COMM_ID_BYTES);
  SC_NCCL(c, g_nccl.CommInitRank(&c->comm, (int)nRanks, id, (int)rank));
  c->nRanks = nRanks;
  c->rank = rank;
  const size_t n = (size_t)nRanks * (kMaxViews + 2);
  SC_CUDA(c, cudaMalloc((void**)&c->dAllCounts, n * 4));
  SC_CUDA(c, cudaMallocHost((void**)&c->hAllCounts, n * 4));
  memset(c->hAllCounts, 0, n * 4);
  return 1;
}

int scgpuCommEnablePeerGather(ScGpuScene* c, uint32_t root, uint32_t capEntries)
{
  if (!enter(c)) return 0;
  if (!c->comm) return (int)fail(c, "scgpuCommEnablePeerGather: scgpuCommInit was not called");
  if (root >= c->nRanks) return (int)fail(c, "scgpuCommEnablePeerGather: root %u >= %u ranks", root, c->nRanks);
  if (c->peerEnabled) return (int)fail(c, "scgpuCommEnablePeerGather: already enabled (root %u)", c->peerRoot);
  if (capEntries == 0) capEntries = std::min<uint64_t>((uint64_t)c->capacityPad * c->maxViews, 4u << 20);
  if (!c->dPeerState && !devAlloc(c, &c->dPeerState, 4, true)) return 0;
  // the root allocates the mailbox and publishes its IPC handle through NCCL (the bootstrap channel we have)
  cudaIpcMemHandle_t handle{};
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  const size_t bytes = PeerBox::bytes(c->nRanks, capEntries);
  void* local = nullptr;
  if (c->rank == root)
  {
    SC_CUDA(c, cudaMalloc(&local, bytes));
    SC_CUDA(c, cudaMemsetAsync(local, 0, bytes, c->stream));
    SC_CUDA(c, cudaIpcGetMemHandle(&handle, local));
  }
  if (!ensure(c, c->scratch, 256)) return 0;
  if (c->rank == root) SC_CUDA(c, cudaMemcpyAsync(c->scratch.ptr, &handle, 64, cudaMemcpyHostToDevice, c->stream));
  SC_NCCL(c, g_nccl.Broadcast(c->scratch.ptr, c->scratch.ptr, 16, ncclUint32, (int)root, c->comm, c->stream));
  SC_CUDA(c, cudaMemcpyAsync(&handle, c->scratch.ptr, 64, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->rank != root)
  {
    SC_CUDA(c, cudaIpcOpenMemHandle(&local, handle, cudaIpcMemLazyEnablePeerAccess));
    c->peerMapped = true;
  }
  c->peerBox.base = (uint32_t*)local;
  c->peerBox.nRanks = c->nRanks;
  c->peerBox.cap = capEntries;
  c->peerRoot = root;
  c->peerSeq = 0;
  c->peerEnabled = true;
  return 1;
}

int scgpuGatherVisible(ScGpuScene* c, uint32_t root)
{
  if (!enter(c)) return 0;
  if (!c->comm) return (int)fail(c, "scgpuGatherVisible: scgpuCommInit was not called");
  if (root >= c->nRanks) return (int)fail(c, "scgpuGatherVisible: root %u >= %u ranks", root, c->nRanks);
  if (!c->updatedOnce) return (int)fail(c, "scgpuGatherVisible: no update issued");
  const size_t row = kMaxViews + 2;
  if (c->peerEnabled && root == c->peerRoot)
  {
    // peer-memory path: the update's last kernel (k_resolve_lists) has already stored this rank's lists into the root's
    // mailbox over NVLink and raised its flag; the root waits for all flags on the device, nobody else does anything
    if (!c->peerPublished)
      return (int)fail(c, "scgpuGatherVisible: scgpuCommEnablePeerGather must precede the scgpuUpdate whose lists are gathered");
    if (c->rank == root)
    {
      SC_CUDA(c, launchPdl(k_peer_wait, 1u, 64u, c->stream, c->peerBox, c->peerSeq, c->dAllCounts, c->dPeerState + 1 + (c->peerSeq & 1u)));
      ++c->launches;
    }
    SC_CUDA(c, cudaGetLastError());
    c->gatheredValid = true;
    c->lastGatherPeer = true;
    return 1;
  }
  c->lastGatherPeer = false;
  // 1) counts of every rank to every rank (V*4 bytes each; one small allgather)
  SC_NCCL(c, g_nccl.AllGather(c->totals, c->dAllCounts, row, ncclUint32, c->comm, c->stream));
  ++c->launches;
  SC_CUDA(c, cudaMemcpyAsync(c->hAllCounts, c->dAllCounts, (size_t)c->nRanks * row * 4, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));  // receive sizes must be known on the host
  // 2) lists: exact-size send/recv to the submitting rank, concatenated in rank order
  if (c->rank == root)
  {
    size_t need = 0;
    for (uint32_t v = 0; v < c->nViews; ++v)
    {
      size_t tot = 0;
      for (uint32_t r = 0; r < c->nRanks; ++r) tot += c->hAllCounts[r * row + v];
      need = std::max(need, tot);
    }
    if (need > c->gatheredCap)
    {
      for (uint32_t v = 0; v < c->maxViews; ++v)
      {
        if (c->gathered[v]) SC_CUDA(c, cudaFree(c->gathered[v]));
        c->gathered[v] = nullptr;
        SC_CUDA(c, cudaMalloc((void**)&c->gathered[v], std::max(need, (size_t)1) * 4));
      }
      c->gatheredCap = need;
    }
  }
  SC_NCCL(c, g_nccl.GroupStart());
  for (uint32_t v = 0; v < c->nViews; ++v)
  {
    if (c->rank == root)
    {
      size_t off = 0;
      for (uint32_t r = 0; r < c->nRanks; ++r)
      {
        const size_t cnt = c->hAllCounts[r * row + v];
        if (r == root)
        {
          if (cnt) SC_CUDA(c, cudaMemcpyAsync(c->gathered[v] + off, c->visEntity[v], cnt * 4, cudaMemcpyDeviceToDevice, c->stream));
        }
        else if (cnt)
        {
          SC_NCCL(c, g_nccl.Recv(c->gathered[v] + off, cnt, ncclUint32, (int)r, c->comm, c->stream));
        }
        off += cnt;
      }
    }
    else
    {
      const size_t cnt = c->hAllCounts[c->rank * row + v];
      if (cnt) SC_NCCL(c, g_nccl.Send(c->visEntity[v], cnt, ncclUint32, (int)root, c->comm, c->stream));
    }
  }
  SC_NCCL(c, g_nccl.GroupEnd());
  ++c->launches;
  c->gatheredValid = true;
  return 1;
}

// peer mode: the counts rows and the error word live on the device until somebody asks
static int peerFetchCounts(ScGpuScene* c)
{
  if (c->rank != c->peerRoot) return (int)fail(c, "peer gather: counts and lists exist on the root (rank %u) only", c->peerRoot);
  uint32_t err = 0;
  SC_CUDA(c, cudaMemcpyAsync(c->hAllCounts, c->dAllCounts, (size_t)c->nRanks * (kMaxViews + 2) * 4, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaMemcpyAsync(&err, c->dPeerState + 1 + (c->peerSeq & 1u), 4, cudaMemcpyDeviceToHost, c->stream));
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (err & 2u) return (int)fail(c, "peer gather: a rank did not deliver within the time limit");
  if (err & 4u) return (int)fail(c, "peer gather: a rank's visible lists exceed the mailbox capacity (scgpuCommEnablePeerGather capEntries)");
  if (err & 8u) return (int)fail(c, "peer gather: a rank could not deliver this frame (the root had not released its mailbox in time)");
  return 1;
}

int scgpuGetGatheredCounts(ScGpuScene* c, uint32_t* outCounts, uint32_t capRanks)
{
  if (!enter(c)) return 0;
  if (!c->gatheredValid || !outCounts) return (int)fail(c, "scgpuGetGatheredCounts: no gather since the last update");
  if (c->lastGatherPeer && !peerFetchCounts(c)) return 0;
  const size_t row = kMaxViews + 2;
  for (uint32_t r = 0; r < c->nRanks && r < capRanks; ++r)
    for (uint32_t v = 0; v < c->nViews; ++v) outCounts[r * c->nViews + v] = c->hAllCounts[r * row + v];
  return 1;
}

int scgpuReadGatheredVisible(ScGpuScene* c, uint32_t view, uint32_t* outEntity, uint32_t cap, uint32_t* outCount)
{
  if (!enter(c)) return 0;
  if (!c->gatheredValid) return (int)fail(c, "scgpuReadGatheredVisible: no gather since the last update");
  if (view >= c->nViews) return (int)fail(c, "scgpuReadGatheredVisible: bad view");
  const size_t row = kMaxViews + 2;
  if (c->lastGatherPeer)
  {
    if (!peerFetchCounts(c)) return 0;
    size_t total = 0, done = 0;
    for (uint32_t r = 0; r < c->nRanks; ++r) total += c->hAllCounts[r * row + view];
    if (outCount) *outCount = (uint32_t)total;
    // rank r's slice of this view starts after its earlier views inside its payload
    for (uint32_t r = 0; r < c->nRanks && outEntity && done < cap; ++r)
    {
      size_t off = 0;
      for (uint32_t v = 0; v < view; ++v) off += c->hAllCounts[r * row + v];
      const size_t cnt = std::min((size_t)c->hAllCounts[r * row + view], (size_t)cap - done);
      if (cnt)
        SC_CUDA(c, cudaMemcpyAsync(outEntity + done, c->peerBox.payload(r, c->peerSeq & 1u) + off, cnt * 4, cudaMemcpyDeviceToHost, c->stream));
      done += cnt;
    }
    SC_CUDA(c, cudaStreamSynchronize(c->stream));
    return 1;
  }
  size_t tot = 0;
  for (uint32_t r = 0; r < c->nRanks; ++r) tot += c->hAllCounts[r * row + view];
  if (outCount) *outCount = (uint32_t)tot;
  if (!c->gathered[view]) return (int)fail(c, "scgpuReadGatheredVisible: this rank was not the gather root");
  const size_t m = std::min(tot, (size_t)cap);
  if (m && outEntity)
  {
    SC_CUDA(c, cudaMemcpyAsync(outEntity, c->gathered[view], m * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  SC_CUDA(c, cudaStreamSynchronize(c->stream));
  return 1;
}

}  // extern "C"
