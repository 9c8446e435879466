// cab_comm.cu -- the multi-GPU data plane behind the C ABI: one cab_ctx per GPU, one host thread or process per context.
//
// The reference runs its plugins single-threaded on one CPU (cloud_algos/src/radius_estimation.cpp:139 "TODO
// parallelize!"); sharding is this implementation's own: queries shard as slabs of rows (cab_grid.cu), the
// cloud is replicated, the results are concatenated.  Nothing here is a library collective on the data path:
//   * the cloud is replicated by copy engines: every rank uploads 1/world of it over its own PCIe link and copies that
//     slice into the peers' buffers over NVLink (cab_comm_upload_cloud);
//   * the concatenation of the results is fused into the RSD kernel: a warp stores its packet's normals, radii and
//     input indices straight into every rank's copy of the concatenated arrays (peer memory, cab_rsd.cu push_results);
//     all that is left of the collective are two tiny flag kernels per step (slice offsets before, completion after).
// Peer memory is plain cudaMalloc memory reached through cudaDeviceEnablePeerAccess (contexts of one process) or CUDA IPC
// handles (one process per GPU, the torchrun layout).  NCCL is used for what it is good at -- the bootstrap (moving the
// IPC handles) and the small integer all-reduce of GRSD histograms (grsd_colorCHLAC_tools.hpp:230-260 accumulates one
// 6x6 matrix per histogram; its integer sums are order independent) -- and is loaded with dlopen so that a single-GPU
// user of the library does not need it.
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "cab_internal.cuh"

namespace cab {

struct Id128 {  // ncclUniqueId
  char b[CAB_COMM_ID_BYTES];
};

namespace {

// ---- NCCL through dlopen (the subset used; declarations follow nccl.h 2.x, whose ABI for these calls is stable) ----
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128 /* ncclUniqueId by value */, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclInt8 = 0, kNcclInt32 = 2, kNcclSum = 0;

NcclApi* nccl_api(std::string* why) {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?");
      return;
    }
    auto sym = [&](const char* n) {
      void* p = dlsym(api.lib, n);
      if (!p && err.empty()) err = std::string("symbol missing in libnccl: ") + n;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  if (!err.empty()) {
    if (why) *why = err;
    return nullptr;
  }
  return &api;
}

// What a rank tells the others about its exchange buffers.
// kBufWork: the rank's WORKING normals (local sorted order, what its kernels read and write: ctx->b_nrm is pointed at
// it while the context belongs to a group), reachable by the neighbours for the halo exchange
constexpr int kBufNrm = 0, kBufRsd = 1, kBufPerm = 2, kBufCloud = 3, kBufSync = 4, kBufWork = 5, kNumBufs = 6;
struct CommBlob {
  int64_t pid;
  int32_t device;
  int32_t rank;
  int64_t cap_points;
  char uuid[16];                          // of the device: two ranks on one GPU are told apart from two GPUs
  uint64_t raw[kNumBufs];                 // device pointers (meaningful inside process `pid`)
  cudaIpcMemHandle_t handle[kNumBufs];    // the same allocations for other processes
};
static_assert(sizeof(CommBlob) <= CAB_COMM_BLOB_BYTES, "CAB_COMM_BLOB_BYTES too small");

// Contexts of one process (one host thread each): the blob table and a host barrier.
struct LocalGroup {
  std::mutex m;
  std::condition_variable cv;
  int world = 0, arrived = 0;
  unsigned gen = 0;
  std::vector<CommBlob> blobs;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++arrived == world) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

// Sync block of a rank (ints / 64-bit words in its own memory, written by the peers).
struct SyncBlock {
  unsigned long long mail[kMaxPeers];   // (seq << 32) | own query count of rank p
  unsigned done[kMaxPeers];             // seq once rank p's pushes of this step are complete
  unsigned cloud[kMaxPeers];            // cloud_seq once rank p's slice of the cloud has arrived
  unsigned normals[kMaxPeers];          // seq once rank p's normals pass is complete and its top layer has been stored into the rank above
  unsigned long long total;             // sum of the ranks' counts of the last step (host reads it back)
  // measured-time feedback of the shard balance + the consistency check of the cuts: with its completion flag every rank
  // leaves how long its normals + RSD passes took (device clock, from the end of its build) and how many queries it
  // answered.  Contiguous: the host reads busy[] and count[] back in one copy.
  unsigned long long busy[kMaxPeers];
  unsigned long long count[kMaxPeers];
  unsigned long long t_begin;           // %globaltimer at the end of this rank's build
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct PeerSync {
  SyncBlock* block[kMaxPeers];
};

// after the slab build: tell every rank how many queries this rank answers
__global__ void post_count_kernel(const SlabInfo* __restrict__ info, PeerSync peers, int rank, int world, unsigned seq) {
  const int p = threadIdx.x;
  if (p >= world) return;
  const unsigned long long count = (unsigned long long)(unsigned)(info->q1 - info->q0);
  st_release_sys(&peers.block[p]->mail[rank], ((unsigned long long)seq << 32) | count);
}

// before the pushes: wait for every rank's count, place this rank's slice in the concatenation
__global__ void wait_counts_kernel(SlabInfo* __restrict__ info, SyncBlock* __restrict__ mine, int rank, int world, unsigned seq,
                                   unsigned long long timeout_ns) {
  __shared__ unsigned long long counts[kMaxPeers];
  __shared__ int failed;
  const int p = threadIdx.x;
  if (p == 0) failed = 0;
  __syncthreads();
  if (p < world) {
    const unsigned long long t0 = global_timer_ns();
    unsigned long long v;
    for (;;) {
      v = ld_acquire_sys(&mine->mail[p]);
      if ((unsigned)(v >> 32) == seq) break;
      if (global_timer_ns() - t0 > timeout_ns) {
        failed = 1;
        v = 0;
        break;
      }
      __nanosleep(200);
    }
    counts[p] = v & 0xffffffffull;
  }
  __syncthreads();
  if (p == 0) {
    unsigned long long base = 0, total = 0;
    for (int q = 0; q < world; ++q) {
      if (q < rank) base += counts[q];
      total += counts[q];
    }
    info->gbase = (int)base;
    if (failed) info->error = 1;
    mine->total = total;
  }
}

__global__ void stamp_kernel(SyncBlock* __restrict__ mine) {
  if (threadIdx.x == 0) mine->t_begin = global_timer_ns();
}

// after the RSD kernel (whose stores to peer memory are complete at the kernel boundary): completion flag, busy time and
// query count to every rank
__global__ void post_done_kernel(PeerSync peers, const SyncBlock* __restrict__ mine, const SlabInfo* __restrict__ info, int rank,
                                 int world, unsigned seq) {
  __shared__ unsigned long long busy_shared;  // ONE reading of the clock: every rank must receive the same number
  const int p = threadIdx.x;
  if (p == 0) busy_shared = global_timer_ns() - mine->t_begin;
  __syncthreads();
  if (p >= world) return;
  const unsigned long long busy = busy_shared;
  peers.block[p]->busy[rank] = busy;
  peers.block[p]->count[rank] = (unsigned long long)(unsigned)(info->q1 - info->q0);
  __threadfence_system();
  st_release_sys(&peers.block[p]->done[rank], seq);
}

// after the RSD kernel (whose stores to peer memory are complete at the kernel boundary): tell every rank
__global__ void post_flag_kernel(PeerSync peers, int rank, int world, unsigned seq, int which) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  unsigned* f = which == 0 ? &peers.block[p]->done[rank] : which == 1 ? &peers.block[p]->cloud[rank] : &peers.block[p]->normals[rank];
  st_release_sys(f, seq);
}

// waits for the flag of every rank (from < 0) or of rank `from` alone
__global__ void wait_flag_kernel(SlabInfo* __restrict__ info, int* __restrict__ error_out, SyncBlock* __restrict__ mine, int world,
                                 unsigned seq, int which, unsigned long long timeout_ns, int from = -1) {
  const int p = threadIdx.x;
  if (p >= world || (from >= 0 && p != from)) return;
  const unsigned* f = which == 0 ? &mine->done[p] : which == 1 ? &mine->cloud[p] : &mine->normals[p];
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(f) != seq) {
    if (global_timer_ns() - t0 > timeout_ns) {
      if (info) info->error = 1;
      if (error_out) *error_out = 1;
      break;
    }
    __nanosleep(200);
  }
}

// Halo exchange, both directions driven by the LOWER rank of a pair (it knows every offset involved: its top layer is
// the upper rank's first rows, and the upper rank's bottom layer starts where that rank's lower halo ends, i.e. after
// as many points as the lower rank's top layer holds).
//   send: my top layer of own rows  -> positions [0, count) of the rank above
__global__ void __launch_bounds__(256) halo_send_kernel(const SlabInfo* __restrict__ info, const float4* __restrict__ mine,
                                                        float4* __restrict__ above) {
  const int first = info->top_src, count = info->q1 - first;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) above[i] = mine[first + i];
}
//   fetch: the bottom layer of the rank above -> my upper halo, positions [q1, n_selected)
__global__ void __launch_bounds__(256) halo_fetch_kernel(const SlabInfo* __restrict__ info, float4* __restrict__ mine,
                                                         const float4* __restrict__ above) {
  const int src = info->q1 - info->top_src, dst = info->q1, count = info->n_selected - info->q1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) mine[dst + i] = above[src + i];
}

// concatenated (sorted-order) results -> one range [j0, j1) of the input order
__global__ void __launch_bounds__(256) range_unpermute_kernel(const int* __restrict__ perm, long long total, int j0, int j1,
                                                              const float4* __restrict__ nrm, const float2* __restrict__ rsd,
                                                              float4* __restrict__ out4, float* __restrict__ out_a,
                                                              float* __restrict__ out_b) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  const int j = perm[s];
  if (j < j0 || j >= j1) return;
  if (out4) out4[j - j0] = nrm[s];
  if (out_a) {
    const float2 v = rsd[s];
    out_a[j - j0] = v.x;
    out_b[j - j0] = v.y;
  }
}

__global__ void __launch_bounds__(256) fill_defaults_kernel(float4* __restrict__ out4, float* __restrict__ out_a,
                                                            float* __restrict__ out_b, int m, float radius) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float nan = __int_as_float(0x7fc00000);
  if (out4) out4[i] = make_float4(nan, nan, nan, nan);
  if (out_a) {
    out_a[i] = radius;
    out_b[i] = radius;
  }
}

// input-range layout: the rank's own range is already in input order; only the radii pairs are split for the caller
__global__ void __launch_bounds__(256) split_radii_kernel(const float2* __restrict__ rsd, int m, float* __restrict__ out_a,
                                                          float* __restrict__ out_b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float2 v = rsd[i];
  out_a[i] = v.x;
  out_b[i] = v.y;
}

}  // namespace

struct CommState {
  int rank = 0, world = 1;
  std::shared_ptr<LocalGroup> local;  // contexts of this process linked by cab_comm_init_local
  void* nccl = nullptr;               // ncclComm_t (bootstrap + integer all-reduce)
  DevBuf own[kNumBufs];
  int64_t cap_points = 0;
  bool connected = false;
  void* peer[kMaxPeers][kNumBufs] = {};
  bool peer_is_ipc[kMaxPeers] = {};
  unsigned seq = 0, cloud_seq = 0;
  bool armed = false;                 // this step's RSD kernel pushes its results
  bool feedback = true;               // measured-time correction of the ranks' shares (cab_comm_set_feedback)
  double share[kMaxPeers];            // the ranks' shares of the modelled cost; the same numbers on every rank
  int layout = CAB_COMM_LAYOUT_REPLICATED;  // where the concatenated results live (cab_comm_set_layout)
  int64_t last_n = 0;                 // points of the cloud of the last step
  int64_t last_total = 0;             // entries of the concatenated arrays after the last step
  double last_plane_radius = 0.1;
  DevBuf stage;                       // device staging for the blob all-gather / integer all-reduce
  unsigned long long timeout_ns = 5ull * 1000 * 1000 * 1000;
};

namespace {

CommState* new_comm_state(int rank, int world) {
  CommState* cs = new CommState();
  cs->rank = rank;
  cs->world = world;
  for (int p = 0; p < kMaxPeers; ++p) cs->share[p] = 1.0 / world;
  if (std::getenv("CAB_NO_FEEDBACK")) cs->feedback = false;  // A/B switch for tuning runs
  return cs;
}

int nccl_fail(cab_ctx* ctx, const char* what, int code) {
  NcclApi* api = nccl_api(nullptr);
  return fail(ctx, CAB_ERR_CUDA, "%s failed: %s", what, api && api->GetErrorString ? api->GetErrorString(code) : "?");
}

void close_peers(cab_ctx* ctx) {
  CommState* cs = ctx->comm;
  for (int p = 0; p < kMaxPeers; ++p) {
    if (cs->peer_is_ipc[p])
      for (int b = 0; b < kNumBufs; ++b)
        if (cs->peer[p][b]) cudaIpcCloseMemHandle(cs->peer[p][b]);
    for (int b = 0; b < kNumBufs; ++b) cs->peer[p][b] = nullptr;
    cs->peer_is_ipc[p] = false;
  }
  cs->connected = false;
}

// (Re)allocates this rank's exchange buffers for clouds of up to n points and describes them.
int make_blob(cab_ctx* ctx, int64_t n, CommBlob* blob) {
  CommState* cs = ctx->comm;
  cudaStreamSynchronize(ctx->stream);
  close_peers(ctx);
  const size_t np = (size_t)std::max<int64_t>(n, 1);
  const size_t bytes[kNumBufs] = {np * sizeof(float4), np * sizeof(float2), np * sizeof(int), np * 3 * sizeof(float) + 64,
                                  sizeof(SyncBlock), np * sizeof(float4)};
  if (ctx->b_nrm.p == cs->own[kBufWork].p) ctx->b_nrm = DevBuf{};  // the working normals move with their buffer
  for (int b = 0; b < kNumBufs; ++b) {
    if (cs->own[b].p && cs->own[b].cap >= bytes[b]) continue;
    if (cs->own[b].p) cudaFree(cs->own[b].p);
    cs->own[b] = DevBuf{};
    // plain cudaMalloc, one allocation per buffer: that is what cudaIpcGetMemHandle can export
    const size_t want = b == kBufSync ? bytes[b] : bytes[b] + bytes[b] / 16;
    cudaError_t e = cudaMalloc(&cs->own[b].p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cs->own[b].p = nullptr;
      return fail(ctx, CAB_ERR_OOM, "cab_comm: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    cs->own[b].cap = want;
    if (b == kBufSync) {
      CAB_CUDA(ctx, cudaMemsetAsync(cs->own[b].p, 0xff, want, ctx->stream));  // no flag equals a sequence number yet
      CAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      cs->seq = cs->cloud_seq = 0;
    }
  }
  cs->cap_points = (int64_t)np;
  // the context's normals live in the exported buffer from now on
  if (ctx->b_nrm.p && ctx->b_nrm.p != cs->own[kBufWork].p) cudaFree(ctx->b_nrm.p);
  ctx->b_nrm = cs->own[kBufWork];
  ctx->have_normals = false;
  std::memset(blob, 0, sizeof(*blob));
  blob->pid = (int64_t)getpid();
  blob->device = ctx->device;
  blob->rank = cs->rank;
  blob->cap_points = cs->cap_points;
  {
    cudaDeviceProp prop;
    CAB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    std::memcpy(blob->uuid, &prop.uuid, 16);
  }
  for (int b = 0; b < kNumBufs; ++b) {
    blob->raw[b] = (uint64_t)(uintptr_t)cs->own[b].p;
    cudaError_t e = cudaIpcGetMemHandle(&blob->handle[b], cs->own[b].p);
    if (e != cudaSuccess) {
      cudaGetLastError();  // contexts of one process do not need the handles
      std::memset(&blob->handle[b], 0, sizeof(blob->handle[b]));
    }
  }
  return CAB_OK;
}

int connect_blobs(cab_ctx* ctx, const CommBlob* blobs) {
  CommState* cs = ctx->comm;
  const int64_t me = (int64_t)getpid();
  // Ranks that share a GPU run their kernels side by side, and a step's kernels wait for the peers' flags.  With CUDA's
  // lazy module loading the FIRST launch of a kernel may have to wait until the device is idle -- i.e. for a peer's
  // kernel that is itself waiting for this rank: the first step would stall until the timeout.  Eager loading (set
  // before CUDA is initialised) is the documented remedy; one GPU per rank, the production layout, is not affected.
  // Likewise the ranks' streams (two per rank) must not share a hardware queue, where a waiting kernel of one rank
  // would hold back the kernels of another: CUDA_DEVICE_MAX_CONNECTIONS (default 8) >= 2 x the ranks on the GPU.
  for (int p = 0; p < cs->world; ++p) {
    int sharing = 0;
    for (int q = 0; q < cs->world; ++q) sharing += std::memcmp(blobs[p].uuid, blobs[q].uuid, 16) == 0 ? 1 : 0;
    if (sharing < 2) continue;
    const char* mode = std::getenv("CUDA_MODULE_LOADING");
    if (!mode || std::strcmp(mode, "EAGER") != 0)
      return fail(ctx, CAB_ERR_STATE, "cab_comm: %d ranks share the GPU of rank %d; set CUDA_MODULE_LOADING=EAGER before CUDA is "
                                      "initialised (lazy kernel loading can stall a step whose kernels wait for each other)", sharing, p);
    const char* conn = std::getenv("CUDA_DEVICE_MAX_CONNECTIONS");
    const int have = conn ? std::atoi(conn) : 8;
    if (have < 2 * sharing)
      return fail(ctx, CAB_ERR_STATE, "cab_comm: %d ranks share the GPU of rank %d; set CUDA_DEVICE_MAX_CONNECTIONS >= %d before "
                                      "CUDA is initialised (streams sharing a hardware queue serialise behind a waiting kernel)",
                  sharing, p, 2 * sharing);
  }
  for (int p = 0; p < cs->world; ++p) {
    const CommBlob& b = blobs[p];
    if (b.rank != p) return fail(ctx, CAB_ERR_ARG, "cab_comm: blob %d describes rank %d", p, b.rank);
    if (b.cap_points != cs->cap_points)
      return fail(ctx, CAB_ERR_ARG, "cab_comm: rank %d reserved %lld points, this rank %lld", p, (long long)b.cap_points,
                  (long long)cs->cap_points);
    if (p == cs->rank) {
      for (int k = 0; k < kNumBufs; ++k) cs->peer[p][k] = cs->own[k].p;
    } else if (b.pid == me) {
      if (b.device != ctx->device) {
        int can = 0;
        CAB_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, b.device));
        if (!can) return fail(ctx, CAB_ERR_CUDA, "cab_comm: device %d cannot access device %d", ctx->device, b.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(ctx, CAB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(e));
        cudaGetLastError();
      }
      for (int k = 0; k < kNumBufs; ++k) cs->peer[p][k] = (void*)(uintptr_t)b.raw[k];
    } else {
      for (int k = 0; k < kNumBufs; ++k) {
        cudaError_t e = cudaIpcOpenMemHandle(&cs->peer[p][k], b.handle[k], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          cudaGetLastError();
          return fail(ctx, CAB_ERR_CUDA, "cudaIpcOpenMemHandle (rank %d, buffer %d): %s", p, k, cudaGetErrorString(e));
        }
      }
      cs->peer_is_ipc[p] = true;
    }
  }
  cs->connected = true;
  return CAB_OK;
}

PeerSync peer_sync(const CommState* cs) {
  PeerSync ps{};
  for (int p = 0; p < cs->world; ++p) ps.block[p] = (SyncBlock*)cs->peer[p][kBufSync];
  return ps;
}

}  // namespace

// Collective: makes sure every rank's exchange buffers hold n points and are mapped on every rank.  A no-op once they
// do (every rank takes the same decision: n is the same everywhere).
int comm_prepare(cab_ctx* ctx, int64_t n) {
  CommState* cs = ctx->comm;
  if (!cs) return fail(ctx, CAB_ERR_STATE, "cab_comm: this context belongs to no group (cab_comm_init / cab_comm_init_local)");
  if (cs->connected && n <= cs->cap_points) return CAB_OK;
  if (!cs->local && !cs->nccl)
    return fail(ctx, CAB_ERR_STATE, "cab_comm: buffers for %lld points are not connected (cab_comm_reserve / _export / _connect)", (long long)n);
  CommBlob mine;
  if (int rc = make_blob(ctx, n + n / 8, &mine)) return rc;
  std::vector<CommBlob> all(cs->world);
  if (cs->local) {
    {
      std::lock_guard<std::mutex> lk(cs->local->m);
      cs->local->blobs.resize(cs->world);
      cs->local->blobs[cs->rank] = mine;
    }
    cs->local->barrier();
    {
      std::lock_guard<std::mutex> lk(cs->local->m);
      all = cs->local->blobs;
    }
    cs->local->barrier();  // nobody overwrites the table before everybody has read it
  } else {
    NcclApi* api = nccl_api(nullptr);
    const size_t sz = sizeof(CommBlob);
    if (int rc = reserve(ctx, cs->stage, sz * (cs->world + 1))) return rc;
    char* d = (char*)cs->stage.p;
    CAB_CUDA(ctx, cudaMemcpyAsync(d, &mine, sz, cudaMemcpyHostToDevice, ctx->stream));
    int e = api->AllGather(d, d + sz, sz, kNcclInt8, cs->nccl, ctx->stream);
    if (e) return nccl_fail(ctx, "ncclAllGather", e);
    CAB_CUDA(ctx, cudaMemcpyAsync(all.data(), d + sz, sz * cs->world, cudaMemcpyDeviceToHost, ctx->stream));
    CAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return connect_blobs(ctx, all.data());
}

void comm_push_targets(const cab_ctx* ctx, PushTargets* out) {
  out->world = 0;
  const CommState* cs = ctx->comm;
  if (!cs || !cs->armed || !cs->connected) return;
  out->world = cs->world;
  out->layout = cs->layout;
  out->n = ctx->n;
  for (int p = 0; p <= cs->world; ++p) out->lo[p] = (int)(ctx->n * p / cs->world);
  for (int p = 0; p < cs->world; ++p) {
    out->nrm[p] = (float4*)cs->peer[p][kBufNrm];
    out->rsd[p] = (float2*)cs->peer[p][kBufRsd];
    out->perm[p] = (int*)cs->peer[p][kBufPerm];
  }
}

void comm_free(cab_ctx* ctx) {
  CommState* cs = ctx->comm;
  if (!cs) return;
  cudaStreamSynchronize(ctx->stream);
  close_peers(ctx);
  if (ctx->b_nrm.p == cs->own[kBufWork].p) {  // the arena gets its own normals buffer back on the next pass
    ctx->b_nrm = DevBuf{};
    ctx->have_normals = false;
  }
  for (auto& b : cs->own)
    if (b.p) cudaFree(b.p);
  if (cs->stage.p) cudaFree(cs->stage.p);
  if (cs->nccl) {
    NcclApi* api = nccl_api(nullptr);
    if (api) api->CommDestroy(cs->nccl);
  }
  delete cs;
  ctx->comm = nullptr;
}

// ---- the pieces of one exchanged step, called by cab_step_normals_rsd (cab_api.cu) ---------------------------------
bool comm_active(const cab_ctx* ctx) { return ctx->comm && ctx->comm->world > 1; }

int comm_step_begin(cab_ctx* ctx) {  // after the slab build
  CommState* cs = ctx->comm;
  if (int rc = comm_prepare(ctx, ctx->n)) return rc;
  cs->seq++;
  cs->armed = true;
  cs->last_n = ctx->n;
  stamp_kernel<<<1, 32, 0, ctx->stream>>>((SyncBlock*)cs->own[kBufSync].p);
  CAB_LAUNCH_CHECK(ctx);
  if (cs->layout == CAB_COMM_LAYOUT_INPUT_RANGES) return CAB_OK;  // a result's place follows from its input index alone
  post_count_kernel<<<1, kMaxPeers, 0, ctx->stream>>>(slab_info_device(ctx), peer_sync(cs), cs->rank, cs->world, cs->seq);
  CAB_LAUNCH_CHECK(ctx);
  return CAB_OK;
}

// Input-range layout, before the slab build: where the selection pass leaves the defaults of this rank's non-finite points
void comm_range_defaults(cab_ctx* ctx, double plane_radius) {
  ctx->range_defaults = RangeDefaults{};
  CommState* cs = ctx->comm;
  if (!cs || cs->world <= 1 || cs->layout != CAB_COMM_LAYOUT_INPUT_RANGES || !cs->connected) return;
  ctx->range_defaults.nrm = (float4*)cs->own[kBufNrm].p;
  ctx->range_defaults.rsd = (float2*)cs->own[kBufRsd].p;
  ctx->range_defaults.lo = (int)(ctx->n * cs->rank / cs->world);
  ctx->range_defaults.hi = (int)(ctx->n * (cs->rank + 1) / cs->world);
  ctx->range_defaults.radius = (float)plane_radius;
}

int comm_step_before_push(cab_ctx* ctx) {  // after the normals pass
  CommState* cs = ctx->comm;
  if (cs->layout == CAB_COMM_LAYOUT_INPUT_RANGES) return CAB_OK;
  wait_counts_kernel<<<1, kMaxPeers, 0, ctx->stream>>>((SlabInfo*)slab_info_device(ctx), (SyncBlock*)cs->own[kBufSync].p, cs->rank,
                                                     cs->world, cs->seq, cs->timeout_ns);
  CAB_LAUNCH_CHECK(ctx);
  return CAB_OK;
}

int comm_halo_send(cab_ctx* ctx) {  // after the normals pass
  CommState* cs = ctx->comm;
  cudaStream_t st = ctx->stream;
  if (cs->rank + 1 < cs->world) {
    halo_send_kernel<<<128, 256, 0, st>>>(slab_info_device(ctx), (const float4*)cs->own[kBufWork].p,
                                          (float4*)cs->peer[cs->rank + 1][kBufWork]);
    CAB_LAUNCH_CHECK(ctx);
  }
  post_flag_kernel<<<1, kMaxPeers, 0, st>>>(peer_sync(cs), cs->rank, cs->world, cs->seq, 2);
  CAB_LAUNCH_CHECK(ctx);
  return CAB_OK;
}

int comm_halo_receive(cab_ctx* ctx, cudaStream_t st) {  // before the boundary packets of the RSD pass
  CommState* cs = ctx->comm;
  SlabInfo* info = (SlabInfo*)slab_info_device(ctx);
  SyncBlock* mine = (SyncBlock*)cs->own[kBufSync].p;
  if (cs->rank + 1 < cs->world) {  // the rank above has finished its normals: fetch its bottom layer
    wait_flag_kernel<<<1, kMaxPeers, 0, st>>>(info, nullptr, mine, cs->world, cs->seq, 2, cs->timeout_ns, cs->rank + 1);
    CAB_LAUNCH_CHECK(ctx);
    halo_fetch_kernel<<<128, 256, 0, st>>>(info, (float4*)cs->own[kBufWork].p, (const float4*)cs->peer[cs->rank + 1][kBufWork]);
    CAB_LAUNCH_CHECK(ctx);
  }
  if (cs->rank > 0) {  // the rank below has stored its top layer into my lower halo
    wait_flag_kernel<<<1, kMaxPeers, 0, st>>>(info, nullptr, mine, cs->world, cs->seq, 2, cs->timeout_ns, cs->rank - 1);
    CAB_LAUNCH_CHECK(ctx);
  }
  return CAB_OK;
}

int comm_step_end(cab_ctx* ctx, double plane_radius) {  // after the RSD kernel
  CommState* cs = ctx->comm;
  cs->armed = false;
  cs->last_plane_radius = plane_radius;
  SyncBlock* mine = (SyncBlock*)cs->own[kBufSync].p;
  post_done_kernel<<<1, kMaxPeers, 0, ctx->stream>>>(peer_sync(cs), mine, slab_info_device(ctx), cs->rank, cs->world, cs->seq);
  CAB_LAUNCH_CHECK(ctx);
  wait_flag_kernel<<<1, kMaxPeers, 0, ctx->stream>>>((SlabInfo*)slab_info_device(ctx), nullptr, mine, cs->world, cs->seq, 0,
                                                   cs->timeout_ns);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_step + kStepComm + 64, mine->busy, 2 * kMaxPeers * sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, ctx->stream));
  // the waits of this step report a peer that did not answer in SlabInfo::error: read it after the last of them
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_step + kStepComm + 8, &slab_info_device(ctx)->error, sizeof(int), cudaMemcpyDeviceToHost,
                                ctx->stream));
  if (cs->layout == CAB_COMM_LAYOUT_REPLICATED)
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_step + kStepComm, &((SyncBlock*)cs->own[kBufSync].p)->total, 8, cudaMemcpyDeviceToHost,
                                  ctx->stream));
  return CAB_OK;
}

int comm_step_finish(cab_ctx* ctx) {  // after the step's synchronisation
  CommState* cs = ctx->comm;
  if (cs->layout == CAB_COMM_LAYOUT_REPLICATED) {
    unsigned long long total;
    std::memcpy(&total, ctx->h_step + kStepComm, 8);
    cs->last_total = (int64_t)total;
  } else {
    cs->last_total = ctx->n * (cs->rank + 1) / cs->world - ctx->n * cs->rank / cs->world;
  }
  std::memcpy(&ctx->slab_info.error, ctx->h_step + kStepComm + 8, sizeof(int));
  if (std::getenv("CAB_COMM_DEBUG"))
    fprintf(stderr, "[cab_comm] rank %d seq %u q0 %d q1 %d gbase %d total %lld n_selected %d exchange %d error %d\n", cs->rank, cs->seq,
            ctx->slab_info.q0, ctx->slab_info.q1, ctx->slab_info.gbase, (long long)cs->last_total, ctx->slab_info.n_selected,
            ctx->slab_info.exchange, ctx->slab_info.error);
  if (ctx->slab_info.error) return fail(ctx, CAB_ERR_STATE, "cab_comm: a peer did not answer within %.1f s", cs->timeout_ns * 1e-9);
  unsigned long long fb[2 * kMaxPeers];
  std::memcpy(fb, ctx->h_step + kStepComm + 64, sizeof(fb));
  const unsigned long long* busy = fb;
  const unsigned long long* count = fb + kMaxPeers;
  // every finite point is exactly one rank's query -- unless the ranks cut the rows differently
  unsigned long long answered = 0;
  for (int p = 0; p < cs->world; ++p) answered += count[p];
  const unsigned long long finite = ctx->dom_count.empty() ? 0ull : (unsigned long long)ctx->dom_count[0];
  if (answered != finite)
    return fail(ctx, CAB_ERR_STATE, "cab_comm: the ranks answered %llu queries of %llu finite points -- they did not cut the "
                                    "rows alike (different cab_comm_set_feedback / clouds / parameters on the ranks?)",
                answered, finite);
  if (cs->feedback) {
    // measured-time feedback: a rank that took longer than the mean gets a smaller share of the modelled cost next
    // time.  Every rank reads the same numbers and does the same arithmetic, so the cuts stay consistent.
    double mean = 0;
    bool ok = true;
    for (int p = 0; p < cs->world; ++p) {
      mean += (double)busy[p];
      ok = ok && busy[p] > 0 && busy[p] < (1ull << 40);
    }
    mean /= cs->world;
    // (steps of less than half a millisecond are launch latency, not work: their times say nothing about the cuts)
    if (ok && mean > 5e5) {
      double sum = 0;
      for (int p = 0; p < cs->world; ++p) {
        // (one step moves a share by a tenth at most: a single mistimed step -- a host hiccup that leaves a rank's kernels
        // waiting -- then cannot hand a rank half the cloud; a real imbalance of a few percent still settles in two steps)
        const double f = std::min(std::max(std::pow(mean / (double)busy[p], 0.75), 0.9), 1.1);
        double s = cs->share[p] * f;
        s = std::min(std::max(s, 0.25 / cs->world), 4.0 / cs->world);
        cs->share[p] = s;
        sum += s;
      }
      for (int p = 0; p < cs->world; ++p) cs->share[p] /= sum;
    }
  }
  return CAB_OK;
}

// before the slab build: the ranks' shares of the modelled cost
void comm_shares(cab_ctx* ctx) {
  ctx->shard_cum.clear();
  CommState* cs = ctx->comm;
  if (!cs || cs->world <= 1) return;
  ctx->shard_cum.assign((size_t)cs->world + 1, 0.0);
  double acc = 0;
  for (int p = 0; p < cs->world; ++p) {
    acc += cs->share[p];
    ctx->shard_cum[(size_t)p + 1] = acc;
  }
  ctx->shard_cum[(size_t)cs->world] = 1.0;
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_comm_get_id(char id[CAB_COMM_ID_BYTES]) {
  std::string why;
  NcclApi* api = nccl_api(&why);
  if (!api) return fail(nullptr, CAB_ERR_CUDA, "cab_comm_get_id: %s", why.c_str());
  std::memset(id, 0, CAB_COMM_ID_BYTES);
  int e = api->GetUniqueId(id);
  if (e) return fail(nullptr, CAB_ERR_CUDA, "ncclGetUniqueId failed: %s", api->GetErrorString(e));
  return CAB_OK;
}

int cab_comm_init(cab_ctx* ctx, const char id[CAB_COMM_ID_BYTES], int32_t rank, int32_t world) {
  if (!ctx || !id) return CAB_ERR_ARG;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(ctx, CAB_ERR_ARG, "cab_comm_init: bad rank %d / world %d (at most %d ranks)", rank, world, kMaxPeers);
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  std::string why;
  NcclApi* api = nccl_api(&why);
  if (!api) return fail(ctx, CAB_ERR_CUDA, "cab_comm_init: %s", why.c_str());
  comm_free(ctx);
  Id128 uid;
  std::memcpy(uid.b, id, CAB_COMM_ID_BYTES);
  void* comm = nullptr;
  int e = api->CommInitRank(&comm, world, uid, rank);
  if (e) return nccl_fail(ctx, "ncclCommInitRank", e);
  ctx->comm = new_comm_state(rank, world);
  ctx->comm->nccl = comm;
  return cab_set_shard(ctx, rank, world);
}

int cab_comm_init_local(cab_ctx** ctxs, int32_t world) {
  if (!ctxs || world < 1 || world > kMaxPeers) return CAB_ERR_ARG;
  auto group = std::make_shared<LocalGroup>();
  group->world = world;
  for (int r = 0; r < world; ++r) {
    cab_ctx* ctx = ctxs[r];
    if (!ctx) return CAB_ERR_ARG;
    cudaSetDevice(ctx->device);
    comm_free(ctx);
    ctx->comm = new_comm_state(r, world);
    ctx->comm->local = group;
    if (int rc = cab_set_shard(ctx, r, world)) return rc;
  }
  return CAB_OK;
}

int cab_comm_reserve(cab_ctx* ctx, int32_t rank, int32_t world, int64_t max_points, void* blob) {
  if (!ctx || !blob) return CAB_ERR_ARG;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(ctx, CAB_ERR_ARG, "cab_comm_reserve: bad rank %d / world %d (at most %d ranks)", rank, world, kMaxPeers);
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->comm || ctx->comm->world != world || ctx->comm->rank != rank) {
    comm_free(ctx);
    ctx->comm = new_comm_state(rank, world);
  }
  if (int rc = cab_set_shard(ctx, rank, world)) return rc;
  std::memset(blob, 0, CAB_COMM_BLOB_BYTES);
  return make_blob(ctx, max_points, (CommBlob*)blob);
}

int cab_comm_connect(cab_ctx* ctx, const void* blobs) {
  if (!ctx || !blobs) return CAB_ERR_ARG;
  if (!ctx->comm) return fail(ctx, CAB_ERR_STATE, "cab_comm_connect: call cab_comm_reserve first");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<CommBlob> all(ctx->comm->world);
  for (int p = 0; p < ctx->comm->world; ++p)
    std::memcpy(&all[p], (const char*)blobs + (size_t)p * CAB_COMM_BLOB_BYTES, sizeof(CommBlob));
  return connect_blobs(ctx, all.data());
}

int cab_comm_set_layout(cab_ctx* ctx, int32_t layout) {
  if (!ctx) return CAB_ERR_ARG;
  if (layout != CAB_COMM_LAYOUT_REPLICATED && layout != CAB_COMM_LAYOUT_INPUT_RANGES)
    return fail(ctx, CAB_ERR_ARG, "cab_comm_set_layout: unknown layout %d", layout);
  if (!ctx->comm) return fail(ctx, CAB_ERR_STATE, "cab_comm_set_layout: this context belongs to no group");
  ctx->comm->layout = layout;
  ctx->comm->last_total = 0;
  return CAB_OK;
}

int cab_comm_set_feedback(cab_ctx* ctx, int32_t on) {
  if (!ctx) return CAB_ERR_ARG;
  if (!ctx->comm) return fail(ctx, CAB_ERR_STATE, "cab_comm_set_feedback: this context belongs to no group");
  ctx->comm->feedback = on != 0;
  for (int p = 0; p < kMaxPeers; ++p) ctx->comm->share[p] = 1.0 / ctx->comm->world;
  return CAB_OK;
}

int cab_comm_set_shares(cab_ctx* ctx, const double* shares, int32_t count) {
  if (!ctx || !shares) return CAB_ERR_ARG;
  if (!ctx->comm) return fail(ctx, CAB_ERR_STATE, "cab_comm_set_shares: this context belongs to no group");
  if (count != ctx->comm->world) return fail(ctx, CAB_ERR_ARG, "cab_comm_set_shares: %d shares for %d ranks", count, ctx->comm->world);
  double sum = 0;
  for (int p = 0; p < count; ++p) {
    if (!(shares[p] > 0) || !std::isfinite(shares[p])) return fail(ctx, CAB_ERR_ARG, "cab_comm_set_shares: shares must be positive");
    sum += shares[p];
  }
  for (int p = 0; p < count; ++p) ctx->comm->share[p] = shares[p] / sum;
  return CAB_OK;
}

int cab_comm_free(cab_ctx* ctx) {
  if (!ctx) return CAB_ERR_ARG;
  cudaSetDevice(ctx->device);
  comm_free(ctx);
  return cab_set_shard(ctx, 0, 1);
}

int cab_comm_upload_cloud(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride) {
  if (!ctx) return CAB_ERR_ARG;
  if (!comm_active(ctx)) return cab_upload_cloud(ctx, xyz, n, stride);
  if (stride != 3) return fail(ctx, CAB_ERR_ARG, "cab_comm_upload_cloud: packed xyz (stride 3) only");
  if (n < 0 || (n > 0 && !xyz)) return fail(ctx, CAB_ERR_ARG, "cab_comm_upload_cloud: bad cloud");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = comm_prepare(ctx, n)) return rc;
  CommState* cs = ctx->comm;
  cudaStream_t st = ctx->stream;
  const int64_t lo = n * cs->rank / cs->world, hi = n * (cs->rank + 1) / cs->world;
  const size_t off = (size_t)lo * 3, bytes = (size_t)(hi - lo) * 3 * sizeof(float);
  float* mine = (float*)cs->own[kBufCloud].p;
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
  if (bytes) {
    CAB_CUDA(ctx, cudaMemcpyAsync(mine + off, xyz + off, bytes, cudaMemcpyHostToDevice, st));
    for (int k = 1; k < cs->world; ++k) {  // start with the right-hand neighbour: the ranks' copies do not collide
      const int p = (cs->rank + k) % cs->world;
      CAB_CUDA(ctx, cudaMemcpyAsync((float*)cs->peer[p][kBufCloud] + off, mine + off, bytes, cudaMemcpyDefault, st));
    }
  }
  cs->cloud_seq++;
  if (int rc = reserve(ctx, ctx->b_misc, 64)) return rc;
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_misc.p, 0, 4, st));
  post_flag_kernel<<<1, kMaxPeers, 0, st>>>(peer_sync(cs), cs->rank, cs->world, cs->cloud_seq, 1);
  CAB_LAUNCH_CHECK(ctx);
  wait_flag_kernel<<<1, kMaxPeers, 0, st>>>(nullptr, (int*)ctx->b_misc.p, (SyncBlock*)cs->own[kBufSync].p, cs->world, cs->cloud_seq, 1,
                                          cs->timeout_ns);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_step + kStepComm + 16, ctx->b_misc.p, 4, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[6], ctx->ev[7]));
  int err;
  std::memcpy(&err, ctx->h_step + kStepComm + 16, 4);
  if (err) return fail(ctx, CAB_ERR_STATE, "cab_comm_upload_cloud: a peer's slice did not arrive within %.1f s", cs->timeout_ns * 1e-9);
  return cab_set_cloud_device(ctx, mine, n, 3);
}

int cab_comm_download_range(cab_ctx* ctx, int64_t j0, int64_t j1, float* nxyz_curv, float* r_min, float* r_max) {
  if (!ctx) return CAB_ERR_ARG;
  if (!comm_active(ctx) || !ctx->comm->connected)
    return fail(ctx, CAB_ERR_STATE, "cab_comm_download_range: no exchanged step has run on this context");
  if (j0 < 0 || j1 < j0 || j1 > ctx->n) return fail(ctx, CAB_ERR_ARG, "cab_comm_download_range: bad range");
  if ((r_min == nullptr) != (r_max == nullptr)) return fail(ctx, CAB_ERR_ARG, "r_min and r_max must be given together");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  CommState* cs = ctx->comm;
  const int m = (int)(j1 - j0);
  if (m == 0 || (!nxyz_curv && !r_min)) return CAB_OK;
  cudaStream_t st = ctx->stream;
  if (cs->layout == CAB_COMM_LAYOUT_INPUT_RANGES) {
    const int64_t lo = ctx->n * cs->rank / cs->world, hi = ctx->n * (cs->rank + 1) / cs->world;
    if (j0 < lo || j1 > hi)
      return fail(ctx, CAB_ERR_ARG, "cab_comm_download_range: [%lld, %lld) is not inside this rank's input range [%lld, %lld) "
                                    "(input-range layout)", (long long)j0, (long long)j1, (long long)lo, (long long)hi);
    if (cs->last_n != ctx->n) return fail(ctx, CAB_ERR_STATE, "cab_comm_download_range: no exchanged step has run on this cloud");
    const float4* nrm = (const float4*)cs->own[kBufNrm].p + (j0 - lo);
    const float2* rsd = (const float2*)cs->own[kBufRsd].p + (j0 - lo);
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
    if (!ctx->slab) {  // no finite point anywhere: nothing was selected, nothing pushed
      if (nxyz_curv)
        if (int rc = reserve(ctx, ctx->b_out4, (size_t)m * sizeof(float4))) return rc;
      if (r_min) {
        if (int rc = reserve(ctx, ctx->b_out1a, (size_t)m * sizeof(float))) return rc;
        if (int rc = reserve(ctx, ctx->b_out1b, (size_t)m * sizeof(float))) return rc;
      }
      fill_defaults_kernel<<<(m + 255) / 256, 256, 0, st>>>(nxyz_curv ? (float4*)ctx->b_out4.p : nullptr,
                                                            r_min ? (float*)ctx->b_out1a.p : nullptr,
                                                            r_min ? (float*)ctx->b_out1b.p : nullptr, m, (float)cs->last_plane_radius);
      CAB_LAUNCH_CHECK(ctx);
      if (nxyz_curv) CAB_CUDA(ctx, cudaMemcpyAsync(nxyz_curv, ctx->b_out4.p, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, st));
      if (r_min) {
        CAB_CUDA(ctx, cudaMemcpyAsync(r_min, ctx->b_out1a.p, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
        CAB_CUDA(ctx, cudaMemcpyAsync(r_max, ctx->b_out1b.p, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
      }
    } else {
      // the normals leave as they are; the radii pairs are split while the normals travel
      if (r_min) {
        if (int rc = reserve(ctx, ctx->b_out1a, (size_t)m * sizeof(float))) return rc;
        if (int rc = reserve(ctx, ctx->b_out1b, (size_t)m * sizeof(float))) return rc;
        split_radii_kernel<<<(m + 255) / 256, 256, 0, st>>>(rsd, m, (float*)ctx->b_out1a.p, (float*)ctx->b_out1b.p);
        CAB_LAUNCH_CHECK(ctx);
      }
      if (nxyz_curv) CAB_CUDA(ctx, cudaMemcpyAsync(nxyz_curv, nrm, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, st));
      if (r_min) {
        CAB_CUDA(ctx, cudaMemcpyAsync(r_min, ctx->b_out1a.p, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
        CAB_CUDA(ctx, cudaMemcpyAsync(r_max, ctx->b_out1b.p, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
      }
    }
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[6], ctx->ev[7]));
    return CAB_OK;
  }
  if (nxyz_curv)
    if (int rc = reserve(ctx, ctx->b_out4, (size_t)m * sizeof(float4))) return rc;
  if (r_min) {
    if (int rc = reserve(ctx, ctx->b_out1a, (size_t)m * sizeof(float))) return rc;
    if (int rc = reserve(ctx, ctx->b_out1b, (size_t)m * sizeof(float))) return rc;
  }
  float4* o4 = nxyz_curv ? (float4*)ctx->b_out4.p : nullptr;
  float* oa = r_min ? (float*)ctx->b_out1a.p : nullptr;
  float* ob = r_min ? (float*)ctx->b_out1b.p : nullptr;
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
  // points that are nobody's query (non-finite) keep the values the single-GPU path gives them
  fill_defaults_kernel<<<(m + 255) / 256, 256, 0, st>>>(o4, oa, ob, m, (float)cs->last_plane_radius);
  CAB_LAUNCH_CHECK(ctx);
  const long long total = cs->last_total;
  if (total > 0) {
    range_unpermute_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const int*)cs->own[kBufPerm].p, total, (int)j0, (int)j1,
                                                                          (const float4*)cs->own[kBufNrm].p,
                                                                          (const float2*)cs->own[kBufRsd].p, o4, oa, ob);
    CAB_LAUNCH_CHECK(ctx);
  }
  if (nxyz_curv) CAB_CUDA(ctx, cudaMemcpyAsync(nxyz_curv, o4, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, st));
  if (r_min) {
    CAB_CUDA(ctx, cudaMemcpyAsync(r_min, oa, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaMemcpyAsync(r_max, ob, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[6], ctx->ev[7]));
  return CAB_OK;
}

void* cab_comm_device_ptr(cab_ctx* ctx, int32_t which, int64_t* count) {
  if (!ctx || !ctx->comm || !ctx->comm->connected) return nullptr;
  if (count) *count = ctx->comm->last_total;
  const bool ranges = ctx->comm->layout == CAB_COMM_LAYOUT_INPUT_RANGES;
  switch (which) {
    case CAB_BUF_NRM_SORTED: return ranges ? nullptr : ctx->comm->own[kBufNrm].p;
    case CAB_BUF_RSD_SORTED: return ranges ? nullptr : ctx->comm->own[kBufRsd].p;
    case CAB_BUF_PERM: return ranges ? nullptr : ctx->comm->own[kBufPerm].p;
    case CAB_BUF_NRM_INPUT_RANGE: return ranges ? ctx->comm->own[kBufNrm].p : nullptr;
    case CAB_BUF_RSD_INPUT_RANGE: return ranges ? ctx->comm->own[kBufRsd].p : nullptr;
    default: return nullptr;
  }
}

int cab_comm_allreduce_i32(cab_ctx* ctx, int32_t* values, int64_t count) {
  if (!ctx || (count > 0 && !values)) return CAB_ERR_ARG;
  if (!ctx->comm || ctx->comm->world <= 1 || count == 0) return CAB_OK;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  CommState* cs = ctx->comm;
  cudaStream_t st = ctx->stream;
  const size_t bytes = (size_t)count * 4;
  if (cs->nccl) {
    NcclApi* api = nccl_api(nullptr);
    if (int rc = reserve(ctx, cs->stage, bytes)) return rc;
    CAB_CUDA(ctx, cudaMemcpyAsync(cs->stage.p, values, bytes, cudaMemcpyHostToDevice, st));
    int e = api->AllReduce(cs->stage.p, cs->stage.p, (size_t)count, kNcclInt32, kNcclSum, cs->nccl, st);
    if (e) return nccl_fail(ctx, "ncclAllReduce", e);
    CAB_CUDA(ctx, cudaMemcpyAsync(values, cs->stage.p, bytes, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    return CAB_OK;
  }
  if (cs->local) {  // contexts of one process: sum on the host, integers are order independent
    static std::mutex m;
    static std::vector<int64_t> acc;
    {
      std::lock_guard<std::mutex> lk(m);
      if ((int64_t)acc.size() != count) acc.assign((size_t)count, 0);
      for (int64_t i = 0; i < count; ++i) acc[(size_t)i] += values[i];
    }
    cs->local->barrier();
    {
      std::lock_guard<std::mutex> lk(m);
      for (int64_t i = 0; i < count; ++i) values[i] = (int32_t)acc[(size_t)i];
    }
    cs->local->barrier();
    if (cs->rank == 0) {
      std::lock_guard<std::mutex> lk(m);
      acc.clear();
    }
    cs->local->barrier();
    return CAB_OK;
  }
  return fail(ctx, CAB_ERR_STATE, "cab_comm_allreduce_i32: needs cab_comm_init (NCCL) or cab_comm_init_local");
}

}  // extern "C"
