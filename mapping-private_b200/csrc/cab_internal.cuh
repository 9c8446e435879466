// cab_internal.cuh -- shared declarations of libcloudalgos_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "cloud_algos_b200.h"

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#error "libcloudalgos_b200 is written for sm_100a (B200) only"
#endif

namespace cab {

constexpr int kWarp = 32;
constexpr int kXBits = 20;            // sub-cell x field of the sort key
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kTruncBins = 64;          // d2 histogram bins of the max_nn selection
constexpr int kTruncCap = 16;           // candidates of the target bin the truncated fast RSD pass settles from a list
constexpr int kStatSlots = 64;         // spread counters (avoids same-address atomic serialisation)
constexpr size_t kStatBytes = (kStatSlots * 2 + 2) * sizeof(unsigned long long);  // + the packet work counter

// One independent neighbourhood domain: the whole cloud, or one segmented cluster.
struct Domain {
  float ox, oy, oz;      // grid origin
  int nx, ny, nz;        // cells per axis
  int xshift;            // x_fine >> xshift == cell x
  int xwide;             // a cell is 2^xwide cell edges wide along x (0 unless the table would dwarf the cloud)
  int swap_yz;           // 1: the grid's y slot holds the physical z axis and its z slot the physical y axis
                         // (the slowest-varying slot gets the longer of the two extents: thinner shard halos)
  int64_t row_base;      // first row id of this domain (row = row_base + cz*ny + cy)
  int64_t cell_base;     // first cell id (cell = cell_base + (cz*ny+cy)*nx + cx); negative for a slab table, whose
                         // first cell is the first cell of row row_lo
  int row_lo, row_hi;    // rows [row_lo, row_hi) of this domain (row = cz*ny + cy) are present in the cell table:
                         // all of them, or the window one rank of a multi-GPU run sorts (cab_set_shard)
};

// What one rank of a sharded run (cab_set_shard, one cloud) knows about its slab.  Rows are (cz, cy) tubes, row id =
// cz*ny + cy; a rank answers the queries of whole rows [own_lo, own_hi), computes normals for one more layer of rows on
// either side (halo: every candidate of its RSD pass then has a locally computed normal) and sorts two layers on either
// side (window: the candidates of the halo rows).  Written on the device by the slab build, read there by the kernels
// that follow it in the stream; the host copy is valid after the next synchronisation.
struct SlabInfo {
  int own_lo, own_hi;    // rows
  int halo_lo, halo_hi;
  int win_lo, win_hi;
  int n_selected;        // points of the window = entries of the local sorted arrays
  int n_packets;         // packets of the window
  int p0, p1;            // own packets
  int ph0, ph1;          // own + halo packets
  int q0, q1;            // own queries: positions in the local sorted arrays
  int gbase;             // position of q0 in the concatenation of all ranks' own slices (result exchange)
  int error;             // != 0: a peer did not answer in time (result exchange)
  // Halo exchange (cab_step_normals_rsd of a group, every rank at least two layers thick): the halo rows' normals are
  // not recomputed but received -- the window is own +- ONE layer, halo = own, and the RSD pass runs in two launches:
  // the packets that read no halo row first (r_int), the two boundary layers (r_blo, r_bhi) once the neighbours'
  // normals have arrived.
  int exchange;          // 1: the mode above
  int top_src;           // local position of the first point of the top layer of own rows (sent to the rank above)
  int r_int[2];          // packets of own rows at least one layer away from either end
  int r_blo[2], r_bhi[2];  // packets of the bottom / top layer of own rows
};

// 32-query work unit: consecutive sorted points of one row.
struct Packet {
  int start, count;
  int row_local;  // cz*ny + cy inside the domain
  int domain;
};

// Read-only view of the search structure handed to kernels.
struct GridView {
  const float4* pos;        // sorted positions (x,y,z,0)
  const int* perm;          // sorted -> input index
  const int* cell_start;    // n_cells + 1
  const Packet* packets;
  const Domain* domains;
  int n_valid;              // finite points (sorted positions [0, n_valid))
  int n_packets;
  float inv_cell;           // 1 / effective cell edge
  float xscale_unused;
};

// ---- the documented epsilon rule -------------------------------------------------------
// d2 = (dx*dx + dy*dy) + dz*dz in fp32 with no FMA contraction; d = p - q.
__device__ __forceinline__ float d2_rule(float px, float py, float pz, float qx, float qy, float qz) {
  float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Cell coordinate along one axis; identical code in the key kernel and in the query kernels.
__device__ __forceinline__ int cell_coord(float v, float origin, float inv_cell, int n) {
  int c = (int)floorf(__fmul_rn(__fsub_rn(v, origin), inv_cell));
  return min(max(c, 0), n - 1);
}
// Row coordinates (cy, cz) of a point: which physical axis fills which slot is the domain's choice.
__device__ __forceinline__ void row_cells(const Domain& dm, float y, float z, float inv_cell, int& cy, int& cz) {
  cy = cell_coord(dm.swap_yz ? z : y, dm.oy, inv_cell, dm.ny);
  cz = cell_coord(dm.swap_yz ? y : z, dm.oz, inv_cell, dm.nz);
}
// Is row (y, z) of the domain inside the grid and present in the cell table?
__device__ __forceinline__ bool row_in_table(const Domain& dm, int y, int z) {
  if (y < 0 || y >= dm.ny || z < 0 || z >= dm.nz) return false;
  const int row = z * dm.ny + y;
  return row >= dm.row_lo && row < dm.row_hi;
}
// Fine x coordinate: steps of 2^-(xshift - xwide) of the cell edge; x_fine >> xshift is the cell along x, which is
// 2^xwide edges wide (xwide > 0 only for clouds much sparser than the cell table, see build_grid).
__device__ __forceinline__ int xfine_coord(float x, float ox, float inv_cell, int nx, int xshift, int xwide) {
  float s = __fmul_rn(__fsub_rn(x, ox), __fmul_rn(inv_cell, (float)(1 << (xshift - xwide))));
  int c = (int)floorf(s);
  return min(max(c, 0), (nx << xshift) - 1);
}

// slots of cab_ctx::h_step
constexpr size_t kStepSlab = 0, kStepThr = 512, kStepStats0 = 1024, kStepStats1 = 3072, kStepComm = 5120, kStepBytes = 8192;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// Result exchange of a multi-GPU run (cab_comm.cu): every rank's copy of the concatenated result arrays, as pointers
// valid on this rank (own memory, a peer device of the same process, or CUDA IPC mappings of the other processes).
constexpr int kMaxPeers = 16;
struct PushTargets {
  int world;   // 0: no exchange
  int layout;  // CAB_COMM_LAYOUT_*: replicated sorted-order arrays on every rank, or input-order ranges (one owner per point)
  long long n; // points of the cloud (input-range layout: rank p owns input indices [lo[p], lo[p + 1]))
  int lo[kMaxPeers + 1];
  float4* nrm[kMaxPeers];
  float2* rsd[kMaxPeers];
  int* perm[kMaxPeers];
};

// Input-range layout: what the slab selection writes for the points of this rank's input range that are nobody's query
// (non-finite coordinates): the values the single-GPU path gives them.
struct RangeDefaults {
  float4* nrm = nullptr;
  float2* rsd = nullptr;
  int lo = 0, hi = 0;
  float radius = 0.f;
};

struct SvmState;   // cab_svm.cu
struct CommState;  // cab_comm.cu

}  // namespace cab

struct cab_ctx {
  cab_config cfg{};
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // device->host copies that overlap the next kernel (cab_normals_rsd)
  cudaEvent_t ev[12]{};
  cudaEvent_t ev_ready = nullptr;      // results staged for the copy stream
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // work forked onto copy_stream inside a step and joined again
  std::string err;
  cab_timings tm{};

  // cloud
  int64_t n = 0;            // points
  int n_valid = 0;
  int n_domains = 0;
  bool have_cloud = false, have_grid = false, have_normals = false, have_rsd = false;
  bool kcount_valid = false;  // b_kcount holds the untruncated in-radius neighbour counts (self included) of radius kcount_r
  float kcount_r = 0.f;
  bool trunc_hist_valid = false;  // b_thr_idx / b_thr_flag hold the d2-histogram codes of radius kcount_r, max_nn trunc_hist_max_nn
  int trunc_hist_max_nn = 0;      // (taken along by the last normals pass for the truncated fast RSD pass)
  bool cloud_external = false;
  const float* xyz_in = nullptr;  // device, stride floats
  int stride = 3;
  std::vector<int32_t> dom_offsets;  // host copy, n_domains + 1
  std::vector<cab::Domain> domains;  // host copy
  std::vector<float> dom_bounds;     // host copy, 6 per domain (min xyz, max xyz); valid if dom_count > 0
  std::vector<uint32_t> dom_count;   // finite points per domain
  float cell = 0.f, cell_eff = 0.f, inv_cell = 0.f;
  int64_t n_rows = 0, n_cells = 0;
  int n_packets = 0;
  int shard_rank = 0, shard_world = 1;
  bool slab = false;               // the grid is this rank's slab of a sharded run (local sorted arrays)
  cab::SlabInfo slab_info{};       // host copy (valid when slab_info_valid)
  bool slab_info_valid = false;
  int64_t n_sorted = 0;            // entries of the sorted arrays: n, or the slab window's points
  int halo_permille = 470;         // share of the normals pass in a packet's cost (shard balance)
  bool want_halo_exchange = false; // set by cab_step_normals_rsd of a group: build_slab may choose the exchange mode
  bool step_input_order = false;   // CAB_STEP_INPUT_ORDER: the RSD kernel of this step also scatters the results into b_in_nrm / b_in_rsd
  bool have_input_order = false;   // ... and they are there
  // cab_normals_rsd, input-order layout: the two pass kernels also store every query's result at its input index (the
  // arrays the host copies leave from), so no permutation pass stands between a kernel and its device-to-host copy
  float4* fuse_nrm_in = nullptr;
  float* fuse_rmin_in = nullptr;
  float* fuse_rmax_in = nullptr;
  bool defer_sync = false;         // run_normals / run_rsd leave the stream running (cab_step_*: one sync per step)
  std::vector<double> shard_cum;   // empty: equal shares; else world + 1 cumulative shares of the modelled cost (0 ... 1), the
                                   // group's measured-time feedback (cab_comm.cu) -- identical on every rank by construction
  cab::RangeDefaults range_defaults{};  // set by cab_step_normals_rsd of a group in the input-range layout (slab selection)

  // device arena (grow-only)
  cab::DevBuf b_xyz, b_domoff, b_domid, b_bounds, b_domains, b_keys[3], b_vals[3], b_cubtmp, b_pos,
      b_perm, b_cellcnt, b_cellstart, b_rowpk, b_packets, b_nrm, b_nrm_in, b_rsd, b_rdif, b_kcount, b_stats,
      b_out4, b_out1a, b_out1b, b_thr_d2, b_thr_idx, b_misc, b_pcost, b_slab, b_sel, b_stats2, b_sorttmp, b_occ, b_in_nrm, b_in_rsd, b_thr_flag, b_knn_avg, b_knn_done, b_pfh[3], b_cluster;
  std::vector<cab::DevBuf> graveyard;  // outgrown arena buffers (see reserve())
  size_t graveyard_bytes = 0;
  // pinned staging
  void* h_pin = nullptr;
  size_t h_pin_cap = 0;
  unsigned char* h_step = nullptr;  // fixed pinned slots read back once per step (cab::kStep*)

  // GRSD results of the last batch
  cab::DevBuf g_vkeys[2], g_vvals[2], g_cent, g_vcount, g_vrad, g_vlabel, g_voff, g_layout, g_layoff,
      g_vgrid, g_hist, g_vfirst;
  cab::DevBuf g_cnrm, g_invperm, g_sig, g_sigdom, g_color, g_vown;
  bool g_own_valid = false;        // g_vown marks the voxels this rank labelled (cab_grsd_cloud on a sharded context)
  int64_t g_nvox = 0;
  std::vector<int64_t> g_vox_offsets;
  std::vector<int32_t> g_min_div;  // host copy of the voxel grids: min_b[3], div_b[3] per cluster
  float g_leaf = 0.f;
  bool g_have_cnrm = false;        // voxel mean normals computed for the last batch

  cab::SvmState* svm = nullptr;    // SVM model + scaling (cab_svm.cu)
  cab::CommState* comm = nullptr;  // multi-GPU group this context belongs to (cab_comm.cu)
};

namespace cab {

int fail(cab_ctx* ctx, int code, const char* fmt, ...);
int reserve(cab_ctx* ctx, DevBuf& b, size_t bytes);
int reserve_pinned(cab_ctx* ctx, size_t bytes);
GridView grid_view(const cab_ctx* ctx);
void packet_range(const cab_ctx* ctx, int* p0, int* p1);
int build_slab(cab_ctx* ctx, bool key32, int xbits);          // cab_grid.cu: one rank's slab of a sharded run
int finish_slab(cab_ctx* ctx);                                 // host copy of the slab ranges once the stream has drained
const int* slab_packet_range(const cab_ctx* ctx, bool halo);   // device {p0, p1}; null unless the grid is a slab
const SlabInfo* slab_info_device(const cab_ctx* ctx);
int finish_pass_stats(cab_ctx* ctx, int pass);                 // deferred mode: timings + counters of pass 0 (normals) / 1 (RSD)
void read_stats(cab_ctx* ctx, const void* staged = nullptr);  // sums the kStatSlots counters staged in ctx->h_pin (or `staged`) into ctx->tm

#define CAB_CUDA(ctx, call)                                                                  \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return cab::fail((ctx), CAB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                                  \
  } while (0)

#define CAB_LAUNCH_CHECK(ctx)                                                                \
  do {                                                                                       \
    (ctx)->tm.kernel_launches++;                                                             \
    CAB_CUDA((ctx), cudaGetLastError());                                                     \
  } while (0)

// stage entry points (each enqueues on ctx->stream; callers synchronise)
int compute_bounds(cab_ctx* ctx);
int build_grid(cab_ctx* ctx, float cell);
int run_normals(cab_ctx* ctx, float r, int max_nn, const float vp[3], const unsigned char* done = nullptr, int hist_max_nn = 0);
// phase 0: the whole pass; 1: prologue + the packets that read no halo row; 2: the boundary packets + epilogue
int run_rsd(cab_ctx* ctx, double r, int max_nn, int ndiv, double plane_radius, int flags, int phase = 0);
int run_thresholds(cab_ctx* ctx, float r, int max_nn, const unsigned char* done = nullptr, bool halo = false);  // max_nn truncation thresholds (halo: also for a slab's halo packets)
int run_nn_hist(cab_ctx* ctx, float r, int max_nn);  // cab_topk.cu: histogram half of the truncated fast RSD pass
int run_knn_mean(cab_ctx* ctx, int k, float cell_hint, double* avg);  // cab_knn.cu
int run_normals_knn(cab_ctx* ctx, int k, const float vp[3], float cell_hint, float* nxyz_curv);  // cab_knn.cu
int run_pfh(cab_ctx* ctx, double radius, int max_nn, int quantum, int flags, float* out);  // cab_pfh.cu
int64_t run_euclidean_clusters(cab_ctx* ctx, double tolerance, int min_pts, int max_pts, int32_t* labels);  // cab_cluster.cu
int64_t run_neighbors_debug(cab_ctx* ctx, float r, int max_nn, int64_t q0, int64_t q1, int64_t* offsets,
                            int32_t* idx, float* d2, int64_t cap);
int run_grsd_batch(cab_ctx* ctx, float leaf, double r_rsd, int rsd_flags, int32_t* hist21, bool labels_only = false);
int permute_normals_in(cab_ctx* ctx, const float* nx, const float* ny, const float* nz);
int download_results(cab_ctx* ctx, float* n4, float* rmin, float* rmax);
void svm_free(cab_ctx* ctx);
void comm_free(cab_ctx* ctx);                                  // cab_comm.cu
int comm_prepare(cab_ctx* ctx, int64_t n);                     // collective: exchange buffers for n points, mapped on every rank
bool comm_active(const cab_ctx* ctx);                          // the context belongs to a group of more than one rank
int comm_step_begin(cab_ctx* ctx);                             // after the slab build: publish this rank's query count
int comm_step_before_push(cab_ctx* ctx);                       // before the RSD kernel: place the slice in the concatenation
int comm_halo_send(cab_ctx* ctx);                              // after the normals pass: top layer to the rank above, "normals done" to all
int comm_halo_receive(cab_ctx* ctx, cudaStream_t st);          // before the boundary packets: fetch the layer above, wait for the layer below
int comm_step_end(cab_ctx* ctx, double plane_radius);          // after the RSD kernel: completion flags
int comm_step_finish(cab_ctx* ctx);                            // after the step's synchronisation
void comm_shares(cab_ctx* ctx);                                // before the slab build: the ranks' shares (measured-time feedback) -> ctx->shard_cum
void comm_range_defaults(cab_ctx* ctx, double plane_radius);   // input-range layout: arms ctx->range_defaults for the slab selection
void comm_push_targets(const cab_ctx* ctx, PushTargets* out);  // world = 0 unless a result exchange is armed for this step

}  // namespace cab
