// cab_grid.cu -- device-built search structure that replaces the reference's kd-tree
// (cloud_kdtree::KdTreeANN at cloud_algos/src/radius_estimation.cpp:107, pcl::KdTreeFLANN at
// color_chlac/include/color_chlac/grsd_colorCHLAC_tools.hpp:79,175).
//
// Layout in HBM (DESIGN.md "Data layout"):
//   rows      : (domain, cell z, cell y) tubes along x; points sorted by (row, fine x)
//   pos       : float4[n]   sorted positions
//   perm      : int[n]      sorted position -> input index
//   cell_start: int[cells+1] dense CSR over (row, cell x)
//   packets   : <=32 consecutive sorted points of one row = the work unit of one warp
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "cab_internal.cuh"

namespace cab {

namespace {

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned u) {
  unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
#ifdef __CUDA_ARCH__
  f = __uint_as_float(b);
#else
  std::memcpy(&f, &b, 4);
#endif
  return f;
}

__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}

struct Chunk {
  int domain, begin, end;
};

// bounds[d] = {minx,miny,minz,maxx,maxy,maxz (ordered-uint encoded), n_finite, unused}
__global__ void __launch_bounds__(256) bounds_kernel(const float* __restrict__ xyz, int stride,
                                                     const Chunk* __restrict__ chunks,
                                                     unsigned* __restrict__ bounds) {
  const Chunk ch = chunks[blockIdx.x];
  unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  unsigned cnt = 0;
  for (int i = ch.begin + threadIdx.x; i < ch.end; i += blockDim.x) {
    const float* p = xyz + (size_t)i * stride;
    float x = p[0], y = p[1], z = p[2];
    if (finite3(x, y, z)) {
      unsigned e[3] = {f2ord(x), f2ord(y), f2ord(z)};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = min(mn[a], e[a]);
        mx[a] = max(mx[a], e[a]);
      }
      ++cnt;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(kFull, mn[a]);
    mx[a] = __reduce_max_sync(kFull, mx[a]);
  }
  cnt = __reduce_add_sync(kFull, cnt);
  if ((threadIdx.x & 31) == 0) {
    unsigned* b = bounds + 8 * (size_t)ch.domain;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(b + a, mn[a]);
      atomicMax(b + 3 + a, mx[a]);
    }
    atomicAdd(b + 6, cnt);
  }
}

__device__ __forceinline__ int find_domain(const int* __restrict__ domoff, int n_domains, int i) {
  int lo = 0, hi = n_domains;  // largest d with domoff[d] <= i
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (domoff[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

template <typename KeyT>
__global__ void __launch_bounds__(256) key_kernel(const float* __restrict__ xyz, int stride, int n,
                                                  const int* __restrict__ domoff, int n_domains,
                                                  const Domain* __restrict__ domains, float inv_cell,
                                                  unsigned long long sentinel_row, int xbits,
                                                  KeyT* __restrict__ keys,
                                                  int* __restrict__ vals, int* __restrict__ cellcnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + (size_t)i * stride;
  float x = p[0], y = p[1], z = p[2];
  unsigned long long key = sentinel_row << xbits;
  if (finite3(x, y, z)) {
    int d = (n_domains > 1) ? find_domain(domoff, n_domains, i) : 0;
    const Domain dm = domains[d];
    int cy = cell_coord(y, dm.oy, inv_cell, dm.ny);
    int cz = cell_coord(z, dm.oz, inv_cell, dm.nz);
    int xf = xfine_coord(x, dm.ox, inv_cell, dm.nx, dm.xshift);
    long long row_local = (long long)cz * dm.ny + cy;
    key = ((unsigned long long)(dm.row_base + row_local) << xbits) | (unsigned)xf;
    atomicAdd(cellcnt + dm.cell_base + row_local * dm.nx + (xf >> dm.xshift), 1);
  }
  keys[i] = (KeyT)key;
  if (vals) vals[i] = i;
}

__global__ void __launch_bounds__(256) reorder_kernel(const float* __restrict__ xyz, int stride, int n,
                                                      const int* __restrict__ perm,
                                                      float4* __restrict__ pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + (size_t)perm[i] * stride;
  pos[i] = make_float4(p[0], p[1], p[2], 0.f);
}

constexpr int kSpanCap = 4;  // a packet never covers more than about this many cells along x

__device__ __forceinline__ int domain_of_cell(const Domain* __restrict__ domains, int n_domains, long long c) {
  int lo = 0, hi = n_domains;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (domains[mid].cell_base <= c) lo = mid; else hi = mid;
  }
  return lo;
}

// A segment is a maximal run of non-empty cells of one row.  Packets are cut inside segments, so a
// packet never straddles an empty stretch of a row (a row that crosses two distant surfaces would
// otherwise produce packets whose x window -- and candidate set -- spans everything in between).
// One thread per cell; the thread of a segment's first cell measures it and sets its packet count.
__global__ void __launch_bounds__(256) segment_kernel(const Domain* __restrict__ domains, int n_domains,
                                                      long long n_cells, const int* __restrict__ cell_start,
                                                      int* __restrict__ segpk, int* __restrict__ seglen) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cells) return;
  int npk = 0, len = 0;
  if (c < n_cells && cell_start[c + 1] > cell_start[c]) {
    const Domain dm = domains[domain_of_cell(domains, n_domains, c)];
    const int cx = (int)((c - dm.cell_base) % dm.nx);
    if (cx == 0 || cell_start[c] == cell_start[c - 1]) {  // previous cell of the row is empty
      len = 1;
      while (cx + len < dm.nx && cell_start[c + len + 1] > cell_start[c + len]) ++len;
      const int pts = cell_start[c + len] - cell_start[c];
      npk = max((pts + kWarp - 1) / kWarp, (len + kSpanCap - 1) / kSpanCap);
      npk = min(npk, pts);
    }
  }
  segpk[c] = npk;
  seglen[c] = len;
}

// one thread per packet: find its segment by binary search over the per-cell packet prefix
__global__ void __launch_bounds__(256) fill_packets_kernel(const Domain* __restrict__ domains, int n_domains,
                                                           long long n_cells, const int* __restrict__ cell_start,
                                                           const int* __restrict__ packet_base,
                                                           const int* __restrict__ seglen, int n_packets,
                                                           Packet* __restrict__ packets) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_packets) return;
  long long lo = 0, hi = n_cells;  // largest cell with packet_base[cell] <= p: the segment's first cell
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (packet_base[mid] <= p) lo = mid; else hi = mid;
  }
  const long long c0 = lo;
  const int d = domain_of_cell(domains, n_domains, c0);
  const Domain dm = domains[d];
  const int start = cell_start[c0];
  const long long pts = cell_start[c0 + seglen[c0]] - start;
  const int pb = packet_base[c0], npk = packet_base[c0 + 1] - pb, k = p - pb;
  const int a = start + (int)(k * pts / npk), b = start + (int)((k + 1) * pts / npk);
  packets[p] = Packet{a, b - a, (int)((c0 - dm.cell_base) / dm.nx), d};
}

// Cells of a row covered by the sorted positions [s0, s1] (both inside the row): the cell holding s is the
// last one whose start is <= s.  Needs only the cell table, not the sorted points.
__device__ __forceinline__ int cell_of_position(const int* __restrict__ row_start, int nx, int s) {
  int lo = 0, hi = nx;  // largest c in [0, nx) with row_start[c] <= s
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (row_start[mid] <= s) lo = mid; else hi = mid;
  }
  return lo;
}

// cost model of a packet for the multi-GPU split: the candidates its queries will be tested against
__global__ void __launch_bounds__(256) packet_cost_kernel(const Domain* __restrict__ domains, const Packet* __restrict__ packets,
                                                          int n_packets, const int* __restrict__ cell_start,
                                                          long long* __restrict__ cost) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_packets) return;
  const Packet pk = packets[p];
  const Domain dm = domains[pk.domain];
  const int cy = pk.row_local % dm.ny, cz = pk.row_local / dm.ny;
  const int* row = cell_start + dm.cell_base + (long long)pk.row_local * dm.nx;
  const int cxlo = max(cell_of_position(row, dm.nx, pk.start) - 1, 0);
  const int cxhi = min(cell_of_position(row, dm.nx, pk.start + pk.count - 1) + 1, dm.nx - 1);
  long long c = 0;
  for (int z = max(cz - 1, 0); z <= min(cz + 1, dm.nz - 1); ++z)
    for (int y = max(cy - 1, 0); y <= min(cy + 1, dm.ny - 1); ++y) {
      const long long b = dm.cell_base + ((long long)z * dm.ny + y) * dm.nx;
      c += cell_start[b + cxhi + 1] - cell_start[b + cxlo];
    }
  cost[p] = c + 64;  // + a constant per packet (setup, fit, eigen-solve)
}

// cells touched by the shard's own packets
__global__ void __launch_bounds__(256) mark_cells_kernel(const Domain* __restrict__ domains, const Packet* __restrict__ packets,
                                                         int p0, int p1, const int* __restrict__ cell_start,
                                                         unsigned char* __restrict__ cell_flag) {
  const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p1) return;
  const Packet pk = packets[p];
  const Domain dm = domains[pk.domain];
  const long long base = dm.cell_base + (long long)pk.row_local * dm.nx;
  const int c0 = cell_of_position(cell_start + base, dm.nx, pk.start);
  const int c1 = cell_of_position(cell_start + base, dm.nx, pk.start + pk.count - 1);
  for (int c = c0; c <= c1; ++c) cell_flag[base + c] = 1;
}

// A packet needs normals on this rank if it is the shard's own or touches a cell adjacent (3x3x3) to
// a cell of the shard: every candidate of the shard's RSD pass then has a locally computed normal
// and no exchange between the two passes is needed.  With need_cell != nullptr the cells that hold the
// candidates of every such packet are marked as well: only their points have to be sorted on this rank.
__global__ void __launch_bounds__(256) flag_halo_kernel(const Domain* __restrict__ domains, const Packet* __restrict__ packets,
                                                        int n_packets, int p0, int p1, const int* __restrict__ cell_start,
                                                        const unsigned char* __restrict__ cell_flag,
                                                        unsigned char* __restrict__ flag, unsigned char* __restrict__ need_cell) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_packets) return;
  unsigned char f = (p >= p0 && p < p1) ? 1 : 0;
  const Packet pk = packets[p];
  const Domain dm = domains[pk.domain];
  const int cy = pk.row_local % dm.ny, cz = pk.row_local / dm.ny;
  const int* row = cell_start + dm.cell_base + (long long)pk.row_local * dm.nx;
  const int c0 = max(cell_of_position(row, dm.nx, pk.start) - 1, 0);
  const int c1 = min(cell_of_position(row, dm.nx, pk.start + pk.count - 1) + 1, dm.nx - 1);
  if (!f) {
    for (int z = max(cz - 1, 0); z <= min(cz + 1, dm.nz - 1) && !f; ++z)
      for (int y = max(cy - 1, 0); y <= min(cy + 1, dm.ny - 1) && !f; ++y) {
        const long long base = dm.cell_base + ((long long)z * dm.ny + y) * dm.nx;
        for (int c = c0; c <= c1; ++c)
          if (cell_flag[base + c]) { f = 1; break; }
      }
  }
  flag[p] = f;
  if (f && need_cell)
    for (int z = max(cz - 1, 0); z <= min(cz + 1, dm.nz - 1); ++z)
      for (int y = max(cy - 1, 0); y <= min(cy + 1, dm.ny - 1); ++y) {
        const long long base = dm.cell_base + ((long long)z * dm.ny + y) * dm.nx;
        for (int c = c0; c <= c1; ++c) need_cell[base + c] = 1;
      }
}

// ---- sharded sort (multi-GPU, one domain): only the points of the needed cells are sorted on this rank ----
// cell of a sort key: row = key >> xbits, cell x = (key & mask) >> xshift
template <typename KeyT>
__device__ __forceinline__ long long cell_of_key(KeyT key, int xbits, int xshift, int nx, unsigned long long n_rows) {
  const unsigned long long row = (unsigned long long)key >> xbits;
  if (row >= n_rows) return -1;  // sentinel row: non-finite point
  const unsigned xf = (unsigned)((unsigned long long)key & ((1ull << xbits) - 1ull));
  return (long long)row * nx + (xf >> xshift);
}

// flag functor of the selection: does point i lie in a cell this shard needs?
template <typename KeyT>
struct NeedPoint {
  const KeyT* keys;
  const unsigned char* need_cell;
  int xbits, xshift, nx;
  unsigned long long n_rows;
  __device__ __forceinline__ bool operator()(int i) const {
    const long long c = cell_of_key(keys[i], xbits, xshift, nx, n_rows);
    return c >= 0 && need_cell[c] != 0;
  }
};

template <typename KeyT>
__global__ void __launch_bounds__(256) gather_keys_kernel(const KeyT* __restrict__ keys, const int* __restrict__ idx, int m,
                                                          KeyT* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) out[j] = keys[idx[j]];
}

__global__ void __launch_bounds__(256) masked_count_kernel(const int* __restrict__ cellcnt, const unsigned char* __restrict__ need_cell,
                                                           long long n_cells, int* __restrict__ out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= n_cells) out[c] = (c < n_cells && need_cell[c]) ? cellcnt[c] : 0;
}

// j-th element of the sorted subset -> its position in the full sorted order: every needed cell is complete
// in the subset, so the rank inside the cell is j minus the subset's start of that cell.
template <typename KeyT>
__global__ void __launch_bounds__(256) scatter_sorted_kernel(const float* __restrict__ xyz, int stride, int m,
                                                             const KeyT* __restrict__ skeys, const int* __restrict__ svals,
                                                             int xbits, int xshift, int nx, unsigned long long n_rows,
                                                             const int* __restrict__ cell_start, const int* __restrict__ sub_start,
                                                             float4* __restrict__ pos, int* __restrict__ perm) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const long long c = cell_of_key(skeys[j], xbits, xshift, nx, n_rows);
  const int gp = cell_start[c] + (j - sub_start[c]);
  const int i = svals[j];
  const float* p = xyz + (size_t)i * stride;
  pos[gp] = make_float4(p[0], p[1], p[2], 0.f);
  perm[gp] = i;
}

// split[g] = first packet whose inclusive cost prefix reaches g/world of the total
__global__ void split_kernel(const long long* __restrict__ cum, int n_packets, int world, int* __restrict__ split) {
  const int g = threadIdx.x;
  if (g > world) return;
  if (g == 0) { split[0] = 0; return; }
  if (g == world) { split[world] = n_packets; return; }
  const long long target = cum[n_packets - 1] / world * g;
  int lo = 0, hi = n_packets;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cum[mid] < target) lo = mid + 1; else hi = mid;
  }
  split[g] = lo;
}

}  // namespace

// Per-domain bounding boxes and finite-point counts of the uploaded cloud -> ctx->dom_bounds / dom_count.
int compute_bounds(cab_ctx* ctx) {
  const int nd = ctx->n_domains;
  cudaStream_t st = ctx->stream;
  // ---- bounds per domain --------------------------------------------------------------
  std::vector<Chunk> chunks;
  const int kChunk = 1 << 13;  // ~2400 blocks for 20 M points: enough loads in flight to stream at HBM speed
  for (int d = 0; d < nd; ++d) {
    int b = ctx->dom_offsets[d], e = ctx->dom_offsets[d + 1];
    if (b == e) chunks.push_back(Chunk{d, b, e});
    for (int s = b; s < e; s += kChunk) chunks.push_back(Chunk{d, s, std::min(e, s + kChunk)});
  }
  if (int rc = reserve(ctx, ctx->b_misc, chunks.size() * sizeof(Chunk))) return rc;
  if (int rc = reserve(ctx, ctx->b_bounds, (size_t)nd * 8 * sizeof(unsigned))) return rc;
  if (int rc = reserve_pinned(ctx, std::max((size_t)nd * 8 * sizeof(unsigned), chunks.size() * sizeof(Chunk)) + 64))
    return rc;
  std::memcpy(ctx->h_pin, chunks.data(), chunks.size() * sizeof(Chunk));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->b_misc.p, ctx->h_pin, chunks.size() * sizeof(Chunk), cudaMemcpyHostToDevice, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  {
    std::vector<unsigned> init((size_t)nd * 8);
    for (int d = 0; d < nd; ++d) {
      unsigned* b = init.data() + 8 * (size_t)d;
      b[0] = b[1] = b[2] = 0xffffffffu;
      b[3] = b[4] = b[5] = 0u;
      b[6] = b[7] = 0u;
    }
    std::memcpy(ctx->h_pin, init.data(), init.size() * sizeof(unsigned));
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->b_bounds.p, ctx->h_pin, init.size() * sizeof(unsigned), cudaMemcpyHostToDevice, st));
  }
  if (!chunks.empty()) {
    bounds_kernel<<<(unsigned)chunks.size(), 256, 0, st>>>(ctx->xyz_in, ctx->stride, (const Chunk*)ctx->b_misc.p,
                                                           (unsigned*)ctx->b_bounds.p);
    CAB_LAUNCH_CHECK(ctx);
  }
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, ctx->b_bounds.p, (size_t)nd * 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));

  const unsigned* hb = (const unsigned*)ctx->h_pin;
  ctx->dom_bounds.assign((size_t)nd * 6, 0.f);
  ctx->dom_count.assign(nd, 0u);
  for (int d = 0; d < nd; ++d) {
    ctx->dom_count[d] = hb[8 * d + 6];
    if (hb[8 * d + 6])
      for (int a = 0; a < 6; ++a) ctx->dom_bounds[6 * (size_t)d + a] = ord2f(hb[8 * d + a]);
  }
  return CAB_OK;
}

int build_grid(cab_ctx* ctx, float cell) {
  if (!ctx->have_cloud) return fail(ctx, CAB_ERR_STATE, "cab_build_grid: no cloud uploaded");
  if (!(cell > 0.f) || !std::isfinite(cell)) return fail(ctx, CAB_ERR_ARG, "cab_build_grid: cell must be > 0");
  const int n = (int)ctx->n;
  const int nd = ctx->n_domains;
  cudaStream_t st = ctx->stream;
  ctx->have_grid = ctx->have_normals = ctx->have_rsd = false;
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));

  if (int rc = compute_bounds(ctx)) return rc;

  // ---- domain geometry (host, tiny) ---------------------------------------------------
  double max_abs = 0;
  int64_t n_valid = 0;
  for (int d = 0; d < nd; ++d) {
    if (ctx->dom_count[d] == 0) continue;
    n_valid += ctx->dom_count[d];
    for (int a = 0; a < 6; ++a) max_abs = std::max(max_abs, (double)std::fabs(ctx->dom_bounds[6 * (size_t)d + a]));
  }
  // effective cell: >= requested + slack for fp32 rounding of (v - origin) * inv_cell
  double ulp = std::ldexp(1.0, (max_abs > 0 ? (int)std::floor(std::log2(max_abs)) : 0) - 23);
  double cell_eff = (double)cell * (1.0 + 1.0 / 1024.0) + 8.0 * ulp;
  ctx->cell = cell;
  ctx->domains.assign(nd, Domain{});
  int64_t rows = 0, cells = 0;
  const int64_t budget = ctx->cfg.max_table_cells > 0 ? ctx->cfg.max_table_cells : ((int64_t)1 << 28);
  // The cell table is dense.  A cloud too spread out for it at the requested cell size (an outdoor scan with a 2 cm
  // radius) gets coarser cells instead of an error: any edge >= the radius is correct, the passes just test more
  // candidates per query.
  for (int attempt = 0;; ++attempt) {
    rows = cells = 0;
    double want = 0;  // cells this edge would need, as a double (it may overflow int64)
    bool too_big = false;
    for (int d = 0; d < nd; ++d) {
      Domain& dm = ctx->domains[d];
      dm.row_base = rows;
      dm.cell_base = cells;
      if (ctx->dom_count[d] == 0) {
        dm.ox = dm.oy = dm.oz = 0.f;
        dm.nx = dm.ny = dm.nz = 1;
      } else {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
          lo[a] = ctx->dom_bounds[6 * (size_t)d + a];
          hi[a] = ctx->dom_bounds[6 * (size_t)d + 3 + a];
        }
        dm.ox = lo[0];
        dm.oy = lo[1];
        dm.oz = lo[2];
        double ext[3] = {(double)hi[0] - lo[0], (double)hi[1] - lo[1], (double)hi[2] - lo[2]};
        double nn[3];
        for (int a = 0; a < 3; ++a) nn[a] = std::floor(ext[a] / cell_eff) + 2;  // +1 slack cell
        want += nn[0] * nn[1] * nn[2];
        if (nn[0] > (double)(1 << kXBits) || nn[1] * nn[2] > std::ldexp(1.0, 40) || want > (double)budget) {
          too_big = true;
          continue;
        }
        dm.nx = (int)nn[0];
        dm.ny = (int)nn[1];
        dm.nz = (int)nn[2];
      }
      rows += (int64_t)dm.ny * dm.nz;
      cells += (int64_t)dm.ny * dm.nz * dm.nx;
    }
    if (!too_big) break;
    if (attempt >= 60 || !std::isfinite(cell_eff))
      return fail(ctx, CAB_ERR_OOM, "cab_build_grid: no cell size fits the dense cell table (budget %lld cells)", (long long)budget);
    cell_eff *= std::max(1.26, std::min(8.0, std::cbrt(want / (double)budget) * 1.02));
  }
  ctx->cell_eff = (float)cell_eff;
  if ((double)ctx->cell_eff < cell_eff) ctx->cell_eff = std::nextafter(ctx->cell_eff, INFINITY);
  ctx->inv_cell = 1.0f / ctx->cell_eff;
  ctx->n_rows = rows;
  ctx->n_cells = cells;
  ctx->n_valid = (int)n_valid;
  // key = (row << xbits) | x_fine.  32-bit keys whenever the rows leave room for the cell x plus at
  // least two sub-cell bits (the radix sort then moves 8 instead of 12 bytes per point and pass).
  int row_bits = 1;
  while (((int64_t)1 << row_bits) <= rows) ++row_bits;  // rows itself is the sentinel row
  int nx_max = 1;
  for (int d = 0; d < nd; ++d) nx_max = std::max(nx_max, ctx->domains[d].nx);
  int nx_bits = 0;
  while ((1 << nx_bits) < nx_max) ++nx_bits;
  const bool key32 = row_bits + nx_bits + 2 <= 32;
  const int xbits = key32 ? std::min(kXBits, 32 - row_bits) : kXBits;
  for (int d = 0; d < nd; ++d) {
    Domain& dm = ctx->domains[d];
    dm.xshift = 0;
    while (dm.xshift < 8 && ((int64_t)dm.nx << (dm.xshift + 1)) <= ((int64_t)1 << xbits)) dm.xshift++;
  }

  if (int rc = reserve(ctx, ctx->b_domains, nd * sizeof(Domain))) return rc;
  if (int rc = reserve_pinned(ctx, nd * sizeof(Domain))) return rc;
  std::memcpy(ctx->h_pin, ctx->domains.data(), nd * sizeof(Domain));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->b_domains.p, ctx->h_pin, nd * sizeof(Domain), cudaMemcpyHostToDevice, st));

  // ---- keys + cell histogram ----------------------------------------------------------
  const size_t ncell1 = (size_t)cells + 1;
  if (int rc = reserve(ctx, ctx->b_keys[0], (size_t)n * 8)) return rc;
  if (int rc = reserve(ctx, ctx->b_keys[1], (size_t)n * 8)) return rc;
  if (int rc = reserve(ctx, ctx->b_vals[0], (size_t)n * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_perm, (size_t)n * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_cellcnt, ncell1 * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_cellstart, ncell1 * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_pos, (size_t)std::max(n, 1) * sizeof(float4))) return rc;
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_cellcnt.p, 0, ncell1 * 4, st));
  if (n > 0) {
    if (key32)
      key_kernel<unsigned><<<(n + 255) / 256, 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, (const int*)ctx->b_domoff.p, nd,
                                                            (const Domain*)ctx->b_domains.p, ctx->inv_cell,
                                                            (unsigned long long)rows, xbits, (unsigned*)ctx->b_keys[0].p,
                                                            (int*)ctx->b_vals[0].p, (int*)ctx->b_cellcnt.p);
    else
      key_kernel<unsigned long long><<<(n + 255) / 256, 256, 0, st>>>(
          ctx->xyz_in, ctx->stride, n, (const int*)ctx->b_domoff.p, nd, (const Domain*)ctx->b_domains.p, ctx->inv_cell,
          (unsigned long long)rows, xbits, (unsigned long long*)ctx->b_keys[0].p, (int*)ctx->b_vals[0].p,
          (int*)ctx->b_cellcnt.p);
    CAB_LAUNCH_CHECK(ctx);
  }

  // ---- cell table, segments and packets: all of it follows from the cell histogram alone -----------
  const int end_bit = std::min(key32 ? 32 : 64, xbits + row_bits);
  size_t tmp_sort = 0, tmp_scan1 = 0;
  if (key32)
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, n, 0, end_bit, st);
  else
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, n, 0, end_bit, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan1, (const int*)nullptr, (int*)nullptr, (int)ncell1, st);
  size_t tmp_bytes = std::max(tmp_sort, tmp_scan1);
  if (int rc = reserve(ctx, ctx->b_cubtmp, tmp_bytes + 16)) return rc;
  CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp_bytes, (const int*)ctx->b_cellcnt.p,
                                              (int*)ctx->b_cellstart.p, (int)ncell1, st));
  ctx->tm.kernel_launches += 2;
  if (int rc = reserve(ctx, ctx->b_rowpk, ncell1 * 4 * 3)) return rc;
  int* segpk = (int*)ctx->b_rowpk.p;
  int* packet_base = segpk + ncell1;
  int* seglen = packet_base + ncell1;
  segment_kernel<<<(unsigned)((ncell1 + 255) / 256), 256, 0, st>>>((const Domain*)ctx->b_domains.p, nd, cells,
                                                                  (const int*)ctx->b_cellstart.p, segpk, seglen);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp_bytes, segpk, packet_base, (int)ncell1, st));
  ctx->tm.kernel_launches += 2;
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, packet_base + cells, 4, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->n_packets = *(const int*)ctx->h_pin;
  if (int rc = reserve(ctx, ctx->b_packets, (size_t)std::max(ctx->n_packets, 1) * sizeof(Packet))) return rc;
  if (ctx->n_packets > 0) {
    fill_packets_kernel<<<(unsigned)((ctx->n_packets + 255) / 256), 256, 0, st>>>(
        (const Domain*)ctx->b_domains.p, nd, cells, (const int*)ctx->b_cellstart.p, packet_base, seglen, ctx->n_packets,
        (Packet*)ctx->b_packets.p);
    CAB_LAUNCH_CHECK(ctx);
  }
  // ---- cost-balanced shard boundaries and the halo (multi-GPU only) ---------------------------------
  ctx->shard_splits.clear();
  ctx->n_halo_packets = -1;
  // one domain: the rank sorts only the points it needs (own cells, the cells of the halo packets and
  // the candidates of those); the full sorted order keeps its layout, unneeded positions stay unwritten
  const bool sharded_sort = ctx->shard_world > 1 && ctx->n_packets > 0 && nd == 1 && n > 0;
  unsigned char* need_cell = nullptr;
  if (ctx->shard_world > 1 && ctx->n_packets > 0) {
    const int np = ctx->n_packets, w = ctx->shard_world;
    if (w + 1 > 64) return fail(ctx, CAB_ERR_ARG, "cab_set_shard: world > 63 not supported");
    if (int rc = reserve(ctx, ctx->b_pcost, (size_t)np * 16 + (w + 1) * 4 + 64)) return rc;
    long long* cost = (long long*)ctx->b_pcost.p;
    long long* cum = cost + np;
    int* split = (int*)(cum + np);
    size_t tmp_cost = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tmp_cost, (const long long*)nullptr, (long long*)nullptr, np, st);
    if (int rc = reserve(ctx, ctx->b_cubtmp, tmp_cost + 16)) return rc;
    packet_cost_kernel<<<(np + 255) / 256, 256, 0, st>>>((const Domain*)ctx->b_domains.p, (const Packet*)ctx->b_packets.p, np,
                                                         (const int*)ctx->b_cellstart.p, cost);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cub::DeviceScan::InclusiveSum(ctx->b_cubtmp.p, tmp_cost, cost, cum, np, st));
    split_kernel<<<1, 64, 0, st>>>(cum, np, w, split);
    CAB_LAUNCH_CHECK(ctx);
    ctx->tm.kernel_launches += 2;
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, split, (w + 1) * 4, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->shard_splits.assign((const int*)ctx->h_pin, (const int*)ctx->h_pin + w + 1);
    // halo packet list for the normals pass (+ the cells whose points this rank needs)
    int p0, p1;
    packet_range(ctx, &p0, &p1);
    if (int rc = reserve(ctx, ctx->b_rowflag, 2 * (size_t)cells + (size_t)np + 64)) return rc;
    if (int rc = reserve(ctx, ctx->b_halo_list, ((size_t)np + 4) * 4)) return rc;
    unsigned char* cell_flag = (unsigned char*)ctx->b_rowflag.p;
    unsigned char* pflag = cell_flag + cells;
    if (sharded_sort) need_cell = pflag + np + 32;
    int* list = (int*)ctx->b_halo_list.p;
    CAB_CUDA(ctx, cudaMemsetAsync(cell_flag, 0, (size_t)cells, st));
    if (need_cell) CAB_CUDA(ctx, cudaMemsetAsync(need_cell, 0, (size_t)cells, st));
    if (p1 > p0) {
      mark_cells_kernel<<<(p1 - p0 + 255) / 256, 256, 0, st>>>((const Domain*)ctx->b_domains.p, (const Packet*)ctx->b_packets.p, p0, p1,
                                                              (const int*)ctx->b_cellstart.p, cell_flag);
      CAB_LAUNCH_CHECK(ctx);
    }
    flag_halo_kernel<<<(np + 255) / 256, 256, 0, st>>>((const Domain*)ctx->b_domains.p, (const Packet*)ctx->b_packets.p, np, p0, p1,
                                                       (const int*)ctx->b_cellstart.p, cell_flag, pflag, need_cell);
    CAB_LAUNCH_CHECK(ctx);
    size_t tmp_sel = 0;
    thrust::counting_iterator<int> ids(0);
    cub::DeviceSelect::Flagged(nullptr, tmp_sel, ids, pflag, list, list + np, np, st);
    if (int rc = reserve(ctx, ctx->b_cubtmp, tmp_sel + 16)) return rc;
    CAB_CUDA(ctx, cub::DeviceSelect::Flagged(ctx->b_cubtmp.p, tmp_sel, ids, pflag, list, list + np, np, st));
    ctx->tm.kernel_launches += 2;
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, list + np, 4, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->n_halo_packets = *(const int*)ctx->h_pin;
  }

  // ---- radix sort by (row, fine x) ----------------------------------------------------
  if (n > 0 && !sharded_sort) {
    if (int rc = reserve(ctx, ctx->b_cubtmp, tmp_sort + 16)) return rc;
    if (key32)
      CAB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_cubtmp.p, tmp_sort, (const unsigned*)ctx->b_keys[0].p,
                                                    (unsigned*)ctx->b_keys[1].p, (const int*)ctx->b_vals[0].p,
                                                    (int*)ctx->b_perm.p, n, 0, end_bit, st));
    else
      CAB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_cubtmp.p, tmp_sort, (const unsigned long long*)ctx->b_keys[0].p,
                                                    (unsigned long long*)ctx->b_keys[1].p, (const int*)ctx->b_vals[0].p,
                                                    (int*)ctx->b_perm.p, n, 0, end_bit, st));
    ctx->tm.kernel_launches += 1 + (end_bit + 7) / 8;  // onesweep: histogram + one kernel per digit
    reorder_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, (const int*)ctx->b_perm.p,
                                                    (float4*)ctx->b_pos.p);
    CAB_LAUNCH_CHECK(ctx);
  } else if (sharded_sort) {
    const Domain& dm = ctx->domains[0];
    if (int rc = reserve(ctx, ctx->b_vals[1], (size_t)n * 4)) return rc;
    if (int rc = reserve(ctx, ctx->b_keys[2], (size_t)n * 8)) return rc;
    if (int rc = reserve(ctx, ctx->b_vals[2], (size_t)n * 4 + 16)) return rc;
    if (int rc = reserve(ctx, ctx->b_substart, ncell1 * 4 * 2)) return rc;
    int* masked = (int*)ctx->b_substart.p;
    int* sub_start = masked + ncell1;
    int* d_m = (int*)((char*)ctx->b_vals[2].p + (size_t)n * 4);  // number of selected points
    masked_count_kernel<<<(unsigned)((ncell1 + 255) / 256), 256, 0, st>>>((const int*)ctx->b_cellcnt.p, need_cell, cells, masked);
    CAB_LAUNCH_CHECK(ctx);
    if (int rc = reserve(ctx, ctx->b_cubtmp, tmp_bytes + 16)) return rc;
    CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp_bytes, masked, sub_start, (int)ncell1, st));
    // stable selection of the indices of the needed points (flags computed on the fly from the keys)
    thrust::counting_iterator<int> all_points(0);
    size_t tmp_selk = 0;
    NeedPoint<unsigned> need32{(const unsigned*)ctx->b_keys[0].p, need_cell, xbits, dm.xshift, dm.nx, (unsigned long long)rows};
    NeedPoint<unsigned long long> need64{(const unsigned long long*)ctx->b_keys[0].p, need_cell, xbits, dm.xshift, dm.nx,
                                         (unsigned long long)rows};
    if (key32) cub::DeviceSelect::If(nullptr, tmp_selk, all_points, (int*)nullptr, d_m, n, need32, st);
    else cub::DeviceSelect::If(nullptr, tmp_selk, all_points, (int*)nullptr, d_m, n, need64, st);
    size_t tmp_sel2 = std::max(tmp_selk, tmp_sort);
    if (int rc = reserve(ctx, ctx->b_cubtmp, tmp_sel2 + 16)) return rc;
    if (key32) CAB_CUDA(ctx, cub::DeviceSelect::If(ctx->b_cubtmp.p, tmp_sel2, all_points, (int*)ctx->b_vals[2].p, d_m, n, need32, st));
    else CAB_CUDA(ctx, cub::DeviceSelect::If(ctx->b_cubtmp.p, tmp_sel2, all_points, (int*)ctx->b_vals[2].p, d_m, n, need64, st));
    ctx->tm.kernel_launches += 3;
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, d_m, 4, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_perm.p, 0xff, (size_t)n * 4, st));  // unneeded positions: perm = -1
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    const int m = *(const int*)ctx->h_pin;
    if (m > 0) {
      if (key32)
        gather_keys_kernel<unsigned><<<(m + 255) / 256, 256, 0, st>>>((const unsigned*)ctx->b_keys[0].p, (const int*)ctx->b_vals[2].p, m,
                                                                     (unsigned*)ctx->b_keys[2].p);
      else
        gather_keys_kernel<unsigned long long><<<(m + 255) / 256, 256, 0, st>>>(
            (const unsigned long long*)ctx->b_keys[0].p, (const int*)ctx->b_vals[2].p, m, (unsigned long long*)ctx->b_keys[2].p);
      CAB_LAUNCH_CHECK(ctx);
      if (key32)
        CAB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_cubtmp.p, tmp_sel2, (const unsigned*)ctx->b_keys[2].p,
                                                      (unsigned*)ctx->b_keys[1].p, (const int*)ctx->b_vals[2].p,
                                                      (int*)ctx->b_vals[1].p, m, 0, end_bit, st));
      else
        CAB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_cubtmp.p, tmp_sel2, (const unsigned long long*)ctx->b_keys[2].p,
                                                      (unsigned long long*)ctx->b_keys[1].p, (const int*)ctx->b_vals[2].p,
                                                      (int*)ctx->b_vals[1].p, m, 0, end_bit, st));
      ctx->tm.kernel_launches += 1 + (end_bit + 7) / 8;
      if (key32)
        scatter_sorted_kernel<unsigned><<<(m + 255) / 256, 256, 0, st>>>(
            ctx->xyz_in, ctx->stride, m, (const unsigned*)ctx->b_keys[1].p, (const int*)ctx->b_vals[1].p, xbits, dm.xshift, dm.nx,
            (unsigned long long)rows, (const int*)ctx->b_cellstart.p, sub_start, (float4*)ctx->b_pos.p, (int*)ctx->b_perm.p);
      else
        scatter_sorted_kernel<unsigned long long><<<(m + 255) / 256, 256, 0, st>>>(
            ctx->xyz_in, ctx->stride, m, (const unsigned long long*)ctx->b_keys[1].p, (const int*)ctx->b_vals[1].p, xbits, dm.xshift,
            dm.nx, (unsigned long long)rows, (const int*)ctx->b_cellstart.p, sub_start, (float4*)ctx->b_pos.p, (int*)ctx->b_perm.p);
      CAB_LAUNCH_CHECK(ctx);
    }
    ctx->tm.n_sorted = m;
  }
  if (!sharded_sort) ctx->tm.n_sorted = ctx->n_valid;
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.build_ms, ctx->ev[0], ctx->ev[1]));
  ctx->tm.n_points = ctx->n;
  ctx->tm.n_valid = ctx->n_valid;
  ctx->tm.n_packets = ctx->n_packets;
  ctx->tm.n_rows = rows;
  ctx->tm.n_cells = cells;
  ctx->have_grid = true;
  return CAB_OK;
}

GridView grid_view(const cab_ctx* ctx) {
  GridView g{};
  g.pos = (const float4*)ctx->b_pos.p;
  g.perm = (const int*)ctx->b_perm.p;
  g.cell_start = (const int*)ctx->b_cellstart.p;
  g.packets = (const Packet*)ctx->b_packets.p;
  g.domains = (const Domain*)ctx->b_domains.p;
  g.n_valid = ctx->n_valid;
  g.n_packets = ctx->n_packets;
  g.inv_cell = ctx->inv_cell;
  return g;
}

void packet_range(const cab_ctx* ctx, int* p0, int* p1) {
  if ((int)ctx->shard_splits.size() == ctx->shard_world + 1) {
    *p0 = ctx->shard_splits[ctx->shard_rank];
    *p1 = ctx->shard_splits[ctx->shard_rank + 1];
    return;
  }
  int64_t P = ctx->n_packets;
  *p0 = (int)(P * ctx->shard_rank / ctx->shard_world);
  *p1 = (int)(P * (ctx->shard_rank + 1) / ctx->shard_world);
}

}  // namespace cab
