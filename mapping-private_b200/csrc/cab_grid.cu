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
#include <thrust/iterator/transform_iterator.h>

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

// sort key of one point (+ its count in the cell histogram); the sentinel row for a non-finite point
__device__ __forceinline__ unsigned long long point_key(float x, float y, float z, const Domain& dm, float inv_cell,
                                                        unsigned long long sentinel_row, int xbits, int* __restrict__ cellcnt) {
  if (!finite3(x, y, z)) return sentinel_row << xbits;
  int cy, cz;
  row_cells(dm, y, z, inv_cell, cy, cz);
  const int xf = xfine_coord(x, dm.ox, inv_cell, dm.nx, dm.xshift, dm.xwide);
  const long long row_local = (long long)cz * dm.ny + cy;
  atomicAdd(cellcnt + dm.cell_base + row_local * dm.nx + (xf >> dm.xshift), 1);
  return ((unsigned long long)(dm.row_base + row_local) << xbits) | (unsigned)xf;
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
  const float x = p[0], y = p[1], z = p[2];
  const int d = (n_domains > 1 && finite3(x, y, z)) ? find_domain(domoff, n_domains, i) : 0;
  keys[i] = (KeyT)point_key(x, y, z, domains[d], inv_cell, sentinel_row, xbits, cellcnt);
  if (vals) vals[i] = i;
}

// ---- vector forms for the common layout: one domain, packed xyz (stride 3), 16-byte aligned base ----
// A thread owns 4 consecutive points = 3 x LDG.128 (48 bytes in flight per thread: the scalar form leaves
// too few bytes in flight to stream at HBM speed).
struct Pts4 {
  float x[4], y[4], z[4];
};
__device__ __forceinline__ Pts4 load_pts4(const float4* __restrict__ v, int g) {
  const float4 a = __ldg(v + 3 * (size_t)g), b = __ldg(v + 3 * (size_t)g + 1), c = __ldg(v + 3 * (size_t)g + 2);
  Pts4 p;
  p.x[0] = a.x; p.y[0] = a.y; p.z[0] = a.z;
  p.x[1] = a.w; p.y[1] = b.x; p.z[1] = b.y;
  p.x[2] = b.z; p.y[2] = b.w; p.z[2] = c.x;
  p.x[3] = c.y; p.y[3] = c.z; p.z[3] = c.w;
  return p;
}

__global__ void __launch_bounds__(256) bounds_vec_kernel(const float* __restrict__ xyz, int n, unsigned* __restrict__ bounds) {
  const float4* v = reinterpret_cast<const float4*>(xyz);
  const int n_groups = n >> 2;
  unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  unsigned cnt = 0;
  auto take = [&](float x, float y, float z) {
    if (finite3(x, y, z)) {
      const unsigned e[3] = {f2ord(x), f2ord(y), f2ord(z)};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = min(mn[a], e[a]);
        mx[a] = max(mx[a], e[a]);
      }
      ++cnt;
    }
  };
#pragma unroll 2
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += gridDim.x * blockDim.x) {
    const Pts4 p = load_pts4(v, g);
#pragma unroll
    for (int k = 0; k < 4; ++k) take(p.x[k], p.y[k], p.z[k]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // the last n mod 4 points
    const float* p = xyz + 3 * (size_t)(4 * n_groups + threadIdx.x);
    take(p[0], p[1], p[2]);
  }
  __shared__ unsigned sh[8][7];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(kFull, mn[a]);
    mx[a] = __reduce_max_sync(kFull, mx[a]);
  }
  cnt = __reduce_add_sync(kFull, cnt);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      sh[w][a] = mn[a];
      sh[w][3 + a] = mx[a];
    }
    sh[w][6] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 7) {  // one set of atomics per block
    unsigned r = sh[0][threadIdx.x];
    for (int k = 1; k < 8; ++k)
      r = threadIdx.x < 3 ? min(r, sh[k][threadIdx.x]) : threadIdx.x < 6 ? max(r, sh[k][threadIdx.x]) : r + sh[k][threadIdx.x];
    if (threadIdx.x < 3) atomicMin(bounds + threadIdx.x, r);
    else if (threadIdx.x < 6) atomicMax(bounds + threadIdx.x, r);
    else atomicAdd(bounds + 6, r);
  }
}

__global__ void init_bounds_kernel(unsigned* __restrict__ bounds) {
  if (threadIdx.x < 8) bounds[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
}

template <typename KeyT>
__global__ void __launch_bounds__(256) key_vec_kernel(const float* __restrict__ xyz, int n, const Domain* __restrict__ domains,
                                                      float inv_cell, unsigned long long sentinel_row, int xbits,
                                                      KeyT* __restrict__ keys, int* __restrict__ vals, int* __restrict__ cellcnt) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_groups = n >> 2;
  const Domain dm = domains[0];
  if (g < n_groups) {
    const Pts4 p = load_pts4(reinterpret_cast<const float4*>(xyz), g);
    KeyT k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = (KeyT)point_key(p.x[j], p.y[j], p.z[j], dm, inv_cell, sentinel_row, xbits, cellcnt);
    if constexpr (sizeof(KeyT) == 4) {
      reinterpret_cast<uint4*>(keys)[g] = make_uint4(k[0], k[1], k[2], k[3]);
    } else {
      reinterpret_cast<ulonglong2*>(keys)[2 * (size_t)g] = make_ulonglong2(k[0], k[1]);
      reinterpret_cast<ulonglong2*>(keys)[2 * (size_t)g + 1] = make_ulonglong2(k[2], k[3]);
    }
    if (vals) reinterpret_cast<int4*>(vals)[g] = make_int4(4 * g, 4 * g + 1, 4 * g + 2, 4 * g + 3);
  } else if (g == n_groups) {
    for (int i = 4 * n_groups; i < n; ++i) {
      const float* p = xyz + 3 * (size_t)i;
      keys[i] = (KeyT)point_key(p[0], p[1], p[2], dm, inv_cell, sentinel_row, xbits, cellcnt);
      if (vals) vals[i] = i;
    }
  }
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
      // (wide cells say nothing about where in them the points sit -- a row that crosses a tilted surface has all of them
      // within two or three edges -- so there the packets are cut by count alone; a full packet of a cloud sparse enough
      // for wide cells spans about six edges)
      npk = (pts + kWarp - 1) / kWarp;
      if (dm.xwide == 0) npk = max(npk, (len + kSpanCap - 1) / kSpanCap);
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
  if (n_packets < 0) n_packets = packet_base[n_cells];  // slab build: the host launches before it knows the count
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

// ---- slab build (multi-GPU, one cloud): a rank indexes only the rows it reads ---------------------------------
// Every rank holds the whole cloud, but keys, sorts and tabulates only its slab: the rows whose queries it answers
// (own), one layer of rows around them whose normals it recomputes instead of receiving them (halo), and one more
// layer, the halo's candidates (window).  The cuts between the ranks must be the same on every rank and balance the
// ranks' work, so they are taken from a statistic every rank computes identically and cheaply: the cell histogram of
// a fixed 1-in-S sample of the points, turned into a cost per row (candidates tested ~ points of a cell times points
// of the 27 cells around it).  Whatever the sample says, the cuts are consistent and complete; balance only needs it
// to be representative.  Inside the window the sort, cell table, segments and packets are the unsharded ones (same
// keys, stable selection, stable sort), so a rank's results equal the single-GPU results bit for bit.

// cell histogram of every S-th run of 256 consecutive points (3 KB: whole DRAM bursts, unlike a stride of single points)
__global__ void __launch_bounds__(256) sample_hist_kernel(const float* __restrict__ xyz, int stride, int n, int sample,
                                                          const Domain* __restrict__ domains, float inv_cell,
                                                          int* __restrict__ cellcnt) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long first = ((t >> 6) * sample * 64 + (t & 63)) * 4;
  if (first >= n) return;
  const Domain dm = domains[0];
  const int last = (int)min((long long)n, first + 4);
  for (int i = (int)first; i < last; ++i) {
    const float* p = xyz + (size_t)i * stride;
    const float x = p[0], y = p[1], z = p[2];
    if (!finite3(x, y, z)) continue;
    int cy, cz;
    row_cells(dm, y, z, inv_cell, cy, cz);
    const int cx = xfine_coord(x, dm.ox, inv_cell, dm.nx, dm.xshift, dm.xwide) >> dm.xshift;
    atomicAdd(cellcnt + ((long long)cz * dm.ny + cy) * dm.nx + cx, 1);  // no return value: a fire-and-forget reduction
  }
}

// cost of a row = sum over its occupied cells of (sampled points) x (estimated candidates per point + constants).  One
// thread per cell of the table, streamed (L2 hits right after the histogram kernel); only the occupied cells -- a few per
// cent of the table of a cloud of surfaces, in runs along x -- read the 3 x 3 x 3 cells around them, and the lanes of a
// run read neighbouring words.  A warp's cells lie in one or two rows: one 64-bit atomic per row and warp.  Integer sums:
// the order does not matter.
__global__ void __launch_bounds__(256) cell_cost_kernel(const Domain* __restrict__ domains, const int* __restrict__ cnt,
                                                        long long n_cells, int sample, unsigned long long* __restrict__ rowcost) {
  const Domain dm = domains[0];
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; base < n_cells; base += stride) {
    const long long cell = base + lane;
    const int c = cell < n_cells ? cnt[cell] : 0;
    const int row = (int)(min(cell, n_cells - 1) / dm.nx);
    unsigned long long cost = 0;
    if (c != 0) {
      const int cx = (int)(cell - (long long)row * dm.nx), cy = row % dm.ny, cz = row / dm.ny;
      // candidates ~ the 3 x 3 rows around the cell, along x the cell itself plus 0.7 of either neighbour: a packet's x
      // window is its own extent + 2 r, about 2.4 cells (a full 3-cell stencil overrates surfaces that run along x,
      // e.g. the walls an end slab consists of, by a quarter against surfaces across x)
      int s10 = 0;
#pragma unroll
      for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int y = cy + dy, z = cz + dz;
          if (y < 0 || y >= dm.ny || z < 0 || z >= dm.nz) continue;
          const int* r = cnt + ((long long)z * dm.ny + y) * dm.nx + cx;
          s10 += 10 * r[0] + 7 * ((cx > 0 ? r[-1] : 0) + (cx + 1 < dm.nx ? r[1] : 0));
        }
      // + a constant per point (fit, eigen-solve: 2 candidates' worth) and per occupied cell (sparse rows make many
      // short packets, each with its own run table and chunk overhead: 16 candidates' worth per cell)
      cost = (unsigned long long)((long long)c * ((long long)s10 * sample + 20) + 160);
    }
    if (__ballot_sync(kFull, c != 0) == 0) continue;
    // the warp's cells are consecutive: lanes of the first row, lanes of the last row, (nx < 16: rows in between)
    const int row_a = __shfl_sync(kFull, row, 0), row_b = __shfl_sync(kFull, row, 31);
    unsigned long long sa = row == row_a ? cost : 0ull, sb = (row == row_b && row_b != row_a) ? cost : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sa += __shfl_xor_sync(kFull, sa, o);
      sb += __shfl_xor_sync(kFull, sb, o);
    }
    if (lane == 0 && sa) atomicAdd(rowcost + row_a, sa);
    if (lane == 31 && sb) atomicAdd(rowcost + row_b, sb);
    if (row != row_a && row != row_b && cost) atomicAdd(rowcost + row, cost);
  }
}

__device__ __forceinline__ long long cum_before(const long long* __restrict__ cum, int p) { return p > 0 ? cum[p - 1] : 0; }

// first index whose inclusive cost prefix reaches t, found by one warp: 32 probes per step instead of a chain of
// dependent loads (the cuts are on the critical path of every sharded build)
__device__ __forceinline__ int cut_at(const long long* __restrict__ cum, int n_items, long long t, int lane) {
  int lo = 0, hi = n_items;  // the answer lies in [lo, hi]
  while (hi - lo > 32) {
    const int step = (hi - lo + 31) / 32;
    const long long idx = (long long)lo + (long long)lane * step;
    const bool reached = idx >= hi || cum[idx] >= t;
    const unsigned m = __ballot_sync(kFull, reached);
    const int first = m ? __ffs(m) - 1 : 32;
    if (first == 0) return lo;
    const long long new_lo = (long long)lo + (long long)(first - 1) * step + 1;
    if (first < 32) hi = (int)min((long long)hi, (long long)lo + (long long)first * step);
    lo = (int)new_lo;
  }
  const int idx = lo + lane;
  const bool reached = idx >= hi || cum[idx] >= t;
  const unsigned m = __ballot_sync(kFull, reached);
  return m ? min(lo + __ffs(m) - 1, hi) : hi;
}

// Row cuts.  A rank's work is its own rows (both passes) plus the normals of its halo rows (one layer = ny + 1 rows on
// either side; the first and last rank have one halo).  One block, one warp per cut; a few fixed-point rounds move the
// cuts until own cost + halo_permille/1000 * halo cost is the same for every rank.  Leaves the rank's row ranges in
// `info` and turns domains[0] into the slab's table geometry.
struct SplitShares {
  double cum[65];  // cumulative share of the total cost the ranks before rank c get; cum[0] = 0, cum[world] = 1
};
__global__ void __launch_bounds__(1024) slab_split_kernel(const long long* __restrict__ cum, int world, int rank,
                                                          int halo_permille, int want_exchange, const SplitShares sh,
                                                          Domain* __restrict__ domains, int* __restrict__ cuts,
                                                          SlabInfo* __restrict__ info) {
  __shared__ int s[65];
  __shared__ long long halo[64];
  __shared__ long long target[65];
  const Domain dm = domains[0];
  const int n_rows = dm.ny * dm.nz, layer = dm.ny + 1;
  const int g = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const long long total = cum[n_rows - 1];
  if (g == 0) {
    s[0] = 0;
    s[world] = n_rows;
  }
  for (int c = 1 + warp; c < world; c += n_warps) {  // one warp per cut
    const int v = cut_at(cum, n_rows, (long long)((double)total * sh.cum[c]), lane);
    if (lane == 0) s[c] = v;
  }
  __syncthreads();
  // exchanged halos cost (almost) nothing: the first cuts stand
  for (int round = 0; round < (want_exchange ? 0 : 4); ++round) {
    if (g < world) {
      long long h = 0;
      const int a = s[g], b = s[g + 1];
      if (g > 0) h += cum_before(cum, a) - cum_before(cum, max(a - layer, 0));
      if (g < world - 1) h += cum_before(cum, min(b + layer, n_rows)) - cum_before(cum, b);
      halo[g] = h * halo_permille / 1000;
    }
    __syncthreads();
    if (g == 0) {
      long long sum = total;
      for (int j = 0; j < world; ++j) sum += halo[j];
      const long long floor_own = total / (8 * (long long)world) + 1;
      double acc = 0;
      target[0] = 0;
      for (int j = 0; j < world; ++j) {
        const long long each = (long long)((double)sum * (sh.cum[j + 1] - sh.cum[j]));
        const long long own = each - halo[j] > floor_own ? each - halo[j] : floor_own;
        acc += (double)own;
        target[j + 1] = (long long)acc;
      }
      const double scale = (double)total / acc;  // the clamp may have moved the sum
      for (int j = 1; j < world; ++j) target[j] = (long long)((double)target[j] * scale);
    }
    __syncthreads();
    for (int c = 1 + warp; c < world; c += n_warps) {
      const int v = cut_at(cum, n_rows, target[c], lane);
      if (lane == 0) s[c] = v;
    }
    __syncthreads();
  }
  if (g == 0)
    for (int c = 1; c <= world; ++c) s[c] = max(s[c], s[c - 1]);  // monotone whatever the rounds did
  __syncthreads();
  if (g <= world) cuts[g] = s[g];
  if (g == 0) {
    SlabInfo si{};
    si.own_lo = s[rank];
    si.own_hi = s[rank + 1];
    // the exchange needs every rank to be two layers thick: a halo then comes from the immediate neighbour alone and
    // there are packets to run while it travels
    bool thick = want_exchange != 0;
    for (int c = 0; c < world; ++c) thick = thick && s[c + 1] - s[c] >= 2 * layer;
    si.exchange = thick ? 1 : 0;
    if (thick) {
      si.halo_lo = si.own_lo;
      si.halo_hi = si.own_hi;
      si.win_lo = max(si.own_lo - layer, 0);
      si.win_hi = min(si.own_hi + layer, n_rows);
    } else if (si.own_hi > si.own_lo) {
      si.halo_lo = max(si.own_lo - layer, 0);
      si.halo_hi = min(si.own_hi + layer, n_rows);
      si.win_lo = max(si.own_lo - 2 * layer, 0);
      si.win_hi = min(si.own_hi + 2 * layer, n_rows);
    } else {
      si.halo_lo = si.halo_hi = si.win_lo = si.win_hi = si.own_lo;
    }
    *info = si;
    domains[0].row_lo = si.win_lo;
    domains[0].row_hi = si.win_hi;
    domains[0].cell_base = -(long long)si.win_lo * dm.nx;
  }
}

// Stable selection of the window's points, read once from the caller's cloud and kept as {x, y, z, input index}.
// Two small kernels instead of a library select (whose look-back pass over 20 M twelve-byte items took 220 us): each
// block compacts its own 2048 consecutive points, in order, into its own stretch of the staging array and leaves its
// count; after a scan of the ~10 k counts the key kernel walks the blocks' stretches and numbers the points
// consecutively -- input order is preserved (the radix sort is stable, so equal keys keep the unsharded order).
struct SelPoint {
  float x, y, z;
  int idx;
};
constexpr int kSelChunk = 2048;  // points per selection block (256 threads x 8 consecutive points)

__device__ __forceinline__ bool in_window(const Domain& dm, float inv_cell, float x, float y, float z) {
  if (!finite3(x, y, z)) return false;
  int cy, cz;
  row_cells(dm, y, z, inv_cell, cy, cz);
  const int row = cz * dm.ny + cy;
  return row >= dm.row_lo && row < dm.row_hi;
}

template <bool kVec>
__global__ void __launch_bounds__(256) slab_select_kernel(const float* __restrict__ xyz, int stride, int n,
                                                          const Domain* __restrict__ domains, float inv_cell,
                                                          SelPoint* __restrict__ stage, int* __restrict__ counts,
                                                          const RangeDefaults rd) {
  __shared__ int warp_sum[8];
  const Domain dm = domains[0];
  const int base = blockIdx.x * kSelChunk + threadIdx.x * 8;  // this thread's 8 consecutive points
  float px[8], py[8], pz[8];
  unsigned keep = 0;
  if (kVec && base + 8 <= n) {
    const float4* v = reinterpret_cast<const float4*>(xyz);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const Pts4 p = load_pts4(v, (base >> 2) + h);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        px[4 * h + k] = p.x[k];
        py[4 * h + k] = p.y[k];
        pz[4 * h + k] = p.z[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) keep |= in_window(dm, inv_cell, px[k], py[k], pz[k]) ? 1u << k : 0u;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (base + k < n) {
        const float* p = xyz + (size_t)(base + k) * stride;
        px[k] = p[0];
        py[k] = p[1];
        pz[k] = p[2];
        keep |= in_window(dm, inv_cell, px[k], py[k], pz[k]) ? 1u << k : 0u;
      }
    }
  }
  if (rd.nrm && base < rd.hi && base + 8 > rd.lo) {
    // input-range layout of a group: this rank owns the results of input indices [lo, hi); the points among them that are
    // nobody's query get the single-GPU path's values here (nobody else ever writes those entries)
    const float nan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = base + k;
      if (j < n && j >= rd.lo && j < rd.hi && !finite3(px[k], py[k], pz[k])) {
        rd.nrm[j - rd.lo] = make_float4(nan, nan, nan, nan);
        rd.rsd[j - rd.lo] = make_float2(rd.radius, rd.radius);
      }
    }
  }
  const int mine = __popc(keep), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  int before = incl - mine;
  for (int w = 0; w < warp; ++w) before += warp_sum[w];
  SelPoint* out = stage + (size_t)blockIdx.x * kSelChunk + before;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (keep & (1u << k)) *out++ = SelPoint{px[k], py[k], pz[k], base + k};
  if (threadIdx.x == 255) counts[blockIdx.x] = before + mine;
}

// keys + the window's cell histogram; one block per selection block, numbering the points consecutively
template <typename KeyT>
__global__ void __launch_bounds__(256) slab_key_kernel(const SelPoint* __restrict__ stage, const int* __restrict__ sel_off,
                                                       const Domain* __restrict__ domains, float inv_cell, int xbits,
                                                       KeyT* __restrict__ keys, int* __restrict__ vals,
                                                       int* __restrict__ cellcnt, SelPoint* __restrict__ compact) {
  const int first = sel_off[blockIdx.x], count = sel_off[blockIdx.x + 1] - first;
  const Domain dm = domains[0];
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const int src = blockIdx.x * kSelChunk + i;
    const SelPoint p = stage[src];
    int cy, cz;
    row_cells(dm, p.y, p.z, inv_cell, cy, cz);
    const int xf = xfine_coord(p.x, dm.ox, inv_cell, dm.nx, dm.xshift, dm.xwide);
    const long long lrow = (long long)cz * dm.ny + cy - dm.row_lo;
    atomicAdd(cellcnt + lrow * dm.nx + (xf >> dm.xshift), 1);
    keys[first + i] = (KeyT)(((unsigned long long)lrow << xbits) | (unsigned)xf);
    vals[first + i] = first + i;
    compact[first + i] = p;  // the staging array is sparse (one stretch per block): the gather after the sort reads this copy
  }
}

// the number of selected points joins the rank's SlabInfo
__global__ void slab_count_kernel(const int* __restrict__ sel_off, int n_blocks, SlabInfo* __restrict__ info) {
  if (threadIdx.x == 0 && blockIdx.x == 0) info->n_selected = sel_off[n_blocks];
}

// packet and query ranges of the rank's own and halo rows (packet_base / cell_start at a row's first cell = packets /
// points of the rows before it)
__global__ void slab_ranges_kernel(const Domain* __restrict__ domains, const int* __restrict__ cell_start,
                                   const int* __restrict__ packet_base, SlabInfo* __restrict__ info) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Domain dm = domains[0];
  SlabInfo si = *info;
  auto first_cell = [&](int row) { return (long long)(row - si.win_lo) * dm.nx; };
  si.n_packets = packet_base[first_cell(si.win_hi)];
  si.p0 = packet_base[first_cell(si.own_lo)];
  si.p1 = packet_base[first_cell(si.own_hi)];
  si.ph0 = packet_base[first_cell(si.halo_lo)];
  si.ph1 = packet_base[first_cell(si.halo_hi)];
  si.q0 = cell_start[first_cell(si.own_lo)];
  si.q1 = cell_start[first_cell(si.own_hi)];
  const int layer = dm.ny + 1;
  const int lo_end = min(si.own_lo + layer, si.own_hi), hi_begin = max(si.own_hi - layer, lo_end);
  si.r_blo[0] = si.p0;
  si.r_blo[1] = si.r_int[0] = packet_base[first_cell(lo_end)];
  si.r_int[1] = si.r_bhi[0] = packet_base[first_cell(hi_begin)];
  si.r_bhi[1] = si.p1;
  si.top_src = cell_start[first_cell(max(si.own_hi - layer, si.own_lo))];
  *info = si;
}

// sorted position -> position, input index
__global__ void __launch_bounds__(256) slab_place_kernel(const SelPoint* __restrict__ sel, const int* __restrict__ svals, int m,
                                                         float4* __restrict__ pos, int* __restrict__ perm) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  const SelPoint p = sel[svals[s]];
  pos[s] = make_float4(p.x, p.y, p.z, 0.f);
  perm[s] = p.idx;
}

}  // namespace

// one domain of packed, 16-byte aligned xyz: the layout the vector kernels read
static bool vector_layout(const cab_ctx* ctx) {
  return ctx->n_domains == 1 && ctx->stride == 3 && (reinterpret_cast<uintptr_t>(ctx->xyz_in) & 15) == 0;
}

// Per-domain bounding boxes and finite-point counts of the uploaded cloud -> ctx->dom_bounds / dom_count.
int compute_bounds(cab_ctx* ctx) {
  const int nd = ctx->n_domains;
  cudaStream_t st = ctx->stream;
  if (vector_layout(ctx) && ctx->n > 0) {  // one domain, packed xyz: no chunk list, one round trip
    if (int rc = reserve(ctx, ctx->b_bounds, 8 * sizeof(unsigned))) return rc;
    if (int rc = reserve_pinned(ctx, 8 * sizeof(unsigned) + 64)) return rc;
    init_bounds_kernel<<<1, 32, 0, st>>>((unsigned*)ctx->b_bounds.p);
    CAB_LAUNCH_CHECK(ctx);
    const int groups = (int)(ctx->n >> 2);
    const int blocks = std::max(1, std::min((groups + 255) / 256, ctx->sm_count * 8));
    bounds_vec_kernel<<<blocks, 256, 0, st>>>(ctx->xyz_in, (int)ctx->n, (unsigned*)ctx->b_bounds.p);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, ctx->b_bounds.p, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    const unsigned* hb = (const unsigned*)ctx->h_pin;
    ctx->dom_bounds.assign(6, 0.f);
    ctx->dom_count.assign(1, hb[6]);
    if (hb[6])
      for (int a = 0; a < 6; ++a) ctx->dom_bounds[a] = ord2f(hb[a]);
    return CAB_OK;
  }
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
  ctx->have_grid = ctx->have_normals = ctx->have_rsd = ctx->kcount_valid = ctx->trunc_hist_valid = false;
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
  // A cloud much sparser than the table at the requested edge (few neighbours per query: tens of cells per point) gets
  // cells that are 2^xwide edges WIDE along x before it gets coarser ones: a row's candidates are a range of its sorted
  // points found by binary search over the fine x coordinate, so wide cells cost two or three more probes per search
  // and no candidates, while the passes over the table (memset, scan, segments) shrink with it.
  constexpr int kMaxWide = 4;
  int xwide = 0;
  for (int attempt = 0;; ++attempt) {
    rows = cells = 0;
    double want = 0;  // cells this edge would need at xwide = 0, as a double (it may overflow int64)
    for (int d = 0; d < nd; ++d) {
      if (ctx->dom_count[d] == 0) continue;
      double nn[3];
      for (int a = 0; a < 3; ++a)
        nn[a] = std::floor(((double)ctx->dom_bounds[6 * (size_t)d + 3 + a] - (double)ctx->dom_bounds[6 * (size_t)d + a]) / cell_eff) + 2;
      want += nn[0] * nn[1] * nn[2];
    }
    xwide = 0;
    while (xwide < kMaxWide && std::ldexp(want, -xwide) > std::max(4.0 * (double)n_valid, 1048576.0)) ++xwide;
    bool too_big = false;
    double have = 0;
    for (int d = 0; d < nd && !too_big; ++d) {
      Domain& dm = ctx->domains[d];
      dm.row_base = rows;
      dm.cell_base = cells;
      dm.xwide = xwide;
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
        nn[0] = std::ceil(std::ldexp(nn[0], -xwide));                          // wide cells along x
        have += nn[0] * nn[1] * nn[2];
        if (nn[0] > (double)(1 << kXBits) || nn[1] * nn[2] > std::ldexp(1.0, 40) || have > (double)budget) {
          too_big = true;
          break;
        }
        dm.nx = (int)nn[0];
        dm.ny = (int)nn[1];
        dm.nz = (int)nn[2];
        // the slowest-varying slot gets the longer of the y and z extents: shards are slabs along it, and
        // the more layers a slab has the smaller the share of its halo
        dm.swap_yz = nn[1] > nn[2] ? 1 : 0;
        if (dm.swap_yz) {
          std::swap(dm.oy, dm.oz);
          std::swap(dm.ny, dm.nz);
        }
      }
      rows += (int64_t)dm.ny * dm.nz;
      cells += (int64_t)dm.ny * dm.nz * dm.nx;
    }
    if (!too_big) break;
    if (attempt >= 60 || !std::isfinite(cell_eff))
      return fail(ctx, CAB_ERR_OOM, "cab_build_grid: no cell size fits the dense cell table (budget %lld cells)", (long long)budget);
    cell_eff *= std::max(1.26, std::min(8.0, std::cbrt(std::ldexp(want, -xwide) / (double)budget) * 1.02));
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
  const bool key32 = row_bits + nx_bits + xwide + 2 <= 32;  // (the sub-cell bits count from the cell EDGE, not the wide cell)
  const int xbits = key32 ? std::min(kXBits, 32 - row_bits) : kXBits;
  for (int d = 0; d < nd; ++d) {
    Domain& dm = ctx->domains[d];
    dm.xshift = dm.xwide;  // x_fine >> xshift = cell x; the fine steps are 2^-(xshift - xwide) of an edge
    while (dm.xshift - dm.xwide < 8 && ((int64_t)dm.nx << (dm.xshift + 1)) <= ((int64_t)1 << xbits)) dm.xshift++;
    dm.row_lo = 0;
    dm.row_hi = dm.ny * dm.nz;
  }

  if (int rc = reserve(ctx, ctx->b_domains, nd * sizeof(Domain))) return rc;
  if (int rc = reserve_pinned(ctx, nd * sizeof(Domain) + sizeof(SlabInfo) + 64)) return rc;
  std::memcpy(ctx->h_pin, ctx->domains.data(), nd * sizeof(Domain));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->b_domains.p, ctx->h_pin, nd * sizeof(Domain), cudaMemcpyHostToDevice, st));

  ctx->slab = false;
  ctx->slab_info_valid = false;
  if (ctx->shard_world > 1 && nd > 1)
    return fail(ctx, CAB_ERR_STATE, "cab_build_grid: query shards apply to one cloud; a batch of clusters shards by cluster "
                                     "(one context per rank, each with its own clusters)");
  if (ctx->shard_world > 1 && n_valid > 0) return build_slab(ctx, key32, xbits);

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
  int* ident = (int*)ctx->b_vals[0].p;
  if (n > 0 && vector_layout(ctx)) {
    const unsigned blocks = (unsigned)(((n >> 2) + 1 + 255) / 256);
    if (key32)
      key_vec_kernel<unsigned><<<blocks, 256, 0, st>>>(ctx->xyz_in, n, (const Domain*)ctx->b_domains.p, ctx->inv_cell,
                                                       (unsigned long long)rows, xbits, (unsigned*)ctx->b_keys[0].p, ident,
                                                       (int*)ctx->b_cellcnt.p);
    else
      key_vec_kernel<unsigned long long><<<blocks, 256, 0, st>>>(ctx->xyz_in, n, (const Domain*)ctx->b_domains.p, ctx->inv_cell,
                                                                 (unsigned long long)rows, xbits,
                                                                 (unsigned long long*)ctx->b_keys[0].p, ident, (int*)ctx->b_cellcnt.p);
    CAB_LAUNCH_CHECK(ctx);
  } else if (n > 0) {
    if (key32)
      key_kernel<unsigned><<<(n + 255) / 256, 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, (const int*)ctx->b_domoff.p, nd,
                                                            (const Domain*)ctx->b_domains.p, ctx->inv_cell,
                                                            (unsigned long long)rows, xbits, (unsigned*)ctx->b_keys[0].p,
                                                            ident, (int*)ctx->b_cellcnt.p);
    else
      key_kernel<unsigned long long><<<(n + 255) / 256, 256, 0, st>>>(
          ctx->xyz_in, ctx->stride, n, (const int*)ctx->b_domoff.p, nd, (const Domain*)ctx->b_domains.p, ctx->inv_cell,
          (unsigned long long)rows, xbits, (unsigned long long*)ctx->b_keys[0].p, ident, (int*)ctx->b_cellcnt.p);
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

  // ---- radix sort by (row, fine x) ----------------------------------------------------
  if (n > 0) {
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
  }
  ctx->n_sorted = n;
  ctx->tm.shard_mode = 0;
  ctx->tm.n_sorted = ctx->n_valid;
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

// The slab build of one rank (see the kernels above).  Two host round trips: the bounds (before this function) and the
// number of points selected for the sort; everything after the second one only enqueues work.  With ctx->defer_sync the
// function returns with the stream still running (cab_step_*): ctx->slab_info and the timings are then filled in by
// finish_slab() after the step's final synchronisation.
int build_slab(cab_ctx* ctx, bool key32, int xbits) {
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  const Domain dm = ctx->domains[0];
  const int64_t rows = ctx->n_rows, cells = ctx->n_cells;
  const int w = ctx->shard_world;
  if (w + 1 > 64) return fail(ctx, CAB_ERR_ARG, "cab_set_shard: world > 63 not supported");
  if (rows >= (int64_t)1 << 31) return fail(ctx, CAB_ERR_OOM, "cab_build_grid: too many rows for a sharded build");
  // ---- cuts from the sampled cell histogram ------------------------------------------------------------
  const int sample = n >= (1 << 22) ? 8 : n >= (1 << 20) ? 4 : n >= (1 << 18) ? 2 : 1;
  const size_t split_ints = (size_t)((w + 1 + 3) & ~3);
  if (int rc = reserve(ctx, ctx->b_cellcnt, ((size_t)cells + 1) * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_pcost, (size_t)rows * 16 + 64)) return rc;
  if (int rc = reserve(ctx, ctx->b_slab, split_ints * 4 + sizeof(SlabInfo) + 64)) return rc;
  if (int rc = reserve(ctx, ctx->b_sel, ((size_t)n + kSelChunk) * sizeof(SelPoint) + 64)) return rc;
  long long* rowcost = (long long*)ctx->b_pcost.p;
  long long* cum = rowcost + rows;
  int* cuts = (int*)ctx->b_slab.p;
  SlabInfo* info = (SlabInfo*)(cuts + split_ints);
  Domain* d_dom = (Domain*)ctx->b_domains.p;
  size_t tmp_cost = 0, tmp_selscan = 0;
  const int sel_blocks = (n + kSelChunk - 1) / kSelChunk;
  cub::DeviceScan::InclusiveSum(nullptr, tmp_cost, (const long long*)nullptr, (long long*)nullptr, (int)rows, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_selscan, (const int*)nullptr, (int*)nullptr, sel_blocks + 1, st);
  if (int rc = reserve(ctx, ctx->b_cubtmp, std::max(tmp_cost, tmp_selscan) + 16)) return rc;
  if (int rc = reserve(ctx, ctx->b_vals[2], ((size_t)sel_blocks + 2) * 8)) return rc;
  int* sel_cnt = (int*)ctx->b_vals[2].p;
  int* sel_off = sel_cnt + sel_blocks + 1;
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_cellcnt.p, 0, ((size_t)cells + 1) * 4, st));
  CAB_CUDA(ctx, cudaMemsetAsync(rowcost, 0, (size_t)rows * 8, st));
  {
    const long long threads = (((long long)n + 256LL * sample - 1) / (256LL * sample)) * 64;
    sample_hist_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, sample, d_dom,
                                                                        ctx->inv_cell, (int*)ctx->b_cellcnt.p);
    CAB_LAUNCH_CHECK(ctx);
    cell_cost_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_dom, (const int*)ctx->b_cellcnt.p, (long long)cells, sample,
                                                        (unsigned long long*)rowcost);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cub::DeviceScan::InclusiveSum(ctx->b_cubtmp.p, tmp_cost, rowcost, cum, (int)rows, st));
    SplitShares sh{};
    for (int c = 0; c <= w; ++c)
      sh.cum[c] = (int)ctx->shard_cum.size() == w + 1 ? ctx->shard_cum[(size_t)c] : (double)c / (double)w;
    slab_split_kernel<<<1, 32 * std::min(32, std::max(2, w)), 0, st>>>(cum, w, ctx->shard_rank, ctx->halo_permille,
                                                                       ctx->want_halo_exchange ? 1 : 0, sh, d_dom, cuts, info);
    CAB_LAUNCH_CHECK(ctx);
    // stable selection of the window's points: per-block compaction, then a scan of the blocks' counts
    CAB_CUDA(ctx, cudaMemsetAsync(sel_cnt + sel_blocks, 0, 4, st));
    if (vector_layout(ctx))
      slab_select_kernel<true><<<sel_blocks, 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, d_dom, ctx->inv_cell, (SelPoint*)ctx->b_sel.p, sel_cnt, ctx->range_defaults);
    else
      slab_select_kernel<false><<<sel_blocks, 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, d_dom, ctx->inv_cell, (SelPoint*)ctx->b_sel.p, sel_cnt, ctx->range_defaults);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp_selscan, sel_cnt, sel_off, sel_blocks + 1, st));
    slab_count_kernel<<<1, 32, 0, st>>>(sel_off, sel_blocks, info);
    CAB_LAUNCH_CHECK(ctx);
    ctx->tm.kernel_launches += 4;
  }
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, info, sizeof(SlabInfo), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  SlabInfo si;
  std::memcpy(&si, ctx->h_pin, sizeof(SlabInfo));
  ctx->slab_info = si;  // rows and mode are final; the packet / query ranges follow with finish_slab()
  const int m = si.n_selected;
  const int64_t lrows = (int64_t)si.win_hi - si.win_lo;
  const int64_t lcells = lrows * dm.nx;
  const size_t ncell1 = (size_t)lcells + 1;
  // host copy of the table geometry the split kernel left on the device
  ctx->domains[0].row_lo = si.win_lo;
  ctx->domains[0].row_hi = si.win_hi;
  ctx->domains[0].cell_base = -(int64_t)si.win_lo * dm.nx;
  // ---- keys + the window's cell table, segments, packets ------------------------------------------------------
  int lrow_bits = 1;
  while (((int64_t)1 << lrow_bits) < std::max<int64_t>(lrows, 1)) ++lrow_bits;
  const int end_bit = std::min(key32 ? 32 : 64, xbits + lrow_bits);
  size_t tmp_sort = 0, tmp_scan = 0;
  if (key32)
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, m, 0, end_bit, st);
  else
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, m, 0, end_bit, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (const int*)nullptr, (int*)nullptr, (int)ncell1, st);
  size_t tmp_bytes = tmp_scan;
  const size_t mm = (size_t)std::max(m, 1);
  const size_t max_packets = std::min<size_t>(mm, mm / kWarp + (size_t)lcells + 1);
  if (int rc = reserve(ctx, ctx->b_cubtmp, tmp_bytes + 16)) return rc;
  if (int rc = reserve(ctx, ctx->b_sorttmp, tmp_sort + 16)) return rc;  // the sort runs beside the scans: its own scratch
  if (int rc = reserve(ctx, ctx->b_keys[0], mm * 8)) return rc;
  if (int rc = reserve(ctx, ctx->b_keys[1], mm * 8)) return rc;
  if (int rc = reserve(ctx, ctx->b_keys[2], mm * sizeof(SelPoint))) return rc;
  if (int rc = reserve(ctx, ctx->b_vals[0], mm * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_vals[1], mm * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_perm, mm * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_pos, mm * sizeof(float4))) return rc;
  if (int rc = reserve(ctx, ctx->b_cellstart, ncell1 * 4)) return rc;
  if (int rc = reserve(ctx, ctx->b_rowpk, ncell1 * 4 * 3)) return rc;
  if (int rc = reserve(ctx, ctx->b_packets, max_packets * sizeof(Packet))) return rc;
  int* cellcnt = (int*)ctx->b_cellcnt.p;  // the sample histogram is no longer needed
  int* segpk = (int*)ctx->b_rowpk.p;
  int* packet_base = segpk + ncell1;
  int* seglen = packet_base + ncell1;
  CAB_CUDA(ctx, cudaMemsetAsync(cellcnt, 0, ncell1 * 4, st));
  if (m > 0) {
    if (key32)
      slab_key_kernel<unsigned><<<sel_blocks, 256, 0, st>>>((const SelPoint*)ctx->b_sel.p, sel_off, d_dom, ctx->inv_cell, xbits,
                                                           (unsigned*)ctx->b_keys[0].p, (int*)ctx->b_vals[0].p, cellcnt,
                                                           (SelPoint*)ctx->b_keys[2].p);
    else
      slab_key_kernel<unsigned long long><<<sel_blocks, 256, 0, st>>>((const SelPoint*)ctx->b_sel.p, sel_off, d_dom, ctx->inv_cell,
                                                                     xbits, (unsigned long long*)ctx->b_keys[0].p,
                                                                     (int*)ctx->b_vals[0].p, cellcnt, (SelPoint*)ctx->b_keys[2].p);
    CAB_LAUNCH_CHECK(ctx);
  }
  // ---- radix sort of the window by (row, fine x), on the copy stream: the four latency-bound passes over a few million
  // keys run beside the cell table, segment and packet kernels below, which need the cell histogram only
  static const bool serial_sort = std::getenv("CAB_SERIAL_SORT") != nullptr;  // A/B switch
  cudaStream_t ss = serial_sort ? st : ctx->copy_stream;
  if (m > 0) {
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    CAB_CUDA(ctx, cudaStreamWaitEvent(ss, ctx->ev_fork, 0));
    if (key32)
      CAB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_sorttmp.p, tmp_sort, (const unsigned*)ctx->b_keys[0].p,
                                                    (unsigned*)ctx->b_keys[1].p, (const int*)ctx->b_vals[0].p,
                                                    (int*)ctx->b_vals[1].p, m, 0, end_bit, ss));
    else
      CAB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_sorttmp.p, tmp_sort, (const unsigned long long*)ctx->b_keys[0].p,
                                                    (unsigned long long*)ctx->b_keys[1].p, (const int*)ctx->b_vals[0].p,
                                                    (int*)ctx->b_vals[1].p, m, 0, end_bit, ss));
    ctx->tm.kernel_launches += 1 + (end_bit + 7) / 8;
    slab_place_kernel<<<(m + 255) / 256, 256, 0, ss>>>((const SelPoint*)ctx->b_keys[2].p, (const int*)ctx->b_vals[1].p, m,
                                                      (float4*)ctx->b_pos.p, (int*)ctx->b_perm.p);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ss));
  }
  CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp_bytes, (const int*)cellcnt, (int*)ctx->b_cellstart.p,
                                              (int)ncell1, st));
  segment_kernel<<<(unsigned)((ncell1 + 255) / 256), 256, 0, st>>>(d_dom, 1, lcells, (const int*)ctx->b_cellstart.p, segpk, seglen);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp_bytes, segpk, packet_base, (int)ncell1, st));
  ctx->tm.kernel_launches += 4;
  fill_packets_kernel<<<(unsigned)((max_packets + 255) / 256), 256, 0, st>>>(d_dom, 1, lcells, (const int*)ctx->b_cellstart.p,
                                                                            packet_base, seglen, -1, (Packet*)ctx->b_packets.p);
  CAB_LAUNCH_CHECK(ctx);
  slab_ranges_kernel<<<1, 32, 0, st>>>(d_dom, (const int*)ctx->b_cellstart.p, packet_base, info);
  CAB_LAUNCH_CHECK(ctx);
  if (m > 0) CAB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));  // the sorted window joins the tables
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_step + kStepSlab, info, sizeof(SlabInfo), cudaMemcpyDeviceToHost, st));
  ctx->slab = true;
  ctx->n_sorted = m;
  ctx->n_valid = m;  // finite points of the local sorted arrays
  ctx->n_cells = lcells;
  ctx->tm.n_sorted = m;
  ctx->tm.n_points = ctx->n;
  ctx->tm.n_valid = m;
  ctx->tm.n_rows = lrows;
  ctx->tm.n_cells = lcells;
  ctx->have_grid = true;
  if (ctx->defer_sync) return CAB_OK;
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  return finish_slab(ctx);
}

// After the stream has drained: the host copy of the slab's ranges and the build time.
int finish_slab(cab_ctx* ctx) {
  std::memcpy(&ctx->slab_info, ctx->h_step + kStepSlab, sizeof(SlabInfo));
  ctx->slab_info_valid = true;
  ctx->n_packets = ctx->slab_info.n_packets;
  ctx->tm.n_packets = ctx->n_packets;
  ctx->tm.shard_mode = ctx->slab_info.exchange ? 2 : 1;
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.build_ms, ctx->ev[0], ctx->ev[1]));
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
  if (ctx->slab && ctx->slab_info_valid) {
    *p0 = ctx->slab_info.p0;
    *p1 = ctx->slab_info.p1;
    return;
  }
  *p0 = 0;
  *p1 = ctx->n_packets;
}

// device-side {p0, p1} of the rank's own (halo = false) or own + halo packets; null when the grid is not a slab
const int* slab_packet_range(const cab_ctx* ctx, bool halo) {
  if (!ctx->slab) return nullptr;
  const size_t split_ints = (size_t)((ctx->shard_world + 1 + 3) & ~3);
  const SlabInfo* info = (const SlabInfo*)((const int*)ctx->b_slab.p + split_ints);
  return halo ? &info->ph0 : &info->p0;
}
const SlabInfo* slab_info_device(const cab_ctx* ctx) {
  if (!ctx->slab) return nullptr;
  const size_t split_ints = (size_t)((ctx->shard_world + 1 + 3) & ~3);
  return (const SlabInfo*)((const int*)ctx->b_slab.p + split_ints);
}

}  // namespace cab
