// cab_grsd.cu -- Global RSD for a batch of segmented clusters.
// Replaces, per cluster, getVoxelGrid + extractGRSDSignature21
// (color_chlac/include/color_chlac/grsd_colorCHLAC_tools.hpp:94-100 and :131-294):
//   pcl::VoxelGrid(leaf, saveLeafLayout)  -> voxel keys, centroids, dense leaf layout
//   pcl::RSDEstimation at the centroids over the full-resolution cloud + normals (:164-180)
//   get_type (:104-116), 26-neighbour transition counting (:230-260), 21-bin packing (:266-276).
// All clusters are processed together: they are independent "domains" of one search grid, so
// neighbourhoods never cross clusters.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include <cfloat>
#include <cmath>
#include <cstring>

#include "cab_internal.cuh"

namespace cab {

namespace {

struct VoxGrid {  // per cluster
  int min_b[3];
  int div_b[3];
  long long layout_base;  // into the concatenated leaf layouts
  int vox_first;          // first voxel (global index) of this cluster
  int pad;
};

__device__ __forceinline__ int find_dom(const int* __restrict__ domoff, int nd, int i) {
  int lo = 0, hi = nd;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (domoff[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// voxel key of every input point: (cluster << 32) | linear voxel index, x fastest.
// pcl::VoxelGrid::applyFilter [EXTERNAL]: ijk = floor(p * inverse_leaf) - min_b, fp32 multiply.
__global__ void __launch_bounds__(256) voxel_key_kernel(const float* __restrict__ xyz, int stride, int n,
                                                        const int* __restrict__ domoff, int nd,
                                                        const VoxGrid* __restrict__ vg, float inv_leaf,
                                                        unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + (size_t)i * stride;
  const float x = p[0], y = p[1], z = p[2];
  unsigned long long key = ~0ull;  // non-finite points sort last and are dropped
  if (isfinite(x) && isfinite(y) && isfinite(z)) {
    const int d = nd > 1 ? find_dom(domoff, nd, i) : 0;
    const VoxGrid g = vg[d];
    const int i0 = (int)floorf(__fmul_rn(x, inv_leaf)) - g.min_b[0];
    const int i1 = (int)floorf(__fmul_rn(y, inv_leaf)) - g.min_b[1];
    const int i2 = (int)floorf(__fmul_rn(z, inv_leaf)) - g.min_b[2];
    const unsigned lin = (unsigned)(i0 + i1 * g.div_b[0] + i2 * g.div_b[0] * g.div_b[1]);
    key = ((unsigned long long)d << 32) | lin;
  }
  keys[i] = key;
  vals[i] = i;
}

// One warp per voxel: the centroid as pcl::VoxelGrid of the reference's era computes it -- an Eigen::VectorXf summed in
// cloud order, then `/= count`, which Eigen 3.0-3.2 evaluates as a multiplication by 1.0f / count (the reference's
// shipped feature vectors pin both: DESIGN.md section 2) -- and the leaf layout.  The lanes load 32 points at a time,
// every lane then adds them in order (the float sum must be sequential).
__global__ void __launch_bounds__(256) centroid_kernel(const float* __restrict__ xyz, int stride,
                                                       const unsigned long long* __restrict__ ukeys,
                                                       const int* __restrict__ ucount, const int* __restrict__ ustart,
                                                       const int* __restrict__ sorted_idx, int nvox,
                                                       const VoxGrid* __restrict__ vg, float4* __restrict__ cent,
                                                       int* __restrict__ layout) {
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (v >= nvox) return;
  const int b = ustart[v], c = ucount[v];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int base = 0; base < c; base += kWarp) {
    float px = 0.f, py = 0.f, pz = 0.f;
    if (base + lane < c) {
      const float* p = xyz + (size_t)sorted_idx[b + base + lane] * stride;
      px = p[0];
      py = p[1];
      pz = p[2];
    }
    const int m = min(kWarp, c - base);
    for (int j = 0; j < m; ++j) {
      sx = __fadd_rn(sx, __shfl_sync(kFull, px, j));
      sy = __fadd_rn(sy, __shfl_sync(kFull, py, j));
      sz = __fadd_rn(sz, __shfl_sync(kFull, pz, j));
    }
  }
  if (lane == 0) {
    const float rcp = __fdiv_rn(1.0f, (float)c);
    cent[v] = make_float4(__fmul_rn(sx, rcp), __fmul_rn(sy, rcp), __fmul_rn(sz, rcp), 0.f);
    const unsigned long long key = ukeys[v];
    const int d = (int)(key >> 32);
    layout[vg[d].layout_base + (unsigned)(key & 0xffffffffu)] = v - vg[d].vox_first;
  }
}

struct VRsdArgs {
  GridView g;
  const float4* nrm;     // sorted order
  const float4* cent;    // voxel centroids
  const unsigned long long* ukeys;
  int nvox;
  float r2;
  int ndiv, flags;
  double radius, plane_radius;
  float skip_thr;        // smallest fp32 d2 whose sqrt((double)d2) > radius
  float bin_thr[9];      // ndiv + 1 entries
  float2* radii;         // out r_min, r_max
  int* labels;
  // one large cloud sharded over the ranks of a group (cab_grsd_cloud): a rank labels the voxels whose centroid lies in
  // one of its own rows (every centroid lies in exactly one rank's rows: the cuts are the same on every rank); `own`
  // remembers which, the other voxels get label -1 until the ranks' labels are merged
  const SlabInfo* slab;
  unsigned char* own;
};

// grsd_colorCHLAC_tools.hpp:104-116: fp32 radii compared against double literals
__device__ __forceinline__ int get_type(float min_radius, float max_radius) {
  if ((double)min_radius > 0.100) return 1;       // PLANE
  if ((double)max_radius > 0.175) return 2;       // CYLINDER
  if ((double)min_radius < 0.015) return 0;       // NOISE
  if ((double)(max_radius - min_radius) < 0.050) return 3;  // SPHERE
  return 4;                                        // EDGE
}

constexpr int kVDiv = 8;  // max nr_subdiv of the voxel RSD (PCL default 5)

// One warp per voxel centroid.  pcl::RSDEstimation semantics [EXTERNAL] (call site :164-180):
// neighbours = surface points within r of the centroid; reference element = the nearest one
// (ties: smallest input index); angles/distances are measured from it; pairs farther than r from
// it are skipped.
__global__ void __launch_bounds__(256) voxel_rsd_kernel(const VRsdArgs a) {
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (v >= a.nvox) return;
  const GridView& g = a.g;
  const float4 q = a.cent[v];
  const int d = (int)(a.ukeys[v] >> 32);
  const Domain dm = g.domains[d];
  int cy, cz;
  row_cells(dm, q.y, q.z, g.inv_cell, cy, cz);
  if (a.own) {
    bool mine = true;
    if (a.slab) {
      const int row = cz * dm.ny + cy;
      mine = row >= a.slab->own_lo && row < a.slab->own_hi;
    }
    if (lane == 0) a.own[v] = mine ? 1 : 0;
    if (!mine) {
      if (lane == 0) {
        a.radii[v] = make_float2(0.f, 0.f);
        a.labels[v] = -1;
      }
      return;
    }
  }
  const int cx = xfine_coord(q.x, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide) >> dm.xshift;
  const int cxlo = max(cx - 1, 0), cxhi = min(cx + 1, dm.nx - 1);
  int rb = 0, re = 0;
  if (lane < 9) {
    const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
    if (y >= 0 && y < dm.ny && z >= 0 && z < dm.nz) {
      const long long c = dm.cell_base + ((long long)z * dm.ny + y) * dm.nx;
      rb = g.cell_start[c + cxlo];
      re = g.cell_start[c + cxhi + 1];
    }
  }
  // pass 1: nearest surface point by (d2, input index)
  float bd2 = INFINITY;
  int bidx = INT_MAX, bpos = -1;
  for (int t = 0; t < 9; ++t) {
    const int b = __shfl_sync(kFull, rb, t), e = __shfl_sync(kFull, re, t);
    for (int j = b + lane; j < e; j += kWarp) {
      const float4 c = g.pos[j];
      const float d2 = d2_rule(c.x, c.y, c.z, q.x, q.y, q.z);
      if (d2 <= a.r2) {
        const int id = g.perm[j];
        if (d2 < bd2 || (d2 == bd2 && id < bidx)) {
          bd2 = d2;
          bidx = id;
          bpos = j;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float od2 = __shfl_xor_sync(kFull, bd2, o);
    const int oid = __shfl_xor_sync(kFull, bidx, o), op = __shfl_xor_sync(kFull, bpos, o);
    if (od2 < bd2 || (od2 == bd2 && oid < bidx)) {
      bd2 = od2;
      bidx = oid;
      bpos = op;
    }
  }
  // per-lane bins: .lo = cosine of smallest |value| (largest angle), .hi = largest |value|
  float lo[kVDiv], hi[kVDiv];
#pragma unroll
  for (int b = 0; b < kVDiv; ++b) {
    lo[b] = INFINITY;
    hi[b] = 0.f;
  }
  if ((a.flags & CAB_RSD_SEED_BIN0) && lane == 0) lo[0] = hi[0] = 1.f;
  if (bpos >= 0) {
    const float4 rp = g.pos[bpos];
    const float4 rn = a.nrm[bpos];
    for (int t = 0; t < 9; ++t) {
      const int b = __shfl_sync(kFull, rb, t), e = __shfl_sync(kFull, re, t);
      for (int j = b + lane; j < e; j += kWarp) {
        if (j == bpos) continue;
        const float4 c = g.pos[j];
        if (!(d2_rule(c.x, c.y, c.z, q.x, q.y, q.z) <= a.r2)) continue;
        const float dd = d2_rule(c.x, c.y, c.z, rp.x, rp.y, rp.z);
        if (dd >= a.skip_thr) continue;  // dist > max_dist
        const float4 nm = a.nrm[j];
        // pcl::computeRSD: normals[*i] * normals[*begin], fp32, left to right
        float cs = __fadd_rn(__fadd_rn(__fmul_rn(nm.x, rn.x), __fmul_rn(nm.y, rn.y)), __fmul_rn(nm.z, rn.z));
        if (cs > 1.f) cs = 1.f;
        if (cs < -1.f) cs = -1.f;
        int bin = 0;
#pragma unroll
        for (int b = 1; b < kVDiv; ++b) bin += (b < a.ndiv && dd >= a.bin_thr[b]) ? 1 : 0;
        const float ac = fabsf(cs);
#pragma unroll
        for (int b = 0; b < kVDiv; ++b)
          if (b == bin) {
            if (ac < fabsf(lo[b])) lo[b] = cs;
            if (ac >= fabsf(hi[b])) hi[b] = cs;
          }
      }
    }
  }
  // warp-reduce the bins (extremes of |cosine|)
#pragma unroll
  for (int b = 0; b < kVDiv; ++b) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ol = __shfl_xor_sync(kFull, lo[b], o), oh = __shfl_xor_sync(kFull, hi[b], o);
      if (fabsf(ol) < fabsf(lo[b])) lo[b] = ol;
      // keep the update rule "later >= wins" irrelevant: only |value| matters for the angle up to 1 ulp
      if (fabsf(oh) > fabsf(hi[b])) hi[b] = oh;
    }
  }
  if (lane == 0) {
    double Amint_Amin = 0, Amint_d = 0, Amaxt_Amax = 0, Amaxt_d = 0;
#pragma unroll
    for (int di = 0; di < kVDiv; ++di) {
      if (di < a.ndiv && fabsf(lo[di]) <= 1.f) {
        double p_min = acos((double)hi[di]), p_max = acos((double)lo[di]);
        if (p_min > M_PI / 2) p_min = M_PI - p_min;
        if (p_max > M_PI / 2) p_max = M_PI - p_max;
        const double f = (di + 0.5) * a.radius / a.ndiv;
        Amint_Amin = __dadd_rn(Amint_Amin, __dmul_rn(p_min, p_min));
        Amint_d = __dadd_rn(Amint_d, __dmul_rn(p_min, f));
        Amaxt_Amax = __dadd_rn(Amaxt_Amax, __dmul_rn(p_max, p_max));
        Amaxt_d = __dadd_rn(Amaxt_d, __dmul_rn(p_max, f));
      }
    }
    const double max_radius = (Amint_Amin == 0) ? a.plane_radius : fmin(Amint_d / Amint_Amin, a.plane_radius);
    const double min_radius = (Amaxt_Amax == 0) ? a.plane_radius : fmin(Amaxt_d / Amaxt_Amax, a.plane_radius);
    float rmin = (float)min_radius, rmax = (float)max_radius;
    if (a.flags & CAB_RSD_SCALE_SORT) {
      const float x = rmax * 1.1f, y = rmin * 0.9f;
      rmin = fminf(x, y);
      rmax = fmaxf(x, y);
    }
    a.radii[v] = make_float2(rmin, rmax);
    a.labels[v] = get_type(rmin, rmax);
  }
}

__constant__ int c_off26[26][3];

// One block per cluster: 6x6 transition matrix in shared memory (integer atomics), then the
// upper triangle is packed into 21 bins (grsd_colorCHLAC_tools.hpp:230-276, hist_num == 1).
__global__ void __launch_bounds__(128) transitions_kernel(const VoxGrid* __restrict__ vg, const int* __restrict__ vox_off,
                                                          const float4* __restrict__ cent, const int* __restrict__ labels,
                                                          const int* __restrict__ layout, float leaf,
                                                          int* __restrict__ hist21) {
  __shared__ int M[36];
  const int d = blockIdx.x;
  if (threadIdx.x < 36) M[threadIdx.x] = 0;
  __syncthreads();
  const VoxGrid g = vg[d];
  const int v0 = vox_off[d], v1 = vox_off[d + 1];
  const int work = (v1 - v0) * 26;
  for (int w = threadIdx.x; w < work; w += blockDim.x) {
    const int v = v0 + w / 26, o = w % 26;
    const float4 c = cent[v];
    // pcl::VoxelGrid::getNeighborCentroidIndices [EXTERNAL]: ijk = floor(ref / leaf) -- a division in the PCL of the
    // reference's era (pinned by its shipped cube / dice feature vectors, whose centroids sit on voxel faces)
    const int i0 = (int)floorf(__fdiv_rn(c.x, leaf)), i1 = (int)floorf(__fdiv_rn(c.y, leaf)),
              i2 = (int)floorf(__fdiv_rn(c.z, leaf));
    const int n0 = i0 + c_off26[o][0] - g.min_b[0], n1 = i1 + c_off26[o][1] - g.min_b[1],
              n2 = i2 + c_off26[o][2] - g.min_b[2];
    int nt = 5;  // EMPTY
    if (n0 >= 0 && n0 < g.div_b[0] && n1 >= 0 && n1 < g.div_b[1] && n2 >= 0 && n2 < g.div_b[2]) {
      const int nb = layout[g.layout_base + n0 + (long long)n1 * g.div_b[0] + (long long)n2 * g.div_b[0] * g.div_b[1]];
      if (nb >= 0) nt = labels[v0 + nb];
    }
    atomicAdd(&M[labels[v] * 6 + nt], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nrf = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) hist21[d * 21 + nrf++] = M[i * 6 + j];
  }
}

// ---- signature variants (GRSD-21 with subdivisions, GRSD-325, PlusGRSD-110) ----------------------

__global__ void __launch_bounds__(256) inverse_perm_kernel(const int* __restrict__ perm, int n, int* __restrict__ inv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = i;
}

// Mean normal of every voxel: pcl::VoxelGrid averages every field of the point type [EXTERNAL]; the
// result is not re-normalised (grsd_colorCHLAC_tools.hpp:558).  One thread per voxel, summed in
// (voxel, input index) order -- the radix sort is stable -- so the fp64 sums are reproducible.
__global__ void __launch_bounds__(128) voxel_normals_kernel(const float4* __restrict__ nrm, const int* __restrict__ inv_perm,
                                                            const int* __restrict__ ucount, const int* __restrict__ ustart,
                                                            const int* __restrict__ sorted_idx, int nvox,
                                                            float4* __restrict__ cnrm) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const int b = ustart[v], c = ucount[v];
  float sx = 0.f, sy = 0.f, sz = 0.f;  // the normals are fields of the same Eigen::VectorXf as the centroid
  for (int s = 0; s < c; ++s) {
    const float4 q = nrm[inv_perm[sorted_idx[b + s]]];
    sx = __fadd_rn(sx, q.x);
    sy = __fadd_rn(sy, q.y);
    sz = __fadd_rn(sz, q.z);
  }
  const float rcp = __fdiv_rn(1.0f, (float)c);
  cnrm[v] = make_float4(__fmul_rn(sx, rcp), __fmul_rn(sy, rcp), __fmul_rn(sz, rcp), 0.f);
}

struct SigDom {  // per cluster
  long long hist_base;  // first histogram of this cluster in the output
  int sb[3];            // subdivisions per axis
  int hist_num;         // 0: offsets exceed the grid, no output
};

constexpr int kSigSmemInts = 8192;  // histograms of a cluster are accumulated in shared memory up to this size

// Eigen's Vector3f::normalize() as used at grsd_colorCHLAC_tools.hpp:559-560: v / sqrt(v.v), fp32
__device__ __forceinline__ float3 unit_normal(float4 q) {
  const float sq = __fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z));
  const float len = __fsqrt_rn(sq);
  return make_float3(__fdiv_rn(q.x, len), __fdiv_rn(q.y, len), __fdiv_rn(q.z, len));
}

// kind 0: GRSD-21 (:230-276), 1: GRSD-325 (:396-431), 2: PlusGRSD-110 (:562-637).  Block (cluster, slice):
// counts go straight to their packed bin (entries below the diagonal of the transition matrices are
// never read by the reference and are dropped), into shared memory when the cluster's histograms fit.
template <int kKind>
__global__ void __launch_bounds__(128) signature_kernel(const VoxGrid* __restrict__ vg, const SigDom* __restrict__ sd,
                                                        const int* __restrict__ vox_off, const float4* __restrict__ cent,
                                                        const float4* __restrict__ cnrm, const int* __restrict__ labels,
                                                        const int* __restrict__ layout, float leaf, float inv_leaf,
                                                        int sub, float inv_sub, int off_x, int off_y, int off_z,
                                                        int* __restrict__ out, const unsigned char* __restrict__ own) {
  constexpr int kDim = kKind == 0 ? 21 : (kKind == 1 ? 325 : 110);
  constexpr int kOff = kKind == 1 ? 13 : 26;
  __shared__ int sh[kSigSmemInts];
  const int d = blockIdx.x;
  const SigDom s = sd[d];
  if (s.hist_num == 0) return;
  const bool use_sh = (long long)s.hist_num * kDim <= kSigSmemInts;
  int* H = use_sh ? sh : out + s.hist_base * kDim;
  if (use_sh) {
    for (int i = threadIdx.x; i < s.hist_num * kDim; i += blockDim.x) sh[i] = 0;
    __syncthreads();
  }
  const VoxGrid g = vg[d];
  const int v0 = vox_off[d], v1 = vox_off[d + 1];
  const long long work = (long long)(v1 - v0) * kOff;
  for (long long w = (long long)blockIdx.y * blockDim.x + threadIdx.x; w < work; w += (long long)gridDim.y * blockDim.x) {
    const int v = v0 + (int)(w / kOff), o = (int)(w % kOff);
    if (own && !own[v]) continue;  // a sharded cloud: the source voxels of this rank only (integer sums: the ranks' parts add up)
    const float4 c = cent[v];
    int hist_idx = 0;
    if (sub > 0) {  // :233-246: floor(x / voxel_size) - min_b - offset, fp32 division
      const int tx = (int)floorf(__fdiv_rn(c.x, leaf)) - g.min_b[0] - off_x;
      const int ty = (int)floorf(__fdiv_rn(c.y, leaf)) - g.min_b[1] - off_y;
      const int tz = (int)floorf(__fdiv_rn(c.z, leaf)) - g.min_b[2] - off_z;
      if (tx < 0 || ty < 0 || tz < 0) continue;
      const int ix = (int)floorf(__fmul_rn((float)tx, inv_sub)), iy = (int)floorf(__fmul_rn((float)ty, inv_sub)),
                iz = (int)floorf(__fmul_rn((float)tz, inv_sub));
      // A centroid on a voxel face can land one voxel beyond the grid here (the division above against the multiply
      // that built the grid): the reference then indexes past its histogram vector (:246-258, undefined behaviour on
      // the host); here the voxel is left out instead of writing into a neighbouring cluster's histograms.
      if (ix >= s.sb[0] || iy >= s.sb[1] || iz >= s.sb[2]) continue;
      hist_idx = ix + iy * s.sb[0] + iz * s.sb[0] * s.sb[1];
    }
    // pcl::VoxelGrid::getNeighborCentroidIndices [EXTERNAL]: ijk = floor(ref * inverse_leaf)
    const int n0 = (int)floorf(__fdiv_rn(c.x, leaf)) + c_off26[o][0] - g.min_b[0],
              n1 = (int)floorf(__fdiv_rn(c.y, leaf)) + c_off26[o][1] - g.min_b[1],
              n2 = (int)floorf(__fdiv_rn(c.z, leaf)) + c_off26[o][2] - g.min_b[2];
    int nb = -1;
    if (n0 >= 0 && n0 < g.div_b[0] && n1 >= 0 && n1 < g.div_b[1] && n2 >= 0 && n2 < g.div_b[2])
      nb = layout[g.layout_base + n0 + (long long)n1 * g.div_b[0] + (long long)n2 * g.div_b[0] * g.div_b[1]];
    const int src = labels[v];
    int bin = -1;
    if (kKind == 0) {
      const int nt = nb >= 0 ? labels[v0 + nb] : 5;
      if (src <= nt) bin = src * 6 - src * (src - 1) / 2 + (nt - src);  // row-major upper triangle of 6x6
    } else if (kKind == 1) {
      if (nb >= 0) bin = src + labels[v0 + nb] * 5 + o * 25;  // EMPTY neighbours are ignored (:427-428)
    } else {
      const float3 sn = unit_normal(cnrm[v]);
      if (!(isfinite(sn.x) && isfinite(sn.y) && isfinite(sn.z))) continue;  // :583
      bin = 105 + src;  // transitions_to_empty (:595-596, :609-610)
      if (nb >= 0) {
        const float3 m = unit_normal(cnrm[v0 + nb]);
        if (isfinite(m.x) && isfinite(m.y) && isfinite(m.z)) {
          // :607-608: min(NR_DIV-1, (int) floor(sqrt(source_normal.cross(nbr).norm()) * NR_DIV))
          const float cx = __fsub_rn(__fmul_rn(sn.y, m.z), __fmul_rn(sn.z, m.y));
          const float cy = __fsub_rn(__fmul_rn(sn.z, m.x), __fmul_rn(sn.x, m.z));
          const float cz = __fsub_rn(__fmul_rn(sn.x, m.y), __fmul_rn(sn.y, m.x));
          const float cn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
          const int ab = min(6, (int)floor(__dmul_rn(sqrt((double)cn), 7.0)));
          const int nt = labels[v0 + nb];
          bin = src <= nt ? ab * 15 + src * 5 - src * (src - 1) / 2 + (nt - src) : -1;
        }
      }
    }
    if (bin >= 0) atomicAdd(&H[(long long)hist_idx * kDim + bin], 1);
  }
  if (use_sh) {
    __syncthreads();
    int* dst = out + s.hist_base * kDim;
    for (int i = threadIdx.x; i < s.hist_num * kDim; i += blockDim.x) {
      const int c = sh[i];
      if (c) atomicAdd(dst + i, c);
    }
  }
}

float d2_threshold(double radius, float r2_hi, bool (*pred)(double, double, int, int), int b, int ndiv) {
  // smallest fp32 d2 in [0, r2_hi] with pred true (pred monotone in d2); INFINITY if none
  if (!pred(std::sqrt((double)r2_hi), radius, b, ndiv)) return INFINITY;
  if (pred(0.0, radius, b, ndiv)) return 0.f;
  uint32_t lo = 0, hi;
  std::memcpy(&hi, &r2_hi, 4);
  while (hi - lo > 1) {
    uint32_t mid = lo + (hi - lo) / 2;
    float f;
    std::memcpy(&f, &mid, 4);
    if (pred(std::sqrt((double)f), radius, b, ndiv)) hi = mid; else lo = mid;
  }
  float f;
  std::memcpy(&f, &hi, 4);
  return f;
}
bool pred_bin(double dist, double radius, int b, int ndiv) { return (int)std::floor(ndiv * dist / radius) >= b; }
bool pred_skip(double dist, double radius, int, int) { return dist > radius; }

// ---- colour half of VOSCH: rotation-invariant Color-CHLAC / C3-HLAC, 117 bins -------------------------------------
// color_chlac/include/color_chlac/color_chlac.hpp:1471-1528 (computeColorCHLAC), :1565-1743 (the four add functions),
// :1745-1782 (normalisation).  The reference adds integer colour products into float bins; the sums pass 2^24, so the
// result depends on the order of the additions.  One block per cluster, thread t owns bin t of every histogram of the
// cluster and walks the voxels in cloud order, the 13 half-stencil neighbours in the reference's order: every bin sees
// exactly the reference's sequence of float additions.

// pcl::VoxelGrid's voxel colour [EXTERNAL]: r, g, b are fields of the same Eigen::VectorXf as the centroid (float sums in
// cloud order, times 1.0f / count), packed back with truncation.
__global__ void __launch_bounds__(128) voxel_color_kernel(const unsigned* __restrict__ rgb, const int* __restrict__ ucount,
                                                          const int* __restrict__ ustart, const int* __restrict__ sorted_idx,
                                                          int nvox, unsigned* __restrict__ vrgb) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const int b = ustart[v], c = ucount[v];
  float sr = 0.f, sg = 0.f, sb = 0.f;
  for (int s = 0; s < c; ++s) {
    const unsigned col = rgb[sorted_idx[b + s]];
    sr = __fadd_rn(sr, (float)((col >> 16) & 0xffu));
    sg = __fadd_rn(sg, (float)((col >> 8) & 0xffu));
    sb = __fadd_rn(sb, (float)(col & 0xffu));
  }
  const float rcp = __fdiv_rn(1.0f, (float)c);
  vrgb[v] = ((unsigned)(int)__fmul_rn(sr, rcp) << 16) | ((unsigned)(int)__fmul_rn(sg, rcp) << 8) | (unsigned)(int)__fmul_rn(sb, rcp);
}

constexpr int kChlacDim = 117;
constexpr int kChlacSmemHists = 64;  // histograms of a cluster kept in shared memory (else accumulated in place in HBM)

__global__ void __launch_bounds__(128) color_chlac_kernel(const VoxGrid* __restrict__ vg, const SigDom* __restrict__ sd,
                                                          const int* __restrict__ vox_off, const float4* __restrict__ cent,
                                                          const unsigned* __restrict__ vrgb, const int* __restrict__ layout,
                                                          const int* __restrict__ lut,  // [2][256]: colour code, its complement
                                                          float leaf, int sub, float inv_sub, int off_x, int off_y, int off_z,
                                                          int thR, int thG, int thB, float* __restrict__ out) {
  __shared__ float sh[kChlacSmemHists * kChlacDim];
  const int d = blockIdx.x, t = threadIdx.x;
  const SigDom s = sd[d];
  if (s.hist_num == 0) return;
  const bool use_sh = s.hist_num <= kChlacSmemHists;
  float* H = use_sh ? sh : out + s.hist_base * kChlacDim;
  for (int i = t; i < s.hist_num * kChlacDim; i += blockDim.x) H[i] = 0.f;  // :1822-1824
  __syncthreads();
  if (t < kChlacDim) {
    // which addition of the reference feeds bin t
    int kind, a = 0, b = 0, c = 0;
    if (t < 6) { kind = 0; a = t; }                                             // addColorCHLAC_0: C[a]
    else if (t < 42) { kind = 1; a = (t - 6) / 6; b = (t - 6) % 6; }            // addColorCHLAC_1: C[a] * N[b]
    else if (t < 63) {                                                          // addColorCHLAC_0: C[a] * C[b], a <= b
      kind = 2;
      int k = t - 42;
      a = 0;
      while (k >= 6 - a) { k -= 6 - a; ++a; }
      b = a + k;
    } else if (t < 69) { kind = 3; a = (t - 63) / 2; b = (t - 63) & 1; }         // 0_bin: channel a set (b = 0) / clear (b = 1)
    else if (t < 105) { kind = 4; a = (t - 69) / 12; b = ((t - 69) % 12) / 6; c = (t - 69) % 6; }  // 1_bin
    else {                                                                      // 0_bin: pairs of channels
      kind = 5;
      const int k = t - 105;
      if (k < 8) { a = k < 4 ? 0 : 1; b = (k & 3) >> 1; c = k & 1; }              // r or !r (a) with g / b (b), plain / negated (c)
      else { a = 2 + ((k - 8) >> 1); b = 2; c = k & 1; }                          // g (2) or !g (3) with b
    }
    const VoxGrid g = vg[d];
    const int v0 = vox_off[d], v1 = vox_off[d + 1];
    for (int v = v0; v < v1; ++v) {
      const float4 ce = cent[v];
      int hist_idx = 0;
      if (s.hist_num != 1) {  // :1476-1500
        const int tx = (int)floorf(__fdiv_rn(ce.x, leaf)) - g.min_b[0] - off_x;
        const int ty = (int)floorf(__fdiv_rn(ce.y, leaf)) - g.min_b[1] - off_y;
        const int tz = (int)floorf(__fdiv_rn(ce.z, leaf)) - g.min_b[2] - off_z;
        if (tx < 0 || ty < 0 || tz < 0) continue;
        const int ix = (int)floorf(__fmul_rn((float)tx, inv_sub)), iy = (int)floorf(__fmul_rn((float)ty, inv_sub)),
                  iz = (int)floorf(__fmul_rn((float)tz, inv_sub));
        if (ix >= s.sb[0] || iy >= s.sb[1] || iz >= s.sb[2]) continue;  // see signature_kernel
        hist_idx = ix + iy * s.sb[0] + iz * s.sb[0] * s.sb[1];
      }
      float* bin = H + (size_t)hist_idx * kChlacDim + t;
      const unsigned col = vrgb[v];
      const int ch[3] = {(int)((col >> 16) & 0xffu), (int)((col >> 8) & 0xffu), (int)(col & 0xffu)};
      const int bins[3] = {ch[0] > thR ? 1 : 0, ch[1] > thG ? 1 : 0, ch[2] > thB ? 1 : 0};
      auto code = [&](const int* v3, int i) { return lut[(i & 1) * 256 + v3[i >> 1]]; };  // r, r_, g, g_, b, b_
      if (kind == 0) *bin = __fadd_rn(*bin, (float)code(ch, a));
      else if (kind == 2) *bin = __fadd_rn(*bin, (float)(code(ch, a) * code(ch, b)));
      else if (kind == 3) { if (bins[a] == 1 - b) *bin = __fadd_rn(*bin, 1.f); }
      else if (kind == 5) {
        bool first, second;
        if (a < 2) { first = bins[0] == 1 - a; second = bins[1 + b] == 1 - c; }
        else { first = bins[1] == 3 - a; second = bins[2] == 1 - c; }
        if (first && second) *bin = __fadd_rn(*bin, 1.f);
      } else {  // the first-order terms: one addition per occupied half-stencil neighbour (:1512-1527)
        if (kind == 4 && bins[a] != 1 - b) continue;
        const int i0 = (int)floorf(__fdiv_rn(ce.x, leaf)), i1 = (int)floorf(__fdiv_rn(ce.y, leaf)),
                  i2 = (int)floorf(__fdiv_rn(ce.z, leaf));
        const int ca = kind == 1 ? code(ch, a) : 0;
        for (int o = 0; o < 13; ++o) {
          const int n0 = i0 + c_off26[o][0] - g.min_b[0], n1 = i1 + c_off26[o][1] - g.min_b[1], n2 = i2 + c_off26[o][2] - g.min_b[2];
          if (n0 < 0 || n0 >= g.div_b[0] || n1 < 0 || n1 >= g.div_b[1] || n2 < 0 || n2 >= g.div_b[2]) continue;
          const int nb = layout[g.layout_base + n0 + (long long)n1 * g.div_b[0] + (long long)n2 * g.div_b[0] * g.div_b[1]];
          if (nb < 0) continue;
          const unsigned nc = vrgb[v0 + nb];
          const int nch[3] = {(int)((nc >> 16) & 0xffu), (int)((nc >> 8) & 0xffu), (int)(nc & 0xffu)};
          if (kind == 1) *bin = __fadd_rn(*bin, (float)(ca * code(nch, b)));
          else {
            const int set = nch[c >> 1] > (c >> 1 == 0 ? thR : (c >> 1 == 1 ? thG : thB)) ? 1 : 0;
            *bin = __fadd_rn(*bin, (float)((c & 1) ? 1 - set : set));
          }
        }
      }
    }
    // normalizeColorCHLAC (:1745-1782): float constants, equal for both classes without ENABLE_THEORY_NORMALIZATION
    const float norm = t < 6 ? (float)(1 / 255.0) : t < 42 ? (float)(1 / 845325.0) : t < 63 ? (float)(1 / 65025.0)
                     : t < 69 ? 1.f : t < 105 ? (float)(1 / 13.0) : 1.f;
    for (int h = 0; h < s.hist_num; ++h) H[(size_t)h * kChlacDim + t] = __fmul_rn(H[(size_t)h * kChlacDim + t], norm);
  }
  if (use_sh) {
    __syncthreads();
    for (int i = t; i < s.hist_num * kChlacDim; i += blockDim.x) out[s.hist_base * kChlacDim + i] = sh[i];
  }
}

}  // namespace

int run_grsd_batch(cab_ctx* ctx, float leaf, double r_rsd, int rsd_flags, int32_t* hist21, bool labels_only) {
  const int n = (int)ctx->n, nd = ctx->n_domains;
  cudaStream_t st = ctx->stream;
  if (!(leaf > 0.f)) return fail(ctx, CAB_ERR_ARG, "cab_grsd_batch: leaf must be > 0");
  const int ndiv = 5;              // PCL RSDEstimation default nr_subdiv, never overridden in the reference tree
  const double plane_radius = 0.2;  // PCL default plane_radius

  // ---- voxel grid descriptors (host, tiny) ---------------------------------------------
  const float inv_leaf = 1.0f / leaf;
  std::vector<VoxGrid> vg(nd);
  long long layout_total = 0;
  for (int d = 0; d < nd; ++d) {
    VoxGrid& g = vg[d];
    g.layout_base = layout_total;
    g.vox_first = 0;
    if (ctx->dom_count[d] == 0) {
      for (int a = 0; a < 3; ++a) g.min_b[a] = 0, g.div_b[a] = 0;
      continue;
    }
    long long cells = 1;
    for (int a = 0; a < 3; ++a) {
      g.min_b[a] = (int)std::floor(ctx->dom_bounds[6 * (size_t)d + a] * inv_leaf);
      const int max_b = (int)std::floor(ctx->dom_bounds[6 * (size_t)d + 3 + a] * inv_leaf);
      g.div_b[a] = max_b - g.min_b[a] + 1;
      cells *= g.div_b[a];
    }
    if (cells > 0x7fffffffll) return fail(ctx, CAB_ERR_OOM, "cab_grsd_batch: voxel grid of cluster %d too large", d);
    layout_total += cells;
  }
  if (layout_total > ((long long)1 << 31)) return fail(ctx, CAB_ERR_OOM, "cab_grsd_batch: leaf layouts too large");

  if (int rc = reserve(ctx, ctx->g_vgrid, nd * sizeof(VoxGrid))) return rc;
  if (int rc = reserve(ctx, ctx->g_vkeys[0], (size_t)std::max(n, 1) * 8)) return rc;
  if (int rc = reserve(ctx, ctx->g_vkeys[1], (size_t)std::max(n, 1) * 8)) return rc;
  if (int rc = reserve(ctx, ctx->g_vvals[0], (size_t)std::max(n, 1) * 4)) return rc;
  if (int rc = reserve(ctx, ctx->g_vvals[1], (size_t)std::max(n, 1) * 4)) return rc;
  if (int rc = reserve(ctx, ctx->g_layout, (size_t)std::max<long long>(layout_total, 1) * 4)) return rc;
  if (int rc = reserve(ctx, ctx->g_hist, (size_t)nd * 21 * 4)) return rc;
  if (int rc = reserve(ctx, ctx->g_voff, (size_t)(nd + 1) * 4)) return rc;
  // unique voxel arrays are bounded by n
  if (int rc = reserve(ctx, ctx->g_vfirst, ((size_t)std::max(n, 1) + 1) * 8 + 16)) return rc;       // unique keys (+ num_runs)
  if (int rc = reserve(ctx, ctx->g_vcount, ((size_t)std::max(n, 1) + 1) * 4 * 2)) return rc;       // counts + starts

  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
  if (int rc = reserve_pinned(ctx, nd * sizeof(VoxGrid) + (nd + 1) * 8)) return rc;
  std::memcpy(ctx->h_pin, vg.data(), nd * sizeof(VoxGrid));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->g_vgrid.p, ctx->h_pin, nd * sizeof(VoxGrid), cudaMemcpyHostToDevice, st));
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->g_layout.p, 0xff, (size_t)std::max<long long>(layout_total, 1) * 4, st));
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->g_hist.p, 0, (size_t)nd * 21 * 4, st));

  unsigned long long* ukeys = (unsigned long long*)ctx->g_vfirst.p;
  int* num_runs = (int*)((char*)ctx->g_vfirst.p + ((size_t)std::max(n, 1) + 1) * 8);
  int* ucount = (int*)ctx->g_vcount.p;
  int* ustart = ucount + (std::max(n, 1) + 1);
  int nvox = 0;
  std::vector<int> vox_off(nd + 1, 0);
  if (n > 0) {
    voxel_key_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, (const int*)ctx->b_domoff.p, nd,
                                                      (const VoxGrid*)ctx->g_vgrid.p, inv_leaf,
                                                      (unsigned long long*)ctx->g_vkeys[0].p, (int*)ctx->g_vvals[0].p);
    CAB_LAUNCH_CHECK(ctx);
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, n, 0, 64, st);
    cub::DeviceRunLengthEncode::Encode(nullptr, t2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                       (int*)nullptr, (int*)nullptr, n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, (const int*)nullptr, (int*)nullptr, n + 1, st);
    size_t tmp = std::max(t1, std::max(t2, t3));
    if (int rc = reserve(ctx, ctx->b_cubtmp, tmp + 16)) return rc;
    int dom_bits = 1;
    while ((1 << dom_bits) < nd) ++dom_bits;
    CAB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_cubtmp.p, tmp, (const unsigned long long*)ctx->g_vkeys[0].p,
                                                  (unsigned long long*)ctx->g_vkeys[1].p, (const int*)ctx->g_vvals[0].p,
                                                  (int*)ctx->g_vvals[1].p, n, 0, 64, st));
    CAB_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(ctx->b_cubtmp.p, tmp, (const unsigned long long*)ctx->g_vkeys[1].p,
                                                     ukeys, ucount, num_runs, n, st));
    ctx->tm.kernel_launches += 9 + 3;
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, num_runs, 4, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    int runs = *(const int*)ctx->h_pin;
    // the last run may be the non-finite sentinel key
    std::vector<unsigned long long> hkeys;
    hkeys.resize(runs);
    if (runs) CAB_CUDA(ctx, cudaMemcpy(hkeys.data(), ukeys, (size_t)runs * 8, cudaMemcpyDeviceToHost));
    nvox = runs;
    if (runs && hkeys[runs - 1] == ~0ull) nvox = runs - 1;
    // voxel range of every cluster
    int v = 0;
    for (int d = 0; d < nd; ++d) {
      vox_off[d] = v;
      while (v < nvox && (int)(hkeys[v] >> 32) == d) ++v;
    }
    vox_off[nd] = v;
    for (int d = 0; d < nd; ++d) vg[d].vox_first = vox_off[d];
    std::memcpy(ctx->h_pin, vg.data(), nd * sizeof(VoxGrid));
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->g_vgrid.p, ctx->h_pin, nd * sizeof(VoxGrid), cudaMemcpyHostToDevice, st));
    CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp, ucount, ustart, runs + 1, st));
    ctx->tm.kernel_launches += 2;
  }
  {
    int* hv = (int*)((char*)ctx->h_pin + nd * sizeof(VoxGrid));
    std::memcpy(hv, vox_off.data(), (nd + 1) * 4);
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->g_voff.p, hv, (nd + 1) * 4, cudaMemcpyHostToDevice, st));
  }
  ctx->g_nvox = nvox;
  ctx->g_vox_offsets.assign(vox_off.begin(), vox_off.end());
  ctx->g_min_div.resize((size_t)nd * 6);
  for (int d = 0; d < nd; ++d)
    for (int a = 0; a < 3; ++a) {
      ctx->g_min_div[6 * (size_t)d + a] = vg[d].min_b[a];
      ctx->g_min_div[6 * (size_t)d + 3 + a] = vg[d].div_b[a];
    }
  ctx->g_leaf = leaf;
  ctx->g_have_cnrm = false;
  ctx->g_own_valid = false;
  if (labels_only)
    if (int rc = reserve(ctx, ctx->g_vown, (size_t)std::max(nvox, 1))) return rc;
  if (int rc = reserve(ctx, ctx->g_cent, (size_t)std::max(nvox, 1) * sizeof(float4))) return rc;
  if (int rc = reserve(ctx, ctx->g_vrad, (size_t)std::max(nvox, 1) * sizeof(float2))) return rc;
  if (int rc = reserve(ctx, ctx->g_vlabel, (size_t)std::max(nvox, 1) * sizeof(int))) return rc;

  if (nvox > 0) {
    const unsigned blocks = (unsigned)(((size_t)nvox * kWarp + 255) / 256);
    centroid_kernel<<<blocks, 256, 0, st>>>(ctx->xyz_in, ctx->stride, ukeys, ucount, ustart, (const int*)ctx->g_vvals[1].p,
                                            nvox, (const VoxGrid*)ctx->g_vgrid.p, (float4*)ctx->g_cent.p,
                                            (int*)ctx->g_layout.p);
    CAB_LAUNCH_CHECK(ctx);
    VRsdArgs a{};
    a.g = grid_view(ctx);
    a.nrm = (const float4*)ctx->b_nrm.p;
    a.cent = (const float4*)ctx->g_cent.p;
    a.ukeys = ukeys;
    a.nvox = nvox;
    const float rf = (float)r_rsd;
    a.r2 = rf * rf;
    a.ndiv = ndiv;
    a.flags = rsd_flags;
    a.radius = r_rsd;
    a.plane_radius = plane_radius;
    // distances from the reference element can reach 2r: thresholds searched up to (2r)^2 * 1.01
    const float hi2 = 4.04f * a.r2;
    a.skip_thr = d2_threshold(r_rsd, hi2, pred_skip, 0, ndiv);
    a.bin_thr[0] = -INFINITY;
    for (int b = 1; b < ndiv; ++b) a.bin_thr[b] = d2_threshold(r_rsd, hi2, pred_bin, b, ndiv);
    for (int b = ndiv; b < 9; ++b) a.bin_thr[b] = INFINITY;
    a.radii = (float2*)ctx->g_vrad.p;
    a.labels = (int*)ctx->g_vlabel.p;
    a.slab = labels_only ? slab_info_device(ctx) : nullptr;  // null unless the grid is one rank's slab
    a.own = labels_only ? (unsigned char*)ctx->g_vown.p : nullptr;
    voxel_rsd_kernel<<<blocks, 256, 0, st>>>(a);
    CAB_LAUNCH_CHECK(ctx);
  }
  {
    static const int off13[13][3] = {{-1, -1, -1}, {-1, 0, -1}, {-1, 1, -1}, {0, -1, -1}, {0, 0, -1}, {0, 1, -1}, {1, -1, -1},
                                     {1, 0, -1},   {1, 1, -1},  {-1, -1, 0}, {0, -1, 0},  {1, -1, 0}, {-1, 0, 0}};
    static int off26[26][3];
    for (int c = 0; c < 13; ++c)
      for (int k = 0; k < 3; ++k) {
        off26[c][k] = off13[c][k];
        off26[13 + c][k] = -off13[c][k];
      }
    CAB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_off26, off26, sizeof(off26), 0, cudaMemcpyHostToDevice, st));
  }
  if (labels_only) {  // the transitions wait for the other ranks' labels (cab_grsd_cloud)
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.grsd_ms, ctx->ev[4], ctx->ev[5]));
    ctx->g_own_valid = ctx->slab;
    return CAB_OK;
  }
  transitions_kernel<<<nd, 128, 0, st>>>((const VoxGrid*)ctx->g_vgrid.p, (const int*)ctx->g_voff.p, (const float4*)ctx->g_cent.p,
                                         (const int*)ctx->g_vlabel.p, (const int*)ctx->g_layout.p, leaf,
                                         (int*)ctx->g_hist.p);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
  if (hist21) CAB_CUDA(ctx, cudaMemcpyAsync(hist21, ctx->g_hist.p, (size_t)nd * 21 * 4, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.grsd_ms, ctx->ev[4], ctx->ev[5]));
  return CAB_OK;
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_grsd_batch(cab_ctx* ctx, const float* xyz, int32_t stride, const int32_t* offsets, int32_t nclusters, float leaf,
                   float r_normals, double rsd_radius_min, int32_t rsd_flags, const float vp[3], const float* nx,
                   const float* ny, const float* nz, int32_t* hist21) {
  if (!ctx) return CAB_ERR_ARG;
  if (!offsets || nclusters < 1) return fail(ctx, CAB_ERR_ARG, "cab_grsd_batch: offsets / nclusters");
  if (!(leaf > 0.f)) return fail(ctx, CAB_ERR_ARG, "cab_grsd_batch: leaf must be > 0");
  const bool have_n = nx && ny && nz;
  if (!have_n && (nx || ny || nz)) return fail(ctx, CAB_ERR_ARG, "cab_grsd_batch: give all of nx, ny, nz or none");
  const int64_t n = offsets[nclusters];
  // A batch shards by cluster (each rank of a group gets its own clusters), never by rows: every cluster of this call
  // is processed here, whatever cab_set_shard says.
  struct WholeBatch {
    cab_ctx* c;
    int r, w;
    explicit WholeBatch(cab_ctx* x) : c(x), r(x->shard_rank), w(x->shard_world) { c->shard_rank = 0; c->shard_world = 1; }
    ~WholeBatch() { c->shard_rank = r; c->shard_world = w; }
  } whole(ctx);
  if (int rc = cab_upload_clusters(ctx, xyz, n, stride, offsets, nclusters)) return rc;
  // grsd_colorCHLAC_tools.hpp:172: std::max(rsd_radius_search, voxel_size/2 * sqrt(3)); float/2 * double
  const double r_rsd = std::max(rsd_radius_min, (double)(leaf / 2) * std::sqrt(3.0));
  float cell = (float)r_rsd;
  if ((double)cell < r_rsd) cell = std::nextafter(cell, INFINITY);
  if (!have_n) cell = std::max(cell, r_normals);
  if (int rc = cab_build_grid(ctx, cell)) return rc;
  if (have_n) {
    if (int rc = cab_set_normals(ctx, nx, ny, nz)) return rc;
  } else {
    if (int rc = cab_normals(ctx, r_normals, 0, vp, nullptr)) return rc;
  }
  return run_grsd_batch(ctx, leaf, r_rsd, rsd_flags, hist21);
}

int64_t cab_grsd_cloud_labels(cab_ctx* ctx, float leaf, float r_normals, double rsd_radius_min, int32_t rsd_flags,
                              const float vp[3], int32_t* labels_plus1, int64_t cap) {
  if (!ctx) return CAB_ERR_ARG;
  if (!ctx->have_cloud) return fail(ctx, CAB_ERR_STATE, "cab_grsd_cloud: no cloud uploaded");
  if (ctx->n_domains != 1) return fail(ctx, CAB_ERR_STATE, "cab_grsd_cloud: works on one cloud (a batch of clusters: cab_grsd_batch)");
  if (!(leaf > 0.f)) return fail(ctx, CAB_ERR_ARG, "cab_grsd_cloud: leaf must be > 0");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cudaSetDevice failed");
  // grsd_colorCHLAC_tools.hpp:172: std::max(rsd_radius_search, voxel_size/2 * sqrt(3)); float/2 * double
  const double r_rsd = std::max(rsd_radius_min, (double)(leaf / 2) * std::sqrt(3.0));
  float cell = (float)r_rsd;
  if ((double)cell < r_rsd) cell = std::nextafter(cell, INFINITY);
  cell = std::max(cell, r_normals);
  // a rank of a sharded context builds its slab of rows (own rows, one layer of halo rows whose normals it computes
  // itself, one more layer of candidates): the points within r_rsd <= cell of a centroid in an own row all have normals
  ctx->want_halo_exchange = false;
  if (int rc = build_grid(ctx, cell)) return rc;
  if (int rc = run_normals(ctx, r_normals, 0, vp)) return rc;
  if (int rc = run_grsd_batch(ctx, leaf, r_rsd, rsd_flags, nullptr, true)) return rc;
  const int64_t nv = ctx->g_nvox;
  if (!labels_plus1 || nv == 0 || nv > cap) return nv;
  std::vector<int32_t> lab((size_t)nv);
  CAB_CUDA(ctx, cudaMemcpy(lab.data(), ctx->g_vlabel.p, (size_t)nv * sizeof(int), cudaMemcpyDeviceToHost));
  for (int64_t v = 0; v < nv; ++v) labels_plus1[v] = lab[(size_t)v] + 1;  // 0: not this rank's voxel
  return nv;
}

int cab_grsd_cloud_set_labels(cab_ctx* ctx, const int32_t* labels_plus1, int64_t count) {
  if (!ctx || !labels_plus1) return CAB_ERR_ARG;
  if (count != ctx->g_nvox) return fail(ctx, CAB_ERR_ARG, "cab_grsd_cloud_set_labels: %lld labels for %lld voxels", (long long)count, (long long)ctx->g_nvox);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cudaSetDevice failed");
  std::vector<int32_t> lab((size_t)count);
  for (int64_t v = 0; v < count; ++v) {
    if (labels_plus1[v] < 1 || labels_plus1[v] > 5)
      return fail(ctx, CAB_ERR_ARG, "cab_grsd_cloud_set_labels: voxel %lld has no label (the ranks' labels were not merged?)", (long long)v);
    lab[(size_t)v] = labels_plus1[v] - 1;
  }
  if (count) CAB_CUDA(ctx, cudaMemcpy(ctx->g_vlabel.p, lab.data(), (size_t)count * sizeof(int), cudaMemcpyHostToDevice));
  return CAB_OK;
}

int cab_grsd_cloud(cab_ctx* ctx, float leaf, float r_normals, double rsd_radius_min, int32_t rsd_flags, const float vp[3],
                   int32_t* hist21) {
  if (!ctx || !hist21) return CAB_ERR_ARG;
  if (ctx->shard_world > 1 && !comm_active(ctx))
    return fail(ctx, CAB_ERR_STATE, "cab_grsd_cloud: a sharded context outside a group -- merge the ranks' labels and histograms "
                                    "yourself (cab_grsd_cloud_labels, cab_grsd_cloud_set_labels, cab_grsd_signatures) or join a group");
  const int64_t nv = cab_grsd_cloud_labels(ctx, leaf, r_normals, rsd_radius_min, rsd_flags, vp, nullptr, 0);
  if (nv < 0) return (int)nv;
  std::memset(hist21, 0, 21 * sizeof(int32_t));
  std::vector<int32_t> lab((size_t)nv);
  if (nv > 0) {
    CAB_CUDA(ctx, cudaMemcpy(lab.data(), ctx->g_vlabel.p, (size_t)nv * sizeof(int), cudaMemcpyDeviceToHost));
    for (auto& l : lab) l += 1;
    // every voxel was labelled by exactly one rank: the sum of the ranks' (label + 1 | 0) arrays is the cloud's labelling
    if (comm_active(ctx))
      if (int rc = cab_comm_allreduce_i32(ctx, lab.data(), nv)) return rc;
    if (int rc = cab_grsd_cloud_set_labels(ctx, lab.data(), nv)) return rc;
  } else if (comm_active(ctx)) {
    if (int rc = cab_comm_allreduce_i32(ctx, lab.data(), 0)) return rc;
  }
  // GRSD-21 of the whole cloud: this rank's source voxels, then the integer all-reduce inside cab_grsd_signatures
  const int64_t got = cab_grsd_signatures(ctx, CAB_SIG_GRSD21, 0, 0, 0, 0, nullptr, nullptr, hist21, 1);
  return got < 0 ? (int)got : CAB_OK;
}

int64_t cab_grsd_voxels(cab_ctx* ctx, int64_t* vox_offsets, float* centroids_xyz, float* r_min, float* r_max,
                        int32_t* labels, int64_t cap) {
  if (!ctx) return CAB_ERR_ARG;
  const int64_t nv = ctx->g_nvox;
  if (vox_offsets)
    for (size_t i = 0; i < ctx->g_vox_offsets.size(); ++i) vox_offsets[i] = ctx->g_vox_offsets[i];
  if (nv == 0 || nv > cap) return nv;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cudaSetDevice failed");
  std::vector<float4> c(nv);
  std::vector<float2> r(nv);
  CAB_CUDA(ctx, cudaMemcpy(c.data(), ctx->g_cent.p, nv * sizeof(float4), cudaMemcpyDeviceToHost));
  CAB_CUDA(ctx, cudaMemcpy(r.data(), ctx->g_vrad.p, nv * sizeof(float2), cudaMemcpyDeviceToHost));
  if (labels) CAB_CUDA(ctx, cudaMemcpy(labels, ctx->g_vlabel.p, nv * sizeof(int), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < nv; ++i) {
    if (centroids_xyz) {
      centroids_xyz[3 * i] = c[i].x;
      centroids_xyz[3 * i + 1] = c[i].y;
      centroids_xyz[3 * i + 2] = c[i].z;
    }
    if (r_min) r_min[i] = r[i].x;
    if (r_max) r_max[i] = r[i].y;
  }
  return nv;
}

int64_t cab_grsd_signatures(cab_ctx* ctx, int32_t kind, int32_t subdivision_size, int32_t off_x, int32_t off_y,
                            int32_t off_z, int64_t* hist_offsets, int32_t* subdiv_b, int32_t* hist, int64_t cap) {
  if (!ctx) return CAB_ERR_ARG;
  if (kind < CAB_SIG_GRSD21 || kind > CAB_SIG_PLUSGRSD110) return fail(ctx, CAB_ERR_ARG, "cab_grsd_signatures: unknown kind %d", kind);
  if (subdivision_size < 0) return fail(ctx, CAB_ERR_ARG, "cab_grsd_signatures: invalid subdivision size %d", subdivision_size);
  const int nd = (int)ctx->g_min_div.size() / 6;
  if (nd == 0 || ctx->g_vox_offsets.size() != (size_t)nd + 1)
    return fail(ctx, CAB_ERR_STATE, "cab_grsd_signatures: run cab_grsd_batch first");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cudaSetDevice failed");
  const int dim = kind == CAB_SIG_GRSD21 ? 21 : (kind == CAB_SIG_GRSD325 ? 325 : 110);
  // subdivision bookkeeping per cluster (grsd_colorCHLAC_tools.hpp:140-161), host, tiny
  const float inv_sub = subdivision_size > 0 ? (float)(1.0 / subdivision_size) : 0.f;
  std::vector<SigDom> sd(nd);
  int64_t total = 0;
  int max_vox = 0;
  for (int d = 0; d < nd; ++d) {
    const int32_t* div_b = ctx->g_min_div.data() + 6 * (size_t)d + 3;
    SigDom& s = sd[d];
    s.hist_base = total;
    s.sb[0] = s.sb[1] = s.sb[2] = 1;
    s.hist_num = 1;
    if (subdivision_size > 0) {
      if (div_b[0] <= off_x || div_b[1] <= off_y || div_b[2] <= off_z) {
        s.sb[0] = s.sb[1] = s.sb[2] = 0;
        s.hist_num = 0;
      } else {
        s.sb[0] = (int)std::ceil((div_b[0] - off_x) * inv_sub);
        s.sb[1] = (int)std::ceil((div_b[1] - off_y) * inv_sub);
        s.sb[2] = (int)std::ceil((div_b[2] - off_z) * inv_sub);
        s.hist_num = s.sb[0] * s.sb[1] * s.sb[2];
      }
    }
    if (hist_offsets) hist_offsets[d] = total;
    if (subdiv_b)
      for (int a = 0; a < 3; ++a) subdiv_b[3 * (size_t)d + a] = s.sb[a];
    total += s.hist_num;
    max_vox = std::max<int>(max_vox, (int)(ctx->g_vox_offsets[d + 1] - ctx->g_vox_offsets[d]));
  }
  if (hist_offsets) hist_offsets[nd] = total;
  if (!hist || total == 0) return total;
  if (total > cap) return fail(ctx, CAB_ERR_ARG, "cab_grsd_signatures: %lld histograms, room for %lld", (long long)total, (long long)cap);
  cudaStream_t st = ctx->stream;
  const int n = (int)ctx->n;
  const int nvox = (int)ctx->g_nvox;
  if (kind == CAB_SIG_PLUSGRSD110 && ctx->g_own_valid)
    return fail(ctx, CAB_ERR_STATE, "cab_grsd_signatures: PlusGRSD needs the normals of every point of a voxel; a sharded cloud "
                                    "(cab_grsd_cloud on a group) holds those of its own rows only");
  if (kind == CAB_SIG_PLUSGRSD110 && !ctx->g_have_cnrm && nvox > 0) {
    if (!ctx->have_normals) return fail(ctx, CAB_ERR_STATE, "cab_grsd_signatures: no normals");
    if (int rc = reserve(ctx, ctx->g_invperm, (size_t)std::max(n, 1) * 4)) return rc;
    if (int rc = reserve(ctx, ctx->g_cnrm, (size_t)nvox * sizeof(float4))) return rc;
    inverse_perm_kernel<<<(n + 255) / 256, 256, 0, st>>>((const int*)ctx->b_perm.p, n, (int*)ctx->g_invperm.p);
    CAB_LAUNCH_CHECK(ctx);
    const int* ucount = (const int*)ctx->g_vcount.p;
    const int* ustart = ucount + (std::max(n, 1) + 1);
    voxel_normals_kernel<<<(nvox + 127) / 128, 128, 0, st>>>((const float4*)ctx->b_nrm.p, (const int*)ctx->g_invperm.p, ucount, ustart,
                                                            (const int*)ctx->g_vvals[1].p, nvox, (float4*)ctx->g_cnrm.p);
    CAB_LAUNCH_CHECK(ctx);
    ctx->g_have_cnrm = true;
  }
  if (int rc = reserve(ctx, ctx->g_sig, (size_t)total * dim * 4)) return rc;
  if (int rc = reserve(ctx, ctx->g_sigdom, (size_t)nd * sizeof(SigDom))) return rc;
  if (int rc = reserve_pinned(ctx, (size_t)nd * sizeof(SigDom))) return rc;
  std::memcpy(ctx->h_pin, sd.data(), (size_t)nd * sizeof(SigDom));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->g_sigdom.p, ctx->h_pin, (size_t)nd * sizeof(SigDom), cudaMemcpyHostToDevice, st));
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->g_sig.p, 0, (size_t)total * dim * 4, st));
  if (nvox > 0) {
    const int noff = kind == CAB_SIG_GRSD325 ? 13 : 26;
    const unsigned slices = (unsigned)std::min<long long>(std::max<long long>(((long long)max_vox * noff + 2047) / 2048, 1), 1024);
    const dim3 grid((unsigned)nd, slices);
    const float leaf = ctx->g_leaf, inv_leaf = 1.0f / leaf;
    const unsigned char* own_mask = ctx->g_own_valid ? (const unsigned char*)ctx->g_vown.p : nullptr;
#define CAB_SIG_LAUNCH(K)                                                                                                  \
  signature_kernel<K><<<grid, 128, 0, st>>>((const VoxGrid*)ctx->g_vgrid.p, (const SigDom*)ctx->g_sigdom.p,                \
                                            (const int*)ctx->g_voff.p, (const float4*)ctx->g_cent.p,                      \
                                            (const float4*)ctx->g_cnrm.p, (const int*)ctx->g_vlabel.p,                    \
                                            (const int*)ctx->g_layout.p, leaf, inv_leaf, subdivision_size, inv_sub, off_x, \
                                            off_y, off_z, (int*)ctx->g_sig.p, own_mask)
    if (kind == CAB_SIG_GRSD21) CAB_SIG_LAUNCH(0);
    else if (kind == CAB_SIG_GRSD325) CAB_SIG_LAUNCH(1);
    else CAB_SIG_LAUNCH(2);
#undef CAB_SIG_LAUNCH
    CAB_LAUNCH_CHECK(ctx);
  }
  CAB_CUDA(ctx, cudaMemcpyAsync(hist, ctx->g_sig.p, (size_t)total * dim * 4, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  // one large cloud sharded over a group (cab_grsd_cloud): this rank counted the transitions of its own voxels; one small
  // integer all-reduce makes every rank's copy the whole cloud's (grsd_colorCHLAC_tools.hpp:230-260 sums the same integers)
  if (ctx->g_own_valid && comm_active(ctx))
    if (int rc = cab_comm_allreduce_i32(ctx, hist, total * dim)) return rc;
  return total;
}

int64_t cab_color_chlac(cab_ctx* ctx, const uint32_t* rgb, int32_t c3, int32_t thR, int32_t thG, int32_t thB,
                        int32_t subdivision_size, int32_t off_x, int32_t off_y, int32_t off_z, int64_t* hist_offsets,
                        int32_t* subdiv_b, float* hist, int64_t cap) {
  if (!ctx) return CAB_ERR_ARG;
  if (thR < 0 || thG < 0 || thB < 0) return fail(ctx, CAB_ERR_ARG, "cab_color_chlac: Invalid color_threshold: %d %d %d", thR, thG, thB);
  if (subdivision_size < 0) return fail(ctx, CAB_ERR_ARG, "cab_color_chlac: Invalid subdivision size: %d", subdivision_size);
  const int nd = (int)ctx->g_min_div.size() / 6;
  if (nd == 0 || ctx->g_vox_offsets.size() != (size_t)nd + 1)
    return fail(ctx, CAB_ERR_STATE, "cab_color_chlac: run cab_grsd_batch first");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cudaSetDevice failed");
  // subdivision bookkeeping per cluster (setVoxelFilter, color_chlac.hpp:181-210), host, tiny
  const float inv_sub = subdivision_size > 0 ? (float)(1.0 / subdivision_size) : 0.f;
  std::vector<SigDom> sd(nd);
  int64_t total = 0;
  for (int d = 0; d < nd; ++d) {
    const int32_t* div_b = ctx->g_min_div.data() + 6 * (size_t)d + 3;
    SigDom& s = sd[d];
    s.hist_base = total;
    s.sb[0] = s.sb[1] = s.sb[2] = 1;
    s.hist_num = 1;
    if (subdivision_size > 0) {
      if (div_b[0] <= off_x || div_b[1] <= off_y || div_b[2] <= off_z) {
        s.sb[0] = s.sb[1] = s.sb[2] = 0;
        s.hist_num = 0;
      } else {
        s.sb[0] = (int)std::ceil((div_b[0] - off_x) * inv_sub);
        s.sb[1] = (int)std::ceil((div_b[1] - off_y) * inv_sub);
        s.sb[2] = (int)std::ceil((div_b[2] - off_z) * inv_sub);
        s.hist_num = s.sb[0] * s.sb[1] * s.sb[2];
      }
    }
    if (hist_offsets) hist_offsets[d] = total;
    if (subdiv_b)
      for (int a = 0; a < 3; ++a) subdiv_b[3 * (size_t)d + a] = s.sb[a];
    total += s.hist_num;
  }
  if (hist_offsets) hist_offsets[nd] = total;
  if (!hist || total == 0) return total;
  if (!rgb) return fail(ctx, CAB_ERR_ARG, "cab_color_chlac: no colours");
  if (total > cap) return fail(ctx, CAB_ERR_ARG, "cab_color_chlac: %lld histograms, room for %lld", (long long)total, (long long)cap);
  cudaStream_t st = ctx->stream;
  const int n = (int)ctx->n;
  const int nvox = (int)ctx->g_nvox;
  // colour code tables (setColor, color_chlac.hpp:148-166).  C3-HLAC: 255 * sin / cos (v * angle_norm) with the float
  // angle_norm = M_PI / 510 (color_chlac.h:9); the unqualified sin / cos of a float argument resolve to the double
  // functions of <math.h> with the toolchain of the reference's era, the product is truncated to int.
  int lut[512];
  const float angle_norm = M_PI / 510;
  for (int v = 0; v < 256; ++v) {
    if (c3) {
      lut[v] = 255 * ::sin((double)(v * angle_norm));
      lut[256 + v] = 255 * ::cos((double)(v * angle_norm));
    } else {
      lut[v] = v;
      lut[256 + v] = 255 - v;
    }
  }
  const size_t rgb_bytes = (size_t)std::max(n, 1) * 4, vrgb_bytes = (size_t)std::max(nvox, 1) * 4;
  if (int rc = reserve(ctx, ctx->g_color, rgb_bytes + vrgb_bytes + sizeof(lut) + 64)) return rc;
  if (int rc = reserve(ctx, ctx->g_sig, (size_t)total * kChlacDim * 4)) return rc;
  if (int rc = reserve(ctx, ctx->g_sigdom, (size_t)nd * sizeof(SigDom))) return rc;
  if (int rc = reserve_pinned(ctx, (size_t)nd * sizeof(SigDom) + sizeof(lut))) return rc;
  unsigned* d_rgb = (unsigned*)ctx->g_color.p;
  unsigned* d_vrgb = d_rgb + std::max(n, 1);
  int* d_lut = (int*)(d_vrgb + std::max(nvox, 1));
  std::memcpy(ctx->h_pin, sd.data(), (size_t)nd * sizeof(SigDom));
  std::memcpy((char*)ctx->h_pin + (size_t)nd * sizeof(SigDom), lut, sizeof(lut));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->g_sigdom.p, ctx->h_pin, (size_t)nd * sizeof(SigDom), cudaMemcpyHostToDevice, st));
  CAB_CUDA(ctx, cudaMemcpyAsync(d_lut, (char*)ctx->h_pin + (size_t)nd * sizeof(SigDom), sizeof(lut), cudaMemcpyHostToDevice, st));
  if (n > 0) CAB_CUDA(ctx, cudaMemcpyAsync(d_rgb, rgb, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->g_sig.p, 0, (size_t)total * kChlacDim * 4, st));
  if (nvox > 0) {
    const int* ucount = (const int*)ctx->g_vcount.p;
    const int* ustart = ucount + (std::max(n, 1) + 1);
    voxel_color_kernel<<<(nvox + 127) / 128, 128, 0, st>>>(d_rgb, ucount, ustart, (const int*)ctx->g_vvals[1].p, nvox, d_vrgb);
    CAB_LAUNCH_CHECK(ctx);
    color_chlac_kernel<<<nd, 128, 0, st>>>((const VoxGrid*)ctx->g_vgrid.p, (const SigDom*)ctx->g_sigdom.p, (const int*)ctx->g_voff.p,
                                           (const float4*)ctx->g_cent.p, d_vrgb, (const int*)ctx->g_layout.p, d_lut, ctx->g_leaf,
                                           subdivision_size, inv_sub, off_x, off_y, off_z, thR, thG, thB, (float*)ctx->g_sig.p);
    CAB_LAUNCH_CHECK(ctx);
  }
  CAB_CUDA(ctx, cudaMemcpyAsync(hist, ctx->g_sig.p, (size_t)total * kChlacDim * 4, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  return total;
}

}  // extern "C"
