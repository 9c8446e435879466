// cab_topk.cu -- max_nn truncation thresholds and the neighbour-set debug path.
//
// The reference keeps at most max_nn_ neighbours per query
// (kdtree_->radiusSearch(cp, radius_, idx, d2, max_nn_), cloud_algos/src/radius_estimation.cpp:120;
// default 150 at radius_estimation.h:82, 75 in launch/pipeline_tmp.launch:20).  The documented rule
// is "the max_nn smallest (d2, input index) pairs".  Instead of materialising and sorting lists,
// every query gets a threshold pair (d2*, idx*); the normals / RSD kernels then accept a candidate iff
// (d2, idx) <= (d2*, idx*).  Two traversals find it (value_select_kernel): a 64-bin histogram of d2 over
// [0, r2] locates the bin holding the max_nn-th neighbour (on a surface the in-radius d2 values are spread
// evenly, so that bin holds k / 64 candidates), the second traversal lists that bin's candidates and ranks
// them by (d2, index).  Packets where a query's bin overflows the list (many equal distances) fall back to
// the radix select over the fp32 bit pattern of d2 (threshold_kernel, 6 bits per traversal).
#include <climits>
#include <cmath>
#include <cstring>

#include <vector>

#include "cab_internal.cuh"
#include "cab_traverse.cuh"

namespace cab {

namespace {

constexpr int kSelWarps = 4;
constexpr int kSelBins = kTruncBins;

struct ThrArgs {
  GridView g;
  int p0, p1;
  float r, r2;
  int max_nn;
  int first_shift;  // 30 when r2 needs bit 30/31, else 24
  float* thr_d2;
  int* thr_idx;
  const unsigned char* done;  // optional, indexed by input index: queries that need no threshold any more
  unsigned char* fallback;    // per packet (relative to p0): value_select_kernel sets it, threshold_kernel honours it
  int use_fallback;           // threshold_kernel: only the flagged packets
  const unsigned char* only;  // optional, per packet (relative to p0): value_select_kernel works on the flagged packets only
  int* code;                  // nn_hist_kernel: per query (sorted order) target bin | rank inside it << 8
  const float4* nrm;          // nn_hist_kernel: normals (sorted order); packets with a candidate without one are flagged
  unsigned long long* stats;  // nn_hist_kernel: packet work counter
};

// (one packet; the kernels below walk the packets with a grid-stride loop: in the truncated fast RSD pass only a few
// packets are flagged, and a grid of one block per four packets costs 0.8 ms on 660 k packets just to be scheduled)
__device__ __forceinline__ void threshold_packet(const ThrArgs& a, int pid) {
  __shared__ ChunkTile tiles[kSelWarps];
  __shared__ unsigned hist[kSelWarps][kSelBins][kWarp];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (pid >= a.p1) return;
  if (a.use_fallback && !a.fallback[pid - a.p0]) return;
  const GridView& g = a.g;
  ChunkTile* tile = &tiles[warp];
  const PacketCtx pc = load_packet(g, pid, lane, a.r, tile);
  const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z, r2 = a.r2;

  unsigned prefix = 0, mask = 0;
  int target = a.max_nn;
  bool need = pc.active;
  if (a.done) {
    need = need && !a.done[g.perm[pc.qi]];
    if (!__any_sync(kFull, need)) return;  // whole packet already resolved (k-NN rounds)
  }
  const bool wanted = need;
  for (int shift = a.first_shift; shift >= 0; shift -= 6) {
    for (int b = 0; b < kSelBins; ++b) hist[warp][b][lane] = 0;
    for_each_chunk(g, pc, lane, tile, [&](int cnt, const float4&, int, bool) {
      for (int m = 0; m < cnt; ++m) {
        const float d2 = d2_rule(tile->x[m], tile->y[m], tile->z[m], qx, qy, qz);
        if (d2 <= r2) {
          const unsigned u = __float_as_uint(d2);
          if ((u & mask) == prefix) hist[warp][(u >> shift) & (kSelBins - 1)][lane]++;
        }
      }
    });
    if (shift == a.first_shift) {  // first pass also yields k
      int k = 0;
      for (int b = 0; b < kSelBins; ++b) k += hist[warp][b][lane];
      if (k <= a.max_nn) need = false;
    }
    if (need) {
      int cum = 0, digit = kSelBins - 1;
      for (int b = 0; b < kSelBins; ++b) {
        const int c = hist[warp][b][lane];
        if (cum + c >= target) {
          digit = b;
          break;
        }
        cum += c;
      }
      target -= cum;
      prefix |= (unsigned)digit << shift;
      mask |= (unsigned)(kSelBins - 1) << shift;
    }
    if (!__any_sync(kFull, need)) break;
  }
  // tie-break on the input index among candidates with d2 == d2*: the target-th smallest
  const float dstar = __uint_as_float(prefix);
  int cur = -1;
  const int rounds = __reduce_max_sync(kFull, need ? target : 0);
  for (int t = 0; t < rounds; ++t) {
    int best = INT_MAX;
    for_each_chunk(g, pc, lane, tile, [&](int cnt, const float4&, int, bool) {
      for (int m = 0; m < cnt; ++m) {
        const float d2 = d2_rule(tile->x[m], tile->y[m], tile->z[m], qx, qy, qz);
        if (d2 == dstar && d2 <= r2) {
          const int id = g.perm[tile->idx[m]];
          if (id > cur && id < best) best = id;
        }
      }
    });
    if (need && t < target) cur = best;
  }
  if (pc.active && wanted) {
    a.thr_d2[pc.qi] = need ? dstar : INFINITY;
    a.thr_idx[pc.qi] = need ? cur : INT_MAX;
  }
}

constexpr int kListCap = kSelBins / 2;  // (d2, index) pairs that fit into a lane's histogram column

// Selection by value histograms (see the file comment).  hist[warp][.][lane] is the lane's private column: first
// the 64 counters, then the list of the target bin's candidates.  A second histogram level subdivides the
// target bin when it is too full for the list (k-NN queries: k << candidates in the radius).
__global__ void __launch_bounds__(kSelWarps * kWarp) threshold_kernel(const ThrArgs a) {
  const int warp = threadIdx.x >> 5;
  for (int pid = a.p0 + blockIdx.x * kSelWarps + warp; pid < a.p1; pid += gridDim.x * kSelWarps) {
    threshold_packet(a, pid);
    __syncwarp();
  }
}

__device__ __forceinline__ void value_select_packet(const ThrArgs& a, int pid) {
  __shared__ ChunkTile tiles[kSelWarps];
  __shared__ unsigned hist[kSelWarps][kSelBins + 1][kWarp];  // row kSelBins: everything beyond the radius
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (pid >= a.p1) return;
  if (a.only && !a.only[pid - a.p0]) return;
  const GridView& g = a.g;
  ChunkTile* tile = &tiles[warp];
  const PacketCtx pc = load_packet(g, pid, lane, a.r, tile);
  const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z, r2 = a.r2;
  bool wanted = pc.active;
  if (a.done) {
    wanted = wanted && !a.done[g.perm[pc.qi]];
    if (!__any_sync(kFull, wanted)) return;
  }
  // bins are monotone maps of d2, so they partition the candidates by distance whatever their exact edges are
  const float scale1 = (float)kSelBins / r2;
  auto bin1 = [&](float d2) { return min(__float2int_rz(d2 * scale1), kSelBins - 1); };
  const unsigned col = (unsigned)__cvta_generic_to_shared(&hist[warp][0][lane]);  // this lane's column, rows 128 bytes apart
  auto count = [&](int row) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(col + 128u * (unsigned)row) : "memory"); };
  auto pick = [&](int target, int& bin, int& before, int& in_bin) {  // bin holding the target-th candidate (1-based)
    int k = 0;
    bin = -1;
    before = in_bin = 0;
    for (int b = 0; b < kSelBins; ++b) {
      const int c = hist[warp][b][lane];
      if (bin < 0 && k + c >= target) {
        bin = b;
        before = k;
        in_bin = c;
      }
      k += c;
    }
    return k;
  };
  for (int b = 0; b <= kSelBins; ++b) hist[warp][b][lane] = 0;
  for_each_chunk(g, pc, lane, tile, [&](int cnt, const float4&, int, bool) {
    // no branch per candidate: whatever is not a neighbour is counted in the extra row (d2 > r2, also inf and NaN)
    for_each_staged_d2(tile, cnt, qx, qy, qz, [&](int, float d2) { count(d2 <= r2 ? bin1(d2) : kSelBins); });
  });
  int t1, before1, in1;
  const int k = pick(a.max_nn, t1, before1, in1);
  const bool need = wanted && k > a.max_nn;  // k <= max_nn: every neighbour is kept
  // second level: 64 sub-bins of bin t1
  const bool deep = need && in1 > kListCap;
  int t2 = -1, before2 = 0;
  const float lo2 = (float)t1 / scale1, scale2 = scale1 * (float)kSelBins;
  auto bin2 = [&](float d2) { return min(max(__float2int_rz((d2 - lo2) * scale2), 0), kSelBins - 1); };
  bool overflow = false;
  if (__any_sync(kFull, deep)) {
    for (int b = 0; b <= kSelBins; ++b) hist[warp][b][lane] = 0;
    for_each_chunk(g, pc, lane, tile, [&](int cnt, const float4&, int, bool) {
      if (!deep) return;
      for_each_staged_d2(tile, cnt, qx, qy, qz, [&](int, float d2) { count(d2 <= r2 && bin1(d2) == t1 ? bin2(d2) : kSelBins); });
    });
    if (deep) {
      int in2;
      pick(a.max_nn - before1, t2, before2, in2);
      overflow = in2 > kListCap;
    }
  }
  if (__any_sync(kFull, overflow)) {  // many equal distances: the radix select redoes this packet
    if (lane == 0) a.fallback[pid - a.p0] = 1;
    return;
  }
  if (__any_sync(kFull, need)) {
    float* ld2 = reinterpret_cast<float*>(&hist[warp][0][lane]);  // entry e: d2 at row 2e, index at row 2e + 1
    int filled = 0;
    for_each_chunk(g, pc, lane, tile, [&](int cnt, const float4&, int, bool) {
      if (!need) return;
      for_each_staged_d2(tile, cnt, qx, qy, qz, [&](int m, float d2) {
        if (d2 <= r2 && bin1(d2) == t1 && (!deep || bin2(d2) == t2) && filled < kListCap) {
          ld2[(2 * filled) * kWarp] = d2;
          reinterpret_cast<int*>(ld2)[(2 * filled + 1) * kWarp] = g.perm[tile->idx[m]];
          ++filled;
        }
      });
    });
    if (need) {
      // the remaining rank among the listed candidates, by (d2, index)
      const int want_rank = a.max_nn - before1 - before2 - 1;
      float sd2 = INFINITY;
      int sidx = INT_MAX;
      for (int e = 0; e < filled; ++e) {
        const float d2 = ld2[(2 * e) * kWarp];
        const int id = reinterpret_cast<const int*>(ld2)[(2 * e + 1) * kWarp];
        int rank = 0;
        for (int f = 0; f < filled; ++f) {
          const float od2 = ld2[(2 * f) * kWarp];
          const int oid = reinterpret_cast<const int*>(ld2)[(2 * f + 1) * kWarp];
          rank += (od2 < d2 || (od2 == d2 && oid < id)) ? 1 : 0;
        }
        if (rank == want_rank) {
          sd2 = d2;
          sidx = id;
        }
      }
      a.thr_d2[pc.qi] = sd2;
      a.thr_idx[pc.qi] = sidx;
    }
  }
  if (pc.active && wanted && !need) {
    a.thr_d2[pc.qi] = INFINITY;
    a.thr_idx[pc.qi] = INT_MAX;
  }
}

__global__ void __launch_bounds__(kSelWarps * kWarp) value_select_kernel(const ThrArgs a) {
  const int warp = threadIdx.x >> 5;
  for (int pid = a.p0 + blockIdx.x * kSelWarps + warp; pid < a.p1; pid += gridDim.x * kSelWarps) {
    value_select_packet(a, pid);
    __syncwarp();
  }
}

// First half of the truncated fast RSD pass (cab_rsd.cu, rsd_fast_kernel<., true>): the 64-bin histogram of d2 alone.
// Per query it leaves the bin that holds the max_nn-th neighbour and how many of that bin's candidates are still kept
// (code = bin | kept << 8; bin 255: the query has at most max_nn neighbours, all are kept).  The RSD traversal then
// accepts every candidate below the bin outright and settles the bin's few candidates from a short list, so the
// truncation costs one extra traversal instead of two plus a slower RSD kernel -- or none at all when the normals pass of
// the same call took the histogram along (cab_normals.cu, kHist).  Packets in which some query's bin holds more
// candidates than the list takes are flagged and go through the exact-threshold path.
__global__ void __launch_bounds__(kSelWarps * kWarp) nn_hist_kernel(const ThrArgs a) {
  __shared__ ChunkTile tiles[kSelWarps];
  __shared__ unsigned hist[kSelWarps][kSelBins + 1][kWarp];  // row kSelBins: everything beyond the radius
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GridView& g = a.g;
  ChunkTile* tile = &tiles[warp];
  const float r2 = a.r2;
  const float scale1 = (float)kSelBins / r2;
  const unsigned col = (unsigned)__cvta_generic_to_shared(&hist[warp][0][lane]);
  for (;;) {
    const int pid = a.p0 + next_packet(a.stats, lane);
    if (pid >= a.p1) break;
    const PacketCtx pc = load_packet(g, pid, lane, a.r, tile);
    const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z;
    for (int b = 0; b <= kSelBins; ++b) hist[warp][b][lane] = 0;
    for_each_chunk(g, pc, lane, tile, [&](int cnt, const float4&, int, bool) {
      for_each_staged_d2(tile, cnt, qx, qy, qz, [&](int, float d2) {
        const int row = d2 <= r2 ? min(__float2int_rz(d2 * scale1), kSelBins - 1) : kSelBins;
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(col + 128u * (unsigned)row) : "memory");
      });
    });
    __syncwarp();
    int k = 0, bin = 255, before = 0, in_bin = 0;
    for (int b = 0; b < kSelBins; ++b) {
      const int c = hist[warp][b][lane];
      if (bin == 255 && k + c >= a.max_nn + 1) {  // the (max_nn + 1)-th neighbour would be the first one dropped
        bin = b;
        before = k;
        in_bin = c;
      }
      k += c;
    }
    // bin: holds neighbour number max_nn + 1, so at least one of its candidates is dropped; kept = max_nn - before of them
    // (the RSD pass lists the bin's candidates by their 16-bit position in the candidate stream)
    const bool over = pc.active && bin != 255 && (in_bin > kTruncCap || pc.total > 65535);
    if (pc.active) a.code[pc.qi] = bin == 255 ? 255 : (bin | ((a.max_nn - before) << 8));
    if (__any_sync(kFull, over) && lane == 0) a.fallback[pid - a.p0] = 1;
    __syncwarp();
  }
}

__global__ void inverse_perm_kernel(const int* __restrict__ perm, int n, int* __restrict__ inv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = i;
}

struct DbgArgs {
  GridView g;
  const float* xyz;
  int stride;
  const int* domoff;
  int n_domains;
  int q0, q1;
  float r2;
  const float* thr_d2;  // sorted order, may be null
  const int* thr_idx;
  const int* inv_perm;
  const long long* offsets;  // null in the counting pass
  int* counts;
  int* idx;
  float* d2out;
};

__global__ void __launch_bounds__(128) neighbors_debug_kernel(const DbgArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = a.q0 + t;
  if (q >= a.q1) return;
  const float* p = a.xyz + (size_t)q * a.stride;
  const float qx = p[0], qy = p[1], qz = p[2];
  int count = 0;
  if (isfinite(qx) && isfinite(qy) && isfinite(qz)) {
    int d = 0;
    if (a.n_domains > 1) {
      int lo = 0, hi = a.n_domains;
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (a.domoff[mid] <= q) lo = mid; else hi = mid;
      }
      d = lo;
    }
    const Domain dm = a.g.domains[d];
    int cy, cz;
    row_cells(dm, qy, qz, a.g.inv_cell, cy, cz);
    const int cx = xfine_coord(qx, dm.ox, a.g.inv_cell, dm.nx, dm.xshift, dm.xwide) >> dm.xshift;
    const int cxlo = max(cx - 1, 0), cxhi = min(cx + 1, dm.nx - 1);
    float td2 = INFINITY;
    int tidx = INT_MAX;
    if (a.thr_d2) {
      const int s = a.inv_perm[q];
      td2 = a.thr_d2[s];
      tidx = a.thr_idx[s];
    }
    const long long off = a.offsets ? a.offsets[t] : 0;
    for (int z = cz - 1; z <= cz + 1; ++z)
      for (int y = cy - 1; y <= cy + 1; ++y) {
        if (!row_in_table(dm, y, z)) continue;
        const long long c = dm.cell_base + ((long long)z * dm.ny + y) * dm.nx;
        const int b = a.g.cell_start[c + cxlo], e = a.g.cell_start[c + cxhi + 1];
        for (int j = b; j < e; ++j) {
          const float4 cp = a.g.pos[j];
          const float d2 = d2_rule(cp.x, cp.y, cp.z, qx, qy, qz);
          if (d2 <= a.r2) {
            const int id = a.g.perm[j];
            if (d2 < td2 || (d2 == td2 && id <= tidx)) {
              if (a.offsets) {
                a.idx[off + count] = id;
                a.d2out[off + count] = d2;
              }
              ++count;
            }
          }
        }
      }
  }
  if (!a.offsets) a.counts[t] = count;
}

}  // namespace

// grid of the selection kernels: enough blocks to fill the GPU several times over, each walking the packets with a stride
static unsigned sel_grid(const cab_ctx* ctx, int np) {
  return (unsigned)std::max(1, std::min((np + kSelWarps - 1) / kSelWarps, ctx->sm_count * 32));
}

int run_thresholds(cab_ctx* ctx, float r, int max_nn, const unsigned char* done, bool halo) {
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  if (ctx->slab && !ctx->slab_info_valid) {  // the launch dimensions below need the slab's packet ranges on the host
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    if (int rc = finish_slab(ctx)) return rc;
  }
  if (int rc = reserve(ctx, ctx->b_thr_d2, (size_t)std::max(n, 1) * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_thr_idx, (size_t)std::max(n, 1) * sizeof(int))) return rc;
  ThrArgs a{};
  a.g = grid_view(ctx);
  packet_range(ctx, &a.p0, &a.p1);
  if (ctx->slab && halo) {
    a.p0 = ctx->slab_info.ph0;
    a.p1 = ctx->slab_info.ph1;
  }
  a.r = r;
  a.r2 = r * r;
  a.max_nn = max_nn;
  uint32_t bits;
  std::memcpy(&bits, &a.r2, 4);
  a.first_shift = (bits >> 30) ? 30 : 24;
  a.thr_d2 = (float*)ctx->b_thr_d2.p;
  a.thr_idx = (int*)ctx->b_thr_idx.p;
  a.done = done;
  const int np = a.p1 - a.p0;
  if (np > 0) {
    if (int rc = reserve(ctx, ctx->b_thr_flag, (size_t)np + 16)) return rc;
    a.fallback = (unsigned char*)ctx->b_thr_flag.p;
    CAB_CUDA(ctx, cudaMemsetAsync(a.fallback, 0, (size_t)np, st));
    value_select_kernel<<<sel_grid(ctx, np), kSelWarps * kWarp, 0, st>>>(a);
    CAB_LAUNCH_CHECK(ctx);
    a.use_fallback = 1;  // packets whose target bin did not fit the list (flag read on the device: no host sync)
    threshold_kernel<<<sel_grid(ctx, np), kSelWarps * kWarp, 0, st>>>(a);
    CAB_LAUNCH_CHECK(ctx);
  }
  return CAB_OK;
}

// Histogram half of the truncated fast RSD pass: per-query codes in b_thr_idx, flags of the packets that need the exact
// thresholds in b_thr_flag, and for those packets the thresholds themselves (b_thr_d2 / b_thr_idx hold either kind).
int run_nn_hist(cab_ctx* ctx, float r, int max_nn) {
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  if (ctx->slab && !ctx->slab_info_valid) {
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    if (int rc = finish_slab(ctx)) return rc;
  }
  if (int rc = reserve(ctx, ctx->b_thr_d2, (size_t)std::max(n, 1) * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_thr_idx, (size_t)std::max(n, 1) * sizeof(int))) return rc;
  if (int rc = reserve(ctx, ctx->b_stats, kStatBytes)) return rc;
  ThrArgs a{};
  a.g = grid_view(ctx);
  packet_range(ctx, &a.p0, &a.p1);
  a.r = r;
  a.r2 = r * r;
  a.max_nn = max_nn;
  uint32_t bits;
  std::memcpy(&bits, &a.r2, 4);
  a.first_shift = (bits >> 30) ? 30 : 24;
  a.thr_d2 = (float*)ctx->b_thr_d2.p;
  a.thr_idx = (int*)ctx->b_thr_idx.p;
  a.code = (int*)ctx->b_thr_idx.p;
  a.nrm = (const float4*)ctx->b_nrm.p;
  a.stats = (unsigned long long*)ctx->b_stats.p;
  const int np = a.p1 - a.p0;
  if (np <= 0) return CAB_OK;
  if (int rc = reserve(ctx, ctx->b_thr_flag, 2 * (size_t)np + 32)) return rc;
  unsigned char* flag_a = (unsigned char*)ctx->b_thr_flag.p;  // packets that leave the fast path
  unsigned char* flag_b = flag_a + np;                        // of those: packets that need the radix select
  // the normals pass of this call may have taken the histogram along (same grid, radius and max_nn): codes and flags are in place
  const bool have = ctx->trunc_hist_valid && !ctx->slab && ctx->kcount_r == r && ctx->trunc_hist_max_nn == max_nn;
  if (!have) {
    CAB_CUDA(ctx, cudaMemsetAsync(flag_a, 0, 2 * (size_t)np, st));
    CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_stats.p, 0, kStatBytes, st));
    a.fallback = flag_a;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nn_hist_kernel, kSelWarps * kWarp, 0);
    const unsigned grid = (unsigned)std::min<long long>((long long)std::max(per_sm, 1) * ctx->sm_count, (np + kSelWarps - 1) / kSelWarps);
    nn_hist_kernel<<<grid, kSelWarps * kWarp, 0, st>>>(a);
    CAB_LAUNCH_CHECK(ctx);
  }
  ctx->trunc_hist_valid = false;  // the thresholds of the flagged packets are about to overwrite their codes
  // exact (d2, index) thresholds for the flagged packets only
  a.only = flag_a;
  a.fallback = flag_b;
  value_select_kernel<<<sel_grid(ctx, np), kSelWarps * kWarp, 0, st>>>(a);
  CAB_LAUNCH_CHECK(ctx);
  a.use_fallback = 1;
  threshold_kernel<<<sel_grid(ctx, np), kSelWarps * kWarp, 0, st>>>(a);
  CAB_LAUNCH_CHECK(ctx);
  return CAB_OK;
}

int64_t run_neighbors_debug(cab_ctx* ctx, float r, int max_nn, int64_t q0, int64_t q1, int64_t* offsets, int32_t* idx,
                            float* d2, int64_t cap) {
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_neighbors_debug: build the grid first");
  if (ctx->slab) return fail(ctx, CAB_ERR_STATE, "cab_neighbors_debug: needs the whole grid, this context holds one shard's slab");
  if (q0 < 0 || q1 < q0 || q1 > ctx->n) return fail(ctx, CAB_ERR_ARG, "cab_neighbors_debug: bad query range");
  if (!offsets) return fail(ctx, CAB_ERR_ARG, "cab_neighbors_debug: offsets is NULL");
  if (!(r > 0.f) || r > ctx->cell * 1.0000001f)
    return fail(ctx, CAB_ERR_ARG, "cab_neighbors_debug: radius %g exceeds the grid cell %g", (double)r, (double)ctx->cell);
  const int nq = (int)(q1 - q0);
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  offsets[0] = 0;
  if (nq == 0) return 0;
  DbgArgs a{};
  a.g = grid_view(ctx);
  a.xyz = ctx->xyz_in;
  a.stride = ctx->stride;
  a.domoff = (const int*)ctx->b_domoff.p;
  a.n_domains = ctx->n_domains;
  a.q0 = (int)q0;
  a.q1 = (int)q1;
  a.r2 = r * r;
  DevBuf& scratch = ctx->b_out4;  // reuse: counts (nq ints) + offsets (nq+1 int64)
  const size_t off_bytes = (size_t)(nq + 1) * 8, cnt_bytes = (size_t)nq * 4;
  if (int rc = reserve(ctx, scratch, off_bytes + cnt_bytes + 64)) return rc;
  long long* d_off = (long long*)scratch.p;
  int* d_cnt = (int*)((char*)scratch.p + off_bytes);
  if (max_nn > 0) {
    if (int rc = run_thresholds(ctx, r, max_nn)) return rc;
    if (int rc2 = reserve(ctx, ctx->b_out1a, (size_t)n * 4)) return rc2;
    inverse_perm_kernel<<<(n + 255) / 256, 256, 0, st>>>((const int*)ctx->b_perm.p, n, (int*)ctx->b_out1a.p);
    CAB_LAUNCH_CHECK(ctx);
    a.thr_d2 = (const float*)ctx->b_thr_d2.p;
    a.thr_idx = (const int*)ctx->b_thr_idx.p;
    a.inv_perm = (const int*)ctx->b_out1a.p;
  }
  a.counts = d_cnt;
  neighbors_debug_kernel<<<(nq + 127) / 128, 128, 0, st>>>(a);
  CAB_LAUNCH_CHECK(ctx);
  std::vector<int> counts(nq);
  CAB_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_cnt, cnt_bytes, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  for (int i = 0; i < nq; ++i) offsets[i + 1] = offsets[i] + counts[i];
  const int64_t total = offsets[nq];
  if (!idx || !d2 || total > cap) return total;
  if (total == 0) return 0;
  if (int rc = reserve(ctx, ctx->b_out1b, (size_t)total * 8)) return rc;
  CAB_CUDA(ctx, cudaMemcpyAsync(d_off, offsets, off_bytes, cudaMemcpyHostToDevice, st));
  a.offsets = d_off;
  a.idx = (int*)ctx->b_out1b.p;
  a.d2out = (float*)ctx->b_out1b.p + total;
  neighbors_debug_kernel<<<(nq + 127) / 128, 128, 0, st>>>(a);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cudaMemcpyAsync(idx, a.idx, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaMemcpyAsync(d2, a.d2out, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  return total;
}

}  // namespace cab
