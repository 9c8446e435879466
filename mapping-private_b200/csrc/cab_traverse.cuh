// cab_traverse.cuh -- the shared neighbour traversal of the normals / RSD / threshold kernels.
//
// One warp owns one packet (<= 32 consecutive sorted queries of one row, one query per lane).
// The candidate set of the packet is 9 contiguous runs of the sorted array: rows (cy+dy, cz+dz),
// cells [cx(xmin)-1, cx(xmax)+1], trimmed to the packet's x window and concatenated into one stream
// of 32-candidate chunks: each lane loads one candidate with a 128-bit load, the chunk is staged in
// the warp's shared-memory tile (SoA: x[32], y[32], z[32]) and every lane tests all staged
// candidates against its own query.
//
// The distance test uses Blackwell's packed fp32x2 pipe (FADD2 / FFMA2, PTX *.f32x2, sm_100+):
// two candidates per instruction, each half rounded exactly like the scalar epsilon rule
//   d2 = (dx*dx + dy*dy) + dz*dz   (fp32, round-to-nearest, no contraction).
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false, so the squares are
// written as fma(d, d, +0), which rounds exactly like a multiply and cannot be fused further.
#pragma once
#include "cab_internal.cuh"

namespace cab {

constexpr int kWarpsPerBlock = 8;

typedef unsigned long long f32x2;  // two packed floats in a 64-bit register

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// exact product (one rounding): fma(a, b, +0) rounds like a multiply and cannot be contracted with a following add
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { return fma2(a, b, 0ull); }
// exact product (one rounding), not contractible by ptxas
__device__ __forceinline__ f32x2 sq2(f32x2 a) { return fma2(a, a, 0ull); }

// Persistent warps pull packets from a global counter (stats[2*kStatSlots + slot]; two launches that run side by side
// use one slot each): blocks never idle on
// their slowest warp and concurrently running warps work on neighbouring packets (L2 locality).
__device__ __forceinline__ int next_packet(unsigned long long* stats, int lane, int slot = 0) {
  int v = 0;
  if (lane == 0) v = (int)atomicAdd(reinterpret_cast<unsigned int*>(stats + 2 * kStatSlots + slot), 1u);
  return __shfl_sync(kFull, v, 0);
}

// Per-warp shared-memory tile: one staged chunk (SoA) + the packet's run table.
struct alignas(16) ChunkTile {
  float x[kWarp], y[kWarp], z[kWarp];
  int idx[kWarp];      // sorted index of the staged candidate (-1: padding)
  int run_begin[12];   // trimmed candidate runs of the packet: first sorted index
  int run_cum[12];     // exclusive prefix of the run lengths; run_cum[9] = total candidates
};

struct PacketCtx {
  int start, count;   // packet
  int qi;             // this lane's sorted query index (clamped to a valid one)
  bool active;        // lane < count
  float4 q;           // this lane's query position
  int total;          // candidates of the packet (all runs, trimmed to the x window)
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// Loads the packet and builds its candidate stream: the 9 runs (rows (cy+dy, cz+dz), cells
// [cx(xmin)-1, cx(xmax)+1]) are trimmed to the packet's x window [xmin - r, xmax + r] by a binary
// search over the fine x coordinate the rows are sorted by (lanes 0..8 search the lower end, lanes
// 16..24 the upper end), then concatenated so that every 32-candidate chunk but the last is full.
__device__ __forceinline__ PacketCtx load_packet(const GridView& g, int pid, int lane, float r, ChunkTile* tile) {
  PacketCtx pc;
  const Packet pk = g.packets[pid];
  const Domain dm = g.domains[pk.domain];
  pc.start = pk.start;
  pc.count = pk.count;
  pc.active = lane < pk.count;
  pc.qi = pk.start + min(lane, pk.count - 1);
  pc.q = g.pos[pc.qi];
  const float xmin = warp_min(pc.q.x), xmax = warp_max(pc.q.x);
  const float rc = r * 1.00001f;
  const int cy = pk.row_local % dm.ny, cz = pk.row_local / dm.ny;
  const int cxlo = max((xfine_coord(xmin, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide) >> dm.xshift) - 1, 0);
  const int cxhi = min((xfine_coord(xmax, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide) >> dm.xshift) + 1, dm.nx - 1);
  // fine-x window: x >= xmin - rc  =>  xf(x) >= xf(xmin - rc) (xf is monotone), same at the top
  const int xf_lo = xfine_coord(xmin - rc, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide);
  const int xf_hi = xfine_coord(xmax + rc, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide);
  const int t = lane & 15;  // run handled by this lane (lanes 0..8 lower bound, 16..24 upper bound)
  int lo = 0, hi = 0;
  if (t < 9) {
    const int y = cy + t % 3 - 1, z = cz + t / 3 - 1;
    if (row_in_table(dm, y, z)) {
      const long long c = dm.cell_base + ((long long)z * dm.ny + y) * dm.nx;
      lo = g.cell_start[c + cxlo];
      hi = g.cell_start[c + cxhi + 1];
    }
  }
  // first index in [lo, hi) whose xf is >= key (lower half-warp) / > key (upper half-warp)
  const int key = lane < 16 ? xf_lo : xf_hi + 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int xf = xfine_coord(g.pos[mid].x, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide);
    if (xf < key) lo = mid + 1; else hi = mid;
  }
  const int end = __shfl_sync(kFull, lo, (lane + 16) & 31);
  int len = (lane < 9) ? max(end - lo, 0) : 0;
  int cum = len;  // inclusive scan over lanes
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    const int v = __shfl_up_sync(kFull, cum, o);
    if (lane >= o) cum += v;
  }
  __syncwarp();
  if (lane < 9) {
    tile->run_begin[lane] = lo;
    tile->run_cum[lane] = cum - len;
  }
  pc.total = __shfl_sync(kFull, cum, 8);
  if (lane == 9) tile->run_cum[9] = pc.total;
  __syncwarp();
  return pc;
}

// Streams the packet's candidates in 32-candidate chunks.  Chunk c holds the stream positions
// {lane * nchunks + c}: a strided, i.e. spatially stratified, sample of the whole candidate region, so
// every chunk gives every query about the same number of hits (consecutive positions would make the
// hit counts of a chunk very uneven across lanes, which the hit-compacting kernels pay for with
// idle lanes).  Neighbouring chunks touch the same cache lines, so the loads stay L1 hits.
// Each lane loads one candidate, the chunk is staged into `tile`, then body(cnt, c, j, valid) runs
// with the tile visible to the whole warp: c / j are this lane's own candidate and its sorted index
// (a far-away sentinel / -1 for the padding lanes, which are always the highest lanes).
// Returns the candidates tested per query.
template <class Body>
__device__ __forceinline__ int for_each_chunk(const GridView& g, const PacketCtx& pc, int lane, ChunkTile* tile,
                                              Body&& body) {
  const int nchunks = (pc.total + kWarp - 1) / kWarp;
#pragma unroll 1
  for (int c0 = 0; c0 < nchunks; ++c0) {
    const int p = lane * nchunks + c0;
    const bool valid = p < pc.total;
    // run containing stream position p: largest t with run_cum[t] <= p (9 runs, branch-free search)
    int t = (p >= tile->run_cum[4]) ? 4 : 0;
    t += (p >= tile->run_cum[t + 2]) ? 2 : 0;
    t += (p >= tile->run_cum[t + 1]) ? 1 : 0;
    t += (t == 7 && p >= tile->run_cum[8]) ? 1 : 0;
    const int j = valid ? tile->run_begin[t] + (p - tile->run_cum[t]) : -1;
    const float4 c = valid ? g.pos[j] : make_float4(3.0e30f, 3.0e30f, 3.0e30f, 0.f);
    __syncwarp();
    tile->x[lane] = c.x;
    tile->y[lane] = c.y;
    tile->z[lane] = c.z;
    tile->idx[lane] = j;
    __syncwarp();
    body(__popc(__ballot_sync(kFull, valid)), c, j, valid);
  }
  return pc.total;
}

// d2 (the documented rule, packed arithmetic) of every staged candidate against this lane's query: fn(m, d2) for
// m = 0 .. 4 * ceil(cnt / 4) - 1; the padding candidates lie far away (d2 = inf).
template <class Fn>
__device__ __forceinline__ void for_each_staged_d2(const ChunkTile* tile, int cnt, float qx, float qy, float qz, Fn&& fn) {
  const f32x2 qx2 = pack2(qx, qx), qy2 = pack2(qy, qy), qz2 = pack2(qz, qz);
  const float4* tx = reinterpret_cast<const float4*>(tile->x);
  const float4* ty = reinterpret_cast<const float4*>(tile->y);
  const float4* tz = reinterpret_cast<const float4*>(tile->z);
  const int groups = (cnt + 3) >> 2;
#pragma unroll 2
  for (int g4 = 0; g4 < groups; ++g4) {
    const float4 X = tx[g4], Y = ty[g4], Z = tz[g4];
    float a, b;
    {
      const f32x2 dx = sub2(pack2(X.x, X.y), qx2), dy = sub2(pack2(Y.x, Y.y), qy2), dz = sub2(pack2(Z.x, Z.y), qz2);
      unpack2(add2(add2(sq2(dx), sq2(dy)), sq2(dz)), a, b);
      fn(4 * g4, a);
      fn(4 * g4 + 1, b);
    }
    {
      const f32x2 dx = sub2(pack2(X.z, X.w), qx2), dy = sub2(pack2(Y.z, Y.w), qy2), dz = sub2(pack2(Z.z, Z.w), qz2);
      unpack2(add2(add2(sq2(dx), sq2(dy)), sq2(dz)), a, b);
      fn(4 * g4 + 2, a);
      fn(4 * g4 + 3, b);
    }
  }
}

// Hit mask of one staged chunk for this lane's query: bit m set iff d2(candidate m, q) <= r2.
// Packed fp32x2 arithmetic, 4 candidates per shared-memory broadcast load.
__device__ __forceinline__ unsigned chunk_hit_mask(const ChunkTile* tile, float qx, float qy, float qz, float r2) {
  const f32x2 qx2 = pack2(qx, qx), qy2 = pack2(qy, qy), qz2 = pack2(qz, qz);
  unsigned mask = 0;
  const float4* tx = reinterpret_cast<const float4*>(tile->x);
  const float4* ty = reinterpret_cast<const float4*>(tile->y);
  const float4* tz = reinterpret_cast<const float4*>(tile->z);
#pragma unroll
  for (int g4 = 0; g4 < kWarp / 4; ++g4) {
    const float4 X = tx[g4], Y = ty[g4], Z = tz[g4];
    {
      const f32x2 dx = sub2(pack2(X.x, X.y), qx2), dy = sub2(pack2(Y.x, Y.y), qy2), dz = sub2(pack2(Z.x, Z.y), qz2);
      const f32x2 d2 = add2(add2(sq2(dx), sq2(dy)), sq2(dz));
      float a, b;
      unpack2(d2, a, b);
      mask |= (a <= r2 ? 1u : 0u) << (4 * g4);
      mask |= (b <= r2 ? 1u : 0u) << (4 * g4 + 1);
    }
    {
      const f32x2 dx = sub2(pack2(X.z, X.w), qx2), dy = sub2(pack2(Y.z, Y.w), qy2), dz = sub2(pack2(Z.z, Z.w), qz2);
      const f32x2 d2 = add2(add2(sq2(dx), sq2(dy)), sq2(dz));
      float a, b;
      unpack2(d2, a, b);
      mask |= (a <= r2 ? 1u : 0u) << (4 * g4 + 2);
      mask |= (b <= r2 ? 1u : 0u) << (4 * g4 + 3);
    }
  }
  return mask;
}

}  // namespace cab
