// cab_traverse.cuh -- the shared neighbour traversal of the normals / RSD / threshold kernels.
//
// One warp owns one packet (<= 32 consecutive sorted queries of one row, one query per lane).
// The candidate set of the packet is 9 contiguous runs of the sorted array: rows (cy+dy, cz+dz),
// cells [cx(xmin)-1, cx(xmax)+1].  Runs are streamed in 32-point chunks: each lane loads one
// candidate with a coalesced 128-bit load, the chunk is staged in the warp's shared-memory tile
// (SoA: x[32], y[32], z[32]) and every lane tests all staged candidates against its own query.
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
// exact product (one rounding), not contractible by ptxas
__device__ __forceinline__ f32x2 sq2(f32x2 a) { return fma2(a, a, 0ull); }

// Persistent warps pull packets from a global counter (stats[2*kStatSlots]): blocks never idle on
// their slowest warp and concurrently running warps work on neighbouring packets (L2 locality).
__device__ __forceinline__ int next_packet(unsigned long long* stats, int lane) {
  int v = 0;
  if (lane == 0) v = (int)atomicAdd(reinterpret_cast<unsigned int*>(stats + 2 * kStatSlots), 1u);
  return __shfl_sync(kFull, v, 0);
}

struct PacketCtx {
  int start, count;   // packet
  int qi;             // this lane's sorted query index (clamped to a valid one)
  bool active;        // lane < count
  float4 q;           // this lane's query position
  float xlo, xhi;     // x window for chunk culling
  int rb, re;         // lane t < 9: candidate run t = [rb, re)
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

__device__ __forceinline__ PacketCtx load_packet(const GridView& g, int pid, int lane, float r) {
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
  pc.xlo = xmin - rc;
  pc.xhi = xmax + rc;
  const int cy = pk.row_local % dm.ny, cz = pk.row_local / dm.ny;
  const int cxlo = max((xfine_coord(xmin, dm.ox, g.inv_cell, dm.nx, dm.xshift) >> dm.xshift) - 1, 0);
  const int cxhi = min((xfine_coord(xmax, dm.ox, g.inv_cell, dm.nx, dm.xshift) >> dm.xshift) + 1, dm.nx - 1);
  pc.rb = pc.re = 0;
  if (lane < 9) {
    const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
    if (y >= 0 && y < dm.ny && z >= 0 && z < dm.nz) {
      const long long c = dm.cell_base + ((long long)z * dm.ny + y) * dm.nx;
      pc.rb = g.cell_start[c + cxlo];
      pc.re = g.cell_start[c + cxhi + 1];
    }
  }
  return pc;
}

// Per-warp SoA tile of one staged chunk.
struct alignas(16) ChunkTile {
  float x[kWarp], y[kWarp], z[kWarp];
};

// Streams every candidate chunk of the packet.  For each chunk that intersects the packet's x
// window: stage(j, valid) lets the caller load per-candidate payload (lane j's candidate), the
// positions are staged into `tile`, then body(base, cnt, c) runs with the tile visible to the whole
// warp; c is this lane's own candidate (sentinel when !valid).  Slots >= cnt hold a far-away
// sentinel that can never pass the distance test.  Returns the candidates tested per query.
template <class Body>
__device__ __forceinline__ int for_each_chunk(const GridView& g, const PacketCtx& pc, int lane, ChunkTile* tile,
                                              Body&& body) {
  int tested = 0;
#pragma unroll 1
  for (int t = 0; t < 9; ++t) {
    const int b = __shfl_sync(kFull, pc.rb, t), e = __shfl_sync(kFull, pc.re, t);
#pragma unroll 1
    for (int base = b; base < e; base += kWarp) {
      const int j = base + lane;
      const bool valid = j < e;
      const float4 c = valid ? g.pos[j] : make_float4(3.0e30f, 3.0e30f, 3.0e30f, 0.f);
      // runs are sorted by x: skip chunks entirely outside the packet's x window
      if (__all_sync(kFull, !valid || c.x < pc.xlo || c.x > pc.xhi)) continue;
      __syncwarp();
      tile->x[lane] = c.x;
      tile->y[lane] = c.y;
      tile->z[lane] = c.z;
      __syncwarp();
      const int cnt = min(kWarp, e - base);
      body(base, cnt, c, valid);
      tested += cnt;
    }
  }
  return tested;
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
