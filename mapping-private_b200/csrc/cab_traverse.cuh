// cab_traverse.cuh -- the shared neighbour traversal of the normals / RSD / debug kernels.
//
// One warp owns one packet (<= 32 consecutive sorted queries of one row, one query per lane).
// The candidate set of the packet is 9 contiguous runs of the sorted array: rows (cy+dy, cz+dz),
// cells [cx(xmin)-1, cx(xmax)+1].  Runs are streamed in 32-point chunks: each lane loads one
// candidate with a coalesced 128-bit load, the chunk is staged in the warp's shared-memory tile
// and every lane tests all staged candidates against its own query (broadcast LDS.128).
#pragma once
#include "cab_internal.cuh"

namespace cab {

constexpr int kWarpsPerBlock = 8;

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

// Streams every candidate chunk of the packet.  `stage(j, valid)` is called by every lane with
// the sorted index it should load (j < run end iff valid) and must store the candidate into the
// warp's tile(s); `body(base, cnt)` then runs with the tile visible to the whole warp.
// Returns the number of candidates tested per query.
template <class Stage, class Body>
__device__ __forceinline__ int for_each_chunk(const GridView& g, const PacketCtx& pc, int lane, Stage&& stage,
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
      stage(j, valid, c);
      __syncwarp();
      const int cnt = min(kWarp, e - base);
      body(base, cnt);
      tested += cnt;
    }
  }
  return tested;
}

}  // namespace cab
