// Device code of F3 marching cubes (see mcubes.cu for the pipeline).  Kept free of host API calls and of CUDA runtime
// types so that tests/emu can run these kernels unmodified on the CPU emulation layer of the build container.
#pragma once
#include <stdint.h>
#include "mp_scan.cuh"
#include "mc_table.inc"

namespace mcubes {


// per node byte: bits 0..2 = owned +x/+y/+z edge active; bits 3..5 = triangle count of the cell whose corner 0 is
// this node (0 when the node is on the +face of the grid).
// Block (32, 8) = eight rows of one z plane, one warp per row; grid (D, ceil(H/8)).  A warp walks its row in groups of
// kClassGroup chunks of 32 nodes: every node's value is loaded ONCE per neighbouring row (4 coalesced loads per chunk,
// all loads of a group in flight together), turned into an inside/outside ballot, and the x+1 neighbour of a node is
// the next bit of the same ballot (bit 0 of the following chunk's for lane 31) -- no second load, no divisions.
constexpr int kClassRows = 8;
constexpr int kClassGroup = 4;
__global__ void __launch_bounds__(32 * kClassRows)
classify_kernel(const float* __restrict__ vol, uint8_t* __restrict__ code, uint8_t* __restrict__ cases, int D, int H,
                int W, float iso) {
  const int z = blockIdx.x, y = blockIdx.y * kClassRows + threadIdx.y;
  if (y >= H) return;                    // (a whole warp: blockDim.x == 32)
  const int lane = threadIdx.x;
  const bool yi = y + 1 < H, zi = z + 1 < D;
  const size_t row = ((size_t)z * H + y) * W;
  // neighbour rows; a missing neighbour aliases the row itself (its bits are masked by yi / zi below)
  const float* r[4];
  r[0] = vol + row;
  r[1] = r[0] + (yi ? (size_t)W : 0);
  r[2] = r[0] + (zi ? (size_t)H * W : 0);
  r[3] = r[2] + (yi ? (size_t)W : 0);
  // inside/outside ballots of chunk `xs` (lanes beyond the row re-read its last node: in bounds, never consumed)
  unsigned m[kClassGroup + 1][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) m[0][q] = __ballot_sync(0xffffffffu, __ldg(r[q] + min(lane, W - 1)) > iso);
  for (int g0 = 0; g0 < W; g0 += 32 * kClassGroup) {
    float v[kClassGroup][4];
#pragma unroll
    for (int c = 0; c < kClassGroup; ++c) {
      const int xs = g0 + 32 * (c + 1);                       // warp-uniform
      const int xl = min(xs + lane, W - 1);
#pragma unroll
      for (int q = 0; q < 4; ++q) v[c][q] = (xs < W) ? __ldg(r[q] + xl) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < kClassGroup; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) m[c + 1][q] = __ballot_sync(0xffffffffu, v[c][q] > iso);
#pragma unroll
    for (int c = 0; c < kClassGroup; ++c) {
      const int x = g0 + 32 * c + lane;
      if (x < W) {
        const bool xi = x + 1 < W;
        // corner (dz, dy, dx): row q = 2*dz + dy, bit lane + dx of that row's ballot
        auto at = [&](int q, int dx) -> bool {
          const int l = lane + dx;
          return ((l < 32 ? (m[c][q] >> l) : m[c + 1][q]) & 1u) != 0u;
        };
        const bool b000 = at(0, 0), b001 = at(0, 1), b010 = at(1, 0), b011 = at(1, 1);
        const bool b100 = at(2, 0), b101 = at(2, 1), b110 = at(3, 0), b111 = at(3, 1);
        uint8_t cd = 0;
        if (xi && (b001 != b000)) cd |= 1;
        if (yi && (b010 != b000)) cd |= 2;
        if (zi && (b100 != b000)) cd |= 4;
        uint8_t cs = 0;
        if (xi && yi && zi) {
          int k = b000 ? 1 : 0;
          k |= b001 ? 2 : 0;
          k |= b010 ? 4 : 0;
          k |= b011 ? 8 : 0;
          k |= b100 ? 16 : 0;
          k |= b101 ? 32 : 0;
          k |= b110 ? 64 : 0;
          k |= b111 ? 128 : 0;
          cs = (uint8_t)k;
          cd |= (uint8_t)(c_mc_ntri[k] << 3);
        }
        code[row + x] = cd;
        cases[row + x] = cs;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) m[0][q] = m[kClassGroup][q];
  }
}

struct CountF {    // low 32: vertices owned by node i, high 32: triangles of cell i
  const uint8_t* code;       // cudaMalloc'ed (8-byte aligned)
  static constexpr bool kVec8 = true;
  static __device__ __forceinline__ unsigned long long counts(uint32_t c) {
    return (unsigned long long)__popc(c & 7u) | ((unsigned long long)(c >> 3) << 32);
  }
  __device__ unsigned long long operator()(long long i) const { return counts(code[i]); }
  __device__ void load8(long long i, unsigned long long (&v)[8]) const {
    uint32_t b[8];
    mpscan::load_bytes8(code, i, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = counts(b[j]);
  }
};

// Stores the exclusive vertex offset of the nodes that own at least one vertex.  Faces look offsets up by node, but only
// for nodes owning the active edge in question, so the other entries of the dense array are never read (and never
// written: 80 k stores instead of 17 M at 257^3).
struct OffsetsEmit {
  uint32_t* voff;
  __device__ void operator()(long long i, unsigned long long v, unsigned long long pre) const {
    if ((uint32_t)v) voff[i] = (uint32_t)pre;
  }
};

// Emission.  Phase 1 = the emit half of the ordered scan (block offsets come from the count pass): every active node
// (owning a vertex or a triangle) is queued in shared memory with its exclusive (vertex, triangle) offsets; the order
// inside the queue is irrelevant because every entry carries its own output positions.  Phase 2 spreads the queue's
// work items -- 3 candidate vertices + MC_MAX_TRI * 3 face corners per entry -- over the CTA.
constexpr int kItemsPerNode = 3 + 3 * MC_MAX_TRI;

__global__ void __launch_bounds__(mpscan::kThreads)
mesh_emit_kernel(const float* __restrict__ vol, const uint8_t* __restrict__ code, const uint8_t* __restrict__ cases,
                 const uint32_t* __restrict__ voff, const unsigned long long* __restrict__ block_offsets,
                 float* __restrict__ verts, int32_t* __restrict__ faces, int H, int W, long long n, float iso) {
  __shared__ int q_node[mpscan::kChunk];
  __shared__ uint32_t q_voff[mpscan::kChunk];
  __shared__ uint32_t q_foff[mpscan::kChunk];
  __shared__ int q_n;
  if (threadIdx.x == 0) q_n = 0;        // (published by the barriers inside block_excl_scan)
  const long long base = (long long)blockIdx.x * mpscan::kChunk + (long long)threadIdx.x * mpscan::kItems;
  unsigned long long v[mpscan::kItems];
  const CountF f{code};
  mpscan::load_items(f, base, n, v);
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < mpscan::kItems; ++j) s += v[j];
  unsigned long long run = block_offsets[blockIdx.x] + mpscan::block_excl_scan(s, nullptr);
#pragma unroll
  for (int j = 0; j < mpscan::kItems; ++j) {
    if (v[j]) {
      const int slot = atomicAdd(&q_n, 1);
      q_node[slot] = (int)(base + j);
      q_voff[slot] = (uint32_t)run;
      q_foff[slot] = (uint32_t)(run >> 32);
    }
    run += v[j];
  }
  __syncthreads();
  const int items = q_n * kItemsPerNode;
  const int plane = H * W;
  for (int w = threadIdx.x; w < items; w += mpscan::kThreads) {
    const int e = w / kItemsPerNode, sub = w - e * kItemsPerNode;
    const int i = q_node[e];
    const uint32_t c = __ldg(code + i);
    if (sub < 3) {
      // vertex on the owned edge along axis `sub`
      if (!((c >> sub) & 1u)) continue;
      const uint32_t vi = q_voff[e] + __popc(c & ((1u << sub) - 1u));
      const int z = i / plane, r = i - z * plane, y = r / W, x = r - y * W;
      const int step = sub == 0 ? 1 : (sub == 1 ? W : plane);
      const float va = __ldg(vol + i), vb = __ldg(vol + i + step);
      const float t = __fdiv_rn(__fsub_rn(iso, va), __fsub_rn(vb, va));
      float p[3] = {(float)x, (float)y, (float)z};
      p[sub] = __fadd_rn(p[sub], t);
      float* o = verts + 3ll * vi;
      o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    } else {
      const int corner = sub - 3, t = corner / 3;
      if (t >= (int)(c >> 3)) continue;
      const int k = __ldg(cases + i);
      const int ed = g_mc_tri[k][corner];
      // edge -> owning node + axis.  edges 0-3 along x at (y,z) offsets, 4-7 along y at (x,z), 8-11 along z at (x,y)
      const int axis = ed >> 2, q = ed & 3;
      int ox = 0, oy = 0, oz = 0;
      if (axis == 0) { oy = q & 1; oz = q >> 1; }
      else if (axis == 1) { ox = q & 1; oz = q >> 1; }
      else { ox = q & 1; oy = q >> 1; }
      const int node = i + (oz * H + oy) * W + ox;
      const uint32_t cn = __ldg(code + node) & 7u;
      const uint32_t rank = __popc(cn & ((1u << axis) - 1u));
      faces[3ll * (q_foff[e] + t) + (corner - 3 * t)] = (int32_t)(__ldg(voff + node) + rank);
    }
  }
}

}  // namespace mcubes
