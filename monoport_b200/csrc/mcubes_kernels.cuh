// Device code of F3 marching cubes (see mcubes.cu for the pipeline).  Kept free of host API calls and of CUDA runtime
// types so that tests/emu can run these kernels unmodified on the CPU emulation layer of the build container.
#pragma once
#include <stdint.h>
#include "mp_scan.cuh"
#include "mc_table.inc"

namespace mcubes {


// per node byte: bits 0..2 = owned +x/+y/+z edge active; bits 3..5 = triangle count of the cell whose corner 0 is
// this node (0 when the node is on the +face of the grid).
// Block (32, 8) = eight rows of one z plane, a warp walks its row 32 nodes at a time (coalesced); grid (D, ceil(H/8)).
// Eight loads per node, four of which hit the lines the neighbouring lane just fetched.  (A variant that loads every
// value once and derives the x+1 neighbours from warp ballots measured SLOWER on B200 -- 100 us instead of 70 us at
// 257^3: the kernel is instruction-bound, not load-bound; profiles/r01_final3_recon_trace_ballot_classify.txt.)
constexpr int kClassRows = 8;
__global__ void __launch_bounds__(32 * kClassRows)
classify_kernel(const float* __restrict__ vol, uint8_t* __restrict__ code, uint8_t* __restrict__ cases, int D, int H,
                int W, float iso) {
  const int z = blockIdx.x, y = blockIdx.y * kClassRows + threadIdx.y;
  if (y >= H) return;
  const bool yi = y + 1 < H, zi = z + 1 < D;
  const size_t row = ((size_t)z * H + y) * W;
  // neighbour rows; a missing neighbour aliases the row itself (its bits are masked by yi / zi below)
  const float* r00 = vol + row;
  const float* r01 = r00 + (yi ? (size_t)W : 0);
  const float* r10 = r00 + (zi ? (size_t)H * W : 0);
  const float* r11 = r10 + (yi ? (size_t)W : 0);
#pragma unroll 2
  for (int x = threadIdx.x; x < W; x += 32) {
    const bool xi = x + 1 < W;
    const int x1 = xi ? x + 1 : x;
    const bool b000 = __ldg(r00 + x) > iso, b001 = __ldg(r00 + x1) > iso;
    const bool b010 = __ldg(r01 + x) > iso, b011 = __ldg(r01 + x1) > iso;
    const bool b100 = __ldg(r10 + x) > iso, b101 = __ldg(r10 + x1) > iso;
    const bool b110 = __ldg(r11 + x) > iso, b111 = __ldg(r11 + x1) > iso;
    uint8_t c = 0;
    if (xi && (b001 != b000)) c |= 1;
    if (yi && (b010 != b000)) c |= 2;
    if (zi && (b100 != b000)) c |= 4;
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
      c |= (uint8_t)(c_mc_ntri[k] << 3);
    }
    code[row + x] = c;
    cases[row + x] = cs;
  }
}

// Opt-in variant (MONOPORT_B200_MC_FAST=1, to be timed on a B200): ~98 % of the 32-node chunks of an occupancy volume lie
// entirely inside or outside the surface.  A warp first loads every node's own column of the four neighbouring rows (4
// loads per node, plus column x0+32 by lane 0) and votes; a uniform chunk stores 32 zero code bytes and is done (its case
// bytes are never read: no triangles).  Only mixed chunks take the full eight-corner path of classify_kernel.
__global__ void __launch_bounds__(32 * kClassRows)
classify_fast_kernel(const float* __restrict__ vol, uint8_t* __restrict__ code, uint8_t* __restrict__ cases, int D, int H,
                     int W, float iso) {
  const int z = blockIdx.x, y = blockIdx.y * kClassRows + threadIdx.y;
  if (y >= H) return;                    // (a whole warp: blockDim.x == 32)
  const int lane = threadIdx.x;
  const bool yi = y + 1 < H, zi = z + 1 < D;
  const size_t row = ((size_t)z * H + y) * W;
  const float* r00 = vol + row;
  const float* r01 = r00 + (yi ? (size_t)W : 0);
  const float* r10 = r00 + (zi ? (size_t)H * W : 0);
  const float* r11 = r10 + (yi ? (size_t)W : 0);
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const int xc = min(x, W - 1);                        // lanes beyond the row repeat its last node
    const int xe = min(x0 + 32, W - 1);                  // the column after the chunk (lane 0 adds it to the vote)
    bool any_in = false, all_in = true;
    {
      const bool a = __ldg(r00 + xc) > iso, b = __ldg(r01 + xc) > iso, c = __ldg(r10 + xc) > iso, d = __ldg(r11 + xc) > iso;
      any_in = a | b | c | d;
      all_in = a & b & c & d;
      if (lane == 0) {
        const bool e = __ldg(r00 + xe) > iso, f = __ldg(r01 + xe) > iso, g = __ldg(r10 + xe) > iso, h = __ldg(r11 + xe) > iso;
        any_in |= e | f | g | h;
        all_in &= e & f & g & h;
      }
    }
    const bool uniform = !__any_sync(0xffffffffu, any_in) || __all_sync(0xffffffffu, all_in);      // warp-uniform
    if (uniform) {
      if (x < W) code[row + x] = 0;
      continue;
    }
    if (x < W) {
      const bool xi = x + 1 < W;
      const int x1 = xi ? x + 1 : x;
      const bool b000 = __ldg(r00 + x) > iso, b001 = __ldg(r00 + x1) > iso;
      const bool b010 = __ldg(r01 + x) > iso, b011 = __ldg(r01 + x1) > iso;
      const bool b100 = __ldg(r10 + x) > iso, b101 = __ldg(r10 + x1) > iso;
      const bool b110 = __ldg(r11 + x) > iso, b111 = __ldg(r11 + x1) > iso;
      uint8_t c = 0;
      if (xi && (b001 != b000)) c |= 1;
      if (yi && (b010 != b000)) c |= 2;
      if (zi && (b100 != b000)) c |= 4;
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
        c |= (uint8_t)(c_mc_ntri[k] << 3);
      }
      code[row + x] = c;
      cases[row + x] = cs;
    }
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
