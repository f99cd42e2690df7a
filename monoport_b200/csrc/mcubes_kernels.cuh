// Device code of F3 marching cubes (see mcubes.cu for the pipeline).  Kept free of host API calls and of CUDA runtime
// types so that tests/emu can run these kernels unmodified on the CPU emulation layer of the build container.
//
// Everything after the first pass works on a BIT VOLUME: one occupancy bit (value > iso) per node in linear node order,
// word w = nodes 32w .. 32w+31 (a word may straddle rows).  At 257^3 that is 2.1 MB instead of 68 MB, so the volume is
// read from HBM exactly once (bits_kernel, streaming); neighbour lookups become funnel shifts of L1/L2-resident words, and
// the ~95 % of words whose eight shifted copies agree (no surface inside) cost a dozen loads and seven XORs.
#pragma once
#include <stdint.h>
#include "mp_scan.cuh"
#include "mc_table.inc"

namespace mcubes {

constexpr int kBitsThreads = 256;
constexpr int kBitsUnroll = 8;      // words (= 128 B lines of the volume) a warp has in flight per iteration

// zero words the bit volume carries behind the last node so that neighbour reads (+1, +W, +H*W and their sums) of the
// last words never leave the allocation
__host__ __device__ inline long long bits_pad_words(int H, int W) { return ((long long)H * W + W + 1) / 32 + 3; }

// ---- pass 1: volume -> bits.  HBM-bound: 4 B read per node, 1 bit written.
__global__ void __launch_bounds__(kBitsThreads)
bits_kernel(const float* __restrict__ vol, uint32_t* __restrict__ bits, long long n, float iso) {
  const long long n_words = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kBitsThreads / 32) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (kBitsThreads / 32);
  for (long long w0 = warp0 * kBitsUnroll; w0 < n_words; w0 += n_warps * kBitsUnroll) {     // (warp-uniform trip count)
    float v[kBitsUnroll];
#pragma unroll
    for (int j = 0; j < kBitsUnroll; ++j) {
      const long long i = (w0 + j) * 32 + lane;
      v[j] = i < n ? __ldg(vol + i) : iso;                  // (iso > iso is false: nodes past the end read as outside)
    }
#pragma unroll
    for (int j = 0; j < kBitsUnroll; ++j) {
      const uint32_t m = __ballot_sync(0xffffffffu, v[j] > iso);
      if (lane == j && w0 + j < n_words) bits[w0 + j] = m;  // lanes 0..7 store eight consecutive words: one 32 B segment
    }
  }
}

// bits of nodes 32w + s .. 32w + s + 31
__device__ __forceinline__ uint32_t shifted_word(const uint32_t* __restrict__ bits, long long w, long long s) {
  const long long i = 32 * w + s;
  const long long q = i >> 5;
  return __funnelshift_r(__ldg(bits + q), __ldg(bits + q + 1), (uint32_t)(i & 31));
}

// the eight corner bit-vectors of the cells whose corner 0 is a node of word w: index c = dx | dy << 1 | dz << 2
__device__ __forceinline__ void corner_words(const uint32_t* __restrict__ bits, long long w, int H, int W, uint32_t (&v)[8]) {
  const long long sy = W, sz = (long long)H * W;
  v[0] = __ldg(bits + w);
  v[1] = shifted_word(bits, w, 1);
  v[2] = shifted_word(bits, w, sy);
  v[3] = shifted_word(bits, w, sy + 1);
  v[4] = shifted_word(bits, w, sz);
  v[5] = shifted_word(bits, w, sz + 1);
  v[6] = shifted_word(bits, w, sz + sy);
  v[7] = shifted_word(bits, w, sz + sy + 1);
}

__device__ __forceinline__ int case_of(const uint32_t (&v)[8], int b) {
  int k = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) k |= (int)((v[c] >> b) & 1u) << c;
  return k;
}

// per word: x / y / z edge masks (bit b: node 32w+b owns an active +x / +y / +z edge) and the triangle count of its cells
struct WordInfo {
  uint32_t ex, ey, ez, nt;
};

// ---- pass 2 (runs inside the scan's first pass): edge masks + triangle count of one word.  Uniform words (all eight
// shifted copies equal) leave at once.
__device__ __forceinline__ WordInfo classify_word(const uint32_t* __restrict__ bits, long long w, long long n, int D, int H, int W) {
  uint32_t v[8];
  corner_words(bits, w, H, W, v);
  uint32_t mixed = 0;
#pragma unroll
  for (int c = 1; c < 8; ++c) mixed |= v[0] ^ v[c];
  WordInfo wi{0u, 0u, 0u, 0u};
  if (mixed) {
    // which nodes of the word have a +x / +y / +z neighbour (a word may straddle rows and planes)
    uint32_t xi = 0, yi = 0, zi = 0;
    const long long i0 = 32 * w;
    int x = (int)(i0 % W);
    const long long t = i0 / W;
    int y = (int)(t % H), z = (int)(t / H);
    for (int b = 0; b < 32 && i0 + b < n; ++b) {
      if (x + 1 < W) xi |= 1u << b;
      if (y + 1 < H) yi |= 1u << b;
      if (z + 1 < D) zi |= 1u << b;
      if (++x == W) { x = 0; if (++y == H) { y = 0; ++z; } }
    }
    wi.ex = (v[0] ^ v[1]) & xi;
    wi.ey = (v[0] ^ v[2]) & yi;
    wi.ez = (v[0] ^ v[4]) & zi;
    uint32_t cm = mixed & xi & yi & zi;
    while (cm) {
      const int b = __ffs((int)cm) - 1;
      cm &= cm - 1;
      wi.nt += (uint32_t)__ldg(&g_mc_tri[case_of(v, b)][15]);      // byte 15 of a table row = its triangle count
    }
  }
  return wi;
}

__device__ __forceinline__ unsigned long long word_counts(const uint4 q) {
  return (unsigned long long)(__popc(q.x) + __popc(q.y) + __popc(q.z)) | ((unsigned long long)q.w << 32);
}

// functor of the scan's FIRST pass: classifies the word, stores its info and returns its (vertex, triangle) counts
struct ClassifyCountF {
  const uint32_t* bits;
  WordInfo* info;
  long long n;
  int D, H, W;
  static constexpr bool kVec8 = false;
  __device__ unsigned long long operator()(long long w) const {
    const WordInfo wi = classify_word(bits, w, n, D, H, W);
    *reinterpret_cast<uint4*>(info + w) = make_uint4(wi.ex, wi.ey, wi.ez, wi.nt);
    return word_counts(make_uint4(wi.ex, wi.ey, wi.ez, wi.nt));
  }
};

// functor of the scan's SECOND pass (and of everything after it): the stored info
struct WordCountF {    // low 32: vertices owned by the word's nodes, high 32: triangles of its cells
  const WordInfo* info;
  static constexpr bool kVec8 = false;
  __device__ unsigned long long operator()(long long w) const {
    // (plain load, not __ldg: the first pass of the same scan wrote it)
    return word_counts(*reinterpret_cast<const uint4*>(info + w));
  }
};

// exclusive (vertex, triangle) prefix of every word; words with a surface inside are also appended to the ACTIVE LIST the
// emission walks (one warp per entry).  The list's order is whatever the atomics give -- it does not matter: every word
// carries its own output offsets, so the mesh is the same for any order.
struct PrefixEmit {
  unsigned long long* prefix;
  uint32_t* active;            // [n_words] word indices
  uint32_t* n_active;          // zeroed before the scan
  __device__ void operator()(long long w, unsigned long long v, unsigned long long pre) const {
    prefix[w] = pre;
    if (v) active[atomicAdd(n_active, 1u)] = (uint32_t)w;
  }
};

// ---- pass 4: emission.  One warp per word with a surface inside (the active list), one lane per node: up to three
// vertices (vertex id = rank of (node, axis) in node order) and up to MC_MAX_TRI triangles (face order = (cell linear
// index, table order); their offsets come from a warp scan).  Walking 32 consecutive words per warp instead left most warps
// with nothing and a few with a dozen dependent chains in a row (63 us at 257^3).
constexpr int kEmitThreads = 256;
__global__ void __launch_bounds__(kEmitThreads)
mesh_emit_kernel(const float* __restrict__ vol, const uint32_t* __restrict__ bits, const WordInfo* __restrict__ info,
                 const unsigned long long* __restrict__ prefix, const uint32_t* __restrict__ active, uint32_t n_active,
                 float* __restrict__ verts, int32_t* __restrict__ faces, int D, int H, int W, long long n, float iso) {
  __shared__ int32_t s_ids[kEmitThreads / 32][12][32];
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (kEmitThreads / 32) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (kEmitThreads / 32);
  const int plane = H * W;
  for (long long e = warp0; e < (long long)n_active; e += n_warps) {             // (warp-uniform trip count)
    const long long w = (long long)__ldg(active + e);
    const uint4 wi = __ldg(reinterpret_cast<const uint4*>(info + w));
    const uint32_t ex = wi.x, ey = wi.y, ez = wi.z, nt = wi.w;
    const unsigned long long pre = __ldg(prefix + w);
    const long long i = 32 * w + lane;                                          // this lane's node
    const int z = (int)(i / plane), r = (int)(i - (long long)z * plane), y = r / W, x = r - y * W;
    // ---- vertices on the owned edges
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t code = ((ex >> lane) & 1u) | (((ey >> lane) & 1u) << 1) | (((ez >> lane) & 1u) << 2);
    if (code) {
      uint32_t vi = (uint32_t)pre + __popc(ex & below) + __popc(ey & below) + __popc(ez & below);
      const float va = __ldg(vol + i);
#pragma unroll
      for (int axis = 0; axis < 3; ++axis) {
        if (!((code >> axis) & 1u)) continue;
        const long long step = axis == 0 ? 1 : (axis == 1 ? W : plane);
        const float vb = __ldg(vol + i + step);
        const float t = __fdiv_rn(__fsub_rn(iso, va), __fsub_rn(vb, va));
        float p[3] = {(float)x, (float)y, (float)z};
        p[axis] = __fadd_rn(p[axis], t);
        float* o = verts + 3ll * vi;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        ++vi;
      }
    }
    // ---- triangles of the cell whose corner 0 is this node
    if (nt) {                                                                   // (warp-uniform)
      uint32_t v[8];
      corner_words(bits, w, H, W, v);
      const bool cell = i < n && x + 1 < W && y + 1 < H && z + 1 < D;
      const int k = cell ? case_of(v, lane) : 0;
      // one 16-byte load: bytes 0..14 = edge ids of the cell's triangles, byte 15 = their number
      const uint4 row = __ldg(reinterpret_cast<const uint4*>(&g_mc_tri[k][0]));
      const uint32_t rw[4] = {row.x, row.y, row.z, row.w};
      const int mytri = (int)(row.w >> 24);
      // exclusive warp scan of the triangle counts
      unsigned long long incl = mpscan::warp_incl_scan((unsigned long long)mytri, lane);
      const uint32_t foff = (uint32_t)(pre >> 32) + (uint32_t)incl - (uint32_t)mytri;
      // ---- ids of the vertices on the cell's 12 edges.  They live on 7 nodes in 4 node rows (dy, dz); the 32 cells of the
      // warp take each row from at most two consecutive words, so (info, prefix) of those words are WARP-UNIFORM loads and
      // every id is a few popcounts on registers -- not a 24-byte lookup per face corner.  The ids go through shared memory
      // ([edge][lane], conflict-free) because the face corners index them by a run-time edge number.
      int32_t* my_ids = s_ids[threadIdx.x >> 5][0];
      const long long last_word = ((n + 31) >> 5) - 1;
#pragma unroll
      for (int rw4 = 0; rw4 < 4; ++rw4) {
        const int oy = rw4 & 1, oz = rw4 >> 1;
        const long long base = 32 * w + ((long long)oz * H + oy) * W;            // node of lane 0 in this row
        const long long wa = min(base >> 5, last_word), wb = min((base >> 5) + 1, last_word);
        const int sft = (int)(base & 31);
        const uint4 qa = __ldg(reinterpret_cast<const uint4*>(info + wa)), qb = __ldg(reinterpret_cast<const uint4*>(info + wb));
        const uint32_t pa = (uint32_t)__ldg(prefix + wa), pb = (uint32_t)__ldg(prefix + wb);
#pragma unroll
        for (int ox = 0; ox < 2; ++ox) {
          if (ox == 1 && rw4 == 3) continue;                                     // node (1,1,1) owns none of the cell's edges
          const int nb = sft + lane + ox;
          const bool hi = nb >= 32;
          const int bb = nb & 31;
          const uint32_t qx = hi ? qb.x : qa.x, qy = hi ? qb.y : qa.y, qz = hi ? qb.z : qa.z;
          const uint32_t bel = (1u << bb) - 1u;
          const uint32_t r0 = (hi ? pb : pa) + __popc(qx & bel) + __popc(qy & bel) + __popc(qz & bel);   // id of its +x vertex
          const uint32_t bx = (qx >> bb) & 1u, by = (qy >> bb) & 1u;
          // edges 0-3: along x at (dy, dz); 4-7: along y at (dx, dz); 8-11: along z at (dx, dy)
          if (ox == 0) my_ids[(0 + oy + 2 * oz) * 32 + lane] = (int32_t)r0;
          if (oy == 0) my_ids[(4 + ox + 2 * oz) * 32 + lane] = (int32_t)(r0 + bx);
          if (oz == 0) my_ids[(8 + ox + 2 * oy) * 32 + lane] = (int32_t)(r0 + bx + by);
        }
      }
      __syncwarp();
#pragma unroll
      for (int corner = 0; corner < 3 * MC_MAX_TRI; ++corner)
        if (corner < 3 * mytri) {
          const int ed = (int)((rw[corner >> 2] >> (8 * (corner & 3))) & 0xFFu);
          faces[3ll * foff + corner] = my_ids[ed * 32 + lane];
        }
      __syncwarp();                                                              // (the next entry overwrites the ids)
    }
  }
}

}  // namespace mcubes
