// Deterministic (ordered) device-wide exclusive scan / stream compaction used by the octree engine,
// marching cubes and the visible-surface kernel.  Output order is always ascending element index --
// never atomics-ordered -- so index lists and mesh topology are bit-reproducible.
//
// Two launches:  block sums, whose last CTA to finish (ticket counter) also scans the block sums and runs an optional
// `post(total)` hook (device-side counts for the next kernel) -> re-evaluate + local scan + emit.
// The per-element value comes from a functor `uint64 f(i)` (two packed 32-bit counters are allowed),
// `emit(i, value, exclusive_prefix)` consumes the result.  HBM-bound: the functor's reads happen twice.
// A functor declares `static constexpr bool kVec8`; when true it also provides `load8(i, v[8])` for eight consecutive
// elements starting at a multiple of 8 (one 8-byte load instead of eight byte loads for the flag / code volumes).
#pragma once
#include <stdint.h>

namespace mpscan {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kChunk = kThreads * kItems;   // elements per CTA

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// exclusive scan of one value per thread across the CTA (of NT threads); returns exclusive prefix, *total = CTA sum
template <int NT = kThreads>
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* total) {
  __shared__ unsigned long long warp_sums[NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    const unsigned long long s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  if (total) *total = tot;
  return base + incl - v;
}

struct NoPost {
  __device__ void operator()(unsigned long long) const {}
};

static_assert(kItems == 8, "load8 functors cover exactly one thread's items");

// the kItems values of the thread whose first element is `base` (a multiple of kItems); elements >= n count as 0
template <class F>
__device__ __forceinline__ void load_items(const F& f, long long base, long long n, unsigned long long (&v)[kItems]) {
  if constexpr (F::kVec8) {
    if (base + kItems <= n) {
      f.load8(base, v);
      return;
    }
  }
#pragma unroll
  for (int j = 0; j < kItems; ++j) v[j] = (base + j < n) ? f(base + j) : 0ull;
}

// eight consecutive bytes at p + i (i a multiple of 8, p 8-byte aligned) as one load; byte j lands in b[j]
__device__ __forceinline__ void load_bytes8(const uint8_t* p, long long i, uint32_t (&b)[kItems]) {
  const uint2 w = __ldg(reinterpret_cast<const uint2*>(p + i));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    b[j] = (w.x >> (8 * j)) & 0xFFu;
    b[4 + j] = (w.y >> (8 * j)) & 0xFFu;
  }
}

// Tail of a scan's first pass, called by ALL threads of every CTA with the CTA's total `tot`: sums[0..nb) <- exclusive scan of
// the chunk totals, total[0] <- grand total.  total[1] is the ticket counter: zero on entry (zero-initialised once by the owner
// of the workspace), reset to zero by the last CTA.  NT = threads per CTA.  By default one CTA = one chunk (slot = blockIdx.x,
// n_ctas = nb); with `accumulate` several CTAs add their totals into one chunk's slot (sums[] must then be zero on entry).
template <int NT = kThreads, class P>
__device__ __forceinline__ void finish_block_sums(unsigned long long tot, unsigned long long* __restrict__ sums, int nb,
                                                  unsigned long long* __restrict__ total, P post, int slot, int n_ctas,
                                                  bool accumulate) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    if (accumulate) atomicAdd(&sums[slot], tot);
    else sums[slot] = tot;
    __threadfence();
    is_last = atomicAdd(&total[1], 1ull) == (unsigned long long)(n_ctas - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // the last CTA turns the nb chunk totals into exclusive offsets: thread t owns the contiguous run [t*per, (t+1)*per)
  // (two independent passes over L2-resident values + ONE block scan, instead of a block scan per 256 entries)
  const int per = (nb + NT - 1) / NT;
  const int b0 = min((int)threadIdx.x * per, nb), b1 = min(b0 + per, nb);
  unsigned long long local = 0;
  for (int i = b0; i < b1; ++i) local += __ldcg(sums + i);
  unsigned long long grand;
  unsigned long long run = block_excl_scan<NT>(local, &grand);
  for (int i = b0; i < b1; ++i) {
    const unsigned long long v = __ldcg(sums + i);
    sums[i] = run;
    run += v;
  }
  if (threadIdx.x == 0) {
    total[0] = grand;
    total[1] = 0;
    post(grand);
  }
}

template <class F, class P>
__global__ void __launch_bounds__(kThreads) block_sums_kernel(F f, long long n, unsigned long long* __restrict__ sums, int nb,
                                                              unsigned long long* __restrict__ total, P post) {
  const long long base = (long long)blockIdx.x * kChunk + (long long)threadIdx.x * kItems;
  unsigned long long v[kItems];
  load_items(f, base, n, v);
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) s += v[j];
  unsigned long long tot;
  block_excl_scan(s, &tot);
  finish_block_sums<kThreads>(tot, sums, nb, total, post, (int)blockIdx.x, nb, false);
}

template <class F, class E>
__global__ void __launch_bounds__(kThreads) emit_kernel(F f, E emit, long long n,
                                                        const unsigned long long* __restrict__ offsets) {
  const long long base = (long long)blockIdx.x * kChunk + (long long)threadIdx.x * kItems;
  unsigned long long v[kItems];
  load_items(f, base, n, v);
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) s += v[j];
  unsigned long long run = offsets[blockIdx.x] + block_excl_scan(s, nullptr);
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const long long i = base + j;
    if (i < n) emit(i, v[j], run);
    run += v[j];
  }
}

inline int num_blocks(long long n) { return (int)((n + kChunk - 1) / kChunk); }

#ifdef __CUDACC__   // (the launcher below is the only part tests/emu cannot take)
// `sums` needs num_blocks(n) entries, `total` two: [0] receives the grand total, [1] is the ticket counter and must be
// zero on entry (it is left at zero).  Both device memory.
template <class F, class E, class P = NoPost>
inline cudaError_t scan_emit(F f, E emit, long long n, unsigned long long* sums, unsigned long long* total,
                             cudaStream_t st, P post = P()) {
  const int nb = num_blocks(n > 0 ? n : 1);
  block_sums_kernel<F, P><<<nb, kThreads, 0, st>>>(f, n > 0 ? n : 0, sums, nb, total, post);
  if (n > 0) emit_kernel<F, E><<<nb, kThreads, 0, st>>>(f, emit, n, sums);
  return cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace mpscan
