// Deterministic (ordered) device-wide exclusive scan / stream compaction used by the octree engine,
// marching cubes and the visible-surface kernel.  Output order is always ascending element index --
// never atomics-ordered -- so index lists and mesh topology are bit-reproducible.
//
// Three launches:  block sums -> scan of block sums (single CTA) -> re-evaluate + local scan + emit.
// The per-element value comes from a functor `uint64 f(i)` (two packed 32-bit counters are allowed),
// `emit(i, value, exclusive_prefix)` consumes the result.  HBM-bound: the functor's reads happen twice.
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

// exclusive scan of one value per thread across the CTA; returns exclusive prefix, *total = CTA sum
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* total) {
  __shared__ unsigned long long warp_sums[kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const unsigned long long s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  if (total) *total = tot;
  return base + incl - v;
}

template <class F>
__global__ void __launch_bounds__(kThreads) block_sums_kernel(F f, long long n, unsigned long long* __restrict__ sums) {
  const long long base = (long long)blockIdx.x * kChunk + (long long)threadIdx.x * kItems;
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const long long i = base + j;
    if (i < n) s += f(i);
  }
  unsigned long long tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// in-place exclusive scan of `nb` block sums by a single CTA; total written to *total (device)
static __global__ void __launch_bounds__(kThreads) scan_sums_kernel(unsigned long long* __restrict__ sums, int nb,
                                                             unsigned long long* __restrict__ total) {
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += kThreads) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < nb ? sums[i] : 0ull;
    unsigned long long tot;
    const unsigned long long ex = block_excl_scan(v, &tot);
    const unsigned long long carry = carry_s;
    if (i < nb) sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

template <class F, class E>
__global__ void __launch_bounds__(kThreads) emit_kernel(F f, E emit, long long n,
                                                        const unsigned long long* __restrict__ offsets) {
  const long long base = (long long)blockIdx.x * kChunk + (long long)threadIdx.x * kItems;
  unsigned long long v[kItems];
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const long long i = base + j;
    v[j] = i < n ? f(i) : 0ull;
    s += v[j];
  }
  unsigned long long run = offsets[blockIdx.x] + block_excl_scan(s, nullptr);
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const long long i = base + j;
    if (i < n) emit(i, v[j], run);
    run += v[j];
  }
}

inline int num_blocks(long long n) { return (int)((n + kChunk - 1) / kChunk); }

// `sums` needs num_blocks(n) entries, `total` one entry (both device memory).
template <class F, class E>
inline cudaError_t scan_emit(F f, E emit, long long n, unsigned long long* sums, unsigned long long* total,
                             cudaStream_t st) {
  if (n <= 0) return cudaMemsetAsync(total, 0, sizeof(unsigned long long), st);
  const int nb = num_blocks(n);
  block_sums_kernel<F><<<nb, kThreads, 0, st>>>(f, n, sums);
  scan_sums_kernel<<<1, kThreads, 0, st>>>(sums, nb, total);
  emit_kernel<F, E><<<nb, kThreads, 0, st>>>(f, emit, n, sums);
  return cudaGetLastError();
}

}  // namespace mpscan
