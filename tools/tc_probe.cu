// Stand-alone probe of the tcgen05 building blocks used by monoport_b200/csrc/query_tc.cu (same tc_ptx.cuh helpers):
//   test 1: SS MMA  D[128x256] = A[128xK] * B[256xK]^T   (K-major SWIZZLE_128B tiles in smem, K = 128 = 2 K-blocks)
//   test 2: TS MMA  A read from TMEM (packed fp16 written with tcgen05.st)
//   test 3: SS MMA  N = 128 into a column offset, accumulate on top of test-1 style result
//   test 4: cta_group::2 MMA (M=256 over a 2-CTA cluster, B split across the pair)
// Each test runs in its own process (`tc_probe <n>`) so that a trap in one does not poison the others.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cooperative_groups.h>

#include "../monoport_b200/csrc/tc_ptx.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int M = 128, N = 256, K = 128;

// smem: A tiles [2][128x64] (2*16KB), B tiles [2][256x64] (2*32KB)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 2 x 16 KB
  uint8_t* sB = smem + 2 * 16384;     // 2 x 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // stage operands into the canonical swizzled layout (generic proxy writes)
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(sA + (k / 64) * 16384 + tc::sw128_offset(r, k % 64)) = A[i];
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(sB + (k / 64) * 32768 + tc::sw128_offset(r, k % 64)) = B[i];
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();      // make the generic-proxy smem writes visible to the tensor core (async proxy)
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (mode == 2) {
    // A -> TMEM as packed fp16: row = lane, 32-bit column j holds k = 2j (low half), 2j+1 (high half); K=128 -> 64 cols at col 256
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
      const int r = tid;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __half2 h = __halves2half2(A[r * K + 2 * (c0 + j)], A[r * K + 2 * (c0 + j) + 1]);
        v[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tc::tmem_st16(tbase + lane_base + 256 + c0, v);
    }
    tc::tmem_st_wait();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
  }

  if (tid == 0) {
    if (mode == 1 || mode == 2) {
      const uint32_t idesc = tc::make_idesc_f16(128, 256);
      for (int kb = 0; kb < 2; ++kb) {
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t bd = tc::make_sdesc_sw128(tc::smem_u32(sB + kb * 32768) + kk * 32, 1024);
          if (mode == 1) {
            const uint64_t ad = tc::make_sdesc_sw128(tc::smem_u32(sA + kb * 16384) + kk * 32, 1024);
            tc::mma_ss(tbase, ad, bd, idesc, (kb | kk) ? 1u : 0u);
          } else {
            tc::mma_ts(tbase, tbase + 256 + (kb * 4 + kk) * 8, bd, idesc, (kb | kk) ? 1u : 0u);
          }
        }
      }
    } else if (mode == 3) {
      // two N=128 MMAs into column halves [0,128) and [128,256): rows 0..127 / 128..255 of B
      const uint32_t idesc = tc::make_idesc_f16(128, 128);
      for (int half = 0; half < 2; ++half)
        for (int kb = 0; kb < 2; ++kb)
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = tc::make_sdesc_sw128(tc::smem_u32(sA + kb * 16384) + kk * 32, 1024);
            const uint64_t bd = tc::make_sdesc_sw128(tc::smem_u32(sB + kb * 32768 + half * 16384) + kk * 32, 1024);
            tc::mma_ss(tbase + half * 128, ad, bd, idesc, (kb | kk) ? 1u : 0u);
          }
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tcgen05_fence_after();
  // read back: thread t owns row t
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld32(tbase + lane_base + c0, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------- cta_group::2
namespace tc2 {
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
}  // namespace tc2

// D[256 x 256] = A[256 x K] * B[256 x K]^T ; CTA r owns A rows [128r,128r+128) and B rows [128r, 128r+128)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 2 x 16 KB  (own 128 rows)
  uint8_t* sB = smem + 2 * 16384;     // 2 x 16 KB  (own 128 of the 256 B rows)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = tc2::cta_rank();
  for (int i = tid; i < 128 * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(sA + (k / 64) * 16384 + tc::sw128_offset(r, k % 64)) = A[(rank * 128 + r) * K + k];
    *reinterpret_cast<__half*>(sB + (k / 64) * 16384 + tc::sw128_offset(r, k % 64)) = B[(rank * 128 + r) * K + k];
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc2::tmem_alloc2(&tmem_base_s, 256); tc2::tmem_relinquish2(); }
  tc::fence_proxy_async_smem();
  tc::tcgen05_fence_before();
  tc2::cluster_sync();
  tc::tcgen05_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = tc::make_idesc_f16(256, 256);
    for (int kb = 0; kb < 2; ++kb)
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = tc::make_sdesc_sw128(tc::smem_u32(sA + kb * 16384) + kk * 32, 1024);
        const uint64_t bd = tc::make_sdesc_sw128(tc::smem_u32(sB + kb * 16384) + kk * 32, 1024);
        tc2::mma_ss2(tbase, ad, bd, idesc, (kb | kk) ? 1u : 0u);
      }
    tc2::mma_commit2(&bar, 3);
  }
  tc::mbar_wait(&bar, 0);
  tc::tcgen05_fence_after();
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld32(tbase + lane_base + c0, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(rank * 128 + tid) * 256 + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tcgen05_fence_before();
  tc2::cluster_sync();
  if (warp == 0) tc2::tmem_dealloc2(tbase, 256);
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;
  const int MM = mode == 4 ? 256 : M;
  std::vector<__half> hA(MM * K), hB(N * K);
  std::vector<float> fA(MM * K), fB(N * K), ref((size_t)MM * N), got((size_t)MM * N);
  srand(1234 + mode);
  for (int i = 0; i < MM * K; ++i) { fA[i] = (float)((rand() % 17) - 8) / 8.0f; hA[i] = __float2half(fA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = (float)((rand() % 13) - 6) / 4.0f; hB[i] = __float2half(fB[i]); }
  for (int m = 0; m < MM; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += fA[m * K + k] * fB[n * K + k];
      ref[(size_t)m * N + n] = s;
    }
  __half *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, got.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, got.size() * 4));
  if (mode == 4) {
    const int smem = 4 * 16384 + 1024;
    CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe2_kernel<<<2, 128, smem>>>(dA, dB, dD);
  } else {
    const int smem = 2 * 16384 + 2 * 32768 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, mode);
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; long bad = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    const double e = fabs((double)got[i] - ref[i]);
    if (!(e <= 1e-3)) ++bad;
    if (e > maxerr || e != e) maxerr = e;
  }
  printf("tc_probe mode %d: max|err| = %g, mismatches = %ld / %zu  -> %s\n", mode, maxerr, bad, got.size(), bad ? "FAIL" : "PASS");
  if (bad) {
    for (int m = 0; m < 4; ++m) { for (int n = 0; n < 8; ++n) printf(" %8.3f/%8.3f", got[(size_t)m * N + n], ref[(size_t)m * N + n]); printf("\n"); }
  }
  return bad ? 1 : 0;
}
