// Rate probe for the layer-1 phase of the fused sample+MLP kernel (monoport_b200/csrc/query_tc.cu): how many cycles does a
// tcgen05.mma (K = 16, N = 256, M = 128 per CTA) take INSIDE the real pipeline -- weights streamed L2 -> smem through a ring,
// A operand chunks written to shared memory by eight worker warps and handed over through mbarriers -- with one CTA per
// tile (cta_group::1) and with a CTA pair (cta_group::2, M = 256, every weight tile split over the two shared memories)?
//
//   tc_rate <cg 1|2> <flags> [stages]      flags: 1 = stream weights (else the ring is filled once and reused)
//                                                 2 = worker warps rewrite the 32 KB A chunk per 16 MMAs + hand-off barriers
//                                                 4 = A operand from tensor memory (the layer-2 / layer-3 hidden parts)
//                                                 8 = weights by tensor-map TMA (cp.async.bulk.tensor) instead of 1-D bulk copies
//                                                     (cg 2: completion of both CTAs' halves lands on the leader's barrier)
//                                                 16 = N = 128 instead of 256          32 = alternate the accumulator every MMA
//                                                 64 = fully unrolled issue loop       128 = a second issuing warp (no streaming:
//                                                      each warp issues half of the MMAs into its own accumulator)
//                                                 512 = warp-uniform issue loop, one elected lane per MMA (the shipped pattern);
//                                                      combines with 1, 2, 4, 8, 16      1024 = (with 2) the workers store 4x the bytes
//                                                 structure of the uniform loop inside the pipeline (with 512): 2048 = every lane
//                                                      polls the barriers (default: an elected poll + __syncwarp); 4096 = ONE elected
//                                                      region per stage (waits, fence, 4 MMAs, commits); 8192 = (with 4096) two stages
//                                                      per region; 16384 = (with 4096) no __syncwarp after it; 32768 = the minimal
//                                                      control loop (streaming only) that reaches the floor
// Prints cycles per MMA (min / mean over the CTAs that issue) and the implied fraction of the 128-cycle floor.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -o tc_rate tools/tc_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#include "../monoport_b200/csrc/tc_ptx.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int kThreads = 384;
constexpr int kStreamStages = 88;            // stages of one tile's weight stream (cycled)
struct Off {
  static constexpr int X = 0;                // 4 K-blocks x 16 KB (A operand, static)
  static constexpr int H0 = 65536;           // 2 x 32 KB chunk buffers
  static constexpr int Wr = 131072;          // 96 KB ring
  static constexpr int Bars = Wr + 98304;
  static constexpr int Total = Bars + 512;
};
enum { B_FULL = 0, B_EMPTY = 6, B_READY0 = 12, B_READY1, B_FREE0, B_FREE1, B_DONE, B_TMEM, B_COUNT };

struct Params {
  const uint8_t* w[2];      // per cluster rank (cg 1: [0])
  CUtensorMap tmap[2];
  unsigned long long* cycles;   // [grid]
  int stages, flags;
};

template <int CG>
__global__ void __launch_bounds__(kThreads, 1) rate_kernel(const __grid_constant__ Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Off::Bars);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + B_TMEM);
  constexpr int Stages = 3 * CG, StageBytes = 32768 / CG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = CG == 2 ? tc::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int flags = prm.flags, n_stages = prm.stages;
  int n_stages_half = n_stages;
  const bool f_stream = flags & 1, f_workers = flags & 2, f_ts = flags & 4, f_tmap = flags & 8;
  const bool f_n128 = flags & 16, f_alt = flags & 32, f_unroll = flags & 64, f_two = flags & 128, f_noacc = flags & 256;   // 256: never accumulate (D is not read)
  const bool f_uniform = flags & 512;   // 512: warp-uniform issue loop, one elected lane per MMA (combines with 1, 2, 4, 8, 16)

  for (int i = tid; i < (Off::Wr + 98304) / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < 6; ++s) { tc::mbar_init(bars + B_FULL + s, 1); tc::mbar_init(bars + B_EMPTY + s, 1); }
    tc::mbar_init(bars + B_READY0, 8 * CG);
    tc::mbar_init(bars + B_READY1, 8 * CG);
    tc::mbar_init(bars + B_FREE0, 1);
    tc::mbar_init(bars + B_FREE1, 1);
    tc::mbar_init(bars + B_DONE, 1);
    tc::fence_barrier_init();
    if (f_tmap) tc::tma_prefetch_desc(&prm.tmap[rank]);
  }
  if (warp == 2) {
    if constexpr (CG == 1) { tc::tmem_alloc(s_tmem, 512); tc::tmem_relinquish(); }
    else { tc::tmem_alloc2(s_tmem, 512); tc::tmem_relinquish2(); }
  }
  tc::fence_proxy_async_smem();
  tc::tcgen05_fence_before();
  if constexpr (CG == 1) __syncthreads(); else tc::cluster_sync_all();
  tc::tcgen05_fence_after();
  const uint32_t tbase = *s_tmem;

  if (warp == 0) {
    // ---- weight producer (every CTA loads its own part of each stage)
    if (lane == 0) {
      const int n_loads = f_stream ? n_stages : Stages;
      for (int it = 0; it < n_loads; ++it) {
        const int slot = it % Stages;
        tc::mbar_wait(bars + B_EMPTY + slot, ((it / Stages) & 1) ^ 1);
        uint8_t* dst = smem + Off::Wr + slot * StageBytes;
        const int s = it % kStreamStages;
        if (f_tmap) {
          if constexpr (CG == 1) {
            tc::mbar_arrive_expect_tx(bars + B_FULL + slot, StageBytes);
            tc::tma_load_2d(dst, &prm.tmap[0], 0, s * (StageBytes / 128), bars + B_FULL + slot);
          } else {
            if (leader) tc::mbar_arrive_expect_tx(bars + B_FULL + slot, 2 * StageBytes);
            tc::tma_load_2d_cg2(dst, &prm.tmap[rank], 0, s * (StageBytes / 128), bars + B_FULL + slot);
          }
        } else {
          // 1-D bulk copies: completion stays CTA-local; with cg 2 the peer forwards it (arrive on the leader's ready... n/a)
          tc::mbar_arrive_expect_tx(bars + B_FULL + slot, StageBytes);
          tc::bulk_g2s(dst, prm.w[rank] + (size_t)s * StageBytes, StageBytes, bars + B_FULL + slot);
        }
      }
    }
  } else if (warp == 1 && leader && f_uniform) {
    // ---- the whole warp runs the loop; only the tcgen05 instructions are predicated on one elected lane
    const uint32_t idesc = tc::make_idesc_f16(128 * CG, f_n128 ? 128 : 256);
    const uint32_t sX = tc::smem_u32(smem + Off::X), sW = tc::smem_u32(smem + Off::Wr);
    for (int sl = 0; sl < Stages; ++sl) tc::mbar_wait(bars + B_FULL + sl, 0);
    tc::tcgen05_fence_after();
    const long long t0 = clock64();
    const uint32_t sH0u = tc::smem_u32(smem + Off::H0);
    const bool f_poll_all = flags & 2048;     // every lane polls the barriers (no elected poll + __syncwarp)
    const bool f_one_region = flags & 4096;   // ONE elected region per stage: waits, fence, 4 MMAs, commits
    const bool f_two_stages = flags & 8192;   // (with 4096) two stages per elected region
    const bool f_no_sync = flags & 16384;     // (with 4096) no __syncwarp after the region
    uint32_t c_ready_u[2] = {0, 0};
    // everything one stage needs, computed by ALL lanes (warp-uniform values, so the tcgen05 operands stay uniform)
    struct St { uint64_t ad0, bd0; uint32_t d, par_ready, par_full, acc0; int slot, b; bool w_ready, w_full, w_free; };
    auto prep = [&](int it) -> St {
      St q;
      q.slot = it % Stages;
      const int chunk = it >> 2, kb = (it >> 1) & 1;                          // 4 stages (2 K-blocks x 2 N-halves) per chunk
      q.b = chunk & 1;
      const uint32_t a_addr = f_workers ? sH0u + q.b * 32768 + kb * 16384 : sX + (it & 3) * 16384;
      q.ad0 = tc::make_sdesc_sw128(a_addr, 1024);
      q.bd0 = tc::make_sdesc_sw128(sW + q.slot * StageBytes, 1024);
      q.d = tbase + (it & 1) * 256;
      q.w_ready = f_workers && (it & 3) == 0;
      q.w_full = f_stream && it >= Stages;
      q.w_free = f_workers && (it & 3) == 3;
      q.par_ready = c_ready_u[q.b] & 1u;
      q.par_full = (it / Stages) & 1;
      q.acc0 = it < 2 ? 0u : 1u;
      if (q.w_ready) ++c_ready_u[q.b];
      return q;
    };
    auto mmas = [&](const St& q) {                                              // (inside an elected region)
      if (f_ts) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc::mma_ts(q.d, tbase + kk * 8, q.bd0 + 2 * kk, idesc, kk == 0 ? q.acc0 : 1u);
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc::mma_ss(q.d, q.ad0 + 2 * kk, q.bd0 + 2 * kk, idesc, kk == 0 ? q.acc0 : 1u);
      }
      if (f_stream) tc::mma_commit(bars + B_EMPTY + q.slot);
      if (q.w_free) tc::mma_commit(bars + B_FREE0 + q.b);
    };
    auto all_in_one = [&](const St& q) {                                        // (inside an elected region)
      if (q.w_ready) tc::mbar_wait(bars + B_READY0 + q.b, q.par_ready);
      if (q.w_full) tc::mbar_wait(bars + B_FULL + q.slot, q.par_full);
      tc::tcgen05_fence_after();
      mmas(q);
    };
    if (flags & 32768) {      // the minimal loop of profiles/r02_call5_* (streaming only: every lane polls), kept as the control
    for (int it = 0; it < n_stages; ++it) {
        const int slot = it % Stages;
        if (f_stream && it >= Stages) { tc::mbar_wait(bars + B_FULL + slot, (it / Stages) & 1); tc::tcgen05_fence_after(); }
        const uint64_t ad0 = tc::make_sdesc_sw128(sX + (it & 3) * 16384, 1024), bd0 = tc::make_sdesc_sw128(sW + slot * StageBytes, 1024);
        const uint32_t d = tbase + (it & 1) * 256;
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if constexpr (CG == 1) tc::mma_ss(d, ad0 + 2 * kk, bd0 + 2 * kk, idesc, (it < 2 && kk == 0) ? 0u : 1u);
            else tc::mma_ss2(d, ad0 + 2 * kk, bd0 + 2 * kk, idesc, (it < 2 && kk == 0) ? 0u : 1u);
          }
          if (f_stream) { if constexpr (CG == 1) tc::mma_commit(bars + B_EMPTY + slot); else tc::mma_commit2(bars + B_EMPTY + slot); }
        }
        __syncwarp();
      }
    } else if (CG == 2 || !f_one_region) {
      for (int it = 0; it < n_stages; ++it) {
        const St q = prep(it);
        // separate regions: waits (elected poll + __syncwarp, or all lanes), then the MMAs + commits
        if (q.w_ready) {
          if (f_poll_all) tc::mbar_wait(bars + B_READY0 + q.b, q.par_ready);
          else { if (tc::elect_one()) tc::mbar_wait(bars + B_READY0 + q.b, q.par_ready); __syncwarp(); }
        }
        if (q.w_full) {
          if (f_poll_all) tc::mbar_wait(bars + B_FULL + q.slot, q.par_full);
          else { if (tc::elect_one()) tc::mbar_wait(bars + B_FULL + q.slot, q.par_full); __syncwarp(); }
        }
        tc::tcgen05_fence_after();
        if (tc::elect_one()) mmas(q);
        __syncwarp();
      }
    } else if (!f_two_stages) {
      for (int it = 0; it < n_stages; ++it) {
        const St q = prep(it);
        if (tc::elect_one()) all_in_one(q);
        if (!f_no_sync) __syncwarp();
      }
    } else {
      for (int it = 0; it < n_stages; it += 2) {
        const St q0 = prep(it), q1 = prep(it + 1);
        if (tc::elect_one()) { all_in_one(q0); all_in_one(q1); }
        if (!f_no_sync) __syncwarp();
      }
    }
    if (tc::elect_one()) { if constexpr (CG == 1) tc::mma_commit(bars + B_DONE); else tc::mma_commit2(bars + B_DONE); }
    __syncwarp();
    tc::mbar_wait(bars + B_DONE, 0);
    if (lane == 0) prm.cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  } else if (warp == 1 && leader) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_f16(128 * CG, f_n128 ? 128 : 256);
      if (f_two) n_stages_half = n_stages / 2;
      const uint32_t sX = tc::smem_u32(smem + Off::X), sH0 = tc::smem_u32(smem + Off::H0), sW = tc::smem_u32(smem + Off::Wr);
      uint32_t c_ready[2] = {0, 0};
      const long long t0 = clock64();
      bool first = true;
      for (int it = 0; it < n_stages_half; ++it) {
        const int chunk = it >> 2, b = chunk & 1, kb = (it >> 1) & 1;       // 4 stages (2 K-blocks x 2 N-halves) per chunk
        if (f_workers && (it & 3) == 0) {
          tc::mbar_wait_cluster(bars + B_READY0 + b, c_ready[b] & 1u);
          ++c_ready[b];
          tc::tcgen05_fence_after();
        }
        const int slot = it % Stages;
        if (f_stream || it < Stages) {
          if constexpr (CG == 1) tc::mbar_wait(bars + B_FULL + slot, (it / Stages) & 1);
          else if (f_tmap) tc::mbar_wait_cluster(bars + B_FULL + slot, (it / Stages) & 1);
          else tc::mbar_wait(bars + B_FULL + slot, (it / Stages) & 1);   // (cg 2 without tensor maps: leader's half only -- rate test, results unused)
          tc::tcgen05_fence_after();
        }
        const uint32_t w = sW + slot * StageBytes;
        const uint32_t d = tbase + (it & 1) * 256;
        if (f_unroll) {
          const uint32_t a = sX + (it & 3) * 16384;
          const uint64_t ad0 = tc::make_sdesc_sw128(a, 1024), bd0 = tc::make_sdesc_sw128(w, 1024);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t dd = f_alt ? tbase + (kk & 1) * 256 : d;
            if constexpr (CG == 1) tc::mma_ss(dd, ad0 + 2 * kk, bd0 + 2 * kk, idesc, first ? 0u : 1u);
            else tc::mma_ss2(dd, ad0 + 2 * kk, bd0 + 2 * kk, idesc, first ? 0u : 1u);
            if (kk >= 1) first = false;
          }
        } else
#pragma unroll 1
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t bd = tc::make_sdesc_sw128(w + kk * 32, 1024);
          if (f_ts) {
            // (A in TMEM overlaps D's columns here: rate only)
            if constexpr (CG == 1) tc::mma_ts(d, tbase + kk * 8, bd, idesc, first ? 0u : 1u);
            else tc::mma_ts2(d, tbase + kk * 8, bd, idesc, first ? 0u : 1u);
          } else {
            const uint32_t a = f_workers ? sH0 + b * 32768 + kb * 16384 : sX + (it & 3) * 16384;
            const uint64_t ad = tc::make_sdesc_sw128(a + kk * 32, 1024);
            const uint32_t dd = f_alt ? tbase + (kk & 1) * 256 : d;
            if constexpr (CG == 1) tc::mma_ss(dd, ad, bd, idesc, ((first && kk < 2) || f_noacc) ? 0u : 1u);
            else tc::mma_ss2(dd, ad, bd, idesc, ((first && kk < 2) || f_noacc) ? 0u : 1u);
          }
          if (kk >= 1) first = false;
        }
        if (f_stream) { if constexpr (CG == 1) tc::mma_commit(bars + B_EMPTY + slot); else tc::mma_commit2(bars + B_EMPTY + slot); }
        if (f_workers && (it & 3) == 3) { if constexpr (CG == 1) tc::mma_commit(bars + B_FREE0 + b); else tc::mma_commit2(bars + B_FREE0 + b); }
      }
      if constexpr (CG == 1) tc::mma_commit(bars + B_DONE); else tc::mma_commit2(bars + B_DONE);
      tc::mbar_wait(bars + B_DONE, 0);
      prm.cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
  } else if (warp == 3 && f_two && CG == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_f16(128, f_n128 ? 128 : 256);
      const uint32_t sX = tc::smem_u32(smem + Off::X), sW = tc::smem_u32(smem + Off::Wr);
      tc::mbar_wait(bars + B_FULL + 2, 0);
      tc::tcgen05_fence_after();
      bool first = true;
      for (int it = 0; it < n_stages / 2; ++it) {
        const uint32_t w = sW + 2 * 32768;
#pragma unroll 1
        for (int kk = 0; kk < 4; ++kk) {
          tc::mma_ss(tbase + 256, tc::make_sdesc_sw128(sX + (it & 3) * 16384 + kk * 32, 1024), tc::make_sdesc_sw128(w + kk * 32, 1024), idesc,
                     first ? 0u : 1u);
          first = false;
        }
      }
      tc::mma_commit(bars + B_READY0);
    }
  } else if (warp >= 4 && f_workers) {
    // ---- workers: rewrite the 32 KB chunk buffer (what the sampled layer-0 chunk generation stores) and hand it over
    const int wk = warp - 4;
    const int n_chunks = n_stages / 4;
    uint32_t c_free[2] = {0, 0};
    for (int c = 0; c < n_chunks; ++c) {
      const int b = c & 1;
      tc::mbar_wait(bars + B_FREE0 + b, (c_free[b] & 1u) ^ 1u);
      ++c_free[b];
      uint8_t* dst = smem + Off::H0 + b * 32768;
      // 16 rows per warp, 2 K-blocks x 128 B per row: 2 x uint4 per lane per row pair (same store count as gen_chunk)
      for (int rep = 0; rep < ((flags & 1024) ? 4 : 1); ++rep) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int p = wk * 16 + 4 * i + (lane >> 3);
          const uint4 v = make_uint4(c, i + rep, lane, wk);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tc::smem_u32(dst + tc::sw128_offset(p, (lane & 7) * 8))), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tc::smem_u32(dst + 16384 + tc::sw128_offset(p, (lane & 7) * 8))), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) tc::mbar_arrive(bars + B_READY0 + b);
        else tc::mbar_arrive_remote(bars + B_READY0 + b, 0);
      }
    }
  }
  __syncwarp();
  tc::tcgen05_fence_before();
  if constexpr (CG == 1) { __syncthreads(); if (warp == 2) tc::tmem_dealloc(tbase, 512); }
  else { tc::cluster_sync_all(); if (warp == 2) tc::tmem_dealloc2(tbase, 512); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int cg = argc > 1 ? atoi(argv[1]) : 1;
  const int flags = argc > 2 ? atoi(argv[2]) : 0;
  const int stages = argc > 3 ? atoi(argv[3]) : 4096;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  Params prm;
  memset(&prm, 0, sizeof(prm));
  prm.stages = stages & ~3;
  prm.flags = flags;
  const int stage_bytes = 32768 / cg;
  std::vector<uint8_t> host((size_t)kStreamStages * stage_bytes, 0);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  for (int r = 0; r < cg; ++r) {
    uint8_t* d = nullptr;
    CK(cudaMalloc(&d, host.size()));
    CK(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
    prm.w[r] = d;
    const cuuint64_t gdim[2] = {64, (cuuint64_t)kStreamStages * (stage_bytes / 128)};
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t box[2] = {64, (cuuint32_t)(stage_bytes / 128)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(&prm.tmap[r], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 2; }
  }
  const int grid = cg == 1 ? sms : (sms / 2) * 2;
  CK(cudaMalloc(&prm.cycles, grid * sizeof(unsigned long long)));
  CK(cudaMemset(prm.cycles, 0, grid * sizeof(unsigned long long)));
  const int smem = Off::Total + 1024;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0.f;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    if (cg == 1) {
      CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      rate_kernel<1><<<grid, kThreads, smem>>>(prm);
    } else {
      CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, rate_kernel<2>, prm));
    }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  std::vector<unsigned long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), prm.cycles, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  double sum = 0, mn = 1e30; int n = 0;
  for (int i = 0; i < grid; ++i) if (cyc[i]) { sum += (double)cyc[i]; mn = std::min(mn, (double)cyc[i]); ++n; }
  const double mmas = (double)prm.stages * 4;
  // per-SM work: an M = 128*cg MMA keeps cg SMs busy for the floor of 128 cycles
  printf("tc_rate cg=%d flags=%d stages=%d: issuing CTAs %d, cycles/MMA min %.1f mean %.1f (floor 128 => %.2f of floor), %.3f ms, %.1f TFLOP/s\n",
         cg, flags, prm.stages, n, mn / mmas, sum / n / mmas, 128.0 / (sum / n / mmas), ms,
         (double)n * cg * mmas * 2.0 * 128 * 256 * 16 / (ms * 1e-3) / 1e12);
  return 0;
}
