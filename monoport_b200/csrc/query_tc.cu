// F1 (tcgen05 flavour): the fused  project -> mask -> bilinear gather -> z-concat -> 5-layer skip-MLP -> sigmoid -> mask
// kernel of MonoPortNet.query (monoport/lib/modeling/MonoPortNet.py:48-91; heads/SurfaceClassifier.py:39-71) on the
// 5th-generation tensor cores.  Specialised for the shipped geometry head PIFuNetGMLP
// (filter_channels [257,1024,512,256,128,R], skip-concat, heads/SurfaceClassifier.py:74-79) on a 256-channel map.
//
// Arithmetic: layers 0-3 fp16 operands x fp16 weights -> fp32 accumulators in TMEM; the depth column (z_feat) and the
// bias of every layer are applied in fp32 in the epilogue (so K is exactly 256-aligned, no padding); the last layer
// (385 -> R) runs in fp32 on CUDA cores from the fp32 layer-3 accumulators and the fp32-sampled features.
//
// One persistent CTA per SM, a tile = 128 points = the M dimension of every MMA (one TMEM lane per point):
//   X  [128 x 256] fp16 sampled features, K-major SWIZZLE_128B, resident in smem for the whole tile (skip operand of
//      every layer; the skip-concat is just extra K-blocks of the same accumulation);
//   weights stream L2 -> smem through a 3-stage ring of 32 KB tiles filled by cp.async.bulk (TMA engine), pre-packed on
//      the host in exactly the order the MMA warp consumes them;
//   layer 0 is produced in 8 chunks of 128 channels: acc0 (TMEM) -> epilogue (bias, z, leaky-relu, fp16) -> smem H0
//      chunk -> immediately consumed as a K-chunk of layer 1, so the 1024-wide activation never exists in full;
//   TMEM (512 columns) cannot hold layer 1's 512 accumulators next to a layer-0 chunk, so layer 1 is evaluated in two
//      halves of 256 outputs and layer 0 is recomputed for the second half (+22% MMA work, documented in DESIGN.md);
//   H1, H2 are written back to TMEM as packed fp16 and consumed by the next layer as the A operand straight from
//      TMEM (tcgen05.mma with A in tensor memory), so they never touch shared memory.
// TMEM map: [0,256) acc1-half / acc2 / acc3 | [256,384) acc0, later H1 (ch 256..511) | [384,512) H1 (ch 0..255), later H2
//
// Warp roles (384 threads): warp 0 weight producer, warp 1 MMA issuer (one lane), warp 2 TMEM allocator, warps 4-11
// workers: all eight sample the tile's X, then act as two epilogue warpgroups (each drains half of the columns).
#include "mp_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kC = 256;                 // feature channels
constexpr int kTile = 128;              // points per tile
constexpr int kThreads = 384;
constexpr int kWorkers = 256;           // warps 4..11
constexpr int kStages = 3;
constexpr int kStageBytes = 32768;
constexpr int kL0 = 1024, kL1 = 512, kL2 = 256, kL3 = 128;
constexpr int kMaxRes = 1;             // output channels handled by the fp32 tail (PIFuNetGMLP: 1)

// per-tile weight stream (32 KB stages), in MMA consumption order
constexpr int kStagesPerHalf = 16 + 16 + 4;                       // L0 (8 chunks x 2) + L1 hidden (8 x 2) + L1 skip
constexpr int kStagesPerTile = 2 * kStagesPerHalf + 12 + 4;       // + L2 (8 hidden + 4 skip) + L3 (2 hidden + 2 skip)

// TMEM columns
constexpr uint32_t kColAcc1 = 0, kColAcc0 = 256, kColH1lo = 384, kColH1hi = 256, kColH2 = 384;

struct TcPack {
  __half* wstream;        // kStagesPerTile * 32 KB
  float* bias[4];         // per hidden layer
  float* wz[4];           // z column of every hidden layer
  float* w4h;             // [R][128]  last layer, hidden part
  float* w4s;             // [R][256]  last layer, feature part
  float* w4z;             // [R]
  float* b4;              // [R]
  int res;
};

struct TcParams {
  const __half* wstream;
  const float* bias[4];
  const float* wz[4];
  const float* w4h;
  const float* w4s;
  const float* w4z;
  const float* b4;
  int res;
  int last_op;
  int H, W;
  const __half* feat;     // NHWC fp16
};

struct Smem {
  // offsets inside the 1024-aligned dynamic shared memory block
  static constexpr int X = 0;                                   // 4 K-blocks x 16 KB
  static constexpr int H0 = X + 65536;                          // 2 buffers x (2 K-blocks x 16 KB)
  static constexpr int Wr = H0 + 65536;                         // kStages x 32 KB
  static constexpr int Small = Wr + kStages * kStageBytes;      // zf[128], inimg[128], s4[kMaxRes][128]
  static constexpr int Bars = Small + (2 + kMaxRes) * kTile * 4;
  static constexpr int NumBars = 2 * kStages + 13;
  static constexpr int TmemPtr = Bars + NumBars * 8;
  static constexpr int Total = TmemPtr + 16;
};
enum Bar { B_WFULL = 0, B_WEMPTY = kStages, B_XREADY = 2 * kStages, B_ACC0_FULL, B_ACC0_FREE, B_H0_READY0, B_H0_READY1,
           B_H0_FREE0, B_H0_FREE1, B_ACC1_FULL, B_H1_READY, B_ACC2_FULL, B_H2_READY, B_ACC3_FULL, B_TILE_DONE };
static_assert(B_TILE_DONE + 1 == Smem::NumBars, "barrier count");
static_assert(Smem::Total + 1024 <= 232448, "shared memory budget (227 KB per CTA)");

struct PhaseCounter {   // number of completed waits on a barrier -> parity to wait for next
  uint32_t n = 0;
  __device__ __forceinline__ uint32_t parity() const { return n & 1u; }
};

__device__ __forceinline__ void wait_bar(uint64_t* bars, int which, uint32_t& count) {
  tc::mbar_wait(bars + which, count & 1u);
  ++count;
}
// "free"-type barriers: the first use must pass without any arrival
__device__ __forceinline__ void wait_free(uint64_t* bars, int which, uint32_t& count) {
  tc::mbar_wait(bars + which, (count & 1u) ^ 1u);
  ++count;
}

// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
query_tc_kernel(TcParams prm, MpPointSrc src, MpCalib cal, MpOutDst dst) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::Bars);
  float* s_zf = reinterpret_cast<float*>(smem + Smem::Small);
  float* s_in = s_zf + kTile;
  float* s_s4 = s_in + kTile;                                    // [res][128]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Smem::TmemPtr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  long long n = src.n;
  if (src.count_dev) {
    const long long c = *src.count_dev;
    n = c < n ? c : n;
  }
  const long long n_tiles = (n + kTile - 1) / kTile;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(bars + B_WFULL + s, 1); tc::mbar_init(bars + B_WEMPTY + s, 1); }
    tc::mbar_init(bars + B_XREADY, kWorkers);
    tc::mbar_init(bars + B_ACC0_FULL, 1);
    tc::mbar_init(bars + B_ACC0_FREE, kWorkers);
    tc::mbar_init(bars + B_H0_READY0, kWorkers);
    tc::mbar_init(bars + B_H0_READY1, kWorkers);
    tc::mbar_init(bars + B_H0_FREE0, 1);
    tc::mbar_init(bars + B_H0_FREE1, 1);
    tc::mbar_init(bars + B_ACC1_FULL, 1);
    tc::mbar_init(bars + B_H1_READY, kWorkers);
    tc::mbar_init(bars + B_ACC2_FULL, 1);
    tc::mbar_init(bars + B_H2_READY, kWorkers);
    tc::mbar_init(bars + B_ACC3_FULL, 1);
    tc::mbar_init(bars + B_TILE_DONE, kTile);
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(s_tmem, 512);
    tc::tmem_relinquish();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tbase = *s_tmem;

  if (warp == 0) {
    // ============================== weight producer ==============================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < kStagesPerTile; ++s, ++it) {
          const int slot = it % kStages;
          const uint32_t use = it / kStages;
          tc::mbar_wait(bars + B_WEMPTY + slot, (use & 1u) ^ 1u);
          tc::mbar_arrive_expect_tx(bars + B_WFULL + slot, kStageBytes);
          tc::bulk_g2s(smem + Smem::Wr + slot * kStageBytes,
                       reinterpret_cast<const uint8_t*>(prm.wstream) + (size_t)s * kStageBytes, kStageBytes,
                       bars + B_WFULL + slot);
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc128 = tc::make_idesc_f16(128, 128);
      const uint32_t idesc256 = tc::make_idesc_f16(128, 256);
      const uint32_t sX = tc::smem_u32(smem + Smem::X);
      const uint32_t sH0 = tc::smem_u32(smem + Smem::H0);
      const uint32_t sW = tc::smem_u32(smem + Smem::Wr);
      uint32_t it = 0;                                  // weight stage counter (mirrors the producer)
      uint32_t c_xready = 0, c_acc0free = 0, c_h0ready[2] = {0, 0}, c_h1ready = 0, c_h2ready = 0, c_tiledone = 0;

      // fetch the next weight stage; returns its smem address
      auto next_stage = [&]() -> uint32_t {
        const int slot = it % kStages;
        tc::mbar_wait(bars + B_WFULL + slot, (it / kStages) & 1u);
        tc::tcgen05_fence_after();
        return sW + slot * kStageBytes;
      };
      auto release_stage = [&]() {
        tc::mma_commit(bars + B_WEMPTY + (it % kStages));
        ++it;
      };
      // 4 MMAs over one 64-wide K-block, A from smem
      auto kblock_ss = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool& first) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          tc::mma_ss(d, tc::make_sdesc_sw128(a_addr + kk * 32, 1024), tc::make_sdesc_sw128(b_addr + kk * 32, 1024), idesc,
                     first ? 0u : 1u);
          first = false;
        }
      };
      auto kblock_ts = [&](uint32_t d, uint32_t a_tmem, uint32_t b_addr, uint32_t idesc, bool& first) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          tc::mma_ts(d, a_tmem + kk * 8, tc::make_sdesc_sw128(b_addr + kk * 32, 1024), idesc, first ? 0u : 1u);
          first = false;
        }
      };

      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        wait_bar(bars, B_XREADY, c_xready);
        tc::tcgen05_fence_after();
        for (int h = 0; h < 2; ++h) {
          bool first1 = true;
          auto issue_l0 = [&](int c) {
            (void)c;
            wait_free(bars, B_ACC0_FREE, c_acc0free);
            tc::tcgen05_fence_after();
            bool first0 = true;
            for (int s = 0; s < 2; ++s) {
              const uint32_t w = next_stage();
              kblock_ss(tbase + kColAcc0, sX + (2 * s) * 16384, w, idesc128, first0);
              kblock_ss(tbase + kColAcc0, sX + (2 * s + 1) * 16384, w + 16384, idesc128, first0);
              release_stage();
            }
            tc::mma_commit(bars + B_ACC0_FULL);
          };
          auto issue_l1 = [&](int c) {
            const int b = c & 1;
            if (c == 0) {
              // the first layer-1 MMA of a half overwrites [0,256): it must have been drained -- by the previous tile's
              // layer-3 epilogue (h == 0) or by this tile's first-half epilogue (h == 1)
              if (h == 0) {
                if (tile != (long long)blockIdx.x) { wait_bar(bars, B_TILE_DONE, c_tiledone); }
              } else {
                wait_bar(bars, B_H1_READY, c_h1ready);
              }
            }
            wait_bar(bars, B_H0_READY0 + b, c_h0ready[b]);
            tc::tcgen05_fence_after();
            for (int kb = 0; kb < 2; ++kb) {
              const uint32_t w = next_stage();
              kblock_ss(tbase + kColAcc1, sH0 + b * 32768 + kb * 16384, w, idesc256, first1);
              release_stage();
            }
            tc::mma_commit(bars + B_H0_FREE0 + b);
          };
          issue_l0(0);
          for (int c = 0; c < 7; ++c) {
            issue_l0(c + 1);
            issue_l1(c);
          }
          issue_l1(7);
          for (int kb = 0; kb < 4; ++kb) {                       // skip part of layer 1: A = X
            const uint32_t w = next_stage();
            kblock_ss(tbase + kColAcc1, sX + kb * 16384, w, idesc256, first1);
            release_stage();
          }
          tc::mma_commit(bars + B_ACC1_FULL);
        }
        // ---- layer 2: A = H1 from TMEM (8 K-blocks) + X (4 K-blocks) -> acc2 [0,256)
        wait_bar(bars, B_H1_READY, c_h1ready);
        tc::tcgen05_fence_after();
        {
          bool first = true;
          for (int kb = 0; kb < 8; ++kb) {
            const uint32_t w = next_stage();
            const uint32_t a = tbase + (kb < 4 ? kColH1lo + kb * 32 : kColH1hi + (kb - 4) * 32);
            kblock_ts(tbase + kColAcc1, a, w, idesc256, first);
            release_stage();
          }
          for (int kb = 0; kb < 4; ++kb) {
            const uint32_t w = next_stage();
            kblock_ss(tbase + kColAcc1, sX + kb * 16384, w, idesc256, first);
            release_stage();
          }
          tc::mma_commit(bars + B_ACC2_FULL);
        }
        // ---- layer 3: A = H2 from TMEM (4 K-blocks) + X (4 K-blocks) -> acc3 [0,128)
        wait_bar(bars, B_H2_READY, c_h2ready);
        tc::tcgen05_fence_after();
        {
          bool first = true;
          for (int s = 0; s < 2; ++s) {
            const uint32_t w = next_stage();
            kblock_ts(tbase + kColAcc1, tbase + kColH2 + (2 * s) * 32, w, idesc128, first);
            kblock_ts(tbase + kColAcc1, tbase + kColH2 + (2 * s + 1) * 32, w + 16384, idesc128, first);
            release_stage();
          }
          for (int s = 0; s < 2; ++s) {
            const uint32_t w = next_stage();
            kblock_ss(tbase + kColAcc1, sX + (2 * s) * 16384, w, idesc128, first);
            kblock_ss(tbase + kColAcc1, sX + (2 * s + 1) * 16384, w + 16384, idesc128, first);
            release_stage();
          }
          tc::mma_commit(bars + B_ACC3_FULL);
        }
      }
    }
  } else if (warp >= 4) {
    // ============================== workers: sampler + epilogue ==============================
    const int wk = warp - 4;                     // 0..7
    const int wg = wk >> 2;                      // epilogue warpgroup: column half
    const int quarter = warp & 3;                // TMEM lane quarter this warp may touch
    const int row = quarter * 32 + lane;         // the point (TMEM lane) this thread owns in the epilogue
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t c_acc0full = 0, c_h0free[2] = {0, 0}, c_acc1full = 0, c_acc2full = 0, c_acc3full = 0;
    const int res = prm.res;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long p0 = tile * kTile;
      // ---- X is free once the previous tile's last MMAs (layer-3 skip) have completed
      if (tile != (long long)blockIdx.x) {
        if (wg == 1) { wait_bar(bars, B_ACC3_FULL, c_acc3full); }   // wg 0 already waited on it in its acc3 drain
      }
      // ---- sampling: warp wk handles points wk*16 .. +15; lane covers 8 consecutive channels
      {
        const int cbase = lane * 8;
        float w4s[kMaxRes][8];
#pragma unroll
        for (int r = 0; r < kMaxRes; ++r)
#pragma unroll
          for (int j = 0; j < 8; ++j) w4s[r][j] = (r < res) ? __ldg(prm.w4s + r * kC + cbase + j) : 0.f;
        for (int q = 0; q < 16; ++q) {
          const int p = wk * 16 + q;
          const long long i = p0 + p;
          float u = 0.f, v = 0.f, w = 0.f;
          const bool valid = i < n;
          if (valid) {
            float x, y, z;
            mp_load_point(src, i, x, y, z);
            mp_project(cal, x, y, z, u, v, w);
          }
          const bool in_img = valid && (u >= -1.f) && (u <= 1.f) && (v >= -1.f) && (v <= 1.f);
          MpTaps t = mp_taps(valid ? u : 0.f, valid ? v : 0.f, prm.H, prm.W);
          if (!valid || !(u == u) || !(v == v)) {
#pragma unroll
            for (int a = 0; a < 4; ++a) { t.off[a] = 0; t.wgt[a] = 0.f; }
          }
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          uint4 raw[4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
            raw[a] = __ldg(reinterpret_cast<const uint4*>(prm.feat + (size_t)t.off[a] * kC + cbase));
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&raw[a]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(h2[j]);
              acc[2 * j] = (a == 0) ? f.x * t.wgt[a] : acc[2 * j] + f.x * t.wgt[a];
              acc[2 * j + 1] = (a == 0) ? f.y * t.wgt[a] : acc[2 * j + 1] + f.y * t.wgt[a];
            }
          }
          uint4 packed;
          packed.x = tc::pack_half2(acc[0], acc[1]);
          packed.y = tc::pack_half2(acc[2], acc[3]);
          packed.z = tc::pack_half2(acc[4], acc[5]);
          packed.w = tc::pack_half2(acc[6], acc[7]);
          const int kb = lane >> 3;
          *reinterpret_cast<uint4*>(smem + Smem::X + kb * 16384 + tc::sw128_offset(p, (lane & 7) * 8)) = packed;
          // fp32 skip part of the last layer: sum_c w4[128 + c] * x_c  (warp-shuffle reduction)
          const float zf = w * cal.z_scale;
#pragma unroll
          for (int r = 0; r < kMaxRes; ++r) {
            if (r < res) {
              float s = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) s = fmaf(w4s[r][j], acc[j], s);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
              if (lane == 0) s_s4[r * kTile + p] = s + __ldg(prm.w4z + r) * zf + __ldg(prm.b4 + r);
            }
          }
          if (lane == 0) {
            s_zf[p] = zf;
            s_in[p] = in_img ? 1.f : 0.f;
          }
        }
      }
      tc::fence_proxy_async_smem();
      // make s_zf/s_in/s_s4 visible to the epilogue role of all worker threads
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float zf = s_zf[row];
      const float inimg = s_in[row];
      float s4[kMaxRes];
#pragma unroll
      for (int r = 0; r < kMaxRes; ++r) s4[r] = (r < res) ? s_s4[r * kTile + row] : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone has read the per-point scalars
      tc::mbar_arrive(bars + B_XREADY);

      // ---- epilogue helper: acc columns [col0, col0+32) of this thread's row -> activated fp32 values
      auto load_act = [&](uint32_t col, const float* __restrict__ bias, const float* __restrict__ wz, int ch0, float (&o)[32]) {
        uint32_t v[32];
        tc::tmem_ld32(tbase + lane_base + col, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias + ch0) + j4);
          const float4 z = __ldg(reinterpret_cast<const float4*>(wz + ch0) + j4);
          o[4 * j4 + 0] = mp_lrelu(__uint_as_float(v[4 * j4 + 0]) + fmaf(z.x, zf, b.x));
          o[4 * j4 + 1] = mp_lrelu(__uint_as_float(v[4 * j4 + 1]) + fmaf(z.y, zf, b.y));
          o[4 * j4 + 2] = mp_lrelu(__uint_as_float(v[4 * j4 + 2]) + fmaf(z.z, zf, b.z));
          o[4 * j4 + 3] = mp_lrelu(__uint_as_float(v[4 * j4 + 3]) + fmaf(z.w, zf, b.w));
        }
      };

      for (int h = 0; h < 2; ++h) {
        // ---- layer-0 chunks -> H0 buffers (smem, A operand of layer 1)
        for (int c = 0; c < 8; ++c) {
          const int b = c & 1;
          wait_bar(bars, B_ACC0_FULL, c_acc0full);
          wait_free(bars, B_H0_FREE0 + b, c_h0free[b]);
          tc::tcgen05_fence_after();
          // this warpgroup drains columns [wg*64, wg*64+64) == K-block `wg` of the chunk
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float o[32];
            const int ch0 = c * 128 + wg * 64 + g * 32;
            load_act(kColAcc0 + wg * 64 + g * 32, prm.bias[0], prm.wz[0], ch0, o);
            uint8_t* dstp = smem + Smem::H0 + b * 32768 + wg * 16384;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 pk;
              pk.x = tc::pack_half2(o[8 * q + 0], o[8 * q + 1]);
              pk.y = tc::pack_half2(o[8 * q + 2], o[8 * q + 3]);
              pk.z = tc::pack_half2(o[8 * q + 4], o[8 * q + 5]);
              pk.w = tc::pack_half2(o[8 * q + 6], o[8 * q + 7]);
              *reinterpret_cast<uint4*>(dstp + tc::sw128_offset(row, g * 32 + q * 8)) = pk;
            }
          }
          tc::fence_proxy_async_smem();
          tc::tcgen05_fence_before();
          tc::mbar_arrive(bars + B_H0_READY0 + b);
          tc::mbar_arrive(bars + B_ACC0_FREE);
        }
        // ---- layer-1 half -> H1 (packed fp16 in TMEM, A operand of layer 2)
        wait_bar(bars, B_ACC1_FULL, c_acc1full);
        tc::tcgen05_fence_after();
        {
          const uint32_t hcol = (h == 0) ? kColH1lo : kColH1hi;
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            float o[32];
            const int lc = wg * 128 + g * 32;                    // column inside the 256-wide half
            load_act(kColAcc1 + lc, prm.bias[1], prm.wz[1], h * 256 + lc, o);
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = tc::pack_half2(o[2 * j], o[2 * j + 1]);
            tc::tmem_st16(tbase + lane_base + hcol + lc / 2, pk);
          }
          tc::tmem_st_wait();
          tc::tcgen05_fence_before();
          tc::mbar_arrive(bars + B_H1_READY);
        }
      }
      // ---- layer 2 -> H2 (packed fp16 in TMEM)
      wait_bar(bars, B_ACC2_FULL, c_acc2full);
      tc::tcgen05_fence_after();
      {
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          float o[32];
          const int lc = wg * 128 + g * 32;
          load_act(kColAcc1 + lc, prm.bias[2], prm.wz[2], lc, o);
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = tc::pack_half2(o[2 * j], o[2 * j + 1]);
          tc::tmem_st16(tbase + lane_base + kColH2 + lc / 2, pk);
        }
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive(bars + B_H2_READY);
      }
      // ---- layer 3 (fp32 accumulators) + layer 4 in fp32 on CUDA cores, warpgroup 0 only
      if (wg == 0) {
        wait_bar(bars, B_ACC3_FULL, c_acc3full);
        tc::tcgen05_fence_after();
        float logit[kMaxRes];
#pragma unroll
        for (int r = 0; r < kMaxRes; ++r) logit[r] = s4[r];
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          float o[32];
          load_act(kColAcc1 + g * 32, prm.bias[3], prm.wz[3], g * 32, o);
#pragma unroll
          for (int r = 0; r < kMaxRes; ++r) {
            if (r < res) {
              const float4* wv = reinterpret_cast<const float4*>(prm.w4h + r * kL3 + g * 32);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 w4 = __ldg(wv + j4);
                logit[r] = fmaf(w4.x, o[4 * j4 + 0], logit[r]);
                logit[r] = fmaf(w4.y, o[4 * j4 + 1], logit[r]);
                logit[r] = fmaf(w4.z, o[4 * j4 + 2], logit[r]);
                logit[r] = fmaf(w4.w, o[4 * j4 + 3], logit[r]);
              }
            }
          }
        }
        tc::tcgen05_fence_before();
        tc::mbar_arrive(bars + B_TILE_DONE);
        const long long i = p0 + row;
        if (i < n) {
#pragma unroll
          for (int r = 0; r < kMaxRes; ++r) {
            if (r < res) {
              const float val = inimg * mp_last_op(logit[r], prm.last_op);      // MonoPortNet.py:89
              if (dst.out) dst.out[(long long)r * dst.ld + i] = val;
              if (dst.scatter_vol && r == 0) dst.scatter_vol[__ldg(src.nodes + i)] = val;
            }
          }
        }
      }
    }
  }
  // ---- teardown
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tbase, 512);
}

// --------------------------------------------------------------------------------------------------------------------
// host side: weight packing
// --------------------------------------------------------------------------------------------------------------------
// writes the [nrows x 64] fp16 K-major SWIZZLE_128B image of W[row0.., col0..col0+64) (W row-major [cout][cin])
void pack_tile(uint8_t* dst, const float* W, int cin, int row0, int nrows, int col0) {
  for (int r = 0; r < nrows; ++r)
    for (int k = 0; k < 64; ++k) {
      const __half h = __float2half_rn(W[(size_t)(row0 + r) * cin + col0 + k]);
      memcpy(dst + tc::sw128_offset(r, k), &h, 2);
    }
}

bool shape_supported(const mp_mlp* m) {
  if (m->n_layers != 5 || !m->skip) return false;
  const int want[5] = {257, 1024, 512, 256, 128};
  for (int l = 0; l < 5; ++l)
    if (m->channels[l] != want[l]) return false;
  return m->channels[5] >= 1 && m->channels[5] <= kMaxRes;
}

}  // namespace

int mp_tc_prepare(mp_mlp* mlp) {
  mlp->tc = nullptr;
  mlp->tc_ok = 0;
  if (!shape_supported(mlp)) return MP_OK;
  int dev = 0, major = 0, max_smem = 0;
  MP_CUDA(cudaGetDevice(&dev));
  MP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MP_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (major != 10 || max_smem < Smem::Total + 1024) return MP_OK;     // tcgen05 is sm_100-family only

  // pull the fp32 weights back to the host and repack
  std::vector<std::vector<float>> W(5), Bv(5);
  for (int l = 0; l < 5; ++l) {
    W[l].resize((size_t)mlp->cin[l] * mlp->cout[l]);
    Bv[l].resize(mlp->cout[l]);
    MP_CUDA(cudaMemcpy(W[l].data(), mlp->w[l], W[l].size() * sizeof(float), cudaMemcpyDeviceToHost));
    MP_CUDA(cudaMemcpy(Bv[l].data(), mlp->bias[l], Bv[l].size() * sizeof(float), cudaMemcpyDeviceToHost));
  }
  std::vector<uint8_t> stream((size_t)kStagesPerTile * kStageBytes, 0);
  size_t st = 0;
  auto stage_ptr = [&]() { return stream.data() + (st++) * kStageBytes; };
  const int cin0 = mlp->cin[0], cin1 = mlp->cin[1], cin2 = mlp->cin[2], cin3 = mlp->cin[3];
  for (int h = 0; h < 2; ++h) {
    auto l0 = [&](int c) {
      for (int s = 0; s < 2; ++s) {
        uint8_t* p = stage_ptr();
        pack_tile(p, W[0].data(), cin0, c * 128, 128, (2 * s) * 64);
        pack_tile(p + 16384, W[0].data(), cin0, c * 128, 128, (2 * s + 1) * 64);
      }
    };
    auto l1 = [&](int c) {
      for (int kb = 0; kb < 2; ++kb) pack_tile(stage_ptr(), W[1].data(), cin1, h * 256, 256, c * 128 + kb * 64);
    };
    l0(0);
    for (int c = 0; c < 7; ++c) { l0(c + 1); l1(c); }
    l1(7);
    for (int kb = 0; kb < 4; ++kb) pack_tile(stage_ptr(), W[1].data(), cin1, h * 256, 256, kL0 + kb * 64);
  }
  for (int kb = 0; kb < 8; ++kb) pack_tile(stage_ptr(), W[2].data(), cin2, 0, 256, kb * 64);
  for (int kb = 0; kb < 4; ++kb) pack_tile(stage_ptr(), W[2].data(), cin2, 0, 256, kL1 + kb * 64);
  for (int s = 0; s < 2; ++s) {
    uint8_t* p = stage_ptr();
    pack_tile(p, W[3].data(), cin3, 0, 128, (2 * s) * 64);
    pack_tile(p + 16384, W[3].data(), cin3, 0, 128, (2 * s + 1) * 64);
  }
  for (int s = 0; s < 2; ++s) {
    uint8_t* p = stage_ptr();
    pack_tile(p, W[3].data(), cin3, 0, 128, kL2 + (2 * s) * 64);
    pack_tile(p + 16384, W[3].data(), cin3, 0, 128, kL2 + (2 * s + 1) * 64);
  }
  if ((int)st != kStagesPerTile) {
    mp_set_error("internal: weight stream has %d stages, expected %d", (int)st, kStagesPerTile);
    return MP_E_INVALID;
  }

  TcPack* pk = new TcPack();
  memset(pk, 0, sizeof(*pk));
  pk->res = mlp->channels[5];
  auto upload = [&](const void* src, size_t bytes, void** dptr) -> cudaError_t {
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
  };
  cudaError_t e = upload(stream.data(), stream.size(), (void**)&pk->wstream);
  const int hid[4] = {0, kL0, kL1, kL2};
  for (int l = 0; l < 4 && e == cudaSuccess; ++l) {
    std::vector<float> wz(mlp->cout[l]);
    const int zcol = hid[l] + kC;
    for (int co = 0; co < mlp->cout[l]; ++co) wz[co] = W[l][(size_t)co * mlp->cin[l] + zcol];
    e = upload(Bv[l].data(), Bv[l].size() * sizeof(float), (void**)&pk->bias[l]);
    if (e == cudaSuccess) e = upload(wz.data(), wz.size() * sizeof(float), (void**)&pk->wz[l]);
  }
  if (e == cudaSuccess) {
    const int R = pk->res, cin4 = mlp->cin[4];
    std::vector<float> w4h((size_t)R * kL3), w4s((size_t)R * kC), w4z(R);
    for (int r = 0; r < R; ++r) {
      for (int j = 0; j < kL3; ++j) w4h[(size_t)r * kL3 + j] = W[4][(size_t)r * cin4 + j];
      for (int j = 0; j < kC; ++j) w4s[(size_t)r * kC + j] = W[4][(size_t)r * cin4 + kL3 + j];
      w4z[r] = W[4][(size_t)r * cin4 + kL3 + kC];
    }
    e = upload(w4h.data(), w4h.size() * sizeof(float), (void**)&pk->w4h);
    if (e == cudaSuccess) e = upload(w4s.data(), w4s.size() * sizeof(float), (void**)&pk->w4s);
    if (e == cudaSuccess) e = upload(w4z.data(), w4z.size() * sizeof(float), (void**)&pk->w4z);
    if (e == cudaSuccess) e = upload(Bv[4].data(), Bv[4].size() * sizeof(float), (void**)&pk->b4);
  }
  mlp->tc = pk;
  if (e != cudaSuccess) {
    mp_set_error("mp_tc_prepare: %s", cudaGetErrorString(e));
    mp_tc_release(mlp);
    return MP_E_CUDA;
  }
  e = cudaFuncSetAttribute(query_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::Total + 1024);
  if (e != cudaSuccess) {
    mp_set_error("mp_tc_prepare: cannot opt in to %d bytes of shared memory: %s", Smem::Total + 1024, cudaGetErrorString(e));
    mp_tc_release(mlp);
    return MP_E_CUDA;
  }
  mlp->tc_ok = 1;
  return MP_OK;
}

void mp_tc_release(mp_mlp* mlp) {
  TcPack* pk = static_cast<TcPack*>(mlp->tc);
  if (!pk) return;
  if (pk->wstream) cudaFree(pk->wstream);
  for (int l = 0; l < 4; ++l) {
    if (pk->bias[l]) cudaFree(pk->bias[l]);
    if (pk->wz[l]) cudaFree(pk->wz[l]);
  }
  if (pk->w4h) cudaFree(pk->w4h);
  if (pk->w4s) cudaFree(pk->w4s);
  if (pk->w4z) cudaFree(pk->w4z);
  if (pk->b4) cudaFree(pk->b4);
  delete pk;
  mlp->tc = nullptr;
  mlp->tc_ok = 0;
}

int mp_launch_query_tc(const mp_mlp* mlp, const mp_feat* feat, const MpPointSrc& src, const MpCalib& cal,
                       const MpOutDst& dst, cudaStream_t st) {
  if (src.n <= 0) return MP_OK;
  const TcPack* pk = static_cast<const TcPack*>(mlp->tc);
  if (!pk || !mlp->tc_ok) {
    mp_set_error("tcgen05 path not prepared for this head");
    return MP_E_UNSUPPORTED;
  }
  if (feat->C != kC) {
    mp_set_error("head expects %d input channels but the feature map has %d (+1 depth)", mlp->channels[0], feat->C);
    return MP_E_INVALID;
  }
  TcParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.wstream = pk->wstream;
  for (int l = 0; l < 4; ++l) { prm.bias[l] = pk->bias[l]; prm.wz[l] = pk->wz[l]; }
  prm.w4h = pk->w4h; prm.w4s = pk->w4s; prm.w4z = pk->w4z; prm.b4 = pk->b4;
  prm.res = pk->res;
  prm.last_op = mlp->last_op;
  prm.H = feat->H; prm.W = feat->W;
  prm.feat = feat->nhwc16;
  int dev = 0, sms = 148;
  MP_CUDA(cudaGetDevice(&dev));
  MP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = (src.n + kTile - 1) / kTile;
  const int grid = (int)(tiles < (long long)sms ? tiles : sms);
  query_tc_kernel<<<grid, kThreads, Smem::Total + 1024, st>>>(prm, src, cal, dst);
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}
