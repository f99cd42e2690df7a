// Shared internals of libmonoport_b200 (sm_100a only).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <vector>

#include "../../include/monoport_b200.h"

// dynamic shared memory of a kernel (tests/emu runs the kernels on the CPU, where it is a buffer the harness hands out --
// one per CTA of a cluster)
#ifdef MP_CUDA_EMU
void* mp_emu_dyn_smem();
#define MP_DYN_SMEM(type, name) type* const name = static_cast<type*>(mp_emu_dyn_smem())
#define MP_DYN_SMEM_ALIGNED(type, name, bytes) type* const name = static_cast<type*>(mp_emu_dyn_smem())
#else
#define MP_DYN_SMEM(type, name) extern __shared__ type name[]
#define MP_DYN_SMEM_ALIGNED(type, name, bytes) extern __shared__ __align__(bytes) type name[]
#endif

// NVTX ranges around the entry points of the hot path (F1 query, F2 octree, F3 marching cubes / visible surface): header-only
// NVTX3, a no-op unless a profiler (nsys / ncu --nvtx) is attached
#ifndef MP_CUDA_EMU
#include <nvtx3/nvToolsExt.h>
struct MpRange {
  explicit MpRange(const char* name) { nvtxRangePushA(name); }
  ~MpRange() { nvtxRangePop(); }
};
#else
struct MpRange {
  explicit MpRange(const char*) {}
};
#endif

#define MP_LEAKY_SLOPE 0.01f   // F.leaky_relu default (heads/SurfaceClassifier.py:58)
#define MP_MAX_LAYERS 8

void mp_set_error(const char* fmt, ...);
void mp_ensure_pool();

#define MP_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      mp_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return MP_E_CUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define MP_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      mp_set_error(__VA_ARGS__);   \
      return MP_E_INVALID;         \
    }                              \
  } while (0)

// ---------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------
struct mp_mlp {
  int n_layers;
  int channels[MP_MAX_LAYERS + 1];   // filter_channels
  int cin[MP_MAX_LAYERS];            // per-layer total input width (hidden + skip)
  int cout[MP_MAX_LAYERS];
  int skip;                          // !no_residual
  int last_op;
  // fp32 path: per layer W^T as [cin][cout] (k-major rows, cout contiguous) + bias
  float* wt[MP_MAX_LAYERS];
  float* bias[MP_MAX_LAYERS];
  // original [cout][cin] fp32 (row-major) kept for the fp32 last layer of the tcgen05 path
  float* w[MP_MAX_LAYERS];
  // tcgen05 path (query_tc.cu): packed fp16 weight tiles + fp32 side vectors; null when unsupported
  void* tc;
  int tc_ok;
  int device;
  unsigned long long gen;            // unique per mp_mlp_create (process-wide counter, never 0): keys per-feature caches
  // validated range of the tensor-core program: frames whose largest |feature| exceeds this limit are evaluated by the
  // exact fp32 kernel instead (decided on the DEVICE, per launch; see mp_query_dispatch).  +inf disables the guard.
  float tc_amax_limit;
};

struct mp_feat {
  int C, H, W;
  float* nhwc32;    // [H][W][C] fp32 (one bilinear tap = one contiguous C*4-byte vector): the CURRENT map -- the handle's own
                    // repacked copy (nhwc_own) or, after mp_feat_bind_nhwc, the caller's channel-last tensor itself
  float* nhwc_own;
  float* staging;   // [C][H][W] device staging for host uploads
  int device;
  // layer-0 pre-activation per texel, G0 = W0[:, :C] . F  ([H*W][g0_n] fp16), built lazily by the tcgen05 v3 path
  // (bilinear sampling is linear, so sampling G0 equals applying W0 to the sampled features); valid for
  // (g0_owner == generation id of the head handle, g0_version == version).  The owner is the head's generation id, not
  // its address: a head rebuilt after a weight change may be handed the freed handle's address again.
  __half* g0;
  __half* f16;      // [H*W][C] fp16 copy of the map (X operand taps of the v3 program), same validity as g0
  float* s4tex;     // [H*W][n_out] fp32: last layer's feature part applied per texel (sampled in fp32 by the v3 program)
  int g0_n;
  unsigned long long g0_owner;
  unsigned long long g0_version;
  unsigned long long version;   // bumped by every mp_feat_upload
  // max |feature| of the current frame as float bits (non-negative floats order like unsigned integers): zeroed by every
  // upload, raised by the staging pass of the per-frame G0 kernels; read by the range guard of the query kernels
  unsigned* amax;
};

// ---------------------------------------------------------------------------------------------
// where the points of a query come from
// ---------------------------------------------------------------------------------------------
enum { MP_SRC_ROWS = 0, MP_SRC_GRID = 1, MP_SRC_NODES = 2 };

struct MpPointSrc {
  int kind;
  // MP_SRC_ROWS: three rows x[n], y[n], z[n]
  const float* px;
  const float* py;
  const float* pz;
  long long pstride;          // element stride between consecutive points
  // MP_SRC_NODES: linear node indices of a res^3 level grid; count may live on the device
  const int32_t* nodes;
  const int32_t* count_dev;   // nullable; when set, n is only an upper bound
  // MP_SRC_GRID / MP_SRC_NODES geometry
  int res;          // nodes per axis of the (level) grid
  int node_stride;  // index-space stride of a node (final-resolution units)
  int r_final;      // final resolution (divisor of the world mapping)
  int z0;           // (host-side bookkeeping only: first plane of a MP_SRC_GRID slab; the kernels use lin0)
  long long lin0;   // MP_SRC_GRID: linear index (z slowest) of the node point 0 stands for -- a z slab starts at z0*res*res,
                    // a balanced shard anywhere (contiguous z-major ranges need not end on plane boundaries)
  // list sharding (multi-GPU): of the n points (after the device-side count) this launch evaluates only the window
  // [rank * per, rank * per + per) with per = ceil(n / world) rounded up to a multiple of 128; 0/1 = everything
  int shard_rank, shard_world;
  float inv_r, half_inv_r;    // 1/r_final, 1/(2 r_final) rounded to fp32
  float bmin[3], bext[3];     // b_min, (b_max - b_min)
  long long n;
};

struct MpCalib {
  float m[12];        // rows 0..2 of [R|t]
  int has;            // 0 => calibs=None (points already in image space)
  int perspective;
  float z_scale;
};

// where results go
#define MP_MAX_PEERS 8
struct MpOutDst {
  float* out;          // MP_SRC_ROWS/GRID: [Res][ld]; value of point i, channel c at out[c*ld + i]
  long long ld;
  float* scatter_vol;  // if non-null: channel 0 is ALSO scattered to scatter_vol[nodes[i]]
  // fused slab exchange of the z-sharded grid query (mp_query_grid_peers): channel 0 of point i is ALSO stored at
  // peer[p][peer_off + i] for p < n_peers -- peer-memory pointers (NVLink P2P) to the full volumes of every rank, the
  // own one included -- so the volume is assembled on every GPU while the tiles are computed, without a collective.
  float* peer[MP_MAX_PEERS] = {};
  long long peer_off = 0;
  int n_peers = 0;
};

// ---------------------------------------------------------------------------------------------
// device helpers shared by the fp32 and tcgen05 kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mp_load_point(const MpPointSrc& s, long long i, float& x, float& y, float& z) {
  if (s.kind == MP_SRC_ROWS) {
    x = __ldg(s.px + i * s.pstride);
    y = __ldg(s.py + i * s.pstride);
    z = __ldg(s.pz + i * s.pstride);
    return;
  }
  int ix, iy, iz;
  if (s.kind == MP_SRC_GRID) {
    const long long plane = (long long)s.res * s.res;
    const long long li = i + s.lin0;
    iz = (int)(li / plane);
    const int r = (int)(li % plane);
    iy = r / s.res;
    ix = r - iy * s.res;
  } else {
    const int lin = __ldg(s.nodes + i);
    const int plane = s.res * s.res;
    iz = lin / plane;
    const int r = lin - iz * plane;
    iy = r / s.res;
    ix = r - iy * s.res;
  }
  // world = (c / R + 1/(2R)) * (b_max-b_min) + b_min, each op rounded separately (no FMA contraction)
  // so that the result is bit-identical to the engine's torch expression (oracle/spec.py:level_points)
  const float cx = (float)(ix * s.node_stride), cy = (float)(iy * s.node_stride), cz = (float)(iz * s.node_stride);
  const float R = (float)s.r_final;
  x = __fadd_rn(__fmul_rn(__fadd_rn(__fdiv_rn(cx, R), s.half_inv_r), s.bext[0]), s.bmin[0]);
  y = __fadd_rn(__fmul_rn(__fadd_rn(__fdiv_rn(cy, R), s.half_inv_r), s.bext[1]), s.bmin[1]);
  z = __fadd_rn(__fmul_rn(__fadd_rn(__fdiv_rn(cz, R), s.half_inv_r), s.bext[2]), s.bmin[2]);
}

// the window [i0, i1) of a query's n points that this launch evaluates (see MpPointSrc::shard_rank)
__device__ __forceinline__ void mp_shard_window(const MpPointSrc& s, long long n, long long& i0, long long& i1) {
  i0 = 0; i1 = n;
  if (s.shard_world > 1) {
    long long per = (n + s.shard_world - 1) / s.shard_world;
    per = (per + 127) / 128 * 128;
    i0 = per * s.shard_rank;
    i1 = i0 + per;
    if (i0 > n) i0 = n;
    if (i1 > n) i1 = n;
  }
}

// geometry.py:19-34 / :37-55
__device__ __forceinline__ void mp_project(const MpCalib& c, float x, float y, float z, float& u, float& v, float& w) {
  if (!c.has) {
    u = x; v = y; w = z;
    return;
  }
  u = c.m[0] * x + c.m[1] * y + c.m[2] * z + c.m[3];
  v = c.m[4] * x + c.m[5] * y + c.m[6] * z + c.m[7];
  w = c.m[8] * x + c.m[9] * y + c.m[10] * z + c.m[11];
  if (c.perspective) {
    u = u / w;
    v = v / w;
  }
}

// Bilinear tap set of grid_sample(align_corners=True, padding zeros)  (geometry.py:15)
struct MpTaps {
  int off[4];     // texel index y*W+x of nw, ne, sw, se (clamped; weight is 0 when the tap is outside)
  float wgt[4];
};

__device__ __forceinline__ MpTaps mp_taps(float u, float v, int H, int W) {
  MpTaps t;
  const float ix = ((u + 1.f) / 2.f) * (float)(W - 1);
  const float iy = ((v + 1.f) / 2.f) * (float)(H - 1);
  const float x0 = floorf(ix), y0 = floorf(iy);
  const float x1 = x0 + 1.f, y1 = y0 + 1.f;
  float w_nw = (x1 - ix) * (y1 - iy);
  float w_ne = (ix - x0) * (y1 - iy);
  float w_sw = (x1 - ix) * (iy - y0);
  float w_se = (ix - x0) * (iy - y0);
  const bool x0ok = (x0 >= 0.f) && (x0 <= (float)(W - 1));
  const bool x1ok = (x1 >= 0.f) && (x1 <= (float)(W - 1));
  const bool y0ok = (y0 >= 0.f) && (y0 <= (float)(H - 1));
  const bool y1ok = (y1 >= 0.f) && (y1 <= (float)(H - 1));
  const int xi0 = min(max((int)x0, 0), W - 1), xi1 = min(max((int)x1, 0), W - 1);
  const int yi0 = min(max((int)y0, 0), H - 1), yi1 = min(max((int)y1, 0), H - 1);
  t.off[0] = yi0 * W + xi0; t.wgt[0] = (x0ok && y0ok) ? w_nw : 0.f;
  t.off[1] = yi0 * W + xi1; t.wgt[1] = (x1ok && y0ok) ? w_ne : 0.f;
  t.off[2] = yi1 * W + xi0; t.wgt[2] = (x0ok && y1ok) ? w_sw : 0.f;
  t.off[3] = yi1 * W + xi1; t.wgt[3] = (x1ok && y1ok) ? w_se : 0.f;
  return t;
}

__device__ __forceinline__ float mp_last_op(float v, int last_op) {
  if (last_op == MP_LAST_SIGMOID) return 1.f / (1.f + expf(-v));
  if (last_op == MP_LAST_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ float mp_lrelu(float v) { return v > 0.f ? v : v * MP_LEAKY_SLOPE; }

// host-side launchers implemented in the kernel files
// `guard`: the launch is conditional on the frame's feature range (feat->amax vs mlp->tc_amax_limit), evaluated on the device:
//   MP_GUARD_NONE always run, MP_GUARD_IN_RANGE run when max|feature| <= limit, MP_GUARD_OUT_OF_RANGE run when it is above.
enum { MP_GUARD_NONE = 0, MP_GUARD_IN_RANGE = 1, MP_GUARD_OUT_OF_RANGE = 2 };
int mp_launch_query_fp32(const mp_mlp* mlp, const mp_feat* feat, const MpPointSrc& src, const MpCalib& cal,
                         const MpOutDst& dst, cudaStream_t st, int guard = MP_GUARD_NONE);
int mp_launch_query_tc(const mp_mlp* mlp, mp_feat* feat, const MpPointSrc& src, const MpCalib& cal,
                       const MpOutDst& dst, cudaStream_t st, int guard = MP_GUARD_NONE);
// device-side predicate shared by the kernels: true = this launch must do nothing
__device__ __forceinline__ bool mp_guard_skips(const unsigned* amax, float limit, int guard) {
  if (guard == MP_GUARD_NONE || amax == nullptr) return false;
  const bool above = __uint_as_float(*amax) > limit;
  return guard == MP_GUARD_IN_RANGE ? above : !above;
}
// colour head only: vertices (X, Y, R - Z) of the visible surface -> world -> colour -> canvas[X, Y, :]   (query_tc.cu)
int mp_launch_colour_surface(const mp_mlp* mlp, mp_feat* feat, const long long* X, const long long* Y, const float* Z, long long n,
                             int R, const float* b_min3, const float* b_max3, const MpCalib& cal, float* canvas, cudaStream_t st);
int mp_tc_prepare(mp_mlp* mlp);     // builds mlp->tc; sets tc_ok
void mp_tc_release(mp_mlp* mlp);

// generic query dispatch (mode handling) used by the API and the octree engine
int mp_query_dispatch(mp_mlp* mlp, mp_feat* feat, const MpPointSrc& src, const MpCalib& cal, const MpOutDst& dst,
                      int mode, cudaStream_t st);
void mp_fill_calib(MpCalib& c, const float* calib12, int projection, float z_scale);
void mp_fill_grid_geom(MpPointSrc& s, int res, int node_stride, int r_final, const float* bmin, const float* bmax);
