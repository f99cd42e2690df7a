// F3: marching cubes on the occupancy volume (absent from the reference -- SURVEY.md finding 3; PARITY UNPINNED,
// bit-exact against oracle/spec.py:marching_cubes_ref which uses the same derived table, tools/gen_mc_table.py).
// HBM-bound byte/integer work.  Indexed, watertight mesh with deterministic ids:
//   vertex id = rank of (node, axis) among active owned edges (every grid edge is owned by its lower node);
//   face order = (cell linear index, table order).
// count:  classify_kernel (1 read of the volume: per node 3 owned-edge bits + per cell case/tri count, packed in
//         one byte per node)  ->  ordered scan of (verts, tris) packed in one uint64  -> totals.
// emit :  the scan is re-run with an emit functor that writes vertices (coalesced by node order) and faces.
#include "mp_common.cuh"
#include "mp_scan.cuh"
#include "mc_table.inc"

namespace {

// per node byte: bits 0..2 = owned +x/+y/+z edge active; bits 3..5 = triangle count of the cell whose corner 0 is
// this node (0 when the node is on the +face of the grid)
__global__ void __launch_bounds__(256)
classify_kernel(const float* __restrict__ vol, uint8_t* __restrict__ code, uint8_t* __restrict__ cases, int D, int H,
                int W, float iso) {
  const long long n = (long long)D * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((long long)W * H));
    const bool xi = x + 1 < W, yi = y + 1 < H, zi = z + 1 < D;
    auto in = [&](int dz, int dy, int dx) { return __ldg(vol + i + ((long long)dz * H + dy) * W + dx) > iso; };
    const bool b0 = in(0, 0, 0);
    uint8_t c = 0;
    if (xi && (in(0, 0, 1) != b0)) c |= 1;
    if (yi && (in(0, 1, 0) != b0)) c |= 2;
    if (zi && (in(1, 0, 0) != b0)) c |= 4;
    uint8_t cs = 0;
    if (xi && yi && zi) {
      int k = b0 ? 1 : 0;
      k |= in(0, 0, 1) ? 2 : 0;
      k |= in(0, 1, 0) ? 4 : 0;
      k |= in(0, 1, 1) ? 8 : 0;
      k |= in(1, 0, 0) ? 16 : 0;
      k |= in(1, 0, 1) ? 32 : 0;
      k |= in(1, 1, 0) ? 64 : 0;
      k |= in(1, 1, 1) ? 128 : 0;
      cs = (uint8_t)k;
      c |= (uint8_t)(c_mc_ntri[k] << 3);
    }
    code[i] = c;
    cases[i] = cs;
  }
}

struct CountF {    // low 32: vertices owned by node i, high 32: triangles of cell i
  const uint8_t* code;
  __device__ unsigned long long operator()(long long i) const {
    const uint32_t c = code[i];
    return (unsigned long long)__popc(c & 7u) | ((unsigned long long)(c >> 3) << 32);
  }
};

struct OffsetsEmit {   // stores the exclusive vertex offset of every node (faces need random access to it)
  uint32_t* voff;
  __device__ void operator()(long long i, unsigned long long, unsigned long long pre) const { voff[i] = (uint32_t)pre; }
};

struct MeshEmit {
  const float* vol; const uint8_t* code; const uint8_t* cases; const uint32_t* voff;
  float* verts; int32_t* faces; int D, H, W; float iso;
  __device__ void operator()(long long i, unsigned long long val, unsigned long long pre) const {
    if (!val) return;
    const uint32_t c = code[i];
    const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((long long)W * H));
    // vertices on the owned edges, axis order
    if (c & 7u) {
      uint32_t v = (uint32_t)pre;
      const float va = vol[i];
      const long long step[3] = {1, W, (long long)W * H};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (!(c >> a & 1u)) continue;
        const float vb = vol[i + step[a]];
        const float t = __fdiv_rn(__fsub_rn(iso, va), __fsub_rn(vb, va));
        float p[3] = {(float)x, (float)y, (float)z};
        p[a] = __fadd_rn(p[a], t);
        verts[3ll * v + 0] = p[0]; verts[3ll * v + 1] = p[1]; verts[3ll * v + 2] = p[2];
        ++v;
      }
    }
    const int nt = (int)(c >> 3);
    if (nt) {
      const int k = cases[i];
      long long f = (long long)(pre >> 32);
      for (int t = 0; t < nt; ++t) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int e = c_mc_tri[k][3 * t + j];
          // edge -> owning node + axis.  edges 0-3 along x at (y,z) offsets, 4-7 along y at (x,z), 8-11 along z at (x,y)
          const int axis = e >> 2, q = e & 3;
          int ox = 0, oy = 0, oz = 0;
          if (axis == 0) { oy = q & 1; oz = q >> 1; }
          else if (axis == 1) { ox = q & 1; oz = q >> 1; }
          else { ox = q & 1; oy = q >> 1; }
          const long long node = i + ((long long)oz * H + oy) * W + ox;
          const uint32_t cn = code[node] & 7u;
          const uint32_t rank = __popc(cn & ((1u << axis) - 1u));
          faces[3 * f + j] = (int32_t)(voff[node] + rank);
        }
        ++f;
      }
    }
  }
};

inline int grid_for(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace

struct mp_mcubes {
  int D, H, W;
  long long n;
  uint8_t* code;
  uint8_t* cases;
  uint32_t* voff;
  unsigned long long* sums;
  unsigned long long* total;
  long long nv, nf;
  int counted;
};

extern "C" int mp_mcubes_destroy(mp_mcubes_t* h) {
  if (!h) return MP_OK;
  if (h->code) cudaFree(h->code);
  if (h->cases) cudaFree(h->cases);
  if (h->voff) cudaFree(h->voff);
  if (h->sums) cudaFree(h->sums);
  if (h->total) cudaFree(h->total);
  delete h;
  return MP_OK;
}

extern "C" int mp_mcubes_create(int D, int H, int W, mp_mcubes_t** out) {
  MP_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  MP_REQUIRE(D >= 2 && H >= 2 && W >= 2 && (long long)D * H * W < (1ll << 31), "bad volume shape %dx%dx%d", D, H, W);
  mp_mcubes* h = new mp_mcubes();
  memset(h, 0, sizeof(*h));
  h->D = D; h->H = H; h->W = W;
  h->n = (long long)D * H * W;
  cudaError_t e = cudaMalloc(&h->code, h->n);
  if (e == cudaSuccess) e = cudaMalloc(&h->cases, h->n);
  if (e == cudaSuccess) e = cudaMalloc(&h->voff, h->n * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->sums, (size_t)(mpscan::num_blocks(h->n) + 1) * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&h->total, 2 * sizeof(unsigned long long));                 // [0] total, [1] scan ticket
  if (e == cudaSuccess) e = cudaMemset(h->total, 0, 2 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    mp_set_error("mp_mcubes_create: %s", cudaGetErrorString(e));
    mp_mcubes_destroy(h);
    return MP_E_NOMEM;
  }
  *out = h;
  return MP_OK;
}

extern "C" int mp_mcubes_count(mp_mcubes_t* h, const float* vol_dev, float iso, int64_t* n_verts, int64_t* n_faces,
                               void* stream) {
  MP_REQUIRE(h && vol_dev && n_verts && n_faces, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  classify_kernel<<<grid_for(h->n), 256, 0, st>>>(vol_dev, h->code, h->cases, h->D, h->H, h->W, iso);
  MP_CUDA(cudaGetLastError());
  CountF f{h->code};
  OffsetsEmit em{h->voff};
  MP_CUDA(mpscan::scan_emit(f, em, h->n, h->sums, h->total, st));
  unsigned long long tot = 0;
  MP_CUDA(cudaMemcpyAsync(&tot, h->total, sizeof(tot), cudaMemcpyDeviceToHost, st));
  MP_CUDA(cudaStreamSynchronize(st));
  h->nv = (long long)(tot & 0xffffffffull);
  h->nf = (long long)(tot >> 32);
  h->counted = 1;
  *n_verts = h->nv;
  *n_faces = h->nf;
  return MP_OK;
}

extern "C" int mp_mcubes_emit(mp_mcubes_t* h, const float* vol_dev, float iso, float* verts_dev, int32_t* faces_dev,
                              void* stream) {
  MP_REQUIRE(h && vol_dev, "NULL argument");
  MP_REQUIRE(h->counted, "mp_mcubes_count must run first");
  if (h->nv == 0 && h->nf == 0) return MP_OK;
  MP_REQUIRE(verts_dev && faces_dev, "NULL output buffers");
  cudaStream_t st = (cudaStream_t)stream;
  CountF f{h->code};
  MeshEmit em{vol_dev, h->code, h->cases, h->voff, verts_dev, faces_dev, h->D, h->H, h->W, iso};
  // block offsets in h->sums are still valid from the count pass: only the emit launch is needed
  mpscan::emit_kernel<CountF, MeshEmit><<<mpscan::num_blocks(h->n), mpscan::kThreads, 0, st>>>(f, em, h->n, h->sums);
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}
