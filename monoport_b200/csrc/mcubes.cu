// F3: marching cubes on the occupancy volume (absent from the reference -- SURVEY.md finding 3; PARITY UNPINNED,
// bit-exact against oracle/spec.py:marching_cubes_ref which uses the same derived table, tools/gen_mc_table.py).
// HBM-bound byte/integer work.  Indexed, watertight mesh with deterministic ids:
//   vertex id = rank of (node, axis) among active owned edges (every grid edge is owned by its lower node);
//   face order = (cell linear index, table order).
// count:  classify_kernel (1 read of the volume, rows walked by warps, no divisions: per node 3 owned-edge bits + the
//         triangle count of its cell packed in one byte, + the cell's case byte)  ->  ordered scan of (verts, tris)
//         packed in one uint64 (8-byte loads of the code bytes); the emit half of the scan stores the exclusive vertex
//         offset of the nodes that own a vertex (sparse: faces look up nothing else)  -> totals.
// emit :  mesh_emit_kernel re-runs the block-local scan, queues the block's active nodes in shared memory and spreads
//         their up to 3 vertices + 15 face corners over the whole CTA (one short dependent chain per thread instead of
//         one thread walking a whole cell).
#include "mp_common.cuh"
#include <stdlib.h>
#include "mcubes_kernels.cuh"

using namespace mcubes;


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
  MP_REQUIRE(D >= 2 && H >= 2 && W >= 2 && (long long)D * H * W < (1ll << 31) && H <= kClassRows * 65535, "bad volume shape %dx%dx%d", D, H, W);
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
  static const int fast = [] { const char* v = getenv("MONOPORT_B200_MC_FAST"); return v ? atoi(v) : 1; }();   // measured 59.6 vs 71.0 us at 257^3 (profiles/r02_call1_*)
  const dim3 cgrid((unsigned)h->D, (unsigned)((h->H + kClassRows - 1) / kClassRows)), cblock(32, kClassRows);
  if (fast) classify_fast_kernel<<<cgrid, cblock, 0, st>>>(vol_dev, h->code, h->cases, h->D, h->H, h->W, iso);
  else classify_kernel<<<cgrid, cblock, 0, st>>>(vol_dev, h->code, h->cases, h->D, h->H, h->W, iso);
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
  // block offsets in h->sums are still valid from the count pass: only the emit launch is needed
  mesh_emit_kernel<<<mpscan::num_blocks(h->n), mpscan::kThreads, 0, st>>>(vol_dev, h->code, h->cases, h->voff, h->sums, verts_dev,
                                                                         faces_dev, h->H, h->W, h->n, iso);
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}
