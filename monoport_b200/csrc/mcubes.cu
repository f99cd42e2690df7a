// F3: marching cubes on the occupancy volume (absent from the reference -- SURVEY.md finding 3; PARITY UNPINNED,
// bit-exact against oracle/spec.py:marching_cubes_ref which uses the same derived table, tools/gen_mc_table.py).
// HBM-bound byte/integer work.  Indexed, watertight mesh with deterministic ids:
//   vertex id = rank of (node, axis) among active owned edges (every grid edge is owned by its lower node);
//   face order = (cell linear index, table order).
// count:  bits_kernel (THE read of the volume: 4 B per node in, one occupancy bit per node out, streaming) ->
//         ordered scan over the 32-node words of the bit volume, (verts, tris) packed in one uint64.  Its first pass
//         (classify_sums_kernel) CLASSIFIES each word on the way (edge masks + triangle count; words with no surface inside
//         leave after a dozen cached loads), its second pass leaves the exclusive prefix per word, the totals, and an
//         unordered list of the words with a surface inside.
// emit :  mesh_emit_kernel, a warp per listed word, a lane per node: vertices from the volume (two loads per vertex); the
//         cells' triangles go through a shared-memory queue and are resolved one thread per face corner (edge -> owning
//         node -> id = the owning word's prefix + popcounts).
// Workspace: 1 bit + 24 B per 32 nodes (15 MB at 257^3; the round-1 byte-per-node design took 102 MB).
#include "mp_common.cuh"
#include <stdlib.h>
#include "mcubes_kernels.cuh"

using namespace mcubes;


struct mp_mcubes {
  int D, H, W;
  long long n, n_words;
  uint32_t* bits;
  WordInfo* info;
  unsigned long long* prefix;
  uint32_t* active;            // words with a surface inside (unordered), filled by the scan's emit half
  uint32_t* n_active_dev;
  uint32_t n_active;
  unsigned long long* sums;
  unsigned long long* total;
  long long nv, nf;
  int counted;
};

extern "C" int mp_mcubes_destroy(mp_mcubes_t* h) {
  if (!h) return MP_OK;
  if (h->bits) cudaFree(h->bits);
  if (h->info) cudaFree(h->info);
  if (h->prefix) cudaFree(h->prefix);
  if (h->active) cudaFree(h->active);
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
  h->n_words = (h->n + 31) >> 5;
  const size_t bit_words = (size_t)(h->n_words + bits_pad_words(H, W));
  cudaError_t e = cudaMalloc(&h->bits, bit_words * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->bits, 0, bit_words * sizeof(uint32_t));     // the padding stays zero for ever
  if (e == cudaSuccess) e = cudaMalloc(&h->info, (size_t)h->n_words * sizeof(WordInfo));
  if (e == cudaSuccess) e = cudaMalloc(&h->prefix, (size_t)h->n_words * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&h->active, (size_t)h->n_words * sizeof(uint32_t));
  // chunk totals of the scan + (last entry) the length of the active list: zeroed together before every count
  if (e == cudaSuccess) e = cudaMalloc(&h->sums, (size_t)(mpscan::num_blocks(h->n_words) + 2) * sizeof(unsigned long long));
  if (e == cudaSuccess) h->n_active_dev = reinterpret_cast<uint32_t*>(h->sums + mpscan::num_blocks(h->n_words) + 1);
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

static int sm_count() {
  static const int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  return sms;
}

extern "C" int mp_mcubes_count(mp_mcubes_t* h, const float* vol_dev, float iso, int64_t* n_verts, int64_t* n_faces,
                               void* stream) {
  MP_REQUIRE(h && vol_dev && n_verts && n_faces, "NULL argument");
  MpRange nvtx("monoport_b200: F3 marching cubes (count)");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // a multiple of the SM count, 8 CTAs of 256 threads resident per SM: 8 x 128 B loads in flight per warp
    const long long warps_needed = (h->n_words + kBitsUnroll - 1) / kBitsUnroll;
    const long long blocks_needed = (warps_needed + kBitsThreads / 32 - 1) / (kBitsThreads / 32);
    const long long cap = (long long)sm_count() * 8;
    bits_kernel<<<(unsigned)(blocks_needed < cap ? blocks_needed : cap), kBitsThreads, 0, st>>>(vol_dev, h->bits, h->n, iso);
  }
  MP_CUDA(cudaGetLastError());
  // the ordered scan over the words, first pass = classify_sums_kernel (classification + chunk totals), second pass = the
  // generic emit half reading the stored info
  const int nb = mpscan::num_blocks(h->n_words);
  MP_CUDA(cudaMemsetAsync(h->sums, 0, (size_t)(nb + 2) * sizeof(unsigned long long), st));
  const unsigned n_ctas = (unsigned)((h->n_words + kClassifyWords - 1) / kClassifyWords);
  classify_sums_kernel<<<n_ctas, kClassifyThreads, 0, st>>>(h->bits, h->info, h->n, h->D, h->H, h->W, h->sums, nb, h->total);
  WordCountF f2{h->info};
  PrefixEmit em{h->prefix, h->active, h->n_active_dev};
  mpscan::emit_kernel<WordCountF, PrefixEmit><<<nb, mpscan::kThreads, 0, st>>>(f2, em, h->n_words, h->sums);
  MP_CUDA(cudaGetLastError());
  unsigned long long tot = 0;
  MP_CUDA(cudaMemcpyAsync(&tot, h->total, sizeof(tot), cudaMemcpyDeviceToHost, st));
  MP_CUDA(cudaMemcpyAsync(&h->n_active, h->n_active_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
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
  MpRange nvtx("monoport_b200: F3 marching cubes (emit)");
  cudaStream_t st = (cudaStream_t)stream;
  // one warp per word with a surface inside (the active list the count pass left behind)
  const long long blocks_needed = ((long long)h->n_active + kEmitThreads / 32 - 1) / (kEmitThreads / 32);
  if (blocks_needed > 0)
    mesh_emit_kernel<<<(unsigned)blocks_needed, kEmitThreads, 0, st>>>(vol_dev, h->bits, h->info, h->prefix, h->active, h->n_active,
                                                                     verts_dev, faces_dev, h->D, h->H, h->W, h->n, iso);
  MP_CUDA(cudaGetLastError());
  return MP_OK;
}
