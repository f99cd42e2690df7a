// TEST INFRASTRUCTURE ONLY: runs the marching-cubes kernels of monoport_b200/csrc/mcubes_kernels.cuh (+ the ordered scan
// of mp_scan.cuh) on the CPU emulation layer, mirroring the launches of mp_mcubes_count / mp_mcubes_emit in mcubes.cu.
//   emu_mcubes D H W iso in.f32 out_verts.f32 out_faces.i32     (prints "nv nf")
#include "cuda_emu.h"

#include "../../monoport_b200/csrc/mcubes_kernels.cuh"

#include <string>

using namespace mcubes;

static std::vector<char> read_file(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> b(n);
  if (fread(b.data(), 1, n, f) != (size_t)n) { perror("read"); exit(2); }
  fclose(f);
  return b;
}

static void write_file(const char* path, const void* p, size_t bytes) {
  FILE* f = fopen(path, "wb");
  if (!f) { perror(path); exit(2); }
  if (bytes && fwrite(p, 1, bytes, f) != bytes) { perror("write"); exit(2); }
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc != 8) { fprintf(stderr, "usage: emu_mcubes D H W iso in.f32 verts.f32 faces.i32\n"); return 2; }
  const int D = atoi(argv[1]), H = atoi(argv[2]), W = atoi(argv[3]);
  const float iso = (float)atof(argv[4]);
  const long long n = (long long)D * H * W;
  std::vector<char> raw = read_file(argv[5]);
  if ((long long)raw.size() != n * 4) { fprintf(stderr, "volume size mismatch\n"); return 2; }
  const float* vol = reinterpret_cast<const float*>(raw.data());

  // workspace of mp_mcubes_create (guard words behind every array)
  const long long n_words = (n + 31) >> 5;
  const long long pad = bits_pad_words(H, W);
  std::vector<uint32_t> bits(n_words + pad, 0u);
  std::vector<WordInfo> info(n_words + 4, WordInfo{0xEEEEEEEEu, 0xEEEEEEEEu, 0xEEEEEEEEu, 0xEEEEEEEEu});
  std::vector<unsigned long long> prefix(n_words + 4, 0xDEADBEEFDEADBEEFull);
  const int nb = mpscan::num_blocks(n_words);
  std::vector<unsigned long long> sums(nb + 1, 0);
  unsigned long long total[2] = {0, 0};

  // ---- mp_mcubes_count
  {
    const long long warps_needed = (n_words + kBitsUnroll - 1) / kBitsUnroll;
    long long blocks = (warps_needed + kBitsThreads / 32 - 1) / (kBitsThreads / 32);
    if (blocks > 3) blocks = 3;                       // force the grid-stride loop
    cuda_emu::launch(dim3((unsigned)blocks), dim3(kBitsThreads), [&] { bits_kernel(vol, bits.data(), n, iso); });
  }
  for (long long w = n_words; w < n_words + pad; ++w)
    if (bits[w] != 0u) { fprintf(stderr, "wrote into the bit padding\n"); return 3; }
  
  std::vector<uint32_t> active(n_words + 4, 0xABABABABu);
  uint32_t n_active = 0;
  WordCountF f{info.data()};
  PrefixEmit em{prefix.data(), active.data(), &n_active};
  std::fill(sums.begin(), sums.end(), 0ull);            // (mp_mcubes_count zeroes the chunk totals: the CTAs accumulate)
  cuda_emu::launch(dim3((unsigned)((n_words + kClassifyWords - 1) / kClassifyWords)), dim3(kClassifyThreads),
                   [&] { classify_sums_kernel(bits.data(), info.data(), n, D, H, W, sums.data(), nb, total); });
  cuda_emu::launch(dim3(nb), dim3(mpscan::kThreads),
                   [&] { mpscan::emit_kernel<WordCountF, PrefixEmit>(f, em, n_words, sums.data()); });
  if (total[1] != 0) { fprintf(stderr, "scan ticket not reset\n"); return 3; }
  const long long nv = (long long)(total[0] & 0xffffffffull), nf = (long long)(total[0] >> 32);

  // ---- mp_mcubes_emit
  std::vector<float> verts((size_t)nv * 3 + 1, -12345.f);
  std::vector<int32_t> faces((size_t)nf * 3 + 1, -7);
  if (nv || nf) {
    long long blocks = ((long long)n_active + kEmitThreads / 32 - 1) / (kEmitThreads / 32);
    if (blocks > 2) blocks = 2;                       // force the grid-stride loop
    if (blocks > 0)
      cuda_emu::launch(dim3((unsigned)blocks), dim3(kEmitThreads), [&] {
        mesh_emit_kernel(vol, bits.data(), info.data(), prefix.data(), active.data(), n_active, verts.data(), faces.data(), D, H, W, n, iso);
      });
  }
  for (long long w = n_words; w < n_words + 4; ++w)
    if (active[w] != 0xABABABABu) { fprintf(stderr, "wrote past the active list\n"); return 3; }
  if (verts[(size_t)nv * 3] != -12345.f || faces[(size_t)nf * 3] != -7) { fprintf(stderr, "wrote past the outputs\n"); return 3; }
  for (long long w = n_words; w < n_words + 4; ++w)
    if (info[w].ex != 0xEEEEEEEEu || prefix[w] != 0xDEADBEEFDEADBEEFull) { fprintf(stderr, "wrote past the word arrays\n"); return 3; }
  for (long long k = 0; k < nf * 3; ++k)
    if (faces[k] < 0 || faces[k] >= nv) { fprintf(stderr, "face index out of range\n"); return 3; }
  write_file(argv[6], verts.data(), (size_t)nv * 3 * sizeof(float));
  write_file(argv[7], faces.data(), (size_t)nf * 3 * sizeof(int32_t));
  printf("%lld %lld\n", nv, nf);
  return 0;
}
