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

  // workspace of mp_mcubes_create; the offset array is poisoned: only entries of vertex-owning nodes may be read
  std::vector<uint8_t> code(n + 64, 0xEE), cases(n + 64, 0xEE);
  std::vector<uint32_t> voff(n, 0xDEADBEEFu);
  const int nb = mpscan::num_blocks(n);
  std::vector<unsigned long long> sums(nb + 1, 0);
  unsigned long long total[2] = {0, 0};

  // ---- mp_mcubes_count
  const bool fast = getenv("MONOPORT_B200_MC_FAST") && atoi(getenv("MONOPORT_B200_MC_FAST"));
  cuda_emu::launch(dim3((unsigned)D, (unsigned)((H + kClassRows - 1) / kClassRows)), dim3(32, kClassRows), [&] {
    if (fast) classify_fast_kernel(vol, code.data(), cases.data(), D, H, W, iso);
    else classify_kernel(vol, code.data(), cases.data(), D, H, W, iso);
  });
  CountF f{code.data()};
  OffsetsEmit em{voff.data()};
  cuda_emu::launch(dim3(nb), dim3(mpscan::kThreads),
                   [&] { mpscan::block_sums_kernel<CountF, mpscan::NoPost>(f, n, sums.data(), nb, total, mpscan::NoPost()); });
  cuda_emu::launch(dim3(nb), dim3(mpscan::kThreads),
                   [&] { mpscan::emit_kernel<CountF, OffsetsEmit>(f, em, n, sums.data()); });
  if (total[1] != 0) { fprintf(stderr, "scan ticket not reset\n"); return 3; }
  const long long nv = (long long)(total[0] & 0xffffffffull), nf = (long long)(total[0] >> 32);

  // ---- mp_mcubes_emit
  std::vector<float> verts((size_t)nv * 3 + 1, -12345.f);
  std::vector<int32_t> faces((size_t)nf * 3 + 1, -7);
  if (nv || nf)
    cuda_emu::launch(dim3(nb), dim3(mpscan::kThreads), [&] {
      mesh_emit_kernel(vol, code.data(), cases.data(), voff.data(), sums.data(), verts.data(), faces.data(), H, W, n, iso);
    });
  if (verts[(size_t)nv * 3] != -12345.f || faces[(size_t)nf * 3] != -7) { fprintf(stderr, "wrote past the outputs\n"); return 3; }
  for (long long i = n; i < n + 64; ++i)
    if (code[i] != 0xEE || cases[i] != 0xEE) { fprintf(stderr, "wrote past the code volume\n"); return 3; }
  write_file(argv[6], verts.data(), (size_t)nv * 3 * sizeof(float));
  write_file(argv[7], faces.data(), (size_t)nf * 3 * sizeof(int32_t));
  printf("%lld %lld\n", nv, nf);
  return 0;
}
