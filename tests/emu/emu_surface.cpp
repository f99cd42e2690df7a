// TEST INFRASTRUCTURE ONLY: the visible-surface kernels (monoport_b200/csrc/surface_kernels.cuh + mp_scan.cuh) on the CPU
// emulation layer, mirroring the launches of mp_forward_vertices in surface.cu.
//   emu_surface R dir vol.f32 out.bin      out = n x { int64 X, int64 Y, float Z, float N[3] } as four consecutive arrays
#include "cuda_emu.h"

#include "../../monoport_b200/csrc/surface_kernels.cuh"

using namespace surface_k;

int main(int argc, char** argv) {
  if (argc != 5) { fprintf(stderr, "usage: emu_surface R dir vol.f32 out.bin\n"); return 2; }
  const int R = atoi(argv[1]), dir = atoi(argv[2]);
  const long long V = (long long)R * R * R, n = (long long)R * R;
  std::vector<float> vol(V);
  FILE* f = fopen(argv[3], "rb");
  if (!f || fread(vol.data(), 4, V, f) != (size_t)V) { perror("vol"); return 2; }
  fclose(f);
  std::vector<int32_t> first_t(n, -12345);
  const int nb = mpscan::num_blocks(n);
  std::vector<unsigned long long> sums(nb + 1, 0);
  unsigned long long total[2] = {0, 0};
  std::vector<long long> X(n, -1), Y(n, -1);
  std::vector<float> Z(n, -1.f), N(3 * n, -1.f);
  const float* vp = vol.data();
  int32_t* ft = first_t.data();
  cuda_emu::launch(dim3((unsigned)std::min<long long>((n + kHitCols - 1) / kHitCols, 37)), dim3(kHitCols * kHitSlices), [&] { first_hit_kernel(vp, R, dir, ft); });
  HitF hf{ft};
  HitEmit em{vp, ft, R, dir, X.data(), Y.data(), Z.data(), N.data()};
  cuda_emu::launch(dim3(nb), dim3(mpscan::kThreads),
                   [&] { mpscan::block_sums_kernel<HitF, mpscan::NoPost>(hf, n, sums.data(), nb, total, mpscan::NoPost()); });
  cuda_emu::launch(dim3(nb), dim3(mpscan::kThreads), [&] { mpscan::emit_kernel<HitF, HitEmit>(hf, em, n, sums.data()); });
  const long long k = (long long)total[0];
  f = fopen(argv[4], "wb");
  fwrite(X.data(), 8, k, f);
  fwrite(Y.data(), 8, k, f);
  fwrite(Z.data(), 4, k, f);
  fwrite(N.data(), 4, 3 * k, f);
  fclose(f);
  printf("%lld\n", k);
  return 0;
}
