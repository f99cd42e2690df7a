// TEST INFRASTRUCTURE ONLY: the ordered scan of monoport_b200/csrc/mp_scan.cuh on the CPU emulation layer.
//   emu_scan n vec in.u8 out_prefix.u64      (prints "total_lo total_hi")
// Element value = (byte & 7) | ((byte >> 3) << 32): two packed 32-bit counters, like marching cubes' (vertices, triangles).
#include "cuda_emu.h"

#include "../../monoport_b200/csrc/mp_scan.cuh"

template <bool VEC>
struct ByteF {
  const uint8_t* b;
  static constexpr bool kVec8 = VEC;
  static unsigned long long val(uint32_t c) { return (unsigned long long)(c & 7u) | ((unsigned long long)(c >> 3) << 32); }
  unsigned long long operator()(long long i) const { return val(b[i]); }
  void load8(long long i, unsigned long long (&v)[8]) const {
    uint32_t t[8];
    mpscan::load_bytes8(b, i, t);
    for (int j = 0; j < 8; ++j) v[j] = val(t[j]);
  }
};
struct StoreEmit {
  unsigned long long* out;
  void operator()(long long i, unsigned long long, unsigned long long pre) const { out[i] = pre; }
};
struct Post {
  unsigned long long* seen;
  void operator()(unsigned long long t) const { *seen = t; }
};

template <bool VEC>
static int run(long long n, const uint8_t* data, unsigned long long* out, unsigned long long* total) {
  const int nb = mpscan::num_blocks(n > 0 ? n : 1);
  std::vector<unsigned long long> sums(nb + 1, 0xABABABABull);
  unsigned long long seen = ~0ull;
  ByteF<VEC> f{data};
  cuda_emu::launch(dim3(nb), dim3(mpscan::kThreads),
                   [&] { mpscan::block_sums_kernel<ByteF<VEC>, Post>(f, n, sums.data(), nb, total, Post{&seen}); });
  if (n > 0)
    cuda_emu::launch(dim3(nb), dim3(mpscan::kThreads),
                     [&] { mpscan::emit_kernel<ByteF<VEC>, StoreEmit>(f, StoreEmit{out}, n, sums.data()); });
  if (total[1] != 0 || seen != total[0]) { fprintf(stderr, "ticket / post hook wrong\n"); return 3; }
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 5) { fprintf(stderr, "usage: emu_scan n vec in.u8 out.u64\n"); return 2; }
  const long long n = atoll(argv[1]);
  const int vec = atoi(argv[2]);
  std::vector<uint8_t> data((size_t)n + 8, 0xFF);       // poisoned tail: the vector path must not run past n
  FILE* f = fopen(argv[3], "rb");
  if (!f || (n && fread(data.data(), 1, n, f) != (size_t)n)) { perror("read"); return 2; }
  fclose(f);
  // 8-byte aligned base like the cudaMalloc'ed volumes
  std::vector<unsigned long long> out((size_t)n + 1, 0);
  unsigned long long total[2] = {0, 0};
  const int rc = vec ? run<true>(n, data.data(), out.data(), total) : run<false>(n, data.data(), out.data(), total);
  if (rc) return rc;
  f = fopen(argv[4], "wb");
  if (n) fwrite(out.data(), 8, n, f);
  fclose(f);
  printf("%llu %llu\n", total[0] & 0xffffffffull, total[0] >> 32);
  return 0;
}
