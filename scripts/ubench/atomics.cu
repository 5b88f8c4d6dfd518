// atomics.cu - how fast are 1 M scattered 64-bit atomics, by target and flavour?  (DESIGN.md 4, K1 of the merge)
//   dense   a u64[rows] side array (20 MB at 2.5 M rows)
//   rows    the last 16-byte chunk of 128-byte table rows (320 MB)
//   ret/red with / without using the returned value (ATOMG vs REDG)
//   touch   a plain-load pass over the same words first (then the atomics hit L2)
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <bool RET, int ILP>
__global__ void __launch_bounds__(256) k_atom(const uint32_t* keys, uint32_t n, unsigned long long* base, uint32_t stride, uint32_t* rank) {
  const uint32_t i0 = blockIdx.x * (256 * ILP) + threadIdx.x;
  uint32_t k[ILP];
  unsigned long long r[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) k[j] = i0 + j * 256 < n ? keys[i0 + j * 256] : 0xFFFFFFFFu;
#pragma unroll
  for (int j = 0; j < ILP; ++j) {
    r[j] = 0;
    if (k[j] != 0xFFFFFFFFu) {
      unsigned long long* p = base + (size_t)k[j] * stride;
      const unsigned long long add = ((unsigned long long)(i0 + j * 256) << 32) | 8ull;
      if (RET) r[j] = atomicAdd(p, add);
      else atomicAdd(p, add);
    }
  }
  if (RET) {
#pragma unroll
    for (int j = 0; j < ILP; ++j)
      if (i0 + j * 256 < n) rank[i0 + j * 256] = (uint32_t)r[j];
  }
}
template <int ILP>
__global__ void __launch_bounds__(256) k_touch(const uint32_t* keys, uint32_t n, const unsigned long long* base, uint32_t stride, uint32_t* sink) {
  const uint32_t i0 = blockIdx.x * (256 * ILP) + threadIdx.x;
  uint32_t k[ILP], acc = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) k[j] = i0 + j * 256 < n ? keys[i0 + j * 256] : 0xFFFFFFFFu;
#pragma unroll
  for (int j = 0; j < ILP; ++j)
    if (k[j] != 0xFFFFFFFFu) acc ^= (uint32_t)__ldcg(base + (size_t)k[j] * stride);
  if (acc == 0x12345u) sink[0] = acc;
}

template <class F>
float time_it(F f, void* flush, size_t fb) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float sum = 0;
  for (int r = 0; r < 7; ++r) {
    CK(cudaMemsetAsync(flush, r, fb));
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (r >= 2) sum += ms;
  }
  CK(cudaGetLastError());
  return sum / 5;
}

int main(int argc, char** argv) {
  const uint64_t rows = argc > 1 ? strtoull(argv[1], 0, 10) : 2500000ull;
  const uint32_t n = 1000000;
  std::mt19937_64 rng(7);
  std::vector<uint32_t> upd(n);
  for (auto& u : upd) u = (uint32_t)(rng() % rows);
  uint32_t *keys, *rank; unsigned long long *dense, *table; void* flush;
  CK(cudaMalloc(&keys, n * 4)); CK(cudaMemcpy(keys, upd.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&rank, n * 4));
  CK(cudaMalloc(&dense, rows * 8)); CK(cudaMemset(dense, 0, rows * 8));
  CK(cudaMalloc(&table, rows * 128)); CK(cudaMemset(table, 0, rows * 128));
  const size_t fb = 256u << 20; CK(cudaMalloc(&flush, fb));
  printf("%llu rows, %u scattered 64-bit atomics\n", (unsigned long long)rows, n);
  auto rep = [&](const char* name, float ms) { printf("  %-58s %8.1f us\n", name, ms * 1e3); };
  unsigned long long* rw = table + 14;  // word at byte 112 of each row
#define G(ILP) ((n + 256 * ILP - 1) / (256 * ILP))
  rep("dense u64, returning, ILP 4", time_it([&] { k_atom<true, 4><<<G(4), 256>>>(keys, n, dense, 1, rank); }, flush, fb));
  rep("dense u64, returning, ILP 1", time_it([&] { k_atom<true, 1><<<G(1), 256>>>(keys, n, dense, 1, rank); }, flush, fb));
  rep("dense u64, returning, ILP 8", time_it([&] { k_atom<true, 8><<<G(8), 256>>>(keys, n, dense, 1, rank); }, flush, fb));
  rep("dense u64, no return (RED), ILP 4", time_it([&] { k_atom<false, 4><<<G(4), 256>>>(keys, n, dense, 1, rank); }, flush, fb));
  rep("row word, returning, ILP 4", time_it([&] { k_atom<true, 4><<<G(4), 256>>>(keys, n, rw, 16, rank); }, flush, fb));
  rep("row word, returning, ILP 1", time_it([&] { k_atom<true, 1><<<G(1), 256>>>(keys, n, rw, 16, rank); }, flush, fb));
  rep("row word, no return (RED), ILP 4", time_it([&] { k_atom<false, 4><<<G(4), 256>>>(keys, n, rw, 16, rank); }, flush, fb));
  rep("row word: touch pass only, ILP 4", time_it([&] { k_touch<4><<<G(4), 256>>>(keys, n, rw, 16, rank); }, flush, fb));
  rep("row word: touch pass only, ILP 8", time_it([&] { k_touch<8><<<G(8), 256>>>(keys, n, rw, 16, rank); }, flush, fb));
  rep("row word: touch kernel + returning atomics kernel", time_it([&] { k_touch<4><<<G(4), 256>>>(keys, n, rw, 16, rank); k_atom<true, 4><<<G(4), 256>>>(keys, n, rw, 16, rank); }, flush, fb));
  rep("dense u64: touch kernel + returning atomics kernel", time_it([&] { k_touch<4><<<G(4), 256>>>(keys, n, dense, 1, rank); k_atom<true, 4><<<G(4), 256>>>(keys, n, dense, 1, rank); }, flush, fb));
  // warm: no flush between runs (a table that stays in L2 between batches)
  {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int r = 0; r < 3; ++r) k_atom<true, 4><<<G(4), 256>>>(keys, n, dense, 1, rank);
    CK(cudaEventRecord(a));
    for (int r = 0; r < 5; ++r) k_atom<true, 4><<<G(4), 256>>>(keys, n, dense, 1, rank);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    rep("dense u64, returning, ILP 4, L2-warm (no flush)", ms / 5);
  }
  printf("done\n");
  return 0;
}
