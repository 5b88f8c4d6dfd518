// row_gather.cu - micro-benchmark behind the merge kernel's staging design:
// how fast can 128-byte table rows at RANDOM row indices be brought into shared memory and
// written back, by which mechanism?
//   v0  cp.async 16 B x 8 lanes per row (LDGSTS), 8-lane 16-byte stores back      (round-1 kernel)
//   v1  one cp.async.bulk (UBLKCP, the TMA engine) per row + mbarrier, bulk store back
//   v2  8 x LDG.128 per thread into registers, 8 x STG.128 back
//   v3  like v1 for the loads, 8-lane 16-byte stores back
// plus the count kernel alternatives (atomicAdd only / atomicAdd + atomicExch list).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o row_gather row_gather.cu
// Run:   ./row_gather [rows_in_table] [rows_touched]
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

constexpr int ROW_Q = 8;

__device__ __forceinline__ void cp_async16(void* s, const void* g) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* s, const void* g, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(s)),
               "l"(g), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* g, const void* s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"((uint32_t)__cvta_generic_to_shared(s)),
               "r"(bytes)
               : "memory");
}

template <int T>
__global__ void __launch_bounds__(T) k_v0(uint4* table, const uint32_t* keys, uint32_t n) {
  __shared__ __align__(16) uint4 s_row[T * ROW_Q];
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const uint32_t i = blockIdx.x * T + tid;
  const uint32_t key = i < n ? keys[i] : 0;
  const uint32_t vmask = __ballot_sync(0xffffffffu, i < n);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = j * 4 + (lane >> 3), c = lane & 7;
    const uint32_t ek = __shfl_sync(0xffffffffu, key, e);
    if ((vmask >> e) & 1u) cp_async16(&s_row[(wbase + e) * ROW_Q + (c ^ (e & 7))], table + (uint64_t)ek * ROW_Q + c);
  }
  cp_async_wait_all();
  __syncwarp();
  if (i < n) {  // touch the row: one thread per row, like the resolver
    uint4 q = s_row[tid * ROW_Q + (0 ^ (tid & 7))];
    q.x += 1;
    s_row[tid * ROW_Q + (0 ^ (tid & 7))] = q;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = j * 4 + (lane >> 3), c = lane & 7;
    const uint32_t ek = __shfl_sync(0xffffffffu, key, e);
    if ((vmask >> e) & 1u) table[(uint64_t)ek * ROW_Q + c] = s_row[(wbase + e) * ROW_Q + (c ^ (e & 7))];
  }
}

// STRIDE_Q = 9: rows 144 bytes apart in shared memory (conflict-free for one-thread-per-row 16-byte accesses)
template <int T, bool BULK_STORE>
__global__ void __launch_bounds__(T) k_v1(uint4* table, const uint32_t* keys, uint32_t n) {
  constexpr int SQ = 9;
  __shared__ __align__(128) uint4 s_row[T * SQ];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const uint32_t i = blockIdx.x * T + tid;
  if (tid == 0) mbar_init(&bar, T);
  const uint32_t key = i < n ? keys[i] : 0;
  __syncthreads();
  if (i < n) {
    mbar_arrive_tx(&bar, 128);
    bulk_g2s(&s_row[tid * SQ], table + (uint64_t)key * ROW_Q, 128, &bar);
  } else {
    mbar_arrive(&bar);
  }
  mbar_wait(&bar, 0);
  if (i < n) {
    uint4 q = s_row[tid * SQ];
    q.x += 1;
    s_row[tid * SQ] = q;
  }
  if (BULK_STORE) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (i < n) {
      bulk_s2g(table + (uint64_t)key * ROW_Q, &s_row[tid * SQ], 128);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    __syncwarp();
    const uint32_t vmask = __ballot_sync(0xffffffffu, i < n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 4 + (lane >> 3), c = lane & 7;
      const uint32_t ek = __shfl_sync(0xffffffffu, key, e);
      if ((vmask >> e) & 1u) table[(uint64_t)ek * ROW_Q + c] = s_row[(wbase + e) * SQ + c];
    }
  }
}

template <int T>
__global__ void __launch_bounds__(T) k_v2(uint4* table, const uint32_t* keys, uint32_t n) {
  const uint32_t i = blockIdx.x * T + threadIdx.x;
  if (i >= n) return;
  const uint32_t key = keys[i];
  uint4 q[ROW_Q];
  uint4* row = table + (uint64_t)key * ROW_Q;
#pragma unroll
  for (int c = 0; c < ROW_Q; ++c) q[c] = row[c];
  q[0].x += 1;
#pragma unroll
  for (int c = 0; c < ROW_Q; ++c) row[c] = q[c];
}

// count kernels: cnt only vs cnt + linked list
__global__ void __launch_bounds__(256) k_count(const uint32_t* keys, uint32_t n, uint32_t* cnt, uint32_t* rank) {
  const uint32_t i0 = blockIdx.x * 1024 + threadIdx.x;
  uint32_t k[4], r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) k[j] = i0 + j * 256 < n ? keys[i0 + j * 256] : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = i0 + j * 256 < n ? atomicAdd(&cnt[k[j]], 1u) : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j * 256 < n) rank[i0 + j * 256] = r[j];
}
__global__ void __launch_bounds__(256) k_count_list(const uint32_t* keys, uint32_t n, uint32_t* cnt, uint32_t* last, uint32_t* rank,
                                                    uint32_t* prev) {
  const uint32_t i0 = blockIdx.x * 1024 + threadIdx.x;
  uint32_t k[4], r[4], p[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) k[j] = i0 + j * 256 < n ? keys[i0 + j * 256] : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r[j] = i0 + j * 256 < n ? atomicAdd(&cnt[k[j]], 1u) : 0;
    p[j] = i0 + j * 256 < n ? atomicExch(&last[k[j]], i0 + j * 256) : 0;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j * 256 < n) {
      rank[i0 + j * 256] = r[j];
      prev[i0 + j * 256] = p[j];
    }
}
// packed 64-bit: one atomicAdd gives (sum of arrival indices mod 2^32 | count) - enough to find the partner of a 2-update path
__global__ void __launch_bounds__(256) k_count64(const uint32_t* keys, uint32_t n, unsigned long long* cnt, uint32_t* rank) {
  const uint32_t i0 = blockIdx.x * 1024 + threadIdx.x;
  uint32_t k[4];
  unsigned long long r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) k[j] = i0 + j * 256 < n ? keys[i0 + j * 256] : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = i0 + j * 256 < n ? atomicAdd(&cnt[k[j]], ((unsigned long long)(i0 + j * 256) << 32) | 1ull) : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j * 256 < n) rank[i0 + j * 256] = (uint32_t)r[j];
}

template <class F>
float time_it(F f, int reps, void* flush, size_t flush_bytes) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e9f, sum = 0;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaMemsetAsync(flush, r, flush_bytes));  // > L2: nothing of the table is cached
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (r >= 2) {
      best = std::min(best, ms);
      sum += ms;
    }
  }
  CK(cudaGetLastError());
  return sum / reps;
}

int main(int argc, char** argv) {
  const uint64_t rows = argc > 1 ? strtoull(argv[1], 0, 10) : 2500000ull;
  const uint32_t n = argc > 2 ? (uint32_t)strtoul(argv[2], 0, 10) : 824000u;
  uint4* table;
  CK(cudaMalloc(&table, rows * 128));
  CK(cudaMemset(table, 0, rows * 128));
  // distinct random rows (a batch's distinct paths)
  std::vector<uint32_t> all(rows);
  std::iota(all.begin(), all.end(), 0u);
  std::mt19937_64 rng(1234);
  for (uint32_t i = 0; i < n; ++i) std::swap(all[i], all[i + rng() % (rows - i)]);
  uint32_t* keys;
  CK(cudaMalloc(&keys, n * 4));
  CK(cudaMemcpy(keys, all.data(), n * 4, cudaMemcpyHostToDevice));
  const size_t flush_bytes = 256u << 20;
  void* flush;
  CK(cudaMalloc(&flush, flush_bytes));
  const double mb = n * 256.0 / 1e6;
  printf("table %llu rows (%.0f MB), %u random distinct rows touched: %.1f MB read+write\n", (unsigned long long)rows, rows * 128 / 1e6, n, mb);
  auto rep = [&](const char* name, float ms) { printf("  %-44s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, mb / ms); };
#define RUN(name, kern, T)                                                   \
  rep(name, time_it([&] { kern<<<(n + T - 1) / T, T>>>(table, keys, n); }, 10, flush, flush_bytes))
  CK(cudaFuncSetAttribute(k_v0<128>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(k_v0<256>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute((k_v1<128, true>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute((k_v1<256, true>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute((k_v1<128, false>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute((k_v1<256, false>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  RUN("v0 cp.async 8 lanes/row, T=128", k_v0<128>, 128);
  RUN("v0 cp.async 8 lanes/row, T=256", k_v0<256>, 256);
  RUN("v1 cp.async.bulk/row load+store, T=128", (k_v1<128, true>), 128);
  RUN("v1 cp.async.bulk/row load+store, T=256", (k_v1<256, true>), 256);
  RUN("v3 cp.async.bulk load, 8-lane store, T=128", (k_v1<128, false>), 128);
  RUN("v3 cp.async.bulk load, 8-lane store, T=256", (k_v1<256, false>), 256);
  RUN("v2 LDG.128 x8 / STG.128 x8 per thread, T=128", k_v2<128>, 128);
  RUN("v2 LDG.128 x8 / STG.128 x8 per thread, T=256", k_v2<256>, 256);

  // count kernels over a 1 M-update batch with duplicates (uniform over the table)
  const uint32_t nb = 1000000;
  std::vector<uint32_t> upd(nb);
  for (auto& u : upd) u = (uint32_t)(rng() % rows);
  uint32_t *d_upd, *cnt, *last, *rank, *prev;
  unsigned long long* cnt64;
  CK(cudaMalloc(&d_upd, nb * 4));
  CK(cudaMemcpy(d_upd, upd.data(), nb * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cnt, rows * 4));
  CK(cudaMalloc(&last, rows * 4));
  CK(cudaMalloc(&cnt64, rows * 8));
  CK(cudaMalloc(&rank, nb * 4));
  CK(cudaMalloc(&prev, nb * 4));
  CK(cudaMemset(cnt, 0, rows * 4));
  CK(cudaMemset(last, 0xFF, rows * 4));
  CK(cudaMemset(cnt64, 0, rows * 8));
  auto rep2 = [&](const char* name, float ms) { printf("  %-44s %8.1f us\n", name, ms * 1e3); };
  rep2("count: atomicAdd u32", time_it([&] { k_count<<<(nb + 1023) / 1024, 256>>>(d_upd, nb, cnt, rank); }, 10, flush, flush_bytes));
  rep2("count: atomicAdd u32 + atomicExch list", time_it([&] { k_count_list<<<(nb + 1023) / 1024, 256>>>(d_upd, nb, cnt, last, rank, prev); }, 10, flush, flush_bytes));
  rep2("count: atomicAdd u64 (count | index sum)", time_it([&] { k_count64<<<(nb + 1023) / 1024, 256>>>(d_upd, nb, cnt64, rank); }, 10, flush, flush_bytes));
  // back-to-back launch gap: 5 empty-ish kernels
  rep2("5 x tiny kernel back to back (launch gaps)", time_it([&] { for (int k = 0; k < 5; ++k) k_count<<<1, 256>>>(d_upd, 256, cnt, rank); }, 10, flush, flush_bytes));
  printf("done\n");
  return 0;
}
