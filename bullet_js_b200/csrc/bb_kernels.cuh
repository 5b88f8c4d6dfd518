// bb_kernels.cuh - sm_100a kernels of the merge pipeline.
//
//   K0  make_keys      (path id, arrival index) -> 64-bit sort items, bounds check
//   K1  radix sort     stable LSD, 8-bit digits, only as many passes as the
//                      table's row-index width needs; moves 8-byte items, never payloads
//   K2  merge          one thread per path segment replays its updates in arrival
//                      order against the 128-byte table row (bb_merge.cuh)
//   K3  compaction     accepted updates -> dense change set in arrival order
//
// All of it is integer / f64 compare-and-move work: HBM-bound, no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bb_merge.cuh"

namespace bb {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 items per CTA
constexpr int RADIX = 256;

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

constexpr int MERGE_THREADS = 128;
constexpr int COMPACT_THREADS = 256;
constexpr int COMPACT_TILE = 4096;

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ---------------------------------------------------------------- K0
__global__ void __launch_bounds__(256) k_make_keys(const uint64_t* __restrict__ path_id, uint64_t n,
                                                   uint64_t capacity, uint64_t* __restrict__ items,
                                                   uint32_t* __restrict__ err) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t p = path_id[i];
  if (p >= capacity) atomicOr(err, 1u);  // the whole batch is rejected: K2/K3 see the flag and do nothing
  items[i] = (p << 32) | i;
}

// ---------------------------------------------------------------- generic exclusive scan (u32)
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = sum
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t inc = warp_inclusive_scan(v);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < THREADS / 32 ? wsum[lane] : 0;
    s = warp_inclusive_scan(s);
    if (lane < THREADS / 32) wsum[lane] = s;
  }
  __syncthreads();
  const uint32_t base = w ? wsum[w - 1] : 0;
  *total = wsum[THREADS / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, uint64_t n,
                                                              uint32_t* __restrict__ tile_sum) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint32_t total;
  block_exclusive_scan<SCAN_THREADS>(s, &total);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// single CTA: exclusive scan of m values in place; writes the grand total to *total_out
__global__ void __launch_bounds__(1024) k_scan_small(uint32_t* __restrict__ data, uint64_t m,
                                                     uint64_t* __restrict__ total_out) {
  uint32_t carry = 0;
  for (uint64_t base = 0; base < m; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const uint32_t v = i < m ? data[i] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan<1024>(v, &total);
    if (i < m) data[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(uint32_t* __restrict__ data, uint64_t n,
                                                             const uint32_t* __restrict__ tile_base) {
  // each thread owns SCAN_ITEMS consecutive elements so the scan is in index order
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = base + k < n ? data[base + k] : 0;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = tile_base[blockIdx.x] + block_exclusive_scan<SCAN_THREADS>(s, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
}

// ---------------------------------------------------------------- K1 radix sort
// counts layout: counts[digit * num_tiles + tile]  (digit-major so one exclusive scan
// over the whole array yields every (digit, tile) base)
__global__ void __launch_bounds__(SORT_THREADS) k_sort_count(const uint64_t* __restrict__ items, uint64_t n,
                                                             int shift, uint32_t num_tiles,
                                                             uint32_t* __restrict__ counts) {
  __shared__ uint32_t hist[RADIX];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const uint64_t i = base + (uint64_t)k * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&hist[(uint32_t)(items[i] >> (32 + shift)) & (RADIX - 1)], 1u);
  }
  __syncthreads();
  counts[(uint64_t)threadIdx.x * num_tiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint64_t* __restrict__ in,
                                                               uint64_t* __restrict__ out, uint64_t n, int shift,
                                                               uint32_t num_tiles,
                                                               const uint32_t* __restrict__ bases) {
  __shared__ uint32_t whist[SORT_WARPS][RADIX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < SORT_WARPS * RADIX; d += SORT_THREADS) (&whist[0][0])[d] = 0;
  __syncthreads();

  // warp w owns the contiguous run [tile + w*512, +512): arrival order inside the
  // tile is (warp, round, lane), which is what makes the pass stable
  const uint64_t wbase = (uint64_t)blockIdx.x * SORT_TILE + (uint64_t)w * (SORT_ITEMS * 32);
  uint64_t kv[SORT_ITEMS];
  uint32_t rank[SORT_ITEMS];
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const uint64_t i = wbase + k * 32 + lane;
    kv[k] = i < n ? in[i] : ~0ull;
  }
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const bool valid = wbase + k * 32 + lane < n;
    const uint32_t d = valid ? ((uint32_t)(kv[k] >> (32 + shift)) & (RADIX - 1)) : RADIX;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader && valid) {
      old = whist[w][d];
      whist[w][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[k] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: turn per-warp counts into global bases
    const int d = threadIdx.x;
    uint32_t b = bases[(uint64_t)d * num_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) {
      const uint32_t t = whist[ww][d];
      whist[ww][d] = b;
      b += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    if (wbase + k * 32 + lane < n) {
      const uint32_t d = (uint32_t)(kv[k] >> (32 + shift)) & (RADIX - 1);
      out[whist[w][d] + rank[k]] = kv[k];
    }
  }
}

// ---------------------------------------------------------------- K2 merge
struct MergeArgs {
  const uint64_t* sorted;  // [n] (path id << 32 | arrival index), stable-sorted by path id
  uint64_t n;
  uint4* table;            // rows, 8 x uint4 each
  const uint4* head;       // [n]
  const uint4* clk;        // [n][2]
  const uint4* val;        // [n][2]
  uint8_t* decision;       // [n]
  uint4* st_head;          // staging at arrival index: [n], [n][2], [n][2]
  uint4* st_clk;
  uint4* st_val;
  uint64_t seq_base;
  const uint32_t* err;     // non-zero: the batch was rejected by K0, leave the table alone
  Params p;
};

__device__ __forceinline__ void load_row(const uint4* __restrict__ row, RowState& r) {
  const uint4 q0 = row[0], q1 = row[1], q2 = row[2], q3 = row[3];
  const uint4 q4 = row[4], q5 = row[5], q6 = row[6], q7 = row[7];
  r.s.val[0] = (uint64_t)q0.x | ((uint64_t)q0.y << 32);
  r.s.val[1] = (uint64_t)q0.z | ((uint64_t)q0.w << 32);
  r.s.val[2] = (uint64_t)q1.x | ((uint64_t)q1.y << 32);
  r.s.val[3] = (uint64_t)q1.z | ((uint64_t)q1.w << 32);
  r.m.cnt[0] = q2.x; r.m.cnt[1] = q2.y; r.m.cnt[2] = q2.z; r.m.cnt[3] = q2.w;
  r.m.cnt[4] = q3.x; r.m.cnt[5] = q3.y; r.m.cnt[6] = q3.z; r.m.cnt[7] = q3.w;
  r.v.cnt[0] = q4.x; r.v.cnt[1] = q4.y; r.v.cnt[2] = q4.z; r.v.cnt[3] = q4.w;
  r.v.cnt[4] = q5.x; r.v.cnt[5] = q5.y; r.v.cnt[6] = q5.z; r.v.cnt[7] = q5.w;
  r.m.order = q6.x;
  r.v.order = q6.y;
  r.s.hdr = (uint64_t)q6.z | ((uint64_t)q6.w << 32);
  r.m.present = (q7.x & BB_ROW_M_PRESENT) != 0;
  r.v.present = (q7.x & BB_ROW_V_PRESENT) != 0;
  r.alias = (q7.x & BB_ROW_ALIAS) != 0;
  r.cseq = (uint64_t)q7.z | ((uint64_t)q7.w << 32);
}

__device__ __forceinline__ void store_row(uint4* __restrict__ row, const RowState& r) {
  row[0] = make_uint4((uint32_t)r.s.val[0], (uint32_t)(r.s.val[0] >> 32), (uint32_t)r.s.val[1],
                      (uint32_t)(r.s.val[1] >> 32));
  row[1] = make_uint4((uint32_t)r.s.val[2], (uint32_t)(r.s.val[2] >> 32), (uint32_t)r.s.val[3],
                      (uint32_t)(r.s.val[3] >> 32));
  row[2] = make_uint4(r.m.cnt[0], r.m.cnt[1], r.m.cnt[2], r.m.cnt[3]);
  row[3] = make_uint4(r.m.cnt[4], r.m.cnt[5], r.m.cnt[6], r.m.cnt[7]);
  row[4] = make_uint4(r.v.cnt[0], r.v.cnt[1], r.v.cnt[2], r.v.cnt[3]);
  row[5] = make_uint4(r.v.cnt[4], r.v.cnt[5], r.v.cnt[6], r.v.cnt[7]);
  row[6] = make_uint4(r.m.order, r.v.order, (uint32_t)r.s.hdr, (uint32_t)(r.s.hdr >> 32));
  const uint32_t flags = (r.m.present ? BB_ROW_M_PRESENT : 0u) | (r.v.present ? BB_ROW_V_PRESENT : 0u) |
                         (r.alias ? BB_ROW_ALIAS : 0u);
  row[7] = make_uint4(flags, 0u, (uint32_t)r.cseq, (uint32_t)(r.cseq >> 32));
}

__device__ __forceinline__ void load_update(const MergeArgs& a, uint32_t idx, uint64_t& uhdr, uint32_t& user,
                                            Clock& c, Value& x) {
  const uint4 h = a.head[idx];
  const uint4 c0 = a.clk[2 * (uint64_t)idx], c1 = a.clk[2 * (uint64_t)idx + 1];
  const uint4 v0 = a.val[2 * (uint64_t)idx], v1 = a.val[2 * (uint64_t)idx + 1];
  uhdr = (uint64_t)h.x | ((uint64_t)h.y << 32);
  user = h.w;
  c.cnt[0] = c0.x; c.cnt[1] = c0.y; c.cnt[2] = c0.z; c.cnt[3] = c0.w;
  c.cnt[4] = c1.x; c.cnt[5] = c1.y; c.cnt[6] = c1.z; c.cnt[7] = c1.w;
  c.order = h.z;
  c.present = 1;
  x.val[0] = (uint64_t)v0.x | ((uint64_t)v0.y << 32);
  x.val[1] = (uint64_t)v0.z | ((uint64_t)v0.w << 32);
  x.val[2] = (uint64_t)v1.x | ((uint64_t)v1.y << 32);
  x.val[3] = (uint64_t)v1.z | ((uint64_t)v1.w << 32);
  x.hdr = uhdr & ~(uint64_t)BB_HDR_FLAVOUR_NET;
}

__device__ __forceinline__ void store_change(const MergeArgs& a, uint32_t idx, uint32_t user, const Value& v,
                                             const Clock& c) {
  a.st_head[idx] = make_uint4((uint32_t)v.hdr, (uint32_t)(v.hdr >> 32), c.order, user);
  a.st_clk[2 * (uint64_t)idx] = make_uint4(c.cnt[0], c.cnt[1], c.cnt[2], c.cnt[3]);
  a.st_clk[2 * (uint64_t)idx + 1] = make_uint4(c.cnt[4], c.cnt[5], c.cnt[6], c.cnt[7]);
  a.st_val[2 * (uint64_t)idx] = make_uint4((uint32_t)v.val[0], (uint32_t)(v.val[0] >> 32), (uint32_t)v.val[1],
                                           (uint32_t)(v.val[1] >> 32));
  a.st_val[2 * (uint64_t)idx + 1] = make_uint4((uint32_t)v.val[2], (uint32_t)(v.val[2] >> 32),
                                               (uint32_t)v.val[3], (uint32_t)(v.val[3] >> 32));
}

__global__ void __launch_bounds__(MERGE_THREADS) k_merge(const MergeArgs a) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n || (*a.err & 1u)) return;
  const uint64_t item = a.sorted[i];
  const uint32_t key = (uint32_t)(item >> 32);
  if (i > 0 && (uint32_t)(a.sorted[i - 1] >> 32) == key) return;  // not a segment head

  uint4* row = a.table + (uint64_t)key * 8;
  RowState r;
  load_row(row, r);
  uint64_t j = i;
  uint64_t cur = item;
  while (true) {
    const uint32_t idx = (uint32_t)cur;
    uint64_t uhdr;
    uint32_t user;
    Clock uc, oc;
    Value x, ov;
    load_update(a, idx, uhdr, user, uc, x);
    const uint32_t code = resolve_step(a.p, r, uhdr, uc, x, a.seq_base + idx, ov, oc);
    a.decision[idx] = (uint8_t)code;
    if (BB_DEC_ACCEPTED(code)) store_change(a, idx, user, ov, oc);
    if (++j >= a.n) break;
    cur = a.sorted[j];
    if ((uint32_t)(cur >> 32) != key) break;
  }
  store_row(row, r);
}

// ---------------------------------------------------------------- K3 compaction
__global__ void __launch_bounds__(COMPACT_THREADS) k_accept_count(const uint8_t* __restrict__ decision,
                                                                  uint64_t n, uint32_t* __restrict__ tile_cnt,
                                                                  const uint32_t* __restrict__ err) {
  const uint64_t base = (uint64_t)blockIdx.x * COMPACT_TILE;
  uint32_t s = 0;
  const bool rejected = (*err & 1u) != 0;
  for (int k = threadIdx.x; k < COMPACT_TILE; k += COMPACT_THREADS) {
    const uint64_t i = base + k;
    if (!rejected && i < n) s += BB_DEC_ACCEPTED(decision[i]);
  }
  uint32_t total;
  block_exclusive_scan<COMPACT_THREADS>(s, &total);
  if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

struct CompactArgs {
  const uint8_t* decision;
  uint64_t n;
  const uint32_t* tile_base;  // exclusive scan of tile counts
  const uint4* st_head;
  const uint4* st_clk;
  const uint4* st_val;
  uint32_t* out_idx;
  uint4* out_head;
  uint4* out_clk;
  uint4* out_val;
  uint64_t cap;
  uint32_t* err;  // bit0 in: batch rejected; bit1 out: cap too small
};

__global__ void __launch_bounds__(COMPACT_THREADS) k_compact(const CompactArgs a) {
  // thread t owns the 16 consecutive updates [tile + 16t, +16): arrival order is kept
  constexpr int PER = COMPACT_TILE / COMPACT_THREADS;
  const uint64_t base = (uint64_t)blockIdx.x * COMPACT_TILE + (uint64_t)threadIdx.x * PER;
  uint32_t flags = 0;
  const bool rejected = (*a.err & 1u) != 0;
#pragma unroll
  for (int k = 0; k < PER; ++k)
    if (!rejected && base + k < a.n && BB_DEC_ACCEPTED(a.decision[base + k])) flags |= 1u << k;
  uint32_t total;
  uint64_t pos = (uint64_t)a.tile_base[blockIdx.x] + block_exclusive_scan<COMPACT_THREADS>(__popc(flags), &total);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    if (!((flags >> k) & 1u)) continue;
    const uint64_t i = base + k;
    if (pos >= a.cap) {
      atomicOr(a.err, 2u);
      return;
    }
    a.out_idx[pos] = (uint32_t)i;
    a.out_head[pos] = a.st_head[i];
    a.out_clk[2 * pos] = a.st_clk[2 * i];
    a.out_clk[2 * pos + 1] = a.st_clk[2 * i + 1];
    a.out_val[2 * pos] = a.st_val[2 * i];
    a.out_val[2 * pos + 1] = a.st_val[2 * i + 1];
    ++pos;
  }
}

// ---------------------------------------------------------------- table import / export
__global__ void __launch_bounds__(256) k_table_scatter(uint4* __restrict__ table, const uint64_t* __restrict__ ids,
                                                       const uint4* __restrict__ rows, uint64_t n,
                                                       uint64_t capacity, uint32_t* __restrict__ err) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i = t >> 3;
  if (i >= n) return;
  const uint64_t p = ids[i];
  if (p >= capacity) {
    atomicOr(err, 1u);
    return;
  }
  table[p * 8 + (t & 7)] = rows[t];
}

__global__ void __launch_bounds__(256) k_table_gather(uint4* __restrict__ table, const uint64_t* __restrict__ ids,
                                                      uint4* __restrict__ rows, uint64_t n, uint64_t capacity,
                                                      int materialise, uint64_t seq, uint32_t* __restrict__ err) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t p = ids[i];
  if (p >= capacity) {
    atomicOr(err, 1u);
    return;
  }
  uint4* row = table + p * 8;
  if (materialise) {  // Bullet._getData's side effect (src/bullet.js:122-124)
    RowState r;
    load_row(row, r);
    const uint32_t k = kind_of(r.s.hdr);
    if (k == BB_KIND_NONE || (k == BB_KIND_PRIM && prim_falsy(tag_of(r.s.hdr, 0), r.s.val[0]))) {
      if (k == BB_KIND_NONE) r.cseq = seq + i + 1;
      materialise_empty_object(r.s);
      store_row(row, r);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) rows[i * 8 + q] = row[q];
}

}  // namespace bb
