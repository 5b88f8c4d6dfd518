// bb_kernels.cuh - sm_100a kernels of the merge pipeline.
//
//   K0  k_keys_hist    (path id, arrival index) -> 64-bit sort items, bounds check,
//                      digit histograms of every radix pass in the same read
//   K1  k_sort_pass    stable LSD radix sort, 8-bit digits, ONE kernel per pass
//                      (chained-scan / decoupled look-back across tiles); only as many
//                      passes as the table's row-index width needs; moves 8-byte items,
//                      never payloads
//   K2  k_merge_stage  one CTA per tile of 128 sorted positions: payloads and rows staged in
//                      shared memory with cp.async, one thread per path segment replays its
//                      updates in arrival order, accepted entries compacted into the change
//                      set, rows written back with 16-byte stores (see the kernel's comment)
//   K4/K5               index build and the equals / range / count scans: bb_index.cuh
//
// All of it is integer / f64 compare-and-move work: HBM-bound, no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bb_merge.cuh"

namespace bb {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 2048 items per CTA: a 1 M batch is one wave
constexpr int RADIX = 256;
constexpr int MAX_PASSES = 4;

constexpr int UPD_Q = 5;         // 16-byte chunks per update payload (head 1 + clk 2 + val 2)
constexpr int ROW_Q = 8;         // 16-byte chunks per table row

// chained-scan tile states: 2 flag bits + 30 value bits in one word
constexpr uint32_t ST_AGG = 1u << 30, ST_PRE = 2u << 30, ST_MASK = 3u << 30, ST_VAL = ~ST_MASK;

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
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

// Exclusive prefix of `agg` over tiles 0..tile-1, published tile by tile (decoupled
// look-back). Called by all 32 lanes of one warp; tile ids come from an atomic ticket
// so every predecessor is already running.
__device__ __forceinline__ uint32_t tile_prefix(uint32_t* state, uint32_t tile, uint32_t agg) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_volatile(state, ST_PRE | agg);
    return 0;
  }
  if (lane == 0) st_volatile(state + tile, ST_AGG | agg);
  uint32_t excl = 0;
  int look = (int)tile - 1;
  while (true) {
    const int i = look - lane;
    uint32_t v = ST_PRE;  // virtual zero prefix in front of tile 0
    if (i >= 0) {
      do {
        v = ld_volatile(state + i);
      } while ((v & ST_MASK) == 0);
    }
    const uint32_t pre = __ballot_sync(0xffffffffu, (v & ST_MASK) == ST_PRE);
    if (pre) {
      const int first = __ffs(pre) - 1;
      excl += warp_sum(lane <= first ? (v & ST_VAL) : 0u);
      break;
    }
    excl += warp_sum(v & ST_VAL);
    look -= 32;
  }
  if (lane == 0) st_volatile(state + tile, ST_PRE | (excl + agg));
  return excl;
}

// ---------------------------------------------------------------- K0
// items[i] = path_id[i] << 32 | i ; ghist[pass][digit] += 1 for every pass
__global__ void __launch_bounds__(SORT_THREADS) k_keys_hist(const uint64_t* __restrict__ path_id, uint64_t n,
                                                            uint64_t capacity, int passes,
                                                            uint64_t* __restrict__ items,
                                                            uint32_t* __restrict__ ghist,
                                                            uint32_t* __restrict__ err) {
  __shared__ uint32_t hist[MAX_PASSES][RADIX];
  for (int p = 0; p < passes; ++p) hist[p][threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
  bool bad = false;
#pragma unroll 4
  for (int k = 0; k < SORT_ITEMS; ++k) {
    const uint64_t i = base + (uint64_t)k * SORT_THREADS + threadIdx.x;
    if (i < n) {
      const uint64_t pid = path_id[i];
      bad |= pid >= capacity;
      items[i] = (pid << 32) | i;
      for (int p = 0; p < passes; ++p) atomicAdd(&hist[p][(uint32_t)(pid >> (8 * p)) & (RADIX - 1)], 1u);
    }
  }
  if (bad) atomicOr(err, 1u);  // the whole batch is rejected: K2 sees the flag and does nothing
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const uint32_t c = hist[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * RADIX + threadIdx.x], c);
  }
}

// whole-batch bounds check ahead of a chunked host call: one bad id rejects every chunk
__global__ void __launch_bounds__(256) k_check_range(const uint64_t* __restrict__ path_id, uint64_t n,
                                                     uint64_t capacity, uint32_t* __restrict__ err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  for (; i < n; i += (uint64_t)gridDim.x * blockDim.x) bad |= path_id[i] >= capacity;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(err, 1u);
}

// one CTA per pass: exclusive scan of its 256-bin histogram, in place
__global__ void __launch_bounds__(RADIX) k_hist_scan(uint32_t* __restrict__ ghist) {
  uint32_t* h = ghist + blockIdx.x * RADIX;
  uint32_t total;
  const uint32_t v = h[threadIdx.x];
  h[threadIdx.x] = block_exclusive_scan<RADIX>(v, &total);
}

// ---------------------------------------------------------------- K1
__global__ void __launch_bounds__(SORT_THREADS, 4) k_sort_pass(const uint64_t* __restrict__ in,
                                                            uint64_t* __restrict__ out, uint64_t n, int shift,
                                                            const uint32_t* __restrict__ gbase,  // [256]
                                                            uint32_t* __restrict__ state,         // [tiles][256]
                                                            uint32_t* __restrict__ ticket) {
  __shared__ uint32_t whist[SORT_WARPS][RADIX];
  __shared__ uint32_t s_base[RADIX];
  __shared__ uint32_t s_tile;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int d = threadIdx.x; d < SORT_WARPS * RADIX; d += SORT_THREADS) (&whist[0][0])[d] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;

  // warp w owns the contiguous run [tile*TILE + w*ITEMS*32, +ITEMS*32): arrival order inside the
  // tile is (warp, round, lane), which is what makes the pass stable
  const uint64_t wbase = (uint64_t)tile * SORT_TILE + (uint64_t)w * (SORT_ITEMS * 32);
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
  {  // thread d: per-warp exclusive offsets, tile count, chained scan over tiles for digit d
    const int d = threadIdx.x;
    uint32_t cnt = 0;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) {
      const uint32_t t = whist[ww][d];
      whist[ww][d] = cnt;
      cnt += t;
    }
    uint32_t excl = 0;
    uint32_t* st = state + d;
    if (tile == 0) {
      st_volatile(st, ST_PRE | cnt);
    } else {
      st_volatile(st + (uint64_t)tile * RADIX, ST_AGG | cnt);
      for (int look = (int)tile - 1;; --look) {
        uint32_t v;
        do {
          v = ld_volatile(st + (uint64_t)look * RADIX);
        } while ((v & ST_MASK) == 0);
        excl += v & ST_VAL;
        if ((v & ST_MASK) == ST_PRE) break;
      }
      st_volatile(st + (uint64_t)tile * RADIX, ST_PRE | (excl + cnt));
    }
    s_base[d] = gbase[d] + excl;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    if (wbase + k * 32 + lane < n) {
      const uint32_t d = (uint32_t)(kv[k] >> (32 + shift)) & (RADIX - 1);
      out[s_base[d] + whist[w][d] + rank[k]] = kv[k];
    }
  }
}

// ---------------------------------------------------------------- K1': counting sort on dense path ids
// The table is direct-indexed, so when the batch is not tiny next to it the sort is a counting sort
// keyed by the row index itself: count (one atomic per update, its return value is the update's
// rank inside its path) -> exclusive scan over the rows -> place.  The rank order inside a path
// is whatever the atomics gave; k_cs_fix restores arrival order: segments of up to
// 8 updates in registers, longer ones (hot keys) by a CTA-wide LSD radix sort on the arrival index.
constexpr int CS_THREADS = 256;
constexpr int CS_SHORT = 8;            // longest segment fixed in registers
constexpr int CS_LONG_CTAS = 592;

constexpr int CS_ILP = 4;  // independent atomics / gathers in flight per thread

__global__ void __launch_bounds__(CS_THREADS) k_cs_count(const uint64_t* __restrict__ path_id, uint64_t n,
                                                         uint64_t capacity, uint32_t* __restrict__ cnt,
                                                         uint32_t* __restrict__ rank, uint32_t* __restrict__ err,
                                                         uint32_t* __restrict__ zero4, uint64_t* __restrict__ zero_n) {
  // first kernel of a batch: it also resets the few words later kernels accumulate into (saves two memsets)
  if (blockIdx.x == 0 && threadIdx.x < 8 && zero4) zero4[threadIdx.x] = 0;  // ctr[0..3] of the grouping kernels, n_hot
  if (blockIdx.x == 0 && threadIdx.x == 0 && zero_n) *zero_n = 0;
  const uint64_t i0 = (uint64_t)blockIdx.x * (CS_THREADS * CS_ILP) + threadIdx.x;
  uint64_t pid[CS_ILP];
  uint32_t r[CS_ILP];
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) pid[k] = i0 + k * CS_THREADS < n ? path_id[i0 + k * CS_THREADS] : 0;
  bool bad = false;
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) {
    const bool in = i0 + k * CS_THREADS < n;
    bad |= in && pid[k] >= capacity;
    r[k] = (in && pid[k] < capacity) ? atomicAdd(&cnt[pid[k]], 1u) : 0u;  // some order inside the path
  }
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k)
    if (i0 + k * CS_THREADS < n) rank[i0 + k * CS_THREADS] = r[k];
  if (bad) atomicOr(err, 1u);  // batch rejected; the scan still runs and clears the counts
}

// Exclusive scan of the counts over the rows, in two launches without any waiting between CTAs:
// per-tile sums, then every tile sums the tiles before it (they are few) and scans its own 4096
// rows.  The second pass also clears the counts, so the array is clean for the next batch.
constexpr int CS_TILE = CS_THREADS * 16;

__device__ __forceinline__ uint32_t cs_load16(const uint32_t* __restrict__ cnt, uint64_t i, uint4 (&v)[4]) {
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = *reinterpret_cast<const uint4*>(cnt + i + 4 * k);  // arrays are padded to whole tiles
    sum += v[k].x + v[k].y + v[k].z + v[k].w;
  }
  return sum;
}

__global__ void __launch_bounds__(CS_THREADS) k_cs_tile_sums(const uint32_t* __restrict__ cnt,
                                                             uint32_t* __restrict__ tile_sum) {
  uint4 v[4];
  uint32_t total;
  const uint32_t sum = cs_load16(cnt, (uint64_t)blockIdx.x * CS_TILE + 16 * threadIdx.x, v);
  block_exclusive_scan<CS_THREADS>(sum, &total);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(CS_THREADS) k_cs_offsets(uint32_t* __restrict__ cnt, const uint32_t* __restrict__ tile_sum,
                                                           uint32_t* __restrict__ off) {
  uint32_t before = 0;
  for (uint32_t t = threadIdx.x; t < blockIdx.x; t += CS_THREADS) before += tile_sum[t];
  uint32_t base;
  block_exclusive_scan<CS_THREADS>(before, &base);
  const uint64_t i = (uint64_t)blockIdx.x * CS_TILE + 16 * threadIdx.x;
  uint4 v[4];
  uint32_t total;
  uint32_t run = base + block_exclusive_scan<CS_THREADS>(cs_load16(cnt, i, v), &total);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint4 o;
    o.x = run; run += v[k].x;
    o.y = run; run += v[k].y;
    o.z = run; run += v[k].z;
    o.w = run; run += v[k].w;
    *reinterpret_cast<uint4*>(off + i + 4 * k) = o;
    *reinterpret_cast<uint4*>(cnt + i + 4 * k) = make_uint4(0, 0, 0, 0);
  }
}

__global__ void __launch_bounds__(CS_THREADS) k_cs_place(const uint64_t* __restrict__ path_id, uint64_t n,
                                                         uint64_t capacity, const uint32_t* __restrict__ rank,
                                                         const uint32_t* __restrict__ off, uint64_t* __restrict__ items,
                                                         const uint32_t* __restrict__ err) {
  if (*err & 1u) return;
  const uint64_t i0 = (uint64_t)blockIdx.x * (CS_THREADS * CS_ILP) + threadIdx.x;
  uint64_t pid[CS_ILP];
  uint32_t dst[CS_ILP];
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) pid[k] = i0 + k * CS_THREADS < n ? path_id[i0 + k * CS_THREADS] : 0;
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k)
    dst[k] = i0 + k * CS_THREADS < n ? off[pid[k]] + rank[i0 + k * CS_THREADS] : 0u;
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k)
    if (i0 + k * CS_THREADS < n) items[dst[k]] = (pid[k] << 32) | (i0 + k * CS_THREADS);
}

// one thread per sorted position; the thread on a segment's first position puts the segment in
// arrival order (<= CS_SHORT updates) or queues it for k_cs_fix_long
__global__ void __launch_bounds__(CS_THREADS) k_cs_fix(uint64_t* __restrict__ items, uint64_t n,
                                                       const uint32_t* __restrict__ off, uint2* __restrict__ long_list,
                                                       uint32_t* __restrict__ n_long, const uint32_t* __restrict__ err) {
  const uint64_t p = (uint64_t)blockIdx.x * CS_THREADS + threadIdx.x;
  if (p >= n || (*err & 1u)) return;
  const uint64_t it = items[p];
  const uint32_t key = (uint32_t)(it >> 32);
  const uint32_t start = off[key];
  if (start != (uint32_t)p) return;
  const uint32_t len = off[key + 1] - start;
  if (len < 2) return;
  if (len > CS_SHORT) {
    long_list[atomicAdd(n_long, 1u)] = make_uint2(start, len);
    return;
  }
  uint32_t v[CS_SHORT];
#pragma unroll
  for (int k = 0; k < CS_SHORT; ++k) v[k] = k < (int)len ? (uint32_t)items[p + k] : 0xFFFFFFFFu;
#pragma unroll
  for (int a = 1; a < CS_SHORT; ++a) {  // insertion sort as a fixed compare-exchange network (registers only)
#pragma unroll
    for (int b = a; b > 0; --b) {
      const uint32_t x = min(v[b - 1], v[b]), y = max(v[b - 1], v[b]);
      v[b - 1] = x;
      v[b] = y;
    }
  }
#pragma unroll
  for (int k = 0; k < CS_SHORT; ++k)
    if (k < (int)len) items[p + k] = ((uint64_t)key << 32) | v[k];
}

// hot keys: each queued segment is sorted on its arrival index by one CTA, LSD radix, 8-bit digits,
// ping-ponging between the item buffer and the scratch buffer at the same offsets
__global__ void __launch_bounds__(CS_THREADS) k_cs_fix_long(uint64_t* __restrict__ items, uint64_t* __restrict__ scratch,
                                                            const uint2* __restrict__ long_list,
                                                            const uint32_t* __restrict__ n_long, uint32_t* __restrict__ next,
                                                            const uint32_t* __restrict__ region_base) {
  __shared__ uint32_t hist[RADIX];
  __shared__ uint32_t whist[CS_THREADS / 32][RADIX];
  __shared__ uint32_t s_seg;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t lt = lanemask_lt();
  while (true) {
    __syncthreads();
    if (tid == 0) s_seg = atomicAdd(next, 1u);
    __syncthreads();
    if (s_seg >= *n_long) return;
    const uint2 seg = long_list[s_seg];
    const uint32_t len = seg.y;
    const uint32_t seg0 = seg.x + (region_base ? *region_base : 0u);
    uint64_t* cur = items + seg0;
    uint64_t* oth = scratch + seg0;
    for (int shift = 0; shift < 32; shift += 8) {
      hist[tid] = 0;
      __syncthreads();
      for (uint32_t i = tid; i < len; i += CS_THREADS) atomicAdd(&hist[((uint32_t)cur[i] >> shift) & 0xFFu], 1u);
      __syncthreads();
      const uint32_t mine = hist[tid];
      const bool trivial = __syncthreads_or(mine == len);  // every index has the same digit: nothing to do
      if (trivial) continue;
      uint32_t total;
      const uint32_t ex = block_exclusive_scan<CS_THREADS>(mine, &total);
      hist[tid] = ex;  // running base of digit `tid`
      __syncthreads();
      for (uint32_t c0 = 0; c0 < len; c0 += CS_THREADS * 8) {  // chunks in order, warps own contiguous runs
        for (int d = lane; d < RADIX; d += 32) whist[w][d] = 0;
        __syncwarp();
        const uint32_t wb = c0 + w * 256;
        uint64_t kv[8];
        uint32_t rank[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t i = wb + k * 32 + lane;
          const bool valid = i < len;
          kv[k] = valid ? cur[i] : 0;
          const uint32_t d = valid ? (((uint32_t)kv[k] >> shift) & 0xFFu) : RADIX;
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
        {  // digit `tid`: offsets of the warps' runs, then advance the running base
          uint32_t run = hist[tid];
#pragma unroll
          for (int ww = 0; ww < CS_THREADS / 32; ++ww) {
            const uint32_t t = whist[ww][tid];
            whist[ww][tid] = run;
            run += t;
          }
          hist[tid] = run;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t i = wb + k * 32 + lane;
          if (i < len) oth[whist[w][((uint32_t)kv[k] >> shift) & 0xFFu] + rank[k]] = kv[k];
        }
        __syncthreads();
      }
      uint64_t* t = cur;
      cur = oth;
      oth = t;
    }
    if (cur != items + seg0) {
      __syncthreads();
      for (uint32_t i = tid; i < len; i += CS_THREADS) oth[i] = cur[i];
    }
  }
}


// ---------------------------------------------------------------- K1'': grouping front end (the default)
// The merge kernel needs a path's updates adjacent and in arrival order; it does not need the paths
// themselves in order.  So after the count (k_cs_count: one atomic per update) the batch is only GROUPED:
//   singles  (the path has one update in the batch: ~2/3 of a uniform batch) keep their arrival order at
//            the front of the item list.  Their payload gathers in the merge kernel then walk the batch
//            arrays almost sequentially (full 64-byte DRAM accesses instead of 16/32-byte pieces of
//            them) and only their 128-byte table rows are accessed at random;
//   multis   the first-counted update of a path claims a run of cnt slots behind the singles with one
//            atomic; k_cg_place drops the path's updates into it, k_cg_fix / k_cs_fix_long put each
//            run in arrival order.
// No pass over the capacity-sized arrays: every kernel is O(batch), and the counters are cleared by
// the threads that used them.  Item-list order (hence change-set layout) is not deterministic;
// BB_CFG_ORDERED_CHANGES keeps the full counting sort.
constexpr uint32_t CG_MULTI = 0x80000000u;
constexpr int CG_CTR_SINGLE = 0, CG_CTR_MULTI = 1, CG_CTR_LONG = 2, CG_CTR_NEXT = 3;

__global__ void __launch_bounds__(CS_THREADS) k_cg_classify(const uint64_t* __restrict__ path_id, uint64_t n,
                                                            uint64_t capacity, uint32_t* __restrict__ cnt,
                                                            uint32_t* __restrict__ rank, uint2* __restrict__ off,
                                                            uint64_t* __restrict__ items, uint32_t* __restrict__ ctr,
                                                            uint2* __restrict__ long_list) {
  __shared__ uint32_t s_w[CS_ILP][CS_THREADS / 32];
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint64_t i0 = (uint64_t)blockIdx.x * (CS_THREADS * CS_ILP) + tid;
  uint64_t pid[CS_ILP];
  uint32_t r[CS_ILP], c[CS_ILP];
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) {
    const uint64_t i = i0 + k * CS_THREADS;
    pid[k] = i < n ? path_id[i] : ~0ull;
    r[k] = i < n ? rank[i] : 0u;
  }
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) c[k] = pid[k] < capacity ? cnt[pid[k]] : 0u;  // final: k_cs_count has completed
  uint32_t before[CS_ILP];
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) {
    const uint32_t m = __ballot_sync(0xffffffffu, c[k] == 1u);
    before[k] = __popc(m & lanemask_lt());
    if (lane == 0) s_w[k][w] = __popc(m);
  }
  __syncthreads();
  if (tid == 0) {  // exclusive scan of the 4 x 8 warp counts in (k, warp) order == arrival order inside the tile
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < CS_ILP; ++k)
#pragma unroll
      for (int ww = 0; ww < CS_THREADS / 32; ++ww) {
        const uint32_t t = s_w[k][ww];
        s_w[k][ww] = run;
        run += t;
      }
    s_base = run ? atomicAdd(&ctr[CG_CTR_SINGLE], run) : 0u;
  }
  __syncthreads();
  const uint32_t base = s_base;
  // runs of the multi-update paths: one claim per CTA (a same-address atomic per path would serialise)
  uint32_t claim = 0;
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k)
    if (c[k] > 1u && r[k] == 0u) claim += c[k];
  uint32_t claimed;
  uint32_t run = block_exclusive_scan<CS_THREADS>(claim, &claimed);
  if (tid == 0) s_base = claimed ? atomicAdd(&ctr[CG_CTR_MULTI], claimed) : 0u;  // s_base was read above, before the scan's barriers
  __syncthreads();
  run += s_base;
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) {
    const uint64_t i = i0 + k * CS_THREADS;
    const bool head = c[k] > 1u && r[k] == 0u;  // the path's first-counted update owns the run
    const bool is_long = head && c[k] > (uint32_t)CS_SHORT;
    const uint32_t lmask = __ballot_sync(0xffffffffu, is_long);
    uint32_t lbase = 0;
    if (lmask) {
      if (lane == __ffs(lmask) - 1) lbase = atomicAdd(&ctr[CG_CTR_LONG], __popc(lmask));
      lbase = __shfl_sync(0xffffffffu, lbase, __ffs(lmask) - 1);
    }
    if (i >= n) continue;
    uint32_t tag = 0;
    if (c[k] == 1u) {
      items[base + s_w[k][w] + before[k]] = (pid[k] << 32) | i;
      cnt[pid[k]] = 0;  // nobody else looks at this counter
    } else if (c[k] > 1u) {
      tag = r[k] | CG_MULTI;
      if (head) {
        off[pid[k]] = make_uint2(run, c[k]);
        if (is_long) long_list[lbase + __popc(lmask & lanemask_lt())] = make_uint2(run, c[k]);
        run += c[k];
      }
    }
    rank[i] = tag;
  }
}

__global__ void __launch_bounds__(CS_THREADS) k_cg_place(const uint64_t* __restrict__ path_id, uint64_t n,
                                                         const uint32_t* __restrict__ rank, const uint2* __restrict__ off,
                                                         uint32_t* __restrict__ cnt, uint64_t* __restrict__ items,
                                                         const uint32_t* __restrict__ ctr) {
  const uint32_t region = ctr[CG_CTR_SINGLE];  // the multi-update runs start behind the singles
  const uint64_t i0 = (uint64_t)blockIdx.x * (CS_THREADS * CS_ILP) + threadIdx.x;
  uint32_t tag[CS_ILP];
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) tag[k] = i0 + k * CS_THREADS < n ? rank[i0 + k * CS_THREADS] : 0u;
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) {
    if (!(tag[k] & CG_MULTI)) continue;
    const uint64_t i = i0 + k * CS_THREADS;
    const uint64_t pid = path_id[i];
    const uint32_t r = tag[k] & ~CG_MULTI;
    items[region + off[pid].x + r] = (pid << 32) | i;
    if (r == 0u) cnt[pid] = 0;  // k_cg_classify was the last reader
  }
}

// one thread per position of the multi-update region; the thread on a run's first slot sorts it (<= CS_SHORT)
__global__ void __launch_bounds__(CS_THREADS) k_cg_fix(uint64_t* __restrict__ items, const uint2* __restrict__ off,
                                                       const uint32_t* __restrict__ ctr) {
  const uint32_t p = blockIdx.x * CS_THREADS + threadIdx.x;
  if (p >= ctr[CG_CTR_MULTI]) return;
  uint64_t* run = items + ctr[CG_CTR_SINGLE] + p;
  const uint32_t key = (uint32_t)(run[0] >> 32);
  const uint2 o = off[key];
  if (o.x != p || o.y > (uint32_t)CS_SHORT) return;
  const int len = (int)o.y;
  uint32_t v[CS_SHORT];
#pragma unroll
  for (int k = 0; k < CS_SHORT; ++k) v[k] = k < len ? (uint32_t)run[k] : 0xFFFFFFFFu;
#pragma unroll
  for (int a = 1; a < CS_SHORT; ++a) {
#pragma unroll
    for (int b = a; b > 0; --b) {
      const uint32_t x = min(v[b - 1], v[b]), y = max(v[b - 1], v[b]);
      v[b - 1] = x;
      v[b] = y;
    }
  }
#pragma unroll
  for (int k = 0; k < CS_SHORT; ++k)
    if (k < len) run[k] = ((uint64_t)key << 32) | v[k];
}

}  // namespace bb

#include "bb_index.cuh"  // needs the warp / look-back helpers above

namespace bb {

// ---------------------------------------------------------------- K2
struct MergeArgs {
  const uint64_t* sorted;  // [n] (path id << 32 | arrival index), stable-sorted by path id
  uint64_t n;
  uint4* table;            // rows, 8 x uint4 each
  const uint4* head;       // [n]
  const uint4* clk;        // [n][2]
  const uint4* val;        // [n][2]
  uint32_t* verdict;       // [n] arrival order: code << 29 | slot
  uint64_t* n_changes;
  uint32_t* out_idx;       // change set, path-major order
  uint4* out_head;
  uint4* out_clk;
  uint4* out_val;
  uint64_t cap;
  uint32_t* st_idx;        // staging for the part of a segment that runs past its tile,
  uint4* st_ent;           // [n][5] indexed by sorted position
  uint32_t* tile_state;    // [num_tiles], zeroed per launch
  uint32_t* ticket;        // zeroed per launch
  uint32_t num_tiles;
  uint4* hot_list;         // (key, next position lo, hi, 0) of segments handed to k_merge_hot
  uint32_t* n_hot;         // zeroed per batch
  uint32_t hot_cap;
  uint64_t seq_base;
  uint32_t idx_base;       // added to the arrival indices this launch reports (chunked host calls)
  const uint64_t* chg_base;  // ORDERED: entries already in the change set when the launch began
  uint32_t* err;           // bit0 in: batch rejected by K0; bit1 out: cap too small
  Params p;
  IndexArgs ix;
};

__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

__device__ __forceinline__ void unpack_row(const uint4* q, RowState& r) {
  const uint4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5], q6 = q[6], q7 = q[7];
  r.s.val[0] = u64_of(q0.x, q0.y); r.s.val[1] = u64_of(q0.z, q0.w);
  r.s.val[2] = u64_of(q1.x, q1.y); r.s.val[3] = u64_of(q1.z, q1.w);
  r.m.cnt[0] = q2.x; r.m.cnt[1] = q2.y; r.m.cnt[2] = q2.z; r.m.cnt[3] = q2.w;
  r.m.cnt[4] = q3.x; r.m.cnt[5] = q3.y; r.m.cnt[6] = q3.z; r.m.cnt[7] = q3.w;
  r.v.cnt[0] = q4.x; r.v.cnt[1] = q4.y; r.v.cnt[2] = q4.z; r.v.cnt[3] = q4.w;
  r.v.cnt[4] = q5.x; r.v.cnt[5] = q5.y; r.v.cnt[6] = q5.z; r.v.cnt[7] = q5.w;
  r.m.order = q6.x;
  r.v.order = q6.y;
  r.s.meta = q6.z;
  r.s.ord = q6.w;
  r.m.present = (q7.x & BB_ROW_M_PRESENT) != 0;
  r.v.present = (q7.x & BB_ROW_V_PRESENT) != 0;
  r.alias = (q7.x & BB_ROW_ALIAS) != 0;
  r.xcnt = q7.y;
  r.cseq = u64_of(q7.z, q7.w);
}

__device__ __forceinline__ void pack_row(uint4* q, const RowState& r) {
  q[0] = make_uint4((uint32_t)r.s.val[0], (uint32_t)(r.s.val[0] >> 32), (uint32_t)r.s.val[1], (uint32_t)(r.s.val[1] >> 32));
  q[1] = make_uint4((uint32_t)r.s.val[2], (uint32_t)(r.s.val[2] >> 32), (uint32_t)r.s.val[3], (uint32_t)(r.s.val[3] >> 32));
  q[2] = make_uint4(r.m.cnt[0], r.m.cnt[1], r.m.cnt[2], r.m.cnt[3]);
  q[3] = make_uint4(r.m.cnt[4], r.m.cnt[5], r.m.cnt[6], r.m.cnt[7]);
  q[4] = make_uint4(r.v.cnt[0], r.v.cnt[1], r.v.cnt[2], r.v.cnt[3]);
  q[5] = make_uint4(r.v.cnt[4], r.v.cnt[5], r.v.cnt[6], r.v.cnt[7]);
  q[6] = make_uint4(r.m.order, r.v.order, r.s.meta, r.s.ord);
  const uint32_t flags = (r.m.present ? BB_ROW_M_PRESENT : 0u) | (r.v.present ? BB_ROW_V_PRESENT : 0u) |
                         (r.alias ? BB_ROW_ALIAS : 0u);
  q[7] = make_uint4(flags, r.xcnt, (uint32_t)r.cseq, (uint32_t)(r.cseq >> 32));
}

// update payload / change entry as 5 x uint4: [head][clk lo][clk hi][val lo][val hi]
__device__ __forceinline__ bool unpack_update(uint4 h, uint4 c0, uint4 c1, uint4 v0, uint4 v1, Clock& c, Value& x) {
  c.cnt[0] = c0.x; c.cnt[1] = c0.y; c.cnt[2] = c0.z; c.cnt[3] = c0.w;
  c.cnt[4] = c1.x; c.cnt[5] = c1.y; c.cnt[6] = c1.z; c.cnt[7] = c1.w;
  c.order = h.z;
  c.present = 1;
  x.val[0] = u64_of(v0.x, v0.y); x.val[1] = u64_of(v0.z, v0.w);
  x.val[2] = u64_of(v1.x, v1.y); x.val[3] = u64_of(v1.z, v1.w);
  x.meta = h.x & ~(uint32_t)BB_HDR_FLAVOUR_NET;
  x.ord = h.y;
  return (h.x & (uint32_t)BB_HDR_FLAVOUR_NET) != 0;
}

__device__ __forceinline__ void pack_change(uint4* q, uint32_t user, const Value& v, const Clock& c) {
  q[0] = make_uint4(v.meta, v.ord, c.order, user);
  q[1] = make_uint4(c.cnt[0], c.cnt[1], c.cnt[2], c.cnt[3]);
  q[2] = make_uint4(c.cnt[4], c.cnt[5], c.cnt[6], c.cnt[7]);
  q[3] = make_uint4((uint32_t)v.val[0], (uint32_t)(v.val[0] >> 32), (uint32_t)v.val[1], (uint32_t)(v.val[1] >> 32));
  q[4] = make_uint4((uint32_t)v.val[2], (uint32_t)(v.val[2] >> 32), (uint32_t)v.val[3], (uint32_t)(v.val[3] >> 32));
}

constexpr uint32_t NO_SLOT = BB_NO_SLOT;
constexpr int MT = 128;    // sorted positions per CTA tile == threads per CTA
constexpr int MT_WARPS = MT / 32;
constexpr int HOT_SERIAL = 8;  // updates of an overrunning segment its owner replays alone before k_merge_hot takes over
constexpr int HOT_CTAS = 64;
// Staged rows sit at their natural 128-byte stride with the 16-byte chunk index XOR-swizzled by the row
// number: conflict-free both for the 8-lanes-per-row copies and for the one-thread-per-row unpack
// (LDS.128 / STS.128 by 32 rows at once), and 2 KB smaller than a padded stride.
__device__ __forceinline__ int row_slot(int r, int c) { return r * ROW_Q + (c ^ (r & 7)); }

// One CTA == one tile of MT sorted positions, in three phases with all global traffic
// asynchronous and coalesced and all resolver work out of shared memory:
//   stage    every update payload of the tile (5 x 16 B, by arrival index) and the table row of
//            every path segment that starts in the tile (8 x 16 B) go global -> shared with
//            cp.async (LDGSTS): ~24 KB in flight per CTA, no registers held
//   resolve  one thread per segment replays its updates in arrival order on the staged row
//            (bb_merge.cuh), overwriting each accepted update's payload slot with its change entry
//   drain    accepted positions ranked by a block scan, tile totals chained with a decoupled
//            look-back, then verdicts (arrival order), change entries (path-major, compacted)
//            and the rows leave with warp-cooperative 16-byte stores
// A segment that runs past its tile is finished by its owner straight from global memory.
// 7 CTAs per SM at 72 registers.  (8 would fit the 27.7 KB of shared memory, but at 64 registers the resolver
// spills and the kernel measured 6 % slower: 104 vs 99 us.)
template <bool ORDERED, bool INDEXED, bool HOT = false>
__global__ void __launch_bounds__(MT, 7) k_merge_stage(const MergeArgs a) {
  __shared__ __align__(16) uint4 s_upd[MT * UPD_Q];
  __shared__ __align__(16) uint4 s_row[MT * ROW_Q];
  __shared__ uint32_t s_idx[MT], s_res[MT];
  __shared__ uint32_t s_hmask[MT_WARPS], s_wsum[MT_WARPS];
  __shared__ uint32_t s_tile, s_over, s_ex, s_nextk;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, wbase = w * 32;
  const uint32_t err_in = *a.err;  // bit 0 was decided before this launch; looked at after the loads are on their way
  uint32_t tile = blockIdx.x;
  if (ORDERED) {  // the look-back chain needs tiles to start in order; otherwise any order will do
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    tile = s_tile;
  }
  if (tid == 0) s_over = 0;
  const uint64_t base = (uint64_t)tile * MT;
  const uint64_t pos = base + tid;
  const bool valid = pos < a.n;
  const int nvalid = (int)min((uint64_t)MT, a.n - base);
  const uint64_t item = valid ? a.sorted[pos] : ~0ull;
  const uint32_t key = (uint32_t)(item >> 32), idx = (uint32_t)item;
  uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
  if (lane == 0) prev = (valid && pos > 0) ? (uint32_t)(a.sorted[pos - 1] >> 32) : ~key;
  const bool is_head = valid && key != prev;
  const uint32_t hmask = __ballot_sync(0xffffffffu, is_head);
  if (lane == 0) s_hmask[w] = hmask;
  s_idx[tid] = idx;
  // the key right behind the tile tells its last segment whether it goes on: fetched now, not in the resolver
  if (tid == MT - 1) s_nextk = base + MT < a.n ? (uint32_t)(a.sorted[base + MT] >> 32) : ~0u;

  // ---- stage: lane pairs fetch whole 32-byte clocks / values, 8 lanes fetch one 128-byte row
  if (valid) cp_async16(&s_upd[tid * UPD_Q], a.head + idx);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int e = j * 16 + (lane >> 1), half = lane & 1;
    const uint32_t eidx = __shfl_sync(0xffffffffu, idx, e);
    if (wbase + e < nvalid) {
      cp_async16(&s_upd[(wbase + e) * UPD_Q + 1 + half], a.clk + 2 * (uint64_t)eidx + half);
      cp_async16(&s_upd[(wbase + e) * UPD_Q + 3 + half], a.val + 2 * (uint64_t)eidx + half);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = j * 4 + (lane >> 3), chunk = lane & 7;
    const uint32_t ekey = __shfl_sync(0xffffffffu, key, e);
    if ((hmask >> e) & 1u) cp_async16(&s_row[row_slot(wbase + e, chunk)], a.table + (uint64_t)ekey * ROW_Q + chunk);
  }
  cp_async_wait_all();
  if (err_in & 1u) return;  // batch rejected by the front end: nothing may be written
  __syncthreads();

  // ---- resolve: thread == segment head
  if (is_head) {
    int end = nvalid;
    bool last = true;  // no later head in this tile
    const uint32_t above = lane < 31 ? (hmask & ~((2u << lane) - 1u)) : 0u;
    if (above) {
      end = wbase + __ffs(above) - 1;
      last = false;
    } else {
      for (int ww = w + 1; ww < MT_WARPS; ++ww) {
        const uint32_t m = s_hmask[ww];
        if (m) {
          end = ww * 32 + __ffs(m) - 1;
          last = false;
          break;
        }
      }
    }
    uint64_t prim[F], prim0[F];  // the node's entries in the dense index columns
    if (INDEXED) {
#pragma unroll
      for (int f = 0; f < F; ++f) prim0[f] = prim[f] = ((a.ix.mask >> f) & 1u) ? a.ix.pcol[f][key] : BB_KEY_NONE;
    }
    RowState r;
    {
      uint4 q[ROW_Q];
#pragma unroll
      for (int c = 0; c < ROW_Q; ++c) q[c] = s_row[row_slot(tid, c)];
      unpack_row(q, r);
    }
    for (int p = tid; p < end; ++p) {
      uint4* u = &s_upd[p * UPD_Q];
      const uint4 h = u[0];
      Clock c, oc;
      Value x, ov;
      const bool net = unpack_update(h, u[1], u[2], u[3], u[4], c, x);
      const uint32_t code = resolve_step(a.p, r, net, c, x, a.seq_base + s_idx[p], ov, oc);
      if (INDEXED) index_hook(a.ix, key, r.s, x, prim, r.xcnt, a.err);
      if (BB_DEC_ACCEPTED(code)) pack_change(u, h.w, ov, oc);
      s_res[p] = code;
    }
    if (last && nvalid == MT && s_nextk == key) {  // the tile's last segment runs on into the next tiles
      uint32_t over = 0;
      for (uint64_t gp = base + MT; gp < a.n; ++gp) {
        const uint64_t it = a.sorted[gp];
        if ((uint32_t)(it >> 32) != key) break;
        if (HOT && gp - (base + MT) >= (uint64_t)HOT_SERIAL) {  // a hot key (BB_CFG_HOT_KEYS): k_merge_hot replays the rest
          const uint32_t slot = atomicAdd(a.n_hot, 1u);
          if (slot < a.hot_cap) {
            a.hot_list[slot] = make_uint4(key, (uint32_t)gp, (uint32_t)(gp >> 32), 0u);
            break;
          }
        }
        const uint32_t ui = (uint32_t)it;
        const uint4 h = a.head[ui];
        Clock c, oc;
        Value x, ov;
        const bool net = unpack_update(h, a.clk[2 * (uint64_t)ui], a.clk[2 * (uint64_t)ui + 1],
                                       a.val[2 * (uint64_t)ui], a.val[2 * (uint64_t)ui + 1], c, x);
        const uint32_t code = resolve_step(a.p, r, net, c, x, a.seq_base + ui, ov, oc);
        if (INDEXED) index_hook(a.ix, key, r.s, x, prim, r.xcnt, a.err);
        if (BB_DEC_ACCEPTED(code)) {  // compacted into the staging slots this segment owns
          const uint64_t sp = base + MT + over;
          pack_change(a.st_ent + sp * UPD_Q, h.w, ov, oc);
          a.st_idx[sp] = ui | (code << 29);
          ++over;
        } else {
          a.verdict[ui] = (code << 29) | NO_SLOT;
        }
      }
      s_over = over;
    }
    {
      uint4 q[ROW_Q];
      pack_row(q, r);
#pragma unroll
      for (int c = 0; c < ROW_Q; ++c) s_row[row_slot(tid, c)] = q[c];
    }
    if (INDEXED) {
#pragma unroll
      for (int f = 0; f < F; ++f)
        if (prim[f] != prim0[f]) a.ix.pcol[f][key] = prim[f];
    }
  }
  __syncthreads();

  // ---- drain: rank the accepted positions, chain the tile totals
  int first = MT;  // positions before the tile's first head continue a segment an earlier tile owns
#pragma unroll
  for (int ww = MT_WARPS - 1; ww >= 0; --ww)
    if (s_hmask[ww]) first = ww * 32 + __ffs(s_hmask[ww]) - 1;
  const bool owned = valid && tid >= first;
  const uint32_t code = owned ? s_res[tid] : 0xFFu;
  const bool acc = owned && BB_DEC_ACCEPTED(code);
  const uint32_t amask = __ballot_sync(0xffffffffu, acc);
  if (lane == 0) s_wsum[w] = __popc(amask);
  __syncthreads();
  uint32_t rank = __popc(amask & lanemask_lt()), in_cnt = 0;
#pragma unroll
  for (int ww = 0; ww < MT_WARPS; ++ww) {
    const uint32_t c = s_wsum[ww];
    if (ww < w) rank += c;
    in_cnt += c;
  }
  const uint32_t over_cnt = s_over;
  // rows first: nothing below them depends on where the tile's entries land
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = j * 4 + (lane >> 3), chunk = lane & 7;
    const uint32_t ekey = __shfl_sync(0xffffffffu, key, e);
    if ((hmask >> e) & 1u) a.table[(uint64_t)ekey * ROW_Q + chunk] = s_row[row_slot(wbase + e, chunk)];
  }
  if (ORDERED) {  // change set in path-major order: chain the tile totals (decoupled look-back)
    if (w == 0) {
      const uint32_t ex = tile_prefix(a.tile_state, tile, in_cnt + over_cnt);
      if (lane == 0) {
        const uint64_t b0 = *a.chg_base;
        s_ex = (uint32_t)b0 + ex;
        if (tile == a.num_tiles - 1) *a.n_changes = b0 + ex + in_cnt + over_cnt;
      }
    }
  } else if (tid == 0) {  // tiles claim their slice of the change set as they finish
    s_ex = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(a.n_changes),
                               (unsigned long long)(in_cnt + over_cnt));
  }
  __syncthreads();
  const uint64_t obase = s_ex;

  // ---- verdicts (arrival order), change set (path-major order), rows
  bool overflow = false;
  const uint64_t dest = obase + rank;
  if (owned) a.verdict[idx] = (code << 29) | (acc ? (uint32_t)dest : NO_SLOT);
  if (acc) {
    if (dest < a.cap) {
      a.out_idx[dest] = a.idx_base + idx;
      a.out_head[dest] = s_upd[tid * UPD_Q];
    } else {
      overflow = true;
    }
  }
  uint32_t wex = 0;  // accepted positions in earlier warps of the tile
#pragma unroll
  for (int ww = 0; ww < MT_WARPS; ++ww)
    if (ww < w) wex += s_wsum[ww];
#pragma unroll
  for (int j = 0; j < 2; ++j) {  // lanes 2k, 2k+1 move the two halves of a 32-byte clock / value
    const int e = j * 16 + (lane >> 1), half = lane & 1;
    const uint64_t edest = obase + wex + __popc(amask & ((1u << e) - 1u));
    if (((amask >> e) & 1u) && edest < a.cap) {
      a.out_clk[2 * edest + half] = s_upd[(wbase + e) * UPD_Q + 1 + half];
      a.out_val[2 * edest + half] = s_upd[(wbase + e) * UPD_Q + 3 + half];
    }
  }
  // ---- entries of a segment tail that ran past the tile
  for (uint32_t k = tid; k < over_cnt; k += MT) {
    const uint64_t sp = base + MT + k, odest = obase + in_cnt + k;
    const uint32_t packed = a.st_idx[sp];
    const uint32_t gi = packed & NO_SLOT;
    a.verdict[gi] = (packed & ~NO_SLOT) | (uint32_t)odest;
    if (odest < a.cap) {
      const uint4* q = a.st_ent + sp * UPD_Q;
      a.out_idx[odest] = a.idx_base + gi;
      a.out_head[odest] = q[0];
      a.out_clk[2 * odest] = q[1];
      a.out_clk[2 * odest + 1] = q[2];
      a.out_val[2 * odest] = q[3];
      a.out_val[2 * odest + 1] = q[4];
    } else {
      overflow = true;
    }
  }
  if (overflow) atomicOr(a.err, 2u);
}

// ---------------------------------------------------------------- K2h: hot keys
// A path that takes thousands of a batch's updates (Zipf) is a serial chain for the thread that owns it:
// ~1 us per update.  But a network-flavour update is decided by (its clock, M, S) alone - V and the alias
// flag, the only things a REJECTED update changes, do not enter - and M, S change only when an update is
// accepted.  So one CTA per hot segment (handed over by k_merge_stage after HOT_SERIAL updates) evaluates the
// next 128 updates in parallel against the row in shared memory; everything in front of the first
// state-changing update (the first accepted one, or the first local put, whose clock IS V) is final, that
// update's own result is exact, and its thread publishes the row for the next round.  Rounds retire ~20-128
// updates instead of one.  Runs after k_merge_stage on the same stream; exits at once when nothing is hot.
__global__ void __launch_bounds__(MT) k_merge_hot(const MergeArgs a) {
  __shared__ __align__(16) uint4 s_row[ROW_Q];
  __shared__ uint32_t s_cnt[MT_WARPS], s_stop[MT_WARPS], s_loc[MT_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (*a.err & 1u) return;
  const uint32_t n_hot = min(*a.n_hot, a.hot_cap);
  for (uint32_t hseg = blockIdx.x; hseg < n_hot; hseg += gridDim.x) {
    const uint4 he = a.hot_list[hseg];
    const uint32_t hkey = he.x;
    uint64_t gp0 = (uint64_t)he.y | ((uint64_t)he.z << 32);
    __syncthreads();
    if (tid < ROW_Q) s_row[tid] = a.table[(uint64_t)hkey * ROW_Q + tid];
    __syncthreads();
    bool overflow = false;
    while (true) {
      const uint64_t gp = gp0 + tid;
      const uint64_t it = gp < a.n ? a.sorted[gp] : ~0ull;
      const bool mine = gp < a.n && (uint32_t)(it >> 32) == hkey;  // the segment's positions are a prefix of the window
      const uint32_t ui = (uint32_t)it;
      uint32_t code = 0;
      bool net = true;
      uint4 h = make_uint4(0, 0, 0, 0);
      RowState r;
      Clock oc;
      Value ov;
      if (mine) {
        h = a.head[ui];
        Clock c;
        Value x;
        net = unpack_update(h, a.clk[2 * (uint64_t)ui], a.clk[2 * (uint64_t)ui + 1], a.val[2 * (uint64_t)ui],
                            a.val[2 * (uint64_t)ui + 1], c, x);
        unpack_row(s_row, r);
        code = resolve_step(a.p, r, net, c, x, a.seq_base + ui, ov, oc);
      }
      const bool stop = mine && (BB_DEC_ACCEPTED(code) || !net);
      const uint32_t bm = __ballot_sync(0xffffffffu, mine), bs = __ballot_sync(0xffffffffu, stop),
                     bl = __ballot_sync(0xffffffffu, mine && !net);
      if (lane == 0) {
        s_cnt[w] = __popc(bm);
        s_stop[w] = bs;
        s_loc[w] = bl;
      }
      __syncthreads();  // also: every thread has unpacked the row
      int nseg = 0, f = MT;
      bool f_local = false;
#pragma unroll
      for (int ww = MT_WARPS - 1; ww >= 0; --ww) {
        nseg += (int)s_cnt[ww];
        if (s_stop[ww]) {
          const int b = __ffs(s_stop[ww]) - 1;
          f = ww * 32 + b;
          f_local = (s_loc[ww] >> b) & 1u;
        }
      }
      if (nseg == 0) break;
      // retired this round: up to and including the first stop - unless that is a local put further in, whose
      // clock depends on the V the updates in front of it leave: it waits for the next round's position 0
      const int retired = f >= nseg ? nseg : ((f_local && f > 0) ? f : f + 1);
      if (tid < retired) {
        if (tid == retired - 1) pack_row(s_row, r);  // its copy is the exact state after the retired updates
        if (BB_DEC_ACCEPTED(code)) {  // only the last retired one can be
          const uint64_t dest = atomicAdd(reinterpret_cast<unsigned long long*>(a.n_changes), 1ull);
          a.verdict[ui] = (code << 29) | (uint32_t)dest;
          if (dest < a.cap) {
            uint4 q[UPD_Q];
            pack_change(q, h.w, ov, oc);
            a.out_idx[dest] = a.idx_base + ui;
            a.out_head[dest] = q[0];
            a.out_clk[2 * dest] = q[1];
            a.out_clk[2 * dest + 1] = q[2];
            a.out_val[2 * dest] = q[3];
            a.out_val[2 * dest + 1] = q[4];
          } else {
            overflow = true;
          }
        } else {
          a.verdict[ui] = (code << 29) | NO_SLOT;
        }
      }
      gp0 += (uint64_t)retired;
      __syncthreads();  // the published row is visible; s_cnt / s_stop / s_loc may be rewritten
    }
    if (overflow) atomicOr(a.err, 2u);
    __syncthreads();
    if (tid < ROW_Q) a.table[(uint64_t)hkey * ROW_Q + tid] = s_row[tid];
  }
}

// ---------------------------------------------------------------- K2': the same merge, software-pipelined
// k_merge_stage's CTAs spend about half of their life waiting for memory with nothing else to do:
// ticket -> sorted items -> payloads and rows are three dependent DRAM round trips before the first
// resolver instruction, and a tile is gone after ~12 us (ncu: 21 % of the stall samples on the
// cp.async wait, 17 % on the loads in front of it).  Here a CTA is PERSISTENT (4 per SM, tiles dealt round-robin)
// and works on three tiles at once:
//     tile i      resolve + drain out of shared-memory stage i & 1
//     tile i + 1  payloads and rows in flight (cp.async group) into stage (i + 1) & 1
//     tile i + 2  its sorted items in flight into registers
// so in steady state no warp waits for DRAM: loads have a whole iteration to land.  Rows are staged at a
// 128-byte stride with the 16-byte chunk index XOR-swizzled by the row number (conflict-free both for the
// 8-lanes-per-row copies and for the one-thread-per-row unpack), which makes two stages fit four times
// into an SM.  Semantics, change-set layout (tiles claim their slice with one atomic) and the handling
// of a segment that runs past its tile are those of k_merge_stage<false, *>.
constexpr int MP_STAGES = 2;
constexpr int MP_CTAS_PER_SM = 4;

struct MergeStage {
  uint4 upd[MT * UPD_Q];
  uint4 row[MT * ROW_Q];
  uint32_t idx[MT];
  uint32_t key[MT];
  uint32_t hmask[MT_WARPS];
};

struct MergePipeSmem {
  MergeStage st[MP_STAGES];
  uint32_t res[MT];
  uint32_t wsum[MT_WARPS];
  uint32_t over, ex;
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <bool INDEXED>
__global__ void __launch_bounds__(MT, MP_CTAS_PER_SM) k_merge_pipe(const MergeArgs a) {
  extern __shared__ __align__(16) unsigned char mp_smem[];
  MergePipeSmem& sm = *reinterpret_cast<MergePipeSmem*>(mp_smem);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, wbase = w * 32;
  if (*a.err & 1u) return;
  const uint32_t G = gridDim.x, T = a.num_tiles;

  // sorted items of a tile -> registers (plus the key in front of the tile, for lane 0 of warp 0 .. 3)
  auto load_items = [&](uint32_t tile, uint64_t& item, uint32_t& prevk) {
    item = ~0ull;
    prevk = 0;
    if (tile >= T) return;
    const uint64_t pos = (uint64_t)tile * MT + tid;
    if (pos < a.n) {
      item = a.sorted[pos];
      if (lane == 0 && pos > 0) prevk = (uint32_t)(a.sorted[pos - 1] >> 32);
    }
  };
  // payloads + rows of a tile -> stage (one cp.async group)
  auto issue = [&](uint32_t tile, MergeStage& st, uint64_t item, uint32_t prevk) {
    if (tile < T) {
      const uint64_t base = (uint64_t)tile * MT;
      const uint64_t pos = base + tid;
      const bool valid = pos < a.n;
      const int nvalid = (int)min((uint64_t)MT, a.n - base);
      const uint32_t key = (uint32_t)(item >> 32), idx = (uint32_t)item;
      uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
      if (lane == 0) prev = (valid && pos > 0) ? prevk : ~key;
      const bool is_head = valid && key != prev;
      const uint32_t hmask = __ballot_sync(0xffffffffu, is_head);
      if (lane == 0) st.hmask[w] = hmask;
      st.idx[tid] = idx;
      st.key[tid] = key;
      if (valid) cp_async16(&st.upd[tid * UPD_Q], a.head + idx);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int e = j * 16 + (lane >> 1), half = lane & 1;
        const uint32_t eidx = __shfl_sync(0xffffffffu, idx, e);
        if (wbase + e < nvalid) {
          cp_async16(&st.upd[(wbase + e) * UPD_Q + 1 + half], a.clk + 2 * (uint64_t)eidx + half);
          cp_async16(&st.upd[(wbase + e) * UPD_Q + 3 + half], a.val + 2 * (uint64_t)eidx + half);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = j * 4 + (lane >> 3), chunk = lane & 7;
        const uint32_t ekey = __shfl_sync(0xffffffffu, key, e);
        if ((hmask >> e) & 1u) cp_async16(&st.row[row_slot(wbase + e, chunk)], a.table + (uint64_t)ekey * ROW_Q + chunk);
      }
    }
    cp_async_commit();  // an empty group keeps the wait_group arithmetic uniform
  };

  uint64_t item_n;
  uint32_t prev_n;
  {
    uint64_t item0;
    uint32_t prev0;
    load_items(blockIdx.x, item0, prev0);
    load_items(blockIdx.x + G, item_n, prev_n);
    issue(blockIdx.x, sm.st[0], item0, prev0);
  }

  int buf = 0;
  for (uint32_t tile = blockIdx.x; tile < T; tile += G, buf ^= 1) {
    MergeStage& st = sm.st[buf];
    // next tile's payloads and rows start now (its stage was drained before the barrier that ended the
    // previous iteration); the tile after that sends for its sorted items
    issue(tile + G, sm.st[buf ^ 1], item_n, prev_n);
    load_items(tile + 2 * G, item_n, prev_n);
    if (tid == 0) sm.over = 0;
    cp_async_wait_group<1>();
    __syncthreads();

    const uint64_t base = (uint64_t)tile * MT;
    const bool valid = base + tid < a.n;
    const int nvalid = (int)min((uint64_t)MT, a.n - base);
    const uint32_t key = st.key[tid], idx = st.idx[tid];
    const uint32_t hmask = st.hmask[w];
    const bool is_head = (hmask >> lane) & 1u;

    // ---- resolve: thread == segment head
    if (is_head) {
      int end = nvalid;
      bool last = true;
      const uint32_t above = lane < 31 ? (hmask & ~((2u << lane) - 1u)) : 0u;
      if (above) {
        end = wbase + __ffs(above) - 1;
        last = false;
      } else {
        for (int ww = w + 1; ww < MT_WARPS; ++ww) {
          const uint32_t m = st.hmask[ww];
          if (m) {
            end = ww * 32 + __ffs(m) - 1;
            last = false;
            break;
          }
        }
      }
      uint64_t prim[F], prim0[F];
      if (INDEXED) {
#pragma unroll
        for (int f = 0; f < F; ++f) prim0[f] = prim[f] = ((a.ix.mask >> f) & 1u) ? a.ix.pcol[f][key] : BB_KEY_NONE;
      }
      RowState r;
      {
        uint4 q[ROW_Q];
#pragma unroll
        for (int c = 0; c < ROW_Q; ++c) q[c] = st.row[row_slot(tid, c)];
        unpack_row(q, r);
      }
      for (int p = tid; p < end; ++p) {
        uint4* u = &st.upd[p * UPD_Q];
        const uint4 h = u[0];
        Clock c, oc;
        Value x, ov;
        const bool net = unpack_update(h, u[1], u[2], u[3], u[4], c, x);
        const uint32_t code = resolve_step(a.p, r, net, c, x, a.seq_base + st.idx[p], ov, oc);
        if (INDEXED) index_hook(a.ix, key, r.s, x, prim, r.xcnt, a.err);
        if (BB_DEC_ACCEPTED(code)) pack_change(u, h.w, ov, oc);
        sm.res[p] = code;
      }
      if (last && nvalid == MT) {  // last segment of a full tile: it may run on into the next tiles
        uint32_t over = 0;
        for (uint64_t gp = base + MT; gp < a.n; ++gp) {
          const uint64_t it = a.sorted[gp];
          if ((uint32_t)(it >> 32) != key) break;
          const uint32_t ui = (uint32_t)it;
          const uint4 h = a.head[ui];
          Clock c, oc;
          Value x, ov;
          const bool net = unpack_update(h, a.clk[2 * (uint64_t)ui], a.clk[2 * (uint64_t)ui + 1],
                                         a.val[2 * (uint64_t)ui], a.val[2 * (uint64_t)ui + 1], c, x);
          const uint32_t code = resolve_step(a.p, r, net, c, x, a.seq_base + ui, ov, oc);
          if (INDEXED) index_hook(a.ix, key, r.s, x, prim, r.xcnt, a.err);
          if (BB_DEC_ACCEPTED(code)) {
            const uint64_t sp = base + MT + over;
            pack_change(a.st_ent + sp * UPD_Q, h.w, ov, oc);
            a.st_idx[sp] = ui | (code << 29);
            ++over;
          } else {
            a.verdict[ui] = (code << 29) | NO_SLOT;
          }
        }
        sm.over = over;
      }
      {
        uint4 q[ROW_Q];
        pack_row(q, r);
#pragma unroll
        for (int c = 0; c < ROW_Q; ++c) st.row[row_slot(tid, c)] = q[c];
      }
      if (INDEXED) {
#pragma unroll
        for (int f = 0; f < F; ++f)
          if (prim[f] != prim0[f]) a.ix.pcol[f][key] = prim[f];
      }
    }
    __syncthreads();

    // ---- drain
    int first = MT;  // positions before the tile's first head continue a segment an earlier tile owns
#pragma unroll
    for (int ww = MT_WARPS - 1; ww >= 0; --ww)
      if (st.hmask[ww]) first = ww * 32 + __ffs(st.hmask[ww]) - 1;
    const bool owned = valid && tid >= first;
    const uint32_t code = owned ? sm.res[tid] : 0xFFu;
    const bool acc = owned && BB_DEC_ACCEPTED(code);
    const uint32_t amask = __ballot_sync(0xffffffffu, acc);
    if (lane == 0) sm.wsum[w] = __popc(amask);
    __syncthreads();
    uint32_t rank = __popc(amask & lanemask_lt()), in_cnt = 0, wex = 0;
#pragma unroll
    for (int ww = 0; ww < MT_WARPS; ++ww) {
      const uint32_t c = sm.wsum[ww];
      if (ww < w) wex += c;
      in_cnt += c;
    }
    rank += wex;
    const uint32_t over_cnt = sm.over;
    if (tid == 0)  // the tile claims its slice of the change set
      sm.ex = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(a.n_changes),
                                  (unsigned long long)(in_cnt + over_cnt));
    // rows go while the claim is in flight: nothing about them depends on where the entries land
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 4 + (lane >> 3), chunk = lane & 7;
      const uint32_t ekey = __shfl_sync(0xffffffffu, key, e);
      if ((hmask >> e) & 1u) a.table[(uint64_t)ekey * ROW_Q + chunk] = st.row[row_slot(wbase + e, chunk)];
    }
    __syncthreads();
    const uint64_t obase = sm.ex;

    bool overflow = false;
    const uint64_t dest = obase + rank;
    if (owned) a.verdict[idx] = (code << 29) | (acc ? (uint32_t)dest : NO_SLOT);
    if (acc) {
      if (dest < a.cap) {
        a.out_idx[dest] = a.idx_base + idx;
        a.out_head[dest] = st.upd[tid * UPD_Q];
      } else {
        overflow = true;
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int e = j * 16 + (lane >> 1), half = lane & 1;
      const uint64_t edest = obase + wex + __popc(amask & ((1u << e) - 1u));
      if (((amask >> e) & 1u) && edest < a.cap) {
        a.out_clk[2 * edest + half] = st.upd[(wbase + e) * UPD_Q + 1 + half];
        a.out_val[2 * edest + half] = st.upd[(wbase + e) * UPD_Q + 3 + half];
      }
    }
    for (uint32_t k = tid; k < over_cnt; k += MT) {
      const uint64_t sp = base + MT + k, odest = obase + in_cnt + k;
      const uint32_t packed = a.st_idx[sp];
      const uint32_t gi = packed & NO_SLOT;
      a.verdict[gi] = (packed & ~NO_SLOT) | (uint32_t)odest;
      if (odest < a.cap) {
        const uint4* q = a.st_ent + sp * UPD_Q;
        a.out_idx[odest] = a.idx_base + gi;
        a.out_head[odest] = q[0];
        a.out_clk[2 * odest] = q[1];
        a.out_clk[2 * odest + 1] = q[2];
        a.out_val[2 * odest] = q[3];
        a.out_val[2 * odest + 1] = q[4];
      } else {
        overflow = true;
      }
    }
    if (overflow) atomicOr(a.err, 2u);
    __syncthreads();  // the stage is free: the next iteration refills it
  }
  cp_async_wait_group<0>();
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
    unpack_row(row, r);
    const uint32_t k = kind_of(r.s.meta);
    if (k == BB_KIND_NONE || falsy_primitive(r.s)) {
      if (k == BB_KIND_NONE) r.cseq = seq + i + 1;
      materialise_empty_object(r.s);
      pack_row(row, r);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) rows[i * 8 + q] = row[q];
}

// ---------------------------------------------------------------- K6: shard routing pack
// A sharded table: path id p lives on rank p % world as local row p / world.  The pack
// is a STABLE partition of the batch by owner - updates for rank 0 first, arrival order kept
// inside every destination - so that the owner, which concatenates what it receives in source-rank
// order, replays each path in (source rank, arrival index) order.  Three small launches: per-tile
// destination counts, one CTA scanning them (tile-major inside destination-major), the scatter.
constexpr int RT_THREADS = 1024;   // updates per tile of the count / scatter kernels
constexpr int RS_THREADS = 256;    // threads of the single scan CTA
constexpr int RT_MAX_WORLD = 16;

__global__ void __launch_bounds__(RT_THREADS) k_route_count(const uint64_t* __restrict__ path_id, uint64_t n,
                                                            uint32_t world, uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t s_cnt[RT_MAX_WORLD];
  if (threadIdx.x < RT_MAX_WORLD) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * RT_THREADS + threadIdx.x;
  const uint32_t d = i < n ? (uint32_t)(path_id[i] % world) : world;
  for (uint32_t r = 0; r < world; ++r) {
    const uint32_t m = __ballot_sync(0xffffffffu, d == r);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt[r], __popc(m));
  }
  __syncthreads();
  if (threadIdx.x < world) tile_cnt[(uint64_t)blockIdx.x * world + threadIdx.x] = s_cnt[threadIdx.x];
}

// one CTA: tile_cnt[tile][r] -> exclusive offsets in the packed order; counts[r] = updates for rank r
__global__ void __launch_bounds__(RS_THREADS) k_route_scan(uint32_t* __restrict__ tile_cnt, uint32_t tiles,
                                                           uint32_t world, uint64_t* __restrict__ counts) {
  __shared__ uint32_t s_run;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (uint32_t r = 0; r < world; ++r) {
    const uint32_t start = s_run;
    for (uint32_t t0 = 0; t0 < tiles; t0 += RS_THREADS) {
      const uint32_t t = t0 + threadIdx.x;
      const uint32_t v = t < tiles ? tile_cnt[(uint64_t)t * world + r] : 0;
      uint32_t total;
      const uint32_t ex = block_exclusive_scan<RS_THREADS>(v, &total);
      if (t < tiles) tile_cnt[(uint64_t)t * world + r] = s_run + ex;
      __syncthreads();
      if (threadIdx.x == 0) s_run += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) counts[r] = s_run - start;
    __syncthreads();
  }
}

struct RouteArgs {
  const uint64_t* path_id; const uint4* head; const uint4* clk; const uint4* val;  // [n] in
  uint64_t* o_path; uint4* o_head; uint4* o_clk; uint4* o_val;                      // [n] packed out
  uint64_t n;
  uint32_t world;
  const uint32_t* tile_off;  // [tiles][world] from k_route_scan
};

// Fused pack + all-to-all: the same stable partition, but every row is stored straight into the
// receive slot of its owner - peer memory mapped over NVLink (cudaIpc) - at the place the owner's
// concatenation in source-rank order gives it.  No send buffer, no separate exchange launch.
struct RouteP2PArgs {
  const uint64_t* path_id; const uint4* head; const uint4* clk; const uint4* val;  // [n] in
  uint64_t* d_path[RT_MAX_WORLD]; uint4* d_head[RT_MAX_WORLD]; uint4* d_clk[RT_MAX_WORLD]; uint4* d_val[RT_MAX_WORLD];
  const uint64_t* matrix;     // [world][world] on the device: row p = what rank p sends to each rank
  uint64_t slot_cap;          // rows a receive slot holds: nothing is stored if some owner would overflow
  uint32_t me;
  uint64_t n;
  uint32_t world;
  uint32_t bulk;  // 1: runs leave with cp.async.bulk (default); 0: with per-thread 16-byte stores
  const uint32_t* tile_off;
};

// Each CTA partitions its 1024 rows by owner in shared memory (88 KB), then streams every owner's
// run out with fully coalesced 16-byte stores: one contiguous run per (tile, owner, array).
// ---- signalling between the ranks' routers through peer-mapped memory (one RouteCtl per rank)
// The two tiny collectives of a route (everybody's counts before the scatter, "all my stores have
// landed" after it) cost 25-40 us each as NCCL all-gathers at 8 GPUs, on the critical path of the step.
// Here a rank stores its words straight into every peer's control block, then a flag (after a
// system-scope fence), and spins on its own flags: a few microseconds.  Flags carry the route's
// epoch and only grow, so a peer that is already one route ahead never confuses a waiter.
struct RouteCtl {
  uint64_t matrix[2][RT_MAX_WORLD * RT_MAX_WORLD];  // [slot][source rank][destination rank]
  uint64_t cflag[2][RT_MAX_WORLD];                  // [slot][source]: that source's counts row is in
  uint64_t bflag[2][RT_MAX_WORLD];                  // [slot][source]: that source's rows have landed
  uint64_t err;
};

struct RouteCtlPeers {
  RouteCtl* ctl[RT_MAX_WORLD];  // [rank]; our own entry is local memory
};

constexpr long long RT_SPIN_LIMIT = 120000000000ll;  // ~60 s of SM clocks: a peer died; fail loudly instead of hanging

__device__ __forceinline__ uint64_t ld_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void spin_until(const uint64_t* flag, uint64_t epoch, uint64_t* err) {
  const long long t0 = clock64();
  while (ld_sys(flag) < epoch) {
    if (clock64() - t0 > RT_SPIN_LIMIT) {
      st_sys(err, 1);
      __trap();
    }
    __nanosleep(64);
  }
}

// counts[W] of this rank -> row `me` of everybody's matrix, then wait for everybody's row
__global__ void __launch_bounds__(RT_MAX_WORLD * RT_MAX_WORLD) k_route_publish(const uint64_t* __restrict__ counts,
                                                                                RouteCtlPeers peers, uint32_t me, uint32_t world,
                                                                                uint32_t slot, uint64_t epoch) {
  const uint32_t t = threadIdx.x;
  if (t < world * world) {
    const uint32_t q = t / world, j = t % world;
    st_sys(&peers.ctl[q]->matrix[slot][(uint64_t)me * world + j], counts[j]);
  }
  __threadfence_system();
  __syncthreads();
  if (t < world) {
    __threadfence_system();
    st_sys(&peers.ctl[t]->cflag[slot][me], epoch);
    spin_until(&peers.ctl[me]->cflag[slot][t], epoch, &peers.ctl[me]->err);
  }
  __threadfence_system();
}

// "every store of my scatter kernel has landed" to everybody, then wait for everybody's
__global__ void __launch_bounds__(32) k_route_barrier(RouteCtlPeers peers, uint32_t me, uint32_t world, uint32_t slot,
                                                      uint64_t epoch) {
  const uint32_t t = threadIdx.x;
  if (t < world) {
    __threadfence_system();  // cumulative: the previous kernel's stores (visible to this thread) go first
    st_sys(&peers.ctl[t]->bflag[slot][me], epoch);
    spin_until(&peers.ctl[me]->bflag[slot][t], epoch, &peers.ctl[me]->err);
  }
  __threadfence_system();
}

constexpr int RT_SMEM = RT_THREADS * 88;

__global__ void __launch_bounds__(RT_THREADS) k_route_scatter_p2p(const RouteP2PArgs a) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint4* s_head = reinterpret_cast<uint4*>(s_raw);                   // [1024]
  uint4* s_clk = s_head + RT_THREADS;                                 // [2048]
  uint4* s_val = s_clk + 2 * RT_THREADS;                              // [2048]
  uint64_t* s_path = reinterpret_cast<uint64_t*>(s_val + 2 * RT_THREADS);  // [1024]
  __shared__ uint32_t s_w[RT_THREADS / 32][RT_MAX_WORLD];
  __shared__ uint32_t s_start[RT_MAX_WORLD + 1];
  __shared__ int64_t s_dst[RT_MAX_WORLD];  // destination row of the owner's run minus its start in the tile
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // where this rank's block starts in every owner's slot comes from the all-gathered counts ON THE DEVICE:
  // the host never waits for them (row in the owner's slot = packed position + adj[owner])
  __shared__ int64_t s_adj[RT_MAX_WORLD];
  __shared__ uint32_t s_bad;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  if (tid < (int)a.world) {
    uint64_t before = 0, so = 0, col = 0;
    for (uint32_t p = 0; p < a.world; ++p) {
      const uint64_t c = a.matrix[(uint64_t)p * a.world + tid];
      col += c;
      if (p < a.me) before += c;
    }
    for (int q = 0; q < tid; ++q) so += a.matrix[(uint64_t)a.me * a.world + q];
    s_adj[tid] = (int64_t)before - (int64_t)so;
    if (col > a.slot_cap) s_bad = 1;
  }
  __syncthreads();
  if (s_bad) return;  // every rank sees the same matrix and skips; bb_router_acquire reports it
  // PERSISTENT on a small grid (bb_router: 64 CTAs): the kernel is NVLink-bound and runs next to the
  // merge of the previous batch; one CTA per tile would take every SM's thread slots away from it
  const uint32_t tiles = (uint32_t)((a.n + RT_THREADS - 1) / RT_THREADS);
  for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
  const uint64_t i = (uint64_t)tile * RT_THREADS + tid;
  const uint64_t p = i < a.n ? a.path_id[i] : 0;
  const uint32_t d = i < a.n ? (uint32_t)(p % a.world) : a.world;
  uint4 h, c0, c1, v0, v1;
  if (i < a.n) {
    h = a.head[i];
    c0 = a.clk[2 * i];
    c1 = a.clk[2 * i + 1];
    v0 = a.val[2 * i];
    v1 = a.val[2 * i + 1];
  }
  uint32_t below = 0;
  for (uint32_t r = 0; r < a.world; ++r) {
    const uint32_t m = __ballot_sync(0xffffffffu, d == r);
    if (d == r) below = __popc(m & lanemask_lt());
    if (lane == 0) s_w[w][r] = __popc(m);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (uint32_t r = 0; r < a.world; ++r) {
      s_start[r] = run;
      s_dst[r] = (int64_t)a.tile_off[(uint64_t)tile * a.world + r] + s_adj[r] - (int64_t)run;
      for (int ww = 0; ww < RT_THREADS / 32; ++ww) run += s_w[ww][r];
    }
    s_start[a.world] = run;
  }
  __syncthreads();
  if (i < a.n) {
    uint32_t lp = s_start[d] + below;
    for (int ww = 0; ww < w; ++ww) lp += s_w[ww][d];
    s_path[lp] = p / a.world;
    s_head[lp] = h;
    s_clk[2 * lp] = c0;
    s_clk[2 * lp + 1] = c1;
    s_val[2 * lp] = v0;
    s_val[2 * lp + 1] = v1;
  }
  __syncthreads();
  const uint32_t rows = s_start[a.world];
  if (a.bulk) {
    // Every (owner, array) run is contiguous in shared memory and in the owner's slot: ONE bulk copy each
    // (cp.async.bulk shared -> global, to peer memory over NVLink).  The copy engine of the SM moves the
    // bytes; no thread, register or LSU slot waits for the remote stores, so a handful of CTAs keeps
    // NVLink busy and the merge kernel running next to them keeps its SMs to itself.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the partition above was written with st.shared
    if (tid < (int)(3 * a.world)) {
      const uint32_t r = tid / 3, arr = tid % 3;
      const uint32_t first = s_start[r], cnt = s_start[r + 1] - first;
      if (cnt) {
        const uint64_t drow = (uint64_t)(s_dst[r] + (int64_t)first);
        const void* src;
        void* dst;
        uint32_t bytes;
        if (arr == 0) {
          src = s_head + first, dst = a.d_head[r] + drow, bytes = cnt * 16u;
        } else if (arr == 1) {
          src = s_clk + 2 * first, dst = a.d_clk[r] + 2 * drow, bytes = cnt * 32u;
        } else {
          src = s_val + 2 * first, dst = a.d_val[r] + 2 * drow, bytes = cnt * 32u;
        }
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                     "r"((uint32_t)__cvta_generic_to_shared(src)), "r"(bytes)
                     : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (tid < (int)rows) {  // the 8-byte local row ids: runs are not 16-byte aligned, plain stores
      uint32_t r = 0;
      while ((uint32_t)tid >= s_start[r + 1]) ++r;
      a.d_path[r][(uint64_t)(s_dst[r] + (int64_t)tid)] = s_path[tid];
    }
    if (tid < (int)(3 * a.world)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // sources read: reusable
  } else {
  {  // rows of the tile in partitioned order: thread j moves row j
    const uint32_t j = tid;
    if (j < rows) {
      uint32_t r = 0;
      while (j >= s_start[r + 1]) ++r;
      const uint64_t dst = (uint64_t)(s_dst[r] + (int64_t)j);
      a.d_path[r][dst] = s_path[j];
      a.d_head[r][dst] = s_head[j];
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // the 32-byte columns as 2048 16-byte pieces
    const uint32_t e = tid + k * RT_THREADS, j = e >> 1;
    if (j < rows) {
      uint32_t r = 0;
      while (j >= s_start[r + 1]) ++r;
      const uint64_t dst = 2 * (uint64_t)(s_dst[r] + (int64_t)j) + (e & 1u);
      a.d_clk[r][dst] = s_clk[e];
      a.d_val[r][dst] = s_val[e];
    }
  }
  }
  __syncthreads();  // shared memory is reused by the next tile
  }
  if (a.bulk && tid < (int)(3 * a.world)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // writes done
}

__global__ void __launch_bounds__(RT_THREADS) k_route_scatter(const RouteArgs a) {
  __shared__ uint32_t s_w[RT_THREADS / 32][RT_MAX_WORLD];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t i = (uint64_t)blockIdx.x * RT_THREADS + threadIdx.x;
  const uint64_t p = i < a.n ? a.path_id[i] : 0;
  const uint32_t d = i < a.n ? (uint32_t)(p % a.world) : a.world;
  uint32_t below = 0;
  for (uint32_t r = 0; r < a.world; ++r) {
    const uint32_t m = __ballot_sync(0xffffffffu, d == r);
    if (d == r) below = __popc(m & lanemask_lt());
    if (lane == 0) s_w[w][r] = __popc(m);
  }
  __syncthreads();
  if (i >= a.n) return;
  uint32_t dst = a.tile_off[(uint64_t)blockIdx.x * a.world + d] + below;
  for (int ww = 0; ww < w; ++ww) dst += s_w[ww][d];
  a.o_path[dst] = p / a.world;
  a.o_head[dst] = a.head[i];
  a.o_clk[2 * (uint64_t)dst] = a.clk[2 * i];
  a.o_clk[2 * (uint64_t)dst + 1] = a.clk[2 * i + 1];
  a.o_val[2 * (uint64_t)dst] = a.val[2 * i];
  a.o_val[2 * (uint64_t)dst + 1] = a.val[2 * i + 1];
}

}  // namespace bb
