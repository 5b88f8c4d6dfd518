// bb_kernels.cuh - sm_100a kernels of the merge pipeline.
//
//   K0  k_keys_hist    (path id, arrival index) -> 64-bit sort items, bounds check,
//                      digit histograms of every radix pass in the same read
//   K1  k_sort_pass    stable LSD radix sort, 8-bit digits, ONE kernel per pass
//                      (chained-scan / decoupled look-back across tiles); only as many
//                      passes as the table's row-index width needs; moves 8-byte items,
//                      never payloads
//   K2  k_merge_stage  one CTA per tile of 128 item-list positions: payloads and rows staged in shared
//                      memory with cp.async, one thread per path segment replays its updates in arrival
//                      order, accepted entries compacted into the change set, rows written back with
//                      16-byte stores.  The default front end that builds the item list (grouping, not
//                      sorting) is bb_group.cuh; hot keys go to k_merge_hot
//   K4/K5               index build and the equals / range / count scans: bb_index.cuh
//
// All of it is integer / f64 compare-and-move work: HBM-bound, no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bb_merge.cuh"
#include "bb_ptx.cuh"

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
      while (((v = ld_volatile(state + i)) & ST_MASK) == 0) BB_SPIN_YIELD();
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

__device__ __forceinline__ void flag_reject(uint32_t* rej, uint32_t* err, uint32_t ordinal) {
  atomicOr(rej, 1u);
  atomicOr(err, ERR_RANGE);
  atomicMin(err + 1, ordinal);
}

// ---------------------------------------------------------------- K0
// items[i] = path_id[i] << 32 | i ; ghist[pass][digit] += 1 for every pass
__global__ void __launch_bounds__(SORT_THREADS) k_keys_hist(const uint64_t* __restrict__ path_id, uint64_t n,
                                                            uint64_t capacity, int passes,
                                                            uint64_t* __restrict__ items,
                                                            uint32_t* __restrict__ ghist,
                                                            uint32_t* __restrict__ rej, uint32_t* __restrict__ err,
                                                            uint32_t ordinal) {
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
  if (bad) flag_reject(rej, err, ordinal);  // the whole batch is rejected: K2s sees the flag and does nothing
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const uint32_t c = hist[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * RADIX + threadIdx.x], c);
  }
}

// whole-batch bounds check ahead of a chunked host call: one bad id rejects every chunk
__global__ void __launch_bounds__(256) k_check_range(const uint64_t* __restrict__ path_id, uint64_t n,
                                                     uint64_t capacity, uint32_t* __restrict__ rej,
                                                     uint32_t* __restrict__ err, uint32_t ordinal) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  for (; i < n; i += (uint64_t)gridDim.x * blockDim.x) bad |= path_id[i] >= capacity;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) flag_reject(rej, err, ordinal);
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
        while (((v = ld_volatile(st + (uint64_t)look * RADIX)) & ST_MASK) == 0) BB_SPIN_YIELD();
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
                                                         uint32_t* __restrict__ rank, uint32_t* __restrict__ rej,
                                                         uint32_t* __restrict__ err, uint32_t ordinal) {
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
  if (bad) flag_reject(rej, err, ordinal);  // batch rejected; the scan still runs and clears the counts
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
                                                         const uint32_t* __restrict__ rej) {
  if (*rej & 1u) return;
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
                                                       uint32_t* __restrict__ n_long, const uint32_t* __restrict__ rej) {
  const uint64_t p = (uint64_t)blockIdx.x * CS_THREADS + threadIdx.x;
  if (p >= n || (*rej & 1u)) return;
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
  pdl_launch_dependents();
  pdl_wait();
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
  uint4* hot_list;         // (key, next position lo, hi, 0) of the segments handed to k_merge_hot
  uint32_t* n_hot;         // zero when the batch begins
  uint32_t hot_cap;
  uint64_t seq_base;
  uint32_t idx_base;       // added to the arrival indices this launch reports (chunked host calls)
  const uint64_t* chg_base;  // ORDERED: entries already in the change set when the launch began
  uint32_t* epoch_col;     // BB_CFG_TRACK_MODIFIED: [capacity] ordinal of the merge call that last wrote the row, or null
  uint32_t epoch;          // this call's ordinal
  uint32_t* err;           // sticky until bb_sync: bit1 change buffer too small, bit2 overflow set full
  const uint32_t* rej;     // bit0: THIS batch was rejected by the front end (path id out of range): touch nothing
  Params p;
  IndexArgs ix;
};

__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// In HBM a row's eight 16-byte chunks are stored in the order 0, 1, 6, 7, 2, 3, 4, 5 of the public bb_row: the values,
// the header word (kind, tags, key order), the flags and the creation sequence - everything an index build or a
// scan of the stored values needs - sit in the FIRST 64 bytes, the two clocks' counts in the second.  A random
// row access is two 64-byte DRAM accesses either way; a pass that only needs the first half moves half the bytes.
// bb_table_load / bb_table_read convert; everything else goes through unpack_row / pack_row.
__host__ __device__ constexpr int row_chunk(int c) { return c < 2 ? c : (c >= 6 ? c - 4 : c + 2); }

__device__ __forceinline__ void unpack_row(const uint4* q, RowState& r) {
  const uint4 q0 = q[row_chunk(0)], q1 = q[row_chunk(1)], q2 = q[row_chunk(2)], q3 = q[row_chunk(3)], q4 = q[row_chunk(4)],
              q5 = q[row_chunk(5)], q6 = q[row_chunk(6)], q7 = q[row_chunk(7)];
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
  q[row_chunk(0)] = make_uint4((uint32_t)r.s.val[0], (uint32_t)(r.s.val[0] >> 32), (uint32_t)r.s.val[1], (uint32_t)(r.s.val[1] >> 32));
  q[row_chunk(1)] = make_uint4((uint32_t)r.s.val[2], (uint32_t)(r.s.val[2] >> 32), (uint32_t)r.s.val[3], (uint32_t)(r.s.val[3] >> 32));
  q[row_chunk(2)] = make_uint4(r.m.cnt[0], r.m.cnt[1], r.m.cnt[2], r.m.cnt[3]);
  q[row_chunk(3)] = make_uint4(r.m.cnt[4], r.m.cnt[5], r.m.cnt[6], r.m.cnt[7]);
  q[row_chunk(4)] = make_uint4(r.v.cnt[0], r.v.cnt[1], r.v.cnt[2], r.v.cnt[3]);
  q[row_chunk(5)] = make_uint4(r.v.cnt[4], r.v.cnt[5], r.v.cnt[6], r.v.cnt[7]);
  q[row_chunk(6)] = make_uint4(r.m.order, r.v.order, r.s.meta, r.s.ord);
  const uint32_t flags = (r.m.present ? BB_ROW_M_PRESENT : 0u) | (r.v.present ? BB_ROW_V_PRESENT : 0u) |
                         (r.alias ? BB_ROW_ALIAS : 0u);
  q[row_chunk(7)] = make_uint4(flags, r.xcnt, (uint32_t)r.cseq, (uint32_t)(r.cseq >> 32));
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

// BB_CFG_COMPACT_CHANGES: is the change entry of an accepted update the update itself (payload u[0..4]), bit for bit?
__device__ __forceinline__ bool echoes_update(const uint4* u, const Value& v, const Clock& c) {
  const uint4 h = u[0], c0 = u[1], c1 = u[2], v0 = u[3], v1 = u[4];
  uint32_t d = ((h.x & ~(uint32_t)BB_HDR_FLAVOUR_NET) ^ v.meta) | (h.y ^ v.ord) | (h.z ^ c.order);
  d |= (c0.x ^ c.cnt[0]) | (c0.y ^ c.cnt[1]) | (c0.z ^ c.cnt[2]) | (c0.w ^ c.cnt[3]);
  d |= (c1.x ^ c.cnt[4]) | (c1.y ^ c.cnt[5]) | (c1.z ^ c.cnt[6]) | (c1.w ^ c.cnt[7]);
  const uint64_t e = (u64_of(v0.x, v0.y) ^ v.val[0]) | (u64_of(v0.z, v0.w) ^ v.val[1]) | (u64_of(v1.x, v1.y) ^ v.val[2]) |
                     (u64_of(v1.z, v1.w) ^ v.val[3]);
  return d == 0 && e == 0;
}

constexpr uint32_t NO_SLOT = BB_NO_SLOT;
constexpr uint32_t SLOT_ECHO = BB_SLOT_ECHO;
constexpr uint32_t RES_ECHO = 0x10u;  // s_res flag: accepted, no entry
constexpr int MT = 128;    // sorted positions per CTA tile == threads per CTA
constexpr int MT_WARPS = MT / 32;
constexpr int HOT_MIN = 24;    // a segment with this many updates inside one tile goes to k_merge_hot whole
constexpr uint32_t RES_HANDED = 0xFEu;
constexpr int HOT_SERIAL = 8;  // updates of a segment that overruns its tile replayed by the owner thread before k_merge_hot takes over
constexpr int HOT_T = 256;     // threads of a k_merge_hot CTA == its speculation window
constexpr int HOT_WARPS = HOT_T / 32;
constexpr int HOT_CTAS = 148;
// Staged rows sit at their natural 128-byte stride with the 16-byte chunk index XOR-swizzled by the row
// number: conflict-free both for the 8-lanes-per-row copies and for the one-thread-per-row unpack
// (LDS.128 / STS.128 by 32 rows at once), and 2 KB smaller than a padded stride.
__device__ __forceinline__ int row_slot(int r, int c) { return r * ROW_Q + (c ^ (r & 7)); }
// Rows staged by the copy engine (cp.async.bulk, one 128-byte copy per row) land as they are in HBM - no swizzle
// possible - so they sit at a 144-byte stride instead: eight consecutive rows' chunk c fall into eight different
// 16-byte bank groups, conflict-free for the one-thread-per-row unpack as well.
constexpr int ROW_QT = ROW_Q + 1;
template <int TMA>
__device__ __forceinline__ int row_at(int r, int c) { return TMA ? r * ROW_QT + c : row_slot(r, c); }

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
// TMA: 1 = rows, 2 = rows + payloads by bulk copy; EXACT: entry tags + hook events for exact Map / Set query order
template <bool ORDERED, bool INDEXED, bool HOT = false, bool COMPACT = false, int TMA = 0, bool EXACT = false>
__global__ void __launch_bounds__(MT, 7) k_merge_stage(const MergeArgs a) {
  __shared__ __align__(16) uint4 s_upd[MT * UPD_Q];
  __shared__ __align__(16) uint4 s_row[MT * (TMA ? ROW_QT : ROW_Q)];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_idx[MT], s_res[MT];
  __shared__ uint32_t s_hmask[MT_WARPS], s_wsum[MT_WARPS];
  __shared__ uint32_t s_tile, s_over, s_ex, s_nextk;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, wbase = w * 32;
  pdl_launch_dependents();
  pdl_wait();  // the front end is complete
  // bit 0 was decided before this launch.  A rejected batch has no valid item list (out-of-range ids never reach it), so
  // nothing may be loaded through it: return before the first dependent access
  if (*a.rej & 1u) return;
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
  if (TMA < 2) {
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
  }
  if (TMA) {  // rows: one 128-byte bulk copy each, issued by the segment's own thread, all completing on one mbarrier
    if (tid == 0) {
      mbar_init(&s_bar, MT);
      mbar_fence_init();
    }
    __syncthreads();
    const uint32_t tx = (is_head ? ROW_Q * 16u : 0u) + (TMA == 2 && valid ? UPD_Q * 16u : 0u);
    if (tx) mbar_arrive_expect_tx(&s_bar, tx);
    else mbar_arrive(&s_bar);
    if (is_head) bulk_g2s(&s_row[row_at<1>(tid, 0)], a.table + (uint64_t)key * ROW_Q, ROW_Q * 16, &s_bar);
    if (TMA == 2 && valid) {  // the update's three pieces: header 16 B, clock 32 B, values 32 B
      bulk_g2s(&s_upd[tid * UPD_Q], a.head + idx, 16, &s_bar);
      bulk_g2s(&s_upd[tid * UPD_Q + 1], a.clk + 2 * (uint64_t)idx, 32, &s_bar);
      bulk_g2s(&s_upd[tid * UPD_Q + 3], a.val + 2 * (uint64_t)idx, 32, &s_bar);
    }
    if (TMA < 2) cp_async_wait_all();
    mbar_wait(&s_bar, 0);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 4 + (lane >> 3), chunk = lane & 7;
      const uint32_t ekey = __shfl_sync(0xffffffffu, key, e);
      if ((hmask >> e) & 1u) cp_async16(&s_row[row_slot(wbase + e, chunk)], a.table + (uint64_t)ekey * ROW_Q + chunk);
    }
    cp_async_wait_all();
  }
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
    // a hot key: a segment this long is a serial chain for its owner thread; k_merge_hot replays it, a CTA per segment
    bool handed = false;
    if (HOT && end - tid >= HOT_MIN) {
      const uint32_t slot = atomicAdd(a.n_hot, 1u);
      if (slot < a.hot_cap) {
        a.hot_list[slot] = make_uint4(key, (uint32_t)pos, (uint32_t)(pos >> 32), 0u);
        for (int p = tid; p < end; ++p) s_res[p] = RES_HANDED;  // not this tile's to report
        handed = true;
      }
    }
    if (!handed) {
    uint64_t prim[F], prim0[F];  // the node's entries in the dense index columns
    if (INDEXED) {
#pragma unroll
      for (int f = 0; f < F; ++f) prim0[f] = prim[f] = ((a.ix.mask >> f) & 1u) ? a.ix.pcol[f][key] : BB_KEY_NONE;
    }
    RowState r;
    {
      uint4 q[ROW_Q];
#pragma unroll
      for (int c = 0; c < ROW_Q; ++c) q[c] = s_row[row_at<TMA>(tid, c)];
      unpack_row(q, r);
    }
    for (int p = tid; p < end; ++p) {
      uint4* u = &s_upd[p * UPD_Q];
      const uint4 h = u[0];
      Clock c, oc;
      Value x, ov;
      const bool net = unpack_update(h, u[1], u[2], u[3], u[4], c, x);
      const uint32_t code = resolve_step(a.p, r, net, c, x, a.seq_base + s_idx[p], ov, oc);
      if (INDEXED) index_hook<EXACT>(a.ix, key, r.s, x, prim, r.xcnt, a.err, a.seq_base + s_idx[p]);
      uint32_t res = code;
      if (BB_DEC_ACCEPTED(code)) {
        if (COMPACT && echoes_update(u, ov, oc)) res |= RES_ECHO;
        else pack_change(u, h.w, ov, oc);
      }
      s_res[p] = res;
    }
    if (last && nvalid == MT && s_nextk == key) {  // the tile's last segment runs on into the next tiles
      uint32_t over = 0;
      for (uint64_t gp = base + MT; gp < a.n; ++gp) {
        const uint64_t it = a.sorted[gp];
        if ((uint32_t)(it >> 32) != key) break;
        if (HOT && gp - (base + MT) >= (uint64_t)HOT_SERIAL) {  // a hot key: k_merge_hot replays the rest, a CTA per segment
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
        if (INDEXED) index_hook<EXACT>(a.ix, key, r.s, x, prim, r.xcnt, a.err, a.seq_base + ui);
        if (BB_DEC_ACCEPTED(code) && a.epoch_col) a.epoch_col[key] = a.epoch;  // meta[path].lastModified (src/bullet.js:201)
        bool echo = false;
        if (COMPACT && BB_DEC_ACCEPTED(code)) {
          uint4 u5[UPD_Q] = {h, a.clk[2 * (uint64_t)ui], a.clk[2 * (uint64_t)ui + 1], a.val[2 * (uint64_t)ui], a.val[2 * (uint64_t)ui + 1]};
          echo = echoes_update(u5, ov, oc);
        }
        if (echo) {
          a.verdict[ui] = (code << 29) | SLOT_ECHO;
        } else if (BB_DEC_ACCEPTED(code)) {  // compacted into the staging slots this segment owns
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
      for (int c = 0; c < ROW_Q; ++c) s_row[row_at<TMA>(tid, c)] = q[c];
    }
    if (TMA) {  // the row goes home with one bulk store, issued by the thread that wrote it (a handed-over row is untouched)
      fence_proxy_async_smem();
      bulk_s2g(a.table + (uint64_t)key * ROW_Q, &s_row[row_at<1>(tid, 0)], ROW_Q * 16);
      bulk_commit();
    }
    if (INDEXED) {
#pragma unroll
      for (int f = 0; f < F; ++f)
        if (prim[f] != prim0[f]) a.ix.pcol[f][key] = prim[f];
    }
    }  // !handed
  }
  __syncthreads();

  // ---- drain: rank the accepted positions, chain the tile totals
  int first = MT;  // positions before the tile's first head continue a segment an earlier tile owns
#pragma unroll
  for (int ww = MT_WARPS - 1; ww >= 0; --ww)
    if (s_hmask[ww]) first = ww * 32 + __ffs(s_hmask[ww]) - 1;
  const bool owned = valid && tid >= first && s_res[tid] != RES_HANDED;
  const uint32_t res = owned ? s_res[tid] : 0xFFu;
  const uint32_t code = res & ~RES_ECHO;
  const bool echo = COMPACT && owned && (res & RES_ECHO);
  const bool acc = owned && !echo && BB_DEC_ACCEPTED(code);  // emits an entry
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
  if (!TMA) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 4 + (lane >> 3), chunk = lane & 7;
      const uint32_t ekey = __shfl_sync(0xffffffffu, key, e);
      if ((hmask >> e) & 1u) a.table[(uint64_t)ekey * ROW_Q + chunk] = s_row[row_slot(wbase + e, chunk)];
    }
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
  if (owned) a.verdict[idx] = (code << 29) | (acc ? (uint32_t)dest : echo ? SLOT_ECHO : NO_SLOT);
  if ((acc || echo) && a.epoch_col) a.epoch_col[key] = a.epoch;  // meta[path].lastModified (src/bullet.js:201), per merge call
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
  if (TMA) bulk_wait_read_all();  // the bulk stores have read their rows: shared memory may go
}

// ---------------------------------------------------------------- K2h: hot keys
// A path that takes thousands of a batch's updates (Zipf) is a serial chain for the thread that owns it: ~1 us per
// update.  But a network-flavour update is decided by (its clock, M, S) alone - V and the alias flag, the only things
// a REJECTED update changes, do not enter - and M, S change only when an update is accepted.  So one CTA per hot
// segment (handed over by k_merge_stage) stages a WINDOW of the segment's next 256 updates in shared memory and
// evaluates all of them in parallel against the row; everything in front of the first state-changing update (the
// first accepted one, or the first local put, whose clock IS V) is final, that update's own result is exact, its
// thread publishes the row - and the REST OF THE WINDOW is evaluated again against the new row, straight from shared
// memory: a window costs one trip to global memory plus one short pass per state change in it, not one trip per
// state change.
// The post-write index hook (query:139-176) of a retired run is the op sequence R(k0) A(a_j) R(k0) A(a_j+1) ... on the
// node's entry set (k0 = key of the unchanged stored value, a_j = key of update j's value), whose net effect is: k0
// removed, every a_j != k0 present, and k0 present again only if the run's last rejected update carried it.  The
// retiring thread applies exactly that, and a per-segment cache of keys KNOWN PRESENT (shared memory, probed by all
// threads in parallel) lets it skip every update whose key is already in the index - on a hot node nearly all of them.
constexpr int HOT_CACHE = 256;            // slots per field of the known-present cache
constexpr int HOT_LOOK = HOT_T;           // window positions classified per pass (the classification is cheap: all of them)

__device__ __forceinline__ uint32_t hc_hash(uint64_t k) { return (uint32_t)(((k ^ (k >> 29)) * 0x9E3779B97F4A7C15ull) >> 56); }

// The hot node's entries of one field, as a WRITE-BACK set in shared memory: what is known about a key's presence in the
// node's overflow entries.  Adds and removes of a retired run change this set only; the global overflow set is
// brought up to date by all threads in parallel when the segment ends (or the set fills up).
constexpr uint8_t HS_CLEAN = 1;   // in the global set
constexpr uint8_t HS_NEW = 2;     // added here, not in the global set yet
constexpr uint8_t HS_GONE = 3;    // in the global set, removed here
constexpr uint8_t HS_ABSENT = 4;  // in neither

__device__ __forceinline__ int hc_find(const uint64_t* c, uint64_t k) {
  uint32_t i = hc_hash(k);
  for (int probe = 0; probe < HOT_CACHE; ++probe, i = (i + 1) & (HOT_CACHE - 1)) {
    const uint64_t v = c[i];
    if (v == k) return (int)i;
    if (v == BB_KEY_NONE) return -1;
  }
  return -1;
}

// slot of k after inserting it with `state` if it was not there; -1: the set is full
__device__ __forceinline__ int hc_put(uint64_t* c, uint8_t* st, uint32_t& used, uint64_t k, uint8_t state) {
  if (used >= (uint32_t)(HOT_CACHE * 3 / 4)) return -1;
  uint32_t i = hc_hash(k);
  for (int probe = 0; probe < HOT_CACHE; ++probe, i = (i + 1) & (HOT_CACHE - 1)) {
    if (c[i] == BB_KEY_NONE) {
      ++used;
      c[i] = k;
      st[i] = state;
      return (int)i;
    }
  }
  return -1;
}

// the key the hook removes / adds for field f of value v (query:153-167): BB_KEY_NONE when there is nothing to do
__device__ __forceinline__ uint64_t hook_key(const Value& v, int f) {
  if (kind_of(v.meta) != BB_KIND_OBJ) return BB_KEY_NONE;
  const uint32_t t = tag_of(v.meta, f);
  if (t == BB_TAG_ABSENT || prim_falsy(t, v.val[f])) return BB_KEY_NONE;
  return canon_key(t, v.val[f]);
}

// Write the set back (called by every thread of the CTA at a uniform point): HS_NEW keys are inserted into the global
// overflow set, HS_GONE keys removed from it, the node's per-field entry counts follow.  `clear`: forget everything.
__device__ __forceinline__ void hot_flush(const IndexArgs& ix, uint32_t node, uint64_t* cache, uint8_t* state, uint32_t* used, int* dx,
                                          uint32_t* xc, uint32_t* flush, bool clear, uint32_t* err) {
  for (int i = threadIdx.x; i < F * HOT_CACHE; i += HOT_T) {
    const int f = i / HOT_CACHE;
    const uint64_t k = cache[i];
    if (k != BB_KEY_NONE) {
      if (state[i] == HS_NEW) {
        if (x_insert(ix, f, node, k) >= 0) atomicAdd(&dx[f], 1);
        else atomicOr(err, ERR_XFULL);
        state[i] = HS_CLEAN;
      } else if (state[i] == HS_GONE) {
        const int64_t slot = x_find(ix, f, node, k);
        if (slot >= 0) {
          x_remove_at(ix, f, slot);
          atomicAdd(&dx[f], -1);
        }
        state[i] = HS_ABSENT;
      }
      if (clear) cache[i] = BB_KEY_NONE;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t v = *xc;
    for (int f = 0; f < F; ++f) {
      const uint32_t o = xcnt_get(v, f);
      if (o != 0xFFu) v = (v & ~(0xFFu << (8 * f))) | ((uint32_t)max(0, min(255, (int)o + dx[f])) << (8 * f));
      dx[f] = 0;
      if (clear) used[f] = 0;
    }
    *xc = v;
    *flush = 0;
  }
  __syncthreads();
}

// one node's entry set, as the retiring thread of k_merge_hot sees it
struct HotIndex {
  const IndexArgs& ix;
  uint32_t node;
  uint64_t* prim;     // [F] shared: the node's entries in the dense columns
  uint64_t* cache;    // [F][HOT_CACHE] shared: the write-back set's keys ...
  uint8_t* state;     // ... and what is known about each
  uint32_t* used;     // [F] shared
  uint32_t* xc;       // shared: the node's GLOBAL overflow entries per field (8 bits each, 0xFF = many), as of the last flush
  uint32_t* flush;    // shared: set when the set is full (the next pass boundary writes it back and empties it)
  uint32_t* err;

  __device__ __forceinline__ void xc_add(int f, int d) {
    const uint32_t v = xcnt_get(*xc, f);
    if (v == 0xFFu) return;
    const uint32_t n = (uint32_t)max(0, min(255, (int)v + d));
    *xc = (*xc & ~(0xFFu << (8 * f))) | (n << (8 * f));
  }
  __device__ __forceinline__ bool in_global(int f, uint64_t k) { return xcnt_get(*xc, f) && x_find(ix, f, node, k) >= 0; }

  __device__ __forceinline__ void remove(int f, uint64_t k) {
    if (k == BB_KEY_NONE) return;
    if (prim[f] == k) {
      prim[f] = BB_KEY_NONE;
      return;
    }
    uint64_t* c = cache + f * HOT_CACHE;
    uint8_t* st = state + f * HOT_CACHE;
    const int i = hc_find(c, k);
    if (i >= 0) {
      if (st[i] == HS_CLEAN) st[i] = HS_GONE;
      else if (st[i] == HS_NEW) st[i] = HS_ABSENT;
      return;
    }
    const int64_t slot = xcnt_get(*xc, f) ? x_find(ix, f, node, k) : -1;
    if (hc_put(c, st, used[f], k, slot >= 0 ? HS_GONE : HS_ABSENT) < 0) {  // full: straight to the global set
      *flush = 1;
      if (slot >= 0) {
        x_remove_at(ix, f, slot);
        xc_add(f, -1);
      }
    }
  }
  // `looked_up`: the key's own thread has looked it up in the global set (`present`) in parallel with the others of its run
  __device__ __forceinline__ void add(int f, uint64_t k, bool looked_up = false, bool present = false) {
    if (k == BB_KEY_NONE || prim[f] == k) return;
    uint64_t* c = cache + f * HOT_CACHE;
    uint8_t* st = state + f * HOT_CACHE;
    const int i = hc_find(c, k);
    if (i >= 0) {
      if (st[i] == HS_GONE) {
        st[i] = HS_CLEAN;
      } else if (st[i] == HS_ABSENT) {
        if (prim[f] == BB_KEY_NONE) prim[f] = k;
        else st[i] = HS_NEW;
      }
      return;
    }
    if (!looked_up) present = in_global(f, k);
    if (present) {
      if (hc_put(c, st, used[f], k, HS_CLEAN) < 0) *flush = 1;  // (not remembered: it will be looked up again)
      return;
    }
    if (prim[f] == BB_KEY_NONE) {
      prim[f] = k;
      return;
    }
    if (hc_put(c, st, used[f], k, HS_NEW) < 0) {  // full: straight to the global set
      *flush = 1;
      if (x_insert(ix, f, node, k) >= 0) xc_add(f, +1);
      else atomicOr(err, ERR_XFULL);
    }
  }
};

template <bool INDEXED, bool COMPACT = false>
__global__ void __launch_bounds__(HOT_T) k_merge_hot(const MergeArgs a) {
  __shared__ __align__(16) uint4 s_row[ROW_Q];
  __shared__ __align__(16) uint4 s_win[HOT_T * UPD_Q];  // payload window
  __shared__ uint64_t s_prim[F], s_k0[F];
  __shared__ uint64_t s_cache[INDEXED ? F * HOT_CACHE : 1];
  __shared__ uint8_t s_cst[INDEXED ? F * HOT_CACHE : 1];
  __shared__ uint32_t s_used[F], s_xc, s_flush;
  __shared__ int s_dx[F];
  __shared__ uint32_t s_cnt[HOT_WARPS], s_stop[HOT_WARPS], s_need[HOT_WARPS];
  __shared__ unsigned long long s_claim;
  __shared__ uint8_t s_found[INDEXED ? HOT_T : 1];  // per window position: fields whose added key the thread found in the index
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  pdl_launch_dependents();
  pdl_wait();
  if (*a.rej & 1u) return;
  const uint32_t n_hot = min(*a.n_hot, a.hot_cap);
  bool overflow = false;
  for (uint32_t hseg = blockIdx.x; hseg < n_hot; hseg += gridDim.x) {
    const uint4 he = a.hot_list[hseg];
    const uint32_t hkey = he.x;
    uint64_t gp0 = (uint64_t)he.y | ((uint64_t)he.z << 32);
    __syncthreads();
    if (tid < ROW_Q) s_row[tid] = a.table[(uint64_t)hkey * ROW_Q + tid];
    if (INDEXED) {
      if (tid < F) {
        s_prim[tid] = ((a.ix.mask >> tid) & 1u) ? a.ix.pcol[tid][hkey] : BB_KEY_NONE;
        s_used[tid] = 0;
        s_dx[tid] = 0;
      }
      if (tid == 0) s_flush = 0;
      for (int i = tid; i < F * HOT_CACHE; i += HOT_T) s_cache[i] = BB_KEY_NONE;
    }
    __syncthreads();
    if (INDEXED && tid == 0) s_xc = s_row[row_chunk(7)].y;  // the node's global overflow entries per field
    while (true) {  // one window of the segment per turn
      const uint64_t gp = gp0 + tid;
      const uint64_t it = gp < a.n ? a.sorted[gp] : ~0ull;
      const bool mine = gp < a.n && (uint32_t)(it >> 32) == hkey;  // the segment's positions are a prefix of the window
      const uint32_t ui = (uint32_t)it;
      uint4* slot = &s_win[tid * UPD_Q];
      {
        const uint64_t gq = gp + HOT_T;  // the next window's payloads: into L2 while this one is replayed
        if (gq < a.n) {
          const uint64_t iq = a.sorted[gq];
          if ((uint32_t)(iq >> 32) == hkey) {
            const uint32_t uq = (uint32_t)iq;
            prefetch_l2(a.head + uq);
            prefetch_l2(a.clk + 2 * (uint64_t)uq);
            prefetch_l2(a.val + 2 * (uint64_t)uq);
          }
        }
      }
      __syncthreads();  // the previous window's slots (and, first turn, the row) are no longer / now in use
      if (mine) {
        cp_async16(slot, a.head + ui);
        cp_async16(slot + 1, a.clk + 2 * (uint64_t)ui);
        cp_async16(slot + 2, a.clk + 2 * (uint64_t)ui + 1);
        cp_async16(slot + 3, a.val + 2 * (uint64_t)ui);
        cp_async16(slot + 4, a.val + 2 * (uint64_t)ui + 1);
      }
      cp_async_wait_all();
      const uint32_t bm = __ballot_sync(0xffffffffu, mine);
      if (lane == 0) s_cnt[w] = __popc(bm);
      __syncthreads();
      int nseg = 0;
#pragma unroll
      for (int ww = 0; ww < HOT_WARPS; ++ww) nseg += (int)s_cnt[ww];
      if (nseg == 0) break;
      uint4 h = make_uint4(0, 0, 0, 0);
      Clock c;
      Value x;
      bool net = true;
      if (mine) {
        h = slot[0];
        net = unpack_update(h, slot[1], slot[2], slot[3], slot[4], c, x);
      }
      uint64_t akey[F];  // what the hook adds for this update
      if (INDEXED) {
#pragma unroll
        for (int f = 0; f < F; ++f) akey[f] = (mine && ((a.ix.mask >> f) & 1u)) ? hook_key(x, f) : BB_KEY_NONE;
      }
      uint32_t fin = BB_DEC_HISTORICAL;  // this position's final code (| RES_ECHO), known when it retires
      int base = 0;  // window positions [0, base) are retired
      while (base < nseg) {
        const int lim = min(nseg, base + HOT_LOOK);
        const bool live = mine && tid >= base && tid < lim;
        // (1) Cheap and uniform: which of these positions are HISTORICAL for sure?  compareVectorClocks(I, M) < 0 with M
        // present (crt:68-95, 251-263): the update is rejected, S and M stay, it leaves nothing behind but V := the merged
        // clock and alias := 0 (crt:187-197) - nothing a later network update looks at.  Everything else (a local put, a
        // dominating / concurrent / tied clock, a node that does not exist or reads as a falsy primitive and is about to
        // be materialised) is a possible state change: the first such position ends the pass.
        bool hist = false;
        if (live && net) {
          const uint4 m0 = s_row[row_chunk(2)], m1 = s_row[row_chunk(3)], hd = s_row[row_chunk(6)], fl = s_row[row_chunk(7)];
          const uint32_t kind = kind_of(hd.z);
          bool plain = kind == BB_KIND_OBJ;
          if (kind == BB_KIND_PRIM) {
            const uint4 v0 = s_row[row_chunk(0)];
            plain = !prim_falsy(tag_of(hd.z, 0), u64_of(v0.x, v0.y));
          }
          const uint32_t mc[P] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
          bool d1 = false, d2 = false;
#pragma unroll
          for (int sl = 0; sl < P; ++sl) {
            d1 |= c.cnt[sl] > mc[sl];
            d2 |= mc[sl] > c.cnt[sl];
          }
          hist = plain && (fl.x & BB_ROW_M_PRESENT) && d2 && !d1;
        }
        if (INDEXED && tid < F) {  // k0: what a rejected update's hook removes - the node as it reads now
          RowState r0;
          unpack_row(s_row, r0);
          if (kind_of(r0.s.meta) == BB_KIND_NONE || falsy_primitive(r0.s)) materialise_empty_object(r0.s);
          s_k0[tid] = ((a.ix.mask >> tid) & 1u) ? hook_key(r0.s, tid) : BB_KEY_NONE;
        }
        const bool stop = live && !hist;
        const uint32_t bs = __ballot_sync(0xffffffffu, stop);
        if (lane == 0) s_stop[w] = bs;
        __syncthreads();  // every live thread has looked at the row
        int first = HOT_T;
#pragma unroll
        for (int ww = HOT_WARPS - 1; ww >= 0; --ww)
          if (s_stop[ww]) first = ww * 32 + __ffs(s_stop[ww]) - 1;
        // retired this pass: up to and including the first stop
        const int end = first >= lim ? lim : first + 1;
        const bool retiring = live && tid < end;
        // (2) ONE thread runs the resolver: the last retired position - the stop itself, or the last of a run of historical
        // updates (whose V is what the row keeps).  Everything in front of it is historical by construction.
        uint32_t code = BB_DEC_HISTORICAL;
        RowState r;
        Clock oc;
        Value ov;
        if (retiring && tid == end - 1) {
          unpack_row(s_row, r);
          if (!net && tid > base) {
            // A local put's clock is V, and V is what the update in front of it left (crt:187-197 stores the merged
            // clock on every call): the state this put meets is the row with the ONE update in front of it applied.
            const uint4* sp = &s_win[(tid - 1) * UPD_Q];
            Clock cp;
            Value xp;
            const bool netp = unpack_update(sp[0], sp[1], sp[2], sp[3], sp[4], cp, xp);
            resolve_step(a.p, r, netp, cp, xp, a.seq_base + ui, ov, oc);
          }
          code = resolve_step(a.p, r, net, c, x, a.seq_base + ui, ov, oc);
        }
        if (INDEXED) {  // which retired updates does the hook have anything to do for?
          bool need = false;
          if (retiring && tid < end - 1) {
            const uint32_t xc = s_xc;  // the node's GLOBAL overflow entries per field
            uint32_t found = 0;
#pragma unroll
            for (int f = 0; f < F; ++f) {
              const uint64_t k = akey[f];
              if (k == BB_KEY_NONE) continue;
              if (k == s_k0[f]) {
                need = need || tid == end - 2;  // added, then removed again by the next update's hook
              } else if (s_prim[f] != k) {
                const int i = hc_find(s_cache + f * HOT_CACHE, k);
                const uint8_t st = i >= 0 ? s_cst[f * HOT_CACHE + i] : 0;
                if (st == HS_CLEAN || st == HS_NEW) continue;  // known present: nothing to do
                need = true;
                // nothing known about it: look it up in the global set NOW, every thread of the run in parallel (the
                // retiring thread would walk the overflow set once per key, one global round trip after the other)
                if (i < 0 && xcnt_get(xc, f) && x_find(a.ix, f, hkey, k) >= 0) found |= 1u << f;
              }
            }
            s_found[tid] = (uint8_t)found;
          }
          const uint32_t nb = __ballot_sync(0xffffffffu, need);
          if (lane == 0) s_need[w] = nb;
          __syncthreads();
        }
        if (retiring) {
          if (tid == end - 1) {  // its copy is the exact state after the retired updates
            if (INDEXED) {
              HotIndex hx{a.ix, hkey, s_prim, s_cache, s_cst, s_used, &s_xc, &s_flush, a.err};
              if (end - 1 > base) {
#pragma unroll
                for (int f = 0; f < F; ++f)
                  if ((a.ix.mask >> f) & 1u) hx.remove(f, s_k0[f]);
                for (int ww = base >> 5; ww <= (end - 2) >> 5; ++ww) {
                  uint32_t m = s_need[ww];
                  while (m) {
                    const int j = ww * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const uint4* sj = &s_win[j * UPD_Q];
                    Clock cj;
                    Value xj;
                    unpack_update(sj[0], sj[1], sj[2], sj[3], sj[4], cj, xj);
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                      if (!((a.ix.mask >> f) & 1u)) continue;
                      const uint64_t k = hook_key(xj, f);
                      if (k == s_k0[f]) {
                        if (j == end - 2) hx.add(f, k);  // k0 was just removed above: the general path
                        continue;
                      }
                      hx.add(f, k, true, (s_found[j] >> f) & 1u);
                    }
                  }
                }
              }
#pragma unroll
              for (int f = 0; f < F; ++f) {  // this update's own hook, against the node as it reads after it
                if (!((a.ix.mask >> f) & 1u)) continue;
                const uint64_t kr = hook_key(r.s, f);
                if (kr != akey[f]) hx.remove(f, kr);  // (remove k, then add k: k is in the set afterwards either way)
                hx.add(f, akey[f]);
              }
            }
            pack_row(s_row, r);
          }
          // nothing leaves for global memory inside the chain of passes (a slot claim is a round trip to L2): the verdict
          // and, if accepted, the entry - written over the update's own payload slot - wait for the end of the window
          fin = code;
          if (BB_DEC_ACCEPTED(code)) {  // only the last retired one can be
            if (COMPACT && echoes_update(slot, ov, oc)) fin = code | RES_ECHO;
            else pack_change(slot, h.w, ov, oc);
          }
        }
        base = end;
        __syncthreads();  // the published row (and the index cache) are visible; s_stop / s_need may be rewritten
        if (INDEXED && s_flush) hot_flush(a.ix, hkey, s_cache, s_cst, s_used, s_dx, &s_xc, &s_flush, true, a.err);
      }
      {  // ---- the window's results: ONE slot claim, then verdicts and entries
        const bool emit = mine && BB_DEC_ACCEPTED(fin & ~RES_ECHO) && !(fin & RES_ECHO);
        const uint32_t em = __ballot_sync(0xffffffffu, emit);
        const uint32_t am = __ballot_sync(0xffffffffu, mine && BB_DEC_ACCEPTED(fin & ~RES_ECHO));
        if (lane == 0) {
          s_cnt[w] = __popc(em);
          s_stop[w] = am;
        }
        __syncthreads();
        uint32_t before = 0, total = 0, any_acc = 0;
#pragma unroll
        for (int ww = 0; ww < HOT_WARPS; ++ww) {
          if (ww < w) before += s_cnt[ww];
          total += s_cnt[ww];
          any_acc |= s_stop[ww];
        }
        if (tid == 0) {
          s_claim = total ? atomicAdd(reinterpret_cast<unsigned long long*>(a.n_changes), (unsigned long long)total) : 0ull;
          if (any_acc && a.epoch_col) a.epoch_col[hkey] = a.epoch;  // meta[path].lastModified (src/bullet.js:201)
        }
        __syncthreads();
        if (mine) {
          const uint32_t fcode = fin & ~RES_ECHO;
          if (emit) {
            const uint64_t dest = s_claim + before + __popc(em & lanemask_lt());
            a.verdict[ui] = (fcode << 29) | (uint32_t)dest;
            if (dest < a.cap) {
              a.out_idx[dest] = a.idx_base + ui;
              a.out_head[dest] = slot[0];
              a.out_clk[2 * dest] = slot[1];
              a.out_clk[2 * dest + 1] = slot[2];
              a.out_val[2 * dest] = slot[3];
              a.out_val[2 * dest + 1] = slot[4];
            } else {
              overflow = true;
            }
          } else {
            a.verdict[ui] = (fcode << 29) | ((fin & RES_ECHO) ? SLOT_ECHO : NO_SLOT);
          }
        }
      }
      gp0 += (uint64_t)nseg;
      if (nseg < HOT_T) break;  // the segment ended inside this window
    }
    __syncthreads();
    if (INDEXED) {  // the write-back set goes home: every thread takes its share of the inserts and removes
      hot_flush(a.ix, hkey, s_cache, s_cst, s_used, s_dx, &s_xc, &s_flush, false, a.err);
      if (tid == 0) reinterpret_cast<uint32_t*>(&s_row[row_chunk(7)])[1] = s_xc;  // the row's count of overflow entries
      __syncthreads();
    }
    if (tid < ROW_Q) a.table[(uint64_t)hkey * ROW_Q + tid] = s_row[tid];
    if (INDEXED && tid < F && ((a.ix.mask >> tid) & 1u)) a.ix.pcol[tid][hkey] = s_prim[tid];
  }
  if (overflow) atomicOr(a.err, ERR_CHANGES);
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
  table[p * 8 + row_chunk((int)(t & 7))] = rows[t];  // public bb_row chunk order -> the order rows have in HBM
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
  for (int q = 0; q < 8; ++q) rows[i * 8 + q] = row[row_chunk(q)];
}

// ---------------------------------------------------------------- sync producer (src/bullet-network-sync.js:592-664)
// _collectFullSyncData(since): every stored path whose meta.lastModified is not older than `since` (:602, :633: an entry
// is skipped only if since > 0 AND it has a lastModified AND that is < since).  lastModified lives here as the ordinal
// of the merge call that last wrote the row (0 = never written by a merge: loaded, i.e. "no lastModified").  A warp
// looks at 32 rows (header chunk + epoch word), the selected ones are compacted with one atomic per warp, and each
// selected row is copied out by 8 lanes as 128 contiguous bytes, in the public bb_row chunk order.
__global__ void __launch_bounds__(256) k_sync_collect(const uint4* __restrict__ table, const uint32_t* __restrict__ epoch_col,
                                                      uint64_t capacity, uint32_t since, uint32_t filter_records, uint64_t cap,
                                                      uint64_t* __restrict__ out_id, uint4* __restrict__ out_rows,
                                                      uint32_t* __restrict__ out_epoch, unsigned long long* __restrict__ n_out) {
  const int lane = threadIdx.x & 31;
  const uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
  bool take = false;
  uint32_t ep = 0;
  if (row < capacity) {
    const uint32_t meta = reinterpret_cast<const uint32_t*>(table + row * ROW_Q + row_chunk(6))[2];
    ep = epoch_col ? epoch_col[row] : 0u;
    // the reference looks lastModified up at the LEAF path: a record written as a whole has leaves without meta, which
    // `since` therefore never filters (:627-636) - unless the caller asks for the records' own lastModified to count
    const uint32_t kind = kind_of(meta);
    take = kind != BB_KIND_NONE && !((kind == BB_KIND_PRIM || filter_records) && since > 0 && ep != 0 && ep < since);
  }
  const uint32_t m = __ballot_sync(0xffffffffu, take);
  if (m == 0) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(n_out, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  const uint64_t dest = base + __popc(m & ((1u << lane) - 1u));
  if (take && dest < cap) {
    out_id[dest] = row;
    out_epoch[dest] = ep;
  }
  const uint64_t row0 = row - lane;
  for (uint32_t rest = m; rest;) {  // four selected rows per pass, 8 lanes each
    const int g = lane >> 3, c = lane & 7;
    int l = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = rest ? __ffs(rest) - 1 : -1;
      if (k == g) l = b;
      if (rest) rest &= rest - 1;
    }
    if (l >= 0) {
      const uint64_t d = base + __popc(m & ((1u << l) - 1u));
      if (d < cap) out_rows[d * ROW_Q + c] = table[(row0 + l) * ROW_Q + row_chunk(c)];
    }
  }
}

}  // namespace bb
