// bb_group.cuh - the default front end of the merge: GROUP the batch, do not sort it.
//
// The merge kernel (k_merge_stage, bb_kernels.cuh) needs a path's updates adjacent and in arrival order; it does
// not need the paths themselves in order.  So after the count (one atomic per update on a dense per-path counter)
// the batch is only grouped:
//   singles  (the path has one update in the batch: ~2/3 of a uniform batch) keep their arrival order at
//            the front of the item list.  Their payload gathers in the merge kernel then walk the batch
//            arrays almost sequentially and only their 128-byte table rows are accessed at random;
//   multis   the first-counted update of a path claims a run of cnt slots behind the singles with one
//            atomic; k_cg_place drops the path's updates into it, k_cg_fix / k_cs_fix_long put each
//            run in arrival order.
// No pass over the capacity-sized arrays: every kernel is O(batch), and the per-path counters are cleared by
// the threads that used them.  Item-list order (hence change-set layout) is not deterministic;
// BB_CFG_ORDERED_CHANGES keeps the full counting sort.
//
// Measured against the alternative of not building an item list at all (bb_direct.cuh: merge tiles of consecutive
// updates where they lie, multi-update paths in a follow-up kernel): 159 us vs 175 us per 1 M-update batch - the
// follow-up kernel is a latency-bound tail this front end does not have.  The five launches are chained with
// programmatic dependent launch (each kernel's CTAs become resident while its predecessor drains and call
// griddepcontrol.wait before touching its output).
#pragma once
#include "bb_kernels.cuh"

namespace bb {

constexpr uint32_t CG_MULTI = 0x80000000u;
// per-batch counters, two sets used alternately: the count kernel of a batch clears the set of the next one
constexpr int CG_CTR_SINGLE = 0, CG_CTR_MULTI = 1, CG_CTR_LONG = 2, CG_CTR_NEXT = 3, CG_CTR_NHOT = 4, CG_CTR_REJ = 5,
              CG_CTR_WORDS = 8;

// one atomic per update: the returned value is the update's rank inside its path, in atomic order
__global__ void __launch_bounds__(CS_THREADS) k_cg_count(const uint64_t* __restrict__ path_id, uint64_t n,
                                                         uint64_t capacity, uint32_t* __restrict__ cnt,
                                                         uint32_t* __restrict__ rank, uint32_t* __restrict__ ctr,
                                                         uint32_t* __restrict__ ctr_next, uint32_t* __restrict__ err,
                                                         uint32_t ordinal, uint64_t* __restrict__ zero_n,
                                                         const uint4* __restrict__ table_prefetch) {
  pdl_launch_dependents();
  const uint64_t i0 = (uint64_t)blockIdx.x * (CS_THREADS * CS_ILP) + threadIdx.x;
  uint64_t pid[CS_ILP];
  uint32_t r[CS_ILP];
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) pid[k] = i0 + k * CS_THREADS < n ? path_id[i0 + k * CS_THREADS] : 0;
  pdl_wait();  // the previous batch's kernels are complete: counters and counts may be touched
  if (blockIdx.x == 0 && threadIdx.x < CG_CTR_WORDS) ctr_next[threadIdx.x] = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && zero_n) *zero_n = 0;
  bool bad = false;
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) {
    const bool in = i0 + k * CS_THREADS < n;
    bad |= in && pid[k] >= capacity;
    if (table_prefetch && in && pid[k] < capacity) prefetch_l2(table_prefetch + pid[k] * ROW_Q);  // the row, for the merge kernel
    r[k] = (in && pid[k] < capacity) ? atomicAdd(&cnt[pid[k]], 1u) : 0u;
  }
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k)
    if (i0 + k * CS_THREADS < n) rank[i0 + k * CS_THREADS] = r[k];
  if (bad) flag_reject(ctr + CG_CTR_REJ, err, ordinal);  // batch rejected; the later kernels still clear the counts
}

__global__ void __launch_bounds__(CS_THREADS) k_cg_classify(const uint64_t* __restrict__ path_id, uint64_t n,
                                                            uint64_t capacity, uint32_t* __restrict__ cnt,
                                                            uint32_t* __restrict__ rank, uint2* __restrict__ off,
                                                            uint64_t* __restrict__ items, uint32_t* __restrict__ ctr,
                                                            uint2* __restrict__ long_list) {
  __shared__ uint32_t s_w[CS_ILP][CS_THREADS / 32];
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  pdl_launch_dependents();
  const uint64_t i0 = (uint64_t)blockIdx.x * (CS_THREADS * CS_ILP) + tid;
  uint64_t pid[CS_ILP];
  uint32_t r[CS_ILP], c[CS_ILP];
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) {
    const uint64_t i = i0 + k * CS_THREADS;
    pid[k] = i < n ? path_id[i] : ~0ull;
  }
  pdl_wait();  // the counts are final
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) r[k] = i0 + k * CS_THREADS < n ? rank[i0 + k * CS_THREADS] : 0u;
#pragma unroll
  for (int k = 0; k < CS_ILP; ++k) c[k] = pid[k] < capacity ? __ldcg(cnt + pid[k]) : 0u;
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
  pdl_launch_dependents();
  const uint64_t i0 = (uint64_t)blockIdx.x * (CS_THREADS * CS_ILP) + threadIdx.x;
  pdl_wait();
  const uint32_t region = ctr[CG_CTR_SINGLE];  // the multi-update runs start behind the singles
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
  pdl_launch_dependents();
  pdl_wait();
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
