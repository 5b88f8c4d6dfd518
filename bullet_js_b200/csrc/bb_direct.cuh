// bb_direct.cuh - the default merge pipeline: NO sort, three launches per batch.
//
// What the reference does per update (src/bullet.js:139-155 -> src/bullet-crt.js:329-385 -> 164-279 ->
// src/bullet.js:184-220, driven by src/bullet-network-sync.js:551-569) only needs a path's updates replayed
// in arrival order; paths are independent.  In a batch most paths occur once, so:
//
//   K1 k_dm_count   one 64-bit atomicAdd per update on cw[path] adds (arrival index << 32 | 1): the low half
//                   counts the path's updates (the returned value is this update's rank inside the path, in
//                   atomic order), the high half sums their arrival indices mod 2^32.  The third arrival of a
//                   path claims a slab of 8 index slots, the ninth queues the path as "long"; the last CTA to
//                   finish lays the long paths' runs out.
//   K2 k_dm_merge   one CTA per tile of 256 CONSECUTIVE updates (arrival order).  The tile's payloads
//                   (16 + 32 + 32 bytes per update) are staged in shared memory by three cp.async.bulk
//                   copies (TMA engine, mbarrier completion) issued before anything else.  Per update:
//                     count 1 (2/3 of a uniform batch)  row -> registers (8 x 16 B, the fastest way to move a
//                                   random 128-byte row: scripts/ubench/row_gather.cu), resolve, row back,
//                                   change entry ranked inside the tile (arrival order), one atomic per CTA
//                     count 2       partner = index sum - own index: the LATER update owns the path, fetches the
//                                   partner's payload (cp.async) next to the row and replays both in order
//                     count 3..8    drops its index into the path's slab; > 8: into the path's run
//   K3 k_dm_multi   one thread per slab (sort <= 8 indices in registers, replay), then one CTA per long path:
//                   radix sort of its run on the arrival index, then rounds of 256 speculative evaluations
//                   against the row in shared memory - everything in front of the first state-changing update
//                   is final (a Zipf batch's hot keys are serial chains by definition, but not every link is)
//
// K2 and K3 are launched with programmatic stream serialisation: their CTAs become resident while the
// previous kernel drains and wait (griddepcontrol.wait) before touching its output; K2 stages payloads first.
// cw[] is all zero between batches: whoever replays a path clears its word.
#pragma once
#include "bb_kernels.cuh"

namespace bb {

constexpr int DM_SHORT = 8;    // longest path one thread replays out of a slab
constexpr int DM_T = 256;      // K2: updates per tile == threads per CTA
constexpr int DM_WARPS = DM_T / 32;
constexpr int DM_ILP = 4;      // K1: atomics in flight per thread
constexpr int DM3_T = 256;     // K3: threads per CTA == speculation window of a long path
constexpr int DM3_WARPS = DM3_T / 32;
// per-batch counters; two sets, used alternately: K1 of a batch clears the set of the next one
constexpr int DC_NSLAB = 0, DC_NLONG = 1, DC_TICKET = 2, DC_LTOTAL = 3, DC_REJ = 4, DC_WORDS = 8;

struct DmArgs {
  const uint64_t* path_id;
  uint64_t n, capacity;
  uint4* table;
  const uint4* head;       // [n]
  const uint4* clk;        // [n][2]
  const uint4* val;        // [n][2]
  unsigned long long* cw;  // [capacity] (sum of arrival indices mod 2^32) << 32 | updates of this path; 0 between batches
  uint32_t* off;           // [capacity] this batch only: slab of a path with 3..8 updates, start of its run if > 8
  uint32_t* rank;          // [n]
  uint32_t* slab;          // [n / 3 + 1][8]
  uint32_t* slab_pid;      // [n / 3 + 1]
  uint32_t* long_pid;      // [n / 9 + 1]
  uint32_t* litems;        // [n] runs of the long paths
  uint32_t* lscratch;      // [n] ping-pong buffer of their radix sort
  uint32_t* ctr;           // [DC_WORDS] this batch
  uint32_t* ctr_next;      // [DC_WORDS] the next batch's set: cleared by K1
  uint32_t* verdict;       // [n] arrival order: code << 29 | slot
  unsigned long long* n_changes;
  uint32_t* out_idx;
  uint4* out_head;
  uint4* out_clk;
  uint4* out_val;
  uint64_t cap;
  uint64_t seq_base;
  uint32_t idx_base;       // added to the arrival indices this launch reports (chunked host calls)
  uint32_t zero_changes;   // K1 clears *n_changes (a new change set starts with this launch)
  uint32_t ordinal;        // number of this batch since the last bb_sync (error reporting)
  uint32_t* err;           // sticky until bb_sync: [0] bits ERR_*, [1] ordinal of the first rejected batch
  const uint32_t* rej;     // bit 0: this batch is rejected (ctr[DC_REJ], or the whole-call word of a chunked host call)
  Params p;
  IndexArgs ix;
};

// ---------------------------------------------------------------- K1
__global__ void __launch_bounds__(256) k_dm_count(const DmArgs a) {
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x, lane = tid & 31;
  if (blockIdx.x == 0) {
    if (tid < DC_WORDS) a.ctr_next[tid] = 0;
    if (tid == 0 && a.zero_changes) *a.n_changes = 0;
  }
  pdl_launch_dependents();  // K2 may start staging its payload tiles
  const uint64_t i0 = (uint64_t)blockIdx.x * (256 * DM_ILP) + tid;
  uint64_t pid[DM_ILP];
  uint32_t r[DM_ILP];
  bool ok[DM_ILP];
#pragma unroll
  for (int k = 0; k < DM_ILP; ++k) pid[k] = i0 + k * 256 < a.n ? a.path_id[i0 + k * 256] : ~0ull;
  bool bad = false;
#pragma unroll
  for (int k = 0; k < DM_ILP; ++k) {
    const uint64_t i = i0 + k * 256;
    ok[k] = i < a.n && pid[k] < a.capacity;
    bad |= i < a.n && !ok[k];
    r[k] = ok[k] ? (uint32_t)atomicAdd(&a.cw[pid[k]], ((unsigned long long)(uint32_t)i << 32) | 1ull) : 0u;
  }
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int k = 0; k < DM_ILP; ++k) {
    if (i0 + k * 256 < a.n) a.rank[i0 + k * 256] = r[k];
    const bool third = ok[k] && r[k] == 2u, ninth = ok[k] && r[k] == (uint32_t)DM_SHORT;
    const uint32_t m3 = __ballot_sync(0xffffffffu, third), m9 = __ballot_sync(0xffffffffu, ninth);
    if (m3) {
      uint32_t base = 0;
      if (lane == __ffs(m3) - 1) base = atomicAdd(&a.ctr[DC_NSLAB], (uint32_t)__popc(m3));
      base = __shfl_sync(0xffffffffu, base, __ffs(m3) - 1);
      if (third) {
        const uint32_t s = base + __popc(m3 & lt);
        a.off[pid[k]] = s;
        a.slab_pid[s] = (uint32_t)pid[k];
      }
    }
    if (m9) {
      uint32_t base = 0;
      if (lane == __ffs(m9) - 1) base = atomicAdd(&a.ctr[DC_NLONG], (uint32_t)__popc(m9));
      base = __shfl_sync(0xffffffffu, base, __ffs(m9) - 1);
      if (ninth) a.long_pid[base + __popc(m9 & lt)] = (uint32_t)pid[k];
    }
  }
  if (bad) flag_reject(a.ctr + DC_REJ, a.err, a.ordinal);

  // the last CTA to get here lays out the runs of the long paths (usually there are none)
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&a.ctr[DC_TICKET], 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const uint32_t nlong = ld_volatile(a.ctr + DC_NLONG);
  uint32_t run = 0;
  for (uint32_t k0 = 0; k0 < nlong; k0 += 256) {
    const uint32_t k = k0 + tid;
    const uint32_t p = k < nlong ? __ldcg(a.long_pid + k) : 0u;
    const uint32_t c = k < nlong ? (uint32_t)__ldcg(a.cw + p) : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan<256>(c, &total);
    if (k < nlong) a.off[p] = run + ex;  // replaces the slab index the path's third arrival stored
    run += total;
  }
  if (tid == 0) a.ctr[DC_LTOTAL] = run;
}

// ---------------------------------------------------------------- K2
__device__ __forceinline__ void load_row_regs(const uint4* row, RowState& r) {
  uint4 q[ROW_Q];
#pragma unroll
  for (int c = 0; c < ROW_Q; ++c) q[c] = ld_stream16(row + c);
  unpack_row(q, r);
}
__device__ __forceinline__ void store_row_regs(uint4* row, const RowState& r) {
  uint4 q[ROW_Q];
  pack_row(q, r);
#pragma unroll
  for (int c = 0; c < ROW_Q; ++c) row[c] = q[c];
}

template <bool INDEXED>
__global__ void __launch_bounds__(DM_T, 3) k_dm_merge(const DmArgs a) {
  __shared__ __align__(128) uint4 s_head[DM_T];      // the tile's payloads; an accepted update's slot is
  __shared__ __align__(128) uint4 s_clk[2 * DM_T];   // overwritten with its change entry
  __shared__ __align__(128) uint4 s_val[2 * DM_T];
  __shared__ __align__(16) uint4 s_pp[DM_T * UPD_Q];  // partner payload of a 2-update path (slot of its owner)
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_wsum[DM_WARPS];
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint64_t base = (uint64_t)blockIdx.x * DM_T;
  const uint64_t i = base + tid;
  const bool valid = i < a.n;
  const uint32_t nvalid = (uint32_t)min((uint64_t)DM_T, a.n - base);

  // ---- stage: the tile's payloads, three bulk copies; nothing here depends on K1
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&s_bar, nvalid * 80u);
    bulk_g2s(s_head, a.head + base, nvalid * 16u, &s_bar);
    bulk_g2s(s_clk, a.clk + 2 * base, nvalid * 32u, &s_bar);
    bulk_g2s(s_val, a.val + 2 * base, nvalid * 32u, &s_bar);
  }
  const uint64_t pid = valid ? a.path_id[i] : ~0ull;
  pdl_launch_dependents();
  __syncthreads();  // the barrier word is initialised for everybody
  pdl_wait();       // K1 is complete: counts, ranks, slabs, runs
  if (*a.rej & 1u) {  // rejected batch: the table stays as it is; only K1's counts are undone
    if (pid < a.capacity) a.cw[pid] = 0ull;
    mbar_wait(&s_bar, 0);  // no copy may be in flight into this CTA's shared memory when it exits
    return;
  }
  const unsigned long long cwv = valid ? __ldcg(a.cw + pid) : 0ull;
  const uint32_t cnt = (uint32_t)cwv;
  const uint32_t partner = (uint32_t)(cwv >> 32) - (uint32_t)i;  // meaningful when cnt == 2
  const bool single = cnt == 1u;
  const bool owner = cnt == 2u && (uint32_t)i > partner;
  const bool replay = single || owner;  // this thread replays the path

  RowState r;
  uint64_t prim[F], prim0[F];
  if (replay) {
    load_row_regs(a.table + pid * ROW_Q, r);
    if (INDEXED) {
#pragma unroll
      for (int f = 0; f < F; ++f) prim0[f] = prim[f] = ((a.ix.mask >> f) & 1u) ? a.ix.pcol[f][pid] : BB_KEY_NONE;
    }
    if (owner) {
      uint4* pp = &s_pp[tid * UPD_Q];
      cp_async16(pp, a.head + partner);
      cp_async16(pp + 1, a.clk + 2 * (uint64_t)partner);
      cp_async16(pp + 2, a.clk + 2 * (uint64_t)partner + 1);
      cp_async16(pp + 3, a.val + 2 * (uint64_t)partner);
      cp_async16(pp + 4, a.val + 2 * (uint64_t)partner + 1);
    }
  } else if (cnt > 2u) {  // 3..8: the path's slab; more: its run.  K3 replays it
    const uint32_t rk = a.rank[i], o = a.off[pid];
    if (cnt <= (uint32_t)DM_SHORT) a.slab[(uint64_t)o * DM_SHORT + rk] = (uint32_t)i;
    else a.litems[o + rk] = (uint32_t)i;
  }
  cp_async_wait_all();
  mbar_wait(&s_bar, 0);

  // ---- resolve: pass 0 = the partner (owners only), pass 1 = the thread's own update
  uint32_t code0 = 0xFFu, code1 = 0xFFu;
  if (replay) {
#pragma unroll 1
    for (int pass = owner ? 0 : 1; pass < 2; ++pass) {
      uint4 h, c0, c1, v0, v1;
      if (pass == 0) {
        const uint4* pp = &s_pp[tid * UPD_Q];
        h = pp[0]; c0 = pp[1]; c1 = pp[2]; v0 = pp[3]; v1 = pp[4];
      } else {
        h = s_head[tid]; c0 = s_clk[2 * tid]; c1 = s_clk[2 * tid + 1]; v0 = s_val[2 * tid]; v1 = s_val[2 * tid + 1];
      }
      Clock c, oc;
      Value x, ov;
      const bool net = unpack_update(h, c0, c1, v0, v1, c, x);
      const uint32_t ui = pass == 0 ? partner : (uint32_t)i;
      const uint32_t code = resolve_step(a.p, r, net, c, x, a.seq_base + ui, ov, oc);
      if (INDEXED) index_hook(a.ix, (uint32_t)pid, r.s, x, prim, r.xcnt, a.err);
      if (BB_DEC_ACCEPTED(code)) {
        uint4 q[UPD_Q];
        pack_change(q, h.w, ov, oc);
        if (pass == 0) {
          uint4* pp = &s_pp[tid * UPD_Q];
#pragma unroll
          for (int k = 0; k < UPD_Q; ++k) pp[k] = q[k];
        } else {
          s_head[tid] = q[0]; s_clk[2 * tid] = q[1]; s_clk[2 * tid + 1] = q[2]; s_val[2 * tid] = q[3]; s_val[2 * tid + 1] = q[4];
        }
      }
      if (pass == 0) code0 = code;
      else code1 = code;
    }
    store_row_regs(a.table + pid * ROW_Q, r);
    a.cw[pid] = 0ull;
    if (INDEXED) {
#pragma unroll
      for (int f = 0; f < F; ++f)
        if (prim[f] != prim0[f]) a.ix.pcol[f][pid] = prim[f];
    }
  }

  // ---- rank the tile's accepted entries (partner before owner: arrival order), one claim per CTA
  const uint32_t acc0 = (code0 != 0xFFu && BB_DEC_ACCEPTED(code0)) ? 1u : 0u;
  const uint32_t acc1 = (code1 != 0xFFu && BB_DEC_ACCEPTED(code1)) ? 1u : 0u;
  const uint32_t mine = acc0 + acc1;
  const uint32_t inc = warp_inclusive_scan(mine);
  if (lane == 31) s_wsum[w] = inc;
  __syncthreads();
  uint32_t before = inc - mine, total = 0;
#pragma unroll
  for (int ww = 0; ww < DM_WARPS; ++ww) {
    const uint32_t c = s_wsum[ww];
    if (ww < w) before += c;
    total += c;
  }
  if (tid == 0) s_base = total ? atomicAdd(a.n_changes, (unsigned long long)total) : 0ull;
  __syncthreads();
  const uint64_t d0 = s_base + before, d1 = d0 + acc0;
  bool overflow = false;
  if (replay) {
    a.verdict[i] = (code1 << 29) | (acc1 ? (uint32_t)d1 : NO_SLOT);
    if (owner) a.verdict[partner] = (code0 << 29) | (acc0 ? (uint32_t)d0 : NO_SLOT);
    if (acc0) {
      if (d0 < a.cap) {
        const uint4* pp = &s_pp[tid * UPD_Q];
        a.out_idx[d0] = a.idx_base + partner;
        a.out_head[d0] = pp[0];
        a.out_clk[2 * d0] = pp[1];
        a.out_clk[2 * d0 + 1] = pp[2];
        a.out_val[2 * d0] = pp[3];
        a.out_val[2 * d0 + 1] = pp[4];
      } else {
        overflow = true;
      }
    }
    if (acc1) {
      if (d1 < a.cap) {
        a.out_idx[d1] = a.idx_base + (uint32_t)i;
        a.out_head[d1] = s_head[tid];
        a.out_clk[2 * d1] = s_clk[2 * tid];
        a.out_clk[2 * d1 + 1] = s_clk[2 * tid + 1];
        a.out_val[2 * d1] = s_val[2 * tid];
        a.out_val[2 * d1 + 1] = s_val[2 * tid + 1];
      } else {
        overflow = true;
      }
    }
  }
  if (overflow) atomicOr(a.err, ERR_CHANGES);
}

// ---------------------------------------------------------------- K3
// CTA-wide stable LSD radix sort of `len` distinct u32 keys (arrival indices), 8-bit digits, ping-ponging between
// `cur` and `oth`; returns the buffer that holds the sorted keys.  Digits every key shares are skipped.
__device__ __forceinline__ uint32_t* cta_radix_sort_u32(uint32_t* cur, uint32_t* oth, uint32_t len, uint32_t max_key) {
  static_assert(DM3_T == RADIX, "one thread per digit");
  __shared__ uint32_t hist[RADIX];
  __shared__ uint32_t whist[DM3_WARPS][RADIX];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t lt = lanemask_lt();
  for (int shift = 0; shift < 32 && (max_key >> shift) != 0u; shift += 8) {
    __syncthreads();
    hist[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < len; i += DM3_T) atomicAdd(&hist[(cur[i] >> shift) & 0xFFu], 1u);
    __syncthreads();
    const uint32_t mine = hist[tid];
    const bool trivial = __syncthreads_or(mine == len);  // every key has the same digit: nothing to move
    if (trivial) continue;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan<DM3_T>(mine, &total);
    hist[tid] = ex;  // running base of digit `tid`
    __syncthreads();
    for (uint32_t c0 = 0; c0 < len; c0 += DM3_T * 8) {  // chunks in order, warps own contiguous runs: stable
      for (int d = lane; d < RADIX; d += 32) whist[w][d] = 0;
      __syncwarp();
      const uint32_t wb = c0 + w * 256;
      uint32_t kv[8], rk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t i = wb + k * 32 + lane;
        const bool valid = i < len;
        kv[k] = valid ? cur[i] : 0u;
        const uint32_t d = valid ? ((kv[k] >> shift) & 0xFFu) : (uint32_t)RADIX;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
          old = whist[w][d];
          whist[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rk[k] = old + __popc(peers & lt);
        __syncwarp();
      }
      __syncthreads();
      {  // digit `tid`: offsets of the warps' runs, then advance the running base
        uint32_t run = hist[tid];
#pragma unroll
        for (int ww = 0; ww < DM3_WARPS; ++ww) {
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
        if (i < len) oth[whist[w][(kv[k] >> shift) & 0xFFu] + rk[k]] = kv[k];
      }
      __syncthreads();
    }
    uint32_t* t = cur;
    cur = oth;
    oth = t;
  }
  __syncthreads();
  return cur;
}

// 8 values ascending, registers only (0xFFFFFFFF pads sort to the end)
__device__ __forceinline__ void sort8(uint32_t (&v)[DM_SHORT]) {
#pragma unroll
  for (int x = 1; x < DM_SHORT; ++x) {
#pragma unroll
    for (int y = x; y > 0; --y) {
      const uint32_t lo = min(v[y - 1], v[y]), hi = max(v[y - 1], v[y]);
      v[y - 1] = lo;
      v[y] = hi;
    }
  }
}

__device__ __forceinline__ void emit_entry(const DmArgs& a, uint64_t dest, uint32_t ui, const uint4 (&q)[UPD_Q], bool& overflow) {
  if (dest < a.cap) {
    a.out_idx[dest] = a.idx_base + ui;
    a.out_head[dest] = q[0];
    a.out_clk[2 * dest] = q[1];
    a.out_clk[2 * dest + 1] = q[2];
    a.out_val[2 * dest] = q[3];
    a.out_val[2 * dest + 1] = q[4];
  } else {
    overflow = true;
  }
}

template <bool INDEXED>
__global__ void __launch_bounds__(DM3_T) k_dm_multi(const DmArgs a) {
  __shared__ uint32_t s_idx[DM_SHORT][DM3_T];  // a slab thread's sorted indices (dynamic index k without local memory)
  __shared__ __align__(16) uint4 s_row[ROW_Q];
  __shared__ __align__(16) uint4 s_win[DM3_T * UPD_Q];  // payload window of a long path
  __shared__ uint64_t s_prim[F];
  __shared__ uint32_t s_cnt[DM3_WARPS], s_stop[DM3_WARPS], s_loc[DM3_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  pdl_wait();  // K2 (and K1 before it) is complete
  if (*a.rej & 1u) return;
  const uint32_t lt = lanemask_lt();
  bool overflow = false;

  // ---- phase A: paths with 3..8 updates, one thread each
  const uint32_t nslab = a.ctr[DC_NSLAB];
  for (uint32_t s0 = blockIdx.x * DM3_T; s0 < nslab; s0 += gridDim.x * DM3_T) {
    const uint32_t s = s0 + tid;
    bool active = s < nslab;
    const uint32_t pid = active ? a.slab_pid[s] : 0u;
    uint32_t cnt = active ? (uint32_t)__ldcg(a.cw + pid) : 0u;
    if (cnt > (uint32_t)DM_SHORT) active = false;  // became a long path: phase B
    if (!active) cnt = 0;
    RowState r;
    uint64_t prim[F], prim0[F];
    if (active) {
      const uint4* sl = reinterpret_cast<const uint4*>(a.slab + (uint64_t)s * DM_SHORT);
      const uint4 lo = sl[0], hi = sl[1];
      uint32_t v[DM_SHORT] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int k = 0; k < DM_SHORT; ++k)
        if (k >= (int)cnt) v[k] = 0xFFFFFFFFu;
      sort8(v);
#pragma unroll
      for (int k = 0; k < DM_SHORT; ++k) s_idx[k][tid] = v[k];
      load_row_regs(a.table + (uint64_t)pid * ROW_Q, r);
      if (INDEXED) {
#pragma unroll
        for (int f = 0; f < F; ++f) prim0[f] = prim[f] = ((a.ix.mask >> f) & 1u) ? a.ix.pcol[f][pid] : BB_KEY_NONE;
      }
    }
    uint4 nh, nc0, nc1, nv0, nv1;  // the next update's payload, fetched one step ahead
    if (cnt > 0u) {
      const uint32_t ui = s_idx[0][tid];
      nh = ld_stream16(a.head + ui);
      nc0 = ld_stream16(a.clk + 2 * (uint64_t)ui); nc1 = ld_stream16(a.clk + 2 * (uint64_t)ui + 1);
      nv0 = ld_stream16(a.val + 2 * (uint64_t)ui); nv1 = ld_stream16(a.val + 2 * (uint64_t)ui + 1);
    }
#pragma unroll 1
    for (int k = 0; k < DM_SHORT; ++k) {
      const bool on = (uint32_t)k < cnt;
      if (!__any_sync(0xffffffffu, on)) break;
      uint32_t code = 0xFFu, ui = 0;
      uint4 q[UPD_Q];
      if (on) {
        ui = s_idx[k][tid];
        const uint4 h = nh, c0 = nc0, c1 = nc1, v0 = nv0, v1 = nv1;
        if ((uint32_t)(k + 1) < cnt) {
          const uint32_t un = s_idx[k + 1][tid];
          nh = ld_stream16(a.head + un);
          nc0 = ld_stream16(a.clk + 2 * (uint64_t)un); nc1 = ld_stream16(a.clk + 2 * (uint64_t)un + 1);
          nv0 = ld_stream16(a.val + 2 * (uint64_t)un); nv1 = ld_stream16(a.val + 2 * (uint64_t)un + 1);
        }
        Clock c, oc;
        Value x, ov;
        const bool net = unpack_update(h, c0, c1, v0, v1, c, x);
        code = resolve_step(a.p, r, net, c, x, a.seq_base + ui, ov, oc);
        if (INDEXED) index_hook(a.ix, pid, r.s, x, prim, r.xcnt, a.err);
        if (BB_DEC_ACCEPTED(code)) pack_change(q, h.w, ov, oc);
      }
      const bool acc = on && BB_DEC_ACCEPTED(code);
      const uint32_t am = __ballot_sync(0xffffffffu, acc);
      unsigned long long wb = 0;
      if (am) {
        if (lane == __ffs(am) - 1) wb = atomicAdd(a.n_changes, (unsigned long long)__popc(am));
        wb = __shfl_sync(0xffffffffu, wb, __ffs(am) - 1);
      }
      if (on) {
        const uint64_t dest = wb + __popc(am & lt);
        a.verdict[ui] = (code << 29) | (acc ? (uint32_t)dest : NO_SLOT);
        if (acc) emit_entry(a, dest, ui, q, overflow);
      }
    }
    if (active) {
      store_row_regs(a.table + (uint64_t)pid * ROW_Q, r);
      a.cw[pid] = 0ull;
      if (INDEXED) {
#pragma unroll
        for (int f = 0; f < F; ++f)
          if (prim[f] != prim0[f]) a.ix.pcol[f][pid] = prim[f];
      }
    }
  }

  // ---- phase B: long paths (hot keys), one CTA each.  A network-flavour update is decided by (its clock, M, S)
  // alone - V and the alias flag, the only things a REJECTED update changes, do not enter - and M, S change only
  // when an update is accepted.  So all DM3_T updates of a window are evaluated in parallel against the row in
  // shared memory; everything in front of the first state-changing update (the first accepted one, or the first
  // local put, whose clock IS V) is final, that update's own result is exact, and its thread publishes the row.
  const uint32_t nlong = a.ctr[DC_NLONG];
  for (uint32_t hseg = blockIdx.x; hseg < nlong; hseg += gridDim.x) {
    const uint32_t pid = a.long_pid[hseg];
    const uint32_t len = (uint32_t)__ldcg(a.cw + pid), start = a.off[pid];
    const uint32_t* run = cta_radix_sort_u32(a.litems + start, a.lscratch + start, len, (uint32_t)(a.n - 1));
    if (tid < ROW_Q) s_row[tid] = a.table[(uint64_t)pid * ROW_Q + tid];
    if (INDEXED && tid < F) s_prim[tid] = ((a.ix.mask >> tid) & 1u) ? a.ix.pcol[tid][pid] : BB_KEY_NONE;
    __syncthreads();
    uint32_t gp0 = 0;
    while (gp0 < len) {
      const uint32_t gp = gp0 + tid;
      const bool mine = gp < len;
      const uint32_t ui = mine ? run[gp] : 0u;
      uint4* slot = &s_win[tid * UPD_Q];
      if (mine) {
        cp_async16(slot, a.head + ui);
        cp_async16(slot + 1, a.clk + 2 * (uint64_t)ui);
        cp_async16(slot + 2, a.clk + 2 * (uint64_t)ui + 1);
        cp_async16(slot + 3, a.val + 2 * (uint64_t)ui);
        cp_async16(slot + 4, a.val + 2 * (uint64_t)ui + 1);
      }
      cp_async_wait_all();
      uint32_t code = 0;
      bool net = true;
      uint4 h = make_uint4(0, 0, 0, 0);
      RowState r;
      Clock oc;
      Value ov;
      if (mine) {
        h = slot[0];
        Clock c;
        Value x;
        net = unpack_update(h, slot[1], slot[2], slot[3], slot[4], c, x);
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
      __syncthreads();  // also: every thread has unpacked the row and its window slot is in shared memory
      int nseg = 0, first = DM3_T;
      bool f_local = false;
#pragma unroll
      for (int ww = DM3_WARPS - 1; ww >= 0; --ww) {
        nseg += (int)s_cnt[ww];
        if (s_stop[ww]) {
          const int b = __ffs(s_stop[ww]) - 1;
          first = ww * 32 + b;
          f_local = (s_loc[ww] >> b) & 1u;
        }
      }
      // retired this round: up to and including the first stop - unless that is a local put further in, whose
      // clock depends on the V the updates in front of it leave: it waits for the next round's position 0
      const int retired = first >= nseg ? nseg : ((f_local && first > 0) ? first : first + 1);
      if (tid < retired) {
        if (tid == retired - 1) {  // its copy is the exact state after the retired updates
          if (INDEXED) {
            // _updateIndices runs after EVERY setData, in order (query:139-176): the retired updates in front
            // of this one were rejected network updates, after which the node reads as the round's initial S
            RowState r0;
            unpack_row(s_row, r0);
            if (kind_of(r0.s.meta) == BB_KIND_NONE || falsy_primitive(r0.s)) materialise_empty_object(r0.s);
            uint64_t prim[F];
#pragma unroll
            for (int f = 0; f < F; ++f) prim[f] = s_prim[f];
            uint32_t xcnt = r0.xcnt;
            for (int j = 0; j < retired; ++j) {
              const uint4* sj = &s_win[j * UPD_Q];
              Clock cj;
              Value xj;
              unpack_update(sj[0], sj[1], sj[2], sj[3], sj[4], cj, xj);
              index_hook(a.ix, pid, j == retired - 1 ? r.s : r0.s, xj, prim, xcnt, a.err);
            }
            r.xcnt = xcnt;
#pragma unroll
            for (int f = 0; f < F; ++f) s_prim[f] = prim[f];
          }
          pack_row(s_row, r);
        }
        if (BB_DEC_ACCEPTED(code)) {  // only the last retired one can be
          const uint64_t dest = atomicAdd(a.n_changes, 1ull);
          a.verdict[ui] = (code << 29) | (uint32_t)dest;
          uint4 q[UPD_Q];
          pack_change(q, h.w, ov, oc);
          emit_entry(a, dest, ui, q, overflow);
        } else {
          a.verdict[ui] = (code << 29) | NO_SLOT;
        }
      }
      gp0 += (uint32_t)retired;
      __syncthreads();  // the published row is visible; the window and s_cnt / s_stop / s_loc may be rewritten
    }
    if (tid < ROW_Q) a.table[(uint64_t)pid * ROW_Q + tid] = s_row[tid];
    if (INDEXED && tid < F && ((a.ix.mask >> tid) & 1u)) a.ix.pcol[tid][pid] = s_prim[tid];
    if (tid == 0) a.cw[pid] = 0ull;
    __syncthreads();
  }
  if (overflow) atomicOr(a.err, ERR_CHANGES);
}

}  // namespace bb
