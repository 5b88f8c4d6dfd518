// bb_direct.cuh - the default merge pipeline: NO sort, three launches per batch.
//
// What the reference does per update (src/bullet.js:139-155 -> src/bullet-crt.js:329-385 -> 164-279 ->
// src/bullet.js:184-220, driven by src/bullet-network-sync.js:551-569) only needs a path's updates replayed
// in arrival order; paths are independent.  In a batch most paths occur once, so:
//
//   K1 k_dm_count   one 64-bit atomicAdd per update ON THE PATH'S TABLE ROW: its last 16-byte chunk holds
//                   flags | count << 3 and, next to it, the sum of the arrival indices mod 2^32 (both 0 between
//                   batches).  The returned value is the update's rank inside its path (atomic order).  Random
//                   row accesses are DRAM-activation-bound (~30 us per million, scripts/ubench/atomics.cu), so
//                   this kernel is where the batch's rows are pulled into L2: persistent CTAs prefetch the row
//                   lines of their next tile while the atomics of the current one run, and walk the tiles in
//                   REVERSE, so that the rows K2 wants first are the freshest.  The second arrival of a path
//                   notes the first one's index (its "partner": the sum so far IS that index) and prefetches
//                   its payload, the third claims a slab of 8 index slots, the ninth queues the path as
//                   "long"; the last CTA lays the long paths' runs out.
//   K2 k_dm_merge   one CTA per tile of 128 CONSECUTIVE updates (arrival order).  The tile's payloads
//                   (16 + 32 + 32 bytes per update) are staged in shared memory by three cp.async.bulk
//                   copies (TMA engine, mbarrier completion) issued before anything else; every update's row
//                   follows with cp.async, 8 lanes per 128-byte row (a divergent 16-byte load per thread
//                   costs 8x the L1 wavefronts: scripts/ubench/row_gather.cu), and carries the path's count:
//                     count 1 (2/3 of a uniform batch)  resolve, row back (count cleared), change entry ranked
//                                   inside the tile, one atomic per CTA
//                     count 2       the SECOND arrival owns the path: it fetched the partner's payload next to
//                                   the row (speculatively, from K1's note) and replays both in arrival order
//                     count 3..8    drops its index into the path's slab; > 8: into the path's run
//   K3 k_dm_multi   one thread per slab (sort <= 8 indices in registers, replay), then one CTA per long path:
//                   radix sort of its run on the arrival index, then rounds of 256 speculative evaluations
//                   against the row in shared memory - everything in front of the first state-changing update
//                   is final (a Zipf batch's hot keys are serial chains by definition, but not every link is)
//
// K2 and K3 are launched with programmatic stream serialisation: their CTAs become resident while the
// previous kernel drains and wait (griddepcontrol.wait) before touching its output; K2 stages payloads first.
#pragma once
#include "bb_kernels.cuh"

namespace bb {

constexpr int DM_SHORT = 8;    // longest path one thread replays out of a slab
constexpr int DM_T = 128;      // K2: updates per tile == threads per CTA
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
  uint32_t* off;           // [capacity] this batch only: slab of a path with 3..8 updates, start of its run if > 8
  uint32_t* rank;          // [n] rank of the update in its path (atomic order); second arrivals: RANK_PAIR | partner
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
  uint32_t tune;           // experiment switches (env BB_DM_TUNE): 1 walk the tiles in reverse, 2 K1 releases K2 late,
                           // 4 second arrivals prefetch the partner's payload, 8 prefetch only the row's last sector
  unsigned long long* tl;  // diagnostics (BB_TIMELINE=1): [k][2] first start / last end of kernel k in globaltimer ns, or null
  Params p;
  IndexArgs ix;
};

__device__ __forceinline__ void tl_start(const DmArgs& a, int k) {
  if (a.tl && threadIdx.x == 0) atomicMin(a.tl + 2 * k, global_timer_ns());
}
__device__ __forceinline__ void tl_end(const DmArgs& a, int k) {
  if (a.tl && threadIdx.x == 0) atomicMax(a.tl + 2 * k + 1, global_timer_ns());
}

constexpr uint32_t RANK_PAIR = 0x80000000u;
constexpr unsigned long long CW_ONE = 1ull << ROW_CNT_SHIFT;  // one more update of this path

// the row's batch word: low half flags | count << 3, high half sum of arrival indices (bb_kernels.cuh: unpack_row)
__device__ __forceinline__ unsigned long long* row_word(uint4* table, uint64_t pid) {
  return reinterpret_cast<unsigned long long*>(table + pid * ROW_Q + 7);
}
__device__ __forceinline__ uint32_t word_count(unsigned long long wv) { return (uint32_t)wv >> ROW_CNT_SHIFT; }

// ---------------------------------------------------------------- K1
constexpr int DM1_T = 256;
constexpr int DM1_CTAS_PER_SM = 4;

__global__ void __launch_bounds__(DM1_T) k_dm_count(const DmArgs a) {
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x, lane = tid & 31;
  tl_start(a, 0);
  if (!(a.tune & 2u)) pdl_launch_dependents();  // K2 may start staging its payload tiles
  const bool rev = (a.tune & 1u) != 0;
  const uint32_t tiles = (uint32_t)((a.n + DM1_T - 1) / DM1_T), G = gridDim.x;
  const uint32_t lt = lanemask_lt();
  bool bad = false;
  // iteration t of this CTA handles tile tiles - 1 - t: the batch is walked back to front
  auto load_pid = [&](uint32_t t) -> uint64_t {
    if (t >= tiles) return ~0ull;
    const uint64_t i = (uint64_t)(rev ? tiles - 1 - t : t) * DM1_T + tid;
    return i < a.n ? a.path_id[i] : ~0ull;
  };
  uint32_t keep = 0;
  auto pull = [&](uint64_t pid) {
    if (pid >= a.capacity) return;
    if (a.tune & 8u) keep ^= touch_l2(a.table + pid * ROW_Q + 7);
    else prefetch_l2(a.table + pid * ROW_Q);
  };
  uint64_t pid_cur = load_pid(blockIdx.x), pid_nxt = load_pid(blockIdx.x + G);
  pdl_wait();  // the previous batch's last kernel is complete: its rows and counters may be touched
  if (blockIdx.x == 0) {
    if (tid < DC_WORDS) a.ctr_next[tid] = 0;
    if (tid == 0 && a.zero_changes) *a.n_changes = 0;
  }
  pull(pid_cur);
  for (uint32_t t = blockIdx.x; t < tiles; t += G) {
    const uint64_t pid_nn = load_pid(t + 2 * G);
    pull(pid_nxt);  // the whole 128-byte row: K2 finds it in L2
    const uint64_t i = (uint64_t)(rev ? tiles - 1 - t : t) * DM1_T + tid;
    const uint64_t pid = pid_cur;
    const bool ok = i < a.n && pid < a.capacity;
    bad |= i < a.n && !ok;
    const unsigned long long old = ok ? atomicAdd(row_word(a.table, pid), ((unsigned long long)(uint32_t)i << 32) | CW_ONE) : 0ull;
    const uint32_t r = word_count(old);
    const bool second = ok && r == 1u;  // knows the first arrival's index: the sum so far IS that index
    const uint32_t partner = (uint32_t)(old >> 32);
    if (i < a.n) a.rank[i] = second ? (RANK_PAIR | partner) : r;
    if (second && (a.tune & 4u)) {  // K2 will most likely replay both updates from this thread's slot: the partner's payload into L2
      prefetch_l2(a.head + partner);
      prefetch_l2(a.clk + 2 * (uint64_t)partner);
      prefetch_l2(a.val + 2 * (uint64_t)partner);
    }
    const bool third = ok && r == 2u, ninth = ok && r == (uint32_t)DM_SHORT;
    const uint32_t m3 = __ballot_sync(0xffffffffu, third), m9 = __ballot_sync(0xffffffffu, ninth);
    if (m3) {
      uint32_t base = 0;
      if (lane == __ffs(m3) - 1) base = atomicAdd(&a.ctr[DC_NSLAB], (uint32_t)__popc(m3));
      base = __shfl_sync(0xffffffffu, base, __ffs(m3) - 1);
      if (third) {
        const uint32_t s = base + __popc(m3 & lt);
        a.off[pid] = s;
        a.slab_pid[s] = (uint32_t)pid;
      }
    }
    if (m9) {
      uint32_t base = 0;
      if (lane == __ffs(m9) - 1) base = atomicAdd(&a.ctr[DC_NLONG], (uint32_t)__popc(m9));
      base = __shfl_sync(0xffffffffu, base, __ffs(m9) - 1);
      if (ninth) a.long_pid[base + __popc(m9 & lt)] = (uint32_t)pid;
    }
    pid_cur = pid_nxt;
    pid_nxt = pid_nn;
  }
  if (bad) flag_reject(a.ctr + DC_REJ, a.err, a.ordinal);
  if (keep == 0xA5C3F00Fu && a.n == ~0ull) a.rank[0] = keep;  // never true: keeps the touch loads alive
  if (a.tune & 2u) pdl_launch_dependents();

  // the last CTA to get here lays out the runs of the long paths (usually there are none).  The CTA barrier orders
  // every thread's stores before thread 0's fence, and the fence (cumulative) before its ticket
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(&a.ctr[DC_TICKET], 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) {
    tl_end(a, 0);
    return;
  }
  __threadfence();
  const uint32_t nlong = ld_volatile(a.ctr + DC_NLONG);
  uint32_t run = 0;
  for (uint32_t k0 = 0; k0 < nlong; k0 += 256) {
    const uint32_t k = k0 + tid;
    const uint32_t p = k < nlong ? __ldcg(a.long_pid + k) : 0u;
    const uint32_t c = k < nlong ? word_count(__ldcg(row_word(a.table, p))) : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan<256>(c, &total);
    if (k < nlong) a.off[p] = run + ex;  // replaces the slab index the path's third arrival stored
    run += total;
  }
  if (tid == 0) a.ctr[DC_LTOTAL] = run;
  tl_end(a, 0);
}

// ---------------------------------------------------------------- K2
__device__ __forceinline__ void store_row_regs(uint4* row, const RowState& r) {
  uint4 q[ROW_Q];
  pack_row(q, r);
#pragma unroll
  for (int c = 0; c < ROW_Q; ++c) row[c] = q[c];
}

// shared memory of one K2 CTA (36 KB, six CTAs per SM)
struct DmSmem {
  uint4 head[DM_T];          // the tile's payloads, staged by cp.async.bulk; an accepted update's slot is
  uint4 clk[2 * DM_T];       // overwritten with its change entry
  uint4 val[2 * DM_T];
  uint4 pp[DM_T * UPD_Q];    // partner payload of a 2-update path (slot of its owner), then its change entry
  uint4 row[DM_T * ROW_Q];   // every update's table row, 128-byte stride, chunk index XOR-swizzled
  uint64_t bar;
  unsigned long long base;
  uint32_t wsum[DM_WARPS];
};

template <bool INDEXED>
__global__ void __launch_bounds__(DM_T, 6) k_dm_merge(const DmArgs a) {
  BB_DYN_SMEM(dm_raw);
  DmSmem& sm = *reinterpret_cast<DmSmem*>(dm_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, wbase = w * 32;
  const uint64_t base = (uint64_t)blockIdx.x * DM_T;
  const uint64_t i = base + tid;
  const bool valid = i < a.n;
  const uint32_t nvalid = (uint32_t)min((uint64_t)DM_T, a.n - base);
  tl_start(a, 1);

  // ---- stage: the tile's payloads, three bulk copies; nothing here depends on K1
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&sm.bar, nvalid * 80u);
    bulk_g2s(sm.head, a.head + base, nvalid * 16u, &sm.bar);
    bulk_g2s(sm.clk, a.clk + 2 * base, nvalid * 32u, &sm.bar);
    bulk_g2s(sm.val, a.val + 2 * base, nvalid * 32u, &sm.bar);
  }
  const uint64_t pid = valid ? a.path_id[i] : ~0ull;
  pdl_launch_dependents();
  __syncthreads();  // the barrier word is initialised for everybody
  pdl_wait();       // K1 is complete: counts (in the rows), ranks, slabs, runs

  // ---- everything K1 left for this update, fetched at once: its row (8 lanes per 128-byte row), its rank word,
  // the batch's reject flag; a second arrival also fetches its partner's payload, before it knows whether the
  // path stayed at two updates
  const uint32_t rej = *a.rej;
  const bool inrange = pid < a.capacity;  // a batch with an id out of range is rejected: no row is fetched through it
  const uint32_t rk = valid ? a.rank[i] : 0u;
  const uint32_t vmask = __ballot_sync(0xffffffffu, inrange);
  const uint32_t key = (uint32_t)pid;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = j * 4 + (lane >> 3), chunk = lane & 7;
    const uint32_t ekey = __shfl_sync(0xffffffffu, key, e);
    if ((vmask >> e) & 1u) cp_async16(&sm.row[row_slot(wbase + e, chunk)], a.table + (uint64_t)ekey * ROW_Q + chunk);
  }
  const bool second = inrange && (rk & RANK_PAIR) != 0u;
  const uint32_t partner = rk & ~RANK_PAIR;
  if (second) {
    uint4* pp = &sm.pp[tid * UPD_Q];
    cp_async16(pp, a.head + partner);
    cp_async16(pp + 1, a.clk + 2 * (uint64_t)partner);
    cp_async16(pp + 2, a.clk + 2 * (uint64_t)partner + 1);
    cp_async16(pp + 3, a.val + 2 * (uint64_t)partner);
    cp_async16(pp + 4, a.val + 2 * (uint64_t)partner + 1);
  }
  uint64_t prim[F], prim0[F];
  if (INDEXED && inrange) {
#pragma unroll
    for (int f = 0; f < F; ++f) prim0[f] = prim[f] = ((a.ix.mask >> f) & 1u) ? a.ix.pcol[f][pid] : BB_KEY_NONE;
  }
  cp_async_wait_all();
  __syncwarp();  // a row's chunks were fetched by eight lanes of this warp
  mbar_wait(&sm.bar, 0);
  if (rej & 1u) {  // rejected batch: the table stays as it is; K1's additions to the rows' batch words are undone
    if (inrange) atomicAdd(row_word(a.table, pid), 0ull - (((unsigned long long)(uint32_t)i << 32) | CW_ONE));
    return;
  }

  uint32_t cnt = 0;
  if (inrange) cnt = sm.row[row_slot(tid, 7)].x >> ROW_CNT_SHIFT;
  const bool single = cnt == 1u;
  const bool owner = cnt == 2u && second;
  const bool replay = single || owner;  // this thread replays the path
  uint32_t slot_off = 0;
  if (cnt > 2u) slot_off = a.off[pid];  // consumed after the resolve: the load hides behind it

  // ---- resolve: pass 0 = the earlier update of a 2-update path (owners only), pass 1 = the later / only one
  uint32_t code0 = 0xFFu, code1 = 0xFFu;
  const bool own_first = owner && (uint32_t)i < partner;  // the second arrival (atomic order) may be the earlier update
  if (replay) {
    RowState r;
    {
      uint4 q[ROW_Q];
#pragma unroll
      for (int c = 0; c < ROW_Q; ++c) q[c] = sm.row[row_slot(tid, c)];
      unpack_row(q, r);
    }
#pragma unroll 1
    for (int pass = owner ? 0 : 1; pass < 2; ++pass) {
      const bool from_pp = owner && ((pass == 0) != own_first);  // this pass replays the partner's update
      uint4 h, c0, c1, v0, v1;
      if (from_pp) {
        const uint4* pp = &sm.pp[tid * UPD_Q];
        h = pp[0]; c0 = pp[1]; c1 = pp[2]; v0 = pp[3]; v1 = pp[4];
      } else {
        h = sm.head[tid]; c0 = sm.clk[2 * tid]; c1 = sm.clk[2 * tid + 1]; v0 = sm.val[2 * tid]; v1 = sm.val[2 * tid + 1];
      }
      Clock c, oc;
      Value x, ov;
      const bool net = unpack_update(h, c0, c1, v0, v1, c, x);
      const uint32_t ui = from_pp ? partner : (uint32_t)i;
      const uint32_t code = resolve_step(a.p, r, net, c, x, a.seq_base + ui, ov, oc);
      if (INDEXED) index_hook(a.ix, key, r.s, x, prim, r.xcnt, a.err);
      if (BB_DEC_ACCEPTED(code)) {
        uint4 q[UPD_Q];
        pack_change(q, h.w, ov, oc);
        if (from_pp) {
          uint4* pp = &sm.pp[tid * UPD_Q];
#pragma unroll
          for (int k = 0; k < UPD_Q; ++k) pp[k] = q[k];
        } else {
          sm.head[tid] = q[0]; sm.clk[2 * tid] = q[1]; sm.clk[2 * tid + 1] = q[2]; sm.val[2 * tid] = q[3]; sm.val[2 * tid + 1] = q[4];
        }
      }
      if (from_pp) code0 = code;  // code0 / code1: the partner's / the thread's own update
      else code1 = code;
    }
    {
      uint4 q[ROW_Q];
      pack_row(q, r);  // count and index sum are written back as zero: the path's batch word is clean again
#pragma unroll
      for (int c = 0; c < ROW_Q; ++c) sm.row[row_slot(tid, c)] = q[c];
    }
    if (INDEXED) {
#pragma unroll
      for (int f = 0; f < F; ++f)
        if (prim[f] != prim0[f]) a.ix.pcol[f][pid] = prim[f];
    }
  } else if (cnt > 2u) {  // 3..8: the path's slab; more: its run.  K3 replays it
    const uint32_t r3 = (rk & RANK_PAIR) ? 1u : rk;
    if (cnt <= (uint32_t)DM_SHORT) a.slab[(uint64_t)slot_off * DM_SHORT + r3] = (uint32_t)i;
    else a.litems[slot_off + r3] = (uint32_t)i;
  }
  __syncwarp();
  // rows back, eight lanes per row
  const uint32_t rmask = __ballot_sync(0xffffffffu, replay);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = j * 4 + (lane >> 3), chunk = lane & 7;
    const uint32_t ekey = __shfl_sync(0xffffffffu, key, e);
    if ((rmask >> e) & 1u) a.table[(uint64_t)ekey * ROW_Q + chunk] = sm.row[row_slot(wbase + e, chunk)];
  }

  // ---- rank the tile's accepted entries (a path's two in arrival order), one claim per CTA
  const uint32_t acc0 = (code0 != 0xFFu && BB_DEC_ACCEPTED(code0)) ? 1u : 0u;
  const uint32_t acc1 = (code1 != 0xFFu && BB_DEC_ACCEPTED(code1)) ? 1u : 0u;
  const uint32_t mine = acc0 + acc1;
  const uint32_t inc = warp_inclusive_scan(mine);
  if (lane == 31) sm.wsum[w] = inc;
  __syncthreads();
  uint32_t before = inc - mine, total = 0;
#pragma unroll
  for (int ww = 0; ww < DM_WARPS; ++ww) {
    const uint32_t c = sm.wsum[ww];
    if (ww < w) before += c;
    total += c;
  }
  if (tid == 0) sm.base = total ? atomicAdd(a.n_changes, (unsigned long long)total) : 0ull;
  __syncthreads();
  // slots: the earlier update's entry first
  const uint64_t dfirst = sm.base + before;
  const uint64_t dpp = own_first ? dfirst + acc1 : dfirst, down = own_first ? dfirst : dfirst + acc0;
  bool overflow = false;
  if (replay) {
    a.verdict[i] = (code1 << 29) | (acc1 ? (uint32_t)down : NO_SLOT);
    if (owner) a.verdict[partner] = (code0 << 29) | (acc0 ? (uint32_t)dpp : NO_SLOT);
    if (acc0) {
      if (dpp < a.cap) {
        const uint4* pp = &sm.pp[tid * UPD_Q];
        a.out_idx[dpp] = a.idx_base + partner;
        a.out_head[dpp] = pp[0];
        a.out_clk[2 * dpp] = pp[1];
        a.out_clk[2 * dpp + 1] = pp[2];
        a.out_val[2 * dpp] = pp[3];
        a.out_val[2 * dpp + 1] = pp[4];
      } else {
        overflow = true;
      }
    }
    if (acc1) {
      if (down < a.cap) {
        a.out_idx[down] = a.idx_base + (uint32_t)i;
        a.out_head[down] = sm.head[tid];
        a.out_clk[2 * down] = sm.clk[2 * tid];
        a.out_clk[2 * down + 1] = sm.clk[2 * tid + 1];
        a.out_val[2 * down] = sm.val[2 * tid];
        a.out_val[2 * down + 1] = sm.val[2 * tid + 1];
      } else {
        overflow = true;
      }
    }
  }
  if (overflow) atomicOr(a.err, ERR_CHANGES);
  tl_end(a, 1);
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
  tl_start(a, 2);
  pdl_launch_dependents();  // the next batch's K1 may load its path ids
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
    uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
    if (active) {  // unclaimed slots hold garbage: masked by the count below
      const uint4* sl = reinterpret_cast<const uint4*>(a.slab + (uint64_t)s * DM_SHORT);
      lo = sl[0];
      hi = sl[1];
    }
    RowState r;
    uint32_t cnt = 0;
    if (active) {
      uint4 q[ROW_Q];
#pragma unroll
      for (int c = 0; c < ROW_Q; ++c) q[c] = __ldcg(a.table + (uint64_t)pid * ROW_Q + c);
      cnt = q[7].x >> ROW_CNT_SHIFT;
      unpack_row(q, r);
    }
    if (cnt > (uint32_t)DM_SHORT) active = false;  // became a long path: phase B
    if (!active) cnt = 0;
    uint64_t prim[F], prim0[F];
    if (active) {
      uint32_t v[DM_SHORT] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int k = 0; k < DM_SHORT; ++k)
        if (k >= (int)cnt) v[k] = 0xFFFFFFFFu;
      sort8(v);
#pragma unroll
      for (int k = 0; k < DM_SHORT; ++k) s_idx[k][tid] = v[k];
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
      store_row_regs(a.table + (uint64_t)pid * ROW_Q, r);  // batch word cleared
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
    const uint32_t len = word_count(__ldcg(row_word(a.table, pid))), start = a.off[pid];
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
    if (tid < ROW_Q) a.table[(uint64_t)pid * ROW_Q + tid] = s_row[tid];  // packed by a retiring thread: batch word cleared
    if (INDEXED && tid < F && ((a.ix.mask >> tid) & 1u)) a.ix.pcol[tid][pid] = s_prim[tid];
    __syncthreads();
  }
  if (overflow) atomicOr(a.err, ERR_CHANGES);
  tl_end(a, 2);
}

}  // namespace bb
