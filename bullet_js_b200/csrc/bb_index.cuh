// bb_index.cuh - BulletQuery on the device: index build, the post-write hook, and the
// equals / range / count scans.
//
// Follows (paths relative to the reference repo):
//   index / _buildIndex / _addToIndex / _getIndexableValue   src/bullet-query.js:30-94, 126-131
//   _updateIndices (post-write hook) / _removeFromIndex       src/bullet-query.js:139-176, 103-118
//   equals / count / range                                    src/bullet-query.js:186-210, 293-313, 221-261
//
// An index is the set of (node, key) pairs the reference's Map<String(value), Set<path>>
// contains (include/bullet_b200.h).  Layout in HBM, per indexed field:
//   pcol[capacity]   u64   one entry per node, BB_KEY_NONE when the node has none: the dense
//                          column the scans stream (8 B per row, hits come out in node order)
//   xkey/xnode[2^k]  u64/u32  open-addressing overflow set (linear probing, tombstones) for the
//                          entries beyond a node's first - the reference never removes the
//                          pre-update value's entry (query:151-167), so they pile up
// A node's entries are only ever touched by the thread that replays that node's updates, so
// the only race is two nodes claiming one free slot: one CAS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bb_merge.cuh"

namespace bb {

constexpr uint32_t X_EMPTY = 0xFFFFFFFFu, X_TOMB = 0xFFFFFFFEu;

struct IndexArgs {
  uint32_t mask;       // bit f: the index on field f is live
  uint64_t* pcol[F];
  uint64_t* xkey[F];
  uint32_t* xnode[F];
  uint32_t xmask[F];   // slots - 1
  uint32_t* xused;     // [F] slots that have left the EMPTY state
  // BB_CFG_EXACT_ORDER: every entry carries the tag of the add that inserted it, and the hook logs its
  // EFFECTIVE adds / removes as events (key, field << 62 | tag) for the per-bucket replay that follows the batch
  uint64_t* pseq[F];   // [capacity] tag of the node's dense entry
  uint64_t* xseq[F];   // [slots]    tag of an overflow entry
  uint64_t* ev_key;    // [ev_cap]
  uint64_t* ev_tag;
  uint32_t* ev_count;
  uint32_t ev_cap;
};

// Event / entry tags: 4 * (sequence of the update) + 2 for its remove, + 3 for its add (the hook removes, then adds:
// query:153-167); an entry put there by an index BUILD carries 4 * (creation sequence of its node): builds walk the
// store in creation order (query:58-66) and everything a build inserts precedes everything a later update inserts.
// (Bit 61 keeps an update's tags above every build tag whatever the two counters are: a table loaded with
// bb_table_load brings its own creation sequences, and the update counter starts at zero.)
constexpr uint64_t TAG_UPDATE = 1ull << 61;
__device__ __forceinline__ uint64_t tag_remove(uint64_t seq) { return TAG_UPDATE | (4 * seq + 2); }
__device__ __forceinline__ uint64_t tag_add(uint64_t seq) { return TAG_UPDATE | (4 * seq + 3); }
__device__ __forceinline__ bool tag_is_add(uint64_t tag) { return (tag & 3u) != 2u; }
constexpr uint64_t TAG_MASK = (1ull << 62) - 1;

__device__ __forceinline__ void log_event(const IndexArgs& ix, int f, uint64_t key, uint64_t tag, uint32_t* err) {
  const uint32_t i = atomicAdd(ix.ev_count, 1u);
  if (i < ix.ev_cap) {
    ix.ev_key[i] = key;
    ix.ev_tag[i] = ((uint64_t)f << 62) | tag;
  } else {
    atomicOr(err, ERR_XFULL);
  }
}

// String(value) as a 64-bit key (query:126-131); value is a non-null primitive
__device__ __forceinline__ uint64_t canon_key(uint32_t tag, uint64_t pay) {
  if (tag == BB_TAG_NUM) {
    if ((pay << 1) == 0) return 0;                                    // String(-0) == "0"
    if ((pay << 1) > 0xFFE0000000000000ull) return BB_KEY_NAN;        // "NaN"
    return pay;
  }
  return (tag == BB_TAG_STR ? BB_KEY_STR : BB_KEY_BOOL) | pay;
}

__device__ __forceinline__ uint32_t x_hash(uint32_t node, uint64_t k) {
  uint64_t h = (k ^ (k >> 31)) * 0x9E3779B97F4A7C15ull + (uint64_t)node * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (uint32_t)h;
}

// slot of (node, k) or -1
__device__ __forceinline__ int64_t x_find(const IndexArgs& ix, int f, uint32_t node, uint64_t k) {
  const uint32_t m = ix.xmask[f];
  uint32_t i = x_hash(node, k) & m;
  for (uint32_t probe = 0; probe <= m; ++probe, i = (i + 1) & m) {
    const uint32_t n = __ldcg(ix.xnode[f] + i);
    if (n == X_EMPTY) return -1;
    if (n == node && __ldcg(ix.xkey[f] + i) == k) return (int64_t)i;
  }
  return -1;
}

__device__ __forceinline__ void x_remove_at(const IndexArgs& ix, int f, int64_t slot) {
  ix.xkey[f][slot] = BB_KEY_NONE;  // scans read keys only
  ix.xnode[f][slot] = X_TOMB;
}

// (node, k) is known to be absent; the slot it went into, or -1 when the set is full
__device__ __forceinline__ int64_t x_insert(const IndexArgs& ix, int f, uint32_t node, uint64_t k) {
  const uint32_t m = ix.xmask[f];
  uint32_t i = x_hash(node, k) & m;
  for (uint32_t probe = 0; probe <= m; ++probe, i = (i + 1) & m) {
    const uint32_t n = __ldcg(ix.xnode[f] + i);
    if (n != X_EMPTY && n != X_TOMB) continue;
    if (n == X_EMPTY && atomicAdd(ix.xused + f, 1u) >= m - (m >> 3)) {  // keep 1/8 of the slots EMPTY
      atomicSub(ix.xused + f, 1u);
      return -1;
    }
    if (atomicCAS(ix.xnode[f] + i, n, node) == n) {
      ix.xkey[f][i] = k;
      return (int64_t)i;
    }
    if (n == X_EMPTY) atomicSub(ix.xused + f, 1u);  // another node took the slot
  }
  return -1;
}

__device__ __forceinline__ uint32_t xcnt_get(uint32_t xcnt, int f) { return (xcnt >> (8 * f)) & 0xFFu; }

// _updateIndices (query:139-176) for one setData on `node`: `after` is the node as _getData
// sees it after the write, `x` the raw incoming value.  prim[f] caches pcol[f][node].
template <bool EXACT = false>
__device__ __forceinline__ void index_hook(const IndexArgs& ix, uint32_t node, const Value& after, const Value& x,
                                           uint64_t (&prim)[F], uint32_t& xcnt, uint32_t* err, uint64_t seq = 0) {
#pragma unroll
  for (int f = 0; f < F; ++f) {
    if (!((ix.mask >> f) & 1u)) continue;
    // The common case - the update won outright, so the node now reads as the update itself - removes key k and adds
    // key k again: k is in the set afterwards either way.  Without exact order (where delete + add moves the entry to
    // the end of its bucket) that is "make sure k is there": one lookup instead of two plus a tombstone and an insert.
    bool same = false;
    if (!EXACT && kind_of(after.meta) == BB_KIND_OBJ && kind_of(x.meta) == BB_KIND_OBJ) {
      const uint32_t ta = tag_of(after.meta, f), tx = tag_of(x.meta, f);
      same = ta != BB_TAG_ABSENT && tx != BB_TAG_ABSENT && !prim_falsy(ta, after.val[f]) && !prim_falsy(tx, x.val[f]) &&
             canon_key(ta, after.val[f]) == canon_key(tx, x.val[f]);
    }
    if (!same && kind_of(after.meta) == BB_KIND_OBJ) {  // if (oldData && oldData[field]) remove
      const uint32_t t = tag_of(after.meta, f);
      if (t != BB_TAG_ABSENT && !prim_falsy(t, after.val[f])) {
        const uint64_t k = canon_key(t, after.val[f]);
        if (prim[f] == k) {
          prim[f] = BB_KEY_NONE;
          if (EXACT) log_event(ix, f, k, tag_remove(seq), err);
        } else if (xcnt_get(xcnt, f)) {
          const int64_t slot = x_find(ix, f, node, k);
          if (slot >= 0) {
            x_remove_at(ix, f, slot);
            if (xcnt_get(xcnt, f) != 0xFFu) xcnt -= 1u << (8 * f);
            if (EXACT) log_event(ix, f, k, tag_remove(seq), err);
          }
        }
      }
    }
    if (kind_of(x.meta) == BB_KIND_OBJ) {  // if (newData && newData[field]) add
      const uint32_t t = tag_of(x.meta, f);
      if (t != BB_TAG_ABSENT && !prim_falsy(t, x.val[f])) {
        const uint64_t k = canon_key(t, x.val[f]);
        if (prim[f] == k) continue;
        if (xcnt_get(xcnt, f) && x_find(ix, f, node, k) >= 0) continue;
        if (prim[f] == BB_KEY_NONE) {
          prim[f] = k;
          if (EXACT) {
            ix.pseq[f][node] = tag_add(seq);
            log_event(ix, f, k, tag_add(seq), err);
          }
        } else {
          const int64_t slot = x_insert(ix, f, node, k);
          if (slot >= 0) {
            if (xcnt_get(xcnt, f) != 0xFFu) xcnt += 1u << (8 * f);
            if (EXACT) {
              ix.xseq[f][slot] = tag_add(seq);
              log_event(ix, f, k, tag_add(seq), err);
            }
          } else {
            atomicOr(err, ERR_XFULL);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- K4: index build
// One thread per row, every requested field in the same pass: `field in value` and value[field] not null -> its key
// (query:58-66, 82-85).  Reads the first 48 bytes of each row (values + header word: one 64-byte DRAM access, see
// row_chunk in bb_kernels.cuh) and writes 8 bytes per row and field; the padding element of an odd-sized column
// gets BB_KEY_NONE here, so the column needs no clearing pass.
struct BuildArgs {
  const uint4* table;
  uint64_t capacity, padded;  // rows; elements per column (capacity rounded up to even)
  uint32_t mask;              // fields to build
  uint64_t* pcol[F];
};

// A warp covers 32 rows: four lanes fetch the first 64 bytes of a row (one full DRAM access, 8 rows per load
// instruction instead of 32 scattered 16-byte pieces), the lane that holds a value slot computes its key, and
// the keys are transposed with shuffles so that every column is written as one coalesced 256-byte run.
__global__ void __launch_bounds__(256) k_index_build(const BuildArgs a) {
  const int lane = threadIdx.x & 31, c = lane & 3;
  const uint64_t row0 = ((uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  if (row0 >= a.padded) return;
  uint4 v[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const uint64_t r = row0 + it * 8 + (lane >> 2);
    v[it] = r < a.capacity ? a.table[r * 8 + c] : make_uint4(0, 0, 0, 0);  // device chunks 0, 1: values; 2: orders + header
  }
  uint64_t key[4][2];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const uint32_t meta = __shfl_sync(0xffffffffu, v[it].z, (lane & ~3) | 2);
    const bool obj = kind_of(meta) == BB_KIND_OBJ;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int f = 2 * c + h;  // the field whose value this lane holds (lanes with c >= 2 hold none)
      const uint32_t t = f < F ? (meta >> (BB_HDR_TAG_SHIFT + 3 * f)) & 7u : BB_TAG_ABSENT;
      const uint64_t val = h ? ((uint64_t)v[it].z | ((uint64_t)v[it].w << 32)) : ((uint64_t)v[it].x | ((uint64_t)v[it].y << 32));
      key[it][h] = (obj && t != BB_TAG_ABSENT && t != BB_TAG_NULL) ? canon_key(t, val) : BB_KEY_NONE;
    }
  }
  const uint64_t row = row0 + lane;
#pragma unroll
  for (int f = 0; f < F; ++f) {
    if (!((a.mask >> f) & 1u)) continue;  // (warp-uniform)
    uint64_t out = BB_KEY_NONE;
#pragma unroll
    for (int it = 0; it < 4; ++it) {  // row `lane` of this warp was fetched in round lane / 8 by lanes 4 * (lane % 8) ..
      const uint64_t k = __shfl_sync(0xffffffffu, key[it][f & 1], (lane & 7) * 4 + (f >> 1));
      if ((lane >> 3) == it) out = k;
    }
    if (row < a.padded) a.pcol[f][row] = out;
  }
}

__global__ void __launch_bounds__(256) k_fill_u64(uint64_t* __restrict__ p, uint64_t n, uint64_t v) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// count live entries of a key column
__global__ void __launch_bounds__(256) k_count_live(const uint64_t* __restrict__ keys, uint64_t n,
                                                    unsigned long long* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = 0;
  for (; i < n; i += (uint64_t)gridDim.x * blockDim.x) c += keys[i] != BB_KEY_NONE;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// ---------------------------------------------------------------- K5: equals / range / count scan
struct Pred {
  uint32_t mode;  // 0 equals, 1 range with numeric bounds, 2 range with a string bound
  uint32_t lo_flags, hi_flags;
  uint64_t eq;           // mode 0: the key; mode 1: ordered image of the lower bound
  uint64_t width;        // mode 1: ordered image of the upper bound minus that of the lower
  double lo, hi;
  uint64_t lo_rank, hi_rank;
};

// order-preserving image of f64 bits in u64: numbers keep their order, every NaN pattern (that is
// every non-numeric key, BB_KEY_NAN and BB_KEY_NONE) lands outside [image(-inf), image(+inf)]
__host__ __device__ __forceinline__ uint64_t ordered_image(uint64_t bits) {
  return bits ^ ((uint64_t)((int64_t)bits >> 63) | 0x8000000000000000ull);
}

// equals: same String(value) (query:200-203).  range (query:238-252): v = Number(key), or the key
// itself when that is NaN; v >= min && v <= max with JS relational semantics.
template <int MODE>
__device__ __forceinline__ bool pred_match(const Pred& p, uint64_t k) {
  if (MODE == 0) return k == p.eq;                             // never BB_KEY_NONE
  if (MODE == 1) return ordered_image(k) - p.eq <= p.width;    // one subtract, one compare
  if (k == BB_KEY_NONE) return false;
  const uint32_t top = (uint32_t)(k >> 48);
  if (top == (uint32_t)(BB_KEY_STR >> 48)) {
    const uint64_t id = k & 0xFFFFFFFFFFFFull;
    return (p.lo_flags & p.hi_flags & BB_BOUND_IS_STRING) && id >= p.lo_rank && id < p.hi_rank;
  }
  if (top == (uint32_t)(BB_KEY_BOOL >> 48)) return (p.lo_flags & p.hi_flags & ((k & 1u) ? BB_BOUND_TRUE : BB_BOUND_FALSE)) != 0;
  if (k == BB_KEY_NAN) return (p.lo_flags & p.hi_flags & BB_BOUND_NAN) != 0;
  const double x = __longlong_as_double((long long)k);
  return x >= p.lo && x <= p.hi;
}

constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_ROUNDS = 8;                          // 16-byte loads in flight per thread
constexpr int SC_TILE = SC_THREADS * SC_ROUNDS * 2;   // 4096 keys (32 KB) per CTA

struct ScanArgs {
  const uint64_t* keys;    // [n] (n even, 16-byte aligned)
  const uint32_t* nodes;   // node of entry i, or nullptr: entry i belongs to node i
  uint64_t n;
  uint32_t* out;           // hit node ids, or nullptr: count only
  uint64_t cap;
  unsigned long long* counters;  // [2]: dense matches, overflow matches
  uint32_t which;          // 0: dense column (writes from 0), 1: overflow set (writes after counters[0])
  uint32_t ref_or;         // OR-ed into the entry index when `nodes` is null (exact order: overflow hits leave as slot | 2^31)
  uint32_t* tile_state;    // [num_tiles] zeroed
  uint32_t* ticket;        // zeroed
  uint32_t num_tiles;
  uint32_t* err;
  Pred p;
};

// Streams the key column with coalesced 16-byte loads (8 in flight per thread), evaluates the
// predicate, ranks the hits in entry order (ballots inside a warp, a 64-entry scan across the
// warps and rounds of the tile, a decoupled look-back across tiles) and writes their node ids
// as one dense run per tile.  Count-only calls skip the ranking and the chain.
// ORDERED (BB_CFG_ORDERED_CHANGES): hits are ranked in entry order (ballots inside a warp, a 64-entry scan
// across the warps and rounds of the tile) and tiles are chained with a decoupled look-back, so the hits of
// the whole column come out in ascending entry order; otherwise a thread's hits form one run, a tile claims
// its slice with one atomicAdd when it is done and never waits (a multiset, in no particular order).
template <bool ORDERED, int MODE>
__global__ void __launch_bounds__(SC_THREADS) k_index_scan(const ScanArgs a) {
  __shared__ uint32_t s_cnt[SC_ROUNDS * SC_WARPS];
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (ORDERED) {
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
  }
  const uint32_t tile = ORDERED ? s_tile : blockIdx.x;
  const uint64_t base = (uint64_t)tile * SC_TILE;
  const uint4* k4 = reinterpret_cast<const uint4*>(a.keys);
  uint4 v[SC_ROUNDS];
#pragma unroll
  for (int j = 0; j < SC_ROUNDS; ++j) {
    const uint64_t e = base + 2 * ((uint64_t)j * SC_THREADS + tid);
    v[j] = e < a.n ? __ldcs(k4 + e / 2) : make_uint4(~0u, ~0u, ~0u, ~0u);
  }
  if (!ORDERED) {
    // Unordered hits (the default): a thread's hits go out as one run, so ranking is ONE warp scan of the
    // per-thread counts instead of two ballots and four popcounts per round - the kernel was issue-bound
    // (78 % issue slots busy at 68 % of the HBM peak).  Order inside a tile is (thread, round), not entry order.
    uint32_t flags = 0;  // 2 hit bits per round
#pragma unroll
    for (int j = 0; j < SC_ROUNDS; ++j) {
      const bool h0 = pred_match<MODE>(a.p, (uint64_t)v[j].x | ((uint64_t)v[j].y << 32));
      const bool h1 = pred_match<MODE>(a.p, (uint64_t)v[j].z | ((uint64_t)v[j].w << 32));
      flags |= ((uint32_t)h0 | ((uint32_t)h1 << 1)) << (2 * j);
    }
    const uint32_t cnt = __popc(flags);
    const uint32_t inc = warp_inclusive_scan(cnt);
    if (lane == 31) s_cnt[w] = inc;
    __syncthreads();
    if (w == 0) {
      const uint32_t c = lane < SC_WARPS ? s_cnt[lane] : 0u;
      const uint32_t incw = warp_inclusive_scan(c);
      const uint32_t tile_total = __shfl_sync(0xffffffffu, incw, SC_WARPS - 1);
      if (lane < SC_WARPS) s_cnt[lane] = incw - c;
      if (lane == 0) {
        if (a.out == nullptr) {
          if (tile_total) atomicAdd(a.counters + a.which, (unsigned long long)tile_total);
        } else {
          s_base = tile_total ? atomicAdd(a.counters + a.which, (unsigned long long)tile_total) : 0ull;
        }
      }
    }
    __syncthreads();
    if (a.out == nullptr || cnt == 0) return;
    uint64_t d = (a.which ? a.counters[0] : 0ull) + s_base + s_cnt[w] + (inc - cnt);
    bool overflow = false;
    while (flags) {
      const int b = __ffs(flags) - 1;
      flags &= flags - 1;
      const uint64_t e = base + 2 * ((uint64_t)(b >> 1) * SC_THREADS + tid) + (b & 1);
      if (d < a.cap) a.out[d] = a.nodes ? a.nodes[e] : ((uint32_t)e | a.ref_or);
      else overflow = true;
      ++d;
    }
    if (overflow) atomicOr(a.err, ERR_HITS);
    return;
  }
  uint32_t flags = 0;                // 2 hit bits per round
  uint32_t lpre_lo = 0, lpre_hi = 0;  // hits of lower lanes in the same round, 8 bits per round
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < SC_ROUNDS; ++j) {
    const bool h0 = pred_match<MODE>(a.p, (uint64_t)v[j].x | ((uint64_t)v[j].y << 32));
    const bool h1 = pred_match<MODE>(a.p, (uint64_t)v[j].z | ((uint64_t)v[j].w << 32));
    const uint32_t b0 = __ballot_sync(0xffffffffu, h0), b1 = __ballot_sync(0xffffffffu, h1);
    flags |= ((uint32_t)h0 | ((uint32_t)h1 << 1)) << (2 * j);
    const uint32_t below = __popc(b0 & lt) + __popc(b1 & lt);
    if (j < 4) lpre_lo |= below << (8 * j);
    else lpre_hi |= below << (8 * (j - 4));
    if (lane == 0) s_cnt[j * SC_WARPS + w] = __popc(b0) + __popc(b1);
  }
  __syncthreads();
  if (w == 0) {  // exclusive scan of the 64 (round, warp) counts, then the tile's place in the output
    const uint32_t c0 = s_cnt[2 * lane], c1 = s_cnt[2 * lane + 1];
    const uint32_t inc = warp_inclusive_scan(c0 + c1);
    s_cnt[2 * lane] = inc - c0 - c1;
    s_cnt[2 * lane + 1] = inc - c1;
    const uint32_t tile_total = __shfl_sync(0xffffffffu, inc, 31);
    if (a.out == nullptr) {
      if (lane == 0 && tile_total) atomicAdd(a.counters + a.which, (unsigned long long)tile_total);
    } else if (ORDERED) {
      const uint32_t ex = tile_prefix(a.tile_state, tile, tile_total);
      if (lane == 0) {
        s_base = ex;
        if (tile == a.num_tiles - 1) a.counters[a.which] = (unsigned long long)ex + tile_total;
      }
    } else if (lane == 0) {
      s_base = tile_total ? atomicAdd(a.counters + a.which, (unsigned long long)tile_total) : 0ull;
    }
  }
  __syncthreads();
  if (a.out == nullptr || flags == 0) return;
  const uint64_t obase = (a.which ? a.counters[0] : 0ull) + s_base;
  bool overflow = false;
#pragma unroll
  for (int j = 0; j < SC_ROUNDS; ++j) {
    const uint32_t hb = (flags >> (2 * j)) & 3u;
    if (!hb) continue;
    const uint32_t lp = ((j < 4 ? lpre_lo >> (8 * j) : lpre_hi >> (8 * (j - 4))) & 0xFFu);
    uint64_t d = obase + s_cnt[j * SC_WARPS + w] + lp;
    const uint64_t e = base + 2 * ((uint64_t)j * SC_THREADS + tid);
    if (hb & 1u) {
      if (d < a.cap) a.out[d] = a.nodes ? a.nodes[e] : ((uint32_t)e | a.ref_or);
      else overflow = true;
      ++d;
    }
    if (hb & 2u) {
      if (d < a.cap) a.out[d] = a.nodes ? a.nodes[e + 1] : ((uint32_t)(e + 1) | a.ref_or);
      else overflow = true;
    }
  }
  if (overflow) atomicOr(a.err, ERR_HITS);
}

// ---------------------------------------------------------------- exact Map / Set order (BB_CFG_EXACT_ORDER)
// The reference returns equals / range results in (bucket creation order, insertion order inside the bucket)
// (query:89-93, 110-116, 204, 237-258).  Entries carry the tag of the add that inserted them (above); a bucket's
// creation tag is that of the add that followed the last time its entry count was zero.  After every batch (and after
// an index build) the logged events are sorted by (key, field, tag) and every bucket's run is replayed against a
// per-field bucket table key -> (count, created); a query sorts its hits by (created, entry tag).
struct BucketTable {
  uint64_t* key;      // [slots] BB_KEY_NONE = empty
  uint32_t* count;    // live entries
  uint64_t* created;  // tag of the add that (re-)created the bucket
  uint32_t mask;      // slots - 1
};

__device__ __forceinline__ uint32_t bucket_hash(uint64_t k) {
  uint64_t h = (k ^ (k >> 33)) * 0xFF51AFD7ED558CCDull;
  h ^= h >> 29;
  return (uint32_t)(h * 0x9E3779B97F4A7C15ull >> 32);
}

// slot of `k`, inserting it if absent; -1 when the table is full
__device__ __forceinline__ int64_t bucket_find_or_insert(const BucketTable& t, uint64_t k) {
  uint32_t i = bucket_hash(k) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe, i = (i + 1) & t.mask) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(t.key + i);
    if (cur == BB_KEY_NONE) cur = atomicCAS(reinterpret_cast<unsigned long long*>(t.key + i), (unsigned long long)BB_KEY_NONE, (unsigned long long)k);
    if (cur == BB_KEY_NONE || cur == k) return (int64_t)i;
  }
  return -1;
}

__device__ __forceinline__ int64_t bucket_find(const BucketTable& t, uint64_t k) {
  uint32_t i = bucket_hash(k) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe, i = (i + 1) & t.mask) {
    const uint64_t cur = t.key[i];
    if (cur == k) return (int64_t)i;
    if (cur == BB_KEY_NONE) return -1;
  }
  return -1;
}

// the entries an index BUILD inserted: tag = 4 * creation sequence of the node, one add event each
__global__ void __launch_bounds__(256) k_exact_build_events(const uint4* __restrict__ table, uint64_t capacity, int f, IndexArgs ix,
                                                            uint32_t* __restrict__ err) {
  const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= capacity) return;
  const uint64_t k = ix.pcol[f][row];
  if (k == BB_KEY_NONE) return;
  const uint4 q = table[row * 8 + 3];  // device chunk 3 = public chunk 7: flags, xcnt, cseq
  const uint64_t cseq = (uint64_t)q.z | ((uint64_t)q.w << 32);
  ix.pseq[f][row] = 4 * cseq;
  log_event(ix, f, k, 4 * cseq, err);
}

// pads [n, padded) of a sort buffer with keys that sort last; n is read on the device (no host round trip)
__global__ void __launch_bounds__(256) k_sort_pad(uint64_t* __restrict__ a, uint64_t* __restrict__ b, const uint32_t* __restrict__ n_dev,
                                                  uint32_t n_host, uint32_t padded) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = n_dev ? min(*n_dev, padded) : n_host;
  if (i >= n && i < padded) {
    a[i] = ~0ull;
    b[i] = ~0ull;
  }
}

// one compare-exchange step of a bitonic sort of (a, b) pairs, lexicographic, with an optional payload
__global__ void __launch_bounds__(256) k_bitonic_step(uint64_t* __restrict__ a, uint64_t* __restrict__ b, uint32_t* __restrict__ v,
                                                      uint32_t padded, uint32_t k, uint32_t j) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t l = i ^ j;
  if (i >= padded || l <= i) return;
  const uint64_t ai = a[i], bi = b[i], al = a[l], bl = b[l];
  const bool greater = ai > al || (ai == al && bi > bl);
  const bool up = (i & k) == 0;
  if (greater == up) {
    a[i] = al; b[i] = bl;
    a[l] = ai; b[l] = bi;
    if (v) {
      const uint32_t t = v[i];
      v[i] = v[l];
      v[l] = t;
    }
  }
}

// sorted events: the thread on the first event of a (key, field) run replays the run against the bucket's state
__global__ void __launch_bounds__(256) k_bucket_replay(const uint64_t* __restrict__ ev_key, const uint64_t* __restrict__ ev_tag,
                                                       const uint32_t* __restrict__ n_dev, uint32_t cap, BucketTable t0, BucketTable t1,
                                                       BucketTable t2, BucketTable t3, uint32_t* __restrict__ err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = min(*n_dev, cap);
  if (i >= n) return;
  const uint64_t k = ev_key[i];
  const uint32_t f = (uint32_t)(ev_tag[i] >> 62);
  if (i > 0 && ev_key[i - 1] == k && (uint32_t)(ev_tag[i - 1] >> 62) == f) return;
  const BucketTable& t = f == 0 ? t0 : f == 1 ? t1 : f == 2 ? t2 : t3;
  const int64_t slot = bucket_find_or_insert(t, k);
  if (slot < 0) {
    atomicOr(err, ERR_XFULL);
    return;
  }
  uint32_t count = t.count[slot];
  uint64_t created = t.created[slot];
  for (uint32_t e = i; e < n && ev_key[e] == k && (uint32_t)(ev_tag[e] >> 62) == f; ++e) {
    const uint64_t tag = ev_tag[e] & TAG_MASK;
    if (tag_is_add(tag)) {
      if (count == 0) created = tag;  // the bucket is (re-)created at the end of the Map order (query:89-93)
      ++count;
    } else if (count) {
      --count;                        // at zero the bucket is deleted (query:114-116)
    }
  }
  t.count[slot] = count;
  t.created[slot] = created;
}

// hit references of a scan (dense: node; overflow: slot | 2^31) -> sort keys (bucket created, entry tag) + node
__global__ void __launch_bounds__(256) k_exact_hit_keys(const uint32_t* __restrict__ refs, uint32_t n, int f, IndexArgs ix, BucketTable t,
                                                        uint64_t* __restrict__ a, uint64_t* __restrict__ b, uint32_t* __restrict__ node_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t ref = refs[i];
  uint64_t key, tag;
  uint32_t node;
  if (ref & 0x80000000u) {
    const uint32_t slot = ref & 0x7FFFFFFFu;
    node = ix.xnode[f][slot];
    key = ix.xkey[f][slot];
    tag = ix.xseq[f][slot];
  } else {
    node = ref;
    key = ix.pcol[f][ref];
    tag = ix.pseq[f][ref];
  }
  const int64_t slot = bucket_find(t, key);
  a[i] = slot >= 0 ? t.created[slot] : ~0ull - 1;
  b[i] = tag;
  node_out[i] = node;
}

}  // namespace bb
