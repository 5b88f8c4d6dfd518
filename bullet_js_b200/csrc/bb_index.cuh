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
};

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

// (node, k) is known to be absent; false when the set is full
__device__ __forceinline__ bool x_insert(const IndexArgs& ix, int f, uint32_t node, uint64_t k) {
  const uint32_t m = ix.xmask[f];
  uint32_t i = x_hash(node, k) & m;
  for (uint32_t probe = 0; probe <= m; ++probe, i = (i + 1) & m) {
    const uint32_t n = __ldcg(ix.xnode[f] + i);
    if (n != X_EMPTY && n != X_TOMB) continue;
    if (n == X_EMPTY && atomicAdd(ix.xused + f, 1u) >= m - (m >> 3)) {  // keep 1/8 of the slots EMPTY
      atomicSub(ix.xused + f, 1u);
      return false;
    }
    if (atomicCAS(ix.xnode[f] + i, n, node) == n) {
      ix.xkey[f][i] = k;
      return true;
    }
    if (n == X_EMPTY) atomicSub(ix.xused + f, 1u);  // another node took the slot
  }
  return false;
}

__device__ __forceinline__ uint32_t xcnt_get(uint32_t xcnt, int f) { return (xcnt >> (8 * f)) & 0xFFu; }

// _updateIndices (query:139-176) for one setData on `node`: `after` is the node as _getData
// sees it after the write, `x` the raw incoming value.  prim[f] caches pcol[f][node].
__device__ __forceinline__ void index_hook(const IndexArgs& ix, uint32_t node, const Value& after, const Value& x,
                                           uint64_t (&prim)[F], uint32_t& xcnt, uint32_t* err) {
#pragma unroll
  for (int f = 0; f < F; ++f) {
    if (!((ix.mask >> f) & 1u)) continue;
    if (kind_of(after.meta) == BB_KIND_OBJ) {  // if (oldData && oldData[field]) remove
      const uint32_t t = tag_of(after.meta, f);
      if (t != BB_TAG_ABSENT && !prim_falsy(t, after.val[f])) {
        const uint64_t k = canon_key(t, after.val[f]);
        if (prim[f] == k) {
          prim[f] = BB_KEY_NONE;
        } else if (xcnt_get(xcnt, f)) {
          const int64_t slot = x_find(ix, f, node, k);
          if (slot >= 0) {
            x_remove_at(ix, f, slot);
            if (xcnt_get(xcnt, f) != 0xFFu) xcnt -= 1u << (8 * f);
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
        } else if (x_insert(ix, f, node, k)) {
          if (xcnt_get(xcnt, f) != 0xFFu) xcnt += 1u << (8 * f);
        } else {
          atomicOr(err, ERR_XFULL);
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
      if (d < a.cap) a.out[d] = a.nodes ? a.nodes[e] : (uint32_t)e;
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
      if (d < a.cap) a.out[d] = a.nodes ? a.nodes[e] : (uint32_t)e;
      else overflow = true;
      ++d;
    }
    if (hb & 2u) {
      if (d < a.cap) a.out[d] = a.nodes ? a.nodes[e + 1] : (uint32_t)(e + 1);
      else overflow = true;
    }
  }
  if (overflow) atomicOr(a.err, ERR_HITS);
}

}  // namespace bb
