// bb_merge.cuh - device-side resolver: one setData() call on bit-packed state.
//
// Follows (paths relative to the reference repo):
//   Bullet._getData falsy materialisation   src/bullet.js:115-129
//   BulletCRT.handleUpdate                  src/bullet-crt.js:329-385
//   incrementVectorClock                    src/bullet-crt.js:56-60 (+33-49)
//   compareVectorClocks / mergeVectorClocks src/bullet-crt.js:68-95 / 103-114
//   compare (default) / mergeValues         src/bullet-crt.js:11-15 / 122-153
//   resolve                                 src/bullet-crt.js:164-279
//   Bullet._applyUpdate (state part)        src/bullet.js:184-220
//
// Everything is kept in registers: clock counts and value payloads are fixed
// arrays that are only ever indexed by unrolled loop counters; key orders and
// tags are 32-bit bit fields (layout: include/bullet_b200.h; the 64-bit header
// word is split into `meta` = low half (kind, tags) and `ord` = high half (key
// order) so that no 64-bit shifts are needed).
#pragma once
#include <stdint.h>

#include "../../include/bullet_b200.h"

namespace bb {

constexpr int P = BB_MAX_PEERS;
constexpr int F = BB_MAX_FIELDS;

// deferred device errors: word 0 of the ctx's error block (sticky until bb_sync), word 1 = ordinal of the first
// rejected batch.  A batch with an out-of-range path id is rejected WHOLE: its own reject word tells the
// kernels of that batch (and only that batch) to touch nothing.
constexpr uint32_t ERR_RANGE = 1u, ERR_CHANGES = 2u, ERR_XFULL = 4u, ERR_HITS = 8u;

struct Clock {
  uint32_t cnt[P];
  uint32_t order;
  uint32_t present;  // the JS object exists
};

struct Value {
  uint64_t val[F];
  uint32_t meta;  // low half of the header word, flavour bit stripped
  uint32_t ord;   // high half: own-key order nibbles
};

struct Params {
  uint64_t rank_object;
  uint32_t me;
  uint32_t post_getdata;
};

__device__ __forceinline__ uint32_t kind_of(uint32_t meta) { return (meta >> BB_HDR_KIND_SHIFT) & 3u; }
__device__ __forceinline__ uint32_t tag_of(uint32_t meta, int f) { return (meta >> (BB_HDR_TAG_SHIFT + 3 * f)) & 7u; }
// one bit per own key, at bit 3f
__device__ __forceinline__ uint32_t present_bits(uint32_t meta) {
  const uint32_t t = meta >> BB_HDR_TAG_SHIFT;
  return (t | (t >> 1) | (t >> 2)) & 0x249249u;
}

__device__ __forceinline__ uint32_t clock_mask(const Clock& c) {
  uint32_t m = 0;
#pragma unroll
  for (int s = 0; s < P; ++s) m |= (c.cnt[s] != 0u) << s;
  return m;
}

// crt:56-60: V absent -> {me:1}; V[me] = (V[me] || 0) + 1 (a new key is appended)
__device__ __forceinline__ void clock_increment(Clock& v, uint32_t me) {
  if (!v.present) {
#pragma unroll
    for (int s = 0; s < P; ++s) v.cnt[s] = (s == (int)me) ? 1u : 0u;
    v.order = me;
    v.present = 1;
  }
  const uint32_t m = clock_mask(v);
  if (!((m >> me) & 1u)) v.order |= me << (4 * __popc(m));
#pragma unroll
  for (int s = 0; s < P; ++s) v.cnt[s] += (s == (int)me) ? 1u : 0u;
}

// crt:103-114: {...c1}, then c2's keys (max); keys new to c1 appended in c2's order
__device__ __forceinline__ void clock_merge(const Clock& c1, const Clock& c2, Clock& out) {
  const uint32_t m1 = clock_mask(c1), m2 = clock_mask(c2);
  uint32_t order = c1.order;
  uint32_t fresh = m2 & ~m1;
  if (fresh) {  // rare: the key sets usually coincide
    uint32_t n = __popc(m1);
    const uint32_t n2 = __popc(m2);
    for (uint32_t i = 0; i < n2; ++i) {
      const uint32_t s = (c2.order >> (4 * i)) & 0xFu;
      if ((fresh >> s) & 1u) {
        order |= s << (4 * n);
        ++n;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < P; ++s) out.cnt[s] = max(c1.cnt[s], c2.cnt[s]);
  out.order = order;
  out.present = 1;
}

__device__ __forceinline__ bool prim_falsy(uint32_t tag, uint64_t pay) {
  if (tag == BB_TAG_NUM) return (pay << 1) == 0 || (pay << 1) > 0xFFE0000000000000ull;  // +-0 or NaN
  if (tag == BB_TAG_BOOL) return pay == 0;
  return tag == BB_TAG_NULL;
}

// crt:11-15 on two primitives (=== then <; anything else, NaN included, is +1)
__device__ __forceinline__ int compare_prim(uint32_t ta, uint64_t pa, uint32_t tb, uint64_t pb) {
  if (ta == BB_TAG_NUM && tb == BB_TAG_NUM) {
    const double da = __longlong_as_double((long long)pa), db = __longlong_as_double((long long)pb);
    return da == db ? 0 : (da < db ? -1 : 1);
  }
  if (ta == tb) {
    if (ta == BB_TAG_NULL || pa == pb) return 0;
    return pa < pb ? -1 : 1;  // STR: dictionary id order == UTF-16 order; BOOL: false < true
  }
  if (ta == BB_TAG_STR || tb == BB_TAG_STR) return 1;  // ToNumber(non-numeric string) is NaN
  const double na = ta == BB_TAG_NUM ? __longlong_as_double((long long)pa) : (double)(ta == BB_TAG_BOOL && pa != 0);
  const double nb = tb == BB_TAG_NUM ? __longlong_as_double((long long)pb) : (double)(tb == BB_TAG_BOOL && pb != 0);
  return na < nb ? -1 : 1;
}

// crt:11-15 on whole values: distinct objects are never ===, and `<` sees an
// object operand as the string "[object Object]".
__device__ __forceinline__ int compare_whole(const Params& p, const Value& x, const Value& cur) {
  const bool xo = kind_of(x.meta) == BB_KIND_OBJ, co = kind_of(cur.meta) == BB_KIND_OBJ;
  if (xo && co) return 1;
  if (xo) return (tag_of(cur.meta, 0) == BB_TAG_STR && cur.val[0] >= p.rank_object) ? -1 : 1;
  if (co) return (tag_of(x.meta, 0) == BB_TAG_STR && x.val[0] < p.rank_object) ? -1 : 1;
  return compare_prim(tag_of(x.meta, 0), x.val[0], tag_of(cur.meta, 0), cur.val[0]);
}

// crt:122-153 (flat records: every leaf compare is primitive vs primitive)
__device__ __forceinline__ void merge_values(const Params& p, const Value& x, const Value& cur, Value& out) {
  if (kind_of(x.meta) != BB_KIND_OBJ || kind_of(cur.meta) != BB_KIND_OBJ) {
    out = compare_whole(p, x, cur) >= 0 ? x : cur;
    return;
  }
  out = cur;  // {...currentValue}
  const uint32_t px = present_bits(x.meta), pc = present_bits(cur.meta);
  const uint32_t fresh = px & ~pc;
  if (fresh) {  // own keys of incoming that current lacks: appended in incoming's order
    uint32_t n = __popc(pc);
    const uint32_t nx = __popc(px);
    for (uint32_t i = 0; i < nx; ++i) {
      const uint32_t f = (x.ord >> (4 * i)) & 0xFu;
      if ((fresh >> (3 * f)) & 1u) {
        out.ord |= f << (4 * n);
        ++n;
      }
    }
  }
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const uint32_t tx = tag_of(x.meta, f), tc = tag_of(cur.meta, f);
    if (tx != BB_TAG_ABSENT && (tc == BB_TAG_ABSENT || compare_prim(tx, x.val[f], tc, cur.val[f]) >= 0)) {
      out.meta = (out.meta & ~(7u << (BB_HDR_TAG_SHIFT + 3 * f))) | (tx << (BB_HDR_TAG_SHIFT + 3 * f));
      out.val[f] = x.val[f];
    }
  }
}

__device__ __forceinline__ void materialise_empty_object(Value& v) {
  v.meta = BB_KIND_OBJ << BB_HDR_KIND_SHIFT;
  v.ord = 0;
#pragma unroll
  for (int f = 0; f < F; ++f) v.val[f] = 0;
}

__device__ __forceinline__ bool falsy_primitive(const Value& v) {
  return kind_of(v.meta) == BB_KIND_PRIM && prim_falsy(tag_of(v.meta, 0), v.val[0]);
}

struct RowState {
  Value s;
  Clock m, v;
  uint32_t alias;
  uint32_t xcnt;  // per-field entry counts in the index overflow sets (bb_index.cuh)
  uint64_t cseq;
};

// One setData(): mutates the row state, returns the decision code, fills the
// emitted (value, clock) when accepted.  `net` = network-with-clock flavour.
__device__ __forceinline__ uint32_t resolve_step(const Params& p, RowState& r, bool net, const Clock& uclk,
                                                 const Value& x, uint64_t seq, Value& out_val, Clock& out_clk) {
  if (kind_of(r.s.meta) == BB_KIND_NONE) {
    r.cseq = seq + 1;
    materialise_empty_object(r.s);
  } else if (falsy_primitive(r.s)) {
    materialise_empty_object(r.s);
  }

  Clock inc;
  if (net) {
    inc = uclk;
  } else {
    clock_increment(r.v, p.me);  // crt:358, in place
    if (r.alias) r.m = r.v;      // M is the same object
    inc = r.v;
  }

  uint32_t code;
  out_val = x;
  if (!r.m.present) {  // crt:172-185
    clock_increment(r.v, p.me);
    out_clk = r.v;
    code = BB_DEC_NO_CURRENT;
  } else {
    // compareVectorClocks (crt:68-95) and mergeVectorClocks (crt:103-114) in one pass over the slots; the
    // key-order work of the merge is only needed when current has a key incoming lacks (rare)
    bool d1 = false, d2 = false, fresh = false;
#pragma unroll
    for (int s = 0; s < P; ++s) {
      const uint32_t ci = inc.cnt[s], cm = r.m.cnt[s];
      d1 |= ci > cm;
      d2 |= cm > ci;
      fresh |= (ci == 0u) & (cm != 0u);
      out_clk.cnt[s] = max(ci, cm);
    }
    out_clk.order = inc.order;
    out_clk.present = 1;
    if (fresh) clock_merge(inc, r.m, out_clk);
    r.v = out_clk;  // crt:197
    r.alias = 0;
    if (d1 != d2) {
      code = d1 ? BB_DEC_INCOMING : BB_DEC_HISTORICAL;
    } else if (!d1 && inc.order == r.m.order) {  // JSON.stringify equal (crt:200-203)
      const int vc = compare_whole(p, x, r.s);
      code = vc == 0 ? BB_DEC_IDENTICAL : (vc > 0 ? BB_DEC_TIE_INCOMING : BB_DEC_TIE_CURRENT);
    } else {
      code = BB_DEC_CONCURRENT;
      merge_values(p, x, r.s, out_val);
    }
  }
  if (BB_DEC_ACCEPTED(code)) {  // _applyUpdate: meta.vectorClock = the resolver's object
    r.s = out_val;
    r.m = out_clk;
    r.v = out_clk;
    r.alias = 1;
  }
  if (p.post_getdata && falsy_primitive(r.s)) materialise_empty_object(r.s);  // hook's _getData (query:151,169)
  return code;
}

}  // namespace bb
