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
// tags are bit fields (layout: include/bullet_b200.h).
#pragma once
#include <stdint.h>

#include "../../include/bullet_b200.h"

namespace bb {

constexpr int P = BB_MAX_PEERS;
constexpr int F = BB_MAX_FIELDS;

struct Clock {
  uint32_t cnt[P];
  uint32_t order;
  uint32_t present;  // the JS object exists
};

struct Value {
  uint64_t val[F];
  uint64_t hdr;  // kind | tags | key order (flavour bit stripped)
};

struct Params {
  uint64_t rank_object;
  uint32_t me;
  uint32_t post_getdata;
};

__device__ __forceinline__ uint32_t kind_of(uint64_t hdr) { return (uint32_t)(hdr >> BB_HDR_KIND_SHIFT) & 3u; }
__device__ __forceinline__ uint32_t tag_of(uint64_t hdr, int f) {
  return (uint32_t)(hdr >> (BB_HDR_TAG_SHIFT + 3 * f)) & 7u;
}
__device__ __forceinline__ uint32_t tags_of(uint64_t hdr) { return (uint32_t)(hdr >> BB_HDR_TAG_SHIFT) & 0xFFFFFFu; }

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
  uint32_t m1 = clock_mask(c1);
  const uint32_t n2 = __popc(clock_mask(c2));
  uint32_t n = __popc(m1);
  uint32_t order = c1.order;
  for (uint32_t i = 0; i < n2; ++i) {
    const uint32_t s = (c2.order >> (4 * i)) & 0xFu;
    if (!((m1 >> s) & 1u)) {
      order |= s << (4 * n);
      ++n;
      m1 |= 1u << s;
    }
  }
#pragma unroll
  for (int s = 0; s < P; ++s) out.cnt[s] = max(c1.cnt[s], c2.cnt[s]);
  out.order = order;
  out.present = 1;
}

__device__ __forceinline__ bool prim_falsy(uint32_t tag, uint64_t pay) {
  if (tag == BB_TAG_NUM) {
    const double d = __longlong_as_double((long long)pay);
    return d == 0.0 || d != d;
  }
  if (tag == BB_TAG_BOOL) return pay == 0;
  return tag == BB_TAG_NULL;
}

// crt:11-15 on two primitives (=== then <; anything else, NaN included, is +1)
__device__ __forceinline__ int compare_prim(uint32_t ta, uint64_t pa, uint32_t tb, uint64_t pb) {
  const double da = __longlong_as_double((long long)pa), db = __longlong_as_double((long long)pb);
  if (ta == tb) {
    const bool eq = ta == BB_TAG_NUM ? (da == db) : (ta == BB_TAG_NULL ? true : pa == pb);
    if (eq) return 0;
  }
  if (ta == BB_TAG_STR || tb == BB_TAG_STR) {
    if (ta == tb) return pa < pb ? -1 : 1;  // UTF-16 order == dictionary id order
    return 1;                               // ToNumber(non-numeric string) is NaN
  }
  const double na = ta == BB_TAG_NUM ? da : (double)(ta == BB_TAG_BOOL && pa != 0);
  const double nb = tb == BB_TAG_NUM ? db : (double)(tb == BB_TAG_BOOL && pb != 0);
  return na < nb ? -1 : 1;
}

// crt:11-15 on whole values: distinct objects are never ===, and `<` sees an
// object operand as the string "[object Object]".
__device__ __forceinline__ int compare_whole(const Params& p, const Value& x, const Value& cur) {
  const bool xo = kind_of(x.hdr) == BB_KIND_OBJ, co = kind_of(cur.hdr) == BB_KIND_OBJ;
  if (xo && co) return 1;
  if (xo) return (tag_of(cur.hdr, 0) == BB_TAG_STR && cur.val[0] >= p.rank_object) ? -1 : 1;
  if (co) return (tag_of(x.hdr, 0) == BB_TAG_STR && x.val[0] < p.rank_object) ? -1 : 1;
  return compare_prim(tag_of(x.hdr, 0), x.val[0], tag_of(cur.hdr, 0), cur.val[0]);
}

__device__ __forceinline__ uint32_t nkeys(uint64_t hdr) {
  const uint32_t t = tags_of(hdr);
  return __popc((t | (t >> 1) | (t >> 2)) & 0x249249u);  // one bit per non-ABSENT slot
}

// crt:122-153 (flat records: every leaf compare is primitive vs primitive)
__device__ __forceinline__ void merge_values(const Params& p, const Value& x, const Value& cur, Value& out) {
  if (kind_of(x.hdr) != BB_KIND_OBJ || kind_of(cur.hdr) != BB_KIND_OBJ) {
    out = compare_whole(p, x, cur) >= 0 ? x : cur;
    return;
  }
  out = cur;  // {...currentValue}
  const uint32_t nx = nkeys(x.hdr);
  uint32_t n = nkeys(cur.hdr);
  for (uint32_t i = 0; i < nx; ++i) {  // own keys of incoming that are new: appended in its order
    const uint32_t f = (uint32_t)(x.hdr >> (BB_HDR_ORDER_SHIFT + 4 * i)) & 0xFu;
    if (tag_of(cur.hdr, f) == BB_TAG_ABSENT) {
      out.hdr |= (uint64_t)f << (BB_HDR_ORDER_SHIFT + 4 * n);
      ++n;
    }
  }
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const uint32_t tx = tag_of(x.hdr, f), tc = tag_of(cur.hdr, f);
    if (tx == BB_TAG_ABSENT) continue;
    if (tc == BB_TAG_ABSENT || compare_prim(tx, x.val[f], tc, cur.val[f]) >= 0) {
      out.hdr = (out.hdr & ~(7ull << (BB_HDR_TAG_SHIFT + 3 * f))) | ((uint64_t)tx << (BB_HDR_TAG_SHIFT + 3 * f));
      out.val[f] = x.val[f];
    }
  }
}

__device__ __forceinline__ void materialise_empty_object(Value& v) {
  v.hdr = (uint64_t)BB_KIND_OBJ << BB_HDR_KIND_SHIFT;
#pragma unroll
  for (int f = 0; f < F; ++f) v.val[f] = 0;
}

struct RowState {
  Value s;
  Clock m, v;
  uint32_t alias;
  uint64_t cseq;
};

// One setData(): mutates the row state, returns the decision code, fills the
// emitted (value, clock) when accepted.
__device__ __forceinline__ uint32_t resolve_step(const Params& p, RowState& r, uint64_t uhdr,
                                                 const Clock& uclk, const Value& x, uint64_t seq,
                                                 Value& out_val, Clock& out_clk) {
  const uint32_t ck = kind_of(r.s.hdr);
  if (ck == BB_KIND_NONE) {
    r.cseq = seq + 1;
    materialise_empty_object(r.s);
  } else if (ck == BB_KIND_PRIM && prim_falsy(tag_of(r.s.hdr, 0), r.s.val[0])) {
    materialise_empty_object(r.s);
  }

  Clock inc;
  if (uhdr & BB_HDR_FLAVOUR_NET) {
    inc = uclk;
    inc.present = 1;
  } else {
    clock_increment(r.v, p.me);  // crt:358, in place
    if (r.alias) r.m = r.v;      // M is the same object
    inc = r.v;
  }

  uint32_t code;
  if (!r.m.present) {  // crt:172-185
    clock_increment(r.v, p.me);
    out_clk = r.v;
    out_val = x;
    code = BB_DEC_NO_CURRENT;
  } else {
    bool d1 = false, d2 = false;
#pragma unroll
    for (int s = 0; s < P; ++s) {
      d1 |= inc.cnt[s] > r.m.cnt[s];
      d2 |= r.m.cnt[s] > inc.cnt[s];
    }
    clock_merge(inc, r.m, out_clk);
    r.v = out_clk;  // crt:197
    r.alias = 0;
    if (!d1 && !d2 && inc.order == r.m.order) {  // JSON.stringify equal (crt:200-203)
      const int vc = compare_whole(p, x, r.s);
      code = vc == 0 ? BB_DEC_IDENTICAL : (vc > 0 ? BB_DEC_TIE_INCOMING : BB_DEC_TIE_CURRENT);
      out_val = x;
    } else if (d1 && !d2) {
      code = BB_DEC_INCOMING;
      out_val = x;
    } else if (d2 && !d1) {
      code = BB_DEC_HISTORICAL;
      out_val = x;
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
  if (p.post_getdata && kind_of(r.s.hdr) == BB_KIND_PRIM && prim_falsy(tag_of(r.s.hdr, 0), r.s.val[0]))
    materialise_empty_object(r.s);  // the index hook's _getData (query:151,169)
  return code;
}

}  // namespace bb
