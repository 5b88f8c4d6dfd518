// bb_group.cuh - the resolver with one path's state spread over an 8-lane group.
//
// Lane g (= lane & 7) of a group owns clock slot g of M, V and the incoming clock,
// lanes 0..3 own value slot g (lanes 4..7 mirror slot g & 3); the header words
// (kind/tags, key orders, flags) are replicated.  A clock compare is then one
// compare per lane plus two group ballots, the merged clock one max per lane, the
// per-field max-wins merge one compare per lane - and every row / payload access is
// a coalesced 32-byte (clock, value) or broadcast 16-byte (header) transaction.
// A warp runs four such groups in lockstep.
//
// Same reference lines as bb_merge.cuh (which keeps the one-thread-per-row form used
// by the table import/export kernels): src/bullet.js:115-129, 184-220;
// src/bullet-crt.js:11-15, 56-60, 68-95, 103-114, 122-153, 164-279, 329-385.
#pragma once
#include <stdint.h>

#include "bb_merge.cuh"

namespace bb {

struct GState {
  uint64_t sval;      // S.val[g & 3]
  uint64_t cseq;      // replicated
  uint32_t m, v;      // M.cnt[g], V.cnt[g]
  uint32_t meta, ord; // replicated: kind/tags, own-key order
  uint32_t m_order, v_order, flags;  // replicated
};

struct GLane {
  uint32_t gm;   // lane mask of my group
  int sh;        // lane & 24
  int g;         // lane & 7
  int lane0;     // first lane of my group
};

__device__ __forceinline__ uint32_t gballot(const GLane& L, bool p) { return (__ballot_sync(L.gm, p) >> L.sh) & 0xFFu; }
__device__ __forceinline__ uint64_t gbcast64(const GLane& L, uint64_t v) {  // value held by the group's lane 0
  const uint32_t lo = __shfl_sync(L.gm, (uint32_t)v, L.lane0), hi = __shfl_sync(L.gm, (uint32_t)(v >> 32), L.lane0);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

// crt:56-60 on V
__device__ __forceinline__ void g_increment(const GLane& L, uint32_t me, GState& r) {
  if (!(r.flags & BB_ROW_V_PRESENT)) {
    r.v = (uint32_t)(L.g == (int)me);
    r.v_order = me;
    r.flags |= BB_ROW_V_PRESENT;
  }
  const uint32_t vm = gballot(L, r.v != 0u);
  if (!((vm >> me) & 1u)) r.v_order |= me << (4 * __popc(vm));
  r.v += (uint32_t)(L.g == (int)me);
}

__device__ __forceinline__ void g_empty_object(GState& r) {
  r.meta = BB_KIND_OBJ << BB_HDR_KIND_SHIFT;
  r.ord = 0;
  r.sval = 0;
}

// crt:11-15 on whole values when at least one side is a primitive (slot 0 = group lane 0)
__device__ __forceinline__ int g_compare_whole(const Params& p, const GLane& L, uint32_t xmeta, uint64_t xval,
                                               uint32_t cmeta, uint64_t cval) {
  const uint64_t x0 = gbcast64(L, xval), c0 = gbcast64(L, cval);
  const bool xo = kind_of(xmeta) == BB_KIND_OBJ, co = kind_of(cmeta) == BB_KIND_OBJ;
  if (xo) return (tag_of(cmeta, 0) == BB_TAG_STR && c0 >= p.rank_object) ? -1 : 1;
  if (co) return (tag_of(xmeta, 0) == BB_TAG_STR && x0 < p.rank_object) ? -1 : 1;
  return compare_prim(tag_of(xmeta, 0), x0, tag_of(cmeta, 0), c0);
}

// One setData() on the group's row.  Inputs: incoming clock slot `icnt` + its key
// order, incoming value (xmeta/xord replicated, xval = slot g & 3).  Outputs: the
// emitted clock slot / order and value (valid when the returned code is accepted).
__device__ __forceinline__ uint32_t g_step(const Params& p, const GLane& L, GState& r, bool net, uint32_t icnt,
                                           uint32_t iorder, uint32_t xmeta, uint32_t xord, uint64_t xval,
                                           uint64_t seq, uint32_t& ocnt, uint32_t& oorder, uint32_t& ometa,
                                           uint32_t& oord, uint64_t& oval) {
  {  // _getData: a missing or falsy value becomes {} (src/bullet.js:122-124)
    const uint32_t k = kind_of(r.meta);
    if (k != BB_KIND_OBJ) {
      bool empty = k == BB_KIND_NONE;
      if (empty) r.cseq = seq + 1;
      else empty = prim_falsy(tag_of(r.meta, 0), gbcast64(L, r.sval));
      if (empty) g_empty_object(r);
    }
  }
  if (!net) {  // local flavour: incrementVectorClock in place (crt:358)
    g_increment(L, p.me, r);
    if (r.flags & BB_ROW_ALIAS) {
      r.m = r.v;
      r.m_order = r.v_order;
    }
    icnt = r.v;
    iorder = r.v_order;
  }
  ometa = xmeta;
  oord = xord;
  oval = xval;
  uint32_t code;
  if (!(r.flags & BB_ROW_M_PRESENT)) {  // crt:172-185
    g_increment(L, p.me, r);
    ocnt = r.v;
    oorder = r.v_order;
    code = BB_DEC_NO_CURRENT;
  } else {
    const bool d1 = gballot(L, icnt > r.m) != 0u, d2 = gballot(L, r.m > icnt) != 0u;
    const uint32_t im = gballot(L, icnt != 0u), mm = gballot(L, r.m != 0u);
    ocnt = max(icnt, r.m);  // crt:103-114
    oorder = iorder;
    const uint32_t fresh = mm & ~im;
    if (fresh) {  // keys of M that the incoming clock lacks: appended in M's order
      uint32_t n = __popc(im);
      const uint32_t n2 = __popc(mm);
      for (uint32_t i = 0; i < n2; ++i) {
        const uint32_t s = (r.m_order >> (4 * i)) & 0xFu;
        if ((fresh >> s) & 1u) {
          oorder |= s << (4 * n);
          ++n;
        }
      }
    }
    r.v = ocnt;  // crt:197
    r.v_order = oorder;
    r.flags = (r.flags | BB_ROW_V_PRESENT) & ~BB_ROW_ALIAS;
    const bool both_obj = kind_of(xmeta) == BB_KIND_OBJ && kind_of(r.meta) == BB_KIND_OBJ;
    if (d1 != d2) {
      code = d1 ? BB_DEC_INCOMING : BB_DEC_HISTORICAL;
    } else if (!d1 && iorder == r.m_order) {  // JSON.stringify equal (crt:200-203)
      const int vc = both_obj ? 1 : g_compare_whole(p, L, xmeta, xval, r.meta, r.sval);
      code = vc == 0 ? BB_DEC_IDENTICAL : (vc > 0 ? BB_DEC_TIE_INCOMING : BB_DEC_TIE_CURRENT);
    } else {
      code = BB_DEC_CONCURRENT;  // mergeValues, crt:122-153
      if (both_obj) {
        const int f = L.g & 3;
        const uint32_t tx = tag_of(xmeta, f), tc = tag_of(r.meta, f);
        const bool take = tx != BB_TAG_ABSENT && (tc == BB_TAG_ABSENT || compare_prim(tx, xval, tc, r.sval) >= 0);
        const uint32_t tmask = gballot(L, take) & 0xFu;
        oval = take ? xval : r.sval;
        ometa = r.meta;
        oord = r.ord;
#pragma unroll
        for (int ff = 0; ff < F; ++ff) {
          const uint32_t fm = 7u << (BB_HDR_TAG_SHIFT + 3 * ff);
          if ((tmask >> ff) & 1u) ometa = (ometa & ~fm) | (xmeta & fm);
        }
        const uint32_t px = present_bits(xmeta), pc = present_bits(r.meta);
        const uint32_t freshf = px & ~pc;
        if (freshf) {  // own keys of incoming that current lacks: appended in incoming's order
          uint32_t n = __popc(pc);
          const uint32_t nx = __popc(px);
          for (uint32_t i = 0; i < nx; ++i) {
            const uint32_t fk = (xord >> (4 * i)) & 0xFu;
            if ((freshf >> (3 * fk)) & 1u) {
              oord |= fk << (4 * n);
              ++n;
            }
          }
        }
      } else if (g_compare_whole(p, L, xmeta, xval, r.meta, r.sval) < 0) {
        ometa = r.meta;
        oord = r.ord;
        oval = r.sval;
      }
    }
  }
  if (BB_DEC_ACCEPTED(code)) {  // _applyUpdate: meta.vectorClock = the resolver's object
    r.sval = oval;
    r.meta = ometa;
    r.ord = oord;
    r.m = ocnt;
    r.v = ocnt;
    r.m_order = oorder;
    r.v_order = oorder;
    r.flags = BB_ROW_M_PRESENT | BB_ROW_V_PRESENT | BB_ROW_ALIAS;
  }
  if (p.post_getdata && kind_of(r.meta) == BB_KIND_PRIM &&
      prim_falsy(tag_of(r.meta, 0), gbcast64(L, r.sval)))
    g_empty_object(r);  // the index hook's _getData (query:151,169)
  return code;
}

// ---- lane-distributed row / payload access (row = 32 x u32, see bb_row) ------------
__device__ __forceinline__ void g_load_row(const uint4* row, const GLane& L, GState& r) {
  const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row);
  r.sval = reinterpret_cast<const uint64_t*>(row)[L.g & 3];
  r.m = r32[8 + L.g];
  r.v = r32[16 + L.g];
  const uint4 q6 = row[6], q7 = row[7];
  r.m_order = q6.x;
  r.v_order = q6.y;
  r.meta = q6.z;
  r.ord = q6.w;
  r.flags = q7.x;
  r.cseq = (uint64_t)q7.z | ((uint64_t)q7.w << 32);
}

__device__ __forceinline__ void g_store_row(uint4* row, const GLane& L, const GState& r) {
  uint32_t* r32 = reinterpret_cast<uint32_t*>(row);
  if (L.g < 4) reinterpret_cast<uint64_t*>(row)[L.g] = r.sval;
  r32[8 + L.g] = r.m;
  r32[16 + L.g] = r.v;
  if (L.g == 0) {
    row[6] = make_uint4(r.m_order, r.v_order, r.meta, r.ord);
    row[7] = make_uint4(r.flags, 0u, (uint32_t)r.cseq, (uint32_t)(r.cseq >> 32));
  }
}

// change entry [head 16][clk 32][val 32] as 5 x uint4 at q (shared or global memory)
__device__ __forceinline__ void g_store_entry(uint4* q, const GLane& L, uint32_t user, uint32_t ocnt,
                                              uint32_t oorder, uint32_t ometa, uint32_t oord, uint64_t oval) {
  if (L.g == 0) q[0] = make_uint4(ometa, oord, oorder, user);
  reinterpret_cast<uint32_t*>(q + 1)[L.g] = ocnt;
  if (L.g < 4) reinterpret_cast<uint64_t*>(q + 3)[L.g] = oval;
}

}  // namespace bb
