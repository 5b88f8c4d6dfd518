/*
 * bullet_oracle.c - typed CPU restatement of the bullet-js merge path.
 * TEST INFRASTRUCTURE ONLY (checker + cpu_baseline); never linked into
 * libbulletb200.so and never called from the product path.
 *
 * PARITY PIN: bullet-js ships no tests or golden vectors of its own (SURVEY.md
 * 8c) and this image has no JS engine, so the reference's own sources are
 * executed by oracle/minijs (an ECMAScript-subset interpreter written for this
 * purpose) and their outputs are committed as tests/golden/*.json.gz
 * (tests/golden/make_golden.py).  tests/test_golden.py checks this file against them:
 * random JS-level streams (decisions, change set, final table, index contents and
 * query results in the reference's exact Map/Set order) and BASELINE config 1 at
 * full size (100 000 updates over 10 000 records).  Caveat: the interpreter is
 * ours, not V8; its language semantics are unit-tested (tests/test_minijs.py) but
 * have not been compared with node.
 *
 * It consumes the same typed struct-of-arrays format as the library
 * (include/bullet_b200.h describes the encoding) but shares no code with the CUDA
 * kernels: values and clocks are unpacked into ordered key lists - the shape of
 * the JS objects they stand for - and the reference is followed statement by
 * statement:
 *   Bullet._getData            src/bullet.js:115-129
 *   BulletCRT.handleUpdate     src/bullet-crt.js:329-385
 *   incrementVectorClock       src/bullet-crt.js:56-60 (+33-49)
 *   compareVectorClocks        src/bullet-crt.js:68-95
 *   mergeVectorClocks          src/bullet-crt.js:103-114
 *   compare (default)          src/bullet-crt.js:11-15
 *   mergeValues                src/bullet-crt.js:122-153
 *   resolve                    src/bullet-crt.js:164-279
 *   Bullet._applyUpdate        src/bullet.js:184-220
 *   _processSyncEntries loop   src/bullet-network-sync.js:551-569
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../include/bullet_b200.h"

/* ---- JS-object-shaped working types ------------------------------------- */
typedef struct {
  int present;  /* the clock object exists (not undefined) */
  int n;        /* number of own keys */
  uint8_t key[BB_MAX_PEERS];
  uint32_t cnt[BB_MAX_PEERS];
} oclock;

typedef struct {
  int kind; /* BB_KIND_* */
  int n;    /* own keys (OBJ) */
  uint8_t key[BB_MAX_FIELDS];
  uint8_t tag[BB_MAX_FIELDS];
  uint64_t pay[BB_MAX_FIELDS];
} ovalue;

static void clock_unpack(oclock* c, const uint32_t* cnt, uint32_t order, int present) {
  int n = 0;
  for (int s = 0; s < BB_MAX_PEERS; ++s) n += cnt[s] != 0;
  c->present = present;
  c->n = n;
  for (int i = 0; i < n; ++i) {
    c->key[i] = (order >> (4 * i)) & 0xF;
    c->cnt[i] = cnt[c->key[i] & 7];
  }
}

static void clock_pack(const oclock* c, uint32_t* cnt, uint32_t* order) {
  uint32_t o = 0;
  memset(cnt, 0, sizeof(uint32_t) * BB_MAX_PEERS);
  for (int i = 0; i < c->n; ++i) {
    cnt[c->key[i]] = c->cnt[i];
    o |= (uint32_t)c->key[i] << (4 * i);
  }
  *order = o;
}

static int clock_find(const oclock* c, int key) {
  for (int i = 0; i < c->n; ++i)
    if (c->key[i] == key) return i;
  return -1;
}

/* clock[key] || 0 */
static uint32_t clock_get(const oclock* c, int key) {
  int i = clock_find(c, key);
  return i < 0 ? 0 : c->cnt[i];
}

/* clock[key] = v  (existing key keeps its position, new key is appended) */
static void clock_set(oclock* c, int key, uint32_t v) {
  int i = clock_find(c, key);
  if (i < 0) {
    i = c->n++;
    c->key[i] = (uint8_t)key;
  }
  c->cnt[i] = v;
}

static void value_unpack(ovalue* v, uint64_t hdr, const uint64_t* val) {
  v->kind = (int)((hdr >> BB_HDR_KIND_SHIFT) & 3);
  v->n = 0;
  if (v->kind == BB_KIND_PRIM) {
    v->n = 1;
    v->key[0] = 0;
    v->tag[0] = (hdr >> BB_HDR_TAG_SHIFT) & 7;
    v->pay[0] = val[0];
  } else if (v->kind == BB_KIND_OBJ) {
    int n = 0;
    for (int f = 0; f < BB_MAX_FIELDS; ++f) n += ((hdr >> (BB_HDR_TAG_SHIFT + 3 * f)) & 7) != 0;
    v->n = n;
    for (int i = 0; i < n; ++i) {
      int f = (hdr >> (BB_HDR_ORDER_SHIFT + 4 * i)) & 0xF;
      v->key[i] = (uint8_t)f;
      v->tag[i] = (hdr >> (BB_HDR_TAG_SHIFT + 3 * f)) & 7;
      v->pay[i] = val[f];
    }
  }
}

static uint64_t value_pack(const ovalue* v, uint64_t* val) {
  uint64_t hdr = (uint64_t)v->kind << BB_HDR_KIND_SHIFT;
  memset(val, 0, sizeof(uint64_t) * BB_MAX_FIELDS);
  if (v->kind == BB_KIND_PRIM) {
    hdr |= (uint64_t)v->tag[0] << BB_HDR_TAG_SHIFT;
    val[0] = v->pay[0];
  } else if (v->kind == BB_KIND_OBJ) {
    for (int i = 0; i < v->n; ++i) {
      int f = v->key[i];
      hdr |= (uint64_t)v->tag[i] << (BB_HDR_TAG_SHIFT + 3 * f);
      hdr |= (uint64_t)f << (BB_HDR_ORDER_SHIFT + 4 * i);
      val[f] = v->pay[i];
    }
  }
  return hdr;
}

/* ---- JS primitive semantics on typed slots ------------------------------ */
static double as_double(uint64_t bits) {
  double d;
  memcpy(&d, &bits, 8);
  return d;
}

/* ToBoolean(v) == false */
static int prim_falsy(int tag, uint64_t pay) {
  switch (tag) {
    case BB_TAG_NUM: {
      double d = as_double(pay);
      return d == 0.0 || d != d;
    }
    case BB_TAG_BOOL: return pay == 0;
    case BB_TAG_NULL: return 1;
    default: return 0; /* dictionary strings are never "" */
  }
}

/* ToNumber for NUM / BOOL / NULL */
static double prim_number(int tag, uint64_t pay) {
  if (tag == BB_TAG_NUM) return as_double(pay);
  if (tag == BB_TAG_BOOL) return pay ? 1.0 : 0.0;
  return 0.0;
}

/* crt:11-15 on two primitives: a===b -> 0 ; a<b -> -1 ; else +1 */
static int compare_prim(int ta, uint64_t pa, int tb, uint64_t pb) {
  if (ta == tb) {
    if (ta == BB_TAG_NUM) {
      if (as_double(pa) == as_double(pb)) return 0;
    } else if (ta == BB_TAG_NULL) {
      return 0;
    } else if (pa == pb) {
      return 0;
    }
  }
  if (ta == BB_TAG_STR && tb == BB_TAG_STR) return pa < pb ? -1 : 1;
  if (ta == BB_TAG_STR || tb == BB_TAG_STR) return 1; /* ToNumber(non-numeric string) = NaN */
  return prim_number(ta, pa) < prim_number(tb, pb) ? -1 : 1;
}

/* crt:11-15 on two whole values; an object operand is "[object Object]" to `<`
 * and two distinct objects are never === */
static int compare_whole(const bb_config* cfg, const ovalue* x, const ovalue* cur) {
  int xo = x->kind == BB_KIND_OBJ, co = cur->kind == BB_KIND_OBJ;
  if (xo && co) return 1;
  if (xo) return (cur->tag[0] == BB_TAG_STR && cur->pay[0] >= cfg->rank_object) ? -1 : 1;
  if (co) return (x->tag[0] == BB_TAG_STR && x->pay[0] < cfg->rank_object) ? -1 : 1;
  return compare_prim(x->tag[0], x->pay[0], cur->tag[0], cur->pay[0]);
}

static int value_find(const ovalue* v, int key) {
  for (int i = 0; i < v->n; ++i)
    if (v->key[i] == key) return i;
  return -1;
}

/* crt:122-153 */
static void merge_values(const bb_config* cfg, const ovalue* inc, const ovalue* cur, ovalue* out) {
  if (inc->kind != BB_KIND_OBJ || cur->kind != BB_KIND_OBJ) {
    *out = compare_whole(cfg, inc, cur) >= 0 ? *inc : *cur;
    return;
  }
  *out = *cur; /* {...currentValue} */
  for (int i = 0; i < inc->n; ++i) {
    int j = value_find(out, inc->key[i]);
    if (j >= 0) { /* key in result: both leaves are primitives in the flat domain */
      if (compare_prim(inc->tag[i], inc->pay[i], out->tag[j], out->pay[j]) >= 0) {
        out->tag[j] = inc->tag[i];
        out->pay[j] = inc->pay[i];
      }
    } else {
      j = out->n++;
      out->key[j] = inc->key[i];
      out->tag[j] = inc->tag[i];
      out->pay[j] = inc->pay[i];
    }
  }
}

/* crt:68-95 */
static int compare_clocks(const oclock* c1, const oclock* c2, int* same_text) {
  int d1 = 0, d2 = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const oclock* c = pass ? c2 : c1;
    for (int i = 0; i < c->n; ++i) {
      uint32_t v1 = clock_get(c1, c->key[i]), v2 = clock_get(c2, c->key[i]);
      if (v1 > v2) d1 = 1;
      else if (v2 > v1) d2 = 1;
    }
  }
  /* JSON.stringify(c1) === JSON.stringify(c2): same keys, same order, same counts */
  int same = c1->n == c2->n;
  for (int i = 0; same && i < c1->n; ++i)
    same = c1->key[i] == c2->key[i] && c1->cnt[i] == c2->cnt[i];
  *same_text = same;
  if (d1 && d2) return 0;
  if (d1) return 1;
  if (d2) return -1;
  return 0;
}

/* crt:103-114 */
static void merge_clocks(const oclock* c1, const oclock* c2, oclock* out) {
  *out = *c1;
  out->present = 1;
  for (int i = 0; i < c2->n; ++i) {
    uint32_t r = clock_get(out, c2->key[i]);
    clock_set(out, c2->key[i], r > c2->cnt[i] ? r : c2->cnt[i]);
  }
}

/* crt:56-60 with getVectorClock/createVectorClock 33-49 */
static void increment_clock(oclock* v, int me) {
  if (!v->present) {
    v->present = 1;
    v->n = 0;
    clock_set(v, me, 1);
  }
  clock_set(v, me, clock_get(v, me) + 1);
}

/* One setData call. Returns the decision code; *accepted = doUpdate. */
static int step(const bb_config* cfg, bb_row* row, uint64_t seq, const bb_head* uh,
                const uint32_t* uclk, const uint64_t* uval, ovalue* out_val, oclock* out_clk,
                int* accepted) {
  const int me = (int)cfg->local_peer;
  ovalue cur, x, res;
  oclock M, V, I, N;
  int alias = (row->flags & BB_ROW_ALIAS) != 0;
  int code;

  value_unpack(&cur, row->hdr, row->val);
  clock_unpack(&M, row->m_cnt, row->m_order, (row->flags & BB_ROW_M_PRESENT) != 0);
  clock_unpack(&V, row->v_cnt, row->v_order, (row->flags & BB_ROW_V_PRESENT) != 0);
  value_unpack(&x, uh->hdr, uval);

  /* _getData: a missing or falsy value is replaced by {} (src/bullet.js:122-124) */
  if (cur.kind == BB_KIND_NONE) {
    row->cseq = seq + 1;
    cur.kind = BB_KIND_OBJ;
    cur.n = 0;
  } else if (cur.kind == BB_KIND_PRIM && prim_falsy(cur.tag[0], cur.pay[0])) {
    cur.kind = BB_KIND_OBJ;
    cur.n = 0;
  }

  if (uh->hdr & BB_HDR_FLAVOUR_NET) {
    clock_unpack(&I, uclk, uh->clk_order, 1);
  } else {
    increment_clock(&V, me); /* crt:358, in place */
    if (alias) M = V;        /* same object */
    I = V;
  }

  if (!M.present) { /* crt:172-185 */
    increment_clock(&V, me);
    res = x;
    *out_clk = V;
    code = BB_DEC_NO_CURRENT;
    *accepted = 1;
  } else {
    int same_text;
    int c = compare_clocks(&I, &M, &same_text);
    merge_clocks(&I, &M, &N);
    V = N; /* crt:197 */
    alias = 0;
    *out_clk = N;
    if (c == 0 && same_text) {
      int vc = compare_whole(cfg, &x, &cur);
      if (vc == 0) {
        code = BB_DEC_IDENTICAL;
        res = cur;
      } else if (vc > 0) {
        code = BB_DEC_TIE_INCOMING;
        res = x;
      } else {
        code = BB_DEC_TIE_CURRENT;
        res = cur;
      }
    } else if (c > 0) {
      code = BB_DEC_INCOMING;
      res = x;
    } else if (c < 0) {
      code = BB_DEC_HISTORICAL;
      res = cur;
    } else {
      code = BB_DEC_CONCURRENT;
      merge_values(cfg, &x, &cur, &res);
    }
    *accepted = code == BB_DEC_TIE_INCOMING || code == BB_DEC_INCOMING || code == BB_DEC_CONCURRENT;
  }

  if (*accepted) { /* _applyUpdate: store value, meta.vectorClock = the resolver's object */
    cur = res;
    M = *out_clk;
    V = *out_clk;
    alias = 1;
  }
  if ((cfg->flags & BB_CFG_POST_GETDATA) && cur.kind == BB_KIND_PRIM &&
      prim_falsy(cur.tag[0], cur.pay[0])) {
    cur.kind = BB_KIND_OBJ; /* the index hook's _getData (query:151,169) */
    cur.n = 0;
  }
  *out_val = res;

  row->hdr = value_pack(&cur, row->val);
  clock_pack(&M, row->m_cnt, &row->m_order);
  clock_pack(&V, row->v_cnt, &row->v_order);
  row->flags = (M.present ? BB_ROW_M_PRESENT : 0) | (V.present ? BB_ROW_V_PRESENT : 0) |
               (alias ? BB_ROW_ALIAS : 0);
  return code;
}

static void emit(bb_changes* out, uint64_t k, uint64_t i, const bb_head* uh, const ovalue* v,
                 const oclock* c) {
  out->idx[k] = (uint32_t)i;
  out->head[k].hdr = value_pack(v, out->val + k * BB_MAX_FIELDS);
  clock_pack(c, out->clk + k * BB_MAX_PEERS, &out->head[k].clk_order);
  out->head[k].user = uh->user;
}

/* The sequential driver: for (const entry of entries) setData(...)  (sync:551-569). */
int bo_merge_batch(const bb_config* cfg, bb_row* table, uint64_t seq_base, const bb_batch* in,
                   bb_changes* out) {
  uint64_t k = 0;
  for (uint64_t i = 0; i < in->n; ++i) {
    ovalue v;
    oclock c;
    int acc;
    if (in->path_id[i] >= cfg->capacity) return BB_ERR_CAPACITY;
    int code = step(cfg, &table[in->path_id[i]], seq_base + i, &in->head[i],
                    in->clk + i * BB_MAX_PEERS, in->val + i * BB_MAX_FIELDS, &v, &c, &acc);
    out->verdict[i] = ((uint32_t)code << 29) | (acc ? (uint32_t)k : BB_NO_SLOT);
    if (acc) {
      if (k >= out->cap) return BB_ERR_CAPACITY;
      emit(out, k++, i, &in->head[i], &v, &c);
    }
  }
  *out->n_changes = k;
  return BB_OK;
}

/* ---- "all host threads" variant for bench.py --impl reference ------------
 * Paths are independent, so thread t replays, in arrival order,
 * the updates whose path id is congruent to t; the change set is compacted in
 * arrival order afterwards.  Same results as bo_merge_batch. */
typedef struct {
  const bb_config* cfg;
  bb_row* table;
  uint64_t seq_base;
  const bb_batch* in;
  bb_changes* out;
  bb_head* thead;
  uint32_t* tclk;
  uint64_t* tval;
  int t, nt, err;
} mt_arg;

static void* mt_worker(void* p) {
  mt_arg* a = (mt_arg*)p;
  const bb_batch* in = a->in;
  for (uint64_t i = 0; i < in->n; ++i) {
    uint64_t pid = in->path_id[i];
    if (pid % (uint64_t)a->nt != (uint64_t)a->t) continue;
    if (pid >= a->cfg->capacity) {
      a->err = BB_ERR_CAPACITY;
      return 0;
    }
    ovalue v;
    oclock c;
    int acc;
    int code = step(a->cfg, &a->table[pid], a->seq_base + i, &in->head[i],
                    in->clk + i * BB_MAX_PEERS, in->val + i * BB_MAX_FIELDS, &v, &c, &acc);
    a->out->verdict[i] = ((uint32_t)code << 29) | BB_NO_SLOT;
    if (acc) {
      a->thead[i].hdr = value_pack(&v, a->tval + i * BB_MAX_FIELDS);
      clock_pack(&c, a->tclk + i * BB_MAX_PEERS, &a->thead[i].clk_order);
      a->thead[i].user = in->head[i].user;
    }
  }
  return 0;
}

int bo_merge_batch_mt(const bb_config* cfg, bb_row* table, uint64_t seq_base, const bb_batch* in,
                      bb_changes* out, int nthreads) {
  if (nthreads <= 1) return bo_merge_batch(cfg, table, seq_base, in, out);
  if (nthreads > 256) nthreads = 256;
  uint64_t n = in->n;
  bb_head* thead = (bb_head*)malloc(sizeof(bb_head) * (n ? n : 1));
  uint32_t* tclk = (uint32_t*)malloc(sizeof(uint32_t) * BB_MAX_PEERS * (n ? n : 1));
  uint64_t* tval = (uint64_t*)malloc(sizeof(uint64_t) * BB_MAX_FIELDS * (n ? n : 1));
  pthread_t th[256];
  mt_arg args[256];
  int err = BB_OK;
  for (int t = 0; t < nthreads; ++t) {
    mt_arg a = {cfg, table, seq_base, in, out, thead, tclk, tval, t, nthreads, 0};
    args[t] = a;
    pthread_create(&th[t], 0, mt_worker, &args[t]);
  }
  for (int t = 0; t < nthreads; ++t) {
    pthread_join(th[t], 0);
    if (args[t].err) err = args[t].err;
  }
  uint64_t k = 0;
  for (uint64_t i = 0; err == BB_OK && i < n; ++i) {
    if (!BB_DEC_ACCEPTED(BB_VERDICT_CODE(out->verdict[i]))) continue;
    if (k >= out->cap) {
      err = BB_ERR_CAPACITY;
      break;
    }
    out->verdict[i] = (out->verdict[i] & ~BB_NO_SLOT) | (uint32_t)k;
    out->idx[k] = (uint32_t)i;
    out->head[k] = thead[i];
    memcpy(out->clk + k * BB_MAX_PEERS, tclk + i * BB_MAX_PEERS, sizeof(uint32_t) * BB_MAX_PEERS);
    memcpy(out->val + k * BB_MAX_FIELDS, tval + i * BB_MAX_FIELDS, sizeof(uint64_t) * BB_MAX_FIELDS);
    ++k;
  }
  *out->n_changes = k;
  free(thead);
  free(tclk);
  free(tval);
  return err;
}

/* _getData's read-side materialisation for bb_table_read(materialise=1). */
void bo_materialise(bb_row* row, uint64_t seq) {
  ovalue cur;
  value_unpack(&cur, row->hdr, row->val);
  if (cur.kind == BB_KIND_NONE) {
    row->cseq = seq + 1;
  } else if (!(cur.kind == BB_KIND_PRIM && prim_falsy(cur.tag[0], cur.pay[0]))) {
    return;
  }
  cur.kind = BB_KIND_OBJ;
  cur.n = 0;
  row->hdr = value_pack(&cur, row->val);
}

/* ======================================================================== *
 * BulletQuery (src/bullet-query.js) on typed values.  The index is kept in the
 * reference's own shape - a Map (buckets in creation order) of Sets (node ids in
 * insertion order) keyed by String(value) - so results come out in the
 * reference's (Map order, Set order).  String(value) is represented by the
 * 64-bit key of include/bullet_b200.h (BB_KEY_*): equal keys <=> equal strings.
 *   index / _buildIndex / _addToIndex      query:30-94
 *   _removeFromIndex                       query:103-118
 *   _updateIndices (field index)           query:139-167
 *   equals / range / count                 query:186-210, 221-261, 293-313
 * ======================================================================== */
typedef struct {
  uint64_t key;
  uint32_t n, cap;
  uint32_t* nodes; /* Set<path>: insertion order */
  int alive;       /* Map.delete() clears it; a re-created bucket is a new one at the end */
} obucket;

typedef struct bo_index {
  int field;
  obucket* b; /* Map entries in creation order (dead ones skipped) */
  uint64_t nb, capb;
  int64_t* slot; /* open-addressing map key -> bucket index + 1 (0 empty, -1 deleted) */
  uint64_t nslot, nused;
} bo_index;

static uint64_t okey_of(int tag, uint64_t pay) {
  if (tag == BB_TAG_NUM) {
    double d = as_double(pay);
    if (d != d) return BB_KEY_NAN; /* String(NaN) */
    if (d == 0.0) return 0;        /* String(-0) === "0" */
    return pay;
  }
  return (tag == BB_TAG_STR ? BB_KEY_STR : BB_KEY_BOOL) | pay;
}

static uint64_t okey_hash(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  return k;
}

static void omap_rehash(bo_index* ix, uint64_t nslot) {
  free(ix->slot);
  ix->slot = (int64_t*)calloc(nslot, sizeof(int64_t));
  ix->nslot = nslot;
  ix->nused = 0;
  for (uint64_t i = 0; i < ix->nb; ++i) {
    if (!ix->b[i].alive) continue;
    uint64_t h = okey_hash(ix->b[i].key) & (nslot - 1);
    while (ix->slot[h]) h = (h + 1) & (nslot - 1);
    ix->slot[h] = (int64_t)i + 1;
    ++ix->nused;
  }
}

/* index.has(key) ? bucket : NULL */
static obucket* omap_get(bo_index* ix, uint64_t key, uint64_t* where) {
  uint64_t h = okey_hash(key) & (ix->nslot - 1);
  while (ix->slot[h]) {
    if (ix->slot[h] > 0 && ix->b[ix->slot[h] - 1].key == key) {
      if (where) *where = h;
      return &ix->b[ix->slot[h] - 1];
    }
    h = (h + 1) & (ix->nslot - 1);
  }
  return 0;
}

bo_index* bo_index_new(int field) {
  bo_index* ix = (bo_index*)calloc(1, sizeof(bo_index));
  ix->field = field;
  omap_rehash(ix, 1024);
  return ix;
}

void bo_index_free(bo_index* ix) {
  if (!ix) return;
  for (uint64_t i = 0; i < ix->nb; ++i) free(ix->b[i].nodes);
  free(ix->b);
  free(ix->slot);
  free(ix);
}

/* _addToIndex (query:82-94); the caller has dealt with null / undefined */
static void oindex_add(bo_index* ix, uint64_t key, uint32_t node) {
  obucket* b = omap_get(ix, key, 0);
  if (!b) { /* index.set(indexValue, new Set()) - appended to the Map order */
    if ((ix->nused + 1) * 2 > ix->nslot) omap_rehash(ix, ix->nslot * 2);
    if (ix->nb == ix->capb) {
      ix->capb = ix->capb ? ix->capb * 2 : 256;
      ix->b = (obucket*)realloc(ix->b, ix->capb * sizeof(obucket));
    }
    b = &ix->b[ix->nb];
    memset(b, 0, sizeof(*b));
    b->key = key;
    b->alive = 1;
    uint64_t h = okey_hash(key) & (ix->nslot - 1);
    while (ix->slot[h] > 0) h = (h + 1) & (ix->nslot - 1);
    if (ix->slot[h] == 0) ++ix->nused;
    ix->slot[h] = (int64_t)ix->nb + 1;
    ++ix->nb;
  }
  for (uint32_t i = 0; i < b->n; ++i)
    if (b->nodes[i] == node) return; /* Set.add of a member keeps its place */
  if (b->n == b->cap) {
    b->cap = b->cap ? b->cap * 2 : 4;
    b->nodes = (uint32_t*)realloc(b->nodes, b->cap * sizeof(uint32_t));
  }
  b->nodes[b->n++] = node;
}

/* _removeFromIndex (query:103-118) */
static void oindex_remove(bo_index* ix, uint64_t key, uint32_t node) {
  uint64_t where = 0;
  obucket* b = omap_get(ix, key, &where);
  if (!b) return;
  for (uint32_t i = 0; i < b->n; ++i) {
    if (b->nodes[i] != node) continue;
    memmove(b->nodes + i, b->nodes + i + 1, (b->n - i - 1) * sizeof(uint32_t));
    --b->n;
    break;
  }
  if (b->n == 0) { /* index.delete(indexValue) */
    b->alive = 0;
    ix->slot[where] = -1;
  }
}

static int cmp_cseq(const void* a, const void* b) {
  const uint64_t* x = (const uint64_t*)a;
  const uint64_t* y = (const uint64_t*)b;
  return x[0] < y[0] ? -1 : x[0] > y[0];
}

/* _buildIndex (query:53-73): Object.entries(store[path]) in own-key order == the order
 * the children were first created in (row.cseq) */
void bo_index_build(const bb_config* cfg, const bb_row* table, bo_index* ix) {
  uint64_t n = 0;
  uint64_t* ord = (uint64_t*)malloc(2 * sizeof(uint64_t) * (cfg->capacity ? cfg->capacity : 1));
  for (uint64_t i = 0; i < cfg->capacity; ++i) {
    if (!table[i].cseq) continue;
    ord[2 * n] = table[i].cseq;
    ord[2 * n + 1] = i;
    ++n;
  }
  qsort(ord, n, 2 * sizeof(uint64_t), cmp_cseq);
  for (uint64_t j = 0; j < n; ++j) {
    const bb_row* row = &table[ord[2 * j + 1]];
    ovalue v;
    value_unpack(&v, row->hdr, row->val);
    if (v.kind != BB_KIND_OBJ) continue; /* typeof value === "object" && value !== null */
    int k = value_find(&v, ix->field);   /* field in value */
    if (k < 0 || v.tag[k] == BB_TAG_NULL) continue;
    oindex_add(ix, okey_of(v.tag[k], v.pay[k]), (uint32_t)ord[2 * j + 1]);
  }
  free(ord);
}

/* _updateIndices (query:139-167) for a path one segment below the indexed base */
static void oindex_hook(bo_index* ix, uint32_t node, const bb_row* row_after, const bb_head* uh,
                        const uint64_t* uval) {
  ovalue old, x;
  value_unpack(&old, row_after->hdr, row_after->val); /* _getData(indexedPath/part) after the write */
  value_unpack(&x, uh->hdr, uval);                     /* the raw newData argument */
  if (old.kind == BB_KIND_OBJ) {
    int k = value_find(&old, ix->field);
    if (k >= 0 && !prim_falsy(old.tag[k], old.pay[k])) oindex_remove(ix, okey_of(old.tag[k], old.pay[k]), node);
  }
  if (x.kind == BB_KIND_OBJ) {
    int k = value_find(&x, ix->field);
    if (k >= 0 && !prim_falsy(x.tag[k], x.pay[k])) oindex_add(ix, okey_of(x.tag[k], x.pay[k]), node);
  }
}

/* setData with the query wrapper around it (query:16-20): merge, then the hook, per update */
int bo_merge_batch_indexed(const bb_config* cfg, bb_row* table, uint64_t seq_base, const bb_batch* in,
                           bb_changes* out, bo_index** idx, int n_idx) {
  uint64_t k = 0;
  for (uint64_t i = 0; i < in->n; ++i) {
    ovalue v;
    oclock c;
    int acc;
    uint64_t pid = in->path_id[i];
    if (pid >= cfg->capacity) return BB_ERR_CAPACITY;
    int code = step(cfg, &table[pid], seq_base + i, &in->head[i], in->clk + i * BB_MAX_PEERS,
                    in->val + i * BB_MAX_FIELDS, &v, &c, &acc);
    out->verdict[i] = ((uint32_t)code << 29) | (acc ? (uint32_t)k : BB_NO_SLOT);
    if (acc) {
      if (k >= out->cap) return BB_ERR_CAPACITY;
      emit(out, k++, i, &in->head[i], &v, &c);
    }
    for (int j = 0; j < n_idx; ++j)
      oindex_hook(idx[j], (uint32_t)pid, &table[pid], &in->head[i], in->val + i * BB_MAX_FIELDS);
  }
  *out->n_changes = k;
  return BB_OK;
}

/* equals (query:186-210): the bucket's paths in Set order. Returns the full count. */
uint64_t bo_index_equals(bo_index* ix, uint64_t key, uint32_t* out, uint64_t cap) {
  obucket* b = omap_get(ix, key, 0);
  if (!b) return 0;
  for (uint32_t i = 0; i < b->n && i < cap; ++i) out[i] = b->nodes[i];
  return b->n;
}

uint64_t bo_index_count(bo_index* ix, uint64_t key) { /* query:293-313 */
  obucket* b = omap_get(ix, key, 0);
  return b ? b->n : 0;
}

/* `value >= min` / `value <= max` of query:246-251 for one bucket key.  value is
 * Number(key) unless that is NaN, then the key string itself. */
static int obound_ok(uint64_t key, const bb_bound* bd, int upper) {
  uint64_t top = key >> 48;
  if (top == (BB_KEY_STR >> 48)) { /* string vs string: UTF-16 order; vs anything else: NaN */
    uint64_t id = key & 0xFFFFFFFFFFFFULL;
    if (!(bd->flags & BB_BOUND_IS_STRING)) return 0;
    return upper ? id < bd->rank : id >= bd->rank;
  }
  if (top == (BB_KEY_BOOL >> 48)) return (bd->flags & ((key & 1) ? BB_BOUND_TRUE : BB_BOUND_FALSE)) != 0;
  if (key == BB_KEY_NAN) return (bd->flags & BB_BOUND_NAN) != 0;
  double x = as_double(key);
  return upper ? x <= bd->num : x >= bd->num;
}

/* range (query:221-261): every bucket in Map order, its paths in Set order */
uint64_t bo_index_range(bo_index* ix, const bb_bound* lo, const bb_bound* hi, uint32_t* out, uint64_t cap) {
  uint64_t n = 0;
  for (uint64_t i = 0; i < ix->nb; ++i) {
    const obucket* b = &ix->b[i];
    if (!b->alive) continue;
    if (!(obound_ok(b->key, lo, 0) && obound_ok(b->key, hi, 1))) continue;
    for (uint32_t j = 0; j < b->n; ++j, ++n)
      if (n < cap) out[n] = b->nodes[j];
  }
  return n;
}

uint64_t bo_index_entries(const bo_index* ix) {
  uint64_t n = 0;
  for (uint64_t i = 0; i < ix->nb; ++i)
    if (ix->b[i].alive) n += ix->b[i].n;
  return n;
}
