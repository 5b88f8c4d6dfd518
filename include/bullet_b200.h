/*
 * bullet_b200.h - C ABI of libbulletb200.so, the B200-native drop-in for ONE hot
 * path of KORandi/bullet-js: batched conflict-resolution merge of graph updates
 * (src/bullet-crt.js as driven by src/bullet-network-sync.js:551-569 through
 * src/bullet.js:139-220) plus index build and equals/range/count
 * (src/bullet-query.js).  Paths below are relative to the reference repo root.
 *
 * The reference has NO FFI / addon / operator registry for this path
 * (SURVEY.md 8b), so there is nothing to mirror symbol-for-symbol; each entry
 * point names the reference function(s) whose batched equivalent it is.  The
 * reference-side binding is js/bullet-b200.js (the shim that answers
 * `bullet.crt.handleUpdate` and `sync._processSyncEntries` from this library on an
 * unmodified Bullet instance) over the N-API addon sketched in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; returns 0 or a negative BB_ERR_*;
 * never aborts, never calls back into the host; host buffers are caller-owned
 * (pinned memory recommended); the library owns all device memory; outputs are
 * only defined on success; a bb_ctx is not thread-safe; results are
 * deterministic (decisions, stored rows and the change set reached through
 * verdict[] are bit-identical across runs).  There is NO CPU fallback: every
 * compute entry point fails with BB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef BULLET_B200_H
#define BULLET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB_ABI_VERSION 1
#define BB_MAX_PEERS 8   /* clock slots per path (peer ids interned by the host) */
#define BB_MAX_FIELDS 4  /* value slots per record in this build (128-byte rows) */

/* ---- error codes ------------------------------------------------------- */
#define BB_OK 0
#define BB_ERR_ARG (-1)      /* null pointer, bad size, bad handle */
#define BB_ERR_CUDA (-2)     /* CUDA runtime / no device; see bb_last_error */
#define BB_ERR_DOMAIN (-3)   /* input outside the bit-exact domain */
#define BB_ERR_CAPACITY (-4) /* path id >= capacity, or output buffer too small */
#define BB_ERR_STATE (-5)    /* e.g. query on an index that was never built */

/* ---- typed JS values ---------------------------------------------------
 * A field slot is (tag, 64-bit payload):
 *   ABSENT  the key is not an own property of the record
 *   NUM     IEEE-754 binary64 bits (NaN, +-0, +-Infinity allowed)
 *   STR     id in the host's order-preserving string dictionary: id order ==
 *           UTF-16 code-unit order (what JS `<` uses for two strings).  The
 *           dictionary must not contain strings Number() can parse, nor
 *           "true", "false", "NaN", "[object Object]" (SURVEY 8a restriction 6)
 *   BOOL    payload 0 / 1
 *   NULL    JS null
 */
#define BB_TAG_ABSENT 0u
#define BB_TAG_NUM 1u
#define BB_TAG_STR 2u
#define BB_TAG_BOOL 3u
#define BB_TAG_NULL 4u

/* A whole value is a flat record (KIND_OBJ: the slots with tag != ABSENT are its
 * own keys; zero keys is `{}`) or a primitive (KIND_PRIM: slot 0).  KIND_NONE is
 * "the path was never written" and only occurs in table rows. */
#define BB_KIND_NONE 0u
#define BB_KIND_OBJ 1u
#define BB_KIND_PRIM 2u

/* Value header word (`hdr`), used by updates, change-set entries and rows:
 *   bit  0      flavour, updates only: 1 = from the network WITH a clock
 *               (src/bullet-crt.js:339-344), 0 = local flavour (also what a
 *               primitive arriving from the network gets, sync:560-563)
 *   bits 1-2    BB_KIND_*
 *   bits 8-31   3-bit tag of slot f at bit 8+3f
 *   bits 32-63  own-key order: nibble i (bit 32+4i) = slot id of the i-th key,
 *               for i < (number of non-ABSENT slots); unused nibbles are 0.
 *               (JS object key order is part of the emitted value.)
 */
#define BB_HDR_FLAVOUR_NET 1ull
#define BB_HDR_KIND_SHIFT 1
#define BB_HDR_TAG_SHIFT 8
#define BB_HDR_ORDER_SHIFT 32

/* Clock: cnt[s] > 0 <=> peer slot s is a key of the clock object; `order` holds
 * the key order (nibble i = slot of the i-th key, unused nibbles 0).  Key order
 * matters: it decides "identical" vs "concurrent" (src/bullet-crt.js:200-203). */

/* 16-byte per-update / per-change head. */
typedef struct bb_head {
  uint64_t hdr;       /* value header word, see above */
  uint32_t clk_order; /* key order of the clock that travels with it */
  uint32_t user;      /* opaque to the library; copied from update to change entry */
} bb_head;

/* Table row, one per interned path id; 128 bytes.  This struct is the IMPORT / EXPORT format of bb_table_load and
 * bb_table_read; in HBM the library keeps the same eight 16-byte chunks in the order 0, 1, 6, 7, 2, 3, 4, 5 (values,
 * header, flags and cseq in the first 64 bytes, the clock counts in the second: an index build reads half a row).
 * State per path is (S, M, V, a) of SURVEY.md 8a:
 *   S = val/hdr, M = meta[path].vectorClock, V = crt.vectorClocks.get(path),
 *   a = M and V are the same JS object (src/bullet.js:198-203 stores the
 *   object the resolver keeps, so in-place increments are shared). */
#define BB_ROW_M_PRESENT 1u
#define BB_ROW_V_PRESENT 2u
#define BB_ROW_ALIAS 4u
typedef struct bb_row {
  uint64_t val[BB_MAX_FIELDS];   /*   0 */
  uint32_t m_cnt[BB_MAX_PEERS];  /*  32 */
  uint32_t v_cnt[BB_MAX_PEERS];  /*  64 */
  uint32_t m_order;              /*  96 */
  uint32_t v_order;              /* 100 */
  uint64_t hdr;                  /* 104: kind, tags, key order (flavour bit unused) */
  uint32_t flags;                /* 112: BB_ROW_* */
  uint32_t xcnt;                 /* 116: byte f = entries this node has in the overflow set of
                                         the index on field f (saturates at 255 = unknown) */
  uint64_t cseq;                 /* 120: 1 + global sequence of the update that first
                                         touched the row (== key order of the parent
                                         collection object, query:61); 0 = never */
} bb_row;

/* Decision codes: the 6 `reason` strings of src/bullet-crt.js:182,216,230,245,
 * 260,276; the :230 one split by the sign of the value comparison. */
#define BB_DEC_NO_CURRENT 0u    /* accepted */
#define BB_DEC_IDENTICAL 1u     /* rejected */
#define BB_DEC_TIE_INCOMING 2u  /* accepted */
#define BB_DEC_TIE_CURRENT 3u   /* rejected */
#define BB_DEC_INCOMING 4u      /* accepted */
#define BB_DEC_HISTORICAL 5u    /* rejected */
#define BB_DEC_CONCURRENT 6u    /* accepted (field-wise max merge) */
#define BB_DEC_ACCEPTED(d) ((0x55u >> (d)) & 1u) /* doUpdate, crt:383 */

typedef struct bb_config {
  uint32_t abi_version; /* BB_ABI_VERSION */
  int32_t device;       /* CUDA ordinal */
  uint32_t n_fields;    /* 1..BB_MAX_FIELDS */
  uint32_t local_peer;  /* slot of bullet.id (`me`) */
  uint64_t capacity;    /* rows; path ids are 0..capacity-1 */
  /* Where the strings JS manufactures on this path would sort in the host's
   * dictionary: the smallest dictionary id that is greater than the string.
   * "[object Object]" is what `<` sees for an object operand (ToPrimitive). */
  uint64_t rank_object;
  uint64_t rank_true;   /* String(true)  as a range() bucket key */
  uint64_t rank_false;  /* String(false) */
  uint64_t rank_nan;    /* String(NaN)   */
  uint32_t flags;       /* BB_CFG_* */
  uint32_t reserved;
} bb_config;
/* The index hook is installed for these paths: `_updateIndices` re-reads the node
 * with `_getData` after every setData (query:151,169), which turns a falsy stored
 * primitive into `{}` right after the write instead of at the next update. */
#define BB_CFG_POST_GETDATA 1u
/* Lay the change set out in path-major order (ascending path id, arrival order within a path) and query hits in
 * ascending node id, bit-identical from run to run: the batch is sorted by path id first (counting sort, or the LSD
 * radix sort when the table is much larger than the batch) and merged by k_merge_stage.  Without it (the default) the
 * batch is not sorted at all: tiles of 256 consecutive updates merge where they lie and claim their slice of the
 * change set as they finish - the same entries, each still found through verdict[], arrival order inside a tile, but
 * the tiles' order in the buffers is not fixed and no tile ever waits for another one. */
#define BB_CFG_ORDERED_CHANGES 2u
/* Sort every batch by path id with the stable LSD radix sort / with the counting sort over the row indices before
 * merging (same decisions, table and change entries as the default; kept for A/B measurements). */
#define BB_CFG_RADIX_SORT 4u
#define BB_CFG_FULL_SORT 8u
/* Hot keys (a path that takes thousands of a batch's updates is a serial chain) are handled by the default pipeline:
 * a whole CTA evaluates 256 of the path's updates per round against the row and retires everything up to the first
 * state-changing one, indexed collections included.  The round-1 opt-in flag is accepted and ignored. */
#define BB_CFG_HOT_KEYS 64u
/* Compact change set.  Most accepted updates win outright ("incoming dominates", "identical clocks: incoming"): what the
 * reference stores (decision.value, decision.vectorClock) is then the incoming update itself, bit for bit, and shipping
 * it back to the caller that just sent it wastes the return link.  With this flag such an update gets NO entry: its
 * verdict carries slot BB_SLOT_ECHO and its entry is by definition
 *   idx = i, head = { in.head[i].hdr & ~BB_HDR_FLAVOUR_NET, in.head[i].clk_order, in.head[i].user },
 *   clk = in.clk[i], val = in.val[i].
 * Entries are emitted only where the stored state differs from the update: first writes (the incoming clock is
 * discarded, crt:172-185), local puts (the clock is V, crt:358), concurrent merges (crt:266-278).  Decisions, table
 * and index are identical to the default; n_changes counts emitted entries only. */
#define BB_CFG_COMPACT_CHANGES 128u
/* Keep meta[path].lastModified (src/bullet.js:201) per row, at merge-call granularity: every accepted update stamps its
 * row with the ordinal of the merge call it arrived in (bb_epoch).  The host maps ordinals to wall-clock time (one
 * Date.now() per call); bb_sync_collect filters on it.  4 bytes per row. */
#define BB_CFG_TRACK_MODIFIED 256u
/* Exact query order.  The reference returns equals / range results in (Map order of the buckets, Set order
 * inside a bucket) = (bucket creation, insertion) order (src/bullet-query.js:89-93, 110-116, 204, 237-258), a node twice
 * when it has entries in two matching buckets.  With this flag every index entry carries the sequence tag of the add
 * that inserted it, the merge logs the hook's effective adds / removes, a per-bucket replay after every batch (and after
 * an index build) keeps key -> (count, creation tag), and bb_query_equals / bb_query_range sort their hits by
 * (bucket creation, entry tag): the reference's exact list.  Costs a sort of the batch's events per merge call and
 * 8 bytes per entry; implies the sorting front end (like BB_CFG_FULL_SORT).  The _dev and router query entry points
 * keep returning multisets. */
#define BB_CFG_EXACT_ORDER 512u

typedef struct bb_ctx bb_ctx;

/* One batch of updates in arrival order == the `entries` argument of
 * BulletNetworkSync._processSyncEntries (sync:551-569), or a run of
 * BulletNode.put / _handlePut calls. Struct-of-arrays, n elements each. */
typedef struct bb_batch {
  uint64_t n;
  const uint64_t* path_id; /* [n] interned path */
  const bb_head* head;     /* [n] */
  const uint32_t* clk;     /* [n][BB_MAX_PEERS] incoming clock counts (net flavour) */
  const uint64_t* val;     /* [n][BB_MAX_FIELDS] */
} bb_batch;

/* Per-update verdict + the emitted change set.
 * verdict[i] (arrival order, every update) = BB_DEC_* code << 29 | slot, where slot
 * is the position of update i's entry in idx/head/clk/val, or BB_NO_SLOT when the
 * update was rejected (doUpdate == false, src/bullet-crt.js:383).
 * The change set == the _applyUpdate calls (src/bullet.js:184-220), one entry per
 * accepted update.  Entries are stored in the order the device resolves them in
 * (see BB_CFG_ORDERED_CHANGES); the reference's arrival order is recovered without a
 * sort by walking verdict[] and following the slots.  A batch is limited to
 * 2^29-2 updates. */
#define BB_NO_SLOT 0x1FFFFFFFu
#define BB_SLOT_ECHO 0x1FFFFFFEu /* BB_CFG_COMPACT_CHANGES: accepted, the entry is the update itself */
#define BB_VERDICT_CODE(v) ((uint32_t)(v) >> 29)
#define BB_VERDICT_SLOT(v) ((uint32_t)(v) & BB_NO_SLOT)
typedef struct bb_changes {
  uint64_t cap;        /* capacity of idx/head/clk/val in entries (n always suffices) */
  uint32_t* verdict;   /* [n] */
  uint64_t* n_changes; /* [1] */
  uint32_t* idx;       /* [cap] index of the update in the batch */
  bb_head* head;       /* [cap] stored value header + key order of the stored clock */
  uint32_t* clk;       /* [cap][BB_MAX_PEERS] stored clock == decision.vectorClock */
  uint64_t* val;       /* [cap][BB_MAX_FIELDS] stored value == decision.value */
} bb_changes;

/* ---- lifecycle ---------------------------------------------------------- */
int bb_abi_version(void);
/* new BulletCRT(bullet) + empty store/meta (src/bullet.js:28-31,62-64) */
int bb_create(const bb_config* cfg, bb_ctx** out);
int bb_destroy(bb_ctx* ctx);
const char* bb_last_error(const bb_ctx* ctx); /* ctx may be NULL: last create error */

/* ---- table import / export (meta.json + store.json shaped state, and the
 *      read side of Bullet._getData, src/bullet.js:115-129) ------------------ */
int bb_table_load(bb_ctx* ctx, uint64_t n, const uint64_t* path_id, const bb_row* rows);
/* materialise != 0 reproduces _getData's side effect (falsy/missing -> {}). */
int bb_table_read(bb_ctx* ctx, uint64_t n, const uint64_t* path_id, bb_row* rows_out,
                  int materialise);
int bb_table_clear(bb_ctx* ctx);

/* Pre-size the device scratch for batches of up to max_batch updates so that no
 * merge call allocates (host_entry != 0: also the device mirrors bb_merge_batch
 * copies through).  Optional: buffers otherwise grow on first use. */
int bb_reserve(bb_ctx* ctx, uint64_t max_batch, int host_entry);

/* ---- the merge: n x [BulletCRT.handleUpdate (crt:329-385) -> resolve (164-279)
 *      -> Bullet._applyUpdate (src/bullet.js:184-220)] in arrival order per path.
 *      Host buffers; H2D / D2H copies are part of the call. ------------------- */
int bb_merge_batch(bb_ctx* ctx, const bb_batch* in, bb_changes* out);
/* Same, all pointers are device pointers on ctx's device (head, clk, val, and the out arrays 16-byte aligned, as
 * cudaMalloc gives them), work is enqueued on `stream` (a cudaStream_t; 0 = ctx's own stream - pass
 * cudaStreamLegacy (0x1) to name the legacy default stream) and NOT synchronised. */
int bb_merge_batch_dev(bb_ctx* ctx, const bb_batch* in, bb_changes* out, void* stream);
/* Synchronise `stream` (0 = ctx's own) and return the deferred status of the
 * *_dev calls enqueued since the last bb_sync: BB_ERR_CAPACITY if a batch held a
 * path id >= capacity or a change buffer was too small.  A batch with a bad path id is rejected WHOLE (table
 * unchanged, its verdicts and change count undefined); batches enqueued after it are merged normally, and
 * bb_last_error names the first rejected batch by its ordinal since the previous bb_sync (1 = first). */
int bb_sync(bb_ctx* ctx, void* stream);

/* ---- indices and queries: BulletQuery (src/bullet-query.js) --------------------
 * One bb_ctx holds the children of ONE base path (`users/<id>` rows), so an index
 * is named by its field slot alone.  An index is the reference's
 * Map<String(value), Set<path>> seen as the set of (node, key) pairs it contains:
 * `key` is the 64-bit canonical form of String(value), so that key equality ==
 * bucket equality (25 == "25" is settled by the host, which passes the canonical
 * number; -0 -> 0; every NaN -> one key):
 *     number x   -> bits(x)                  (after -0 -> +0, NaN -> BB_KEY_NAN)
 *     string id  -> BB_KEY_STR  | id         (dictionary strings are non-numeric)
 *     boolean b  -> BB_KEY_BOOL | b
 * Because the post-write hook never removes the pre-update value's entry
 * (query:151-167), a node can sit in any number of buckets; the device keeps one
 * entry per node in a dense column (8 bytes per row, what range/equals stream) and
 * the rest in an open-addressing overflow set of `extra_capacity` slots.
 * Results: node ids; first the matches of the dense column (in no particular order; the
 * whole column ascending with BB_CFG_ORDERED_CHANGES), then the matches of the overflow set.  A node with entries in two matching
 * buckets appears twice, as in the reference (query:237-258).  The reference's
 * (Map order, Set order) result order is not reproduced: compare as multisets. */
#define BB_KEY_NAN 0x7FF8000000000000ull
#define BB_KEY_STR 0xFFF9000000000000ull
#define BB_KEY_BOOL 0xFFFA000000000000ull
#define BB_KEY_NONE 0xFFFFFFFFFFFFFFFFull

/* BulletQuery.index(path, field) + _buildIndex (query:30-73): entries for every
 * stored record that has the field with a value other than null (0, "" and false
 * ARE indexed here, unlike in the hook).  From then on every merge call also runs
 * _updateIndices (query:139-176) after each update, accepted or not; the ctx must
 * have been created with BB_CFG_POST_GETDATA (the hook's _getData re-read).
 * Creating an index that exists is a no-op (query:33-35).  There is no drop, as in
 * the reference.  BB_ERR_CAPACITY is reported (by the merge call or bb_sync) when
 * the overflow set is full. */
int bb_index_create(bb_ctx* ctx, uint32_t field, uint64_t extra_capacity);
/* Several indices at once (bit f of field_mask = field slot f): ONE pass over the table builds all of them. */
int bb_index_create_fields(bb_ctx* ctx, uint32_t field_mask, uint64_t extra_capacity);

/* One side of range(): ToNumber(bound) for numeric keys (NaN never matches), and,
 * when the bound is a JS string, its place in the dictionary order for string
 * keys: `rank` = number of dictionary strings < bound for a lower bound, number of
 * dictionary strings <= bound for an upper bound; BB_BOUND_TRUE/FALSE/NAN say
 * whether the keys "true", "false", "NaN" pass this side. */
#define BB_BOUND_IS_STRING 1u
#define BB_BOUND_TRUE 2u
#define BB_BOUND_FALSE 4u
#define BB_BOUND_NAN 8u
typedef struct bb_bound {
  double num;
  uint64_t rank;
  uint32_t flags;
  uint32_t reserved;
} bb_bound;

typedef struct bb_hits {
  uint64_t cap;        /* capacity of `node` in entries */
  uint32_t* node;      /* [cap] */
  uint64_t* n_dense;   /* [1] matches from the dense column */
  uint64_t* n_extra;   /* [1] matches from the overflow set, stored after them */
} bb_hits;

/* equals (query:186-210) / count (query:293-313) / range (query:221-261, both
 * ends inclusive, JS relational semantics).  BB_ERR_STATE if the index does not
 * exist (the reference would create it; the host shim calls bb_index_create
 * first).  Host buffers; the call synchronises.  count needs no node buffer. */
int bb_query_equals(bb_ctx* ctx, uint32_t field, uint64_t key, bb_hits* out);
int bb_query_count(bb_ctx* ctx, uint32_t field, uint64_t key, uint64_t* count);
int bb_query_range(bb_ctx* ctx, uint32_t field, const bb_bound* lo, const bb_bound* hi, bb_hits* out);
/* Same, device buffers (out->node, n_dense, n_extra on ctx's device), enqueued on
 * `stream`, not synchronised; too-small buffers are reported by bb_sync. */
int bb_query_equals_dev(bb_ctx* ctx, uint32_t field, uint64_t key, bb_hits* out, void* stream);
int bb_query_range_dev(bb_ctx* ctx, uint32_t field, const bb_bound* lo, const bb_bound* hi, bb_hits* out,
                       void* stream);
/* Entries currently held by the index (dense + overflow); synchronises. */
int bb_index_stats(bb_ctx* ctx, uint32_t field, uint64_t* n_dense, uint64_t* n_extra);

/* ---- sharding: the table is split over `world` ranks.  With the default sharding function path id p
 *      lives on rank p % world as local row p / world; bb_router_set_sharding(r, key_bits) switches a router to
 *      hashed sharding: s = mix(p) is a splitmix64-style finaliser restricted to key_bits bits (a bijection of
 *      [0, 2^key_bits), so a dense local row index still exists), owner = s % world, local row = s / world, and a
 *      shard holds ceil(2^key_bits / world) rows.  Strided or clustered ids then spread evenly.  bb_route_pack_dev is the send side of the update
 *      routing: a stable partition of a device-resident batch by owner rank (arrival order kept
 *      inside every destination), path ids rewritten to local rows, counts[r] = updates for rank
 *      r (device, [world]).  The caller exchanges counts and the four packed arrays with an
 *      all-to-all (NCCL over NVLink in bullet_js_b200/shard.py) and hands what it received,
 *      concatenated in source-rank order, to bb_merge_batch_dev.  world <= 16. */
int bb_route_pack_dev(bb_ctx* ctx, uint32_t world, const bb_batch* in, bb_batch* out, uint64_t* counts,
                      void* stream);

/* A router owns the exchange: two receive slots, its own streams (routing batch i+1 overlaps merging
 * batch i) and an NCCL communicator used for bootstrap and as the fallback transport (libnccl.so.2 is
 * opened at run time; a process that never creates a router does not need it).  When the ranks' slots can be
 * mapped into each other (cudaIpc: same box, NVLink / NVSwitch) the exchange is ONE kernel per rank that
 * partitions the batch in shared memory and stores every update straight into its owner's slot
 * (cp.async.bulk to peer memory); counts and completion travel through peer-mapped epoch flags.
 * Lifecycle per batch:
 *   bb_router_route_dev(r, batch, slot, &n)   collective: count by owner, exchange the counts, pack +
 *                                             store into receive slot `slot` everywhere.  Peer-store
 *                                             path: returns without waiting, *n = 2^64 - 1 (not known
 *                                             yet).  NCCL path (BB_ROUTER_NO_P2P): one host round trip,
 *                                             grouped ncclSend/ncclRecv, *n = rows received
 *   bb_router_acquire(r, slot, stream, &b)    `stream` waits for the slot; b = the received batch, in
 *                                             (source rank, arrival index) order; BB_ERR_CAPACITY if
 *                                             the slot was too small (then nothing was exchanged)
 *   bb_merge_batch_dev(ctx, &b, out, stream)  merge it into this rank's shard
 *   bb_router_release(r, slot, stream)        the slot may be overwritten once `stream` gets here
 * Every rank must call route in the same order.  The batch passed to route must stay unchanged until
 * the slot has been acquired. */
typedef struct bb_router bb_router;
#define BB_NCCL_ID_BYTES 128
int bb_router_unique_id(char id[BB_NCCL_ID_BYTES]); /* rank 0; hand the bytes to every rank */
int bb_router_create(int32_t device, uint32_t world, uint32_t rank, const char id[BB_NCCL_ID_BYTES],
                     uint64_t max_batch, uint64_t recv_capacity, bb_router** out);
int bb_router_destroy(bb_router* r);
/* every rank must call it with the same key_bits (0 = id % world) before the first route; ids must be < 2^key_bits */
int bb_router_set_sharding(bb_router* r, uint32_t key_bits);
const char* bb_router_last_error(const bb_router* r);
/* `in`: device batch; `in_stream`: stream the batch was produced on (0 = already complete). */
int bb_router_route_dev(bb_router* r, const bb_batch* in, uint32_t slot, uint64_t* n_recv, void* in_stream);
int bb_router_acquire(bb_router* r, uint32_t slot, void* stream, bb_batch* received);
int bb_router_release(bb_router* r, uint32_t slot, void* stream);
/* Sync producer side: BulletNetworkSync._collectFullSyncData(since) (src/bullet-network-sync.js:592-664).
 * bb_epoch: ordinal of the most recent merge call on this ctx (1, 2, ...; 0 before the first).
 * bb_sync_collect: one pass over the table selects every stored path (kind != none) except those the reference skips -
 * "since > 0 && meta.lastModified && meta.lastModified < since" (:602, :633), with lastModified = the ordinal of the
 * call that last wrote the row; 0 = never written by a merge (loaded rows: no lastModified, never skipped) - and
 * copies only the selected rows to the host: path ids (ascending within a warp's 32 rows, warps in any order), rows in
 * the bb_row format, ordinals.  *n_out = rows selected; BB_ERR_CAPACITY if that exceeds cap (then nothing is copied).
 * The entry list itself (leaf walk, wire shape :606-612) is the host's formatting of these rows
 * (bullet_js_b200/persist.py).  since_epoch > 0 needs BB_CFG_TRACK_MODIFIED.
 * The reference looks meta up at the LEAF path (:599, :628), so the leaves of a record that was written as a whole
 * have no lastModified and `since` never filters them: by default a record row (kind object) is always selected and
 * only primitive rows are filtered - exactly the reference's entry list.  BB_COLLECT_FILTER_RECORDS filters record rows
 * by their own lastModified as well (what a delta sync between two B200 peers wants). */
#define BB_COLLECT_FILTER_RECORDS 1u
uint64_t bb_epoch(const bb_ctx* ctx);
int bb_sync_collect(bb_ctx* ctx, uint64_t since_epoch, uint32_t flags, uint64_t cap, uint64_t* path_id_out,
                    bb_row* rows_out, uint32_t* epoch_out, uint64_t* n_out);
/* Host entry of the sharded path (the batched ingress of src/bullet-network-sync.js:551-569 for a table that spans the
 * router's ranks).  Collective: every rank passes its own batch in HOST memory (pinned recommended) and the same
 * `chunks` (1..16, 0 = 4).  Each rank's batch is cut into `chunks` pieces in arrival order; piece j of every rank is
 * copied in, exchanged (fused pack + all-to-all over NVLink) and merged by the owning shards while piece j+1 is on its
 * way in and piece j-1's results are on their way out.  A shard replays what it receives piece by piece, inside a
 * piece in (source rank, arrival index) order - the same as one peer replaying rank 0's piece 0, rank 1's piece 0, ...,
 * rank 0's piece 1, ...  Results stay sharded: `out` (host) receives verdict[] for the *n_received updates this shard
 * received, in that replay order, and their change entries (idx = position in the replay order);
 * recv_counts[j * world + q] (optional) = updates of rank q's piece j this shard received, which is what maps a
 * verdict back to (source rank, arrival index).  ctx must not have BB_CFG_ORDERED_CHANGES; BB_CFG_COMPACT_CHANGES
 * works as usual (BB_SLOT_ECHO refers to the received update).  Synchronous.  The call is a sequence of collectives:
 * if it fails on one rank (too small a buffer, a CUDA error) the ranks are no longer in step - destroy the router and
 * create a new one on every rank before routing again. */
int bb_router_merge_batch(bb_router* r, bb_ctx* ctx, const bb_batch* in, bb_changes* out, uint32_t chunks,
                          uint64_t* n_received, uint64_t* recv_counts);
/* Sharded queries (src/bullet-query.js:186-210, 221-261 over a table that spans the router's ranks).  Collective:
 * every rank calls with the same field and predicate and its own shard's ctx.  Each rank scans its shard
 * (k_index_scan), the counts travel through the peer-mapped control blocks, and one kernel per rank stores its u32
 * LOCAL hit ids straight into every rank's result buffer over NVLink at the offset the counts give it: on return
 * out->node (device memory owned by the router, valid until the next query) holds rank 0's hits, then rank 1's, ...;
 * out->offset[q] .. offset[q+1] is rank q's run.  A hit h of rank q is the node with scrambled id h * world + q
 * (path id = that, or its shard_mix inverse with hashed sharding).  The call synchronises `stream`.
 * bb_router_query_reserve (collective, once): max_total_hits = capacity of the gathered result on every rank (a
 * query that matches more returns BB_ERR_CAPACITY on every rank).  Needs the peers' memory mapped (same box); W = 1
 * works too. */
typedef struct bb_gathered_hits {
  const uint32_t* node; /* device */
  uint64_t offset[17];
  uint64_t total;
} bb_gathered_hits;
int bb_router_query_reserve(bb_router* r, uint64_t max_total_hits);
int bb_router_query_range(bb_router* r, bb_ctx* ctx, uint32_t field, const bb_bound* lo, const bb_bound* hi,
                          bb_gathered_hits* out, void* stream);
int bb_router_query_equals(bb_router* r, bb_ctx* ctx, uint32_t field, uint64_t key, bb_gathered_hits* out, void* stream);
/* convenience: copy n gathered ids starting at `first` to host memory (synchronous) */
int bb_router_query_fetch(bb_router* r, uint64_t first, uint64_t n, uint32_t* host_out);
/* Telemetry of the most recent route, ms: device time of [count by owner, counts exchange (+ wait for the
 * exchange stream's turn), pack + exchange, completion barrier], then host time of the call (twice). */
int bb_router_last_ms(bb_router* r, double out[6]);
/* bytes this rank has sent to other ranks, kernels it has launched */
uint64_t bb_router_sent_bytes(const bb_router* r);
uint64_t bb_router_launch_count(const bb_router* r);

/* ---- telemetry ---------------------------------------------------------- */
/* Kernels launched by this ctx since creation (for bench.py's gpu_launches). */
uint64_t bb_launch_count(const bb_ctx* ctx);
/* on != 0: also record the event between the front end and the merge kernels of every later merge call, so that
 * the "sort" and "merge" phases below can be told apart.  Off by default: that event sits between two launches the
 * library chains with programmatic dependent launch, and an event in between serialises them. */
int bb_phase_events(bb_ctx* ctx, int on);
/* Device time of the named phase of the most recent *_dev / host call, in ms,
 * from CUDA events on the launching stream; -1 if unknown. Synchronises. */
double bb_last_phase_ms(bb_ctx* ctx, const char* phase);
/* Same for the merge call issued `calls_ago` calls before the most recent one
 * (0 = most recent; the last 64 calls are kept).  Phases: "h2d", "sort", "merge",
 * "d2h", "device" (sort+merge), "total"; the query calls record "scan".  "sort" and "merge" need
 * bb_phase_events(ctx, 1) on the default (unsorted) pipeline. */
double bb_phase_ms(bb_ctx* ctx, const char* phase, uint32_t calls_ago);

#ifdef __cplusplus
}
#endif
#endif /* BULLET_B200_H */
