/*
 * bullet_b200_napi.c - the N-API addon js/bullet-b200.js loads as `native`: a thin pass-through of typed arrays to
 * the C ABI of include/bullet_b200.h.  Everything host-side (interning, dictionary, packing, decoding, the
 * reference's result shapes) is JavaScript (js/pack.js, js/bullet-b200.js).
 *
 *   native.create({ capacity, nFields, localPeer, flags, rankObject, rankTrue, rankFalse, rankNaN, device }) -> ctx
 *   native.mergeBatch(ctx, n, pathId, head, clk, val, out) -> number of change entries
 *       pathId Uint32Array(2n) | head Uint32Array(4n) | clk Uint32Array(8n) | val Uint32Array(8n)  == bb_batch
 *       out { verdict Uint32Array(n), idx Uint32Array(n), head Uint32Array(4n), clk Uint32Array(8n), val Uint32Array(8n) }
 *   native.tableRead(ctx, pathId) -> Uint32Array(32)      one 128-byte bb_row
 *   native.indexCreate(ctx, field) | queryEquals(ctx, field, keyLo, keyHi) -> Uint32Array | queryCount(...) -> number |
 *   queryRange(ctx, field, loNum, loRank, loFlags, hiNum, hiRank, hiFlags) -> Uint32Array   (keys / bounds: js/pack.js)
 *   native.destroy(ctx)
 *
 * Build (node-gyp or by hand):
 *   cc -shared -fPIC native/bullet_b200_napi.c -I include -I "$(node -p 'process.execPath')/../../include/node" \
 *      -L bullet_js_b200/csrc -lbulletb200 -Wl,-rpath,'$ORIGIN/../bullet_js_b200/csrc' -o build/bullet_b200.node
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no node, no node_api.h).  The same surface is implemented over the same
 * C structs in tests/js_bridge.py, which is what tests/test_js_shim.py runs the JavaScript against.
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>

#include "bullet_b200.h"

#define NAPI_OK(call)                                                          \
  do {                                                                         \
    if ((call) != napi_ok) {                                                   \
      napi_throw_error(env, NULL, "bullet_b200: N-API call failed: " #call);   \
      return NULL;                                                             \
    }                                                                          \
  } while (0)

static napi_value fail(napi_env env, bb_ctx* ctx, int rc) {
  char msg[256];
  const char* what = bb_last_error(ctx);
  snprintf(msg, sizeof msg, "bullet_b200: error %d: %s", rc, what ? what : "");
  napi_throw_error(env, NULL, msg);
  return NULL;
}

static int get_u32_prop(napi_env env, napi_value obj, const char* name, uint32_t* out) {
  napi_value v;
  bool has = false;
  if (napi_has_named_property(env, obj, name, &has) != napi_ok || !has) return 0;
  if (napi_get_named_property(env, obj, name, &v) != napi_ok) return 0;
  double d = 0;
  if (napi_get_value_double(env, v, &d) != napi_ok) return 0;
  *out = (uint32_t)d;
  return 1;
}

/* a typed array argument -> its data pointer (any element type; byte length returned) */
static void* typed_data(napi_env env, napi_value v, size_t* bytes) {
  napi_typedarray_type type;
  size_t length = 0, offset = 0;
  void* data = NULL;
  napi_value buffer;
  if (napi_get_typedarray_info(env, v, &type, &length, &data, &buffer, &offset) != napi_ok) return NULL;
  static const size_t width[] = {1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8};
  *bytes = length * width[type];
  return data;
}

static void destroy_ctx(napi_env env, void* data, void* hint) {
  (void)env;
  (void)hint;
  if (data) bb_destroy((bb_ctx*)data);
}

static napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  bb_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = BB_ABI_VERSION;
  uint32_t v = 0;
  double capacity = 0;
  napi_value cap;
  NAPI_OK(napi_get_named_property(env, argv[0], "capacity", &cap));
  NAPI_OK(napi_get_value_double(env, cap, &capacity));
  cfg.capacity = (uint64_t)capacity;
  cfg.n_fields = get_u32_prop(env, argv[0], "nFields", &v) ? v : BB_MAX_FIELDS;
  cfg.local_peer = get_u32_prop(env, argv[0], "localPeer", &v) ? v : 0;
  cfg.flags = get_u32_prop(env, argv[0], "flags", &v) ? v : 0;
  cfg.device = get_u32_prop(env, argv[0], "device", &v) ? (int32_t)v : 0;
  cfg.rank_object = get_u32_prop(env, argv[0], "rankObject", &v) ? v : 0;
  cfg.rank_true = get_u32_prop(env, argv[0], "rankTrue", &v) ? v : 0;
  cfg.rank_false = get_u32_prop(env, argv[0], "rankFalse", &v) ? v : 0;
  cfg.rank_nan = get_u32_prop(env, argv[0], "rankNaN", &v) ? v : 0;
  bb_ctx* ctx = NULL;
  const int rc = bb_create(&cfg, &ctx);
  if (rc != BB_OK) return fail(env, NULL, rc);
  napi_value ext;
  NAPI_OK(napi_create_external(env, ctx, destroy_ctx, NULL, &ext));
  return ext;
}

static napi_value MergeBatch(napi_env env, napi_callback_info info) {
  size_t argc = 7;
  napi_value a[7];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  bb_ctx* ctx = NULL;
  NAPI_OK(napi_get_value_external(env, a[0], (void**)&ctx));
  uint32_t n = 0;
  NAPI_OK(napi_get_value_uint32(env, a[1], &n));
  size_t bytes[4];
  void* in[4];
  for (int k = 0; k < 4; ++k) {
    in[k] = typed_data(env, a[2 + k], &bytes[k]);
    if (n && !in[k]) return fail(env, ctx, BB_ERR_ARG);
  }
  if (bytes[0] < 8u * n || bytes[1] < 16u * n || bytes[2] < 32u * n || bytes[3] < 32u * n) return fail(env, ctx, BB_ERR_ARG);
  static const char* names[5] = {"verdict", "idx", "head", "clk", "val"};
  static const size_t need[5] = {4, 4, 16, 32, 32};
  void* out[5];
  for (int k = 0; k < 5; ++k) {
    napi_value v;
    size_t b = 0;
    NAPI_OK(napi_get_named_property(env, a[6], names[k], &v));
    out[k] = typed_data(env, v, &b);
    if (n && (!out[k] || b < need[k] * n)) return fail(env, ctx, BB_ERR_ARG);
  }
  uint64_t n_changes = 0;
  bb_batch batch = {n, (const uint64_t*)in[0], (const bb_head*)in[1], (const uint32_t*)in[2], (const uint64_t*)in[3]};
  bb_changes changes = {n, (uint32_t*)out[0], &n_changes, (uint32_t*)out[1], (bb_head*)out[2], (uint32_t*)out[3], (uint64_t*)out[4]};
  const int rc = bb_merge_batch(ctx, &batch, &changes);
  if (rc != BB_OK) return fail(env, ctx, rc);
  napi_value r;
  NAPI_OK(napi_create_double(env, (double)n_changes, &r));
  return r;
}

static napi_value TableRead(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value a[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  bb_ctx* ctx = NULL;
  NAPI_OK(napi_get_value_external(env, a[0], (void**)&ctx));
  double id = 0;
  NAPI_OK(napi_get_value_double(env, a[1], &id));
  const uint64_t path_id = (uint64_t)id;
  void* data = NULL;
  napi_value buffer, array;
  NAPI_OK(napi_create_arraybuffer(env, sizeof(bb_row), &data, &buffer));
  const int rc = bb_table_read(ctx, 1, &path_id, (bb_row*)data, 0);
  if (rc != BB_OK) return fail(env, ctx, rc);
  NAPI_OK(napi_create_typedarray(env, napi_uint32_array, sizeof(bb_row) / 4, buffer, 0, &array));
  return array;
}

/* ---- index + queries: bb_index_create, bb_query_equals / count / range ------------------------------- */
static int args_ctx_numbers(napi_env env, napi_callback_info info, size_t want, bb_ctx** ctx, double* num) {
  size_t argc = 9;
  napi_value a[9];
  if (napi_get_cb_info(env, info, &argc, a, NULL, NULL) != napi_ok || argc < want + 1) return 0;
  if (napi_get_value_external(env, a[0], (void**)ctx) != napi_ok) return 0;
  for (size_t k = 0; k < want; ++k)
    if (napi_get_value_double(env, a[1 + k], &num[k]) != napi_ok) return 0;
  return 1;
}

static napi_value hits_to_array(napi_env env, bb_ctx* ctx, uint32_t* node, uint64_t n) {
  void* data = NULL;
  napi_value buffer, array;
  if (napi_create_arraybuffer(env, (size_t)n * 4, &data, &buffer) != napi_ok) {
    free(node);
    return fail(env, ctx, BB_ERR_ARG);
  }
  memcpy(data, node, (size_t)n * 4);
  free(node);
  NAPI_OK(napi_create_typedarray(env, napi_uint32_array, (size_t)n, buffer, 0, &array));
  return array;
}

static napi_value IndexCreate(napi_env env, napi_callback_info info) {
  bb_ctx* ctx = NULL;
  double v[1];
  if (!args_ctx_numbers(env, info, 1, &ctx, v)) return fail(env, NULL, BB_ERR_ARG);
  /* room for the stale entries the reference's hook leaves behind (query:151-167): as the Python host, 4 per row */
  const int rc = bb_index_create(ctx, (uint32_t)v[0], 1ull << 22);
  if (rc != BB_OK) return fail(env, ctx, rc);
  napi_value u;
  NAPI_OK(napi_get_undefined(env, &u));
  return u;
}

static napi_value QueryCount(napi_env env, napi_callback_info info) {
  bb_ctx* ctx = NULL;
  double v[3];
  if (!args_ctx_numbers(env, info, 3, &ctx, v)) return fail(env, NULL, BB_ERR_ARG);
  uint64_t count = 0;
  const uint64_t key = (uint64_t)(uint32_t)v[1] | ((uint64_t)(uint32_t)v[2] << 32);
  const int rc = bb_query_count(ctx, (uint32_t)v[0], key, &count);
  if (rc != BB_OK) return fail(env, ctx, rc);
  napi_value r;
  NAPI_OK(napi_create_double(env, (double)count, &r));
  return r;
}

/* two passes: count the hits, then fetch them into a buffer of that size */
static napi_value QueryEquals(napi_env env, napi_callback_info info) {
  bb_ctx* ctx = NULL;
  double v[3];
  if (!args_ctx_numbers(env, info, 3, &ctx, v)) return fail(env, NULL, BB_ERR_ARG);
  const uint64_t key = (uint64_t)(uint32_t)v[1] | ((uint64_t)(uint32_t)v[2] << 32);
  uint64_t count = 0, n_dense = 0, n_extra = 0;
  int rc = bb_query_count(ctx, (uint32_t)v[0], key, &count);
  if (rc != BB_OK) return fail(env, ctx, rc);
  uint32_t* node = (uint32_t*)malloc((size_t)(count ? count : 1) * 4);
  bb_hits hits = {count ? count : 1, node, &n_dense, &n_extra};
  rc = bb_query_equals(ctx, (uint32_t)v[0], key, &hits);
  if (rc != BB_OK) {
    free(node);
    return fail(env, ctx, rc);
  }
  return hits_to_array(env, ctx, node, n_dense + n_extra);
}

/* range(field, lo{num, rank, flags}, hi{num, rank, flags}); the hit buffer is sized by the table (bb_index_stats) */
static napi_value QueryRange(napi_env env, napi_callback_info info) {
  bb_ctx* ctx = NULL;
  double v[7];
  if (!args_ctx_numbers(env, info, 7, &ctx, v)) return fail(env, NULL, BB_ERR_ARG);
  const bb_bound lo = {v[1], (uint64_t)v[2], (uint32_t)v[3], 0}, hi = {v[4], (uint64_t)v[5], (uint32_t)v[6], 0};
  uint64_t dense = 0, extra = 0, n_dense = 0, n_extra = 0;
  int rc = bb_index_stats(ctx, (uint32_t)v[0], &dense, &extra);
  if (rc != BB_OK) return fail(env, ctx, rc);
  const uint64_t cap = dense + extra + 1;
  uint32_t* node = (uint32_t*)malloc((size_t)cap * 4);
  bb_hits hits = {cap, node, &n_dense, &n_extra};
  rc = bb_query_range(ctx, (uint32_t)v[0], &lo, &hi, &hits);
  if (rc != BB_OK) {
    free(node);
    return fail(env, ctx, rc);
  }
  return hits_to_array(env, ctx, node, n_dense + n_extra);
}

static napi_value Destroy(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value a[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  (void)a; /* the context is released by the external's finalizer */
  napi_value u;
  NAPI_OK(napi_get_undefined(env, &u));
  return u;
}

NAPI_MODULE_INIT() {
  static const struct {
    const char* name;
    napi_callback fn;
  } fns[] = {{"create", Create},           {"mergeBatch", MergeBatch},   {"tableRead", TableRead},
             {"indexCreate", IndexCreate}, {"queryEquals", QueryEquals}, {"queryCount", QueryCount},
             {"queryRange", QueryRange},   {"destroy", Destroy}};
  for (size_t i = 0; i < sizeof fns / sizeof fns[0]; ++i) {
    napi_value f;
    if (napi_create_function(env, fns[i].name, NAPI_AUTO_LENGTH, fns[i].fn, NULL, &f) != napi_ok) return NULL;
    if (napi_set_named_property(env, exports, fns[i].name, f) != napi_ok) return NULL;
  }
  return exports;
}
