// bullet_b200.cu - the C ABI of include/bullet_b200.h over the sm_100a kernels.
// One bb_ctx == one GPU-resident shard of the graph table (rows of 128 bytes,
// row index == interned path id) plus the scratch the pipeline needs.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "bb_kernels.cuh"

static_assert(sizeof(bb_row) == 128, "table rows are one 128-byte line");
static_assert(sizeof(bb_head) == 16, "heads are one 16-byte vector");

namespace {

thread_local std::string g_create_error;

enum { EV_H2D0, EV_START, EV_SORT, EV_MERGE, EV_D2H, EV_Q0, EV_Q1, EV_COUNT };
constexpr int EV_RING = 64;  // merge calls whose phase timings can still be queried
constexpr int MAX_CHUNKS = 8;          // a host call is pipelined as up to this many chunks
constexpr uint64_t MIN_CHUNK = 1 << 16;  // updates

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 1024;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct bb_ctx {
  bb_config cfg{};
  cudaStream_t stream = nullptr;
  uint4* table = nullptr;
  uint64_t seq = 0;  // updates (and materialising reads) seen so far
  int key_bits = 0;
  uint64_t launches = 0;
  std::string err;
  cudaEvent_t ev[EV_RING][EV_COUNT]{};
  bool ev_valid[EV_RING][EV_COUNT]{};
  uint64_t calls = 0;  // merge calls started; ring slot = (calls - 1) % EV_RING
  // pipeline scratch
  DevBuf<uint64_t> items_a, items_b;
  DevBuf<uint32_t> zero;  // zeroed per call: [digit histograms | tickets | sort tile states | merge tile states]
  DevBuf<uint32_t> st_idx;
  DevBuf<uint4> st_ent;
  uint32_t* d_err = nullptr;    // bit0: path id out of range, bit1: change buffer too small
  uint32_t* h_err = nullptr;    // pinned
  // device mirrors of the host-call buffers
  DevBuf<uint64_t> io_path;
  DevBuf<uint4> io_head, io_clk, io_val, io_out_head, io_out_clk, io_out_val, io_rows;
  DevBuf<uint32_t> io_verdict, io_out_idx;
  uint32_t* cs_cnt = nullptr;  // counting sort: per-row update counts (all zero between calls)
  uint32_t* cs_off = nullptr;  // and their exclusive scan, [capacity + 1] (both padded to 1024)
  DevBuf<uint2> cs_long;       // segments longer than CS_SHORT, queued for k_cs_fix_long
  DevBuf<uint32_t> cs_tile;    // per-4096-row sums of cs_cnt
  uint64_t* d_nchanges = nullptr;
  uint64_t* d_chg_base = nullptr;
  uint64_t* h_nchanges = nullptr;  // pinned [MAX_CHUNKS]: running total after each chunk of a host call
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_in[MAX_CHUNKS]{}, ev_done[MAX_CHUNKS]{}, ev_cnt[MAX_CHUNKS]{};
  // indices (bb_index.cuh): dense key column + overflow set per indexed field
  struct IndexDev {
    bool live = false;
    uint64_t* pcol = nullptr;
    uint64_t* xkey = nullptr;
    uint32_t* xnode = nullptr;
    uint64_t xslots = 0;
  } index[BB_MAX_FIELDS];
  uint32_t index_mask = 0;
  uint32_t* d_xused = nullptr;             // [BB_MAX_FIELDS]
  unsigned long long* d_counters = nullptr;  // [2] dense / overflow matches of the running query
  unsigned long long* h_counters = nullptr;  // pinned [2]
  DevBuf<uint32_t> route_tiles;            // [tiles][world] of bb_route_pack_dev
  DevBuf<uint32_t> scan_zero;              // [ticket x 2 | tile states of both scans]
  DevBuf<uint32_t> io_hits;
};

namespace {

int fail(bb_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) {
      c->err += ": ";
      c->err += cudaGetErrorString(e);
    }
  }
  return code;
}

#define BB_CUDA(c, call)                                         \
  do {                                                           \
    cudaError_t e_ = (call);                                     \
    if (e_ != cudaSuccess) return fail((c), BB_ERR_CUDA, #call, e_); \
  } while (0)

#define BB_LAUNCH(c, kernel, grid, block, stream, ...)                  \
  do {                                                                  \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);              \
    ++(c)->launches;                                                    \
    cudaError_t e_ = cudaGetLastError();                                \
    if (e_ != cudaSuccess) return fail((c), BB_ERR_CUDA, #kernel, e_);  \
  } while (0)

inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

void begin_call(bb_ctx* c) {
  ++c->calls;
  const int slot = (int)((c->calls - 1) % EV_RING);
  for (int i = 0; i < EV_COUNT; ++i) c->ev_valid[slot][i] = false;
}

void mark(bb_ctx* c, int which, cudaStream_t s) {
  const int slot = (int)((c->calls - 1) % EV_RING);
  cudaEventRecord(c->ev[slot][which], s);
  c->ev_valid[slot][which] = true;
}

struct ZeroLayout {
  uint32_t passes, sort_tiles, merge_tiles;
  size_t hist, tickets, sort_state, merge_state, cs_state, cs_ctr, total;  // offsets in uint32_t units
};

// counting sort when the scan over the rows is cheap next to the batch, radix sort otherwise
bool use_counting_sort(const bb_ctx* c, uint64_t n) {
  if (c->cfg.flags & BB_CFG_RADIX_SORT) return false;
  return c->cfg.capacity <= 64 * (n < 4096 ? 4096 : n);
}

ZeroLayout zero_layout(const bb_ctx* c, uint64_t n) {
  using namespace bb;
  ZeroLayout z;
  z.passes = (uint32_t)((c->key_bits + 7) / 8);
  z.sort_tiles = div_up(n, SORT_TILE);
  z.merge_tiles = div_up(n, MT);
  z.hist = 0;
  z.tickets = z.hist + (size_t)MAX_PASSES * RADIX;
  z.sort_state = z.tickets + 8;
  if (use_counting_sort(c, n)) z.passes = 0;
  z.merge_state = z.sort_state + (size_t)z.passes * z.sort_tiles * RADIX;
  z.cs_state = z.merge_state + z.merge_tiles;
  z.cs_ctr = z.cs_state;
  z.total = z.cs_ctr + 2;
  return z;
}

// size every scratch buffer of the device pipeline for batches of up to n updates
int reserve_dev(bb_ctx* c, uint64_t n) {
  const ZeroLayout z = zero_layout(c, n);
  BB_CUDA(c, c->items_a.ensure(n));
  BB_CUDA(c, c->items_b.ensure(n));
  BB_CUDA(c, c->zero.ensure(z.total));
  BB_CUDA(c, c->st_idx.ensure(n));
  BB_CUDA(c, c->st_ent.ensure(5 * n));
  BB_CUDA(c, c->cs_long.ensure(n / 8 + 1));
  return BB_OK;
}

int reserve_io(bb_ctx* c, uint64_t n) {
  BB_CUDA(c, c->io_path.ensure(n));
  BB_CUDA(c, c->io_head.ensure(n));
  BB_CUDA(c, c->io_clk.ensure(2 * n));
  BB_CUDA(c, c->io_val.ensure(2 * n));
  BB_CUDA(c, c->io_verdict.ensure(n));
  BB_CUDA(c, c->io_out_idx.ensure(n));
  BB_CUDA(c, c->io_out_head.ensure(n));
  BB_CUDA(c, c->io_out_clk.ensure(2 * n));
  BB_CUDA(c, c->io_out_val.ensure(2 * n));
  return BB_OK;
}

// One batch (or one chunk of a host call: `idx_base` = arrival index of its first update,
// `append` = keep adding to *out->n_changes instead of starting a new change set).
int merge_dev(bb_ctx* c, const bb_batch* in, bb_changes* out, cudaStream_t s, uint32_t idx_base = 0,
              bool append = false) {
  using namespace bb;
  const uint64_t n = in->n;
  if (n >= BB_NO_SLOT) return fail(c, BB_ERR_ARG, "batch larger than 2^29-2 updates");
  if (!append) {
    mark(c, EV_START, s);
    BB_CUDA(c, cudaMemsetAsync(out->n_changes, 0, sizeof(uint64_t), s));
  }
  if (n == 0) {
    if (!append) {
      mark(c, EV_SORT, s);
      mark(c, EV_MERGE, s);
    }
    return BB_OK;
  }
  {
    int rc = reserve_dev(c, n);  // no-op once the scratch is large enough
    if (rc) return rc;
  }
  const ZeroLayout z = zero_layout(c, n);
  uint32_t* zp = c->zero.p;
  BB_CUDA(c, cudaMemsetAsync(zp, 0, z.total * sizeof(uint32_t), s));

  uint64_t* src = c->items_a.p;
  if (use_counting_sort(c, n)) {
    // K1': counting sort keyed by the row index (bb_kernels.cuh), arrival order restored per segment
    const uint32_t g = div_up(n, CS_THREADS);
    const uint64_t cap = c->cfg.capacity;
    const uint32_t tiles = div_up(cap + 1, CS_TILE);  // + 1: off[capacity] = the batch size
    BB_LAUNCH(c, k_cs_count, div_up(n, CS_THREADS * CS_ILP), CS_THREADS, s, in->path_id, n, cap, c->cs_cnt, c->st_idx.p, c->d_err);
    BB_LAUNCH(c, k_cs_tile_sums, tiles, CS_THREADS, s, c->cs_cnt, c->cs_tile.p);
    BB_LAUNCH(c, k_cs_offsets, tiles, CS_THREADS, s, c->cs_cnt, c->cs_tile.p, c->cs_off);
    BB_LAUNCH(c, k_cs_place, div_up(n, CS_THREADS * CS_ILP), CS_THREADS, s, in->path_id, n, cap, c->st_idx.p, c->cs_off, src, c->d_err);
    BB_LAUNCH(c, k_cs_fix, g, CS_THREADS, s, src, n, c->cs_off, c->cs_long.p, zp + z.cs_ctr, c->d_err);
    BB_LAUNCH(c, k_cs_fix_long, CS_LONG_CTAS, CS_THREADS, s, src, c->items_b.p, c->cs_long.p, zp + z.cs_ctr,
              zp + z.cs_ctr + 1);
  } else {
    // K0 + K1: stable LSD radix sort of (path id, arrival index) by path id
    BB_LAUNCH(c, k_keys_hist, z.sort_tiles, SORT_THREADS, s, in->path_id, n, c->cfg.capacity, (int)z.passes,
              c->items_a.p, zp + z.hist, c->d_err);
    BB_LAUNCH(c, k_hist_scan, z.passes, RADIX, s, zp + z.hist);
    uint64_t* dst = c->items_b.p;
    for (uint32_t pass = 0; pass < z.passes; ++pass) {
      BB_LAUNCH(c, k_sort_pass, z.sort_tiles, SORT_THREADS, s, src, dst, n, (int)(8 * pass),
                zp + z.hist + (size_t)pass * RADIX, zp + z.sort_state + (size_t)pass * z.sort_tiles * RADIX,
                zp + z.tickets + pass);
      uint64_t* t = src;
      src = dst;
      dst = t;
    }
  }
  if (!append) mark(c, EV_SORT, s);
  if (c->cfg.flags & BB_CFG_ORDERED_CHANGES)
    BB_CUDA(c, cudaMemcpyAsync(c->d_chg_base, out->n_changes, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));

  // K2: per-path sequential replay against the table + change-set compaction
  MergeArgs ma;
  ma.sorted = src;
  ma.n = n;
  ma.table = c->table;
  ma.head = reinterpret_cast<const uint4*>(in->head);
  ma.clk = reinterpret_cast<const uint4*>(in->clk);
  ma.val = reinterpret_cast<const uint4*>(in->val);
  ma.verdict = out->verdict;
  ma.n_changes = out->n_changes;
  ma.out_idx = out->idx;
  ma.out_head = reinterpret_cast<uint4*>(out->head);
  ma.out_clk = reinterpret_cast<uint4*>(out->clk);
  ma.out_val = reinterpret_cast<uint4*>(out->val);
  ma.cap = out->cap;
  ma.st_idx = c->st_idx.p;
  ma.st_ent = c->st_ent.p;
  ma.tile_state = zp + z.merge_state;
  ma.ticket = zp + z.tickets + MAX_PASSES;
  ma.num_tiles = z.merge_tiles;
  ma.seq_base = c->seq;
  ma.idx_base = idx_base;
  ma.chg_base = c->d_chg_base;
  ma.err = c->d_err;
  ma.p.rank_object = c->cfg.rank_object;
  ma.p.me = c->cfg.local_peer;
  ma.p.post_getdata = (c->cfg.flags & BB_CFG_POST_GETDATA) != 0;
  ma.ix.mask = c->index_mask;
  ma.ix.xused = c->d_xused;
  for (int f = 0; f < BB_MAX_FIELDS; ++f) {
    ma.ix.pcol[f] = c->index[f].pcol;
    ma.ix.xkey[f] = c->index[f].xkey;
    ma.ix.xnode[f] = c->index[f].xnode;
    ma.ix.xmask[f] = c->index[f].live ? (uint32_t)(c->index[f].xslots - 1) : 0u;
  }
  const bool ordered = (c->cfg.flags & BB_CFG_ORDERED_CHANGES) != 0;
  if (c->index_mask) {
    if (ordered) BB_LAUNCH(c, (k_merge_stage<true, true>), z.merge_tiles, MT, s, ma);
    else BB_LAUNCH(c, (k_merge_stage<false, true>), z.merge_tiles, MT, s, ma);
  } else {
    if (ordered) BB_LAUNCH(c, (k_merge_stage<true, false>), z.merge_tiles, MT, s, ma);
    else BB_LAUNCH(c, (k_merge_stage<false, false>), z.merge_tiles, MT, s, ma);
  }
  if (!append) mark(c, EV_MERGE, s);
  c->seq += n;
  return BB_OK;
}

// fetch + clear the deferred device error word; stream must be idle afterwards
int collect_device_error(bb_ctx* c, cudaStream_t s) {
  BB_CUDA(c, cudaMemcpyAsync(c->h_err, c->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  BB_CUDA(c, cudaMemsetAsync(c->d_err, 0, sizeof(uint32_t), s));
  BB_CUDA(c, cudaStreamSynchronize(s));
  const uint32_t e = *c->h_err;
  if (e & bb::ERR_RANGE) return fail(c, BB_ERR_CAPACITY, "path id >= capacity (batch rejected, table unchanged)");
  if (e & bb::ERR_CHANGES) return fail(c, BB_ERR_CAPACITY, "change-set buffer too small");
  if (e & bb::ERR_XFULL) return fail(c, BB_ERR_CAPACITY, "index overflow set is full (bb_index_create extra_capacity)");
  if (e & bb::ERR_HITS) return fail(c, BB_ERR_CAPACITY, "hit buffer too small");
  return BB_OK;
}

// (re)initialise the index on field f from the current table
int index_fill(bb_ctx* c, int f, cudaStream_t s) {
  bb_ctx::IndexDev& ix = c->index[f];
  const uint64_t cap2 = (c->cfg.capacity + 1) & ~1ull;
  BB_CUDA(c, cudaMemsetAsync(ix.pcol, 0xFF, cap2 * sizeof(uint64_t), s));
  BB_CUDA(c, cudaMemsetAsync(ix.xkey, 0xFF, ix.xslots * sizeof(uint64_t), s));
  BB_CUDA(c, cudaMemsetAsync(ix.xnode, 0xFF, ix.xslots * sizeof(uint32_t), s));
  BB_CUDA(c, cudaMemsetAsync(c->d_xused + f, 0, sizeof(uint32_t), s));
  BB_LAUNCH(c, bb::k_index_build, div_up(c->cfg.capacity, 256), 256, s, c->table, c->cfg.capacity, f, ix.pcol);
  return BB_OK;
}

int launch_scan(bb_ctx* c, const bb::ScanArgs& a, uint32_t tiles, bool ordered, cudaStream_t s) {
  using namespace bb;
  switch (a.p.mode * 2 + (ordered ? 1 : 0)) {
    case 0: BB_LAUNCH(c, (k_index_scan<false, 0>), tiles, SC_THREADS, s, a); break;
    case 1: BB_LAUNCH(c, (k_index_scan<true, 0>), tiles, SC_THREADS, s, a); break;
    case 2: BB_LAUNCH(c, (k_index_scan<false, 1>), tiles, SC_THREADS, s, a); break;
    case 3: BB_LAUNCH(c, (k_index_scan<true, 1>), tiles, SC_THREADS, s, a); break;
    case 4: BB_LAUNCH(c, (k_index_scan<false, 2>), tiles, SC_THREADS, s, a); break;
    default: BB_LAUNCH(c, (k_index_scan<true, 2>), tiles, SC_THREADS, s, a); break;
  }
  return BB_OK;
}

// both scans of one query; `hits` (device or null), counters land in c->d_counters
int scan_dev(bb_ctx* c, uint32_t field, const bb::Pred& pred, uint32_t* hits, uint64_t cap, cudaStream_t s) {
  using namespace bb;
  if (field >= c->cfg.n_fields || !c->index[field].live)
    return fail(c, BB_ERR_STATE, "no index on this field (bb_index_create first)");
  if (c->cfg.capacity >= (1ull << 30)) return fail(c, BB_ERR_ARG, "queries support up to 2^30 rows per shard");
  const bb_ctx::IndexDev& ix = c->index[field];
  const uint64_t n0 = (c->cfg.capacity + 1) & ~1ull, n1 = ix.xslots;
  const uint32_t t0 = div_up(n0, SC_TILE), t1 = div_up(n1, SC_TILE);
  BB_CUDA(c, c->scan_zero.ensure(2 + (size_t)t0 + t1));
  mark(c, EV_Q0, s);
  BB_CUDA(c, cudaMemsetAsync(c->scan_zero.p, 0, (2 + (size_t)t0 + t1) * sizeof(uint32_t), s));
  BB_CUDA(c, cudaMemsetAsync(c->d_counters, 0, 2 * sizeof(unsigned long long), s));
  ScanArgs a;
  a.out = hits;
  a.cap = cap;
  a.counters = c->d_counters;
  a.err = c->d_err;
  a.p = pred;
  a.keys = ix.pcol; a.nodes = nullptr; a.n = n0; a.which = 0;
  a.ticket = c->scan_zero.p; a.tile_state = c->scan_zero.p + 2; a.num_tiles = t0;
  const bool ordered = (c->cfg.flags & BB_CFG_ORDERED_CHANGES) != 0;
  {
    int rc = launch_scan(c, a, t0, ordered, s);
    if (rc) return rc;
  }
  a.keys = ix.xkey; a.nodes = ix.xnode; a.n = n1; a.which = 1;
  a.ticket = c->scan_zero.p + 1; a.tile_state = c->scan_zero.p + 2 + t0; a.num_tiles = t1;
  {
    int rc = launch_scan(c, a, t1, ordered, s);
    if (rc) return rc;
  }
  mark(c, EV_Q1, s);
  return BB_OK;
}

bb::Pred range_pred(const bb_bound* lo, const bb_bound* hi) {
  bb::Pred p{};
  p.lo = lo->num; p.hi = hi->num;
  p.lo_rank = lo->rank; p.hi_rank = hi->rank;
  p.lo_flags = lo->flags; p.hi_flags = hi->flags;
  if ((lo->flags | hi->flags) & BB_BOUND_IS_STRING) {
    p.mode = 2;  // string keys can match: the general predicate
    return p;
  }
  // numeric bounds only: x >= lo && x <= hi on the ordered images of the bits (keys hold no -0)
  p.mode = 1;
  p.eq = 1;     // an empty range: image 1 belongs to an unused NaN pattern
  p.width = 0;
  if (lo->num == lo->num && hi->num == hi->num && lo->num <= hi->num) {
    double l = lo->num == 0.0 ? 0.0 : lo->num, h = hi->num == 0.0 ? 0.0 : hi->num;  // -0 -> +0
    uint64_t lb, hb;
    memcpy(&lb, &l, 8);
    memcpy(&hb, &h, 8);
    p.eq = bb::ordered_image(lb);
    p.width = bb::ordered_image(hb) - p.eq;
  }
  return p;
}

int query_host(bb_ctx* c, uint32_t field, const bb::Pred& pred, bb_hits* out) {
  if (!out || !out->n_dense || !out->n_extra || (out->cap && !out->node)) return fail(c, BB_ERR_ARG, "null argument");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  cudaStream_t s = c->stream;
  BB_CUDA(c, c->io_hits.ensure(out->cap ? out->cap : 1));
  begin_call(c);
  int rc = scan_dev(c, field, pred, c->io_hits.p, out->cap, s);
  if (rc) return rc;
  BB_CUDA(c, cudaMemcpyAsync(c->h_counters, c->d_counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  rc = collect_device_error(c, s);  // synchronises
  if (rc) return rc;
  const uint64_t k = c->h_counters[0] + c->h_counters[1];
  if (k) BB_CUDA(c, cudaMemcpyAsync(out->node, c->io_hits.p, k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  BB_CUDA(c, cudaStreamSynchronize(s));
  *out->n_dense = c->h_counters[0];
  *out->n_extra = c->h_counters[1];
  return BB_OK;
}

int query_dev(bb_ctx* c, uint32_t field, const bb::Pred& pred, bb_hits* out, void* stream) {
  if (!out || !out->n_dense || !out->n_extra || !out->node) return fail(c, BB_ERR_ARG, "null argument");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  begin_call(c);
  int rc = scan_dev(c, field, pred, out->node, out->cap, s);
  if (rc) return rc;
  BB_CUDA(c, cudaMemcpyAsync(out->n_dense, c->d_counters, 8, cudaMemcpyDeviceToDevice, s));
  BB_CUDA(c, cudaMemcpyAsync(out->n_extra, c->d_counters + 1, 8, cudaMemcpyDeviceToDevice, s));
  return BB_OK;
}

}  // namespace

extern "C" {

int bb_abi_version(void) { return BB_ABI_VERSION; }

const char* bb_last_error(const bb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int bb_create(const bb_config* cfg, bb_ctx** out) {
  if (!cfg || !out) {
    g_create_error = "null argument";
    return BB_ERR_ARG;
  }
  *out = nullptr;
  if (cfg->abi_version != BB_ABI_VERSION || cfg->n_fields < 1 || cfg->n_fields > BB_MAX_FIELDS ||
      cfg->local_peer >= BB_MAX_PEERS || cfg->capacity == 0 || cfg->capacity >= 0xFFFFFFFFull) {
    g_create_error = "bad bb_config (abi_version / n_fields / local_peer / capacity)";
    return BB_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = std::string("no usable CUDA device (there is no CPU fallback): ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
    return BB_ERR_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10) {
    g_create_error = "device is not sm_100 (this library is built for B200 only)";
    return BB_ERR_CUDA;
  }
  bb_ctx* c = new (std::nothrow) bb_ctx();
  if (!c) {
    g_create_error = "out of host memory";
    return BB_ERR_ARG;
  }
  c->cfg = *cfg;
  int bits = 1;
  while (bits < 32 && (1ull << bits) < cfg->capacity) ++bits;
  c->key_bits = bits;
  const size_t cs_words = ((size_t)cfg->capacity + 1 + bb::CS_TILE - 1) / bb::CS_TILE * bb::CS_TILE;
  bool ok = cudaSetDevice(cfg->device) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc((void**)&c->table, cfg->capacity * sizeof(bb_row)) == cudaSuccess &&
            cudaMemsetAsync(c->table, 0, cfg->capacity * sizeof(bb_row), c->stream) == cudaSuccess &&
            cudaMalloc((void**)&c->d_err, sizeof(uint32_t)) == cudaSuccess &&
            cudaMemsetAsync(c->d_err, 0, sizeof(uint32_t), c->stream) == cudaSuccess &&
            cudaMalloc((void**)&c->cs_cnt, cs_words * sizeof(uint32_t)) == cudaSuccess &&
            cudaMalloc((void**)&c->cs_off, cs_words * sizeof(uint32_t)) == cudaSuccess &&
            cudaMemsetAsync(c->cs_cnt, 0, cs_words * sizeof(uint32_t), c->stream) == cudaSuccess &&
            cudaMemsetAsync(c->cs_off, 0, cs_words * sizeof(uint32_t), c->stream) == cudaSuccess &&
            c->cs_tile.ensure(cs_words / bb::CS_TILE) == cudaSuccess &&
            cudaMalloc((void**)&c->d_nchanges, sizeof(uint64_t)) == cudaSuccess &&
            cudaMalloc((void**)&c->d_chg_base, sizeof(uint64_t)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc((void**)&c->d_xused, BB_MAX_FIELDS * sizeof(uint32_t)) == cudaSuccess &&
            cudaMemsetAsync(c->d_xused, 0, BB_MAX_FIELDS * sizeof(uint32_t), c->stream) == cudaSuccess &&
            cudaMalloc((void**)&c->d_counters, 2 * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMallocHost((void**)&c->h_counters, 2 * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMallocHost((void**)&c->h_err, sizeof(uint32_t)) == cudaSuccess &&
            cudaMallocHost((void**)&c->h_nchanges, MAX_CHUNKS * sizeof(uint64_t)) == cudaSuccess;
  for (int i = 0; ok && i < MAX_CHUNKS; ++i)
    ok = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_cnt[i], cudaEventDisableTiming) == cudaSuccess;
  for (int r = 0; ok && r < EV_RING; ++r)
    for (int i = 0; ok && i < EV_COUNT; ++i) ok = cudaEventCreate(&c->ev[r][i]) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
  if (!ok) {
    g_create_error = std::string("CUDA allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    bb_destroy(c);
    return BB_ERR_CUDA;
  }
  *out = c;
  return BB_OK;
}

int bb_destroy(bb_ctx* c) {
  if (!c) return BB_ERR_ARG;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  c->items_a.release(); c->items_b.release(); c->zero.release(); c->st_idx.release();
  c->st_ent.release();
  c->io_path.release(); c->io_head.release(); c->io_clk.release(); c->io_val.release();
  c->io_out_head.release(); c->io_out_clk.release(); c->io_out_val.release(); c->io_rows.release();
  c->io_verdict.release(); c->io_out_idx.release();
  c->scan_zero.release(); c->io_hits.release(); c->route_tiles.release();
  for (int f = 0; f < BB_MAX_FIELDS; ++f) {
    if (c->index[f].pcol) cudaFree(c->index[f].pcol);
    if (c->index[f].xkey) cudaFree(c->index[f].xkey);
    if (c->index[f].xnode) cudaFree(c->index[f].xnode);
  }
  if (c->d_xused) cudaFree(c->d_xused);
  if (c->d_counters) cudaFree(c->d_counters);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->table) cudaFree(c->table);
  if (c->d_err) cudaFree(c->d_err);
  if (c->cs_cnt) cudaFree(c->cs_cnt);
  if (c->cs_off) cudaFree(c->cs_off);
  c->cs_long.release(); c->cs_tile.release();
  if (c->d_nchanges) cudaFree(c->d_nchanges);
  if (c->d_chg_base) cudaFree(c->d_chg_base);
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
    if (c->ev_cnt[i]) cudaEventDestroy(c->ev_cnt[i]);
  }
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  if (c->h_err) cudaFreeHost(c->h_err);
  if (c->h_nchanges) cudaFreeHost(c->h_nchanges);
  for (int r = 0; r < EV_RING; ++r)
    for (int i = 0; i < EV_COUNT; ++i)
      if (c->ev[r][i]) cudaEventDestroy(c->ev[r][i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return BB_OK;
}

int bb_table_clear(bb_ctx* c) {
  if (!c) return BB_ERR_ARG;
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  BB_CUDA(c, cudaMemsetAsync(c->table, 0, c->cfg.capacity * sizeof(bb_row), c->stream));
  for (int f = 0; f < BB_MAX_FIELDS; ++f)
    if (c->index[f].live) {
      int rc = index_fill(c, f, c->stream);
      if (rc) return rc;
    }
  BB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->seq = 0;
  return BB_OK;
}

int bb_table_load(bb_ctx* c, uint64_t n, const uint64_t* path_id, const bb_row* rows) {
  if (!c || (n && (!path_id || !rows))) return fail(c, BB_ERR_ARG, "null argument");
  if (n == 0) return BB_OK;
  if (c->index_mask) return fail(c, BB_ERR_STATE, "load the table before creating indices");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  cudaStream_t s = c->stream;
  BB_CUDA(c, c->io_path.ensure(n));
  BB_CUDA(c, c->io_rows.ensure(n * 8));
  BB_CUDA(c, cudaMemcpyAsync(c->io_path.p, path_id, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  BB_CUDA(c, cudaMemcpyAsync(c->io_rows.p, rows, n * sizeof(bb_row), cudaMemcpyHostToDevice, s));
  BB_LAUNCH(c, bb::k_table_scatter, div_up(n * 8, 256), 256, s, c->table, c->io_path.p, c->io_rows.p, n,
            c->cfg.capacity, c->d_err);
  return collect_device_error(c, s);
}

int bb_table_read(bb_ctx* c, uint64_t n, const uint64_t* path_id, bb_row* rows_out, int materialise) {
  if (!c || (n && (!path_id || !rows_out))) return fail(c, BB_ERR_ARG, "null argument");
  if (n == 0) return BB_OK;
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  cudaStream_t s = c->stream;
  BB_CUDA(c, c->io_path.ensure(n));
  BB_CUDA(c, c->io_rows.ensure(n * 8));
  BB_CUDA(c, cudaMemcpyAsync(c->io_path.p, path_id, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  BB_CUDA(c, cudaMemsetAsync(c->io_rows.p, 0, n * sizeof(bb_row), s));
  BB_LAUNCH(c, bb::k_table_gather, div_up(n, 256), 256, s, c->table, c->io_path.p, c->io_rows.p, n,
            c->cfg.capacity, materialise, c->seq, c->d_err);
  BB_CUDA(c, cudaMemcpyAsync(rows_out, c->io_rows.p, n * sizeof(bb_row), cudaMemcpyDeviceToHost, s));
  if (materialise) c->seq += n;
  return collect_device_error(c, s);
}

int bb_merge_batch_dev(bb_ctx* c, const bb_batch* in, bb_changes* out, void* stream) {
  if (!c || !in || !out) return fail(c, BB_ERR_ARG, "null argument");
  if (in->n && (!in->path_id || !in->head || !in->clk || !in->val || !out->verdict || !out->idx ||
                !out->head || !out->clk || !out->val))
    return fail(c, BB_ERR_ARG, "null buffer");
  if (!out->n_changes) return fail(c, BB_ERR_ARG, "null n_changes");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  begin_call(c);
  return merge_dev(c, in, out, stream ? (cudaStream_t)stream : c->stream);
}

int bb_reserve(bb_ctx* c, uint64_t max_batch, int host_entry) {
  if (!c || max_batch >= 0xFFFFFFFFull) return fail(c, BB_ERR_ARG, "bad argument");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  int rc = reserve_dev(c, max_batch);
  if (rc == BB_OK && host_entry) rc = reserve_io(c, max_batch);
  return rc;
}

int bb_sync(bb_ctx* c, void* stream) {
  if (!c) return BB_ERR_ARG;
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  return collect_device_error(c, stream ? (cudaStream_t)stream : c->stream);
}

// Host entry.  The batch is cut into up to MAX_CHUNKS chunks in arrival order and pipelined over
// three streams - H2D of chunk i+1, sort + merge of chunk i and D2H of chunk i-1 overlap - so both
// PCIe directions are busy at once.  Chunks are merged strictly in order on one stream, so the
// per-path arrival order (and therefore every decision) is that of the unsplit batch.
int bb_merge_batch(bb_ctx* c, const bb_batch* in, bb_changes* out) {
  if (!c || !in || !out || !out->n_changes) return fail(c, BB_ERR_ARG, "null argument");
  const uint64_t n = in->n;
  if (n && (!in->path_id || !in->head || !in->clk || !in->val || !out->verdict))
    return fail(c, BB_ERR_ARG, "null buffer");
  if (n >= BB_NO_SLOT) return fail(c, BB_ERR_ARG, "batch larger than 2^29-2 updates");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  cudaStream_t s = c->stream;
  begin_call(c);
  mark(c, EV_H2D0, s);
  if (n == 0) {
    *out->n_changes = 0;
    mark(c, EV_START, s); mark(c, EV_SORT, s); mark(c, EV_MERGE, s); mark(c, EV_D2H, s);
    return BB_OK;
  }
  uint64_t chunk = (n + MAX_CHUNKS - 1) / MAX_CHUNKS;
  if (chunk < MIN_CHUNK) chunk = MIN_CHUNK;
  if (c->cfg.flags & BB_CFG_ORDERED_CHANGES) chunk = n;  // one path-major run, as promised
  const int nchunks = (int)((n + chunk - 1) / chunk);
  {
    int rc = reserve_io(c, n);
    if (rc) return rc;
    rc = reserve_dev(c, chunk);
    if (rc) return rc;
  }
  BB_CUDA(c, cudaEventRecord(c->ev_in[0], s));  // the side streams start after what is queued on ours
  BB_CUDA(c, cudaStreamWaitEvent(c->s_h2d, c->ev_in[0], 0));
  BB_CUDA(c, cudaStreamWaitEvent(c->s_d2h, c->ev_in[0], 0));
  // every path id first: a bad one must reject the batch before any chunk touches the table
  BB_CUDA(c, cudaMemcpyAsync(c->io_path.p, in->path_id, n * 8, cudaMemcpyHostToDevice, c->s_h2d));
  for (int i = 0; i < nchunks; ++i) {
    const uint64_t o = (uint64_t)i * chunk, m = (o + chunk <= n) ? chunk : n - o;
    BB_CUDA(c, cudaMemcpyAsync(c->io_head.p + o, in->head + o, m * 16, cudaMemcpyHostToDevice, c->s_h2d));
    BB_CUDA(c, cudaMemcpyAsync(c->io_clk.p + 2 * o, in->clk + 8 * o, m * 32, cudaMemcpyHostToDevice, c->s_h2d));
    BB_CUDA(c, cudaMemcpyAsync(c->io_val.p + 2 * o, in->val + 4 * o, m * 32, cudaMemcpyHostToDevice, c->s_h2d));
    BB_CUDA(c, cudaEventRecord(c->ev_in[i], c->s_h2d));
  }
  bb_changes dout{n, c->io_verdict.p, c->d_nchanges, c->io_out_idx.p,
                  reinterpret_cast<bb_head*>(c->io_out_head.p), reinterpret_cast<uint32_t*>(c->io_out_clk.p),
                  reinterpret_cast<uint64_t*>(c->io_out_val.p)};
  BB_CUDA(c, cudaStreamWaitEvent(s, c->ev_in[0], 0));
  mark(c, EV_START, s);
  BB_CUDA(c, cudaMemsetAsync(c->d_nchanges, 0, sizeof(uint64_t), s));
  if (nchunks > 1) BB_LAUNCH(c, bb::k_check_range, 296, 256, s, c->io_path.p, n, c->cfg.capacity, c->d_err);
  for (int i = 0; i < nchunks; ++i) {
    const uint64_t o = (uint64_t)i * chunk, m = (o + chunk <= n) ? chunk : n - o;
    bb_batch din{m, c->io_path.p + o, reinterpret_cast<const bb_head*>(c->io_head.p + o),
                 reinterpret_cast<const uint32_t*>(c->io_clk.p + 2 * o),
                 reinterpret_cast<const uint64_t*>(c->io_val.p + 2 * o)};
    bb_changes dchunk = dout;
    dchunk.verdict = c->io_verdict.p + o;
    BB_CUDA(c, cudaStreamWaitEvent(s, c->ev_in[i], 0));
    int rc = merge_dev(c, &din, &dchunk, s, (uint32_t)o, true);
    if (rc) return rc;
    BB_CUDA(c, cudaEventRecord(c->ev_done[i], s));
  }
  mark(c, EV_SORT, s);  // (phases of a chunked call: "sort" is 0, "merge" = sort + merge of every chunk)
  mark(c, EV_MERGE, s);
  uint64_t done = 0;
  int rc = BB_OK;
  for (int i = 0; i < nchunks; ++i) {  // a chunk's entries are final once its kernels are: ship them
    const uint64_t o = (uint64_t)i * chunk, m = (o + chunk <= n) ? chunk : n - o;
    BB_CUDA(c, cudaStreamWaitEvent(c->s_d2h, c->ev_done[i], 0));
    BB_CUDA(c, cudaMemcpyAsync(c->h_nchanges + i, c->d_nchanges, 8, cudaMemcpyDeviceToHost, c->s_d2h));
    BB_CUDA(c, cudaEventRecord(c->ev_cnt[i], c->s_d2h));
    BB_CUDA(c, cudaMemcpyAsync(out->verdict + o, c->io_verdict.p + o, m * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    BB_CUDA(c, cudaEventSynchronize(c->ev_cnt[i]));
    const uint64_t k = c->h_nchanges[i];
    if (k > out->cap) {
      rc = fail(c, BB_ERR_CAPACITY, "change-set buffer too small");
      break;
    }
    if (k > done && (!out->idx || !out->head || !out->clk || !out->val)) {
      rc = fail(c, BB_ERR_ARG, "null buffer");
      break;
    }
    if (k > done) {
      const uint64_t cnt = k - done;
      BB_CUDA(c, cudaMemcpyAsync(out->idx + done, c->io_out_idx.p + done, cnt * 4, cudaMemcpyDeviceToHost, c->s_d2h));
      BB_CUDA(c, cudaMemcpyAsync(out->head + done, c->io_out_head.p + done, cnt * 16, cudaMemcpyDeviceToHost, c->s_d2h));
      BB_CUDA(c, cudaMemcpyAsync(out->clk + 8 * done, c->io_out_clk.p + 2 * done, cnt * 32, cudaMemcpyDeviceToHost, c->s_d2h));
      BB_CUDA(c, cudaMemcpyAsync(out->val + 4 * done, c->io_out_val.p + 2 * done, cnt * 32, cudaMemcpyDeviceToHost, c->s_d2h));
    }
    done = k;
  }
  BB_CUDA(c, cudaEventRecord(c->ev_cnt[0], c->s_d2h));
  BB_CUDA(c, cudaStreamWaitEvent(s, c->ev_cnt[0], 0));  // our stream is "done" when the last copy is
  mark(c, EV_D2H, s);
  const int drc = collect_device_error(c, s);  // synchronises everything
  if (rc) return rc;
  if (drc) return drc;
  *out->n_changes = done;
  return BB_OK;
}

int bb_route_pack_dev(bb_ctx* c, uint32_t world, const bb_batch* in, bb_batch* out, uint64_t* counts, void* stream) {
  using namespace bb;
  if (!c || !in || !out || !counts) return fail(c, BB_ERR_ARG, "null argument");
  if (world < 1 || world > RT_MAX_WORLD) return fail(c, BB_ERR_ARG, "world must be 1..16");
  const uint64_t n = in->n;
  if (n >= 0xFFFFFFFFull) return fail(c, BB_ERR_ARG, "batch too large");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  if (n == 0) {
    BB_CUDA(c, cudaMemsetAsync(counts, 0, world * sizeof(uint64_t), s));
    return BB_OK;
  }
  if (!in->path_id || !in->head || !in->clk || !in->val || !out->path_id || !out->head || !out->clk || !out->val)
    return fail(c, BB_ERR_ARG, "null buffer");
  const uint32_t tiles = div_up(n, RT_THREADS);
  BB_CUDA(c, c->route_tiles.ensure((size_t)tiles * world));
  BB_LAUNCH(c, k_route_count, tiles, RT_THREADS, s, in->path_id, n, world, c->route_tiles.p);
  BB_LAUNCH(c, k_route_scan, 1, RT_THREADS, s, c->route_tiles.p, tiles, world, counts);
  RouteArgs a;
  a.path_id = in->path_id;
  a.head = reinterpret_cast<const uint4*>(in->head);
  a.clk = reinterpret_cast<const uint4*>(in->clk);
  a.val = reinterpret_cast<const uint4*>(in->val);
  a.o_path = const_cast<uint64_t*>(out->path_id);
  a.o_head = reinterpret_cast<uint4*>(const_cast<bb_head*>(out->head));
  a.o_clk = reinterpret_cast<uint4*>(const_cast<uint32_t*>(out->clk));
  a.o_val = reinterpret_cast<uint4*>(const_cast<uint64_t*>(out->val));
  a.n = n;
  a.world = world;
  a.tile_off = c->route_tiles.p;
  BB_LAUNCH(c, k_route_scatter, tiles, RT_THREADS, s, a);
  return BB_OK;
}

int bb_index_create(bb_ctx* c, uint32_t field, uint64_t extra_capacity) {
  if (!c) return BB_ERR_ARG;
  if (field >= c->cfg.n_fields) return fail(c, BB_ERR_ARG, "field slot out of range");
  if (c->index[field].live) return BB_OK;  // query:33-35
  if (!(c->cfg.flags & BB_CFG_POST_GETDATA))
    return fail(c, BB_ERR_STATE, "indices need a ctx created with BB_CFG_POST_GETDATA");
  if (extra_capacity >= (1ull << 31)) return fail(c, BB_ERR_ARG, "extra_capacity too large");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  bb_ctx::IndexDev& ix = c->index[field];
  uint64_t slots = 1024;
  while (slots - (slots >> 3) <= extra_capacity + 1) slots <<= 1;
  const uint64_t cap2 = (c->cfg.capacity + 1) & ~1ull;
  if (cudaMalloc((void**)&ix.pcol, cap2 * sizeof(uint64_t)) != cudaSuccess ||
      cudaMalloc((void**)&ix.xkey, slots * sizeof(uint64_t)) != cudaSuccess ||
      cudaMalloc((void**)&ix.xnode, slots * sizeof(uint32_t)) != cudaSuccess) {
    if (ix.pcol) cudaFree(ix.pcol);
    if (ix.xkey) cudaFree(ix.xkey);
    if (ix.xnode) cudaFree(ix.xnode);
    ix = bb_ctx::IndexDev();
    return fail(c, BB_ERR_CUDA, "index allocation failed", cudaGetLastError());
  }
  ix.xslots = slots;
  begin_call(c);
  mark(c, EV_Q0, c->stream);
  int rc = index_fill(c, (int)field, c->stream);
  if (rc) return rc;
  mark(c, EV_Q1, c->stream);
  BB_CUDA(c, cudaStreamSynchronize(c->stream));
  ix.live = true;
  c->index_mask |= 1u << field;
  return BB_OK;
}

int bb_query_equals(bb_ctx* c, uint32_t field, uint64_t key, bb_hits* out) {
  if (!c) return BB_ERR_ARG;
  bb::Pred p{};
  p.mode = 0;
  p.eq = key;
  return query_host(c, field, p, out);
}

int bb_query_range(bb_ctx* c, uint32_t field, const bb_bound* lo, const bb_bound* hi, bb_hits* out) {
  if (!c) return BB_ERR_ARG;
  if (!lo || !hi) return fail(c, BB_ERR_ARG, "null argument");
  return query_host(c, field, range_pred(lo, hi), out);
}

int bb_query_equals_dev(bb_ctx* c, uint32_t field, uint64_t key, bb_hits* out, void* stream) {
  if (!c) return BB_ERR_ARG;
  bb::Pred p{};
  p.mode = 0;
  p.eq = key;
  return query_dev(c, field, p, out, stream);
}

int bb_query_range_dev(bb_ctx* c, uint32_t field, const bb_bound* lo, const bb_bound* hi, bb_hits* out,
                       void* stream) {
  if (!c) return BB_ERR_ARG;
  if (!lo || !hi) return fail(c, BB_ERR_ARG, "null argument");
  return query_dev(c, field, range_pred(lo, hi), out, stream);
}

int bb_query_count(bb_ctx* c, uint32_t field, uint64_t key, uint64_t* count) {
  if (!c) return BB_ERR_ARG;
  if (!count) return fail(c, BB_ERR_ARG, "null argument");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  bb::Pred p{};
  p.mode = 0;
  p.eq = key;
  begin_call(c);
  int rc = scan_dev(c, field, p, nullptr, 0, c->stream);
  if (rc) return rc;
  BB_CUDA(c, cudaMemcpyAsync(c->h_counters, c->d_counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  BB_CUDA(c, cudaStreamSynchronize(c->stream));
  *count = c->h_counters[0] + c->h_counters[1];
  return BB_OK;
}

int bb_index_stats(bb_ctx* c, uint32_t field, uint64_t* n_dense, uint64_t* n_extra) {
  if (!c) return BB_ERR_ARG;
  if (!n_dense || !n_extra) return fail(c, BB_ERR_ARG, "null argument");
  if (field >= c->cfg.n_fields || !c->index[field].live) return fail(c, BB_ERR_STATE, "no index on this field");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  cudaStream_t s = c->stream;
  const bb_ctx::IndexDev& ix = c->index[field];
  BB_CUDA(c, cudaMemsetAsync(c->d_counters, 0, 2 * sizeof(unsigned long long), s));
  BB_LAUNCH(c, bb::k_count_live, 1184, 256, s, ix.pcol, c->cfg.capacity, c->d_counters);
  BB_LAUNCH(c, bb::k_count_live, 1184, 256, s, ix.xkey, ix.xslots, c->d_counters + 1);
  BB_CUDA(c, cudaMemcpyAsync(c->h_counters, c->d_counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  BB_CUDA(c, cudaStreamSynchronize(s));
  *n_dense = c->h_counters[0];
  *n_extra = c->h_counters[1];
  return BB_OK;
}

uint64_t bb_launch_count(const bb_ctx* c) { return c ? c->launches : 0; }

double bb_last_phase_ms(bb_ctx* c, const char* phase) { return bb_phase_ms(c, phase, 0); }

double bb_phase_ms(bb_ctx* c, const char* phase, uint32_t calls_ago) {
  if (!c || !phase || calls_ago >= EV_RING || calls_ago >= c->calls) return -1.0;
  const int slot = (int)((c->calls - 1 - calls_ago) % EV_RING);
  int a = -1, b = -1;
  if (!strcmp(phase, "h2d")) { a = EV_H2D0; b = EV_START; }
  else if (!strcmp(phase, "sort")) { a = EV_START; b = EV_SORT; }
  else if (!strcmp(phase, "merge")) { a = EV_SORT; b = EV_MERGE; }
  else if (!strcmp(phase, "d2h")) { a = EV_MERGE; b = EV_D2H; }
  else if (!strcmp(phase, "device")) { a = EV_START; b = EV_MERGE; }
  else if (!strcmp(phase, "total")) { a = EV_H2D0; b = EV_D2H; }
  else if (!strcmp(phase, "scan")) { a = EV_Q0; b = EV_Q1; }
  if (a < 0 || !c->ev_valid[slot][a] || !c->ev_valid[slot][b]) return -1.0;
  cudaSetDevice(c->cfg.device);
  if (cudaEventSynchronize(c->ev[slot][b]) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev[slot][a], c->ev[slot][b]) != cudaSuccess) return -1.0;
  return (double)ms;
}

}  // extern "C"
