// bullet_b200.cu - the C ABI of include/bullet_b200.h over the sm_100a kernels.
// One bb_ctx == one GPU-resident shard of the graph table (rows of 128 bytes,
// row index == interned path id) plus the scratch the pipeline needs.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: libnccl.so.2 is opened at run time

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "bb_group.cuh"
#include "bb_kernels.cuh"
#include "bb_route.cuh"

static_assert(sizeof(bb_row) == 128, "table rows are one 128-byte line");
static_assert(sizeof(bb_head) == 16, "heads are one 16-byte vector");

namespace {

thread_local std::string g_create_error;

enum { EV_H2D0, EV_START, EV_SORT, EV_MERGE, EV_D2H, EV_Q0, EV_Q1, EV_COUNT };
constexpr int EV_RING = 64;  // merge calls whose phase timings can still be queried
constexpr int MAX_CHUNKS = 16;         // a host call is pipelined as up to this many chunks (default: host_chunks)
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
  int n_sm = 148;
  uint64_t launches = 0;
  std::string err;
  cudaEvent_t ev[EV_RING][EV_COUNT]{};
  bool ev_valid[EV_RING][EV_COUNT]{};
  uint64_t calls = 0;  // merge calls started; ring slot = (calls - 1) % EV_RING
  // pipeline scratch
  DevBuf<uint64_t> items_a, items_b;
  DevBuf<uint32_t> zero;  // zeroed per call: [digit histograms | tickets | sort tile states | merge tile states]
  DevBuf<uint32_t> st_idx;     // an update's rank in its path (count kernels), then staging of overrunning segments
  DevBuf<uint4> st_ent;
  uint32_t* d_err = nullptr;    // sticky until bb_sync: [0] ERR_* bits, [1] ordinal of the first rejected batch
  uint32_t* h_err = nullptr;    // pinned [2]
  uint32_t* d_callrej = nullptr;  // bit0: the chunked host call in flight holds a path id >= capacity (every chunk is skipped)
  uint32_t batches_since_sync = 0;
  // device mirrors of the host-call buffers
  DevBuf<uint64_t> io_path;
  DevBuf<uint4> io_head, io_clk, io_val, io_out_head, io_out_clk, io_out_val, io_rows;
  DevBuf<uint32_t> io_verdict, io_out_idx;
  // per-row scratch of the front ends, allocated on first use.  Sorted paths: cs_cnt u32[capacity] (all zero between
  // calls; the grouping front end uses only this one) and its exclusive scan cs_off u32[capacity + 1], both padded
  // to whole tiles.
  uint32_t* cs_cnt = nullptr;
  uint32_t* cs_off = nullptr;
  uint2* cg_off = nullptr;     // grouping front end: (start, length) of a multi-update path's run, uint2[capacity]
  uint32_t* cg_ctr = nullptr;  // [2][CG_CTR_WORDS] per-batch counters, sets used alternately
  uint32_t cg_parity = 0;
  DevBuf<uint4> hot_list;      // segments k_merge_stage hands to k_merge_hot
  DevBuf<uint2> cs_long;       // sorted paths: segments longer than CS_SHORT, queued for k_cs_fix_long
  DevBuf<uint32_t> cs_tile;    // per-4096-row sums of cs_cnt
  uint32_t tune = 0;  // env BB_TUNE (experiments): 16 = the count kernel prefetches the rows into L2
  bool phase_events = false;   // record the event between the count and the merge kernels (it serialises them)
  uint64_t* d_nchanges = nullptr;
  uint64_t* d_chunk_total = nullptr;  // [MAX_CHUNKS] host calls: the change count as it stood when chunk i was done
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
    // BB_CFG_EXACT_ORDER: entry tags and the bucket table key -> (count, created)
    uint64_t* pseq = nullptr;
    uint64_t* xseq = nullptr;
    uint64_t* bkey = nullptr;
    uint32_t* bcount = nullptr;
    uint64_t* bcreated = nullptr;
    uint64_t bslots = 0;
  } index[BB_MAX_FIELDS];
  DevBuf<uint64_t> ev_key, ev_tag;   // hook events of the running batch, then sort buffers of a query
  DevBuf<uint64_t> q_a, q_b;         // a query's sort keys (bucket created, entry tag)
  DevBuf<uint32_t> ev_node;          // ... and its payload
  uint32_t* d_evcount = nullptr;
  uint32_t ev_padded = 0;            // events the buffers hold: a power of two (the sort pads in place)
  uint32_t index_mask = 0;
  uint32_t* d_xused = nullptr;             // [BB_MAX_FIELDS]
  uint32_t* epoch_col = nullptr;  // BB_CFG_TRACK_MODIFIED: per row, the ordinal of the merge call that last wrote it
  uint64_t epoch = 0;             // ordinal of the most recent merge call
  uint32_t host_chunks = 8;  // pieces a bb_merge_batch call is pipelined in (env BB_HOST_CHUNKS, 1..16)
  int rows_tma = 1;       // rows of k_merge_stage staged and written back with cp.async.bulk (BB_MERGE_TMA=0: cp.async, for A/B)
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

#define BB_LAUNCH(c, kernel, grid, block, stream, ...)                                        \
  do {                                                                                        \
    cudaError_t e_ = bb_launch(kernel, (uint32_t)(grid), (uint32_t)(block), 0, (stream), false, __VA_ARGS__); \
    ++(c)->launches;                                                                          \
    if (e_ == cudaSuccess) e_ = cudaGetLastError();                                           \
    if (e_ != cudaSuccess) return fail((c), BB_ERR_CUDA, #kernel, e_);                        \
  } while (0)
// same, with programmatic stream serialisation: the kernel calls pdl_wait() before it touches its predecessor's output
#define BB_LAUNCH_PDL(c, kernel, grid, block, smem, stream, ...)                              \
  do {                                                                                        \
    cudaError_t e_ = bb_launch(kernel, (uint32_t)(grid), (uint32_t)(block), (smem), (stream), true, __VA_ARGS__); \
    ++(c)->launches;                                                                          \
    if (e_ == cudaSuccess) e_ = cudaGetLastError();                                           \
    if (e_ != cudaSuccess) return fail((c), BB_ERR_CUDA, #kernel, e_);                        \
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

// How a batch is brought into "a path's updates adjacent, in arrival order":
//   grouping (default)  O(batch) passes, singles stay in arrival order (bb_group.cuh)
//   counting sort       BB_CFG_ORDERED_CHANGES / BB_CFG_FULL_SORT, when the scan over the rows is cheap
//                       next to the batch: the item list is ascending in path id
//   radix sort          otherwise, or with BB_CFG_RADIX_SORT
bool ordered_cfg(const bb_ctx* c) { return (c->cfg.flags & BB_CFG_ORDERED_CHANGES) != 0; }

bool use_grouping(const bb_ctx* c) {
  return !(c->cfg.flags & (BB_CFG_RADIX_SORT | BB_CFG_ORDERED_CHANGES | BB_CFG_FULL_SORT | BB_CFG_EXACT_ORDER));
}


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
  z.total = z.cs_ctr + 8;  // k_cs_fix counters [0..1], this batch's reject word [5]
  return z;
}

// size every scratch buffer of the device pipeline for batches of up to n updates
int reserve_dev(bb_ctx* c, uint64_t n) {
  BB_CUDA(c, c->st_idx.ensure(n));
  if (use_grouping(c)) {
    if (!c->cs_cnt) {  // per-path update counts (zero between calls) and run descriptors
      BB_CUDA(c, cudaMalloc((void**)&c->cs_cnt, c->cfg.capacity * sizeof(uint32_t)));
      BB_CUDA(c, cudaMalloc((void**)&c->cg_off, c->cfg.capacity * sizeof(uint2)));
      BB_CUDA(c, cudaMemsetAsync(c->cs_cnt, 0, c->cfg.capacity * sizeof(uint32_t), c->stream));
      BB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    BB_CUDA(c, c->items_a.ensure(n));
    BB_CUDA(c, c->items_b.ensure(n));
    BB_CUDA(c, c->st_ent.ensure(5 * n));
    BB_CUDA(c, c->cs_long.ensure(n / 8 + 1));
    BB_CUDA(c, c->hot_list.ensure(n / bb::HOT_MIN + 16));
    return BB_OK;
  }
  if (!c->cs_cnt) {
    const size_t cs_words = ((size_t)c->cfg.capacity + 1 + bb::CS_TILE - 1) / bb::CS_TILE * bb::CS_TILE;
    BB_CUDA(c, cudaMalloc((void**)&c->cs_cnt, cs_words * sizeof(uint32_t)));
    BB_CUDA(c, cudaMalloc((void**)&c->cs_off, cs_words * sizeof(uint32_t)));
    BB_CUDA(c, cudaMemsetAsync(c->cs_cnt, 0, cs_words * sizeof(uint32_t), c->stream));
    BB_CUDA(c, cudaMemsetAsync(c->cs_off, 0, cs_words * sizeof(uint32_t), c->stream));
    BB_CUDA(c, c->cs_tile.ensure(cs_words / bb::CS_TILE));
    BB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  const ZeroLayout z = zero_layout(c, n);
  BB_CUDA(c, c->items_a.ensure(n));
  BB_CUDA(c, c->items_b.ensure(n));
  BB_CUDA(c, c->zero.ensure(z.total));
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

void fill_params(const bb_ctx* c, bb::Params& p, bb::IndexArgs& ix) {
  p.rank_object = c->cfg.rank_object;
  p.me = c->cfg.local_peer;
  p.post_getdata = (c->cfg.flags & BB_CFG_POST_GETDATA) != 0;
  ix.mask = c->index_mask;
  ix.xused = c->d_xused;
  for (int f = 0; f < BB_MAX_FIELDS; ++f) {
    ix.pcol[f] = c->index[f].pcol;
    ix.xkey[f] = c->index[f].xkey;
    ix.xnode[f] = c->index[f].xnode;
    ix.xmask[f] = c->index[f].live ? (uint32_t)(c->index[f].xslots - 1) : 0u;
    ix.pseq[f] = c->index[f].pseq;
    ix.xseq[f] = c->index[f].xseq;
  }
  ix.ev_key = c->ev_key.p;
  ix.ev_tag = c->ev_tag.p;
  ix.ev_count = c->d_evcount;
  ix.ev_cap = c->ev_padded;
}

bool exact_order(const bb_ctx* c) { return (c->cfg.flags & BB_CFG_EXACT_ORDER) != 0; }

uint32_t pow2_at_least(uint64_t n) {
  uint32_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

bb::BucketTable bucket_table(const bb_ctx* c, int f) {
  const bb_ctx::IndexDev& ix = c->index[f];
  return bb::BucketTable{ix.bkey, ix.bcount, ix.bcreated, ix.bslots ? (uint32_t)(ix.bslots - 1) : 0u};
}

// bitonic sort of the first `padded` (a power of two) pairs of (a, b), lexicographic, with an optional payload
int bitonic_sort(bb_ctx* c, uint64_t* a, uint64_t* b, uint32_t* v, uint32_t padded, cudaStream_t s) {
  for (uint32_t k = 2; k <= padded; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1)
      BB_LAUNCH(c, bb::k_bitonic_step, div_up(padded, 256), 256, s, a, b, v, padded, k, j);
  return BB_OK;
}

// BB_CFG_EXACT_ORDER: the events the hook (or an index build) logged -> sorted by (key, field, tag) -> replayed per bucket
int process_events(bb_ctx* c, cudaStream_t s) {
  const uint32_t padded = c->ev_padded;
  if (padded == 0) return BB_OK;
  BB_LAUNCH(c, bb::k_sort_pad, div_up(padded, 256), 256, s, c->ev_key.p, c->ev_tag.p, c->d_evcount, 0u, padded);
  int rc = bitonic_sort(c, c->ev_key.p, c->ev_tag.p, nullptr, padded, s);
  if (rc) return rc;
  BB_LAUNCH(c, bb::k_bucket_replay, div_up(padded, 256), 256, s, c->ev_key.p, c->ev_tag.p, c->d_evcount, padded, bucket_table(c, 0),
            bucket_table(c, 1), bucket_table(c, 2), bucket_table(c, 3), c->d_err);
  BB_CUDA(c, cudaMemsetAsync(c->d_evcount, 0, sizeof(uint32_t), s));
  return BB_OK;
}

// room for `n` events (a power of two, so that the sort can pad in place)
int reserve_events(bb_ctx* c, uint64_t n) {
  if (n > 0x40000000ull) return fail(c, BB_ERR_ARG, "batch too large for BB_CFG_EXACT_ORDER");
  const uint32_t p = pow2_at_least(std::max<uint64_t>(n, 256));
  if (p <= c->ev_padded) return BB_OK;
  BB_CUDA(c, c->ev_key.ensure(p));
  BB_CUDA(c, c->ev_tag.ensure(p));
  c->ev_padded = p;
  return BB_OK;
}

// One batch (or one chunk of a host call: `idx_base` = arrival index of its first update,
// `append` = keep adding to *out->n_changes instead of starting a new change set; `call_rej` = the word
// that says the whole host call is rejected).
int merge_dev(bb_ctx* c, const bb_batch* in, bb_changes* out, cudaStream_t s, uint32_t idx_base = 0,
              bool append = false, const uint32_t* call_rej = nullptr) {
  using namespace bb;
  const uint64_t n = in->n;
  if (n >= BB_NO_SLOT) return fail(c, BB_ERR_ARG, "batch larger than 2^29-2 updates");
  const bool grouped = use_grouping(c);
  if (!append) {
    mark(c, EV_START, s);
    if (!grouped || n == 0) BB_CUDA(c, cudaMemsetAsync(out->n_changes, 0, sizeof(uint64_t), s));
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
  ++c->batches_since_sync;
  const ZeroLayout z = zero_layout(c, n);
  uint32_t* zp = c->zero.p;
  if (!grouped) BB_CUDA(c, cudaMemsetAsync(zp, 0, z.total * sizeof(uint32_t), s));
  uint32_t* rej = grouped ? nullptr : zp + z.cs_ctr + 5;  // this batch's reject word
  const uint32_t ordinal = c->batches_since_sync;
  uint32_t* gctr = nullptr;

  uint64_t* src = c->items_a.p;
  if (grouped) {
    // default: count, then group (singles in arrival order, multi-update paths in claimed runs behind them); the
    // five launches and the merge behind them are chained with programmatic dependent launch
    const uint32_t g4 = div_up(n, CS_THREADS * CS_ILP);
    gctr = c->cg_ctr + (size_t)c->cg_parity * CG_CTR_WORDS;
    uint32_t* gnext = c->cg_ctr + (size_t)(c->cg_parity ^ 1u) * CG_CTR_WORDS;
    c->cg_parity ^= 1u;
    rej = gctr + CG_CTR_REJ;
    BB_LAUNCH_PDL(c, k_cg_count, g4, CS_THREADS, 0, s, in->path_id, n, c->cfg.capacity, c->cs_cnt, c->st_idx.p, gctr, gnext,
                  c->d_err, ordinal, append ? (uint64_t*)nullptr : out->n_changes,
                  (const uint4*)((c->tune & 16u) ? c->table : nullptr));
    BB_LAUNCH_PDL(c, k_cg_classify, g4, CS_THREADS, 0, s, in->path_id, n, c->cfg.capacity, c->cs_cnt, c->st_idx.p, c->cg_off, src,
                  gctr, c->cs_long.p);
    BB_LAUNCH_PDL(c, k_cg_place, g4, CS_THREADS, 0, s, in->path_id, n, c->st_idx.p, c->cg_off, c->cs_cnt, src, gctr);
    BB_LAUNCH_PDL(c, k_cg_fix, div_up(n, CS_THREADS), CS_THREADS, 0, s, src, c->cg_off, gctr);
    BB_LAUNCH_PDL(c, k_cs_fix_long, CS_LONG_CTAS, CS_THREADS, 0, s, src, c->items_b.p, c->cs_long.p, gctr + CG_CTR_LONG,
                  gctr + CG_CTR_NEXT, gctr + CG_CTR_SINGLE);
  } else if (use_counting_sort(c, n)) {
    // K1': counting sort keyed by the row index (bb_kernels.cuh), arrival order restored per segment
    const uint32_t g = div_up(n, CS_THREADS);
    const uint64_t cap = c->cfg.capacity;
    const uint32_t tiles = div_up(cap + 1, CS_TILE);  // + 1: off[capacity] = the batch size
    BB_LAUNCH(c, k_cs_count, div_up(n, CS_THREADS * CS_ILP), CS_THREADS, s, in->path_id, n, cap, c->cs_cnt, c->st_idx.p, rej,
              c->d_err, ordinal);
    BB_LAUNCH(c, k_cs_tile_sums, tiles, CS_THREADS, s, c->cs_cnt, c->cs_tile.p);
    BB_LAUNCH(c, k_cs_offsets, tiles, CS_THREADS, s, c->cs_cnt, c->cs_tile.p, c->cs_off);
    BB_LAUNCH(c, k_cs_place, div_up(n, CS_THREADS * CS_ILP), CS_THREADS, s, in->path_id, n, cap, c->st_idx.p, c->cs_off, src, rej);
    BB_LAUNCH(c, k_cs_fix, g, CS_THREADS, s, src, n, c->cs_off, c->cs_long.p, zp + z.cs_ctr, rej);
    BB_LAUNCH(c, k_cs_fix_long, CS_LONG_CTAS, CS_THREADS, s, src, c->items_b.p, c->cs_long.p, zp + z.cs_ctr,
              zp + z.cs_ctr + 1, (const uint32_t*)nullptr);
  } else {
    // K0 + K1: stable LSD radix sort of (path id, arrival index) by path id
    BB_LAUNCH(c, k_keys_hist, z.sort_tiles, SORT_THREADS, s, in->path_id, n, c->cfg.capacity, (int)z.passes,
              c->items_a.p, zp + z.hist, rej, c->d_err, ordinal);
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
  if (!append && (!grouped || c->phase_events)) mark(c, EV_SORT, s);  // (an event here ends the overlap of the launches)
  if (c->cfg.flags & BB_CFG_ORDERED_CHANGES)
    BB_CUDA(c, cudaMemcpyAsync(c->d_chg_base, out->n_changes, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));

  // K2s: per-path sequential replay over the sorted item list + change-set compaction
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
  ma.tile_state = grouped ? nullptr : zp + z.merge_state;
  ma.ticket = grouped ? nullptr : zp + z.tickets + MAX_PASSES;
  ma.num_tiles = z.merge_tiles;
  ma.hot_list = c->hot_list.p;
  ma.n_hot = grouped ? gctr + CG_CTR_NHOT : nullptr;
  ma.hot_cap = grouped ? (uint32_t)std::min<uint64_t>(c->hot_list.cap, 0x7FFFFFFFull) : 0u;
  ma.seq_base = c->seq;
  ma.idx_base = idx_base;
  ma.chg_base = c->d_chg_base;
  ma.epoch_col = c->epoch_col;
  ma.epoch = (uint32_t)c->epoch;
  ma.err = c->d_err;
  ma.rej = call_rej ? call_rej : rej;
  const bool exact = exact_order(c) && c->index_mask;
  if (exact) {  // at most one effective remove + one effective add per update and index
    int rc = reserve_events(c, 2 * n * (uint64_t)__builtin_popcount(c->index_mask));
    if (rc) return rc;
  }
  fill_params(c, ma.p, ma.ix);
  const bool ordered = (c->cfg.flags & BB_CFG_ORDERED_CHANGES) != 0;
  if (grouped) {  // hot keys are handed to k_merge_hot, a CTA per segment (exits at once when there are none)
    // rows staged and written back by the copy engine (cp.async.bulk + mbarrier); BB_MERGE_TMA=0 keeps the cp.async
    // (LDGSTS) staging of the plain variant for A/B measurements
    const bool compact = (c->cfg.flags & BB_CFG_COMPACT_CHANGES) != 0;
    if (compact && c->index_mask) {
      BB_LAUNCH_PDL(c, (k_merge_stage<false, true, true, true, 1>), z.merge_tiles, MT, 0, s, ma);
      BB_LAUNCH_PDL(c, (k_merge_hot<true, true>), HOT_CTAS, HOT_T, 0, s, ma);
    } else if (compact) {
      BB_LAUNCH_PDL(c, (k_merge_stage<false, false, true, true, 1>), z.merge_tiles, MT, 0, s, ma);
      BB_LAUNCH_PDL(c, (k_merge_hot<false, true>), HOT_CTAS, HOT_T, 0, s, ma);
    } else if (c->index_mask) {
      BB_LAUNCH_PDL(c, (k_merge_stage<false, true, true, false, 1>), z.merge_tiles, MT, 0, s, ma);
      BB_LAUNCH_PDL(c, k_merge_hot<true>, HOT_CTAS, HOT_T, 0, s, ma);
    } else if (c->rows_tma == 2) {
      BB_LAUNCH_PDL(c, (k_merge_stage<false, false, true, false, 2>), z.merge_tiles, MT, 0, s, ma);
      BB_LAUNCH_PDL(c, k_merge_hot<false>, HOT_CTAS, HOT_T, 0, s, ma);
    } else if (c->rows_tma) {
      BB_LAUNCH_PDL(c, (k_merge_stage<false, false, true, false, 1>), z.merge_tiles, MT, 0, s, ma);
      BB_LAUNCH_PDL(c, k_merge_hot<false>, HOT_CTAS, HOT_T, 0, s, ma);
    } else {
      BB_LAUNCH_PDL(c, (k_merge_stage<false, false, true>), z.merge_tiles, MT, 0, s, ma);
      BB_LAUNCH_PDL(c, k_merge_hot<false>, HOT_CTAS, HOT_T, 0, s, ma);
    }
  } else if (exact) {  // entry tags + hook events, then the per-bucket replay (hot keys: serial, in order, by their owner thread)
    BB_LAUNCH(c, (k_merge_stage<false, true, false, false, 0, true>), z.merge_tiles, MT, s, ma);
    int rc = process_events(c, s);
    if (rc) return rc;
  } else if (c->index_mask) {
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

// fetch + clear the deferred device error words; stream must be idle afterwards
int collect_device_error(bb_ctx* c, cudaStream_t s) {
  static const uint32_t clear[2] = {0u, 0xFFFFFFFFu};
  BB_CUDA(c, cudaMemcpyAsync(c->h_err, c->d_err, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  BB_CUDA(c, cudaMemcpyAsync(c->d_err, clear, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  BB_CUDA(c, cudaStreamSynchronize(s));
  const uint32_t e = c->h_err[0], first = c->h_err[1];
  c->batches_since_sync = 0;
  if (e & bb::ERR_RANGE) {
    char msg[160];
    snprintf(msg, sizeof msg, "path id >= capacity: batch %u since the last sync was rejected whole, table unchanged (later batches were merged)", first);
    return fail(c, BB_ERR_CAPACITY, msg);
  }
  if (e & bb::ERR_CHANGES) return fail(c, BB_ERR_CAPACITY, "change-set buffer too small");
  if (e & bb::ERR_XFULL) return fail(c, BB_ERR_CAPACITY, "index overflow set is full (bb_index_create extra_capacity)");
  if (e & bb::ERR_HITS) return fail(c, BB_ERR_CAPACITY, "hit buffer too small");
  return BB_OK;
}

// (re)initialise the indices on the fields of `mask` from the current table: ONE pass over the rows for all of them
int index_fill(bb_ctx* c, uint32_t mask, cudaStream_t s) {
  bb::BuildArgs a;
  a.table = c->table;
  a.capacity = c->cfg.capacity;
  a.padded = (c->cfg.capacity + 1) & ~1ull;
  a.mask = mask;
  for (int f = 0; f < BB_MAX_FIELDS; ++f) {
    bb_ctx::IndexDev& ix = c->index[f];
    a.pcol[f] = ix.pcol;
    if (!((mask >> f) & 1u)) continue;
    BB_CUDA(c, cudaMemsetAsync(ix.xkey, 0xFF, ix.xslots * sizeof(uint64_t), s));
    BB_CUDA(c, cudaMemsetAsync(ix.xnode, 0xFF, ix.xslots * sizeof(uint32_t), s));
    BB_CUDA(c, cudaMemsetAsync(c->d_xused + f, 0, sizeof(uint32_t), s));
  }
  BB_LAUNCH(c, bb::k_index_build, div_up(a.padded, 256), 256, s, a);  // 8 warps x 32 rows per CTA
  if (exact_order(c)) {  // the entries the build inserted enter the bucket tables in creation order (query:58-66)
    for (int f = 0; f < BB_MAX_FIELDS; ++f) {
      if (!((mask >> f) & 1u)) continue;
      bb_ctx::IndexDev& ix = c->index[f];
      BB_CUDA(c, cudaMemsetAsync(ix.bkey, 0xFF, ix.bslots * sizeof(uint64_t), s));
      BB_CUDA(c, cudaMemsetAsync(ix.bcount, 0, ix.bslots * sizeof(uint32_t), s));
      BB_CUDA(c, cudaMemsetAsync(ix.bcreated, 0, ix.bslots * sizeof(uint64_t), s));
      int rc = reserve_events(c, c->cfg.capacity);
      if (rc) return rc;
      bb::Params p;
      bb::IndexArgs ia;
      fill_params(c, p, ia);
      BB_LAUNCH(c, bb::k_exact_build_events, div_up(c->cfg.capacity, 256), 256, s, c->table, c->cfg.capacity, f, ia, c->d_err);
      rc = process_events(c, s);
      if (rc) return rc;
    }
  }
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
int scan_dev(bb_ctx* c, uint32_t field, const bb::Pred& pred, uint32_t* hits, uint64_t cap, cudaStream_t s, bool refs = false) {
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
  a.ref_or = 0;
  a.keys = ix.pcol; a.nodes = nullptr; a.n = n0; a.which = 0;
  a.ticket = c->scan_zero.p; a.tile_state = c->scan_zero.p + 2; a.num_tiles = t0;
  const bool ordered = (c->cfg.flags & BB_CFG_ORDERED_CHANGES) != 0;
  {
    int rc = launch_scan(c, a, t0, ordered, s);
    if (rc) return rc;
  }
  a.keys = ix.xkey; a.nodes = ix.xnode; a.n = n1; a.which = 1;
  if (refs) {  // exact order: hits leave as entry references (overflow: slot | 2^31), resolved by k_exact_hit_keys
    a.nodes = nullptr;
    a.ref_or = 0x80000000u;
  }
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
  const bool exact = exact_order(c);
  int rc = scan_dev(c, field, pred, c->io_hits.p, out->cap, s, exact);
  if (rc) return rc;
  BB_CUDA(c, cudaMemcpyAsync(c->h_counters, c->d_counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  rc = collect_device_error(c, s);  // synchronises
  if (rc) return rc;
  const uint64_t k = c->h_counters[0] + c->h_counters[1];
  const uint32_t* result = c->io_hits.p;
  if (exact && k) {
    // the reference's order (query:204, 237-258): buckets in creation order, inside a bucket in insertion order
    if (k > out->cap) return fail(c, BB_ERR_CAPACITY, "hit buffer too small");
    const uint32_t padded = pow2_at_least(k);
    BB_CUDA(c, c->q_a.ensure(padded));
    BB_CUDA(c, c->q_b.ensure(padded));
    BB_CUDA(c, c->ev_node.ensure(padded));
    bb::Params pp;
    bb::IndexArgs ia;
    fill_params(c, pp, ia);
    BB_LAUNCH(c, bb::k_exact_hit_keys, div_up(k, 256), 256, s, c->io_hits.p, (uint32_t)k, (int)field, ia, bucket_table(c, (int)field),
              c->q_a.p, c->q_b.p, c->ev_node.p);
    BB_LAUNCH(c, bb::k_sort_pad, div_up(padded, 256), 256, s, c->q_a.p, c->q_b.p, (const uint32_t*)nullptr, (uint32_t)k, padded);
    rc = bitonic_sort(c, c->q_a.p, c->q_b.p, c->ev_node.p, padded, s);
    if (rc) return rc;
    result = c->ev_node.p;
  }
  if (k) BB_CUDA(c, cudaMemcpyAsync(out->node, result, k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
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
  if ((cfg->flags & BB_CFG_COMPACT_CHANGES) && (cfg->flags & (BB_CFG_ORDERED_CHANGES | BB_CFG_RADIX_SORT | BB_CFG_FULL_SORT))) {
    g_create_error = "BB_CFG_COMPACT_CHANGES goes with the default pipeline only (not ORDERED_CHANGES / RADIX_SORT / FULL_SORT)";
    return BB_ERR_ARG;
  }
  if ((cfg->flags & BB_CFG_EXACT_ORDER) && (cfg->flags & (BB_CFG_ORDERED_CHANGES | BB_CFG_COMPACT_CHANGES))) {
    g_create_error = "BB_CFG_EXACT_ORDER does not combine with ORDERED_CHANGES / COMPACT_CHANGES";
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
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  // everything below (function attributes included) is per device: select it first
  if (cudaSetDevice(cfg->device) != cudaSuccess) {
    g_create_error = "cudaSetDevice failed";
    delete c;
    return BB_ERR_CUDA;
  }
  // k_merge_stage: 7 x 27.7 KB of static shared memory per SM: ask for the full carve-out
  cudaFuncSetAttribute(bb::k_merge_stage<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<false, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<false, false, true, false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<false, true, true, false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<false, false, true, true, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<false, true, true, true, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(bb::k_merge_stage<false, false, true, false, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (const char* e = getenv("BB_MERGE_TMA")) c->rows_tma = e[0] - '0';
  if (const char* e = getenv("BB_HOST_CHUNKS")) {
    const long v = strtol(e, nullptr, 10);
    if (v >= 1 && v <= MAX_CHUNKS) c->host_chunks = (uint32_t)v;
  }
 // 0: cp.async, 1: rows by bulk copy (default), 2: rows + payloads
  cudaFuncSetAttribute(bb::k_merge_stage<false, true, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  {
    int n_sm = 0;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && n_sm > 0) c->n_sm = n_sm;
  }
  static const uint32_t err_clear[2] = {0u, 0xFFFFFFFFu};
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc((void**)&c->table, cfg->capacity * sizeof(bb_row)) == cudaSuccess &&
            cudaMemsetAsync(c->table, 0, cfg->capacity * sizeof(bb_row), c->stream) == cudaSuccess &&
            cudaMalloc((void**)&c->d_err, 2 * sizeof(uint32_t)) == cudaSuccess &&
            cudaMemcpyAsync(c->d_err, err_clear, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
            cudaMalloc((void**)&c->d_callrej, sizeof(uint32_t)) == cudaSuccess &&
            cudaMemsetAsync(c->d_callrej, 0, sizeof(uint32_t), c->stream) == cudaSuccess &&
            cudaMalloc((void**)&c->cg_ctr, 2 * bb::CG_CTR_WORDS * sizeof(uint32_t)) == cudaSuccess &&
            cudaMemsetAsync(c->cg_ctr, 0, 2 * bb::CG_CTR_WORDS * sizeof(uint32_t), c->stream) == cudaSuccess &&
            cudaMalloc((void**)&c->d_nchanges, sizeof(uint64_t)) == cudaSuccess &&
            cudaMalloc((void**)&c->d_chunk_total, MAX_CHUNKS * sizeof(uint64_t)) == cudaSuccess &&
            cudaMalloc((void**)&c->d_chg_base, sizeof(uint64_t)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc((void**)&c->d_xused, BB_MAX_FIELDS * sizeof(uint32_t)) == cudaSuccess &&
            cudaMemsetAsync(c->d_xused, 0, BB_MAX_FIELDS * sizeof(uint32_t), c->stream) == cudaSuccess &&
            cudaMalloc((void**)&c->d_counters, 2 * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMalloc((void**)&c->d_evcount, sizeof(uint32_t)) == cudaSuccess &&
            cudaMemsetAsync(c->d_evcount, 0, sizeof(uint32_t), c->stream) == cudaSuccess &&
            cudaMallocHost((void**)&c->h_counters, 2 * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMallocHost((void**)&c->h_err, 2 * sizeof(uint32_t)) == cudaSuccess &&
            cudaMallocHost((void**)&c->h_nchanges, MAX_CHUNKS * sizeof(uint64_t)) == cudaSuccess;
  for (int i = 0; ok && i < MAX_CHUNKS; ++i)
    ok = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_cnt[i], cudaEventDisableTiming) == cudaSuccess;
  for (int r = 0; ok && r < EV_RING; ++r)
    for (int i = 0; ok && i < EV_COUNT; ++i) ok = cudaEventCreate(&c->ev[r][i]) == cudaSuccess;
  if (ok && (cfg->flags & BB_CFG_TRACK_MODIFIED))
    ok = cudaMalloc((void**)&c->epoch_col, cfg->capacity * sizeof(uint32_t)) == cudaSuccess &&
         cudaMemsetAsync(c->epoch_col, 0, cfg->capacity * sizeof(uint32_t), c->stream) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
  if (!ok) {
    g_create_error = std::string("CUDA allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    bb_destroy(c);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return BB_ERR_CUDA;
  }
  if (const char* e = getenv("BB_TUNE")) c->tune = (uint32_t)strtoul(e, nullptr, 10);
  if (prev_dev >= 0 && prev_dev != cfg->device) cudaSetDevice(prev_dev);  // every entry point selects the ctx's device itself
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
    if (c->index[f].pseq) cudaFree(c->index[f].pseq);
    if (c->index[f].xseq) cudaFree(c->index[f].xseq);
    if (c->index[f].bkey) cudaFree(c->index[f].bkey);
    if (c->index[f].bcount) cudaFree(c->index[f].bcount);
    if (c->index[f].bcreated) cudaFree(c->index[f].bcreated);
  }
  c->q_a.release(); c->q_b.release();
  if (c->d_xused) cudaFree(c->d_xused);
  if (c->d_counters) cudaFree(c->d_counters);
  if (c->d_evcount) cudaFree(c->d_evcount);
  c->ev_key.release(); c->ev_tag.release(); c->ev_node.release();
  if (c->epoch_col) cudaFree(c->epoch_col);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->table) cudaFree(c->table);
  if (c->d_err) cudaFree(c->d_err);
  if (c->cs_cnt) cudaFree(c->cs_cnt);
  if (c->cs_off) cudaFree(c->cs_off);
  if (c->cg_off) cudaFree(c->cg_off);
  if (c->cg_ctr) cudaFree(c->cg_ctr);
  c->hot_list.release();
  if (c->d_callrej) cudaFree(c->d_callrej);
  if (c->d_chunk_total) cudaFree(c->d_chunk_total);
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
  if (c->epoch_col) BB_CUDA(c, cudaMemsetAsync(c->epoch_col, 0, c->cfg.capacity * sizeof(uint32_t), c->stream));
  c->epoch = 0;
  if (c->index_mask) {
    int rc = index_fill(c, c->index_mask, c->stream);
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

uint64_t bb_epoch(const bb_ctx* c) { return c ? c->epoch : 0; }

// _collectFullSyncData(since) on the device: filter + compaction of the table rows (k_sync_collect), then only the
// selected rows cross PCIe.
int bb_sync_collect(bb_ctx* c, uint64_t since_epoch, uint32_t flags, uint64_t cap, uint64_t* path_id_out, bb_row* rows_out,
                    uint32_t* epoch_out, uint64_t* n_out) {
  if (!c || !n_out || (cap && (!path_id_out || !rows_out || !epoch_out))) return fail(c, BB_ERR_ARG, "null argument");
  if (since_epoch && !c->epoch_col)
    return fail(c, BB_ERR_STATE, "a `since` filter needs a ctx created with BB_CFG_TRACK_MODIFIED");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  cudaStream_t s = c->stream;
  const uint64_t m = cap ? cap : 1;
  BB_CUDA(c, c->io_path.ensure(m));
  BB_CUDA(c, c->io_rows.ensure(m * 8));
  BB_CUDA(c, c->io_verdict.ensure(m));
  begin_call(c);
  mark(c, EV_Q0, s);
  BB_CUDA(c, cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned long long), s));
  BB_LAUNCH(c, bb::k_sync_collect, div_up(c->cfg.capacity, 256), 256, s, c->table, c->epoch_col, c->cfg.capacity,
            (uint32_t)std::min<uint64_t>(since_epoch, 0xFFFFFFFFull), flags & BB_COLLECT_FILTER_RECORDS, cap, c->io_path.p, c->io_rows.p, c->io_verdict.p, c->d_counters);
  mark(c, EV_Q1, s);
  BB_CUDA(c, cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  int rc = collect_device_error(c, s);  // synchronises
  if (rc) return rc;
  const uint64_t k = c->h_counters[0];
  *n_out = k;
  if (k > cap) return fail(c, BB_ERR_CAPACITY, "more rows selected than the output buffers hold (n_out = how many)");
  if (k) {
    BB_CUDA(c, cudaMemcpyAsync(path_id_out, c->io_path.p, k * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    BB_CUDA(c, cudaMemcpyAsync(rows_out, c->io_rows.p, k * sizeof(bb_row), cudaMemcpyDeviceToHost, s));
    BB_CUDA(c, cudaMemcpyAsync(epoch_out, c->io_verdict.p, k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    BB_CUDA(c, cudaStreamSynchronize(s));
  }
  return BB_OK;
}

int bb_merge_batch_dev(bb_ctx* c, const bb_batch* in, bb_changes* out, void* stream) {
  if (!c || !in || !out) return fail(c, BB_ERR_ARG, "null argument");
  if (in->n && (!in->path_id || !in->head || !in->clk || !in->val || !out->verdict || !out->idx ||
                !out->head || !out->clk || !out->val))
    return fail(c, BB_ERR_ARG, "null buffer");
  if (!out->n_changes) return fail(c, BB_ERR_ARG, "null n_changes");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  begin_call(c);
  ++c->epoch;
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
  ++c->epoch;
  mark(c, EV_H2D0, s);
  if (n == 0) {
    *out->n_changes = 0;
    mark(c, EV_START, s); mark(c, EV_SORT, s); mark(c, EV_MERGE, s); mark(c, EV_D2H, s);
    return BB_OK;
  }
  uint64_t chunk = (n + c->host_chunks - 1) / c->host_chunks;
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
  const uint32_t* call_rej = nullptr;
  if (nchunks > 1) {  // one bad id anywhere rejects every chunk before any of them touches the table
    BB_CUDA(c, cudaMemsetAsync(c->d_callrej, 0, sizeof(uint32_t), s));
    BB_LAUNCH(c, bb::k_check_range, 296, 256, s, c->io_path.p, n, c->cfg.capacity, c->d_callrej, c->d_err,
              c->batches_since_sync + 1);
    call_rej = c->d_callrej;
  }
  for (int i = 0; i < nchunks; ++i) {
    const uint64_t o = (uint64_t)i * chunk, m = (o + chunk <= n) ? chunk : n - o;
    bb_batch din{m, c->io_path.p + o, reinterpret_cast<const bb_head*>(c->io_head.p + o),
                 reinterpret_cast<const uint32_t*>(c->io_clk.p + 2 * o),
                 reinterpret_cast<const uint64_t*>(c->io_val.p + 2 * o)};
    bb_changes dchunk = dout;
    dchunk.verdict = c->io_verdict.p + o;
    BB_CUDA(c, cudaStreamWaitEvent(s, c->ev_in[i], 0));
    int rc = merge_dev(c, &din, &dchunk, s, (uint32_t)o, true, call_rej);
    if (rc) return rc;
    // the change count as it stands NOW: later chunks keep adding to d_nchanges, and entries past this snapshot
    // are not written yet when the copy-out of chunk i runs
    BB_CUDA(c, cudaMemcpyAsync(c->d_chunk_total + i, c->d_nchanges, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    BB_CUDA(c, cudaEventRecord(c->ev_done[i], s));
  }
  mark(c, EV_SORT, s);  // (phases of a chunked call: "sort" is 0, "merge" = sort + merge of every chunk)
  mark(c, EV_MERGE, s);
  uint64_t done = 0;
  int rc = BB_OK;
  for (int i = 0; i < nchunks; ++i) {  // a chunk's entries are final once its kernels are: ship them
    const uint64_t o = (uint64_t)i * chunk, m = (o + chunk <= n) ? chunk : n - o;
    BB_CUDA(c, cudaStreamWaitEvent(c->s_d2h, c->ev_done[i], 0));
    BB_CUDA(c, cudaMemcpyAsync(c->h_nchanges + i, c->d_chunk_total + i, 8, cudaMemcpyDeviceToHost, c->s_d2h));
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

// the three pack launches; `tiles_buf` holds [tiles][world] words
static cudaError_t launch_pack(uint32_t world, uint32_t key_bits, const bb_batch* in, bb_batch* out, uint64_t* counts,
                               uint32_t* tiles_buf, cudaStream_t s) {
  using namespace bb;
  const uint64_t n = in->n;
  const uint32_t tiles = div_up(n, RT_THREADS);
  bb_launch(k_route_count, tiles, RT_THREADS, 0, s, false, in->path_id, n, world, key_bits, tiles_buf);
  bb_launch(k_route_scan, 1, RS_THREADS, 0, s, false, tiles_buf, tiles, world, counts);
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
  a.key_bits = key_bits;
  a.tile_off = tiles_buf;
  bb_launch(k_route_scatter, tiles, RT_THREADS, 0, s, false, a);
  return cudaGetLastError();
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
  BB_CUDA(c, c->route_tiles.ensure((size_t)div_up(n, RT_THREADS) * world));
  c->launches += 3;
  BB_CUDA(c, launch_pack(world, 0, in, out, counts, c->route_tiles.p, s));
  return BB_OK;
}

int bb_index_create_fields(bb_ctx* c, uint32_t field_mask, uint64_t extra_capacity) {
  if (!c) return BB_ERR_ARG;
  if (field_mask == 0 || (field_mask >> c->cfg.n_fields)) return fail(c, BB_ERR_ARG, "field slot out of range");
  const uint32_t todo = field_mask & ~c->index_mask;  // creating an index that exists is a no-op (query:33-35)
  if (!todo) return BB_OK;
  if (!(c->cfg.flags & BB_CFG_POST_GETDATA))
    return fail(c, BB_ERR_STATE, "indices need a ctx created with BB_CFG_POST_GETDATA");
  if (extra_capacity >= (1ull << 31)) return fail(c, BB_ERR_ARG, "extra_capacity too large");
  BB_CUDA(c, cudaSetDevice(c->cfg.device));
  uint64_t slots = 1024;
  while (slots - (slots >> 3) <= extra_capacity + 1) slots <<= 1;
  const uint64_t cap2 = (c->cfg.capacity + 1) & ~1ull;
  for (int f = 0; f < BB_MAX_FIELDS; ++f) {
    if (!((todo >> f) & 1u)) continue;
    bb_ctx::IndexDev& ix = c->index[f];
    if (cudaMalloc((void**)&ix.pcol, cap2 * sizeof(uint64_t)) != cudaSuccess ||
        cudaMalloc((void**)&ix.xkey, slots * sizeof(uint64_t)) != cudaSuccess ||
        cudaMalloc((void**)&ix.xnode, slots * sizeof(uint32_t)) != cudaSuccess) {
      for (int g = 0; g <= f; ++g) {
        if (!((todo >> g) & 1u)) continue;
        bb_ctx::IndexDev& iy = c->index[g];
        if (iy.pcol) cudaFree(iy.pcol);
        if (iy.xkey) cudaFree(iy.xkey);
        if (iy.xnode) cudaFree(iy.xnode);
        iy = bb_ctx::IndexDev();
      }
      return fail(c, BB_ERR_CUDA, "index allocation failed", cudaGetLastError());
    }
    ix.xslots = slots;
    if (exact_order(c)) {  // entry tags + bucket table (every entry could be a bucket of its own)
      ix.bslots = pow2_at_least(2 * (c->cfg.capacity + extra_capacity) + 16);
      if (cudaMalloc((void**)&ix.pseq, cap2 * sizeof(uint64_t)) != cudaSuccess ||
          cudaMalloc((void**)&ix.xseq, slots * sizeof(uint64_t)) != cudaSuccess ||
          cudaMalloc((void**)&ix.bkey, ix.bslots * sizeof(uint64_t)) != cudaSuccess ||
          cudaMalloc((void**)&ix.bcount, ix.bslots * sizeof(uint32_t)) != cudaSuccess ||
          cudaMalloc((void**)&ix.bcreated, ix.bslots * sizeof(uint64_t)) != cudaSuccess)
        return fail(c, BB_ERR_CUDA, "index allocation failed (exact order)", cudaGetLastError());
    }
  }
  begin_call(c);
  mark(c, EV_Q0, c->stream);
  int rc = index_fill(c, todo, c->stream);
  if (rc) return rc;
  mark(c, EV_Q1, c->stream);
  BB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int f = 0; f < BB_MAX_FIELDS; ++f)
    if ((todo >> f) & 1u) c->index[f].live = true;
  c->index_mask |= todo;
  return BB_OK;
}

int bb_index_create(bb_ctx* c, uint32_t field, uint64_t extra_capacity) {
  if (!c) return BB_ERR_ARG;
  if (field >= c->cfg.n_fields) return fail(c, BB_ERR_ARG, "field slot out of range");
  return bb_index_create_fields(c, 1u << field, extra_capacity);
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

int bb_phase_events(bb_ctx* c, int on) {
  if (!c) return BB_ERR_ARG;
  c->phase_events = on != 0;
  return BB_OK;
}

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

/* ======================================================================== *
 * bb_router: update routing between the shards of one box.
 * NCCL is reached through dlopen so that single-GPU users carry no dependency.
 * ======================================================================== */
namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
thread_local std::string g_router_error;

bool nccl_load() {
  if (g_nccl.lib) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host process already uses, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW);
  if (!h) return false;
  NcclApi a;
  a.lib = h;
#define BB_SYM(field, name) *(void**)(&a.field) = dlsym(h, name)
  BB_SYM(GetUniqueId, "ncclGetUniqueId");
  BB_SYM(CommInitRank, "ncclCommInitRank");
  BB_SYM(CommDestroy, "ncclCommDestroy");
  BB_SYM(AllGather, "ncclAllGather");
  BB_SYM(Send, "ncclSend");
  BB_SYM(Recv, "ncclRecv");
  BB_SYM(GroupStart, "ncclGroupStart");
  BB_SYM(GroupEnd, "ncclGroupEnd");
  BB_SYM(GetErrorString, "ncclGetErrorString");
#undef BB_SYM
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.Send || !a.Recv || !a.GroupStart ||
      !a.GroupEnd || !a.GetErrorString)
    return false;
  g_nccl = a;
  return true;
}

constexpr size_t ROUTE_W[4] = {8, 16, 32, 32};  // bytes per update of path / head / clk / val

}  // namespace

struct bb_router {
  int device = 0;
  uint32_t world = 0, rank = 0;
  uint32_t key_bits = 0;  // sharding function, bb_router_set_sharding
  uint64_t max_batch = 0, cap = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  char* send[4]{};
  char* recv[2][4]{};
  uint64_t n_recv[2]{};
  cudaEvent_t ready[2]{}, merged[2]{}, ev_in = nullptr;
  bool p2p = false;               // peers' receive slots are mapped here: rows are stored directly
  char* peer[bb::RT_MAX_WORLD][2][4]{};  // [rank][slot][array]; our own entry = recv
  uint64_t* d_bar = nullptr;      // [1 + world] barrier token + gather target
  uint64_t* d_counts = nullptr;   // [world] this rank's send counts
  uint64_t* d_matrix = nullptr;   // [world][world] everybody's
  uint64_t* h_matrix = nullptr;   // pinned [2][world][world]: one copy per receive slot
  cudaEvent_t counts_ev[2]{};     // the slot's copy of the matrix has landed on the host
  bool pending[2]{};              // n_recv[slot] still to be derived from it (bb_router_acquire)
  uint64_t routed_n[2]{};
  uint32_t* tiles = nullptr;
  uint64_t sent_bytes = 0, launches = 0;
  uint32_t scatter_ctas = 192;  // grid of the fused pack + exchange kernel (env BB_ROUTE_CTAS): 256-thread CTAs, 45 KB each
  bool bulk = true;            // runs leave with cp.async.bulk (env BB_ROUTE_BULK=0: per-thread stores)
  bool flag_sync = true;       // counts / completion through peer-mapped flags (env BB_ROUTE_NCCL_SYNC=1: NCCL all-gathers)
  bb::RouteCtl* ctl = nullptr; // this rank's control block (peer-mapped by everybody)
  bb::RouteCtlPeers ctl_peers{};
  uint64_t epoch = 0;
  // sharded queries: local hits, peer-mapped result buffers, counts for the host
  uint32_t* q_local = nullptr;
  uint64_t q_local_cap = 0;
  uint32_t* q_result = nullptr;            // this rank's copy of the gathered result
  uint32_t* q_peer[bb::RT_MAX_WORLD]{};    // everybody's; our own entry = q_result
  uint64_t q_cap = 0;
  uint64_t* d_qcounts = nullptr;           // [world + 1]
  uint64_t* h_qcounts = nullptr;           // pinned
  uint64_t q_epoch = 0;
  // two-stream route (flag_sync): `prep` counts + publishes the next batch while `stream` still exchanges
  cudaStream_t prep = nullptr;
  uint32_t* tiles2[2]{};          // per-slot tile offsets (the scatter of one batch and the count of the next overlap)
  uint64_t* d_counts2[2]{};
  cudaEvent_t prep_done[2]{};
  cudaEvent_t tev[5]{};   // telemetry of the last route: start, packed, counts known, exchanged, own rows copied
  double host_ms[2]{};    // host time of the last route: until the counts are known, whole call
  std::string err;
};

namespace {

int rfail(bb_router* r, int code, const std::string& what) {
  if (r) r->err = what;
  else g_router_error = what;
  return code;
}

#define BB_RCUDA(r, call)                                                                           \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return rfail((r), BB_ERR_CUDA, std::string(#call ": ") + cudaGetErrorString(e_)); \
  } while (0)
#define BB_RNCCL(r, call)                                                                            \
  do {                                                                                              \
    ncclResult_t e_ = (call);                                                                       \
    if (e_ != ncclSuccess) return rfail((r), BB_ERR_CUDA, std::string(#call ": ") + g_nccl.GetErrorString(e_)); \
  } while (0)

}  // namespace

namespace {

// Exchange cudaIpc handles of the receive slots over the communicator and map every peer's slots.
void router_map_peers(bb_router* r) {
  const uint32_t W = r->world, me = r->rank;
  for (int sl = 0; sl < 2; ++sl)
    for (int k = 0; k < 4; ++k) r->peer[me][sl][k] = r->recv[sl][k];
  r->p2p = false;
  if (getenv("BB_ROUTER_NO_P2P")) return;
  r->ctl_peers.ctl[me] = r->ctl;
  if (W == 1) {
    r->p2p = true;
    return;
  }
  constexpr size_t HB = sizeof(cudaIpcMemHandle_t);
  const size_t mine = 9 * HB;  // 2 slots x 4 arrays + the control block
  std::string host(mine * W, '\0');
  char* d_all = nullptr;
  bool ok = cudaMalloc((void**)&d_all, mine * W) == cudaSuccess;
  for (int sl = 0; ok && sl < 2; ++sl)
    for (int k = 0; ok && k < 4; ++k) {
      cudaIpcMemHandle_t h;
      ok = cudaIpcGetMemHandle(&h, r->recv[sl][k]) == cudaSuccess;
      memcpy(&host[mine * me + (sl * 4 + k) * HB], &h, HB);
    }
  if (ok) {
    cudaIpcMemHandle_t h;
    ok = cudaIpcGetMemHandle(&h, r->ctl) == cudaSuccess;
    memcpy(&host[mine * me + 8 * HB], &h, HB);
  }
  r->ctl_peers.ctl[me] = r->ctl;
  ok = ok && cudaMemcpyAsync(d_all + mine * me, &host[mine * me], mine, cudaMemcpyHostToDevice, r->stream) == cudaSuccess;
  // every rank must take part in the collective even if its own handles failed
  const bool sent = g_nccl.AllGather(d_all ? d_all + mine * me : nullptr, d_all, mine, ncclUint8, r->comm, r->stream) == ncclSuccess;
  ok = ok && sent && cudaMemcpyAsync(&host[0], d_all, mine * W, cudaMemcpyDeviceToHost, r->stream) == cudaSuccess &&
       cudaStreamSynchronize(r->stream) == cudaSuccess;
  for (uint32_t q = 0; ok && q < W; ++q) {
    if (q == me) continue;
    for (int j = 0; ok && j < 9; ++j) {
      cudaIpcMemHandle_t h;
      memcpy(&h, &host[mine * q + j * HB], HB);
      void* p = nullptr;
      ok = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      if (j < 8) r->peer[q][j / 4][j % 4] = (char*)p;
      else r->ctl_peers.ctl[q] = reinterpret_cast<bb::RouteCtl*>(p);
    }
  }
  if (d_all) cudaFree(d_all);
  cudaGetLastError();
  // all ranks must agree, or one would wait for rows that are sent the other way
  uint64_t flag = ok ? 1 : 0;
  if (cudaMemcpyAsync(r->d_bar, &flag, 8, cudaMemcpyHostToDevice, r->stream) == cudaSuccess &&
      g_nccl.AllGather(r->d_bar, r->d_bar + 1, 1, ncclUint64, r->comm, r->stream) == ncclSuccess) {
    uint64_t flags[bb::RT_MAX_WORLD] = {0};
    if (cudaMemcpyAsync(flags, r->d_bar + 1, W * 8, cudaMemcpyDeviceToHost, r->stream) == cudaSuccess &&
        cudaStreamSynchronize(r->stream) == cudaSuccess) {
      bool all = true;
      for (uint32_t q = 0; q < W; ++q) all = all && flags[q] == 1;
      r->p2p = all;
    }
  }
}

}  // namespace

extern "C" {

int bb_router_unique_id(char id[BB_NCCL_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) <= BB_NCCL_ID_BYTES, "ncclUniqueId fits the id buffer");
  if (!id) return rfail(nullptr, BB_ERR_ARG, "null argument");
  if (!nccl_load()) return rfail(nullptr, BB_ERR_CUDA, "libnccl.so.2 not found");
  ncclUniqueId u;
  BB_RNCCL(nullptr, g_nccl.GetUniqueId(&u));
  memset(id, 0, BB_NCCL_ID_BYTES);
  memcpy(id, &u, sizeof(u));
  return BB_OK;
}

const char* bb_router_last_error(const bb_router* r) { return r ? r->err.c_str() : g_router_error.c_str(); }

int bb_router_destroy(bb_router* r) {
  if (!r) return BB_ERR_ARG;
  cudaSetDevice(r->device);
  if (r->stream) cudaStreamSynchronize(r->stream);
  if (r->comm) g_nccl.CommDestroy(r->comm);
  for (int k = 0; k < 4; ++k) {
    if (r->send[k]) cudaFree(r->send[k]);
    for (int sl = 0; sl < 2; ++sl)
      if (r->recv[sl][k]) cudaFree(r->recv[sl][k]);
  }
  for (int sl = 0; sl < 2; ++sl) {
    if (r->ready[sl]) cudaEventDestroy(r->ready[sl]);
    if (r->merged[sl]) cudaEventDestroy(r->merged[sl]);
  }
  if (r->ev_in) cudaEventDestroy(r->ev_in);
  for (int i = 0; i < 5; ++i)
    if (r->tev[i]) cudaEventDestroy(r->tev[i]);
  for (uint32_t q = 0; q < r->world; ++q)
    for (int j = 0; j < 8; ++j)
      if (q != r->rank && r->peer[q][j / 4][j % 4]) cudaIpcCloseMemHandle(r->peer[q][j / 4][j % 4]);
  for (uint32_t q = 0; q < r->world; ++q)
    if (q != r->rank && r->q_peer[q]) cudaIpcCloseMemHandle(r->q_peer[q]);
  if (r->q_result) cudaFree(r->q_result);
  if (r->q_local) cudaFree(r->q_local);
  if (r->d_qcounts) cudaFree(r->d_qcounts);
  if (r->h_qcounts) cudaFreeHost(r->h_qcounts);
  if (r->d_bar) cudaFree(r->d_bar);
  if (r->d_counts) cudaFree(r->d_counts);
  if (r->d_matrix) cudaFree(r->d_matrix);
  if (r->ctl) cudaFree(r->ctl);
  if (r->prep) cudaStreamDestroy(r->prep);
  for (int i = 0; i < 2; ++i) {
    if (r->tiles2[i]) cudaFree(r->tiles2[i]);
    if (r->d_counts2[i]) cudaFree(r->d_counts2[i]);
    if (r->prep_done[i]) cudaEventDestroy(r->prep_done[i]);
  }
  if (r->h_matrix) cudaFreeHost(r->h_matrix);
  for (int i = 0; i < 2; ++i)
    if (r->counts_ev[i]) cudaEventDestroy(r->counts_ev[i]);
  if (r->tiles) cudaFree(r->tiles);
  if (r->stream) cudaStreamDestroy(r->stream);
  delete r;
  return BB_OK;
}

int bb_router_create(int32_t device, uint32_t world, uint32_t rank, const char id[BB_NCCL_ID_BYTES],
                     uint64_t max_batch, uint64_t recv_capacity, bb_router** out) {
  if (!out || !id) return rfail(nullptr, BB_ERR_ARG, "null argument");
  *out = nullptr;
  if (world < 1 || world > bb::RT_MAX_WORLD || rank >= world || max_batch == 0 || max_batch >= 0xFFFFFFFFull)
    return rfail(nullptr, BB_ERR_ARG, "bad world / rank / max_batch");
  if (world > 1 && !nccl_load()) return rfail(nullptr, BB_ERR_CUDA, "libnccl.so.2 not found");  // one rank: no communicator
  if (cudaSetDevice(device) != cudaSuccess) return rfail(nullptr, BB_ERR_CUDA, "cudaSetDevice failed");
  bb_router* r = new (std::nothrow) bb_router();
  if (!r) return rfail(nullptr, BB_ERR_ARG, "out of host memory");
  r->device = device;
  if (const char* e = getenv("BB_ROUTE_CTAS")) {
    const long v = strtol(e, nullptr, 10);
    if (v > 0 && v < 65536) r->scatter_ctas = (uint32_t)v;
  }
  if (const char* e = getenv("BB_ROUTE_BULK")) r->bulk = e[0] != '0';
  if (const char* e = getenv("BB_ROUTE_NCCL_SYNC")) r->flag_sync = e[0] == '0';
  r->world = world;
  r->rank = rank;
  r->max_batch = max_batch;
  r->cap = recv_capacity ? recv_capacity : max_batch * world;  // a rank may own every update of every batch
  int prio_lo = 0, prio_hi = 0;  // routing runs beside a merge that fills the GPU: let its small kernels in first
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (const char* e = getenv("BB_ROUTE_PRIO")) {  // 0: the router's streams at default priority
    if (e[0] == '0') prio_hi = 0;
  }
  bool ok = cudaStreamCreateWithPriority(&r->stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
            cudaEventCreateWithFlags(&r->ev_in, cudaEventDisableTiming) == cudaSuccess &&
            cudaMalloc((void**)&r->d_counts, world * sizeof(uint64_t)) == cudaSuccess &&
            cudaMalloc((void**)&r->d_bar, (1 + world) * sizeof(uint64_t)) == cudaSuccess &&
            cudaMalloc((void**)&r->d_matrix, (size_t)world * world * sizeof(uint64_t)) == cudaSuccess &&
            cudaMalloc((void**)&r->ctl, sizeof(bb::RouteCtl)) == cudaSuccess &&
            cudaStreamCreateWithPriority(&r->prep, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
            cudaMemset(r->ctl, 0, sizeof(bb::RouteCtl)) == cudaSuccess &&
            cudaMallocHost((void**)&r->h_matrix, 2 * (size_t)world * world * sizeof(uint64_t)) == cudaSuccess &&
            cudaEventCreateWithFlags(&r->counts_ev[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&r->counts_ev[1], cudaEventDisableTiming) == cudaSuccess &&
            cudaMalloc((void**)&r->tiles, (size_t)div_up(max_batch, bb::RT_THREADS) * world * sizeof(uint32_t)) == cudaSuccess;
  for (int k = 0; ok && k < 4; ++k) {
    ok = cudaMalloc((void**)&r->send[k], max_batch * ROUTE_W[k]) == cudaSuccess;
    for (int sl = 0; ok && sl < 2; ++sl) ok = cudaMalloc((void**)&r->recv[sl][k], r->cap * ROUTE_W[k]) == cudaSuccess;
  }
  for (int i = 0; ok && i < 5; ++i) ok = cudaEventCreate(&r->tev[i]) == cudaSuccess;
  for (int sl = 0; ok && sl < 2; ++sl)
    ok = cudaMalloc((void**)&r->tiles2[sl], (size_t)div_up(max_batch, bb::RT_THREADS) * world * sizeof(uint32_t)) == cudaSuccess &&
         cudaMalloc((void**)&r->d_counts2[sl], world * sizeof(uint64_t)) == cudaSuccess &&
         cudaEventCreateWithFlags(&r->prep_done[sl], cudaEventDisableTiming) == cudaSuccess;
  for (int sl = 0; ok && sl < 2; ++sl)
    ok = cudaEventCreateWithFlags(&r->ready[sl], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&r->merged[sl], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    g_router_error = std::string("router allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    bb_router_destroy(r);
    return BB_ERR_CUDA;
  }
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclResult_t e = world > 1 ? g_nccl.CommInitRank(&r->comm, (int)world, u, (int)rank) : ncclSuccess;
  if (e != ncclSuccess) {
    g_router_error = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(e);
    r->comm = nullptr;
    bb_router_destroy(r);
    return BB_ERR_CUDA;
  }
  cudaFuncSetAttribute(bb::k_route_scatter_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, bb::RT_SMEM);
  router_map_peers(r);  // falls back to ncclSend/ncclRecv when the slots cannot be mapped
  *out = r;
  return BB_OK;
}

int bb_router_route_dev(bb_router* r, const bb_batch* in, uint32_t slot, uint64_t* n_recv, void* in_stream) {
  if (!r || !in || !n_recv || slot > 1) return rfail(r, BB_ERR_ARG, "bad argument");
  const uint64_t n = in->n;
  if (n > r->max_batch) return rfail(r, BB_ERR_CAPACITY, "batch larger than the router's max_batch");
  if (n && (!in->path_id || !in->head || !in->clk || !in->val)) return rfail(r, BB_ERR_ARG, "null buffer");
  BB_RCUDA(r, cudaSetDevice(r->device));
  cudaStream_t s = r->stream;
  const uint32_t W = r->world, me = r->rank;
  const bool flags = r->p2p && r->flag_sync && W > 1;
  // flag-sync mode: counting and publishing this batch run on `prep`, concurrently with the exchange of the
  // previous batch on `stream`; they may start once the previous route on this slot has finished EVERYWHERE
  // (ready[slot] follows its barrier): until then some peer may still read that route's counts
  cudaStream_t sp = flags ? r->prep : s;
  uint32_t* tile_buf = flags ? r->tiles2[slot] : r->tiles;
  uint64_t* cnt_buf = flags ? r->d_counts2[slot] : r->d_counts;
  if (in_stream) {
    BB_RCUDA(r, cudaEventRecord(r->ev_in, (cudaStream_t)in_stream));
    BB_RCUDA(r, cudaStreamWaitEvent(sp, r->ev_in, 0));
    if (flags) BB_RCUDA(r, cudaStreamWaitEvent(s, r->ev_in, 0));
  }
  if (flags) {
    BB_RCUDA(r, cudaStreamWaitEvent(sp, r->ready[slot], 0));
    // Publishing this route's counts is what lets every peer's scatter kernel start storing into OUR copy of the
    // slot: not before our merge of what the slot held has finished reading it (a peer only needs its own
    // merged[slot] and everybody's counts flag, so without this a fast peer would overwrite a slow shard's input).
    BB_RCUDA(r, cudaStreamWaitEvent(sp, r->merged[slot], 0));
  }
  BB_RCUDA(r, cudaStreamWaitEvent(s, r->merged[slot], 0));  // the merge that last read this slot is done
  const auto h0 = std::chrono::steady_clock::now();
  cudaEventRecord(r->tev[0], sp);
  const uint32_t tiles = div_up(n, bb::RT_THREADS);
  if (n && r->p2p) {  // the scatter waits until every rank's counts are known
    bb_launch(bb::k_route_count, tiles, bb::RT_THREADS, 0, sp, false, in->path_id, n, W, r->key_bits, tile_buf);
    bb_launch(bb::k_route_scan, 1, bb::RS_THREADS, 0, sp, false, tile_buf, tiles, W, cnt_buf);
    r->launches += 2;
    BB_RCUDA(r, cudaGetLastError());
  } else if (n) {
    bb_batch packed{n, reinterpret_cast<uint64_t*>(r->send[0]), reinterpret_cast<bb_head*>(r->send[1]),
                    reinterpret_cast<uint32_t*>(r->send[2]), reinterpret_cast<uint64_t*>(r->send[3])};
    BB_RCUDA(r, launch_pack(W, r->key_bits, in, &packed, r->d_counts, r->tiles, s));
    r->launches += 3;
  } else {
    BB_RCUDA(r, cudaMemsetAsync(cnt_buf, 0, W * sizeof(uint64_t), sp));
  }
  cudaEventRecord(r->tev[1], sp);
  // everybody's counts: row q = what rank q sends to each rank
  const uint64_t epoch = ++r->epoch;
  const uint64_t* d_matrix = r->d_matrix;
  if (flags) {
    bb_launch(bb::k_route_publish, 1, bb::RT_MAX_WORLD * bb::RT_MAX_WORLD, 0, sp, false, cnt_buf, r->ctl_peers, me, W, slot, epoch);
    r->launches += 1;
    BB_RCUDA(r, cudaGetLastError());
    d_matrix = r->ctl->matrix[slot];
  } else {
    if (W > 1) BB_RNCCL(r, g_nccl.AllGather(r->d_counts, r->d_matrix, W, ncclUint64, r->comm, s));
    else BB_RCUDA(r, cudaMemcpyAsync(r->d_matrix, r->d_counts, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
  }
  uint64_t* hm = r->h_matrix + (size_t)slot * W * W;
  BB_RCUDA(r, cudaMemcpyAsync(hm, d_matrix, (size_t)W * W * sizeof(uint64_t), cudaMemcpyDeviceToHost, sp));
  if (flags) {
    BB_RCUDA(r, cudaEventRecord(r->counts_ev[slot], sp));
    BB_RCUDA(r, cudaEventRecord(r->prep_done[slot], sp));
    BB_RCUDA(r, cudaStreamWaitEvent(s, r->prep_done[slot], 0));
  }
  cudaEventRecord(r->tev[2], s);
  if (r->p2p) {
    // Fused pack + exchange, fully asynchronous: the scatter kernel reads the all-gathered counts on the
    // device, the host picks its copy up in bb_router_acquire.  Nobody stores into a slot before everyone
    // has passed the counts all-gather above, i.e. before every owner has finished merging what it held.
    if (!flags) BB_RCUDA(r, cudaEventRecord(r->counts_ev[slot], s));
    if (n) {
      bb::RouteP2PArgs a;
      a.path_id = in->path_id;
      a.head = reinterpret_cast<const uint4*>(in->head);
      a.clk = reinterpret_cast<const uint4*>(in->clk);
      a.val = reinterpret_cast<const uint4*>(in->val);
      for (uint32_t q = 0; q < W; ++q) {
        a.d_path[q] = reinterpret_cast<uint64_t*>(r->peer[q][slot][0]);
        a.d_head[q] = reinterpret_cast<uint4*>(r->peer[q][slot][1]);
        a.d_clk[q] = reinterpret_cast<uint4*>(r->peer[q][slot][2]);
        a.d_val[q] = reinterpret_cast<uint4*>(r->peer[q][slot][3]);
      }
      a.matrix = d_matrix;
      a.slot_cap = r->cap;
      a.me = me;
      a.n = n;
      a.world = W;
      a.key_bits = r->key_bits;
      a.bulk = r->bulk ? 1u : 0u;
      a.tile_off = tile_buf;
      bb_launch(bb::k_route_scatter_p2p, std::min<uint32_t>(tiles, r->scatter_ctas), bb::RT_BLOCK, bb::RT_SMEM, s, false, a);
      r->launches += 1;
      BB_RCUDA(r, cudaGetLastError());
    }
    cudaEventRecord(r->tev[3], s);
    // every rank's stores are complete when its scatter kernel is: a barrier across the ranks follows
    if (flags) {
      bb_launch(bb::k_route_barrier, 1, 32, 0, s, false, r->ctl_peers, me, W, slot, epoch);
      r->launches += 1;
      BB_RCUDA(r, cudaGetLastError());
    } else if (W > 1) {
      BB_RNCCL(r, g_nccl.AllGather(r->d_bar, r->d_bar + 1, 1, ncclUint64, r->comm, s));
    }
    cudaEventRecord(r->tev[4], s);
    BB_RCUDA(r, cudaEventRecord(r->ready[slot], s));
    r->host_ms[0] = r->host_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
    r->pending[slot] = true;
    r->routed_n[slot] = n;
    *n_recv = ~0ull;  // not known on the host yet: bb_router_acquire returns it
    return BB_OK;
  }
  BB_RCUDA(r, cudaStreamSynchronize(s));  // NCCL send/recv path: the host needs the counts to post the exchange
  r->host_ms[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
  uint64_t so[bb::RT_MAX_WORLD + 1], ro[bb::RT_MAX_WORLD + 1];
  so[0] = ro[0] = 0;
  for (uint32_t q = 0; q < W; ++q) {
    so[q + 1] = so[q] + hm[(size_t)me * W + q];
    ro[q + 1] = ro[q] + hm[(size_t)q * W + me];
  }
  if (ro[W] > r->cap) return rfail(r, BB_ERR_CAPACITY, "receive slot too small for this batch");
  BB_RNCCL(r, g_nccl.GroupStart());
  for (int k = 0; k < 4; ++k) {
    const size_t w = ROUTE_W[k];
    for (uint32_t q = 0; q < W; ++q) {
      if (q == me) continue;
      if (so[q + 1] > so[q])
        BB_RNCCL(r, g_nccl.Send(r->send[k] + so[q] * w, (so[q + 1] - so[q]) * w, ncclUint8, (int)q, r->comm, s));
      if (ro[q + 1] > ro[q])
        BB_RNCCL(r, g_nccl.Recv(r->recv[slot][k] + ro[q] * w, (ro[q + 1] - ro[q]) * w, ncclUint8, (int)q, r->comm, s));
    }
  }
  BB_RNCCL(r, g_nccl.GroupEnd());
  cudaEventRecord(r->tev[3], s);
  for (int k = 0; k < 4; ++k)  // this rank's own rows never leave the device
    if (so[me + 1] > so[me])
      BB_RCUDA(r, cudaMemcpyAsync(r->recv[slot][k] + ro[me] * ROUTE_W[k], r->send[k] + so[me] * ROUTE_W[k],
                                  (so[me + 1] - so[me]) * ROUTE_W[k], cudaMemcpyDeviceToDevice, s));
  cudaEventRecord(r->tev[4], s);
  BB_RCUDA(r, cudaEventRecord(r->ready[slot], s));
  r->host_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
  r->sent_bytes += (so[W] - (so[me + 1] - so[me])) * 88;
  r->n_recv[slot] = ro[W];
  *n_recv = ro[W];
  return BB_OK;
}

int bb_router_set_sharding(bb_router* r, uint32_t key_bits) {
  if (!r || key_bits > 40) return rfail(r, BB_ERR_ARG, "key_bits must be 0 (id % world) or 1..40");
  r->key_bits = key_bits;
  return BB_OK;
}

int bb_router_acquire(bb_router* r, uint32_t slot, void* stream, bb_batch* received) {
  if (!r || !received || slot > 1 || !stream) return rfail(r, BB_ERR_ARG, "bad argument (an explicit stream is required)");
  BB_RCUDA(r, cudaSetDevice(r->device));
  if (r->pending[slot]) {  // the counts of this slot's route: long on the host by now
    BB_RCUDA(r, cudaEventSynchronize(r->counts_ev[slot]));
    const uint32_t W = r->world, me = r->rank;
    const uint64_t* hm = r->h_matrix + (size_t)slot * W * W;
    uint64_t recv = 0, sent = 0;
    bool fits = true;
    for (uint32_t q = 0; q < W; ++q) {
      recv += hm[(size_t)q * W + me];
      if (q != me) sent += hm[(size_t)me * W + q];
      uint64_t col = 0;
      for (uint32_t p = 0; p < W; ++p) col += hm[(size_t)p * W + q];
      fits = fits && col <= r->cap;
    }
    r->pending[slot] = false;
    if (!fits) return rfail(r, BB_ERR_CAPACITY, "receive slot too small for this batch (nothing was exchanged)");
    r->n_recv[slot] = recv;
    r->sent_bytes += sent * 88;
  }
  BB_RCUDA(r, cudaStreamWaitEvent((cudaStream_t)stream, r->ready[slot], 0));
  received->n = r->n_recv[slot];
  received->path_id = reinterpret_cast<uint64_t*>(r->recv[slot][0]);
  received->head = reinterpret_cast<bb_head*>(r->recv[slot][1]);
  received->clk = reinterpret_cast<uint32_t*>(r->recv[slot][2]);
  received->val = reinterpret_cast<uint64_t*>(r->recv[slot][3]);
  return BB_OK;
}

int bb_router_release(bb_router* r, uint32_t slot, void* stream) {
  if (!r || slot > 1 || !stream) return rfail(r, BB_ERR_ARG, "bad argument (an explicit stream is required)");
  BB_RCUDA(r, cudaSetDevice(r->device));
  BB_RCUDA(r, cudaEventRecord(r->merged[slot], (cudaStream_t)stream));
  return BB_OK;
}

/* ---- host entry of the sharded path ------------------------------------------------------------------------- */
// Collective.  Every rank's batch (HOST buffers, arrival order) is cut into `chunks` pieces; piece j of every rank is
// routed (H2D on the copy-in stream -> pack + NVLink all-to-all into receive slot j % 2), merged into the owning
// shards, and the verdicts + change entries of what THIS shard received go back to the host - H2D of piece j+1, the
// exchange of piece j+1, the merge of piece j and the D2H of piece j-1 all overlap.
int bb_router_merge_batch(bb_router* r, bb_ctx* c, const bb_batch* in, bb_changes* out, uint32_t chunks,
                          uint64_t* n_received, uint64_t* recv_counts) {
  if (!r || !c || !in || !out || !out->n_changes || !out->verdict || !n_received)
    return rfail(r, BB_ERR_ARG, "null argument");
  if (chunks == 0) chunks = 4;
  if (chunks > (uint32_t)MAX_CHUNKS) return rfail(r, BB_ERR_ARG, "at most 16 chunks");
  const uint64_t n = in->n;
  if (n && (!in->path_id || !in->head || !in->clk || !in->val)) return rfail(r, BB_ERR_ARG, "null buffer");
  if (c->cfg.flags & BB_CFG_ORDERED_CHANGES) return rfail(r, BB_ERR_STATE, "the sharded host entry needs the default change-set layout");
  const uint64_t chunk = (n + chunks - 1) / chunks;
  if (chunk > r->max_batch) return rfail(r, BB_ERR_CAPACITY, "a chunk is larger than the router's max_batch");
  BB_RCUDA(r, cudaSetDevice(r->device));
  const uint32_t W = r->world, me = r->rank;
  cudaStream_t s = c->stream;
  begin_call(c);
  ++c->epoch;
  mark(c, EV_H2D0, s);
  {
    int rc = reserve_io(c, std::max<uint64_t>(std::max<uint64_t>(n, out->cap), 1));
    if (rc == BB_OK) rc = reserve_dev(c, r->cap);
    if (rc) return rfail(r, rc, bb_last_error(c));
  }
  BB_RCUDA(r, cudaEventRecord(c->ev_in[0], s));  // the side streams start after what is queued on ours
  BB_RCUDA(r, cudaStreamWaitEvent(c->s_h2d, c->ev_in[0], 0));
  BB_RCUDA(r, cudaStreamWaitEvent(c->s_d2h, c->ev_in[0], 0));
  mark(c, EV_START, s);
  BB_RCUDA(r, cudaMemsetAsync(c->d_nchanges, 0, sizeof(uint64_t), s));
  bb_changes dout{out->cap, c->io_verdict.p, c->d_nchanges, c->io_out_idx.p, reinterpret_cast<bb_head*>(c->io_out_head.p),
                  reinterpret_cast<uint32_t*>(c->io_out_clk.p), reinterpret_cast<uint64_t*>(c->io_out_val.p)};
  auto copy_in_and_route = [&](uint32_t j) -> int {
    const uint64_t o = std::min<uint64_t>((uint64_t)j * chunk, n), m = std::min<uint64_t>(chunk, n - o);
    if (m) {
      BB_RCUDA(r, cudaMemcpyAsync(c->io_path.p + o, in->path_id + o, m * 8, cudaMemcpyHostToDevice, c->s_h2d));
      BB_RCUDA(r, cudaMemcpyAsync(c->io_head.p + o, in->head + o, m * 16, cudaMemcpyHostToDevice, c->s_h2d));
      BB_RCUDA(r, cudaMemcpyAsync(c->io_clk.p + 2 * o, in->clk + 8 * o, m * 32, cudaMemcpyHostToDevice, c->s_h2d));
      BB_RCUDA(r, cudaMemcpyAsync(c->io_val.p + 2 * o, in->val + 4 * o, m * 32, cudaMemcpyHostToDevice, c->s_h2d));
    }
    bb_batch din{m, c->io_path.p + o, reinterpret_cast<const bb_head*>(c->io_head.p + o),
                 reinterpret_cast<const uint32_t*>(c->io_clk.p + 2 * o), reinterpret_cast<const uint64_t*>(c->io_val.p + 2 * o)};
    uint64_t dummy = 0;
    return bb_router_route_dev(r, &din, j & 1u, &dummy, c->s_h2d);
  };
  int rc = copy_in_and_route(0);
  if (rc) return rc;
  uint64_t voff = 0, done = 0;
  for (uint32_t j = 0; j < chunks; ++j) {
    if (j + 1 < chunks) {
      rc = copy_in_and_route(j + 1);
      if (rc) return rc;
    }
    bb_batch rb;
    rc = bb_router_acquire(r, j & 1u, s, &rb);
    if (rc) return rc;
    const uint64_t m = rb.n;
    if (recv_counts) {
      const uint64_t* hm = r->h_matrix + (size_t)(j & 1u) * W * W;
      for (uint32_t q = 0; q < W; ++q) recv_counts[(size_t)j * W + q] = hm[(size_t)q * W + me];
    }
    if (voff + m > out->cap) return rfail(r, BB_ERR_CAPACITY, "change-set buffers too small for what this shard received");
    bb_changes dchunk = dout;
    dchunk.verdict = c->io_verdict.p + voff;
    rc = merge_dev(c, &rb, &dchunk, s, (uint32_t)voff, true, nullptr);
    if (rc) return rfail(r, rc, bb_last_error(c));
    rc = bb_router_release(r, j & 1u, s);
    if (rc) return rc;
    BB_RCUDA(r, cudaMemcpyAsync(c->d_chunk_total + j, c->d_nchanges, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    BB_RCUDA(r, cudaEventRecord(c->ev_done[j], s));
    // ship piece j's verdicts and entries while the next piece is exchanged and merged
    BB_RCUDA(r, cudaStreamWaitEvent(c->s_d2h, c->ev_done[j], 0));
    BB_RCUDA(r, cudaMemcpyAsync(c->h_nchanges + j, c->d_chunk_total + j, 8, cudaMemcpyDeviceToHost, c->s_d2h));
    BB_RCUDA(r, cudaEventRecord(c->ev_cnt[j], c->s_d2h));
    if (m) BB_RCUDA(r, cudaMemcpyAsync(out->verdict + voff, c->io_verdict.p + voff, m * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    voff += m;
    if (j > 0) {  // entries of piece j-1: its count is on the host by now (no stall: piece j is already queued)
      BB_RCUDA(r, cudaEventSynchronize(c->ev_cnt[j - 1]));
      const uint64_t k = c->h_nchanges[j - 1];
      if (k > out->cap) return rfail(r, BB_ERR_CAPACITY, "change-set buffer too small");
      if (k > done) {
        if (!out->idx || !out->head || !out->clk || !out->val) return rfail(r, BB_ERR_ARG, "null buffer");
        const uint64_t cnt = k - done;
        BB_RCUDA(r, cudaMemcpyAsync(out->idx + done, c->io_out_idx.p + done, cnt * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        BB_RCUDA(r, cudaMemcpyAsync(out->head + done, c->io_out_head.p + done, cnt * 16, cudaMemcpyDeviceToHost, c->s_d2h));
        BB_RCUDA(r, cudaMemcpyAsync(out->clk + 8 * done, c->io_out_clk.p + 2 * done, cnt * 32, cudaMemcpyDeviceToHost, c->s_d2h));
        BB_RCUDA(r, cudaMemcpyAsync(out->val + 4 * done, c->io_out_val.p + 2 * done, cnt * 32, cudaMemcpyDeviceToHost, c->s_d2h));
        done = k;
      }
    }
  }
  mark(c, EV_SORT, s);
  mark(c, EV_MERGE, s);
  {
    BB_RCUDA(r, cudaEventSynchronize(c->ev_cnt[chunks - 1]));
    const uint64_t k = c->h_nchanges[chunks - 1];
    if (k > out->cap) return rfail(r, BB_ERR_CAPACITY, "change-set buffer too small");
    if (k > done) {
      if (!out->idx || !out->head || !out->clk || !out->val) return rfail(r, BB_ERR_ARG, "null buffer");
      const uint64_t cnt = k - done;
      BB_RCUDA(r, cudaMemcpyAsync(out->idx + done, c->io_out_idx.p + done, cnt * 4, cudaMemcpyDeviceToHost, c->s_d2h));
      BB_RCUDA(r, cudaMemcpyAsync(out->head + done, c->io_out_head.p + done, cnt * 16, cudaMemcpyDeviceToHost, c->s_d2h));
      BB_RCUDA(r, cudaMemcpyAsync(out->clk + 8 * done, c->io_out_clk.p + 2 * done, cnt * 32, cudaMemcpyDeviceToHost, c->s_d2h));
      BB_RCUDA(r, cudaMemcpyAsync(out->val + 4 * done, c->io_out_val.p + 2 * done, cnt * 32, cudaMemcpyDeviceToHost, c->s_d2h));
      done = k;
    }
  }
  BB_RCUDA(r, cudaEventRecord(c->ev_cnt[0], c->s_d2h));
  BB_RCUDA(r, cudaStreamWaitEvent(s, c->ev_cnt[0], 0));
  mark(c, EV_D2H, s);
  rc = collect_device_error(c, s);  // synchronises everything
  if (rc) return rfail(r, rc, bb_last_error(c));
  *out->n_changes = done;
  *n_received = voff;
  return BB_OK;
}

/* ---- sharded queries ------------------------------------------------------------------------------------ */
int bb_router_query_reserve(bb_router* r, uint64_t max_total_hits) {
  if (!r || max_total_hits == 0) return rfail(r, BB_ERR_ARG, "bad argument");
  if (!r->p2p) return rfail(r, BB_ERR_STATE, "sharded queries need the peers' memory mapped (cudaIpc); BB_ROUTER_NO_P2P is set or mapping failed");
  if (r->q_result) return rfail(r, BB_ERR_STATE, "query buffers already reserved");
  BB_RCUDA(r, cudaSetDevice(r->device));
  const uint32_t W = r->world, me = r->rank;
  BB_RCUDA(r, cudaMalloc((void**)&r->q_local, max_total_hits * sizeof(uint32_t)));  // one rank's hits <= everybody's
  r->q_local_cap = max_total_hits;
  BB_RCUDA(r, cudaMalloc((void**)&r->q_result, max_total_hits * sizeof(uint32_t)));
  r->q_cap = max_total_hits;
  BB_RCUDA(r, cudaMalloc((void**)&r->d_qcounts, (W + 1) * sizeof(uint64_t)));
  BB_RCUDA(r, cudaMallocHost((void**)&r->h_qcounts, (W + 1) * sizeof(uint64_t)));
  r->q_peer[me] = r->q_result;
  if (W == 1) return BB_OK;
  constexpr size_t HB = sizeof(cudaIpcMemHandle_t);
  std::string host(HB * W, '\0');
  char* d_all = nullptr;
  bool ok = cudaMalloc((void**)&d_all, HB * W) == cudaSuccess;
  cudaIpcMemHandle_t h;
  ok = ok && cudaIpcGetMemHandle(&h, r->q_result) == cudaSuccess;
  memcpy(&host[HB * me], &h, HB);
  ok = ok && cudaMemcpyAsync(d_all + HB * me, &host[HB * me], HB, cudaMemcpyHostToDevice, r->stream) == cudaSuccess;
  const bool sent = g_nccl.AllGather(d_all ? d_all + HB * me : nullptr, d_all, HB, ncclUint8, r->comm, r->stream) == ncclSuccess;
  ok = ok && sent && cudaMemcpyAsync(&host[0], d_all, HB * W, cudaMemcpyDeviceToHost, r->stream) == cudaSuccess &&
       cudaStreamSynchronize(r->stream) == cudaSuccess;
  for (uint32_t q = 0; ok && q < W; ++q) {
    if (q == me) continue;
    memcpy(&h, &host[HB * q], HB);
    void* p = nullptr;
    ok = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    r->q_peer[q] = (uint32_t*)p;
  }
  if (d_all) cudaFree(d_all);
  if (!ok) return rfail(r, BB_ERR_CUDA, "mapping the peers' query result buffers failed");
  return BB_OK;
}

static int router_query(bb_router* r, bb_ctx* c, uint32_t field, const bb::Pred& pred, bb_gathered_hits* out, void* stream) {
  if (!r || !c || !out || !stream) return rfail(r, BB_ERR_ARG, "bad argument (an explicit stream is required)");
  if (!r->q_result) return rfail(r, BB_ERR_STATE, "bb_router_query_reserve first");
  BB_RCUDA(r, cudaSetDevice(r->device));
  cudaStream_t s = (cudaStream_t)stream;
  begin_call(c);
  int rc = scan_dev(c, field, pred, r->q_local, r->q_local_cap, s);
  if (rc) return rfail(r, rc, bb_last_error(c));
  bb::QueryPushArgs a;
  a.hits = r->q_local;
  a.n_hits = c->d_counters;
  a.hits_cap = r->q_local_cap;
  for (uint32_t q = 0; q < r->world; ++q) a.dst[q] = r->q_peer[q];
  a.dst_cap = r->q_cap;
  a.peers = r->ctl_peers;
  a.counts_out = r->d_qcounts;
  a.me = r->rank;
  a.world = r->world;
  a.epoch = ++r->q_epoch;
  bb_launch(bb::k_query_publish, 1, 32, 0, s, false, a);
  bb_launch(bb::k_query_push, (uint32_t)(c->n_sm * 4), 256, 0, s, false, a);
  bb_launch(bb::k_query_barrier, 1, 32, 0, s, false, r->ctl_peers, r->rank, r->world, a.epoch);
  r->launches += 3;
  BB_RCUDA(r, cudaGetLastError());
  BB_RCUDA(r, cudaMemcpyAsync(r->h_qcounts, r->d_qcounts, (r->world + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  BB_RCUDA(r, cudaStreamSynchronize(s));
  if (r->h_qcounts[r->world]) return rfail(r, BB_ERR_CAPACITY, "gathered result larger than bb_router_query_reserve's max_total_hits");
  out->node = r->q_result;
  out->offset[0] = 0;
  for (uint32_t q = 0; q < r->world; ++q) out->offset[q + 1] = out->offset[q] + r->h_qcounts[q];
  out->total = out->offset[r->world];
  return BB_OK;
}

int bb_router_query_range(bb_router* r, bb_ctx* ctx, uint32_t field, const bb_bound* lo, const bb_bound* hi,
                          bb_gathered_hits* out, void* stream) {
  if (!lo || !hi) return rfail(r, BB_ERR_ARG, "null argument");
  return router_query(r, ctx, field, range_pred(lo, hi), out, stream);
}

int bb_router_query_equals(bb_router* r, bb_ctx* ctx, uint32_t field, uint64_t key, bb_gathered_hits* out, void* stream) {
  bb::Pred p{};
  p.mode = 0;
  p.eq = key;
  return router_query(r, ctx, field, p, out, stream);
}

int bb_router_query_fetch(bb_router* r, uint64_t first, uint64_t n, uint32_t* host_out) {
  if (!r || !r->q_result || (n && !host_out) || first + n > r->q_cap) return rfail(r, BB_ERR_ARG, "bad argument");
  BB_RCUDA(r, cudaSetDevice(r->device));
  if (n) BB_RCUDA(r, cudaMemcpy(host_out, r->q_result + first, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return BB_OK;
}

int bb_router_last_ms(bb_router* r, double out[6]) {
  if (!r || !out) return BB_ERR_ARG;
  cudaSetDevice(r->device);
  if (cudaEventSynchronize(r->tev[4]) != cudaSuccess) return BB_ERR_STATE;
  for (int i = 0; i < 4; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r->tev[i], r->tev[i + 1]);
    out[i] = ms;
  }
  out[4] = r->host_ms[0];
  out[5] = r->host_ms[1];
  return BB_OK;
}

uint64_t bb_router_sent_bytes(const bb_router* r) { return r ? r->sent_bytes : 0; }
uint64_t bb_router_launch_count(const bb_router* r) { return r ? r->launches : 0; }

}  // extern "C"
