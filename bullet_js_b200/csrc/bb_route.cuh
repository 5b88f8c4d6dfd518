// bb_route.cuh - sharded runs: stable partition of a batch by owner rank and the fused
// pack + all-to-all over NVLink peer memory, with flag-based signalling between the ranks.
#pragma once
#include "bb_kernels.cuh"

namespace bb {

// ---------------------------------------------------------------- K6: shard routing pack
// A sharded table: path id p lives on rank p % world as local row p / world.  The pack
// is a STABLE partition of the batch by owner - updates for rank 0 first, arrival order kept
// inside every destination - so that the owner, which concatenates what it receives in source-rank
// order, replays each path in (source rank, arrival index) order.  Three small launches: per-tile
// destination counts, one CTA scanning them (tile-major inside destination-major), the scatter.
// Owner and local row of a path id.  key_bits == 0: owner = id % world, row = id / world (ids that are dense and
// evenly spread).  Otherwise the id (< 2^key_bits) first goes through a splitmix64-style finaliser restricted to
// key_bits bits - odd multiplications and xor-shifts, each a bijection of [0, 2^key_bits) - so that strided or
// clustered ids still spread evenly, and the scrambled id is split the same way; a shard then holds
// ceil(2^key_bits / world) rows.  bullet_js_b200/shard.py mirrors it for the host.
__host__ __device__ __forceinline__ uint64_t shard_mix(uint64_t id, uint32_t key_bits) {
  if (key_bits == 0) return id;
  const uint64_t mask = key_bits >= 64 ? ~0ull : ((1ull << key_bits) - 1ull);
  const uint32_t s = (key_bits + 1) / 2;
  uint64_t x = id & mask;
  x = (x * 0x9E3779B97F4A7C15ull) & mask;
  x ^= x >> s;
  x = (x * 0xBF58476D1CE4E5B9ull) & mask;
  x ^= x >> s;
  x = (x * 0x94D049BB133111EBull) & mask;
  x ^= x >> s;
  return x;
}

constexpr int RT_THREADS = 512;    // updates per tile of the count / scatter kernels
constexpr int RT_BLOCK = 256;      // threads of a k_route_scatter_p2p CTA: RT_RPT rows of the tile each.  The kernel runs BESIDE
constexpr int RT_RPT = RT_THREADS / RT_BLOCK;  // the merge: a 1024-thread CTA with 88 KB of tile took 6 of an SM's 7 merge CTAs
constexpr int RS_THREADS = 256;    // threads of the single scan CTA
constexpr int RT_MAX_WORLD = 16;

__global__ void __launch_bounds__(RT_THREADS) k_route_count(const uint64_t* __restrict__ path_id, uint64_t n,
                                                            uint32_t world, uint32_t key_bits,
                                                            uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t s_cnt[RT_MAX_WORLD];
  if (threadIdx.x < RT_MAX_WORLD) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * RT_THREADS + threadIdx.x;
  const uint32_t d = i < n ? (uint32_t)(shard_mix(path_id[i], key_bits) % world) : world;
  for (uint32_t r = 0; r < world; ++r) {
    const uint32_t m = __ballot_sync(0xffffffffu, d == r);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt[r], __popc(m));
  }
  __syncthreads();
  if (threadIdx.x < world) tile_cnt[(uint64_t)blockIdx.x * world + threadIdx.x] = s_cnt[threadIdx.x];
}

// one CTA: tile_cnt[tile][r] -> exclusive offsets in the packed order; counts[r] = updates for rank r
__global__ void __launch_bounds__(RS_THREADS) k_route_scan(uint32_t* __restrict__ tile_cnt, uint32_t tiles,
                                                           uint32_t world, uint64_t* __restrict__ counts) {
  __shared__ uint32_t s_run;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (uint32_t r = 0; r < world; ++r) {
    const uint32_t start = s_run;
    for (uint32_t t0 = 0; t0 < tiles; t0 += RS_THREADS) {
      const uint32_t t = t0 + threadIdx.x;
      const uint32_t v = t < tiles ? tile_cnt[(uint64_t)t * world + r] : 0;
      uint32_t total;
      const uint32_t ex = block_exclusive_scan<RS_THREADS>(v, &total);
      if (t < tiles) tile_cnt[(uint64_t)t * world + r] = s_run + ex;
      __syncthreads();
      if (threadIdx.x == 0) s_run += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) counts[r] = s_run - start;
    __syncthreads();
  }
}

struct RouteArgs {
  const uint64_t* path_id; const uint4* head; const uint4* clk; const uint4* val;  // [n] in
  uint64_t* o_path; uint4* o_head; uint4* o_clk; uint4* o_val;                      // [n] packed out
  uint64_t n;
  uint32_t world;
  uint32_t key_bits;
  const uint32_t* tile_off;  // [tiles][world] from k_route_scan
};

// Fused pack + all-to-all: the same stable partition, but every row is stored straight into the
// receive slot of its owner - peer memory mapped over NVLink (cudaIpc) - at the place the owner's
// concatenation in source-rank order gives it.  No send buffer, no separate exchange launch.
struct RouteP2PArgs {
  const uint64_t* path_id; const uint4* head; const uint4* clk; const uint4* val;  // [n] in
  uint64_t* d_path[RT_MAX_WORLD]; uint4* d_head[RT_MAX_WORLD]; uint4* d_clk[RT_MAX_WORLD]; uint4* d_val[RT_MAX_WORLD];
  const uint64_t* matrix;     // [world][world] on the device: row p = what rank p sends to each rank
  uint64_t slot_cap;          // rows a receive slot holds: nothing is stored if some owner would overflow
  uint32_t me;
  uint64_t n;
  uint32_t world;
  uint32_t key_bits;
  uint32_t bulk;  // 1: runs leave with cp.async.bulk (default); 0: with per-thread 16-byte stores
  const uint32_t* tile_off;
};

// Each CTA partitions its 1024 rows by owner in shared memory (88 KB), then streams every owner's
// run out with fully coalesced 16-byte stores: one contiguous run per (tile, owner, array).
// ---- signalling between the ranks' routers through peer-mapped memory (one RouteCtl per rank)
// The two tiny collectives of a route (everybody's counts before the scatter, "all my stores have
// landed" after it) cost 25-40 us each as NCCL all-gathers at 8 GPUs, on the critical path of the step.
// Here a rank stores its words straight into every peer's control block, then a flag (after a
// system-scope fence), and spins on its own flags: a few microseconds.  Flags carry the route's
// epoch and only grow, so a peer that is already one route ahead never confuses a waiter.
struct RouteCtl {
  uint64_t matrix[2][RT_MAX_WORLD * RT_MAX_WORLD];  // [slot][source rank][destination rank]
  uint64_t cflag[2][RT_MAX_WORLD];                  // [slot][source]: that source's counts row is in
  uint64_t bflag[2][RT_MAX_WORLD];                  // [slot][source]: that source's rows have landed
  uint64_t err;
  // sharded queries: every rank's hit count, then "my hit ids have landed everywhere"
  uint64_t qcount[RT_MAX_WORLD];
  uint64_t qcflag[RT_MAX_WORLD];
  uint64_t qbflag[RT_MAX_WORLD];
};

struct RouteCtlPeers {
  RouteCtl* ctl[RT_MAX_WORLD];  // [rank]; our own entry is local memory
};

constexpr long long RT_SPIN_LIMIT = 120000000000ll;  // ~60 s of SM clocks: a peer died; fail loudly instead of hanging

__device__ __forceinline__ void spin_until(const uint64_t* flag, uint64_t epoch, uint64_t* err) {
  const long long t0 = clock64();
  while (ld_sys(flag) < epoch) {
    if (clock64() - t0 > RT_SPIN_LIMIT) {
      st_sys(err, 1);
      __trap();
    }
    __nanosleep(64);
    BB_SPIN_YIELD();
  }
}

// counts[W] of this rank -> row `me` of everybody's matrix, then wait for everybody's row
__global__ void __launch_bounds__(RT_MAX_WORLD * RT_MAX_WORLD) k_route_publish(const uint64_t* __restrict__ counts,
                                                                                RouteCtlPeers peers, uint32_t me, uint32_t world,
                                                                                uint32_t slot, uint64_t epoch) {
  const uint32_t t = threadIdx.x;
  if (t < world * world) {
    const uint32_t q = t / world, j = t % world;
    st_sys(&peers.ctl[q]->matrix[slot][(uint64_t)me * world + j], counts[j]);
  }
  __threadfence_system();
  __syncthreads();
  if (t < world) {
    __threadfence_system();
    st_sys(&peers.ctl[t]->cflag[slot][me], epoch);
    spin_until(&peers.ctl[me]->cflag[slot][t], epoch, &peers.ctl[me]->err);
  }
  __threadfence_system();
}

// "every store of my scatter kernel has landed" to everybody, then wait for everybody's
__global__ void __launch_bounds__(32) k_route_barrier(RouteCtlPeers peers, uint32_t me, uint32_t world, uint32_t slot,
                                                      uint64_t epoch) {
  const uint32_t t = threadIdx.x;
  if (t < world) {
    __threadfence_system();  // cumulative: the previous kernel's stores (visible to this thread) go first
    st_sys(&peers.ctl[t]->bflag[slot][me], epoch);
    spin_until(&peers.ctl[me]->bflag[slot][t], epoch, &peers.ctl[me]->err);
  }
  __threadfence_system();
}

// ---------------------------------------------------------------- sharded queries: all-gather(v) of the hit ids
// through peer memory.  Every rank has scanned its shard (k_index_scan) into a local buffer; the counts go to
// everybody's control block, then ONE kernel per rank stores its hit ids straight into every rank's result buffer
// (peer-mapped, NVLink) at the offset the counts give it - rank-major, each rank's hits as one run - and an
// epoch-flag barrier tells the readers that all runs are in.
struct QueryPushArgs {
  const uint32_t* hits;              // this rank's local hit ids
  const unsigned long long* n_hits;  // [2] dense + overflow matches of the local scan (device)
  uint64_t hits_cap;                 // elements of `hits`: a scan that matched more publishes a poisoned count
  uint32_t* dst[RT_MAX_WORLD];       // every rank's result buffer
  uint64_t dst_cap;                  // elements
  RouteCtlPeers peers;
  uint64_t* counts_out;              // [world + 1] local copy of everybody's count + overflow marker, for the host
  uint32_t me, world;
  uint64_t epoch;
};

__global__ void __launch_bounds__(32) k_query_publish(QueryPushArgs a) {
  const uint32_t t = threadIdx.x;
  uint64_t mine = a.n_hits[0] + a.n_hits[1];
  if (mine > a.hits_cap) mine = 1ull << 62;  // more matches than the local buffer holds: every rank reports BB_ERR_CAPACITY
  if (t < a.world) {
    st_sys(&a.peers.ctl[t]->qcount[a.me], mine);
    __threadfence_system();
    st_sys(&a.peers.ctl[t]->qcflag[a.me], a.epoch);
    spin_until(&a.peers.ctl[a.me]->qcflag[t], a.epoch, &a.peers.ctl[a.me]->err);
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(256) k_query_push(QueryPushArgs a) {
  __shared__ uint64_t s_off, s_n, s_total;
  if (threadIdx.x == 0) {
    uint64_t off = 0, total = 0;
    for (uint32_t q = 0; q < a.world; ++q) {
      const uint64_t c = ld_sys(&a.peers.ctl[a.me]->qcount[q]);
      if (q < a.me) off += c;
      total += c;
      if (blockIdx.x == 0) a.counts_out[q] = c;  // (private copy: a fast peer may publish its next query's count)
    }
    if (blockIdx.x == 0) a.counts_out[a.world] = total > a.dst_cap ? 1ull : 0ull;
    s_off = off;
    s_n = a.n_hits[0] + a.n_hits[1];  // (== qcount[me] whenever total fits)
    s_total = total;
  }
  __syncthreads();
  if (s_total > a.dst_cap) return;  // every rank sees the same counts and skips: reported to the host
  const uint64_t n = s_n, off = s_off;
  // 16-byte stores over NVLink: the run starts at the same offset in every destination, so one scalar head brings all
  // of them to a 16-byte boundary; the source side is read with scalar (coalesced) loads
  const uint64_t h4 = (4 - (off & 3)) & 3, head = h4 < n ? h4 : n, quads = (n - head) >> 2, tail0 = head + (quads << 2);
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = t; g < quads; g += nt) {
    const uint32_t* src = a.hits + head + (g << 2);
    const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
    for (uint32_t q = 0; q < a.world; ++q) {
      const uint32_t d = (a.me + q) % a.world;  // every rank starts with itself: the peers' ingress is spread
      *reinterpret_cast<uint4*>(a.dst[d] + off + head + (g << 2)) = v;
    }
  }
  if (t < head + (n - tail0)) {
    const uint64_t i = t < head ? t : tail0 + (t - head);
    const uint32_t v = a.hits[i];
    for (uint32_t q = 0; q < a.world; ++q) a.dst[q][off + i] = v;
  }
}

__global__ void __launch_bounds__(32) k_query_barrier(RouteCtlPeers peers, uint32_t me, uint32_t world, uint64_t epoch) {
  const uint32_t t = threadIdx.x;
  if (t < world) {
    __threadfence_system();  // cumulative: the push kernel's stores go first
    st_sys(&peers.ctl[t]->qbflag[me], epoch);
    spin_until(&peers.ctl[me]->qbflag[t], epoch, &peers.ctl[me]->err);
  }
  __threadfence_system();
}

constexpr int RT_SMEM = RT_THREADS * 88;

__global__ void __launch_bounds__(RT_BLOCK) k_route_scatter_p2p(const RouteP2PArgs a) {
  BB_DYN_SMEM(s_raw);
  uint4* s_head = reinterpret_cast<uint4*>(s_raw);                   // [RT_THREADS]
  uint4* s_clk = s_head + RT_THREADS;                                 // [2 * RT_THREADS]
  uint4* s_val = s_clk + 2 * RT_THREADS;                              // [2 * RT_THREADS]
  uint64_t* s_path = reinterpret_cast<uint64_t*>(s_val + 2 * RT_THREADS);  // [RT_THREADS]
  __shared__ uint32_t s_w[RT_THREADS / 32][RT_MAX_WORLD];  // [warp-sized run of the tile][owner]
  __shared__ uint32_t s_start[RT_MAX_WORLD + 1];
  __shared__ int64_t s_dst[RT_MAX_WORLD];  // destination row of the owner's run minus its start in the tile
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // where this rank's block starts in every owner's slot comes from the all-gathered counts ON THE DEVICE:
  // the host never waits for them (row in the owner's slot = packed position + adj[owner])
  __shared__ int64_t s_adj[RT_MAX_WORLD];
  __shared__ uint32_t s_bad;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  if (tid < (int)a.world) {
    uint64_t before = 0, so = 0, col = 0;
    for (uint32_t p = 0; p < a.world; ++p) {
      const uint64_t c = a.matrix[(uint64_t)p * a.world + tid];
      col += c;
      if (p < a.me) before += c;
    }
    for (int q = 0; q < tid; ++q) so += a.matrix[(uint64_t)a.me * a.world + q];
    s_adj[tid] = (int64_t)before - (int64_t)so;
    if (col > a.slot_cap) s_bad = 1;
  }
  __syncthreads();
  if (s_bad) return;  // every rank sees the same matrix and skips; bb_router_acquire reports it
  // PERSISTENT on a small grid (bb_router: 64 CTAs): the kernel is NVLink-bound and runs next to the
  // merge of the previous batch; one CTA per tile would take every SM's thread slots away from it
  const uint32_t tiles = (uint32_t)((a.n + RT_THREADS - 1) / RT_THREADS);
  // a tile's rows travel global -> registers -> shared (partitioned) -> peer memory; the NEXT tile's rows are
  // fetched into the registers while the copy engine still drains this tile's runs out of shared memory.
  // Thread t holds rows t, t + RT_BLOCK, ... of the tile: "virtual warp" k * (RT_BLOCK / 32) + w covers 32 consecutive rows.
  uint64_t pid_raw[RT_RPT];
  uint4 h[RT_RPT], c0[RT_RPT], c1[RT_RPT], v0[RT_RPT], v1[RT_RPT];
  auto fetch = [&](uint32_t tile) {
    if (tile >= tiles) return;
#pragma unroll
    for (int k = 0; k < RT_RPT; ++k) {
      const uint64_t i = (uint64_t)tile * RT_THREADS + k * RT_BLOCK + tid;
      if (i < a.n) {
        pid_raw[k] = a.path_id[i];
        h[k] = a.head[i];
        c0[k] = a.clk[2 * i];
        c1[k] = a.clk[2 * i + 1];
        v0[k] = a.val[2 * i];
        v1[k] = a.val[2 * i + 1];
      }
    }
  };
  fetch(blockIdx.x);
  for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    uint64_t p[RT_RPT];
    uint32_t d[RT_RPT], below[RT_RPT];
#pragma unroll
    for (int k = 0; k < RT_RPT; ++k) {
      const uint64_t i = (uint64_t)tile * RT_THREADS + k * RT_BLOCK + tid;
      p[k] = i < a.n ? shard_mix(pid_raw[k], a.key_bits) : 0;
      d[k] = i < a.n ? (uint32_t)(p[k] % a.world) : a.world;
      below[k] = 0;
      for (uint32_t r = 0; r < a.world; ++r) {
        const uint32_t m = __ballot_sync(0xffffffffu, d[k] == r);
        if (d[k] == r) below[k] = __popc(m & lanemask_lt());
        if (lane == 0) s_w[k * (RT_BLOCK / 32) + w][r] = __popc(m);
      }
    }
    __syncthreads();
    if (w == 0) {  // lane r sums owner r's run counts, then an exclusive scan over the owners
      uint32_t cnt = 0;
      if (lane < (int)a.world)
        for (int ww = 0; ww < RT_THREADS / 32; ++ww) cnt += s_w[ww][lane];
      uint32_t run = cnt;
#pragma unroll
      for (int o = 1; o < RT_MAX_WORLD; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run += t;
      }
      if (lane < (int)a.world) {
        s_start[lane] = run - cnt;
        s_dst[lane] = (int64_t)a.tile_off[(uint64_t)tile * a.world + lane] + s_adj[lane] - (int64_t)(run - cnt);
        if (lane == (int)a.world - 1) s_start[a.world] = run;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RT_RPT; ++k) {
      const uint64_t i = (uint64_t)tile * RT_THREADS + k * RT_BLOCK + tid;
      if (i < a.n) {
        uint32_t lp = s_start[d[k]] + below[k];
        for (int ww = 0; ww < k * (RT_BLOCK / 32) + w; ++ww) lp += s_w[ww][d[k]];
        s_path[lp] = p[k] / a.world;
        s_head[lp] = h[k];
        s_clk[2 * lp] = c0[k];
        s_clk[2 * lp + 1] = c1[k];
        s_val[2 * lp] = v0[k];
        s_val[2 * lp + 1] = v1[k];
      }
    }
    __syncthreads();
    const uint32_t rows = s_start[a.world];
    if (a.bulk) {
      // Every (owner, array) run is contiguous in shared memory and in the owner's slot: ONE bulk copy each
      // (cp.async.bulk shared -> global, to peer memory over NVLink).  The copy engine of the SM moves the
      // bytes; no thread, register or LSU slot waits for the remote stores, so a handful of CTAs keeps
      // NVLink busy and the merge kernel running next to them keeps its SMs to itself.
      fence_proxy_async_smem();  // the partition above was written with st.shared
      if (tid < (int)(3 * a.world)) {
        const uint32_t r = tid / 3, arr = tid % 3;
        const uint32_t first = s_start[r], cnt = s_start[r + 1] - first;
        if (cnt) {
          const uint64_t drow = (uint64_t)(s_dst[r] + (int64_t)first);
          const void* src;
          void* dst;
          uint32_t bytes;
          if (arr == 0) {
            src = s_head + first, dst = a.d_head[r] + drow, bytes = cnt * 16u;
          } else if (arr == 1) {
            src = s_clk + 2 * first, dst = a.d_clk[r] + 2 * drow, bytes = cnt * 32u;
          } else {
            src = s_val + 2 * first, dst = a.d_val[r] + 2 * drow, bytes = cnt * 32u;
          }
          bulk_s2g(dst, src, bytes);
        }
        bulk_commit();
      }
      for (uint32_t j = tid; j < rows; j += RT_BLOCK) {  // the 8-byte local row ids: runs are not 16-byte aligned, plain stores
        uint32_t r = 0;
        while (j >= s_start[r + 1]) ++r;
        a.d_path[r][(uint64_t)(s_dst[r] + (int64_t)j)] = s_path[j];
      }
      fetch(tile + gridDim.x);  // the next tile's rows: in flight while the copy engine reads this tile's runs
      if (tid < (int)(3 * a.world)) bulk_wait_read_all();  // sources read: reusable
    } else {
      for (uint32_t j = tid; j < rows; j += RT_BLOCK) {  // rows of the tile in partitioned order
        uint32_t r = 0;
        while (j >= s_start[r + 1]) ++r;
        const uint64_t dst = (uint64_t)(s_dst[r] + (int64_t)j);
        a.d_path[r][dst] = s_path[j];
        a.d_head[r][dst] = s_head[j];
      }
      for (uint32_t e = tid; e < 2 * rows; e += RT_BLOCK) {  // the 32-byte columns as 16-byte pieces
        const uint32_t j = e >> 1;
        uint32_t r = 0;
        while (j >= s_start[r + 1]) ++r;
        const uint64_t dst = 2 * (uint64_t)(s_dst[r] + (int64_t)j) + (e & 1u);
        a.d_clk[r][dst] = s_clk[e];
        a.d_val[r][dst] = s_val[e];
      }
      fetch(tile + gridDim.x);
    }
    __syncthreads();  // shared memory is reused by the next tile
  }
  if (a.bulk && tid < (int)(3 * a.world)) bulk_wait_all();  // writes done
}

__global__ void __launch_bounds__(RT_THREADS) k_route_scatter(const RouteArgs a) {
  __shared__ uint32_t s_w[RT_THREADS / 32][RT_MAX_WORLD];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t i = (uint64_t)blockIdx.x * RT_THREADS + threadIdx.x;
  const uint64_t p = i < a.n ? shard_mix(a.path_id[i], a.key_bits) : 0;
  const uint32_t d = i < a.n ? (uint32_t)(p % a.world) : a.world;
  uint32_t below = 0;
  for (uint32_t r = 0; r < a.world; ++r) {
    const uint32_t m = __ballot_sync(0xffffffffu, d == r);
    if (d == r) below = __popc(m & lanemask_lt());
    if (lane == 0) s_w[w][r] = __popc(m);
  }
  __syncthreads();
  if (i >= a.n) return;
  uint32_t dst = a.tile_off[(uint64_t)blockIdx.x * a.world + d] + below;
  for (int ww = 0; ww < w; ++ww) dst += s_w[ww][d];
  a.o_path[dst] = p / a.world;
  a.o_head[dst] = a.head[i];
  a.o_clk[2 * (uint64_t)dst] = a.clk[2 * i];
  a.o_clk[2 * (uint64_t)dst + 1] = a.clk[2 * i + 1];
  a.o_val[2 * (uint64_t)dst] = a.val[2 * i];
  a.o_val[2 * (uint64_t)dst + 1] = a.val[2 * i + 1];
}


}  // namespace bb
