"""Key-sharded table across the GPUs of one box.

One process per GPU.  Path id `p` lives on rank `owner_of(p)` as local row `local_row(p)`: `p % world` / `p // world`
by default, or - `key_bits` given - the same split of the id hashed by a bijective splitmix64-style finaliser
(`shard_mix`), which spreads strided or clustered ids evenly (SURVEY 8e specifies a splitmix64 hash).  A step of the product
path (`Router` -> the library's native router, bb_router_* in include/bullet_b200.h):

  1. every rank counts its batch by owner (k_route_count / k_route_scan) and publishes the counts straight into
     every peer's control block (peer-mapped memory, epoch flags) - on its own stream, while the previous batch
     is still being exchanged;
  2. ONE persistent kernel per rank partitions the batch in shared memory and stores every update's four SoA
     pieces into its owner's receive slot over NVLink (cp.async.bulk to peer memory): (world-1)/world of every
     batch crosses NVLink, nothing is staged in between; an epoch-flag barrier tells the owners their rows are in;
  3. every rank merges what it received, concatenated in source-rank order, into its shard (bb_merge_batch_dev)
     while the next batch is being routed.  The per-path replay order is therefore (source rank, arrival index):
     the same as one peer replaying rank 0's batch, then rank 1's, ... - which is what the parity runs check
     (tests/test_shard_gloo.py on CPU, tests/check_shard_gpu.py on 2 / 4 / 8 GPUs).

torch.distributed is only used to hand rank 0's NCCL id to the other ranks (NCCL bootstraps the communicator that
carries the IPC handles and is the fallback transport: BB_ROUTER_NO_P2P, BB_ROUTE_NCCL_SYNC).

The reference's transport is JSON over WebSocket between peers (src/bullet-network.js:
404-418, sync chunks of 50 entries, src/bullet-network-sync.js:713-723); this module is
its B200 equivalent for ONE logical peer whose table spans several GPUs.  The product
path is `Router`, a thin wrapper over the library's native router (CUDA pack kernels +
NCCL called from C).  `Exchange` / `route_on_host` restate the same routing over
torch.distributed so that it runs under gloo on CPU in the tests, with numpy packing
and the merge injected.
"""
from __future__ import annotations

import numpy as np

from . import capi

ROW_BYTES = {"path": 8, "head": 16, "clk": 32, "val": 32}


def shard_mix(path_id, key_bits: int = 0):
    """The library's sharding function (csrc/bb_route.cuh shard_mix): identity for key_bits == 0, otherwise a
    splitmix64-style finaliser restricted to key_bits bits - a bijection of [0, 2**key_bits)."""
    x = np.asarray(path_id, dtype=np.uint64)
    if key_bits == 0:
        return x
    mask = np.uint64((1 << key_bits) - 1)
    s = np.uint64((key_bits + 1) // 2)
    x = x & mask
    with np.errstate(over="ignore"):
        for mul in (0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB):
            x = (x * np.uint64(mul)) & mask
            x = x ^ (x >> s)
    return x


def owner_of(path_id, world: int, key_bits: int = 0):
    return shard_mix(path_id, key_bits) % np.uint64(world)


def local_row(path_id, world: int, key_bits: int = 0):
    return shard_mix(path_id, key_bits) // np.uint64(world)


def shard_capacity(world: int, key_bits: int, n_ids: int = 0) -> int:
    """Rows a shard's table needs: ceil(2**key_bits / world) with hashed sharding, ceil(n_ids / world) without."""
    return -(-(1 << key_bits) // world) if key_bits else -(-n_ids // world)


class Exchange:
    """counts all-to-all + ONE grouped exchange of the four SoA arrays (every array's slice for
    every peer is a send/recv of the same NCCL group, so it is a single launch on the wire)."""

    def __init__(self, dist, world: int, rank: int):
        self.dist, self.world, self.rank = dist, world, rank

    def counts(self, send_counts):
        """send_counts: int64 tensor [world] on the compute device -> recv counts [world] (same device)."""
        import torch

        recv = torch.empty_like(send_counts)
        self.dist.all_to_all_single(recv, send_counts)
        return recv

    def payload(self, send: dict, send_counts, recv: dict, recv_counts):
        """send / recv: {name: uint8 tensor}; rows for peer r start at sum(counts[:r]) in every array."""
        dist = self.dist
        so = np.concatenate([[0], np.cumsum(send_counts)]).tolist()
        ro = np.concatenate([[0], np.cumsum(recv_counts)]).tolist()
        ops = []
        for k, w in ROW_BYTES.items():
            me = self.rank
            recv[k][ro[me] * w: ro[me + 1] * w].copy_(send[k][so[me] * w: so[me + 1] * w])  # own rows stay here
            for r in range(self.world):
                if r == me:
                    continue
                if recv_counts[r]:
                    ops.append(dist.P2POp(dist.irecv, recv[k][ro[r] * w: ro[r + 1] * w], r))
                if send_counts[r]:
                    ops.append(dist.P2POp(dist.isend, send[k][so[r] * w: so[r + 1] * w], r))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


class Router:
    """The library's native router (bb_router_*: pack kernels, NCCL all-gather of the counts,
    grouped ncclSend/ncclRecv, two receive slots on its own stream).  torch.distributed is only
    used to hand rank 0's NCCL id to the other ranks.  `route` and `merge` are separate so that
    routing batch i+1 overlaps merging batch i."""

    def __init__(self, world: int, rank: int, batch: int, device_index: int, dist=None, recv_capacity: int = 0,
                 key_bits: int = 0):
        import ctypes as C

        if dist is None and world > 1:
            import torch.distributed as dist
        self.lib = capi.load()
        self.world, self.rank = world, rank
        idbuf = C.create_string_buffer(capi.NCCL_ID_BYTES)
        if rank == 0 and world > 1:  # a one-rank router needs no communicator
            rc = self.lib.bb_router_unique_id(idbuf)
            if rc:
                raise capi.BulletB200Error(rc, (self.lib.bb_router_last_error(None) or b"").decode())
        box = [idbuf.raw]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        h = C.c_void_p()
        rc = self.lib.bb_router_create(device_index, world, rank, box[0], batch, recv_capacity, C.byref(h))
        if rc:
            raise capi.BulletB200Error(rc, (self.lib.bb_router_last_error(None) or b"").decode())
        self._h = h
        self.key_bits = key_bits
        if key_bits:
            self._check(self.lib.bb_router_set_sharding(self._h, key_bits))

    def _check(self, rc):
        if rc:
            raise capi.BulletB200Error(rc, (self.lib.bb_router_last_error(self._h) or b"").decode())

    def route(self, bs: capi.BBBatch, slot: int, in_stream: int = 0) -> int:
        """Collective: pack the device batch by owner and exchange it into receive slot `slot`.  Asynchronous when
        the peers' slots are mapped (P2P): the received count is then only known in `merge` (returns 2**64 - 1)."""
        import ctypes as C

        n = C.c_uint64(0)
        self._check(self.lib.bb_router_route_dev(self._h, C.byref(bs), slot, C.byref(n), C.c_void_p(in_stream)))
        return int(n.value)

    def merge(self, engine, slot: int, cs: capi.BBChanges, stream: int) -> int:
        """Merge receive slot `slot` into this rank's shard on `stream` (a raw cudaStream_t)."""
        import ctypes as C

        rb = capi.BBBatch()
        self._check(self.lib.bb_router_acquire(self._h, slot, C.c_void_p(stream), C.byref(rb)))
        engine.merge_dev(rb, cs, stream)
        self._check(self.lib.bb_router_release(self._h, slot, C.c_void_p(stream)))
        return int(rb.n)

    def merge_batch(self, engine, bs: capi.BBBatch, cs: capi.BBChanges, chunks: int = 0):
        """Collective host entry (bb_router_merge_batch): this rank's HOST batch in, verdicts + change entries of what
        this shard received out (host buffers of `cs`).  Returns (n_received, recv_counts[chunks, world])."""
        import ctypes as C

        k = chunks or 4
        counts = np.zeros((k, self.world), np.uint64)
        n = C.c_uint64(0)
        self._check(self.lib.bb_router_merge_batch(self._h, engine._h, C.byref(bs), C.byref(cs), k, C.byref(n),
                                                   counts.ctypes.data))
        return int(n.value), counts

    # ---- sharded queries (src/bullet-query.js:186-210, 221-261 over a table that spans the ranks)
    def query_reserve(self, max_total_hits: int):
        """Collective, once: buffers for this rank's scan and for the gathered result on every rank."""
        self._check(self.lib.bb_router_query_reserve(self._h, int(max_total_hits)))

    def query_range(self, engine, field: int, lo: capi.BBBound, hi: capi.BBBound, stream: int) -> capi.BBGatheredHits:
        """Collective: every rank scans its shard and stores its local hit ids into every rank's result buffer over
        NVLink.  The result (device memory) is rank 0's hits, then rank 1's, ...: offset[q] .. offset[q + 1]."""
        import ctypes as C

        out = capi.BBGatheredHits()
        self._check(self.lib.bb_router_query_range(self._h, engine._h, field, C.byref(lo), C.byref(hi), C.byref(out),
                                                   C.c_void_p(stream)))
        return out

    def query_equals(self, engine, field: int, key: int, stream: int) -> capi.BBGatheredHits:
        import ctypes as C

        out = capi.BBGatheredHits()
        self._check(self.lib.bb_router_query_equals(self._h, engine._h, field, int(key), C.byref(out), C.c_void_p(stream)))
        return out

    def query_fetch(self, first: int, n: int) -> np.ndarray:
        """n gathered local hit ids starting at `first`, copied to the host."""
        host = np.empty(max(int(n), 1), np.uint32)
        self._check(self.lib.bb_router_query_fetch(self._h, int(first), int(n), host.ctypes.data))
        return host[: int(n)]

    def last_ms(self) -> dict:
        import ctypes as C

        out = (C.c_double * 6)()
        self._check(self.lib.bb_router_last_ms(self._h, out))
        return dict(zip(("pack", "counts", "exchange", "own_copy", "host_until_counts", "host_call"), out))

    @property
    def sent_bytes(self) -> int:
        return int(self.lib.bb_router_sent_bytes(self._h))

    @property
    def launches(self) -> int:
        return int(self.lib.bb_router_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.bb_router_destroy(self._h)
            self._h = None


def route_on_host(world: int, rank: int, batch, dist, merge_fn, key_bits: int = 0):
    """The same routing with numpy packing (stable partition by owner) and torch.distributed on
    CPU tensors - used by the gloo tests to check the exchange logic and the replay order; the
    product path is `Router` (CUDA pack + NCCL)."""
    import torch

    from . import codec

    owner = owner_of(batch.path_id, world, key_bits).astype(np.int64)
    order = np.argsort(owner, kind="stable")
    send_counts = np.bincount(owner, minlength=world).astype(np.int64)
    packed = codec.Batch(local_row(batch.path_id[order], world, key_bits), batch.head[order], batch.clk[order],
                         batch.val[order])
    ex = Exchange(dist, world, rank)
    rc = ex.counts(torch.from_numpy(send_counts)).tolist()
    sc = send_counts.tolist()
    n_recv = sum(rc)
    got = codec.Batch.empty(n_recv)
    arrays = {"path": (packed.path_id, got.path_id), "head": (packed.head, got.head),
              "clk": (packed.clk, got.clk), "val": (packed.val, got.val)}
    send = {k: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()) for k, (a, _) in arrays.items()}
    recv = {k: torch.zeros(max(n_recv, 1) * w, dtype=torch.uint8) for k, w in ROW_BYTES.items()}
    ex.payload(send, sc, recv, rc)
    for k, (_, dst) in arrays.items():
        dst.view(np.uint8).reshape(-1)[:] = recv[k].numpy()[: n_recv * ROW_BYTES[k]]
    return merge_fn(got), got


def merge_batch_on_host(world: int, rank: int, batch, dist, merge_fn, pieces: int = 4, key_bits: int = 0):
    """The replay order of bb_router_merge_batch (include/bullet_b200.h) restated over `route_on_host`: every rank's
    batch is cut into `pieces` pieces in arrival order; piece j of every rank is routed and merged before piece j + 1 -
    a shard replays piece by piece, inside a piece in (source rank, arrival index) order.  -> [(changes, received batch)]
    per piece.  For the gloo tests; the product path is the native router."""
    chunk = -(-batch.n // pieces) if batch.n else 0
    out = []
    for j in range(pieces):
        lo, hi = min(j * chunk, batch.n), min((j + 1) * chunk, batch.n)
        out.append(route_on_host(world, rank, batch.slice(lo, hi), dist, merge_fn, key_bits))
    return out
