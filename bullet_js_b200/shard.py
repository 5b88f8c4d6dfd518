"""Key-sharded table across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).  Path id `p`
lives on rank `p % world` as local row `p // world`.  A step:

  1. every rank packs its own batch by owner with the library's stable partition
     (bb_route_pack_dev: three small launches, 88 B per update moved once);
  2. counts all-to-all (world x world int64), then one all-to-all per SoA array with
     exact split sizes - (world-1)/world of every batch crosses NVLink;
  3. every rank merges what it received, concatenated in source-rank order, into its
     shard (bb_merge_batch_dev).  The per-path replay order is therefore
     (source rank, arrival index): the same as one peer replaying rank 0's batch,
     then rank 1's, ... - which is what the parity test checks.

The reference's transport is JSON over WebSocket between peers (src/bullet-network.js:
404-418, sync chunks of 50 entries, src/bullet-network-sync.js:713-723); this module is
its B200 equivalent for ONE logical peer whose table spans several GPUs.  The exchange
itself (`Exchange`) is plumbing over torch.distributed so that the same routing logic
runs under gloo on CPU in the tests, with the pack and the merge injected.
"""
from __future__ import annotations

import numpy as np

from . import capi

ROW_BYTES = {"path": 8, "head": 16, "clk": 32, "val": 32}


def owner_of(path_id, world: int):
    return path_id % world


def local_row(path_id, world: int):
    return path_id // world


class Exchange:
    """counts all-to-all + four variable-size all-to-alls on byte tensors."""

    def __init__(self, dist, world: int, rank: int):
        self.dist, self.world, self.rank = dist, world, rank

    def counts(self, send_counts):
        """send_counts: int64 tensor [world] on the compute device -> recv counts [world] (same device)."""
        import torch

        recv = torch.empty_like(send_counts)
        self.dist.all_to_all_single(recv, send_counts)
        return recv

    def payload(self, send, send_counts, recv, recv_counts, row_bytes: int):
        """send / recv: uint8 tensors; counts: python lists of rows per peer."""
        self.dist.all_to_all_single(
            recv[: sum(recv_counts) * row_bytes], send[: sum(send_counts) * row_bytes],
            output_split_sizes=[c * row_bytes for c in recv_counts],
            input_split_sizes=[c * row_bytes for c in send_counts])


class Router:
    """Routes device-resident batches to their owner shard and merges them there."""

    def __init__(self, engine, world: int, rank: int, batch: int, device, dist=None, recv_factor: float = None):
        import torch

        if dist is None:
            import torch.distributed as dist
        self.torch, self.world, self.rank, self.dev = torch, world, rank, device
        self.ex = Exchange(dist, world, rank)
        self.batch = batch
        self.cap = batch * world if recv_factor is None else int(batch * recv_factor)  # a rank may own every update

        def buf(rows, width):
            return torch.zeros(max(rows, 1) * width, dtype=torch.uint8, device=device)

        self.send = {k: buf(batch, w) for k, w in ROW_BYTES.items()}
        self.recv = {k: buf(self.cap, w) for k, w in ROW_BYTES.items()}
        self.d_counts = torch.zeros(world, dtype=torch.int64, device=device)
        self.nvlink_bytes = 0  # bytes this rank sent to other ranks so far

    def route_and_merge(self, engine, p, h, c, v, cs: capi.BBChanges, stream: int) -> int:
        """p/h/c/v: uint8 device tensors of one batch (bb_batch arrays). Returns updates merged here."""
        torch = self.torch
        n = p.numel() // 8
        bs = capi.BBBatch(n=n, path_id=p.data_ptr(), head=h.data_ptr(), clk=c.data_ptr(), val=v.data_ptr())
        out = capi.BBBatch(n=n, path_id=self.send["path"].data_ptr(), head=self.send["head"].data_ptr(),
                           clk=self.send["clk"].data_ptr(), val=self.send["val"].data_ptr())
        engine.route_pack_dev(self.world, bs, out, self.d_counts.data_ptr(), stream)
        recv_counts_t = self.ex.counts(self.d_counts)          # NCCL, on torch's current stream
        both = torch.stack([self.d_counts, recv_counts_t]).cpu()  # the one host sync of the step
        send_counts, recv_counts = both[0].tolist(), both[1].tolist()
        n_recv = sum(recv_counts)
        if n_recv > self.cap:
            raise capi.BulletB200Error(capi.ERR_CAPACITY, f"rank {self.rank} received {n_recv} updates > {self.cap}")
        for k, w in ROW_BYTES.items():
            self.ex.payload(self.send[k], send_counts, self.recv[k], recv_counts, w)
        self.nvlink_bytes += (n - send_counts[self.rank]) * sum(ROW_BYTES.values())
        rb = capi.BBBatch(n=n_recv, path_id=self.recv["path"].data_ptr(), head=self.recv["head"].data_ptr(),
                          clk=self.recv["clk"].data_ptr(), val=self.recv["val"].data_ptr())
        engine.merge_dev(rb, cs, stream)
        return n_recv


def route_on_host(world: int, rank: int, batch, dist, merge_fn):
    """The same routing with numpy packing (stable partition by owner) and torch.distributed on
    CPU tensors - used by the gloo tests to check the exchange logic and the replay order; the
    product path is `Router` (CUDA pack + NCCL)."""
    import torch

    from . import codec

    owner = owner_of(batch.path_id, np.uint64(world)).astype(np.int64)
    order = np.argsort(owner, kind="stable")
    send_counts = np.bincount(owner, minlength=world).astype(np.int64)
    packed = codec.Batch(local_row(batch.path_id[order], np.uint64(world)), batch.head[order], batch.clk[order],
                         batch.val[order])
    ex = Exchange(dist, world, rank)
    rc = ex.counts(torch.from_numpy(send_counts)).tolist()
    sc = send_counts.tolist()
    n_recv = sum(rc)
    got = codec.Batch.empty(n_recv)
    for name, arr_s, arr_r in (("path", packed.path_id, got.path_id), ("head", packed.head, got.head),
                               ("clk", packed.clk, got.clk), ("val", packed.val, got.val)):
        s = torch.from_numpy(np.ascontiguousarray(arr_s).view(np.uint8).reshape(-1).copy())
        r = torch.zeros(max(n_recv, 1) * ROW_BYTES[name], dtype=torch.uint8)
        ex.payload(s, sc, r, rc, ROW_BYTES[name])
        arr_r.view(np.uint8).reshape(-1)[:] = r.numpy()[: n_recv * ROW_BYTES[name]]
    return merge_fn(got), got
