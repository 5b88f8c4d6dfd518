#!/usr/bin/env python
"""bench.py - the BASELINE.json metric on BASELINE.json's configs.

A "step" = one pass of the hot path over one synthetic batch of conflicting updates (F=4 fields each).

  N = 1   config 2 (BASELINE.json configs[1]): 1 M updates merged into a resident table of 2.5 M
          records (10 M fields), uniform keys - the variant the 8d roofline figure is worked on.
  N > 1   config 3 (configs[2]): the table is sharded by a hash of the path id over the N GPUs, 31.25 M records
          (125 M fields, 4 GB of rows) and a 2 M-update batch per GPU and step - 250 M records / 1 B fields / 16 M
          updates per step at N = 8, exactly the stated config; every batch is routed to its owning shards by the
          library's fused pack + all-to-all over NVLink peer memory (routing of batch i+1 overlaps the merge of
          batch i).  Weak scaling: the per-GPU shard is fixed.

  value   field-merges/s, whole job, inputs already resident in HBM (bb_merge_batch_dev / bb_router_*; CUDA events
          on the launching stream, max over ranks)
  e2e     same metric through the reference-facing C-ABI call with pinned HOST buffers: H2D of the batch and D2H
          of decisions + change set inside the timed region
  parity  step 0 of the very workload that is timed, compared bit for bit (decisions, change set, every touched
          row) with the typed C oracle on the host; per shard at N > 1; query hit multisets against numpy.  A
          mismatch fails the run.
  roofline      the dominant kernel (k_merge_stage): algorithmic bytes / its mean launch duration / measured HBM peak,
                plus the same for the whole step (pipeline_frac)
  cpu_baseline  the typed C oracle (oracle/bullet_oracle.c, a restatement of the reference's JS: kind "port"),
                1 thread, bounded sample
  zipf    the "conflicting" Zipf(0.8) variant of config 2, default configuration, without and with an index
  query   config 4: index build + range(20,30) + equals(role,'admin') over --query-records nodes per GPU
  --impl reference   the same C oracle on every host thread.  The reference itself is JavaScript and no JS engine
                exists on the box: every ratio against this arm is against the C PORT, not Node.js.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints ONE line

F = 4
N_BATCHES = 4            # distinct pre-generated batches, cycled
CFG2 = dict(records=2_500_000, batch=1_000_000)
CFG3 = dict(records=31_250_000, batch=2_000_000, image=1 << 22)  # per GPU; the table image is tiled over the shard
REFERENCE_KIND = "C port of the reference's JavaScript (oracle/bullet_oracle.c); Node.js itself was not run: no JS engine on the box"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=0, help="records per GPU (default: config 2 at N=1, config 3 at N>1)")
    ap.add_argument("--batch", type=int, default=0, help="updates per GPU and step")
    ap.add_argument("--keys", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--query-records", type=int, default=100_000_000,
                    help="nodes per GPU for the index build + range/equals scans (BASELINE config 4)")
    ap.add_argument("--no-query", action="store_true")
    ap.add_argument("--no-zipf", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the step-0 comparison with the oracle (debugging only)")
    ap.add_argument("--config", default="merge", choices=["merge", "mesh"],
                    help="merge: configs 2 / 3 / 4 (the default line); mesh: BASELINE config 5, mesh-topology sync replay (bench_mesh.py)")
    ap.add_argument("--mesh-ops", type=int, default=10_000_000, help="--config mesh: log entries per peer")
    ap.add_argument("--mesh-batch", type=int, default=1_000_000, help="--config mesh: updates per replayed batch")
    ap.add_argument("--front-end", default="group", choices=["group", "full", "radix"],
                    help="how a batch is grouped by path: the library default, BB_CFG_FULL_SORT or BB_CFG_RADIX_SORT")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    base = CFG2 if world == 1 else CFG3
    a.records = a.records or base["records"]
    a.batch = a.batch or base["batch"]
    return a


def ncu_traffic(kernel, expect_default):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu capture
    (profiles/traffic.json); null when the run is not the captured workload."""
    if not expect_default:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return float(json.load(f)[kernel]["bytes_per_launch"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        loaded = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {
            "sm_mhz": float(np.median(loaded)) if loaded else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ---------------------------------------------------------------------------------------------- workloads
def key_bits_for(world, records):
    """Hashed sharding needs ids < 2^key_bits with 2^key_bits >= world * records (bullet_js_b200/shard.py)."""
    return int(np.ceil(np.log2(world * records)))


def shard_mix_inverse(x, key_bits):
    """Inverse of shard.shard_mix (each step of the finaliser is a bijection of [0, 2^key_bits))."""
    x = np.asarray(x, np.uint64)
    mask = np.uint64((1 << key_bits) - 1)
    s = np.uint64((key_bits + 1) // 2)
    with np.errstate(over="ignore"):
        for mul in (0x94D049BB133111EB, 0xBF58476D1CE4E5B9, 0x9E3779B97F4A7C15):
            x = x ^ (x >> s)  # 2 s >= key_bits: the xor-shift is its own inverse
            x = (x * np.uint64(pow(mul, -1, 1 << key_bits))) & mask
    return x


def make_workload(args, rank, world):
    """-> (image, batches, meta).  N = 1: the table IS the image (row i == path id i).  N > 1 (config 3): every shard
    holds `cap` rows, local row r = image row r % len(image) (the image tiled); a batch row drawn against image row p
    goes to a uniformly drawn owner q and tile t: local row r = t * len(image) + p, scrambled id x = r * world + q,
    global path id = shard_mix^-1(x).  So every shard sees ~batch updates spread uniformly over ALL of its rows,
    with clocks built against the row they hit, and the path ids look like the hashed ids they are."""
    from bullet_js_b200 import synth

    rng = synth.rng_for(2 if world == 1 else 3, salt=rank)
    if world == 1:
        image = synth.make_table(args.records, rng)
        batches = [synth.make_batch(image, args.batch, rng, keys=args.keys) for _ in range(N_BATCHES)]
        return image, batches, dict(capacity=args.records, key_bits=0, tiles=1)
    kb = key_bits_for(world, args.records)
    cap = (1 << kb) // world
    n_img = min(CFG3["image"], cap)
    assert cap % n_img == 0
    image = synth.make_table(n_img, synth.rng_for(3, salt=1000))  # the same image on every rank
    batches = [sharded_batch(image, args.batch, synth.rng_for(3, salt=rank * 16 + j), args.keys, world, kb, cap)
               for j in range(N_BATCHES)]
    return image, batches, dict(capacity=cap, key_bits=kb, tiles=cap // n_img)


def sharded_batch(image, n, rng, keys, world, kb, cap):
    from bullet_js_b200 import synth

    b = synth.make_batch(image, n, rng, keys=keys)
    tiles = cap // image.n
    r = rng.integers(0, tiles, n).astype(np.uint64) * np.uint64(image.n) + b.path_id
    q = rng.integers(0, world, n).astype(np.uint64)
    b.path_id[:] = shard_mix_inverse(r * np.uint64(world) + q, kb)
    return b


def workload_config(args, world, meta=None):
    if world == 1:
        wl = (f"config2: {args.records * F // 1_000_000}M-field table ({args.records} records x {F} fields, 128 B rows), "
              f"{args.batch}-update conflicting batch per step, {args.keys} keys, clock mix 40/20/25/5/5/5")
    else:
        wl = (f"config3: {world * args.records * F / 1e9:.2f}B-field table sharded by key hash over {world} GPUs "
              f"({args.records} records = {args.records * 128 / 2**30:.1f} GiB of rows per GPU), {world * args.batch}-update "
              f"batch per step ({args.batch} submitted per GPU, routed all-to-all), {args.keys} keys, clock mix 40/20/25/5/5/5")
    return {
        "workload": wl, "records_per_gpu": args.records, "batch_per_gpu": args.batch, "fields": F, "peers": 8,
        "keys": args.keys, "front_end": args.front_end,
        "sharding": "none" if world == 1 else f"splitmix64-style bijective hash of the path id (key_bits {meta['key_bits'] if meta else '?'}) % {world}",
        "l2": f"working set {args.records * 128 // 2**20} MiB table + {N_BATCHES} x {args.batch * 88 // 2**20} MiB batches > 126 MB L2; "
              "every step merges into a pristine copy of the table",
    }


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's algorithm on the host cores: typed C restatement, all threads (see REFERENCE_KIND)."""
    if rank != 0:
        return
    from bullet_js_b200 import capi, synth
    from oracle.typed import TypedOracle

    ref_args = argparse.Namespace(**vars(args))
    if world > 1:  # one rank's share of config 3 is not a sensible CPU sample: the single-GPU workload instead
        ref_args.records, ref_args.batch = CFG2["records"], CFG2["batch"]
    image, batches, _ = make_workload(ref_args, 0, 1)
    cores = os.cpu_count() or 1
    cfg = capi.make_config(ref_args.records, **synth.synth_ranks(ref_args.records))
    orc = TypedOracle(cfg)
    out = capi.ChangeBuffers(ref_args.batch)
    dt = 0.0
    for i in range(args.warmup + args.steps):
        orc.table[:] = image.rows  # every step merges into the pristine table (not timed)
        t0 = time.perf_counter()
        orc.merge(batches[i % N_BATCHES], threads=cores, out=out)
        if i >= args.warmup:
            dt += time.perf_counter() - t0
    v = args.steps * ref_args.batch * F / dt
    line = {
        "impl": "reference", "metric": "crdt_field_merges_per_sec", "value": v, "unit": "field-merges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64",
        "data": "synthetic", "config": workload_config(ref_args, 1),
        "cpu_baseline": {"value": v, "unit": "field-merges/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x {ref_args.batch} updates on the same table, path-sharded over {cores} threads"},
        "e2e": {"value": v, "unit": "field-merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_kind": REFERENCE_KIND,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- helpers (GPU)
class DevOut:
    """Device-resident change-set buffers of one merge call."""

    def __init__(self, torch, dev, cap):
        from bullet_js_b200 import capi

        self.cap = cap
        self.ver = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.n = torch.zeros(1, dtype=torch.int64, device=dev)
        self.idx = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.head = torch.zeros(cap * 16, dtype=torch.uint8, device=dev)
        self.clk = torch.zeros(cap * 32, dtype=torch.uint8, device=dev)
        self.val = torch.zeros(cap * 32, dtype=torch.uint8, device=dev)
        self.cs = capi.BBChanges(cap=cap, verdict=self.ver.data_ptr(), n_changes=self.n.data_ptr(), idx=self.idx.data_ptr(),
                                 head=self.head.data_ptr(), clk=self.clk.data_ptr(), val=self.val.data_ptr())

    def changes(self, m):
        """The first m verdicts + the change set, copied to the host, in arrival order."""
        from bullet_js_b200 import codec

        k = int(self.n.item())
        return codec.Changes.from_verdicts(
            self.ver[:m].cpu().numpy().view(np.uint32), self.idx[:k].cpu().numpy().view(np.uint32),
            self.head[: k * 16].cpu().numpy().view(codec.HEAD_DTYPE),
            self.clk[: k * 32].cpu().numpy().view(np.uint32).reshape(k, 8),
            self.val[: k * 32].cpu().numpy().view(np.uint64).reshape(k, 4))


def to_dev(torch, dev, a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)


def dev_batch(torch, dev, b):
    from bullet_js_b200 import capi

    t = (to_dev(torch, dev, b.path_id), to_dev(torch, dev, b.head), to_dev(torch, dev, b.clk), to_dev(torch, dev, b.val))
    return t, capi.BBBatch(n=b.n, path_id=t[0].data_ptr(), head=t[1].data_ptr(), clk=t[2].data_ptr(), val=t[3].data_ptr())


def oracle_check(cfg, rows, batch, got, read_rows, indexed_field=None):
    """Replay `batch` over `rows` (row i of `rows` is path id i of the batch) with the typed oracle and compare
    decisions, change set and every touched row with what the device produced.  -> (ok, detail)."""
    from oracle.typed import TypedOracle

    orc = TypedOracle(cfg)
    orc.table[: len(rows)] = rows
    if indexed_field is not None:
        orc.index_create(indexed_field)
    want = orc.merge(batch, threads=1 if indexed_field is not None else (os.cpu_count() or 1))
    if not got.same_as(want):
        same_shape = got.decision.shape == want.decision.shape
        bad = int(np.argmax(got.decision != want.decision)) if same_shape and (got.decision != want.decision).any() else -1
        return False, f"decisions / change set differ (first differing decision at update {bad})"
    touched = np.unique(batch.path_id)
    have, expect = read_rows(touched), orc.table[touched.astype(np.int64)]
    have["xcnt"] = 0  # device-private index bookkeeping (include/bullet_b200.h), not part of the reference's state
    expect["xcnt"] = 0
    if not np.array_equal(have, expect):
        return False, "table rows differ after the batch"
    return True, f"{batch.n} updates, {len(want.idx)} accepted, {touched.size} rows compared bit for bit"


def compact_for_oracle(image_rows, tile_rows, batch_local):
    """Oracle-sized view of a huge shard: the distinct local rows a batch touches become ids 0..d-1."""
    from bullet_js_b200 import codec

    uniq, inv = np.unique(batch_local.path_id, return_inverse=True)
    rows = image_rows[(uniq % np.uint64(tile_rows)).astype(np.int64)]
    b = codec.Batch(inv.astype(np.uint64), batch_local.head, batch_local.clk, batch_local.val)
    return uniq, rows, b


# ---------------------------------------------------------------------------------------------- config 4
def run_queries(args, rank, world, local_rank, dev, image, dist):
    """BASELINE config 4: index('users','age') + index('users','role') build, range(age,20,30), equals(role,'admin')
    over --query-records nodes per GPU (the image tiled), every rank scanning its shard; at world > 1 the hit ids are
    all-gathered.  Parity: the hit MULTISETS of this rank's scans against a numpy evaluation of the table image."""
    import torch

    from bullet_js_b200 import capi, codec, synth
    from bullet_js_b200.engine import Engine

    nq = args.query_records
    eng = Engine(nq, device=local_rank, post_getdata=True, **synth.synth_ranks(image.n))
    base = np.arange(image.n, dtype=np.uint64)
    for off in range(0, nq, image.n):
        m = min(image.n, nq - off)
        eng.table_load(base[:m] + np.uint64(off), image.rows[:m])
    t0 = time.perf_counter()
    eng.index_create_fields((0, 2), extra_capacity=1 << 16)  # age and role: one pass over the table builds both
    build_ms_both = eng.phase_ms("scan")
    build_wall = time.perf_counter() - t0
    side = torch.cuda.current_stream(dev)
    stream = side.cuda_stream
    cap = nq
    hits = torch.zeros(cap, dtype=torch.int32, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    hs = capi.BBHits(cap=cap, node=hits.data_ptr(), n_dense=cnt.data_ptr(), n_extra=cnt[1:].data_ptr())
    lo = capi.BBBound(num=20.0, rank=0, flags=0, reserved=0)
    hi = capi.BBBound(num=30.0, rank=0, flags=0, reserved=0)
    admin = codec.KEY_STR | 0

    def range_dev():
        eng._check(eng.lib.bb_query_range_dev(eng._h, 0, C.byref(lo), C.byref(hi), C.byref(hs), C.c_void_p(stream)))

    def equals_dev():
        eng._check(eng.lib.bb_query_equals_dev(eng._h, 2, admin, C.byref(hs), C.c_void_p(stream)))

    # what the reference would return, as a set of node ids: numpy over the image, tiled like the table
    age = image.rows["val"][:, 0].view(np.float64)
    role = image.rows["val"][:, 2]
    want = {"range": np.nonzero((age >= 20.0) & (age <= 30.0))[0], "equals": np.nonzero(role == 0)[0]}

    def expected_count(sel):
        full, rest = divmod(nq, image.n)
        return full * sel.size + int(np.searchsorted(sel, rest))

    out, parity = {}, {}
    reps = 10
    for name, fn in (("range", range_dev), ("equals", equals_dev)):
        for _ in range(3):
            fn()
        eng.sync(stream)
        nh = int(cnt.sum().item())
        if not args.no_parity:  # every hit id maps back to an image row that satisfies the predicate, each node once
            got = np.sort(hits[:nh].cpu().numpy().view(np.uint32).astype(np.int64))
            ok = nh == expected_count(want[name])
            if ok:
                ok = bool(np.isin(got % image.n, want[name]).all()) and bool((np.diff(got) > 0).all())
            parity[name] = "ok" if ok else "MISMATCH"
            del got
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        kernel_ms = float(np.mean([eng.phase_ms("scan", j) for j in range(reps)]))
        out[name] = {"ms": ms, "scan_ms": kernel_ms, "hits": nh}
    range_dev()
    eng.sync(stream)
    # e2e through the host entry points (pinned output, D2H of the hits inside the timed region)
    hn = out["range"]["hits"] + 16
    h_node = torch.zeros(hn, dtype=torch.int32).pin_memory()
    h_cnt = torch.zeros(2, dtype=torch.int64).pin_memory()
    hhs = capi.BBHits(cap=hn, node=h_node.data_ptr(), n_dense=h_cnt.data_ptr(), n_extra=h_cnt[1:].data_ptr())
    eng.query_range_raw(0, lo, hi, hhs)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.query_range_raw(0, lo, hi, hhs)
    e2e_ms = (time.perf_counter() - t0) / 5 * 1e3
    if int(h_cnt.sum()) != out["range"]["hits"]:
        parity["range"] = "MISMATCH"
    if dist is not None:
        # sharded query in the LIBRARY (bb_router_query_range): scan + counts through the peer-mapped control blocks +
        # one kernel per rank that stores its u32 local hit ids into every rank's result buffer over NVLink
        from bullet_js_b200 import shard

        t = torch.tensor([out["range"]["ms"], out["equals"]["ms"], e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["range"]["ms"], out["equals"]["ms"], e2e_ms = (float(x) for x in t)
        tot = torch.tensor([out["range"]["hits"]], device=dev, dtype=torch.int64)
        dist.all_reduce(tot)
        qr = shard.Router(world, rank, 1024, local_rank)
        qr.query_reserve(int(tot.item()) + 4096)
        for _ in range(3):
            g = qr.query_range(eng, 0, lo, hi, stream)
        if not args.no_parity:
            # every rank holds the same tiled image, so every rank's run must be this rank's own hit multiset
            class _View:
                def __init__(self, ptr, n):
                    self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}

            allh = torch.as_tensor(_View(g.node, int(g.total)), device=dev)
            range_dev()
            eng.sync(stream)
            mine = torch.sort(hits[: int(cnt.sum().item())]).values
            ok = int(g.total) == int(tot.item())
            for q in range(world):
                run = allh[int(g.offset[q]): int(g.offset[q + 1])]
                ok = ok and run.numel() == mine.numel() and bool(torch.equal(torch.sort(run).values, mine))
            parity["sharded_range"] = "ok" if ok else "MISMATCH"
            del allh, mine
        dist.barrier()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(reps):
            g = qr.query_range(eng, 0, lo, hi, stream)
        g1.record()
        torch.cuda.synchronize()
        t = torch.tensor([g0.elapsed_time(g1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["sharded_range"] = {
            "ms": float(t[0]), "total_hits": int(g.total), "rows_per_sec": nq * world / (float(t[0]) * 1e-3),
            "nvlink_bytes_out_per_gpu": out["range"]["hits"] * 4 * (world - 1),
            "api": "bb_router_query_range: k_index_scan on every shard, counts through peer-mapped control blocks, k_query_push "
                   "stores each rank's u32 local hit ids into every rank's result buffer over NVLink (16-byte stores), epoch "
                   "barrier; result resident on every GPU; each call synchronises its stream"}
        qr.close()
    peak, _ = peaks()
    res = {
        "workload": f"config4: index(age)+index(role) build, range(age,20,30), equals(role,'admin') over {nq} nodes/GPU",
        "nodes_per_gpu": nq, "n_gpus": world,
        "range_rows_per_sec": nq * world / (out["range"]["ms"] * 1e-3),
        "equals_rows_per_sec": nq * world / (out["equals"]["ms"] * 1e-3),
        "range_hits": out["range"]["hits"], "equals_hits": out["equals"]["hits"],
        "range_ms": out["range"]["ms"], "equals_ms": out["equals"]["ms"],
        "index_build_ms": {"age_and_role_one_pass": build_ms_both, "per_index": build_ms_both / 2, "wall_both": build_wall * 1e3},
        "index_build_rows_per_sec": 2 * nq * world / (build_ms_both * 1e-3),
        "index_build_roofline": {"algorithmic_bytes_per_row_and_index": 20.0, "indices": 2,
                                 "achieved": 2 * 20.0 * nq / (build_ms_both * 1e-3) / 1e9,
                                 "frac": 2 * 20.0 * nq / (build_ms_both * 1e-3) / 1e9 / peak,
                                 "traffic": ncu_traffic("k_index_build", world == 1 and nq == 100_000_000),
                                 "note": "SURVEY 8d counts a columnar source: 8 B read + 12 B written per row and index.  The table is row-major "
                                         "(128-byte rows): one pass builds both indices, touching the first 64 B of every row, but DRAM delivers "
                                         "whole 128-byte lines (ncu: 12.8 GB read + 1.6 GB written per 100 M rows) - `traffic_frac` is that traffic "
                                         "over the kernel time against the same peak"},
        "e2e_range": {"rows_per_sec": nq * world / (e2e_ms * 1e-3), "ms": e2e_ms, "d2h_bytes": out["range"]["hits"] * 4 + 16,
                      "api": "bb_query_range (host hit buffer, synchronous)"},
        "parity": parity,
        "roofline": {"bound": "hbm", "kernel": "k_index_scan", "unit": "GB/s", "peak": peak,
                     "achieved": (8.0 * nq + 4.0 * out["range"]["hits"]) / (out["range"]["scan_ms"] * 1e-3) / 1e9,
                     "bytes_per_row": 8.0 + 4.0 * out["range"]["hits"] / nq, "kernel_ms": out["range"]["scan_ms"],
                     "traffic": ncu_traffic("k_index_scan", world == 1 and nq == 100_000_000)},
    }
    if res["index_build_roofline"]["traffic"]:
        res["index_build_roofline"]["traffic_frac"] = res["index_build_roofline"]["traffic"] / (build_ms_both * 1e-3) / 1e9 / peak
    res["roofline"]["frac"] = res["roofline"]["achieved"] / peak
    # the measured peak is a COPY (half reads, half writes); this kernel is a pure read stream and can exceed it:
    # also report it against the data-sheet HBM3e figure the profiling recipe quotes
    res["roofline"]["frac_of_nominal_7700"] = res["roofline"]["achieved"] / 7700.0
    if "sharded_range" in out:
        res["sharded_range"] = out["sharded_range"]
    eng.close()
    return res


def pcie_probe(torch, dev, nbytes, dist):
    """What the box's PCIe path gives this process while every rank does the same: H2D alone, D2H alone and both at once
    (pinned buffers, two streams, `nbytes` each way) - the ceiling the e2e numbers live under."""
    h_a, h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a, d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    out = {}
    for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        for timed in (False, True):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                if up:
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_a, non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        h_b.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 4
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name + "_gb_per_s_per_gpu_per_direction"] = nbytes / float(t[0]) / 1e9
    out["note"] = "measured in this run with every rank copying at once; max time over ranks"
    return out


# ---------------------------------------------------------------------------------------------- zipf section
def run_zipf(args, local_rank, dev, image, stream, torch):
    """The "conflicting" variant SURVEY 8d names: Zipf(0.8) keys (the hottest path takes ~1 % of the batch), default
    configuration, without and with an index on `age`; each checked against the oracle."""
    from bullet_js_b200 import synth
    from bullet_js_b200.engine import Engine

    n = args.batch
    rng = synth.rng_for(2, salt=77)
    zb = [synth.make_batch(image, n, rng, keys="zipf") for _ in range(2)]
    ids = np.arange(args.records, dtype=np.uint64)
    res = {"workload": f"{n}-update Zipf(0.8) batch over {args.records} records",
           "hottest_path_updates": int(np.bincount(zb[0].path_id.astype(np.int64)).max())}
    for label, indexed in (("default", False), ("indexed_age", True)):
        engines = []
        for _ in range(4):
            e = Engine(args.records, device=local_rank, post_getdata=indexed, **synth.synth_ranks(args.records))
            e.table_load(ids, image.rows)
            if indexed:
                e.index_create(0, extra_capacity=2 * n)
            e.reserve(n, host_entry=False)
            engines.append(e)
        out = DevOut(torch, dev, n)
        d_in = [dev_batch(torch, dev, b) for b in zb]
        engines[0].merge_dev(d_in[0][1], out.cs, stream)
        engines[0].sync(stream)
        ok, detail = (True, "skipped") if args.no_parity else oracle_check(
            engines[0].cfg, image.rows, zb[0], out.changes(n), lambda t: engines[0].table_read(t), indexed_field=0 if indexed else None)
        ms = []
        for j, e in enumerate(engines[1:]):
            e.merge_dev(d_in[(j + 1) % 2][1], out.cs, stream)
            e.sync(stream)
            ms.append(e.phase_ms("device"))
        res[label] = {"ms_per_batch": float(np.mean(ms)), "field_merges_per_sec": n * F / (float(np.mean(ms)) * 1e-3),
                      "parity": "ok" if ok else "MISMATCH: " + detail}
        for e in engines:
            e.close()
    return res


# ---------------------------------------------------------------------------------------------- main
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if args.config == "mesh":
        import bench_mesh

        args.records = CFG2["records"] if args.records in (CFG2["records"], CFG3["records"]) else args.records
        bench_mesh.run(args, rank, world, local_rank, sys.modules[__name__])
        return

    import torch

    from bullet_js_b200 import capi, codec, shard, synth
    from bullet_js_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bullet_js_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    image, batches, meta = make_workload(args, rank, world)
    n, K, W = args.batch, args.steps, args.warmup
    capacity, kb = meta["capacity"], meta["key_bits"]
    ranks_kw = synth.synth_ranks(image.n)
    # Every step merges into a PRISTINE copy of the table (one bb_ctx per step, all loaded with the same image): the
    # batches' clocks are built relative to that image, so re-merging into an already-merged table would turn the
    # workload into "all historical".  It also means no step ever finds its table in L2.
    per_engine = capacity * (128 + 12) + n * 2 * 60 + (160 << 20)  # table + per-path scratch + per-batch scratch + load staging
    total_mem = torch.cuda.get_device_properties(dev).total_memory
    if (W + K) * per_engine > 0.75 * total_mem:
        K = max(3, int(0.75 * total_mem // per_engine) - W)
        print(f"[bench] {W + args.steps} pristine tables of {per_engine >> 20} MiB do not fit: timing {K} steps", file=sys.stderr)
    ids = np.arange(capacity, dtype=np.uint64)

    def load_pristine(e):
        step = 1 << 20  # (bounded staging buffers inside the library)
        for t in range(meta["tiles"]):  # the image, tiled over the shard
            for o in range(0, image.n, step):
                e.table_load(ids[t * image.n + o: t * image.n + min(o + step, image.n)], image.rows[o:o + step])

    engines = []
    for _ in range(W + K):
        e = Engine(capacity, device=local_rank, full_sort=args.front_end == "full", radix_sort=args.front_end == "radix", **ranks_kw)
        load_pristine(e)
        e.reserve(n * (2 if world > 1 else 1), host_entry=(world == 1))
        engines.append(e)
    eng = engines[0]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # an explicit stream: the library treats handle 0 as "the ctx's own stream"
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    stream = side.cuda_stream

    d_in = [dev_batch(torch, dev, b) for b in batches]
    cap = n * (2 if world > 1 else 1)  # a shard receives ~n updates of a uniform batch; twice that is the slot size
    out = DevOut(torch, dev, cap)

    router = None
    if world > 1:
        router = shard.Router(world, rank, n, local_rank, recv_capacity=cap, key_bits=kb)

    def step_dev(i, last):
        """One step.  Sharded: merge batch i (already routed into slot i % 2), then route batch i + 1 while that
        merge runs - the routing never depends on the table."""
        if router is None:
            engines[i].merge_dev(d_in[i % N_BATCHES][1], out.cs, stream)
            return n
        m = router.merge(engines[i], i % 2, out.cs, stream)
        if not last:
            router.route(d_in[(i + 1) % N_BATCHES][1], (i + 1) % 2)
        return m

    # ---- warm-up; step 0 doubles as the parity check of the timed workload
    parity = {}
    sampler = ClockSampler(local_rank)
    if router is not None:
        router.route(d_in[0][1], 0)
    for i in range(W):
        m0 = step_dev(i, False)
        if i == 0 and not args.no_parity:
            eng.sync(stream)
            got = out.changes(m0)
            if world == 1:
                ok, detail = oracle_check(eng.cfg, image.rows, batches[0], got, lambda t: eng.table_read(t))
                parity["merge"] = "ok" if ok else "MISMATCH: " + detail
                parity["n"] = detail
            else:
                # what this shard received, in (source rank, arrival) order: every rank's first batch, regenerated
                recv = []
                for src in range(world):
                    b = batches[0] if src == rank else sharded_batch(image, n, synth.rng_for(3, salt=src * 16), args.keys, world, kb, capacity)
                    mine = np.nonzero(shard.owner_of(b.path_id, world, kb) == rank)[0]
                    recv.append(codec.Batch(shard.local_row(b.path_id[mine], world, kb), b.head[mine], b.clk[mine], b.val[mine]))
                rb = codec.Batch(*(np.concatenate([getattr(x, f) for x in recv]) for f in ("path_id", "head", "clk", "val")))
                uniq, rows, cb = compact_for_oracle(image.rows, image.n, rb)
                ok = rb.n == m0
                detail = f"received {m0}, expected {rb.n}"
                if ok:
                    ocfg = capi.make_config(max(len(uniq), 1), **ranks_kw)
                    ok, detail = oracle_check(ocfg, rows, cb, got, lambda t: eng.table_read(uniq[t.astype(np.int64)]))
                parity["shard"] = "ok" if ok else "MISMATCH: " + detail
                parity["n"] = "rank 0: " + detail
    eng.sync(stream)
    barrier()
    sampler.start()
    launches0 = sum(e.launch_count() for e in engines) + (router.launches if router else 0)
    sent0 = router.sent_bytes if router else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    merged = 0
    for i in range(K):
        merged += step_dev(W + i, i == K - 1)
    e1.record()
    barrier()
    launches = sum(e.launch_count() for e in engines) - launches0 + ((router.launches if router else 0))
    for e in engines:
        e.sync(stream)
    dev_ms = e0.elapsed_time(e1)
    sent = (router.sent_bytes - sent0) if router else 0
    acc_frac = float(out.n.item()) / max(1, (merged // K))
    ph = {"device": float(np.mean([engines[W + j].phase_ms("device") for j in range(K)]))}
    if world == 1:  # per-kernel split: a few extra steps with the event between the front end and the merge recorded
        Kp = min(K, 5)
        for e in engines[:Kp]:
            load_pristine(e)
            e.phase_events(True)
        for i in range(Kp):
            step_dev(i, False)
        eng.sync(stream)
        torch.cuda.synchronize()
        for name in ("sort", "merge"):
            ph[name] = float(np.mean([engines[j].phase_ms(name) for j in range(Kp)]))
        for e in engines[:Kp]:
            e.phase_events(False)

    # ---- the same shard WITHOUT routing: every rank merges batches it owns entirely (what one GPU does alone on this
    # shard shape); value / (N x this) is the efficiency of the routed job on config 3 itself
    unrouted = None
    if world > 1:
        Ku = min(K, 6)
        for e in engines[:Ku + 1]:
            load_pristine(e)
        lb = []
        for j in range(2):
            b = synth.make_batch(image, n, synth.rng_for(3, salt=5000 + rank * 16 + j), keys=args.keys)
            b.path_id[:] = synth.rng_for(3, salt=6000 + rank * 16 + j).integers(0, meta["tiles"], n).astype(np.uint64) * np.uint64(image.n) + b.path_id
            lb.append(dev_batch(torch, dev, b))
        engines[0].merge_dev(lb[0][1], out.cs, stream)
        engines[0].sync(stream)
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for i in range(Ku):
            engines[1 + i].merge_dev(lb[i % 2][1], out.cs, stream)
        u1.record()
        torch.cuda.synchronize()
        t = torch.tensor([u0.elapsed_time(u1) / Ku], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        unrouted = {"ms_per_step": float(t[0]), "per_gpu_field_merges_per_sec": n * F / (float(t[0]) * 1e-3), "steps": Ku,
                    "note": "local row ids, bb_merge_batch_dev only: the per-GPU rate on this shard shape with no exchange"}
    # ---- e2e through the reference-facing host entry points, pinned host buffers, copies inside the timed region.
    # The contexts of this section are created with BB_CFG_COMPACT_CHANGES (an accepted update that is stored exactly as
    # it came in gets no entry - the caller has it already): same decisions, table and (rebuilt) change set, about half
    # the bytes on the return link.
    def pinned(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory()

    for e in engines:
        e.close()
    Ke = min(K, 20 if world == 1 else 10)
    e_eng = []
    for _ in range(W + Ke):
        e = Engine(capacity, device=local_rank, compact_changes=True, **ranks_kw)
        load_pristine(e)
        e.reserve(cap, host_entry=True)
        e_eng.append(e)
    h_in = [tuple(pinned(x) for x in (b.path_id, b.head, b.clk, b.val)) for b in batches]
    hp = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
    h_ver, h_n, h_idx, h_head, h_clk, h_val = hp(4 * cap), hp(8), hp(4 * cap), hp(16 * cap), hp(32 * cap), hp(32 * cap)
    hcs = capi.BBChanges(cap=cap, verdict=h_ver.data_ptr(), n_changes=h_n.data_ptr(), idx=h_idx.data_ptr(),
                         head=h_head.data_ptr(), clk=h_clk.data_ptr(), val=h_val.data_ptr())
    hbs = [capi.BBBatch(n=n, path_id=a.data_ptr(), head=b_.data_ptr(), clk=c_.data_ptr(), val=d_.data_ptr())
           for a, b_, c_, d_ in h_in]

    def host_changes(m, rb):
        k = int(h_n.view(torch.int64)[0])
        return codec.Changes.from_verdicts(h_ver.numpy().view(np.uint32)[:m], h_idx.numpy().view(np.uint32)[:k],
                                           h_head.numpy().view(codec.HEAD_DTYPE)[:k], h_clk.numpy().view(np.uint32).reshape(-1, 8)[:k],
                                           h_val.numpy().view(np.uint64).reshape(-1, 4)[:k], rb)

    pcie = pcie_probe(torch, dev, n * 88, dist)
    if world == 1:
        for i in range(W):
            e_eng[i].merge_raw(hbs[i % N_BATCHES], hcs)
        if not args.no_parity:  # the host entry's output (chunked, three streams, compact) of step W-1 against the oracle too
            hg = host_changes(n, batches[(W - 1) % N_BATCHES])
            okh, detail = oracle_check(eng.cfg, image.rows, batches[(W - 1) % N_BATCHES], hg, lambda t: e_eng[W - 1].table_read(t))
            parity["host_entry"] = "ok" if okh else "MISMATCH: " + detail
        torch.cuda.synchronize()
        d2h = 0
        t0 = time.perf_counter()
        for i in range(Ke):
            e_eng[W + i].merge_raw(hbs[(W + i) % N_BATCHES], hcs)
            k = int(h_n.view(torch.int64)[0])
            d2h += 4 * n + 8 + k * 84
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": Ke * n * F / dt, "unit": "field-merges/s", "h2d_bytes_per_step": n * 88,
               "d2h_bytes_per_step": d2h // Ke, "ms_per_step": dt / Ke * 1e3, "steps": Ke,
               "api": "bb_merge_batch (pinned host buffers, synchronous; ctx with BB_CFG_COMPACT_CHANGES)", "pcie": pcie}
    else:
        # sharded: every rank's batch starts in pinned HOST memory; bb_router_merge_batch copies it in piece by piece,
        # routes every piece all-to-all over NVLink, merges on the owning shards and returns each shard's verdicts +
        # change entries to its rank's host buffers - all inside the timed region
        hrouter = shard.Router(world, rank, n, local_rank, recv_capacity=cap, key_bits=kb)
        PIECES = 4

        def step_e2e(i):
            m, counts = hrouter.merge_batch(e_eng[i], hbs[i % N_BATCHES], hcs, PIECES)
            return m, counts, 4 * m + 8 + 84 * int(h_n.view(torch.int64)[0])

        for i in range(W):
            m, counts, _ = step_e2e(i)
            if i == 0 and not args.no_parity:
                # what this shard received: piece by piece, inside a piece by source rank (include/bullet_b200.h)
                chunk = -(-n // PIECES)
                recv = []
                for j in range(PIECES):
                    for src in range(world):
                        b = batches[0] if src == rank else sharded_batch(image, n, synth.rng_for(3, salt=src * 16), args.keys, world, kb, capacity)
                        b = b.slice(j * chunk, min((j + 1) * chunk, n))
                        mine = np.nonzero(shard.owner_of(b.path_id, world, kb) == rank)[0]
                        recv.append(codec.Batch(shard.local_row(b.path_id[mine], world, kb), b.head[mine], b.clk[mine], b.val[mine]))
                rb = codec.Batch(*(np.concatenate([getattr(x, f) for x in recv]) for f in ("path_id", "head", "clk", "val")))
                okh = rb.n == m and counts.ravel().tolist() == [x.n for x in recv]
                detail = f"received {m}, expected {rb.n}"
                if okh:
                    uniq, rows, cb = compact_for_oracle(image.rows, image.n, rb)
                    ocfg = capi.make_config(max(len(uniq), 1), **ranks_kw)
                    okh, detail = oracle_check(ocfg, rows, cb, host_changes(m, rb), lambda t: e_eng[0].table_read(uniq[t.astype(np.int64)]))
                parity["host_entry"] = "ok" if okh else "MISMATCH: " + detail
        barrier()
        t0 = time.perf_counter()
        d2h = me = 0
        for i in range(Ke):
            m, _, b_ = step_e2e(W + i)
            me += m
            d2h += b_
        barrier()
        dt = time.perf_counter() - t0
        hrouter.close()
        t = torch.tensor([dt, float(me), float(d2h)], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e = {"value": float(t[1]) * F / float(tmax[0]), "unit": "field-merges/s", "h2d_bytes_per_step": n * 88,
               "d2h_bytes_per_step": int(float(t[2]) / world / Ke), "ms_per_step": float(tmax[0]) / Ke * 1e3, "steps": Ke,
               "api": f"bb_router_merge_batch (pinned host buffers in and out, {PIECES} pipelined pieces: H2D -> pack + NVLink all-to-all -> "
                      "merge on the owning shard -> D2H of that shard's verdicts + compact change set), per rank, max over ranks",
               "pcie": pcie}
    for e in e_eng:
        e.close()
    clocks = sampler.stop()

    # ---- max over ranks
    if dist is not None:
        t = torch.tensor([dev_ms, float(merged)], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, merged_total = float(tmax[0]), float(tsum[1])
        if not args.no_parity:
            bad = torch.tensor([0 if parity.get("shard") == "ok" else 1], device=dev)
            dist.all_reduce(bad)
            if int(bad.item()) and parity.get("shard") == "ok":
                parity["shard"] = "MISMATCH on another rank"
            parity["shards_checked"] = world
    else:
        merged_total = float(merged)

    # ---- roofline of the dominant kernel (k_merge_stage) and of the whole step, SURVEY 8d figure
    peak, peak_src = peaks()
    if world == 1:
        distinct = float(np.mean([np.unique(b.path_id).size for b in batches])) / n
    else:  # what a shard sees per update: ~n updates uniform over its `capacity` rows
        distinct = float(capacity) * (1.0 - np.exp(-float(n) / capacity)) / n
    bytes_per_update = 84 + 68 * acc_frac + 256 * distinct
    per_launch_updates = merged / K
    pipeline = bytes_per_update * per_launch_updates / (ph["device"] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "k_merge_stage", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_per_update * per_launch_updates, "bytes_per_update": bytes_per_update,
            "accepted_frac": acc_frac, "distinct_paths_per_update": distinct,
            "pipeline_achieved": pipeline, "pipeline_frac": pipeline / peak, "phase_ms": ph,
            "traffic": ncu_traffic("k_merge_stage", world == 1 and args.records == CFG2["records"] and n == CFG2["batch"]
                                   and args.keys == "uniform" and args.front_end == "group")}
    if "merge" in ph:
        roof["achieved"] = bytes_per_update * per_launch_updates / (ph["merge"] * 1e-3) / 1e9
        roof["kernel_ms"] = ph["merge"]
    else:  # sharded runs do not split the step: the whole device step stands in for the kernel
        roof["achieved"] = pipeline
        roof["kernel_ms"] = ph["device"]
    roof["frac"] = roof["achieved"] / peak

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.typed import TypedOracle

        orc = TypedOracle(eng.cfg)
        cout = capi.ChangeBuffers(n)
        done, dtc = 0, 0.0
        while dtc < args.cpu_seconds and done < 64:
            orc.table[:] = image.rows  # pristine table, not timed
            t0 = time.perf_counter()
            orc.merge(batches[done % N_BATCHES], out=cout)
            dtc += time.perf_counter() - t0
            done += 1
        cpu = {"value": done * n * F / dtc, "unit": "field-merges/s", "cores": 1, "kind": "port",
               "sample": f"{done} x {n}-update batches of the same workload on the same table, "
                         f"oracle/bullet_oracle.c single thread, {dtc:.1f} s", "reference_kind": REFERENCE_KIND}

    for e in engines:
        e.close()
    engines = []
    if router is not None:
        router.close()
    zipf = None
    if world == 1 and not args.no_zipf and args.keys == "uniform":
        zipf = run_zipf(args, local_rank, dev, image, stream, torch)
    query = None
    if not args.no_query:
        query = run_queries(args, rank, world, local_rank, dev, image, dist)
        if not args.no_parity:
            parity["query"] = "ok" if all(v == "ok" for v in query["parity"].values()) else "MISMATCH"
            if dist is not None:
                bad = torch.tensor([0 if parity["query"] == "ok" else 1], device=dev)
                dist.all_reduce(bad)
                if int(bad.item()):
                    parity["query"] = "MISMATCH"

    failed = [k for k, v in parity.items() if isinstance(v, str) and v.startswith("MISMATCH")]
    if zipf:
        failed += [f"zipf.{k}" for k in ("default", "indexed_age") if zipf[k]["parity"] != "ok"]
    if rank == 0:
        line = {
            "metric": "crdt_field_merges_per_sec", "value": merged_total * F / (dev_ms * 1e-3),
            "unit": "field-merges/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64",
            "data": "synthetic", "config": workload_config(args, world, meta),
            "updates_per_sec": merged_total / (dev_ms * 1e-3),
            "e2e": e2e, "gpu_launches": launches, "parity": parity,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "zipf": zipf, "query": query,
            "reference_kind": REFERENCE_KIND,
        }
        if world > 1:  # the all-to-all of rank 0, against the measured NVLink peer-copy rate (B200_PROFILING.md)
            line["unrouted_shard"] = unrouted
            line["routing_efficiency_on_config3"] = line["value"] / (world * unrouted["per_gpu_field_merges_per_sec"])
            gbs = sent / (dev_ms * 1e-3) / 1e9
            line["alltoall"] = {"sent_bytes_per_step": sent // K, "gb_per_s_per_gpu": gbs, "nvlink_peak": 770.0,
                                "frac": gbs / 770.0, "overlapped_with_merge": True,
                                "api": "bb_router_route_dev (counts through peer-mapped flags, then one pack kernel storing rows "
                                       "straight into the owners' receive slots over NVLink)"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if failed:
        print(f"[bench] PARITY FAILURE: {failed}: {parity} {zipf}", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
