#!/usr/bin/env python
"""bench.py - the BASELINE.json metric on BASELINE.json's config.

A "step" = one pass of the hot path over one synthetic batch: 1 M conflicting
updates (F=4 fields each => 4 M field-merges) merged into a resident table of
2.5 M records (10 M fields) per GPU  (BASELINE.json configs[1]; SURVEY.md 8d
"Config 2", uniform-key variant - the one the 8d roofline figure is worked on).

  value   field-merges/s, whole job, inputs already resident in HBM
          (bb_merge_batch_dev; CUDA events on the launching stream, max over ranks)
  e2e     same metric through the reference-facing C-ABI call bb_merge_batch with
          pinned HOST buffers: H2D of the batch and D2H of decisions + change set
          inside the timed region
  roofline      the dominant kernel (k_merge_stage): algorithmic bytes / its mean launch
                duration inside the timed region / measured HBM peak
  cpu_baseline  the typed C oracle (oracle/bullet_oracle.c, a restatement of the
                reference's JS: kind "port"), 1 thread, bounded sample
  --impl reference   the same oracle on every host thread (the reference itself is
                JavaScript and there is no JS engine on the box: see DESIGN.md)

N > 1 (torchrun, one rank per GPU): the table is sharded by path id (global table =
N x the per-GPU table), every rank submits its own 1 M batch, updates are routed to
their owner by the library's fused pack + all-to-all over NVLink peer memory and
merged there, routing of batch i+1 overlapping the merge of batch i; weak scaling
(per-GPU work fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints ONE line

N_RECORDS = 2_500_000   # per GPU: 10 M fields
BATCH = 1_000_000       # updates per step per GPU
N_BATCHES = 4           # distinct pre-generated batches, cycled
F = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=N_RECORDS)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--keys", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--query-records", type=int, default=100_000_000,
                    help="nodes per GPU for the index build + range/equals scans (BASELINE config 4)")
    ap.add_argument("--no-query", action="store_true")
    ap.add_argument("--merge-kernel", default="stage", choices=["stage", "pipe"],
                    help="k_merge_stage (one CTA per 128-update tile; default) or k_merge_pipe (BB_CFG_CTA_PIPE)")
    ap.add_argument("--front-end", default="group", choices=["group", "full", "radix"],
                    help="how a batch is grouped by path: the library default, BB_CFG_FULL_SORT or BB_CFG_RADIX_SORT")
    return ap.parse_args()


def ncu_traffic(kernel, expect_default):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu capture
    (profiles/traffic.json, written by scripts/make_profiles.py); null when the run is not the captured workload."""
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


def make_workload(args, rank, world=1):
    """Table image of one shard (every shard is loaded with the same image) and this rank's batches.
    Sharded runs keep the per-GPU workload of the single-GPU run (weak scaling): the global table has
    world x records rows, a batch row drawn for image row r goes to a uniformly drawn owner q as global
    path id r * world + q (owner = id % world, local row = id // world = r), so every shard still sees
    ~batch updates spread over all of its `records` rows, with clocks built against the row they hit."""
    from bullet_js_b200 import synth

    rng = synth.rng_for(2, salt=rank)
    table = synth.make_table(args.records, rng if world == 1 else synth.rng_for(2, salt=1000))
    batches = [synth.make_batch(table, args.batch, rng, keys=args.keys) for _ in range(N_BATCHES)]
    if world > 1:
        for b in batches:
            b.path_id[:] = b.path_id * np.uint64(world) + rng.integers(0, world, b.n).astype(np.uint64)
    return table, batches


def run_reference(args, rank, world):
    """The reference's algorithm on the host cores: typed C restatement, all threads."""
    if rank != 0:
        return
    from bullet_js_b200 import capi, synth
    from oracle.typed import TypedOracle

    table, batches = make_workload(args, 0)  # one rank's share: the single-GPU workload
    cores = os.cpu_count() or 1
    cfg = capi.make_config(args.records, **synth.synth_ranks(args.records))
    orc = TypedOracle(cfg)
    out = capi.ChangeBuffers(args.batch)
    dt = 0.0
    for i in range(args.warmup + args.steps):
        orc.table[:] = table.rows  # every step merges into the pristine table (not timed)
        t0 = time.perf_counter()
        orc.merge(batches[i % N_BATCHES], threads=cores, out=out)
        if i >= args.warmup:
            dt += time.perf_counter() - t0
    v = args.steps * args.batch * F / dt
    line = {
        "impl": "reference", "metric": "crdt_field_merges_per_sec", "value": v, "unit": "field-merges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64",
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": v, "unit": "field-merges/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x {args.batch} updates on the same table, path-sharded over {cores} threads"},
        "e2e": {"value": v, "unit": "field-merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is JavaScript; no JS engine on the box, so this is the C restatement (oracle/bullet_oracle.c)",
    }
    print(json.dumps(line))


def run_queries(args, rank, world, local_rank, dev, table, dist):
    """BASELINE config 4: index('users','age') build + range(20,30) + equals(role,'admin') over
    --query-records nodes per GPU (the 2.5 M-record image tiled), every rank scanning its shard;
    results all-gathered (counts, then padded payload) when world > 1."""
    import torch

    from bullet_js_b200 import capi, codec, synth
    from bullet_js_b200.engine import Engine

    nq = args.query_records
    eng = Engine(nq, device=local_rank, post_getdata=True, **synth.synth_ranks(args.records))
    base = np.arange(table.n, dtype=np.uint64)
    for off in range(0, nq, table.n):
        m = min(table.n, nq - off)
        eng.table_load(base[:m] + np.uint64(off), table.rows[:m])
    t0 = time.perf_counter()
    eng.index_create(0, extra_capacity=1 << 16)
    build_ms_age = eng.phase_ms("scan")
    eng.index_create(2, extra_capacity=1 << 16)
    build_ms_role = eng.phase_ms("scan")
    build_wall = time.perf_counter() - t0
    side = torch.cuda.current_stream(dev)
    stream = side.cuda_stream
    cap = nq
    hits = torch.zeros(cap, dtype=torch.int32, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    hs = capi.BBHits(cap=cap, node=hits.data_ptr(), n_dense=cnt.data_ptr(), n_extra=cnt[1:].data_ptr())
    import ctypes as C

    lo = capi.BBBound(num=20.0, rank=0, flags=0, reserved=0)
    hi = capi.BBBound(num=30.0, rank=0, flags=0, reserved=0)
    admin = codec.KEY_STR | 0

    def range_dev():
        eng._check(eng.lib.bb_query_range_dev(eng._h, 0, C.byref(lo), C.byref(hi), C.byref(hs), C.c_void_p(stream)))

    def equals_dev():
        eng._check(eng.lib.bb_query_equals_dev(eng._h, 2, admin, C.byref(hs), C.c_void_p(stream)))

    out = {}
    reps = 10
    for name, fn in (("range", range_dev), ("equals", equals_dev)):
        for _ in range(3):
            fn()
        eng.sync(stream)
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
        nh = int(cnt.sum().item())
        out[name] = {"ms": ms, "scan_ms": kernel_ms, "hits": nh}
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
    assert int(h_cnt.sum()) == out["range"]["hits"]
    if dist is not None:  # all-gather(v) of the result ids: counts first, then the padded payload
        t = torch.tensor([out["range"]["ms"], out["equals"]["ms"], e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["range"]["ms"], out["equals"]["ms"], e2e_ms = (float(x) for x in t)
        n_local = torch.tensor([out["range"]["hits"]], device=dev, dtype=torch.int64)
        counts = torch.zeros(world, device=dev, dtype=torch.int64)
        dist.all_gather_into_tensor(counts, n_local)
        mx = int(counts.max().item())
        range_dev()
        gids = hits[:mx].to(torch.int64) * world + rank  # local row -> global path id
        allh = torch.zeros(world * mx, device=dev, dtype=torch.int64)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(allh, gids)
        g1.record()
        torch.cuda.synchronize()
        out["allgather_ms"] = g0.elapsed_time(g1)
        out["allgather_bytes"] = int(world * mx * 8)
    peak, _ = peaks()
    res = {
        "workload": f"config4: index(age)+index(role) build, range(age,20,30), equals(role,'admin') over {nq} nodes/GPU",
        "nodes_per_gpu": nq, "n_gpus": world,
        "range_rows_per_sec": nq * world / (out["range"]["ms"] * 1e-3),
        "equals_rows_per_sec": nq * world / (out["equals"]["ms"] * 1e-3),
        "range_hits": out["range"]["hits"], "equals_hits": out["equals"]["hits"],
        "range_ms": out["range"]["ms"], "equals_ms": out["equals"]["ms"],
        "index_build_ms": {"age": build_ms_age, "role": build_ms_role, "wall_both": build_wall * 1e3},
        "index_build_rows_per_sec": nq * world / (build_ms_age * 1e-3),
        "e2e_range": {"rows_per_sec": nq * world / (e2e_ms * 1e-3), "ms": e2e_ms, "d2h_bytes": out["range"]["hits"] * 4 + 16,
                      "api": "bb_query_range (host hit buffer, synchronous)"},
        "roofline": {"bound": "hbm", "kernel": "k_index_scan", "unit": "GB/s", "peak": peak,
                     "achieved": (8.0 * nq + 4.0 * out["range"]["hits"]) / (out["range"]["scan_ms"] * 1e-3) / 1e9,
                     "bytes_per_row": 8.0 + 4.0 * out["range"]["hits"] / nq, "kernel_ms": out["range"]["scan_ms"],
                     "traffic": ncu_traffic("k_index_scan", world == 1 and nq == 100_000_000)},
    }
    res["roofline"]["frac"] = res["roofline"]["achieved"] / peak
    # the measured peak is a COPY (half reads, half writes); this kernel is a pure read stream and can exceed it:
    # also report it against the data-sheet HBM3e figure the profiling recipe quotes
    res["roofline"]["frac_of_nominal_7700"] = res["roofline"]["achieved"] / 7700.0
    for k in ("allgather_ms", "allgather_bytes"):
        if k in out:
            res[k] = out[k]
    eng.close()
    return res


def workload_config(args, world):
    return {
        "workload": f"config2: {args.records * F // 1_000_000}M-field table/GPU ({args.records} records x {F} fields, 128 B rows), "
                    f"{args.batch}-update conflicting batch/GPU/step, {args.keys} keys, clock mix 40/20/25/5/5/5",
        "records_per_gpu": args.records, "batch_per_gpu": args.batch, "fields": F, "peers": 8,
        "keys": args.keys, "front_end": args.front_end, "merge_kernel": args.merge_kernel, "hot_keys": args.keys == "zipf", "sharding": f"path_id % {world}" if world > 1 else "none",
        "l2": f"working set {args.records * 128 // 2**20} MiB table + {N_BATCHES} x {args.batch * 88 // 2**20} MiB batches > 126 MB L2",
    }


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    from bullet_js_b200 import capi, codec, synth
    from bullet_js_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bullet_js_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    table, batches = make_workload(args, rank, world)
    n, K, W = args.batch, args.steps, args.warmup
    # Every step merges into a PRISTINE copy of the table (one bb_ctx per step, all
    # loaded with the same image): the batches' clocks are built relative to that image,
    # so re-merging into an already-merged table would turn the workload into "all
    # historical".  It also means no step ever finds its table in L2.
    ids = np.arange(args.records, dtype=np.uint64)
    per_engine = args.records * (128 + 12) + args.batch * 120  # table + counters + per-batch scratch, bytes
    if (W + K) * per_engine > 0.8 * torch.cuda.get_device_properties(dev).total_memory:
        raise SystemExit(f"--steps {K}: {W + K} pristine tables of {per_engine >> 20} MiB do not fit this GPU; use fewer steps")
    engines = []
    for _ in range(W + K):
        e = Engine(args.records, device=local_rank, full_sort=args.front_end == "full",
                   radix_sort=args.front_end == "radix", cta_pipe=args.merge_kernel == "pipe", hot_keys=args.keys == "zipf", **synth.synth_ranks(args.records))
        e.table_load(ids, table.rows)
        e.reserve(args.batch * world, host_entry=(world == 1))
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

    # ---- device-resident inputs and outputs
    def to_dev(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)

    d_in = [(to_dev(b.path_id), to_dev(b.head), to_dev(b.clk), to_dev(b.val)) for b in batches]
    cap = n * (world if world > 1 else 1)  # a rank may receive more than it sent
    o_ver = torch.zeros(cap, dtype=torch.int32, device=dev)
    o_n = torch.zeros(1, dtype=torch.int64, device=dev)
    o_idx = torch.zeros(cap, dtype=torch.int32, device=dev)
    o_head = torch.zeros(cap * 16, dtype=torch.uint8, device=dev)
    o_clk = torch.zeros(cap * 32, dtype=torch.uint8, device=dev)
    o_val = torch.zeros(cap * 32, dtype=torch.uint8, device=dev)
    cs = capi.BBChanges(cap=cap, verdict=o_ver.data_ptr(), n_changes=o_n.data_ptr(), idx=o_idx.data_ptr(),
                        head=o_head.data_ptr(), clk=o_clk.data_ptr(), val=o_val.data_ptr())

    router = None
    if world > 1:
        from bullet_js_b200.shard import Router

        router = Router(world, rank, n, local_rank)
        r_in = [capi.BBBatch(n=n, path_id=p.data_ptr(), head=h.data_ptr(), clk=c.data_ptr(), val=v.data_ptr())
                for p, h, c, v in d_in]

    def step_dev(i, last):
        """One step.  Sharded: merge batch i (already routed into slot i % 2), then route batch
        i + 1 while that merge runs - the routing never depends on the table."""
        p, h, c, v = d_in[i % N_BATCHES]
        if router is None:
            bs = capi.BBBatch(n=n, path_id=p.data_ptr(), head=h.data_ptr(), clk=c.data_ptr(), val=v.data_ptr())
            engines[i].merge_dev(bs, cs, stream)
            return n
        m = router.merge(engines[i], i % 2, cs, stream)
        if not last:
            router.route(r_in[(i + 1) % N_BATCHES], (i + 1) % 2)
        return m

    sampler = ClockSampler(local_rank)
    if router is not None:
        router.route(r_in[0], 0)
    for i in range(W):
        step_dev(i, False)
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
    launches = sum(e.launch_count() for e in engines) - launches0 + (router.launches if router else 0)
    for e in engines:
        e.sync(stream)
    dev_ms = e0.elapsed_time(e1)
    sent = (router.sent_bytes - sent0) if router else 0
    # phase timings of the timed steps (events recorded inside the library on the same stream)
    ph = {"device": float(np.mean([engines[W + j].phase_ms("device") for j in range(K)]))}
    if world == 1:  # per-kernel split: a few extra steps with the event between the front end and the merge recorded
        Kp = min(K, 5)
        for e in engines[:Kp]:
            e.table_load(ids, table.rows)
            e.phase_events(True)
        for i in range(Kp):
            step_dev(i, False)
        eng.sync(stream)
        torch.cuda.synchronize()
        for name in ("sort", "merge"):
            ph[name] = float(np.mean([engines[j].phase_ms(name) for j in range(Kp)]))
        ph["device_with_phase_events"] = float(np.mean([engines[j].phase_ms("device") for j in range(Kp)]))
        for e in engines[:Kp]:
            e.phase_events(False)
    else:
        ph["sort"] = ph["merge"] = float("nan")
    acc_frac = float(o_n.item()) / max(1, (merged // K))

    # ---- e2e through bb_merge_batch with pinned host buffers
    def pinned(a):
        t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory()
        return t

    e2e = None
    if world == 1:
        h_in = [tuple(pinned(x) for x in (b.path_id, b.head, b.clk, b.val)) for b in batches]
        hp = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
        h_ver, h_n, h_idx, h_head, h_clk, h_val = hp(4 * n), hp(8), hp(4 * n), hp(16 * n), hp(32 * n), hp(32 * n)
        hcs = capi.BBChanges(cap=n, verdict=h_ver.data_ptr(), n_changes=h_n.data_ptr(), idx=h_idx.data_ptr(),
                             head=h_head.data_ptr(), clk=h_clk.data_ptr(), val=h_val.data_ptr())
        hbs = [capi.BBBatch(n=n, path_id=a.data_ptr(), head=b_.data_ptr(), clk=c_.data_ptr(), val=d_.data_ptr())
               for a, b_, c_, d_ in h_in]
        Ke = min(K, 20)
        for e in engines[: W + Ke]:
            e.table_load(ids, table.rows)  # pristine again
        for i in range(W):
            engines[i].merge_raw(hbs[i % N_BATCHES], hcs)
        torch.cuda.synchronize()
        d2h = 0
        t0 = time.perf_counter()
        for i in range(Ke):
            engines[W + i].merge_raw(hbs[(W + i) % N_BATCHES], hcs)
            k = int(h_n.view(torch.int64)[0])
            d2h += 4 * n + 8 + k * 84
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": Ke * n * F / dt, "unit": "field-merges/s", "h2d_bytes_per_step": n * 88,
               "d2h_bytes_per_step": d2h // Ke, "ms_per_step": dt / Ke * 1e3, "steps": Ke,
               "api": "bb_merge_batch (pinned host buffers, synchronous)"}
    elif world > 1:
        # sharded e2e: every rank's batch starts in pinned HOST memory; H2D, route (all-to-all over NVLink),
        # merge on the owning shards, then D2H of the decisions and the change set - all inside the timed region
        h_in = [tuple(pinned(x) for x in (b.path_id, b.head, b.clk, b.val)) for b in batches]
        hp = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
        h_ver, h_idx, h_head, h_clk, h_val = hp(4 * cap), hp(4 * cap), hp(16 * cap), hp(32 * cap), hp(32 * cap)
        stage = tuple(torch.empty_like(x) for x in d_in[0])  # device landing buffers of the H2D copies
        r_stage = capi.BBBatch(n=n, path_id=stage[0].data_ptr(), head=stage[1].data_ptr(), clk=stage[2].data_ptr(),
                               val=stage[3].data_ptr())
        Ke = min(K, 10)
        for e in engines[: W + Ke]:
            e.table_load(ids, table.rows)  # pristine again

        def step_e2e(i):
            for dst, src in zip(stage, h_in[i % N_BATCHES]):
                dst.copy_(src, non_blocking=True)
            router.route(r_stage, i % 2, stream)
            m = router.merge(engines[i], i % 2, cs, stream)
            k = int(o_n.item())  # D2H of the count (synchronises the stream)
            h_ver[: 4 * m].copy_(o_ver.view(torch.uint8)[: 4 * m], non_blocking=True)
            h_idx[: 4 * k].copy_(o_idx.view(torch.uint8)[: 4 * k], non_blocking=True)
            h_head[: 16 * k].copy_(o_head[: 16 * k], non_blocking=True)
            h_clk[: 32 * k].copy_(o_clk[: 32 * k], non_blocking=True)
            h_val[: 32 * k].copy_(o_val[: 32 * k], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return m, 4 * m + 8 + 84 * k

        for i in range(W):
            step_e2e(i)
        barrier()
        t0 = time.perf_counter()
        d2h = me = 0
        for i in range(Ke):
            m, b_ = step_e2e(W + i)
            me += m
            d2h += b_
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt, float(me), float(d2h)], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e = {"value": float(t[1]) * F / float(tmax[0]), "unit": "field-merges/s", "h2d_bytes_per_step": n * 88,
               "d2h_bytes_per_step": int(float(t[2]) / world / Ke), "ms_per_step": float(tmax[0]) / Ke * 1e3, "steps": Ke,
               "api": "pinned host batch -> H2D -> bb_router_route_dev -> bb_merge_batch_dev -> D2H of verdicts + change set, "
                      "per rank, max over ranks"}
    clocks = sampler.stop()

    # ---- max over ranks
    if dist is not None:
        t = torch.tensor([dev_ms, float(merged)], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, merged_total = float(tmax[0]), float(tsum[1])
    else:
        merged_total = float(merged)

    # ---- roofline of the dominant kernel (k_merge_stage), SURVEY 8d figure
    peak, peak_src = peaks()
    distinct = float(np.mean([np.unique(b.path_id).size for b in batches])) / n
    bytes_per_update = 84 + 68 * acc_frac + 256 * distinct
    per_launch_updates = merged / K
    achieved = bytes_per_update * per_launch_updates / (ph["merge"] * 1e-3) / 1e9
    pipeline = bytes_per_update * per_launch_updates / (ph["device"] * 1e-3) / 1e9

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.typed import TypedOracle

        orc = TypedOracle(eng.cfg)
        out = capi.ChangeBuffers(n)
        done, dtc = 0, 0.0
        while dtc < args.cpu_seconds and done < 64:
            orc.table[:] = table.rows  # pristine table, not timed
            t0 = time.perf_counter()
            orc.merge(batches[done % N_BATCHES], out=out)
            dtc += time.perf_counter() - t0
            done += 1
        cpu = {"value": done * n * F / dtc, "unit": "field-merges/s", "cores": 1, "kind": "port",
               "sample": f"{done} x {n}-update batches of the same workload on the same table, "
                         f"oracle/bullet_oracle.c single thread, {dtc:.1f} s"}

    for e in engines:
        e.close()
    engines = []
    query = None
    if not args.no_query:
        query = run_queries(args, rank, world, local_rank, dev, table, dist)

    if rank == 0:
        line = {
            "metric": "crdt_field_merges_per_sec", "value": merged_total * F / (dev_ms * 1e-3),
            "unit": "field-merges/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64",
            "data": "synthetic", "config": workload_config(args, world),
            "updates_per_sec": merged_total / (dev_ms * 1e-3),
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_merge_stage", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": ncu_traffic("k_merge_stage", world == 1 and args.records == N_RECORDS and args.batch == BATCH
                                                and args.keys == "uniform" and args.front_end == "group"
                                                and args.merge_kernel == "stage"),
                         "algorithmic_bytes_per_launch": bytes_per_update * per_launch_updates, "peak_source": peak_src,
                         "bytes_per_update": bytes_per_update, "accepted_frac": acc_frac,
                         "distinct_paths_per_update": distinct, "kernel_ms": ph["merge"],
                         "pipeline_achieved": pipeline, "pipeline_frac": pipeline / peak,
                         "phase_ms": ph},
            "cpu_baseline": cpu, "clocks": clocks, "query": query,
        }
        if world > 1:  # the all-to-all of rank 0, against the measured NVLink peer-copy rate (B200_PROFILING.md)
            gbs = sent / (dev_ms * 1e-3) / 1e9
            line["alltoall"] = {"sent_bytes_per_step": sent // K, "gb_per_s_per_gpu": gbs, "nvlink_peak": 770.0,
                                "frac": gbs / 770.0, "overlapped_with_merge": True, "last_route_ms": router.last_ms(),
                                "api": "bb_router_route_dev (counts all-gather over NCCL, then one pack kernel storing rows straight into the owners' receive slots over NVLink)"}
        print(json.dumps(line))
    for e in engines:
        e.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
