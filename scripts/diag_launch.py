"""Diagnose the gap between per-kernel (ncu) time and back-to-back step time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bullet_js_b200 import capi, synth
from bullet_js_b200.engine import Engine

N_REC, N = 2_500_000, 1_000_000
rng = synth.rng_for(2)
table = synth.make_table(N_REC, rng)
b = synth.make_batch(table, N, rng)
dev = torch.device("cuda:0")
ids = np.arange(N_REC, dtype=np.uint64)
K = 12
engines = []
for _ in range(K):
    e = Engine(N_REC, **synth.synth_ranks(N_REC)); e.table_load(ids, table.rows); e.reserve(N, False); engines.append(e)
t = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
p, h, c, v = t(b.path_id), t(b.head), t(b.clk), t(b.val)
o = [torch.zeros(N * k, dtype=torch.uint8, device=dev) for k in (1, 8, 4, 16, 32, 32)]
bs = capi.BBBatch(n=N, path_id=p.data_ptr(), head=h.data_ptr(), clk=c.data_ptr(), val=v.data_ptr())
cs = capi.BBChanges(cap=N, decision=o[0].data_ptr(), n_changes=o[1].data_ptr(), idx=o[2].data_ptr(), head=o[3].data_ptr(), clk=o[4].data_ptr(), val=o[5].data_ptr())

def run(label, stream, reload=True):
    if reload:
        for e in engines: e.table_load(ids, table.rows)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream() if stream != 0 or True else None
    t0 = time.perf_counter(); e0.record()
    for e in engines: e.merge_dev(bs, cs, stream)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    for e in engines: e.sync(stream)
    ph = {k: np.mean([e.phase_ms(k) for e in engines[2:]]) for k in ("sort", "merge", "compact", "device")}
    print(f"{label:28s} enqueue {1e3*(t1-t0)/K:.3f} ms/step  wall {1e3*(t2-t0)/K:.3f}  events {e0.elapsed_time(e1)/K:.3f}  phases " + " ".join(f"{k}={x:.3f}" for k, x in ph.items()), flush=True)

s_legacy = torch.cuda.current_stream().cuda_stream
print("torch current stream handle:", s_legacy)
side = torch.cuda.Stream()
run("legacy stream (warm)", s_legacy)
run("legacy stream", s_legacy)
run("ctx own streams", 0)
with torch.cuda.stream(side):
    run("torch side stream", side.cuda_stream)
# same engine repeatedly (table not pristine: timing only)
eng = engines[0]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(K): eng.merge_dev(bs, cs, side.cuda_stream)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"one engine x{K}: enqueue {1e3*(t1-t0)/K:.3f} wall {1e3*(t2-t0)/K:.3f} device {eng.phase_ms('device'):.3f}")
import threading
try:
    import pynvml; pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    stop = threading.Event()
    def samp():
        while not stop.is_set():
            pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetUtilizationRates(hnd); pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd); time.sleep(0.005)
    th = threading.Thread(target=samp, daemon=True); th.start()
    run("side stream + NVML sampler", side.cuda_stream)
    stop.set(); th.join()
    t0 = time.perf_counter(); pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM); t1 = time.perf_counter(); pynvml.nvmlDeviceGetUtilizationRates(hnd); t2 = time.perf_counter(); pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd); t3 = time.perf_counter()
    print(f"nvml call ms: clock {1e3*(t1-t0):.3f} util {1e3*(t2-t1):.3f} reasons {1e3*(t3-t2):.3f}")
except Exception as ex:
    print("nvml:", ex)
