import ctypes as C, json, sys, os
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from bullet_js_b200 import capi, synth
from bullet_js_b200.engine import Engine
n_rec, n = 2_500_000, 1_000_000
rng = synth.rng_for(2)
table = synth.make_table(n_rec, rng)
batches = [synth.make_batch(table, n, rng) for _ in range(2)]
dev = torch.device('cuda', 0)
ids = np.arange(n_rec, dtype=np.uint64)
engs = []
for _ in range(6):
    e = Engine(n_rec, **synth.synth_ranks(n_rec)); e.table_load(ids, table.rows); e.reserve(n, host_entry=False); engs.append(e)
to_dev = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
d_in = [(to_dev(b.path_id), to_dev(b.head), to_dev(b.clk), to_dev(b.val)) for b in batches]
o = [torch.zeros(n * k, dtype=torch.uint8, device=dev) for k in (4, 8, 4, 16, 32, 32)]
cs = capi.BBChanges(cap=n, verdict=o[0].data_ptr(), n_changes=o[1].data_ptr(), idx=o[2].data_ptr(), head=o[3].data_ptr(), clk=o[4].data_ptr(), val=o[5].data_ptr())
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side)
lib = capi.load(); lib.bb_debug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
for i, e in enumerate(engs):
    p, h, c, v = d_in[i % 2]
    bs = capi.BBBatch(n=n, path_id=p.data_ptr(), head=h.data_ptr(), clk=c.data_ptr(), val=v.data_ptr())
    e.merge_dev(bs, cs, side.cuda_stream)
torch.cuda.synchronize()
for e in engs:
    out = (C.c_double * 6)()
    rc = lib.bb_debug_timeline(e._h, out)
    print(rc, ["%.1f" % (x / 1e3) for x in out])
