"""tests/meshgen.py (the batched mesh simulator behind `bench.py --config mesh`, BASELINE config 5): its per-peer logs
replayed by a FRESH oracle in other batch sizes give the recorded expectations, replicas differ (the rule is
order-dependent), every decision class the mesh can produce shows up; (gpu) the CUDA path replays the logs."""
import numpy as np
import pytest

from bullet_js_b200 import capi, synth
from oracle.typed import TypedOracle
from tests import meshgen

N_REC, PEERS, ROUNDS, LOCAL, BATCH = 5000, 4, 4, 1500, 4000


@pytest.fixture(scope="module")
def mesh():
    image = synth.make_table(N_REC, synth.rng_for(5, salt=3))
    return image, meshgen.run_mesh_rounds(image, PEERS, ROUNDS, LOCAL, seed=11, batch=BATCH)


def replay(image, m, p, make, cut):
    eng = make(capi.make_config(image.n, local_peer=p, **synth.synth_ranks(image.n)))
    ids = np.arange(image.n, dtype=np.uint64)
    (eng.load if hasattr(eng, "load") else eng.table_load)(ids, image.rows)
    log = m["logs"][p]
    dec, entries = [], []
    for lo in range(0, log.n, cut):
        ch = eng.merge(log.slice(lo, min(lo + cut, log.n)))
        dec.append(ch.decision)
        entries.append((ch.idx.astype(np.int64) + lo, ch.head, ch.clk, ch.val))
    dec = np.concatenate(dec)
    gi = np.concatenate([e[0] for e in entries])
    gh, gc, gv = (np.concatenate([e[k] for e in entries]) for k in (1, 2, 3))
    for j, (n, hist, k, cs) in enumerate(m["expect"][p]):
        lo, hi = j * BATCH, j * BATCH + n
        a, b = np.searchsorted(gi, lo), np.searchsorted(gi, hi)
        assert np.bincount(dec[lo:hi], minlength=7)[:7].tolist() == hist, (p, j)
        assert b - a == k and meshgen.entries_checksum((gi[a:b] - lo).astype(np.uint32), gh[a:b], gc[a:b], gv[a:b]) == cs, (p, j)
    return eng


def test_logs_replay_to_the_recorded_expectations(mesh):
    image, m = mesh
    seen = np.zeros(7, np.int64)
    for p in range(PEERS):
        orc = replay(image, m, p, TypedOracle, 2777)  # another batching than the generator's rounds and the expectations'
        assert np.array_equal(orc.table, m["tables"][p])
        for _n, hist, _k, _cs in m["expect"][p]:
            seen += np.array(hist)
    assert seen[2] and seen[4] and seen[5] and seen[6]  # tie, dominating, historical (late deliveries), concurrent
    assert not np.array_equal(m["tables"][0], m["tables"][1])  # replicas differ: "converged" means "equals the replay"
    assert meshgen.entries_checksum(np.zeros(0, np.uint32), np.zeros(0, capi.codec.HEAD_DTYPE), np.zeros((0, 8), np.uint32),
                                    np.zeros((0, 4), np.uint64)) == 0


@pytest.mark.gpu
def test_gpu_replays_the_mesh_logs(mesh):
    from bullet_js_b200.engine import Engine

    image, m = mesh
    for p in range(PEERS):
        def make(cfg):
            return Engine(image.n, local_peer=p, **synth.synth_ranks(image.n))

        eng = replay(image, m, p, make, 3000)
        assert np.array_equal(eng.table_read(np.arange(image.n, dtype=np.uint64)), m["tables"][p])
        eng.close()
