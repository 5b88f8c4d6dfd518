"""Design validation for SURVEY 8f-1 (exact Map / Set result order on the device), on the CPU.

The scheme a device implementation would use:
  * every index entry (bucket key, node) carries the sequence number of the add that inserted it (a re-add of an
    existing entry changes nothing; delete + add gives it a new one);
  * the hook of a batch emits its EFFECTIVE adds / removes as events (key, 2*seq + is_add, +-1) - effectiveness only
    depends on the node's own entries, which the merge kernel's per-path replay knows;
  * per bucket, the events sorted by sequence are scanned with the running entry count: the bucket's creation
    sequence is that of the add that followed the last time the count was zero (a bucket dies when its Set empties
    and is re-created at the END of the Map order by the next add, src/bullet-query.js:89-93, 110-116);
  * a query sorts its hits by (bucket creation sequence, entry sequence).
This test attaches those tags to the literal oracle's index (by wrapping its two mutators) and checks that the order
they induce IS the Map / Set iteration order of the reference's data structures, on streams that the golden
fixtures tie to the reference itself - including index builds over an existing store and late index builds.
"""
import itertools

import pytest

from oracle.js_literal import RefBullet, RefQuery
from oracle.jsvalue import UNDEFINED
from tests import golden_io, streamgen
from tests.test_oracle_query import BOUNDS

STREAMS = golden_io.load("streams.json.gz")["cases"]


class Tagged(RefQuery):
    """RefQuery + the event log a device implementation would produce (it never looks at dict order)."""

    def __init__(self, bullet):
        super().__init__(bullet)
        self.entry_seq = {}   # (id(index), bucket key, node) -> sequence of the add that inserted the entry
        self.events = {}      # id(index) -> {bucket key: [(2 * seq + is_add, +-1)]}
        self.clock = 0        # one tick per mutator call: stands for (update sequence, remove-before-add)

    def _addToIndex(self, index, value, nodePath):
        self.clock += 1
        if value is None or value is UNDEFINED:
            return
        k = self._getIndexableValue(value)
        had = k in index and nodePath in index[k]
        super()._addToIndex(index, value, nodePath)
        if not had:  # effective add
            self.entry_seq[(id(index), k, nodePath)] = self.clock
            self.events.setdefault(id(index), {}).setdefault(k, []).append((self.clock, +1))

    def _removeFromIndex(self, index, value, nodePath):
        self.clock += 1
        if value is None or value is UNDEFINED:
            return
        k = self._getIndexableValue(value)
        had = k in index and nodePath in index[k]
        super()._removeFromIndex(index, value, nodePath)
        if had:  # effective remove
            del self.entry_seq[(id(index), k, nodePath)]
            self.events.setdefault(id(index), {}).setdefault(k, []).append((self.clock, -1))

    def bucket_creation(self, index):
        """The per-bucket temporal scan: creation sequence of every LIVE bucket from its event list alone."""
        out = {}
        for k, evs in self.events.get(id(index), {}).items():
            count, created = 0, None
            for seq, d in sorted(evs):
                if d > 0 and count == 0:
                    created = seq
                count += d
            if count > 0:
                out[k] = created
        return out

    def ordered(self, index, want_bucket):
        """Hits of the buckets `want_bucket` accepts, ordered by (bucket creation, entry sequence) only."""
        created = self.bucket_creation(index)
        hits = [(created[k], s, node) for (i, k, node), s in self.entry_seq.items() if i == id(index) and want_bucket(k)]
        return [node for _, _, node in sorted(hits)]


def replay(case):
    ref = RefBullet("p0", enable_indexing=True)
    ref.query = Tagged(ref)
    for f in case["index_fields"]:
        ref.index("users", f)
    for k, op in enumerate(golden_io.ops_of(case)):
        for f, at in case["late_index"].items():
            if at == k:
                ref.index("users", f)  # built from the store: entries enter in Object.entries order
        streamgen.apply_op(ref, op)
    return ref


@pytest.mark.parametrize("k", [i for i, c in enumerate(STREAMS) if c["index_fields"] or c["late_index"]][:4])
def test_sequence_tags_reproduce_map_and_set_order(k):
    case = STREAMS[k]
    ref = replay(case)
    q = ref.query
    for key, index in q.indices.items():
        assert list(q.bucket_creation(index)) and set(q.bucket_creation(index)) == set(index)
        # Map order == ascending creation sequence; Set order == ascending entry sequence
        created = q.bucket_creation(index)
        assert sorted(index, key=created.get) == list(index), key
        for bk, paths in index.items():
            assert sorted(paths, key=lambda p: q.entry_seq[(id(index), bk, p)]) == list(paths), (key, bk)
        # and the golden index dump of the reference itself is that order
        assert [[bk, list(p)] for bk, p in index.items()] == case["index"][key]
    pairs = list(itertools.product(BOUNDS, BOUNDS))
    for name, want in case["queries"].items():
        index = q.indices[f"users:{name}"]
        checked = 0
        for (lo, hi), expect in list(zip(pairs, want["range"]))[::7]:  # every 7th pair of bounds: all shapes, a few seconds
            if expect:
                selected = {b for b in index if _in_range(ref, name, b, lo, hi)}
                assert q.ordered(index, selected.__contains__) == expect, (name, lo, hi)
                checked += 1
        assert checked > 3


def _in_range(ref, name, bucket, lo, hi):
    """Does range(lo, hi) select this bucket?  Asked of the oracle one bucket at a time."""
    probe = RefQuery(ref)
    probe.indices = {f"users:{name}": {bucket: {"x": True}}}
    probe.indexedPaths = {"users": True}
    return probe.range("users", name, lo, hi) == ["x"]
