"""Host-side typed encoding of JS values, clocks, updates and rows.

This is the packing step the north-star assigns to the host ("packs update
batches into struct-of-arrays buffers: interned path IDs, field IDs, state
clocks, value hashes or offsets"); the bit layout is the one documented in
include/bullet_b200.h.  JS values are modelled as: float (number), str, bool,
None (null), dict (flat record, insertion-ordered own keys).

Reference behaviour being encoded (paths relative to the reference repo):
  value shapes   src/bullet-network-sync.js:551-569 (entries), src/bullet-network.js:332-346
  clock objects  src/bullet-crt.js:33-60 (ordered own keys peer -> count)
"""
from __future__ import annotations

import bisect
import struct
from dataclasses import dataclass, field

import numpy as np

MAX_PEERS = 8
MAX_FIELDS = 4

TAG_ABSENT, TAG_NUM, TAG_STR, TAG_BOOL, TAG_NULL = 0, 1, 2, 3, 4
KIND_NONE, KIND_OBJ, KIND_PRIM = 0, 1, 2
HDR_FLAVOUR_NET = 1
HDR_KIND_SHIFT, HDR_TAG_SHIFT, HDR_ORDER_SHIFT = 1, 8, 32
ROW_M_PRESENT, ROW_V_PRESENT, ROW_ALIAS = 1, 2, 4
CFG_POST_GETDATA = 1
CFG_ORDERED_CHANGES = 2
CFG_RADIX_SORT = 4
CFG_FULL_SORT = 8
CFG_HOT_KEYS = 64
CFG_COMPACT_CHANGES = 128
CFG_TRACK_MODIFIED = 256
CFG_EXACT_ORDER = 512
COLLECT_FILTER_RECORDS = 1

DEC_NO_CURRENT, DEC_IDENTICAL, DEC_TIE_INCOMING, DEC_TIE_CURRENT = 0, 1, 2, 3
DEC_INCOMING, DEC_HISTORICAL, DEC_CONCURRENT = 4, 5, 6
DEC_ACCEPTED_MASK = 0x55

HEAD_DTYPE = np.dtype([("hdr", "<u8"), ("clk_order", "<u4"), ("user", "<u4")])
ROW_DTYPE = np.dtype(
    [
        ("val", "<u8", (MAX_FIELDS,)),
        ("m_cnt", "<u4", (MAX_PEERS,)),
        ("v_cnt", "<u4", (MAX_PEERS,)),
        ("m_order", "<u4"),
        ("v_order", "<u4"),
        ("hdr", "<u8"),
        ("flags", "<u4"),
        ("xcnt", "<u4"),
        ("cseq", "<u8"),
    ]
)
assert HEAD_DTYPE.itemsize == 16 and ROW_DTYPE.itemsize == 128


class DomainError(ValueError):
    """Input outside the bit-exact typed domain (SURVEY.md 8a restrictions)."""


def _utf16_key(s: str) -> bytes:
    return s.encode("utf-16-be", "surrogatepass")


def _js_numeric_string(s: str) -> bool:
    """True if Number(s) is not NaN (such strings may not enter the dictionary:
    range() would treat them as numbers, src/bullet-query.js:240-243)."""
    import re

    t = s.strip(_JS_WS)
    if t == "" or t in ("Infinity", "+Infinity", "-Infinity"):
        return True
    if re.fullmatch(r"0[xX][0-9a-fA-F]+|0[oO][0-7]+|0[bB][01]+", t):
        return True
    return re.fullmatch(r"[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", t) is not None


_JS_WS = (
    "\t\n\v\f\r \u00a0\u1680\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008"
    "\u2009\u200a\u2028\u2029\u202f\u205f\u3000\ufeff"
)
_FORBIDDEN = ("true", "false", "NaN", "[object Object]")


def js_string_to_number(s: str) -> float:
    """ToNumber of a JS string (ECMA-262 7.1.4.1.1): what `<`, `>=` and Number() apply."""
    import re

    t = s.strip(_JS_WS)
    if t == "":
        return 0.0
    if t in ("Infinity", "+Infinity"):
        return float("inf")
    if t == "-Infinity":
        return float("-inf")
    m = re.fullmatch(r"0[xX]([0-9a-fA-F]+)|0[oO]([0-7]+)|0[bB]([01]+)", t)
    if m:
        return float(int(m.group(1) or m.group(2) or m.group(3), 16 if m.group(1) else 8 if m.group(2) else 2))
    if re.fullmatch(r"[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", t):
        return float(t)
    return float("nan")


def js_number_to_string(x: float) -> str:
    """String(x) for a JS number (Number::toString, radix 10): shortest digits that
    round-trip, plain notation for 1e-7 <= |x| < 1e21, exponent form otherwise."""
    if x != x:
        return "NaN"
    if x == 0:
        return "0"
    if x < 0:
        return "-" + js_number_to_string(-x)
    if x == float("inf"):
        return "Infinity"
    mant, _, e10 = ("%r" % x).partition("e")           # repr() is shortest round-trip
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    point = len(ip) + (int(e10) if e10 else 0)           # x = 0.<ip fp> * 10**point before stripping
    point -= len(ip + fp) - len((ip + fp).lstrip("0"))   # leading zeros removed
    digits = digits.rstrip("0") or "0"
    k, n = len(digits), point
    if k <= n <= 21:
        return digits + "0" * (n - k)
    if 0 < n <= 21:
        return digits[:n] + "." + digits[n:]
    if -6 < n <= 0:
        return "0." + "0" * (-n) + digits
    e = n - 1
    es = ("+" if e > 0 else "-") + str(abs(e))
    return (digits if k == 1 else digits[0] + "." + digits[1:]) + "e" + es


KEY_NAN = 0x7FF8000000000000
KEY_STR = 0xFFF9000000000000
KEY_BOOL = 0xFFFA000000000000
KEY_NONE = 0xFFFFFFFFFFFFFFFF
BOUND_IS_STRING, BOUND_TRUE, BOUND_FALSE, BOUND_NAN = 1, 2, 4, 8

BOUND_DTYPE = np.dtype([("num", "<f8"), ("rank", "<u8"), ("flags", "<u4"), ("reserved", "<u4")])


class StringDict:
    """Order-preserving dictionary: id order == UTF-16 code-unit order."""

    def __init__(self, strings):
        uniq = sorted(set(strings), key=_utf16_key)
        for s in uniq:
            if _js_numeric_string(s) or s in _FORBIDDEN:
                raise DomainError(f"string {s!r} cannot be dictionary-encoded bit-exactly")
        self.strings = uniq
        self._keys = [_utf16_key(s) for s in uniq]
        self._ids = {s: i for i, s in enumerate(uniq)}

    def id(self, s: str) -> int:
        try:
            return self._ids[s]
        except KeyError:
            raise DomainError(f"string {s!r} is not in the dictionary") from None

    def string(self, i: int) -> str:
        return self.strings[int(i)]

    def rank(self, s: str) -> int:
        """Smallest id whose string is greater than s."""
        return bisect.bisect_right(self._keys, _utf16_key(s))

    def __len__(self):
        return len(self.strings)


class Interner:
    """Dense ids in first-seen order (paths)."""

    def __init__(self):
        self.ids: dict[str, int] = {}
        self.names: list[str] = []

    def id(self, name: str) -> int:
        i = self.ids.get(name)
        if i is None:
            i = len(self.names)
            self.ids[name] = i
            self.names.append(name)
        return i

    def name(self, i: int) -> str:
        return self.names[int(i)]

    def __len__(self):
        return len(self.names)


def _f64_bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def _bits_f64(b: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", int(b)))[0]


@dataclass
class Schema:
    """Field-name -> slot, peer-id -> clock slot and the string dictionary."""

    fields: list[str]
    peers: list[str]
    strings: StringDict
    local_peer: str
    paths: Interner = field(default_factory=Interner)

    def __post_init__(self):
        if len(self.fields) > MAX_FIELDS:
            raise DomainError(f"more than {MAX_FIELDS} fields per record")
        if len(self.peers) > MAX_PEERS:
            raise DomainError(f"more than {MAX_PEERS} peers per clock")
        self._fslot = {f: i for i, f in enumerate(self.fields)}
        self._pslot = {p: i for i, p in enumerate(self.peers)}
        if self.local_peer not in self._pslot:
            raise DomainError("local peer must have a clock slot")

    # ---- primitives
    def enc_prim(self, v):
        if isinstance(v, bool):
            return TAG_BOOL, int(v)
        if v is None:
            return TAG_NULL, 0
        if isinstance(v, (int, float)):
            return TAG_NUM, _f64_bits(float(v))
        if isinstance(v, str):
            return TAG_STR, self.strings.id(v)
        raise DomainError(f"unsupported primitive {v!r}")

    def dec_prim(self, tag, pay):
        if tag == TAG_NUM:
            return _bits_f64(pay)
        if tag == TAG_STR:
            return self.strings.string(pay)
        if tag == TAG_BOOL:
            return bool(pay)
        if tag == TAG_NULL:
            return None
        raise ValueError(f"bad tag {tag}")

    # ---- whole values
    def enc_value(self, v):
        """-> (hdr bits without flavour, [payload]*MAX_FIELDS)."""
        val = [0] * MAX_FIELDS
        if isinstance(v, dict):
            hdr = KIND_OBJ << HDR_KIND_SHIFT
            for i, (k, x) in enumerate(v.items()):
                f = self._fslot.get(k)
                if f is None:
                    raise DomainError(f"field {k!r} not in schema")
                if isinstance(x, dict):
                    raise DomainError("nested objects are outside the typed domain")
                tag, pay = self.enc_prim(x)
                hdr |= tag << (HDR_TAG_SHIFT + 3 * f)
                hdr |= f << (HDR_ORDER_SHIFT + 4 * i)
                val[f] = pay
            return hdr, val
        tag, pay = self.enc_prim(v)
        val[0] = pay
        return (KIND_PRIM << HDR_KIND_SHIFT) | (tag << HDR_TAG_SHIFT), val

    def dec_value(self, hdr, val):
        hdr = int(hdr)
        kind = (hdr >> HDR_KIND_SHIFT) & 3
        if kind == KIND_NONE:
            return None  # caller distinguishes via kind
        if kind == KIND_PRIM:
            return self.dec_prim((hdr >> HDR_TAG_SHIFT) & 7, int(val[0]))
        tags = [(hdr >> (HDR_TAG_SHIFT + 3 * f)) & 7 for f in range(MAX_FIELDS)]
        n = sum(t != 0 for t in tags)
        out = {}
        for i in range(n):
            f = (hdr >> (HDR_ORDER_SHIFT + 4 * i)) & 0xF
            out[self.fields[f]] = self.dec_prim(tags[f], int(val[f]))
        return out

    # ---- clocks
    def enc_clock(self, clock: dict):
        cnt = [0] * MAX_PEERS
        order = 0
        for i, (p, c) in enumerate(clock.items()):
            s = self._pslot.get(p)
            if s is None:
                raise DomainError(f"peer {p!r} has no clock slot")
            c = int(c)
            if c < 1 or c > 0xFFFFFFFF:
                raise DomainError("clock entries must be integers >= 1")
            cnt[s] = c
            order |= s << (4 * i)
        return cnt, order

    def dec_clock(self, cnt, order) -> dict:
        order = int(order)
        n = sum(int(c) != 0 for c in cnt)
        out = {}
        for i in range(n):
            s = (order >> (4 * i)) & 0xF
            out[self.peers[s]] = float(cnt[s])
        return out

    # ---- query arguments (src/bullet-query.js:126-131, 238-252)
    def index_key(self, value):
        """The 64-bit key of String(value) (include/bullet_b200.h), or None when no
        stored value can have that string.  equals/count are type-blind: 25 and "25"
        name the same bucket (query:130)."""
        if isinstance(value, bool):
            return KEY_BOOL | int(value)
        if isinstance(value, (int, float)):
            x = float(value)
            if x != x:
                return KEY_NAN
            return 0 if x == 0 else _f64_bits(x)
        if value is None:
            value = "null"
        if isinstance(value, dict):
            return None  # JSON.stringify(object): nested values are outside the typed domain
        if not isinstance(value, str):
            raise DomainError(f"unsupported query value {value!r}")
        if value in ("true", "false"):
            return KEY_BOOL | int(value == "true")
        if value == "NaN":
            return KEY_NAN
        x = js_string_to_number(value)
        if x == x and js_number_to_string(x) == value:
            return 0 if x == 0 else _f64_bits(x)
        i = self.strings._ids.get(value)
        return None if i is None else KEY_STR | i

    def bound(self, value, upper: bool):
        """One side of range() as a bb_bound record (query:246-251)."""
        b = np.zeros((), BOUND_DTYPE)
        if isinstance(value, str):
            k = _utf16_key(value)
            b["num"] = js_string_to_number(value)
            b["rank"] = (bisect.bisect_right if upper else bisect.bisect_left)(self.strings._keys, k)
            fl = BOUND_IS_STRING
            for name, bit in (("true", BOUND_TRUE), ("false", BOUND_FALSE), ("NaN", BOUND_NAN)):
                kk = _utf16_key(name)
                if (kk <= k) if upper else (kk >= k):
                    fl |= bit
            b["flags"] = fl
        elif isinstance(value, bool):
            b["num"] = float(value)
        elif value is None:
            b["num"] = 0.0  # ToNumber(null)
        elif isinstance(value, (int, float)):
            b["num"] = float(value)
        else:
            b["num"] = float("nan")  # objects: ToNumber("[object Object]")
        return b

    def config_ranks(self):
        return dict(
            rank_object=self.strings.rank("[object Object]"),
            rank_true=self.strings.rank("true"),
            rank_false=self.strings.rank("false"),
            rank_nan=self.strings.rank("NaN"),
        )


@dataclass
class Batch:
    """Struct-of-arrays update batch (bb_batch)."""

    path_id: np.ndarray  # u64[n]
    head: np.ndarray     # HEAD_DTYPE[n]
    clk: np.ndarray      # u32[n, 8]
    val: np.ndarray      # u64[n, 4]

    @property
    def n(self):
        return int(self.path_id.shape[0])

    @staticmethod
    def empty(n: int) -> "Batch":
        return Batch(
            np.zeros(n, np.uint64), np.zeros(n, HEAD_DTYPE),
            np.zeros((n, MAX_PEERS), np.uint32), np.zeros((n, MAX_FIELDS), np.uint64),
        )

    def slice(self, lo, hi) -> "Batch":
        return Batch(self.path_id[lo:hi], self.head[lo:hi], self.clk[lo:hi], self.val[lo:hi])


NO_SLOT = 0x1FFFFFFF
SLOT_ECHO = 0x1FFFFFFE  # BB_CFG_COMPACT_CHANGES: accepted, the entry is the update itself


@dataclass
class Changes:
    """Decisions + emitted change set in ARRIVAL order (the reference's order)."""

    decision: np.ndarray  # u8[n]
    idx: np.ndarray       # u32[k]
    head: np.ndarray      # HEAD_DTYPE[k]
    clk: np.ndarray       # u32[k, 8]
    val: np.ndarray       # u64[k, 4]

    @staticmethod
    def from_verdicts(verdict, idx, head, clk, val, batch: "Batch | None" = None) -> "Changes":
        """bb_changes -> arrival order: walk verdict[] and follow the slots (the
        library stores entries path-major; include/bullet_b200.h).  With BB_CFG_COMPACT_CHANGES an
        accepted update whose entry would repeat it bit for bit carries SLOT_ECHO instead of a slot:
        its entry is rebuilt from `batch` (the caller's own input), as the header defines it."""
        verdict = np.asarray(verdict, np.uint32)
        decision = (verdict >> 29).astype(np.uint8)
        slot = verdict & NO_SLOT
        echo = slot == SLOT_ECHO
        emitted = (slot != NO_SLOT) & ~echo
        acc = np.nonzero(emitted)[0]
        order = slot[acc].astype(np.int64)
        k = len(idx)
        if acc.size != k or (k and (np.sort(order) != np.arange(k)).any()):
            raise ValueError("verdict slots are not a permutation of the change set")
        if k and not np.array_equal(np.asarray(idx)[order], acc.astype(np.uint32)):
            raise ValueError("verdict slot does not point at the update's own entry")
        if not echo.any():
            return Changes(decision, np.asarray(idx)[order].copy(), np.asarray(head)[order].copy(),
                           np.asarray(clk)[order].copy(), np.asarray(val)[order].copy())
        if batch is None:
            raise ValueError("compact change set: the input batch is needed to rebuild the echoed entries")
        every = np.nonzero(emitted | echo)[0]
        o_head = np.zeros(every.size, HEAD_DTYPE)
        o_clk = np.zeros((every.size, MAX_PEERS), np.uint32)
        o_val = np.zeros((every.size, MAX_FIELDS), np.uint64)
        is_echo = echo[every]
        src = every[is_echo]
        o_head[is_echo] = batch.head[src]
        o_head["hdr"][is_echo] &= ~np.uint64(HDR_FLAVOUR_NET)
        o_clk[is_echo] = batch.clk[src]
        o_val[is_echo] = batch.val[src]
        o_head[~is_echo] = np.asarray(head)[order]
        o_clk[~is_echo] = np.asarray(clk)[order]
        o_val[~is_echo] = np.asarray(val)[order]
        return Changes(decision, every.astype(np.uint32), o_head, o_clk, o_val)

    def same_as(self, other: "Changes") -> bool:
        return (
            np.array_equal(self.decision, other.decision)
            and np.array_equal(self.idx, other.idx)
            and np.array_equal(self.head, other.head)
            and np.array_equal(self.clk, other.clk)
            and np.array_equal(self.val, other.val)
        )


def encode_updates(schema: Schema, updates) -> Batch:
    """updates: iterable of (path, value, clock_or_None). A clock makes the update
    the network-with-clock flavour when the value is an object (crt:339-344);
    primitives never carry one (sync:560-563)."""
    updates = list(updates)
    b = Batch.empty(len(updates))
    for i, (path, value, clock) in enumerate(updates):
        b.path_id[i] = schema.paths.id(path)
        hdr, val = schema.enc_value(value)
        if clock is not None and isinstance(value, dict):
            cnt, order = schema.enc_clock(clock)
            hdr |= HDR_FLAVOUR_NET
            b.clk[i] = cnt
            b.head[i]["clk_order"] = order
        b.head[i]["hdr"] = hdr
        b.val[i] = val
    return b


def decode_changes(schema: Schema, batch: Batch, ch: Changes):
    out = []
    for k in range(len(ch.idx)):
        i = int(ch.idx[k])
        out.append(
            dict(
                seq=i,
                path=schema.paths.name(batch.path_id[i]),
                value=schema.dec_value(ch.head[k]["hdr"], ch.val[k]),
                vectorClock=schema.dec_clock(ch.clk[k], ch.head[k]["clk_order"]),
            )
        )
    return out


def decode_row(schema: Schema, row):
    """-> dict(kind, value, M, V, alias, cseq) with clocks None when absent."""
    hdr = int(row["hdr"])
    flags = int(row["flags"])
    kind = (hdr >> HDR_KIND_SHIFT) & 3
    return dict(
        kind=kind,
        value=schema.dec_value(hdr, row["val"]) if kind != KIND_NONE else None,
        M=schema.dec_clock(row["m_cnt"], row["m_order"]) if flags & ROW_M_PRESENT else None,
        V=schema.dec_clock(row["v_cnt"], row["v_order"]) if flags & ROW_V_PRESENT else None,
        alias=bool(flags & ROW_ALIAS),
        cseq=int(row["cseq"]),
    )
