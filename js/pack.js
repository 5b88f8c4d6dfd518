/**
 * pack.js - the Node host's half of the batch boundary (include/bullet_b200.h): interned path ids, field and
 * clock slots, the order-preserving string dictionary, and the struct-of-arrays packing of a batch of updates
 * (bb_batch) / unpacking of its change set (bb_changes).  Plain typed arrays; 64-bit words are written as two
 * little-endian 32-bit halves, so no BigInt is needed.  Byte-for-byte what bullet_js_b200/codec.py produces
 * (tests/test_js_shim.py compares the buffers).
 *
 * Typed domain (SURVEY.md 8a): flat records over the schema's fields; values are numbers, booleans, null or
 * dictionary strings; clock entries are integers >= 1 of the schema's peers.  Anything else throws DomainError:
 * there is no CPU fallback, such collections stay with the stock BulletCRT.
 */
const TAG_ABSENT = 0, TAG_NUM = 1, TAG_STR = 2, TAG_BOOL = 3, TAG_NULL = 4;
const KIND_NONE = 0, KIND_OBJ = 1, KIND_PRIM = 2;
const HDR_FLAVOUR_NET = 1, HDR_KIND_SHIFT = 1, HDR_TAG_SHIFT = 8;
const NO_SLOT = 0x1fffffff;
const MAX_PEERS = 8, MAX_FIELDS = 4;
const FORBIDDEN = ["true", "false", "NaN", "[object Object]"];
const KEY_NAN_HI = 0x7ff80000, KEY_STR_HI = 0xfff90000, KEY_BOOL_HI = 0xfffa0000; // high words of BB_KEY_*
const BOUND_IS_STRING = 1, BOUND_TRUE = 2, BOUND_FALSE = 4, BOUND_NAN = 8;

class DomainError extends Error {
  constructor(message) {
    super(message);
    this.name = "DomainError";
  }
}

const f64 = new Float64Array(1);
const u32 = new Uint32Array(f64.buffer);

class Schema {
  constructor({ fields, peers, strings, localPeer }) {
    if (fields.length > MAX_FIELDS) throw new DomainError(`more than ${MAX_FIELDS} fields per record`);
    if (peers.length > MAX_PEERS) throw new DomainError(`more than ${MAX_PEERS} peers per clock`);
    this.fields = fields;
    this.peers = peers;
    this.fslot = new Map(fields.map((f, i) => [f, i]));
    this.pslot = new Map(peers.map((p, i) => [p, i]));
    if (!this.pslot.has(localPeer)) throw new DomainError("local peer must have a clock slot");
    this.localPeer = localPeer;
    // id order == UTF-16 code-unit order: exactly what the default sort of strings is
    this.strings = [...new Set(strings)].sort();
    for (const s of this.strings) {
      if (!Number.isNaN(Number(s)) || FORBIDDEN.includes(s)) {
        throw new DomainError(`string ${JSON.stringify(s)} cannot be dictionary-encoded bit-exactly`);
      }
    }
    this.sid = new Map(this.strings.map((s, i) => [s, i]));
    this.paths = [];
    this.pid = new Map();
  }

  /** bb_config ranks: ids at which the strings JS compares records / booleans / NaN as would sort */
  ranks() {
    const rank = (s) => this.strings.filter((x) => x <= s).length;
    return { rankObject: rank("[object Object]"), rankTrue: rank("true"), rankFalse: rank("false"), rankNaN: rank("NaN") };
  }

  pathId(path) {
    let id = this.pid.get(path);
    if (id === undefined) {
      id = this.paths.length;
      this.paths.push(path);
      this.pid.set(path, id);
    }
    return id;
  }

  /**
   * The 64-bit key of String(value) as [lo, hi] (include/bullet_b200.h), or null when no stored value can have that
   * string.  equals / count are type-blind: 25 and "25" name the same bucket (src/bullet-query.js:126-131).
   */
  indexKey(value) {
    if (typeof value === "boolean") return [value ? 1 : 0, KEY_BOOL_HI];
    if (typeof value === "number") return this._numberKey(value);
    if (value === null) value = "null";
    if (typeof value === "object" || value === undefined) return null; // JSON.stringify(object): outside the typed domain
    if (typeof value !== "string") throw new DomainError(`unsupported query value ${String(value)}`);
    if (value === "true" || value === "false") return [value === "true" ? 1 : 0, KEY_BOOL_HI];
    if (value === "NaN") return [0, KEY_NAN_HI];
    const x = Number(value);
    if (!Number.isNaN(x) && String(x) === value) return this._numberKey(x);
    const id = this.sid.get(value);
    return id === undefined ? null : [id, KEY_STR_HI];
  }

  _numberKey(x) {
    if (Number.isNaN(x)) return [0, KEY_NAN_HI];
    if (x === 0) return [0, 0]; // String(-0) is "0"
    f64[0] = x;
    return [u32[0], u32[1]];
  }

  /** One side of range() as a bb_bound { num, rank, flags } (src/bullet-query.js:238-252: JS relational semantics). */
  bound(value, upper) {
    if (typeof value === "string") {
      let flags = BOUND_IS_STRING;
      for (const [name, bit] of [["true", BOUND_TRUE], ["false", BOUND_FALSE], ["NaN", BOUND_NAN]]) {
        if (upper ? name <= value : name >= value) flags |= bit;
      }
      const rank = this.strings.filter((s) => (upper ? s <= value : s < value)).length;
      return { num: Number(value), rank, flags };
    }
    if (typeof value === "boolean") return { num: value ? 1 : 0, rank: 0, flags: 0 };
    if (value === null) return { num: 0, rank: 0, flags: 0 }; // ToNumber(null)
    if (typeof value === "number") return { num: value, rank: 0, flags: 0 };
    return { num: NaN, rank: 0, flags: 0 }; // objects: ToNumber("[object Object]")
  }

  /** -> [tag, lo, hi] */
  encPrim(v) {
    if (typeof v === "boolean") return [TAG_BOOL, v ? 1 : 0, 0];
    if (v === null) return [TAG_NULL, 0, 0];
    if (typeof v === "number") {
      f64[0] = v;
      return [TAG_NUM, u32[0], u32[1]];
    }
    if (typeof v === "string") {
      const id = this.sid.get(v);
      if (id === undefined) throw new DomainError(`string ${JSON.stringify(v)} is not in the dictionary`);
      return [TAG_STR, id, 0];
    }
    throw new DomainError(`unsupported primitive ${String(v)}`);
  }

  decPrim(tag, lo, hi) {
    if (tag === TAG_NUM) {
      u32[0] = lo;
      u32[1] = hi;
      return f64[0];
    }
    if (tag === TAG_STR) return this.strings[lo];
    if (tag === TAG_BOOL) return lo !== 0;
    if (tag === TAG_NULL) return null;
    throw new Error(`bad tag ${tag}`);
  }
}

/**
 * entries: [{ path, data, vectorClock | undefined, local }] in arrival order (js/bullet-b200.js).
 * -> { n, pathId: Uint32Array(2n), head: Uint32Array(4n), clk: Uint32Array(8n), val: Uint32Array(8n) }
 *    = bb_batch { path_id u64[n], head {hdr u64, clk_order u32, user u32}[n], clk u32[n][8], val u64[n][4] }
 */
function packEntries(schema, entries) {
  const n = entries.length;
  const pathId = new Uint32Array(2 * n), head = new Uint32Array(4 * n);
  const clk = new Uint32Array(8 * n), val = new Uint32Array(8 * n);
  for (let i = 0; i < n; ++i) {
    const e = entries[i];
    pathId[2 * i] = schema.pathId(e.path);
    let lo = 0, hi = 0; // header word: kind and tags below, own-key order above
    const v = e.data;
    const isRecord = typeof v === "object" && v !== null;
    if (isRecord) {
      if (Array.isArray(v)) throw new DomainError("arrays are outside the typed domain");
      lo = KIND_OBJ << HDR_KIND_SHIFT;
      let k = 0;
      for (const key of Object.keys(v)) {
        const f = schema.fslot.get(key);
        if (f === undefined) throw new DomainError(`field ${JSON.stringify(key)} not in schema`);
        if (typeof v[key] === "object" && v[key] !== null) throw new DomainError("nested objects are outside the typed domain");
        const [tag, plo, phi] = schema.encPrim(v[key]);
        lo |= tag << (HDR_TAG_SHIFT + 3 * f);
        hi |= f << (4 * k);
        val[8 * i + 2 * f] = plo;
        val[8 * i + 2 * f + 1] = phi;
        ++k;
      }
    } else {
      const [tag, plo, phi] = schema.encPrim(v);
      lo = (KIND_PRIM << HDR_KIND_SHIFT) | (tag << HDR_TAG_SHIFT);
      val[8 * i] = plo;
      val[8 * i + 1] = phi;
    }
    if (!e.local && e.vectorClock && isRecord) {
      // network-with-clock flavour (src/bullet-crt.js:339-353); primitives never carry one (sync:560-563)
      let order = 0, k = 0;
      for (const peer of Object.keys(e.vectorClock)) {
        const s = schema.pslot.get(peer);
        const c = e.vectorClock[peer];
        if (s === undefined) throw new DomainError(`peer ${JSON.stringify(peer)} has no clock slot`);
        if (!Number.isInteger(c) || c < 1 || c > 0xffffffff) throw new DomainError("clock entries must be integers >= 1");
        clk[8 * i + s] = c;
        order |= s << (4 * k);
        ++k;
      }
      lo |= HDR_FLAVOUR_NET;
      head[4 * i + 2] = order;
    }
    head[4 * i] = lo;
    head[4 * i + 1] = hi;
    head[4 * i + 3] = i; // `user`: echoed in the change entry
  }
  return { n, pathId, head, clk, val };
}

function decodeValue(schema, lo, hi, val, at) {
  const kind = (lo >>> HDR_KIND_SHIFT) & 3;
  if (kind === KIND_NONE) return undefined;
  if (kind === KIND_PRIM) return schema.decPrim((lo >>> HDR_TAG_SHIFT) & 7, val[at], val[at + 1]);
  const tags = [];
  let count = 0;
  for (let f = 0; f < MAX_FIELDS; ++f) {
    tags.push((lo >>> (HDR_TAG_SHIFT + 3 * f)) & 7);
    if (tags[f] !== TAG_ABSENT) ++count;
  }
  const out = {};
  for (let k = 0; k < count; ++k) {
    const f = (hi >>> (4 * k)) & 0xf;
    out[schema.fields[f]] = schema.decPrim(tags[f], val[at + 2 * f], val[at + 2 * f + 1]);
  }
  return out;
}

function decodeClock(schema, clk, at, order) {
  let count = 0;
  for (let s = 0; s < MAX_PEERS; ++s) if (clk[at + s] !== 0) ++count;
  const out = {};
  for (let k = 0; k < count; ++k) {
    const s = (order >>> (4 * k)) & 0xf;
    out[schema.peers[s]] = clk[at + s];
  }
  return out;
}

/**
 * out = { verdict: Uint32Array(n), nChanges, idx: Uint32Array, head: Uint32Array(4k), clk: Uint32Array(8k), val: Uint32Array(8k) }
 * -> { codes: [decision code per entry], changes: [{ i, value, vectorClock }] in ARRIVAL order }
 * (verdict[i] = code << 29 | slot of the entry's change, or NO_SLOT; the library stores entries in its own order)
 */
function unpackChanges(schema, n, out) {
  const codes = [], changes = [];
  for (let i = 0; i < n; ++i) {
    const v = out.verdict[i];
    codes.push(v >>> 29);
    const slot = v & NO_SLOT;
    if (slot !== NO_SLOT) {
      const lo = out.head[4 * slot], hi = out.head[4 * slot + 1];
      changes.push({
        i,
        value: decodeValue(schema, lo & ~HDR_FLAVOUR_NET, hi, out.val, 8 * slot),
        vectorClock: decodeClock(schema, out.clk, 8 * slot, out.head[4 * slot + 2]),
      });
    }
  }
  return { codes, changes };
}

module.exports = { Schema, DomainError, packEntries, unpackChanges, decodeValue, decodeClock, NO_SLOT };
