/**
 * bullet-b200.js - the reference-side shim: plugs libbulletb200 in under an unchanged Bullet instance.
 *
 * Constructed like the reference's own plugins (a constructor that receives the Bullet instance and replaces
 * public properties: src/bullet-query.js:13-21, src/bullet-middleware.js:23-135):
 *
 *     const Bullet = require("bullet-js");
 *     const BulletB200 = require("./js/bullet-b200");
 *     const bullet = new Bullet({ ... });
 *     new BulletB200(bullet, native, { capacity: 1 << 20, collections: { users: ["age", "score", "role", "name"] } });
 *
 * `native` is the N-API addon over include/bullet_b200.h (INTEGRATION.md): a thin pass-through of typed arrays.
 *     native.create({ capacity, nFields, localPeer, flags, rankObject, rankTrue, rankFalse, rankNaN }) -> ctx   bb_create
 *     native.mergeBatch(ctx, n, pathId, head, clk, val, out) -> number of change entries                       bb_merge_batch
 *         in:  Uint32Array views of bb_batch (js/pack.js packs them: interned path ids, heads, clocks, values)
 *         out: { verdict: Uint32Array(n), idx: Uint32Array(n), head: Uint32Array(4n), clk: Uint32Array(8n), val: Uint32Array(8n) }
 *     native.tableRead(ctx, pathId) -> Uint32Array(32)   one 128-byte bb_row                                   bb_table_read
 *     native.indexCreate(ctx, field) / queryEquals(ctx, field, keyLo, keyHi) -> Uint32Array of node ids /
 *     queryCount(...) -> number / queryRange(ctx, field, loNum, loRank, loFlags, hiNum, hiRank, hiFlags) -> Uint32Array
 *                                                        bb_index_create, bb_query_equals / count / range
 * All host-side logic - interning, dictionary, packing, decoding, the reference's result shapes - is JavaScript
 * (this file and js/pack.js).
 *
 * What is replaced, and nothing else:
 *   bullet.crt.handleUpdate(path, data, isFromNetwork)   src/bullet-crt.js:329-385 - one update, decided on the device;
 *       setData (src/bullet.js:139-155), the middleware and query wrappers around it, _applyUpdate, listeners,
 *       log and broadcast all stay the reference's own code and keep bullet.store / bullet.meta as before
 *   bullet.network.sync._processSyncEntries(entries)     src/bullet-network-sync.js:551-569 - the batched ingress:
 *       ONE native call per chunk, then the reference's own effects for every entry, in arrival order
 *   bullet.crt.getVectorClock / vectorClocks              read through to the device (crt clocks live there)
 *   bullet.query.index / equals / range / count           with options.deviceQueries = "<collection>": the device index
 * Values outside the typed domain (nested records, integer-like keys, setCompare; SURVEY.md 8a) make the addon
 * throw: there is no CPU fallback, such collections keep the stock BulletCRT.
 */
const { Schema, packEntries, unpackChanges, decodeClock } = require("./pack");

const BB_CFG_POST_GETDATA = 1;
const BB_CFG_EXACT_ORDER = 512;
const ROW_M_PRESENT = 1, ROW_V_PRESENT = 2;

const REASONS = [
  "no current state",
  "identical clocks and values",
  "identical clocks, decided by value comparison",
  "identical clocks, decided by value comparison",
  "incoming vector clock dominates",
  "current vector clock dominates (incoming is historical)",
  "concurrent modifications, merged objects",
];

class BulletB200 {
  constructor(bullet, native, options = {}) {
    this.bullet = bullet;
    this.native = native;
    this.schema = new Schema({
      fields: options.fields,
      peers: options.peers,
      strings: options.strings || [],
      localPeer: bullet.id,
    });
    this.ctx = native.create({
      capacity: options.capacity,
      nFields: options.fields.length,
      localPeer: this.schema.pslot.get(bullet.id),
      // postGetData: set it when an index hook is installed (query:151,169); exactOrder: device queries return the
      // reference's exact (Map order, Set order) lists (BB_CFG_EXACT_ORDER) instead of the same nodes in device order
      flags: (options.postGetData ? BB_CFG_POST_GETDATA : 0) | (options.exactOrder ? BB_CFG_EXACT_ORDER : 0),
      ...this.schema.ranks(),
    });
    this.calls = 0; // native merge calls (telemetry)
    this._installCrt();
    this._installSync();
    if (options.deviceQueries) this._installQueries(options.deviceQueries);
  }

  /** one bb_merge_batch: pack, call, decode -> { codes, changes: [{ i, value, vectorClock }] } in arrival order */
  merge(entries) {
    const b = packEntries(this.schema, entries);
    const n = b.n;
    const out = {
      verdict: new Uint32Array(n),
      idx: new Uint32Array(n),
      head: new Uint32Array(4 * n),
      clk: new Uint32Array(8 * n),
      val: new Uint32Array(8 * n),
    };
    this.native.mergeBatch(this.ctx, n, b.pathId, b.head, b.clk, b.val, out);
    this.calls++;
    return unpackChanges(this.schema, n, out);
  }

  /** the two clock maps the device holds for a path: { meta, crt } (undefined when absent) */
  clocks(path) {
    const id = this.schema.pid.get(path);
    if (id === undefined) return { meta: undefined, crt: undefined };
    const row = this.native.tableRead(this.ctx, id); // bb_row as 32 words: val 0-7, m_cnt 8-15, v_cnt 16-23, orders 24-25, flags 28
    const flags = row[28];
    return {
      meta: flags & ROW_M_PRESENT ? decodeClock(this.schema, row, 8, row[24]) : undefined,
      crt: flags & ROW_V_PRESENT ? decodeClock(this.schema, row, 16, row[25]) : undefined,
    };
  }

  /** the decision object of src/bullet-crt.js:164-279 for a device decision code */
  _decision(code, value, vectorClock) {
    // flag for flag what src/bullet-crt.js returns: :174-184 (no current state), :208-219 (identical), :222-232 (tie,
    // by the sign of the value comparison), :237-248 (incoming dominates), :252-263 (historical), :268-278
    // (concurrent: neither `incoming` nor `current`); `converge` is true in every branch
    return {
      defer: false,
      historical: code === 5,
      converge: true,
      incoming: code === 0 || code === 2 || code === 4,
      current: code === 3 || code === 5,
      concurrent: code === 6,
      vectorClock,
      reason: REASONS[code],
      value,
    };
  }

  _entry(path, incomingData, isFromNetwork) {
    // flavour: network-with-clock iff an object carrying a truthy __vectorClock arrives from the network
    // (src/bullet-crt.js:339-353); everything else, primitives from the network included, is a local update
    if (
      isFromNetwork &&
      incomingData &&
      typeof incomingData === "object" &&
      incomingData.__vectorClock
    ) {
      const { __vectorClock, ...data } = incomingData;
      return { path, data, vectorClock: __vectorClock, local: false };
    }
    return { path, data: incomingData, vectorClock: undefined, local: true };
  }

  _result(entry, code, change) {
    // a rejected update leaves the stored value in place and only moves the crt clock: read them back like the
    // reference would (inside a batch the crt clock read here is the one after the whole batch; setData does
    // not use the clock of a rejected update)
    const value = change ? change.value : this.bullet._getData(entry.path);
    const vectorClock = change
      ? change.vectorClock
      : this.clocks(entry.path).crt;
    let broadcastData = value;
    if (typeof broadcastData === "object" && broadcastData !== null) {
      broadcastData = { ...broadcastData, __vectorClock: vectorClock }; // src/bullet-crt.js:371-376
    }
    const decision = this._decision(code, value, vectorClock);
    // src/bullet-crt.js:383: decision.incoming || !currentVectorClock || decision.concurrent
    const doUpdate = code === 0 || code === 2 || code === 4 || code === 6;
    return { value, vectorClock, broadcastData, decision, doUpdate };
  }

  _installCrt() {
    const crt = this.bullet.crt;
    const self = this;
    crt.handleUpdate = function (path, incomingData, isFromNetwork = false) {
      self.bullet._getData(path); // src/bullet-crt.js:331: the read that turns a stored falsy value into {}
      const entry = self._entry(path, incomingData, isFromNetwork);
      const out = self.merge([entry]);
      if (entry.local && !out.changes[0] && self.bullet.meta[path]) {
        // a rejected LOCAL put still bumps the clock it was compared with: incrementVectorClock works in place on
        // the object meta[path].vectorClock aliases after an accepted write (src/bullet-crt.js:56-60, 358;
        // src/bullet.js:198-203).  The device keeps that aliasing; mirror what it holds.
        const held = self.clocks(path).meta;
        if (held) self.bullet.meta[path].vectorClock = held;
      }
      return self._result(entry, out.codes[0], out.changes[0]);
    };
    crt.getVectorClock = (key) => self.clocks(key).crt;
  }

  /**
   * bullet.query.index / equals / range / count on the device (src/bullet-query.js:30-45, 186-261, 293-313) for ONE
   * collection (`base`, e.g. "users").  The index hook then runs inside the merge kernel on the raw incoming value
   * (stale entries kept, as in query:139-176); the stock hook finds no JS-side index and does nothing.  Results
   * are BulletNode[] like the reference's; their order is the reference's (Map, Set) order only up to a
   * permutation (multiset) - see include/bullet_b200.h.
   */
  _installQueries(base) {
    const q = this.bullet.query, self = this;
    const slot = (field) => {
      const f = self.schema.fslot.get(field);
      if (f === undefined) throw new Error(`field ${field} not in schema`);
      return f;
    };
    const nodes = (ids) => Array.from(ids, (id) => new self.bullet.BulletNode(self.bullet, self.schema.paths[id]));
    const indexed = new Set();
    const ensure = (field) => {
      if (!indexed.has(field)) {
        self.native.indexCreate(self.ctx, slot(field));
        indexed.add(field);
      }
    };
    const own = (path) => path === base;
    const stock = { index: q.index.bind(q), equals: q.equals.bind(q), range: q.range.bind(q), count: q.count.bind(q) };
    q.index = (path, field = null) => {
      if (!own(path) || !field) return stock.index(path, field);
      ensure(field);
      return q;
    };
    q.equals = (path, field, value) => {
      if (!own(path)) return stock.equals(path, field, value);
      ensure(field); // the reference creates the index on first use (query:194-196)
      const key = self.schema.indexKey(value);
      return key === null ? [] : nodes(self.native.queryEquals(self.ctx, slot(field), key[0], key[1]));
    };
    q.count = (path, field, value) => {
      if (!own(path)) return stock.count(path, field, value);
      ensure(field);
      const key = self.schema.indexKey(value);
      return key === null ? 0 : self.native.queryCount(self.ctx, slot(field), key[0], key[1]);
    };
    q.range = (path, field, min, max) => {
      if (!own(path)) return stock.range(path, field, min, max);
      ensure(field);
      if (min === undefined || max === undefined) return []; // query:246-251: an undefined bound matches nothing
      const lo = self.schema.bound(min, false), hi = self.schema.bound(max, true);
      return nodes(self.native.queryRange(self.ctx, slot(field), lo.num, lo.rank, lo.flags, hi.num, hi.rank, hi.flags));
    };
  }

  _installSync() {
    const sync = this.bullet.network && this.bullet.network.sync;
    if (sync) sync._processSyncEntries = (entries, peerId) => this.processSyncEntries(entries, peerId);
  }

  /**
   * BulletNetworkSync._processSyncEntries (src/bullet-network-sync.js:551-569) with ONE device call per chunk.
   * The reference's per-entry effects run afterwards in arrival order through its own setData: crt.handleUpdate
   * is answered from the batch result instead of calling the device again.
   */
  processSyncEntries(entries) {
    // what setData would hand to handleUpdate for each entry: __fromNetwork already stripped (src/bullet.js:161-178)
    const packed = entries.map((e) =>
      e.deleted
        ? this._entry(e.path, null, false)
        : this._entry(
            e.path,
            typeof e.data === "object" && e.data !== null
              ? { ...e.data, __vectorClock: e.vectorClock }
              : e.data,
            typeof e.data === "object" && e.data !== null
          )
    );
    const out = this.merge(packed);
    const byIndex = new Map(out.changes.map((c) => [c.i, c]));
    const crt = this.bullet.crt;
    const live = crt.handleUpdate;
    let k = 0;
    crt.handleUpdate = (path) => {
      this.bullet._getData(path);
      const i = k++;
      return this._result(packed[i], out.codes[i], byIndex.get(i));
    };
    try {
      for (const entry of entries) {
        const { path, data, vectorClock, deleted } = entry;
        if (deleted) {
          this.bullet.setData(path, null, false);
        } else {
          const networkData =
            typeof data === "object" && data !== null
              ? { ...data, __fromNetwork: true, __vectorClock: vectorClock }
              : data;
          this.bullet.setData(path, networkData, false);
        }
      }
    } finally {
      crt.handleUpdate = live;
    }
  }
}

module.exports = BulletB200;
