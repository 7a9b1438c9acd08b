"""k-mer hashing facade and the k-mer -> unitig index, mirroring the reference's
``spacegraphcats/cdbg/hashing.py`` (``hash_sequence`` :14-20, ``MPHF_KmerIndex`` :23-130) with the
arithmetic on the GPU.

``GpuKmerTable`` stands where the reference uses ``bbhash_table.BBHashTable`` (:8, :34-37, :69,
:115, :126): an exact map ``u64 hash -> u32 unitig id`` with ``None`` on a miss.  Scalar calls work
(one kernel launch each); the batch methods (``hash_sequences``, ``lookup_many``,
``count_matches_array``, ``label_reads``) are what the pipeline entry points use.
"""
import os
import pickle

import numpy as np

from ..device import DeviceMphf, DeviceTable, get_context, pack_sequences
from .._lib import SGC_MISS, SGC_MAX_K


STORE_MAX_BASES = (1 << 30) - 129  # positions in the unitig sequence store are 30 bits of the slot's aux word


def side_default(n_bases=0):
    """Tables built from unitig sequences keep the unitig sequence store (0.25 B + 1 bit per base) unless
    SGC_SIDE=0 or the sequences are too long for 30-bit positions: then every k-mer probes the table."""
    return os.environ.get("SGC_SIDE", "1") != "0" and n_bases < STORE_MAX_BASES

UNSET_VALUE = 2**32 - 1  # the reference's BBHashTable "unset" value (index_cdbg_by_kmer.py:24)


def _as_bytes(seq):
    return seq.encode("ascii") if isinstance(seq, str) else bytes(seq)


def hash_sequences(seqs, ksize, ctx=None, check_bases=True):
    """Batched ``hash_sequence``: hashes of all k-mers of all sequences.

    Returns ``(hashes uint64[total], kmer_off int64[n+1])``; sequence ``i`` owns
    ``hashes[kmer_off[i]:kmer_off[i+1]]`` (empty when it is shorter than k -- the per-sequence
    ``ValueError`` of the scalar call is the caller's guard, cf. index_reads.py:143)."""
    ctx = ctx or get_context()
    if not 1 <= int(ksize) <= SGC_MAX_K:
        raise ValueError("ksize must be in 1..%d" % SGC_MAX_K)
    buf, off = pack_sequences(seqs)
    hashes, koff, status = ctx.hash_batch(buf, off, ksize)
    if check_bases and status is not None and status.any():
        bad = int(np.flatnonzero(status)[0])
        raise ValueError("Invalid base in read (sequence %d of the batch)" % bad)
    return hashes, koff


def hash_sequence(seq, ksize):
    """Drop-in for cdbg/hashing.py:14-20: list of the L-k+1 canonical k-mer hashes of ``seq``
    (khmer ``Nodetable(k,1,1).get_kmer_hashes``), as Python ints in k-mer order."""
    seq = _as_bytes(seq)
    if len(seq) < ksize:
        raise ValueError(
            "sequence length ({}) must >= the hashtable k-mer size ({})".format(len(seq), ksize)
        )
    hashes, _ = hash_sequences([seq], ksize)
    return hashes.tolist()


class GpuKmerTable:
    """Exact ``hash -> unitig id`` map resident in HBM (BBHashTable stand-in)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or get_context()
        self._dev = None
        self._pending_h, self._pending_v = [], []
        self._host_pairs = None  # cached export
        self._stats = None

    # ---- construction ----------------------------------------------------
    @classmethod
    def from_pairs(cls, hashes, values, ctx=None):
        t = cls(ctx)
        hashes = np.ascontiguousarray(hashes, np.uint64)
        values = np.ascontiguousarray(values, np.uint32)
        t._dev = DeviceTable(t.ctx, max(len(hashes), 1))
        if len(hashes):
            t._dev.insert_pairs(hashes, values)
        return t

    @classmethod
    def from_sequences(cls, buf, off, ksize, ids=None, ctx=None, side=None):
        """build_mphf's two passes fused: hash every unitig k-mer and insert it under the
        unitig's id; hashes seen under two different ids are dropped (multicounts).
        ``side`` (default on; ``SGC_SIDE=0`` or ``side=False`` turns it off) also keeps the unitig sequences
        themselves, 2 bits per base, so that read labelling hashes and probes one k-mer in sixteen and settles
        the others by comparing bases: identical labels, a quarter of the hashes and table lines."""
        t = cls(ctx)
        if isinstance(off, np.ndarray) or not hasattr(off, "device"):
            n_kmers = int(np.maximum(np.diff(off) - ksize + 1, 0).sum())
        else:  # CUDA tensors (already resident sequences)
            n_kmers = int((off[1:] - off[:-1] - (ksize - 1)).clamp_(min=0).sum().item())
        if side is None:
            side = side_default(int(buf.numel()) if hasattr(buf, "numel") else len(buf))
        t._dev = DeviceTable(t.ctx, max(n_kmers, 1))
        if side:
            t._dev.enable_side()
        status = t._dev.insert_sequences(buf, off, ksize, ids=ids)
        t._build_status = t.ctx.to_host(status)
        return t

    def initialize(self, hashes):
        """BBHashTable.initialize: declare the key set; values start unset."""
        hashes = np.fromiter((int(h) for h in hashes), np.uint64) if not isinstance(hashes, np.ndarray) else hashes
        self._dev = DeviceTable(self.ctx, max(len(hashes), 1))
        self._declared = len(hashes)
        self._pending_h, self._pending_v = [], []
        self._invalidate()

    def _invalidate(self):
        self._host_pairs = None
        self._stats = None

    def _flush(self):
        if self._dev is None:
            raise RuntimeError("table is not initialized")
        if self._pending_h:
            h = np.array(self._pending_h, np.uint64)
            v = np.array(self._pending_v, np.uint32)
            self._pending_h, self._pending_v = [], []
            self._dev.insert_pairs(h, v)
            self._invalidate()

    def __setitem__(self, kmer_hash, value):
        value = int(value)
        if not 0 <= value < UNSET_VALUE - 1:
            raise ValueError("unitig id out of range")
        self._pending_h.append(int(kmer_hash))
        self._pending_v.append(value)
        if len(self._pending_h) >= 1 << 20:
            self._flush()

    # ---- scalar API ---------------------------------------------------------
    def __getitem__(self, kmer_hash):
        self._flush()
        v = int(self._dev.lookup(np.array([int(kmer_hash)], np.uint64))[0])
        return None if v == SGC_MISS else v

    def stats(self):
        self._flush()
        if self._stats is None:
            self._stats = self._dev.stats()
        return self._stats  # (live, multicount, max_id)

    def __len__(self):
        return self.stats()[0]

    @property
    def n_ids(self):
        live, _, mx = self.stats()
        return mx + 1 if live else 0

    def _pairs(self):
        self._flush()
        if self._host_pairs is None:
            self._host_pairs = self._dev.export()
        return self._host_pairs

    @property
    def mphf_to_hash(self):
        return self._pairs()[0]

    @property
    def mphf_to_value(self):
        return self._pairs()[1]

    # ---- batch API -------------------------------------------------------------
    def lookup_many(self, hashes):
        """uint32 ids, SGC_MISS (0xFFFFFFFF) where the hash is absent."""
        self._flush()
        return self._dev.lookup(np.ascontiguousarray(hashes, np.uint64))

    def count_matches_array(self, hashes, distinct=False, n_ids=None):
        """-> (dense count_by_id uint32[n_ids], n_miss, n_distinct)"""
        self._flush()
        n_ids = self.n_ids if n_ids is None else n_ids
        return self._dev.count_matches(hashes, max(n_ids, 1), distinct=distinct)

    def get_unique_values(self, hashes, require_exist=False):
        """BBHashTable.get_unique_values: ``{unitig id: number of input hashes that map to it}``;
        an input list counts duplicates, a set does not (hashing.py:67-69)."""
        if not isinstance(hashes, np.ndarray):
            hashes = np.fromiter((int(h) for h in hashes), np.uint64)
        counts, n_miss, _ = self.count_matches_array(hashes)
        if require_exist and n_miss:
            raise ValueError("hashval not found.")
        nz = np.flatnonzero(counts)
        return dict(zip(nz.tolist(), counts[nz].tolist()))

    def count_matches_batch(self, queries, ksize):
        """Batched ``Query`` + ``count_cdbg_matches`` (search/query_by_sequence.py:153-189): ``queries`` is a list
        of queries, each a list of sequences (str / bytes; those shorter than k contribute nothing, :174).
        One hash launch, one sort, one lookup/count pass for ALL queries.
        -> list of ``(cdbg_match_counts dict, number of distinct query k-mers)``, one per query."""
        self._flush()
        seqs, seq_query = [], []
        for qi, q in enumerate(queries):
            for s in q:
                seqs.append(s)
                seq_query.append(qi)
        buf, off = pack_sequences(seqs)
        res = self._dev.count_matches_batch(buf, off, np.array(seq_query, np.int32), len(queries), ksize)
        if res["status"].any():
            bad = int(np.flatnonzero(res["status"])[0])
            raise ValueError("Invalid base in read (sequence %d of query %d)" % (bad, seq_query[bad]))
        keys, counts = res["keys"], res["counts"]
        qs = (keys >> np.uint64(32)).astype(np.int64)
        ids = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
        bounds = np.searchsorted(qs, np.arange(len(queries) + 1))
        out = []
        for qi in range(len(queries)):
            a, b = int(bounds[qi]), int(bounds[qi + 1])
            out.append((dict(zip(ids[a:b].tolist(), counts[a:b].tolist())), int(res["q_distinct"][qi])))
        return out

    def label_reads(self, buf, off, ksize, packed=False):
        self._flush()
        return self._dev.label_reads(buf, off, ksize, packed=packed)

    def label_reads_stream(self, batches, ksize, **kw):
        """Pipelined labelling of a stream of host batches (pinned double buffers, H2D / kernels / D2H of
        consecutive batches overlapped): see :meth:`DeviceTable.label_stream`.  What ``index_reads.main``
        and ``bench.py``'s end-to-end arm run."""
        self._flush()
        return self._dev.label_stream(batches, ksize, **kw)

    def lookup_sequences(self, buf, off, ksize, n_ids=0, want_ids=True):
        self._flush()
        return self._dev.lookup_sequences(buf, off, ksize, n_ids=n_ids, want_ids=want_ids)

    # ---- persistence (the reference's on-disk trio, hashing.py:109-130) -------
    def save(self, mphf_filename, array_filename):
        """BBHashTable.save: a boomphf dump of the live hashes plus the hash / id arrays in MPHF
        order, i.e. files the reference's own ``BBHashTable.load`` accepts.  The MPHF is built on
        the device (sgc_mphf_build); this package's ``load`` only needs the arrays."""
        self._flush()
        d_h, d_v = self._dev.export(to_host=False)
        mphf = DeviceMphf.build(self._dev.ctx, d_h)
        try:
            h, v = mphf.order(d_h, d_v) if d_h.numel() else (np.zeros(0, np.uint64), np.zeros(0, np.uint32))
            raw = mphf.dumps()
        finally:
            mphf.close()
        with open(array_filename, "wb") as fp:
            np.savez_compressed(fp, mphf_to_hash=h, mphf_to_value=v)
        with open(mphf_filename, "wb") as fp:
            fp.write(raw)

    @classmethod
    def load(cls, mphf_filename, array_filename, ctx=None):
        with np.load(array_filename) as z:
            h, v = z["mphf_to_hash"], z["mphf_to_value"]
        keep = v != UNSET_VALUE
        return cls.from_pairs(h[keep], v[keep], ctx=ctx)


def _catlas_csr(catlas):
    """CSR view of a CAtlas for the per-dominator kernel, cached on the object:
    (member_off, member, level_nodes, level_off, n_nodes)."""
    cached = getattr(catlas, "_sgc_csr", None)
    if cached is not None:
        return cached
    nodes = sorted(catlas.levels, key=lambda v: (catlas.levels[v], v))
    n_nodes = (max(nodes) + 1) if nodes else 0
    member_off = np.zeros(n_nodes + 1, np.int64)
    members = [()] * n_nodes
    for v in nodes:
        if catlas.levels[v] == 1:
            m = catlas.layer1_to_cdbg.get(v)
            if m is None:
                raise KeyError("level-1 catlas node %d has no cDBG nodes (first_doms.txt)" % v)
        else:
            m = catlas.children[v]
        members[v] = sorted(m)
        member_off[v + 1] = len(members[v])
    np.cumsum(member_off, out=member_off)
    member = np.fromiter((x for m in members for x in m), np.uint32, int(member_off[-1]))
    max_level = max(catlas.levels.values()) if nodes else 0
    level_nodes = np.array(nodes, np.uint32)
    lv = np.array([catlas.levels[v] for v in nodes], np.int64)
    level_off = np.searchsorted(lv, np.arange(1, max_level + 2)).astype(np.int64)
    csr = (member_off, member, level_nodes, level_off, n_nodes)
    catlas._sgc_csr = csr
    return csr


class MPHF_KmerIndex:
    """Support kmer -> cDBG node id queries (reference: cdbg/hashing.py:23-130)."""

    def __init__(self, ksize, bbhash_table, sizes):
        self.ksize = ksize
        self.table = bbhash_table
        self.sizes = sizes

    def __len__(self):
        return len(self.table)

    def get_cdbg_id(self, kmer_hash):
        return self.table[kmer_hash]

    def get_cdbg_size(self, cdbg_id):
        return self.sizes[cdbg_id]

    # ---- per-dominator sums (K6) ------------------------------------------------
    def _sum_over_catlas(self, dense_by_cdbg, catlas):
        ctx = self.table.ctx
        member_off, member, level_nodes, level_off, n_nodes = _catlas_csr(catlas)
        need = int(member[: int(member_off[-1])].max()) + 1 if len(member) else 0
        if len(dense_by_cdbg) < need:
            dense_by_cdbg = np.concatenate([dense_by_cdbg, np.zeros(need - len(dense_by_cdbg), np.uint32)])
        d_cnt = ctx.to_device(dense_by_cdbg, np.uint32)
        d_out = ctx.count_doms(d_cnt, member_off, member, level_nodes, level_off, n_nodes)
        return ctx.to_host(d_out, np.uint32), level_nodes

    def _dense_sizes(self):
        n = (max(self.sizes) + 1) if self.sizes else 0
        a = np.zeros(n, np.uint32)
        for k, v in self.sizes.items():
            a[k] = v
        return a

    def sum_by_catlas_node(self, counts_by_cdbg, catlas):
        """Raw K6: per-catlas-node sums of a per-unitig count vector (dict or dense array), with
        no assertions -- the per-dominator abundance sums of count_dominator_abundance.py:90-97."""
        if isinstance(counts_by_cdbg, dict):
            n = (max(counts_by_cdbg) + 1) if counts_by_cdbg else 0
            dense = np.zeros(n, np.uint32)
            for k, v in counts_by_cdbg.items():
                dense[k] = v
        else:
            dense = np.ascontiguousarray(counts_by_cdbg, np.uint32)
        out, _ = self._sum_over_catlas(dense, catlas)
        return {v: int(out[v]) for v in catlas}

    def build_catlas_node_sizes(self, catlas):
        """Number of k-mers in the shadow of each catlas node (hashing.py:42-65)."""
        out, nodes = self._sum_over_catlas(self._dense_sizes(), catlas)
        reach = list(catlas)
        return {v: int(out[v]) for v in reach}

    def count_cdbg_matches(self, query_kmers, verbose=True, require_exist=False):
        "Return a dictionary containing cdbg_id -> # of matches in query_kmers (hashing.py:67-69)"
        return self.table.get_unique_values(query_kmers, require_exist=require_exist)

    def count_catlas_matches(self, cdbg_matches, catlas):
        """Matches summed per dominator and propagated up the catlas (hashing.py:71-107);
        nodes without matches are omitted."""
        n = (max(cdbg_matches) + 1) if cdbg_matches else 0
        dense = np.zeros(n, np.uint32)
        for k, v in cdbg_matches.items():
            dense[k] = v
        out, _ = self._sum_over_catlas(dense, catlas)
        sizes = getattr(catlas, "_sgc_node_sizes", None)
        if sizes is None:
            sizes, _ = self._sum_over_catlas(self._dense_sizes(), catlas)
            catlas._sgc_node_sizes = sizes
        dom_matches = {}
        for v in catlas:
            c = int(out[v])
            if c:
                dom_matches[v] = c
                if catlas.levels[v] == 1:
                    assert int(sizes[v]) >= c
        return dom_matches

    # ---- persistence ----------------------------------------------------------------
    def save(self, location):
        "Save kmer index: contigs.mphf, contigs.indices, contigs.sizes (hashing.py:109-117)."
        self.table.save(os.path.join(location, "contigs.mphf"), os.path.join(location, "contigs.indices"))
        with open(os.path.join(location, "contigs.sizes"), "wb") as fp:
            pickle.dump((self.ksize, self.sizes), fp)

    @classmethod
    def from_directory(cls, location):
        """Load a kmer index written by this package or by the reference (hashing.py:119-130).

        ``contigs.indices`` holds the (hash, id) pairs in MPHF order: enough for exact lookups, but without the
        unitig order the fast read-labelling kernel needs.  When the unitig database the index was built from
        lies next to it (``bcalm.unitigs.db``, the layout conf/Snakefile:288-301 produces) the table is rebuilt
        from those sequences instead -- one fused hash+insert pass on the GPU -- and checked against the stored
        pairs count; otherwise (or if they disagree) the pairs are loaded as they are."""
        with open(os.path.join(location, "contigs.sizes"), "rb") as fp:
            ksize, sizes = pickle.load(fp)
        table = None
        db = os.path.join(location, "bcalm.unitigs.db")
        if os.path.exists(db) and os.environ.get("SGC_B200_REBUILD_FROM_UNITIGS", "1") != "0":
            table = cls._table_from_unitig_db(db, ksize, os.path.join(location, "contigs.indices"))
        if table is None:
            table = GpuKmerTable.load(os.path.join(location, "contigs.mphf"), os.path.join(location, "contigs.indices"))
        return cls(ksize, table, sizes)

    @staticmethod
    def _table_from_unitig_db(db_path, ksize, indices_path):
        import sqlite3

        try:
            con = sqlite3.connect("file:%s?mode=ro" % db_path, uri=True)
            rows = con.execute("SELECT id, sequence FROM sequences ORDER BY id").fetchall()
            con.close()
        except sqlite3.Error:
            return None
        if not rows:
            return None
        buf, off = pack_sequences([r[1] for r in rows])
        ids = np.array([int(r[0]) for r in rows], np.uint32)
        table = GpuKmerTable.from_sequences(buf, off, ksize, ids=ids)
        with np.load(indices_path) as z:
            n_stored = int((z["mphf_to_value"] != UNSET_VALUE).sum())
        if len(table) != n_stored or table._build_status.any():
            return None  # not the unitigs this index was built from
        return table
