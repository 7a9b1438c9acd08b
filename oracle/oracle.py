"""CPU ORACLE -- test infrastructure, NOT product code.

A restatement of spacegraphcats' k-mer -> unitig indexing path in plain Python /
numpy (+ ``libsgc_oracle.so`` built from ``sgc_oracle.c`` for anything larger
than a few thousand k-mers).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module.  ``spacegraphcats_b200`` never does.

What is restated, with the reference location each function follows
(paths relative to the reference checkout):

* ``murmur3_x64_128`` / ``hash_kmer`` / ``hash_sequence``  -- khmer 3.0.0a3
  ``Nodetable.get_kmer_hashes`` as called from ``spacegraphcats/cdbg/hashing.py:14-20``
  (third-party, absent from the checkout; published algorithm: smhasher
  MurmurHash3_x64_128 seed 0, forward XOR reverse complement, low word).
* ``OracleTable``  -- ``bbhash_table.BBHashTable`` as used at ``hashing.py:34,37,69``
  and ``index_cdbg_by_kmer.py:63-64,89,106`` (exact map, miss = None).
* ``build_mphf``  -- ``spacegraphcats/cdbg/index_cdbg_by_kmer.py:27-115``.
* ``count_cdbg_matches`` / ``count_catlas_matches`` / ``build_catlas_node_sizes``
  -- ``hashing.py:67-69, 71-107, 42-65``.
* ``OracleCAtlas``  -- ``spacegraphcats/search/catlas.py:43-99, 142-154, 174-211``.
* ``label_reads``  -- the loop of ``spacegraphcats/cdbg/index_reads.py:105-167``.
* ``query_kmers`` / ``query_execute``  -- ``search/query_by_sequence.py:153-227``.
* ``dominator_abundance``  -- ``search/count_dominator_abundance.py:80-97``.

PARITY PINS: see tests/test_oracle.py (dory k=21 fixture: 212,475 pairs, 4 KATs,
query response files).  UNPINNED: non-ACGT input bytes (passed through the
reverse complement unchanged here and in the CUDA path; the host policy decides
whether to raise).
"""
import ctypes
import gzip
import os
from collections import defaultdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MISS = 0xFFFFFFFF
_M64 = (1 << 64) - 1


# --------------------------------------------------------------------------
# pure-Python Murmur3 (independent of the C file, so a shared bug cannot hide)
# --------------------------------------------------------------------------
def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & _M64


def _fmix(k):
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & _M64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & _M64
    k ^= k >> 33
    return k


def murmur3_x64_128(data, seed=0):
    """smhasher MurmurHash3_x64_128 -> (out[0], out[1]) as unsigned ints."""
    c1, c2 = 0x87C37B91114253D5, 0x4CF5AD432745937F
    n = len(data)
    h1 = h2 = seed
    nblocks = n // 16
    for b in range(nblocks):
        k1 = int.from_bytes(data[16 * b : 16 * b + 8], "little")
        k2 = int.from_bytes(data[16 * b + 8 : 16 * b + 16], "little")
        k1 = (k1 * c1) & _M64
        k1 = _rotl(k1, 31)
        k1 = (k1 * c2) & _M64
        h1 ^= k1
        h1 = _rotl(h1, 27)
        h1 = (h1 + h2) & _M64
        h1 = (h1 * 5 + 0x52DCE729) & _M64
        k2 = (k2 * c2) & _M64
        k2 = _rotl(k2, 33)
        k2 = (k2 * c1) & _M64
        h2 ^= k2
        h2 = _rotl(h2, 31)
        h2 = (h2 + h1) & _M64
        h2 = (h2 * 5 + 0x38495AB5) & _M64
    tail = data[16 * nblocks :]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * c2) & _M64
        k2 = _rotl(k2, 33)
        k2 = (k2 * c1) & _M64
        h2 ^= k2
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * c1) & _M64
        k1 = _rotl(k1, 31)
        k1 = (k1 * c2) & _M64
        h1 ^= k1
    h1 ^= n
    h2 ^= n
    h1 = (h1 + h2) & _M64
    h2 = (h2 + h1) & _M64
    h1 = _fmix(h1)
    h2 = _fmix(h2)
    h1 = (h1 + h2) & _M64
    h2 = (h2 + h1) & _M64
    return h1, h2


_COMP = bytes.maketrans(b"ACGT", b"TGCA")  # every other byte maps to itself


def revcomp(kmer):
    return kmer.translate(_COMP)[::-1]


def hash_kmer_py(kmer):
    if isinstance(kmer, str):
        kmer = kmer.encode()
    return murmur3_x64_128(kmer)[0] ^ murmur3_x64_128(revcomp(kmer))[0]


def hash_sequence_py(seq, ksize):
    """hashing.py:14-20 in pure Python (small inputs only)."""
    if isinstance(seq, str):
        seq = seq.encode()
    if len(seq) < ksize:
        raise ValueError("sequence length (%d) must >= the hashtable k-mer size (%d)" % (len(seq), ksize))
    return [hash_kmer_py(seq[i : i + ksize]) for i in range(len(seq) - ksize + 1)]


def minhash_kmer_py(kmer, seed=42):
    """sourmash DNA k-mer hash: Murmur3_x64_128(min(kmer, revcomp(kmer)), seed).h1 -- the hash behind
    MinHash.add_sequence (sourmash is third-party and absent; pinned by the reference's own
    ``min_hashval`` column, cdbg/sort_bcalm_unitigs.py:56-128, tests/golden/dory_k21_min_hashval.json)."""
    if isinstance(kmer, str):
        kmer = kmer.encode()
    return murmur3_x64_128(min(kmer, revcomp(kmer)), seed)[0]


def minhash_sequence_py(seq, ksize, seed=42):
    if isinstance(seq, str):
        seq = seq.encode()
    return [minhash_kmer_py(seq[i : i + ksize], seed) for i in range(len(seq) - ksize + 1)]


# --------------------------------------------------------------------------
# C oracle (sgc_oracle.c) through ctypes
# --------------------------------------------------------------------------
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libsgc_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libsgc_oracle.so missing: run `make -C oracle` (or __graft_entry__.build())")
        L = ctypes.CDLL(path)
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.sgc_oracle_hash_sequence.restype = i64
        L.sgc_oracle_hash_sequence.argtypes = [vp, i64, i32, vp]
        L.sgc_oracle_hash_batch.restype = i32
        L.sgc_oracle_hash_batch.argtypes = [vp, vp, i64, i32, vp, vp, i32]
        L.sgc_oracle_count_invalid.restype = i64
        L.sgc_oracle_count_invalid.argtypes = [vp, i64]
        L.sgc_oracle_table_new.restype = vp
        L.sgc_oracle_table_new.argtypes = [vp, vp, i64]
        L.sgc_oracle_table_free.restype = None
        L.sgc_oracle_table_free.argtypes = [vp]
        L.sgc_oracle_table_len.restype = i64
        L.sgc_oracle_table_len.argtypes = [vp]
        L.sgc_oracle_lookup.restype = None
        L.sgc_oracle_lookup.argtypes = [vp, vp, i64, vp]
        L.sgc_oracle_count_matches.restype = i64
        L.sgc_oracle_count_matches.argtypes = [vp, vp, i64, vp]
        L.sgc_oracle_label_reads.restype = i64
        L.sgc_oracle_label_reads.argtypes = [vp, vp, vp, i64, i32, vp, vp, i64, vp, vp, i32]
        L.sgc_oracle_label_reads_checksum.restype = i64
        L.sgc_oracle_label_reads_checksum.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, i32]
        L.sgc_oracle_query_counts.restype = i64
        L.sgc_oracle_query_counts.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, i32]
        L.sgc_oracle_num_threads.restype = i32
        L.sgc_oracle_murmur3_x64_128.restype = None
        L.sgc_oracle_murmur3_x64_128.argtypes = [vp, i32, ctypes.c_uint32, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def murmur3_c(data, seed=0):
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(2, np.uint64)
    lib().sgc_oracle_murmur3_x64_128(_ptr(buf), len(data), seed, _ptr(out))
    return int(out[0]), int(out[1])


def hash_sequence(seq, ksize):
    """hashing.py:14-20 -> numpy uint64[L-k+1] (C oracle)."""
    if isinstance(seq, str):
        seq = seq.encode()
    if len(seq) < ksize:
        raise ValueError("sequence length (%d) must >= the hashtable k-mer size (%d)" % (len(seq), ksize))
    buf = np.frombuffer(seq, dtype=np.uint8)
    out = np.empty(len(seq) - ksize + 1, np.uint64)
    n = lib().sgc_oracle_hash_sequence(_ptr(buf), len(seq), ksize, _ptr(out))
    assert n == len(out)
    return out


def pack_sequences(seqs):
    """list of str/bytes -> (uint8 concatenation, int64 offsets[n+1])."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8) if off[-1] else np.zeros(0, np.uint8)
    return buf, off


def kmer_offsets(off, ksize):
    """Per-sequence output offsets: n_kmers = max(0, L-k+1)."""
    n = np.maximum(np.diff(off) - ksize + 1, 0)
    out = np.zeros(len(off), np.int64)
    out[1:] = np.cumsum(n)
    return out


def hash_batch(buf, off, ksize, threads=0):
    out_off = kmer_offsets(off, ksize)
    out = np.empty(int(out_off[-1]), np.uint64)
    pad = buf if len(buf) else np.zeros(1, np.uint8)
    lib().sgc_oracle_hash_batch(_ptr(pad), _ptr(off), len(off) - 1, ksize, _ptr(out if len(out) else np.zeros(1, np.uint64)), _ptr(out_off), threads)
    return out, out_off


class COracleTable:
    """Flat open-addressing u64->u32 map in C, for sizes where a dict is slow."""

    def __init__(self, hashes, ids):
        self._h = np.ascontiguousarray(hashes, np.uint64)
        self._v = np.ascontiguousarray(ids, np.uint32)
        self._t = lib().sgc_oracle_table_new(_ptr(self._h), _ptr(self._v), len(self._h))

    def __del__(self):
        if getattr(self, "_t", None):
            lib().sgc_oracle_table_free(self._t)
            self._t = None

    def __len__(self):
        return lib().sgc_oracle_table_len(self._t)

    def lookup(self, hashes):
        hashes = np.ascontiguousarray(hashes, np.uint64)
        out = np.empty(len(hashes), np.uint32)
        if len(hashes):
            lib().sgc_oracle_lookup(self._t, _ptr(hashes), len(hashes), _ptr(out))
        return out

    def count_matches(self, hashes, n_ids):
        hashes = np.ascontiguousarray(hashes, np.uint64)
        cnt = np.zeros(n_ids, np.uint32)
        miss = lib().sgc_oracle_count_matches(self._t, _ptr(hashes), len(hashes), _ptr(cnt)) if len(hashes) else 0
        return cnt, miss

    def get_unique_values_list(self, hashes):
        """BBHashTable.get_unique_values at the reference's granularity: any Python iterable of ints in,
        ``{id: count}`` out (bbhash_table iterates the iterable in Cython; cdbg/hashing.py:67-69)."""
        h = np.fromiter(hashes, np.uint64)
        ids = self.lookup(h)
        ids = ids[ids != 0xFFFFFFFF]
        u, c = np.unique(ids, return_counts=True)
        return dict(zip(u.tolist(), c.tolist()))

    def label_reads(self, buf, off, ksize, threads=0):
        nseq = len(off) - 1
        ids_off = np.zeros(nseq + 1, np.int64)
        nk = np.zeros(1, np.int64)
        nh = np.zeros(1, np.int64)
        pad = buf if len(buf) else np.zeros(1, np.uint8)
        tot = lib().sgc_oracle_label_reads(self._t, _ptr(pad), _ptr(off), nseq, ksize, _ptr(ids_off), None, 0, _ptr(nk), _ptr(nh), threads)
        ids = np.empty(max(int(tot), 1), np.uint32)
        tot2 = lib().sgc_oracle_label_reads(self._t, _ptr(pad), _ptr(off), nseq, ksize, _ptr(ids_off), _ptr(ids), len(ids), _ptr(nk), _ptr(nh), threads)
        assert tot2 == tot
        return ids_off, ids[:tot], int(nk[0]), int(nh[0])

    def query_counts(self, buf, off, qseq, ksize, threads=0):
        """Batched Query + count_cdbg_matches in C (queries in parallel): query q owns sequences
        [qseq[q], qseq[q+1]).  -> (list of {id: count} dicts, n_distinct per query, total k-mers)"""
        off = np.ascontiguousarray(off, np.int64)
        qseq = np.ascontiguousarray(qseq, np.int64)
        nq = len(qseq) - 1
        nk = np.maximum(np.diff(off) - int(ksize) + 1, 0)
        per_q = np.add.reduceat(np.concatenate([nk, [0]]), qseq[:-1]) if nq else np.zeros(0, np.int64)
        per_q = np.where(np.diff(qseq) > 0, per_q, 0) if nq else per_q
        kq = np.zeros(nq + 1, np.int64)
        np.cumsum(per_q, out=kq[1:])
        out_ids = np.zeros(max(int(kq[-1]), 1), np.uint32)
        out_cnt = np.zeros(max(int(kq[-1]), 1), np.uint32)
        n_pairs = np.zeros(max(nq, 1), np.int64)
        n_dist = np.zeros(max(nq, 1), np.int64)
        pad = buf if len(buf) else np.zeros(1, np.uint8)
        total = lib().sgc_oracle_query_counts(self._t, _ptr(pad), _ptr(off), _ptr(qseq), nq, int(ksize), _ptr(kq), _ptr(out_ids),
                                              _ptr(out_cnt), _ptr(n_pairs), _ptr(n_dist), threads)
        dicts = []
        for q in range(nq):
            a, n = int(kq[q]), int(n_pairs[q])
            dicts.append(dict(zip(out_ids[a:a + n].tolist(), out_cnt[a:a + n].tolist())))
        return dicts, n_dist[:nq], int(total)

    def label_reads_checksum(self, buf, off, ksize, threads=0):
        cs = np.zeros(1, np.uint64)
        nh = np.zeros(1, np.int64)
        nl = np.zeros(1, np.int64)
        nk = lib().sgc_oracle_label_reads_checksum(self._t, _ptr(buf), _ptr(off), len(off) - 1, ksize, _ptr(cs), _ptr(nh), _ptr(nl), threads)
        return int(nk), int(nh[0]), int(nl[0]), int(cs[0])


def labels_checksum(ids_off, ids):
    """Same order-independent checksum sgc_oracle_label_reads_checksum computes."""
    ids_off = np.asarray(ids_off, np.int64)
    ids = np.asarray(ids, np.uint64)
    read_idx = np.repeat(np.arange(1, len(ids_off), dtype=np.uint64), np.diff(ids_off))
    with np.errstate(over="ignore"):
        v = (read_idx * np.uint64(0x9E3779B97F4A7C15)) ^ (ids * np.uint64(0xC2B2AE3D27D4EB4F))
        return int(v.sum(dtype=np.uint64))


# --------------------------------------------------------------------------
# BBHashTable stand-in and the reference's Python-level semantics
# --------------------------------------------------------------------------
class OracleTable:
    """Exact dict semantics of bbhash_table.BBHashTable (see module docstring)."""

    def __init__(self):
        self._d = {}

    def initialize(self, hashes):
        self._d = {int(h): MISS for h in hashes}

    def __setitem__(self, h, v):
        h = int(h)
        if h not in self._d:
            raise ValueError("hash is not in the table")
        self._d[h] = int(v)

    def __getitem__(self, h):
        v = self._d.get(int(h))
        return None if v is None or v == MISS else v

    def __len__(self):
        return len(self._d)

    @property
    def mphf_to_hash(self):
        return np.fromiter(self._d.keys(), np.uint64, len(self._d))

    @property
    def mphf_to_value(self):
        return np.fromiter(self._d.values(), np.uint32, len(self._d))

    def get_unique_values(self, hashes, require_exist=False):
        counts = defaultdict(int)
        for h in hashes:
            v = self[h]
            if v is None:
                if require_exist:
                    raise ValueError("hashval not found.")
                continue
            counts[v] += 1
        return dict(counts)

    @classmethod
    def from_arrays(cls, mphf_to_hash, mphf_to_value):
        t = cls()
        t._d = {int(h): int(v) for h, v in zip(mphf_to_hash, mphf_to_value)}
        return t


def build_mphf(ksize, records_iter_fn, hash_fn=hash_sequence, table_cls=None):
    """index_cdbg_by_kmer.py:27-115.  Returns (OracleTable, sizes, n_multicounts).  ``table_cls``: any class with
    the BBHashTable surface (the tests drive the library's ctypes binding through this very loop)."""
    seen, multi = set(), set()
    n = -1
    for n, rec in enumerate(records_iter_fn()):
        these = set(int(h) for h in hash_fn(rec.sequence, ksize))  # :38-40
        multi |= seen & these  # :43-45
        seen |= these  # :47
    n_contigs = n + 1
    seen -= multi  # :52-57
    table = (table_cls or OracleTable)()
    table.initialize(list(seen))  # :63-64
    sizes = {}
    max_id = 0
    for n, rec in enumerate(records_iter_fn()):
        cdbg_id = int(rec.name)  # :82
        kmers = [int(h) for h in hash_fn(rec.sequence, ksize)]
        for h in set(kmers) - multi:  # :88-89
            table[h] = cdbg_id
        sizes[cdbg_id] = len(kmers)  # :91
        max_id = max(max_id, cdbg_id)
    vals = table.mphf_to_value
    assert n <= MISS  # :103
    assert int(vals.max()) != MISS  # :106-107
    assert n == max_id  # :110
    assert n == int(vals.max())  # :113
    assert n_contigs == n + 1
    return table, sizes, len(multi)


def count_cdbg_matches(table, query_kmers, require_exist=False):
    """hashing.py:67-69."""
    return table.get_unique_values(query_kmers, require_exist=require_exist)


class Record:
    __slots__ = ("name", "sequence", "quality")

    def __init__(self, name, sequence, quality=None):
        self.name, self.sequence, self.quality = name, sequence, quality


def read_records(path):
    """Minimal FASTA/FASTQ(.gz) reader; names are the first whitespace-delimited token
    stripped of nothing else (screed keeps the full header line as .name)."""
    op = gzip.open if str(path).endswith(".gz") else open
    recs = []
    with op(path, "rt") as fp:
        lines = [ln.rstrip("\r\n") for ln in fp]
    i = 0
    while i < len(lines):
        ln = lines[i]
        if not ln:
            i += 1
        elif ln[0] == ">":
            j = i + 1
            parts = []
            while j < len(lines) and not lines[j].startswith(">"):
                parts.append(lines[j].strip())
                j += 1
            recs.append(Record(ln[1:], "".join(parts)))
            i = j
        elif ln[0] == "@":
            recs.append(Record(ln[1:], lines[i + 1].strip(), lines[i + 3].strip()))
            i += 4
        else:
            raise ValueError("unrecognised sequence file line: %r" % ln[:40])
    return recs


def iterate_records(data):
    """search/search_utils.py:67-154 (iterate_bgzf, my_fasta_iter, my_fastq_iter) over an in-memory
    byte string: yields (Record, position of the record's first line).  ``tell()`` of the reference's
    reader is the byte position here (the BGZF virtual offset is derived from it separately)."""
    data = bytes(data)
    lines, starts, pos = [], [], 0
    while pos < len(data):  # readline()
        e = data.find(b"\n", pos)
        e = len(data) if e < 0 else e + 1
        lines.append(data[pos:e].decode("ascii"))
        starts.append(pos)
        pos = e
    if not data:
        return
    ch = chr(data[0])
    if ch == ">":
        i = 0
        while i < len(lines):
            line = lines[i].strip()
            if not line.startswith(">"):
                raise IOError("Bad FASTA format: no '>' at beginning of line: {}".format(line))
            name = line[1:].strip()
            j = i + 1
            parts = []
            while j < len(lines) and not lines[j].startswith(">"):
                parts.append(lines[j].strip())
                j += 1
            yield Record(name, "".join(parts)), starts[i]
            i = j
    elif ch == "@":
        i = 0
        while i + 3 < len(lines):
            assert lines[i].startswith("@"), lines[i]
            assert lines[i + 2].strip() == "+"
            yield Record(lines[i].strip()[1:], lines[i + 1].strip(), lines[i + 3].strip()), starts[i]
            i += 4
    else:
        raise Exception("unknown start chr {}".format(ch))


def read_unitigs_tsv(path):
    op = gzip.open if str(path).endswith(".gz") else open
    recs = []
    with op(path, "rt") as fp:
        for ln in fp:
            i, s = ln.rstrip("\n").split("\t")
            recs.append(Record(i, s))
    return recs


class OracleCAtlas:
    """search/catlas.py:43-99 (loading), :142-154 (child-before-parent order),
    :174-211 (leaves / shadow)."""

    def __init__(self, catlas_csv, first_doms):
        self.parent, self.children, self.levels = {}, defaultdict(set), {}
        self._cdbg_to_catlas = {}
        self.max_level, self.root = -1, -1
        with open(catlas_csv) as fp:
            for ln in fp:
                node, cdbg, level, kids = ln.strip().split(",")
                node, level = int(node), int(level)
                kids = kids.strip()
                if kids:
                    ks = set(int(x) for x in kids.split(" "))
                    self.children[node] = ks
                    for c in ks:
                        self.parent[c] = node
                self.levels[node] = level
                if level > self.max_level:
                    self.max_level, self.root = level, node
                if level == 1:
                    self._cdbg_to_catlas[int(cdbg)] = node
        self.layer1_to_cdbg, self.cdbg_to_layer1 = {}, {}
        with open(first_doms) as fp:
            for ln in fp:
                dom, *beneath = ln.strip().split(" ")
                node = self._cdbg_to_catlas[int(dom)]
                bs = set(int(x) for x in beneath)
                self.layer1_to_cdbg[node] = bs
                for c in bs:
                    self.cdbg_to_layer1[c] = node

    def __iter__(self):
        q = [self.root]
        pos = 0
        while pos < len(self.levels):
            q.extend(self.children[q[pos]])
            pos += 1
        return reversed(q)

    def leaves(self, nodes=None):
        nodes = [self.root] if nodes is None else nodes
        out, seen = set(), set()
        stack = list(nodes)
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            ks = self.children[v]
            if ks:
                stack.extend(ks)
            else:
                out.add(v)
        return out

    def shadow(self, nodes):
        out = set()
        for leaf in self.leaves(nodes):
            out |= self.layer1_to_cdbg[leaf]
        return out


def build_catlas_node_sizes(sizes, catlas):
    """hashing.py:42-65."""
    out = {}
    for node in catlas:
        if catlas.levels[node] == 1:
            out[node] = sum(sizes[c] for c in catlas.layer1_to_cdbg.get(node))
        else:
            out[node] = sum(out[c] for c in catlas.children[node])
    return out


def count_catlas_matches(cdbg_matches, sizes, catlas):
    """hashing.py:71-107 (zero-count nodes are omitted)."""
    dom = {}
    for node in catlas:
        if catlas.levels[node] == 1:
            q = sum(cdbg_matches.get(c, 0) for c in catlas.layer1_to_cdbg.get(node))
            if q:
                dom[node] = q
                assert sum(sizes[c] for c in catlas.layer1_to_cdbg.get(node)) >= q
        else:
            s = sum(dom.get(c, 0) for c in catlas.children[node])
            if s:
                dom[node] = s
    return dom


def query_kmers(records, ksize, hash_fn=hash_sequence):
    """query_by_sequence.py:170-176: one SET over all records with len >= k."""
    kmers = set()
    for rec in records:
        if len(rec.sequence) >= int(ksize):
            kmers.update(int(h) for h in hash_fn(rec.sequence, ksize))
    return kmers


def query_execute(kmers, table, sizes, catlas):
    """query_by_sequence.py:187-227 -> (cdbg_match_counts, catlas_match_counts, doms, shadow)."""
    cdbg = count_cdbg_matches(table, kmers)
    for c, v in cdbg.items():
        assert v <= sizes[c]
    cat = count_catlas_matches(cdbg, sizes, catlas)
    if cat:
        assert sum(cdbg.values()) == cat[catlas.root]
    doms = set(catlas.cdbg_to_layer1[c] for c in cdbg)
    return cdbg, cat, doms, catlas.shadow(doms)


def label_reads(records_with_offsets, table, ksize, ignore_paired=False, hash_fn=hash_sequence):
    """index_reads.py:105-167.  Yields nothing; returns (rows, n_paired, total_ids) where rows
    is the list of (offset, cdbg_id) tuples the reference INSERTs, in its order up to
    set-iteration order (callers compare as multisets)."""
    rows = []
    total = set()
    last_record = last_offset = None
    n_paired = 0
    cdbg_ids = set()
    for record, offset in records_with_offsets:
        this_name = record.name
        is_paired = False
        if last_record is not None and last_record is not False:
            last_name = last_record.name
            if last_name.endswith("/1") and this_name.endswith("/2"):
                if last_name[:-2] == this_name[:-2] and len(last_name) > 2:
                    is_paired = True
            elif last_name == this_name and last_name:
                is_paired = True
        if is_paired:
            offsets = [last_offset, offset]
            n_paired += 1
        else:
            offsets = [offset]
            cdbg_ids = set()
        if len(record.sequence) < ksize:  # :143-144 (skipped before becoming last_record)
            continue
        cdbg_ids.update(count_cdbg_matches(table, hash_fn(record.sequence, ksize)))
        for o in offsets:
            for c in cdbg_ids:
                rows.append((o, c))
        total.update(cdbg_ids)
        if not ignore_paired:
            last_record, last_offset = record, offset
    return rows, n_paired, total


def dominator_abundance(records, table, ksize, catlas, hash_fn=hash_sequence):
    """count_dominator_abundance.py:80-97: per-k-mer get_cdbg_id, no de-duplication,
    misses under key None; then per-level-1-dominator sums."""
    cdbg_counts = defaultdict(int)
    for rec in records:
        if len(rec.sequence) < ksize:
            continue
        for h in hash_fn(rec.sequence, ksize):
            cdbg_counts[table[h]] += 1
    dom_counts = {}
    for dom, beneath in catlas.layer1_to_cdbg.items():
        dom_counts[dom] = sum(cdbg_counts.get(c, 0) for c in beneath)
    return dict(cdbg_counts), dom_counts
