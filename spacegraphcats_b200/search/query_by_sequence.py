"""Query a catlas with sequences: the hot-path half of the reference's
``spacegraphcats/search/query_by_sequence.py`` -- ``Query.__init__`` :153-185 (hash every record's
k-mers into ONE set) and ``Query.execute`` :187-227 (per-unitig match counts -> per-dominator
counts -> dominators -> neighbourhood).  Of ``QueryOutput`` (:27-150) the arithmetic is here too -- the scaled
MinHash of the query and of the retrieved contigs (sourmash's Murmur3, seed 42: cdbg/minhash.py) and their
containment / similarity (:41-49) --; signature files, CSV writing and the sqlite contig store stay on the CPU
reference.

All records of a query are hashed in one batch; the set() is a device sort + adjacent-unique
fused into the lookup/count kernel (``sgc_count_matches(distinct=1)``).
"""
import numpy as np

from ..cdbg.hashing import get_context, pack_sequences
from ..utils.seqio import read_records


class QueryResult:
    """What ``Query.execute`` hands to the reference's ``QueryOutput``: the matched dominators
    and, derived from them, the neighbourhood ("shadow") of cDBG nodes."""

    def __init__(self, query, catlas, kmer_idx, doms, catlas_name):
        self.query = query
        self.catlas = catlas
        self.kmer_idx = kmer_idx
        self.doms = doms
        self.catlas_name = catlas_name
        self.total_bp = 0
        self.total_seq = 0
        self.query_sig = getattr(query, "sig", None)
        # here is where we go from the dominators to the actual cDBG nodes (:34)
        self.shadow = catlas.shadow(doms)

    # ---- the two sourmash columns of results.csv (:41-49, :52-89) --------------------------------
    def retrieve_contigs(self, sequences):
        """``QueryOutput.retrieve_contigs`` given the neighbourhood's contig sequences (the reference reads them from
        the sqlite contig store by ``self.shadow``): their scaled MinHash -- every k-mer hashed on the GPU -- plus
        the bp / contig totals of the results row."""
        from ..cdbg.minhash import scaled_minhash

        seqs = [s for s in sequences]
        self.total_seq = len(seqs)
        self.total_bp = sum(len(s) for s in seqs)
        scaled = getattr(self.query, "scaled", None) or 1
        long_enough = [s for s in seqs if len(s) >= self.query.ksize]
        self.contigs_hashes = scaled_minhash(long_enough, self.query.ksize, scaled) if long_enough else np.zeros(0, np.uint64)
        return self.contigs_hashes

    def containment(self):
        """``query_sig.minhash.contained_by(contigs_minhash)`` (:44-45): |Q n C| / |Q| over the scaled sketches."""
        q = self.query.minhash_hashes
        return float(len(np.intersect1d(q, self.contigs_hashes, assume_unique=True)) / len(q)) if len(q) else 0.0

    def similarity(self):
        """``query_sig.minhash.similarity(contigs_minhash)`` (:47-48): Jaccard index of the scaled sketches."""
        q, c = self.query.minhash_hashes, self.contigs_hashes
        u = len(np.union1d(q, c))
        return float(len(np.intersect1d(q, c, assume_unique=True)) / u) if u else 0.0

    def response_curve_lines(self):
        """The ``*.response.txt`` curve (search_utils.output_response_curve, search/search_utils.py:268-329):
        level-1 dominators with matches, by decreasing matched k-mers, with cumulative containment and
        overhead.  The per-dominator match counts and sizes are the K6 sums already on hand."""
        q, cat = self.query, self.catalog_sizes()
        curve = []
        for node, level in self.catlas.levels.items():
            if level == 1:
                n_cont = q.catlas_match_counts.get(node, 0)
                if n_cont:
                    curve.append((n_cont, cat[node] - n_cont, node))
        curve.sort(reverse=True)
        total = sum(c[0] for c in curve)
        lines = ["sum_cont relative_cont relative_overhead sum_cont2 sum_oh catlas_id"]
        step = max(int(len(curve) / 200), 1)
        sofar = total_oh = 0
        for pos, (n_cont, n_oh, node) in enumerate(curve):
            sofar += n_cont
            total_oh += n_oh
            if pos % step == 0 or pos + 1 == len(curve):
                lines.append("{} {} {} {} {} {}".format(sofar, sofar / total, total_oh / total, n_cont, n_oh, node))
        return lines

    def catalog_sizes(self):
        sizes = getattr(self.catlas, "index_sizes", None)
        if sizes is None:
            self.catlas.decorate_with_index_sizes(self.kmer_idx)
            sizes = self.catlas.index_sizes
        return sizes

    def summary(self):
        """The index-derived columns of the ``results.csv`` row (query_by_sequence.py:91-124): bp and
        contigs of the neighbourhood, number of query k-mers and the three upper bounds.  The two
        sourmash columns (containment, similarity of the retrieved contigs) stay on the CPU reference."""
        k = self.kmer_idx.ksize
        bp = sum(self.kmer_idx.get_cdbg_size(c) + k - 1 for c in self.shadow)
        best, cdbg_oh, catlas_oh = (self.query.con_sim_upper_bounds(self.catlas, self.kmer_idx)
                                    if self.query.n_kmers else (0.0, 0, 0))
        return {"bp": bp, "contigs": len(self.shadow), "ksize": k, "num_query_kmers": self.query.n_kmers,
                "best_containment": best, "cdbg_min_overhead": cdbg_oh, "catlas_min_overhead": catlas_oh,
                "catlas_name": self.catlas_name}

    # the two id lists QueryOutput.write saves (:126-132)
    def cdbg_ids_lines(self):
        return [str(x) for x in sorted(self.shadow)]

    def frontier_lines(self):
        return [str(x) for x in sorted(self.doms)]


def pseudo_unitigs(sequences, ksize, piece_kmers=100):
    """Chop sequences into consecutive pieces that overlap by k-1 bases, so that every k-mer
    position lies in exactly one piece (stand-in for a BCALM cDBG when bcalm is not
    available; k-mers repeated across pieces become multicounts, which build_mphf drops)."""
    out = []
    for s in sequences:
        for a in range(0, max(len(s) - ksize + 1, 0), piece_kmers):
            out.append(s[a : a + piece_kmers + ksize - 1])
    return out


def query_batch(query_files, ksize, catlas, kmer_idx, scaled=None, catlas_name=None, debug=False, distributed=False):
    """Batched ``Query(...)`` + ``execute`` over many query files (BASELINE.json configs[2]: batched
    count_cdbg_matches over many query sequences).  ALL records of ALL queries go through one hash launch,
    one device sort and one lookup/count pass (``sgc_query_count_batch``): per query the distinct-hash set and
    its per-unitig counts, exactly ``Query.__init__`` :170-176 + ``count_cdbg_matches`` :189.
    ``distributed=True`` (inside an initialised ``torch.distributed`` job): the queries are sharded over the ranks.
    -> list of (Query, QueryResult or None when catlas is None)."""
    queries, per_query = [], []
    for qf in query_files:
        q = Query.__new__(Query)
        q.filename, q.ksize, q.name, q.catlas_name, q.debug, q.sig = qf, int(ksize), None, catlas_name, debug, None
        q.scaled, q._mh = scaled, None
        seqs = []
        for record in read_records(qf):
            if q.name is None:
                q.name = record.name
            if len(record.sequence) >= q.ksize:
                seqs.append(record.sequence)
        q._seqs = seqs
        per_query.append(seqs)
        queries.append(q)
    if distributed:
        # one process per GPU, every rank holds the table: rank r counts its contiguous range of the queries and
        # the sparse results are all-gathered (parallel.count_matches_batch_sharded)
        from ..parallel import count_matches_batch_sharded

        counted = count_matches_batch_sharded(kmer_idx.table.count_matches_batch, per_query, ksize,
                                              device=get_context().device)
    else:
        counted = kmer_idx.table.count_matches_batch(per_query, ksize)
    out = []
    for q, (counts, n_distinct) in zip(queries, counted):
        q._ctx, q._kmers, q._d_hashes = get_context(), None, None
        q._n_distinct = n_distinct
        q.cdbg_match_counts = counts
        q.catlas_match_counts = None
        out.append((q, q.finish(catlas, kmer_idx) if catlas is not None else None))
    return out


class Query:
    def __init__(self, query_file, ksize, scaled=None, catlas_name=None, debug=True, records=None):
        self.filename = query_file
        self.ksize = int(ksize)
        self.name = None
        self.catlas_name = catlas_name
        self.debug = debug
        self.scaled = scaled
        self.sig = None  # the sourmash signature OBJECT stays on the CPU reference; its hashes: minhash_hashes

        seqs = []
        for record in (records if records is not None else read_records(query_file)):
            if self.name is None:
                self.name = record.name
            if len(record.sequence) >= self.ksize:  # :174
                seqs.append(record.sequence)
        ctx = get_context()
        buf, off = pack_sequences(seqs)
        # hashes stay on the device; the distinct set is only materialised on demand
        d_hashes, _, d_status = ctx.hash_batch(buf, off, self.ksize, to_host=False)
        if d_status is not None and d_status.numel() and bool(d_status.any().item()):
            raise ValueError("Invalid base in read (query %s)" % query_file)
        self._d_hashes = d_hashes
        self._ctx = ctx
        self._seqs = seqs
        self._mh = None
        self._kmers = None
        self._n_distinct = None
        self.cdbg_match_counts = None
        self.catlas_match_counts = None

    @property
    def minhash_hashes(self):
        """The hashes of the query's scaled MinHash (``self.mh`` of the reference, :164-176: sourmash
        ``MinHash(0, ksize, scaled=scaled)`` + ``add_sequence`` of every record that is at least k long): sorted
        distinct sourmash hashes below 2^64 / scaled, computed on the GPU (cdbg/minhash.py)."""
        if self._mh is None:
            from ..cdbg.minhash import scaled_minhash

            self._mh = scaled_minhash(self._seqs, self.ksize, self.scaled or 1) if self._seqs else np.zeros(0, np.uint64)
        return self._mh

    # the reference exposes a Python set of ints; build it lazily
    @property
    def kmers(self):
        if self._kmers is None:
            if self._d_hashes is None:  # built by query_batch: hash this query's records on demand
                buf, off = pack_sequences(self._seqs)
                self._d_hashes, _, _ = self._ctx.hash_batch(buf, off, self.ksize, to_host=False)
            h = self._ctx.to_host(self._d_hashes, np.uint64)
            self._kmers = set(np.unique(h).tolist())
            self._n_distinct = len(self._kmers)
        return self._kmers

    def __len__(self):
        return self.n_kmers

    @property
    def n_kmers(self):
        if self._n_distinct is None:
            len(self.kmers)
        return self._n_distinct

    def execute(self, catlas, kmer_idx):
        table = kmer_idx.table
        counts, n_miss, n_distinct = table.count_matches_array(self._d_hashes, distinct=True)
        self._n_distinct = n_distinct
        nz = np.flatnonzero(counts)
        self.cdbg_match_counts = dict(zip(nz.tolist(), counts[nz].tolist()))  # :189
        return self.finish(catlas, kmer_idx)

    def finish(self, catlas, kmer_idx):
        """``execute`` after ``cdbg_match_counts`` is known (:191-227)."""
        for k, v in self.cdbg_match_counts.items():  # :191-192
            assert v <= kmer_idx.get_cdbg_size(k), k

        self.catlas_match_counts = kmer_idx.count_catlas_matches(self.cdbg_match_counts, catlas)  # :195

        if self.debug and self.catlas_match_counts:  # :199-211
            assert sum(self.cdbg_match_counts.values()) == self.catlas_match_counts[catlas.root]
            index_sizes = getattr(catlas, "index_sizes", None)
            if index_sizes is not None:
                for v, match_amount in self.catlas_match_counts.items():
                    assert match_amount <= index_sizes[v], v
                self.con_sim_upper_bounds(catlas, kmer_idx)

        doms = {catlas.cdbg_to_layer1[c] for c in self.cdbg_match_counts}  # :214-220
        return QueryResult(self, catlas, kmer_idx, doms, self.catlas_name)

    def con_sim_upper_bounds(self, catlas, kmer_idx):
        """Best possible containment / overheads (:229-270) -- host-side ratios."""
        root = catlas.root
        total_match = sum(self.cdbg_match_counts.values())
        best_containment = total_match / self.n_kmers if self.n_kmers else 0.0
        in_nodes = sum(kmer_idx.get_cdbg_size(c) for c in self.cdbg_match_counts)
        cdbg_min_overhead = (in_nodes - total_match) / in_nodes if in_nodes else 0
        catlas_min_overhead = 0
        cm = self.catlas_match_counts
        if cm and cm.get(root):
            in_doms = sum(
                catlas.index_sizes[v] for v, lvl in catlas.levels.items() if lvl == 1 and cm.get(v)
            )
            catlas_min_overhead = (in_doms - cm[root]) / in_doms
        return best_containment, cdbg_min_overhead, catlas_min_overhead
