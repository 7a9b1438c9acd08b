"""CPU model of the label kernel's algorithm (spacegraphcats_b200/csrc/sgc_labelseq.cuh), checked against per-k-mer
lookups: the EXACTNESS ARGUMENT of the kernel, independent of CUDA.

The kernel hashes and probes only some k-mers of a read (an anchor per 16, the first k-mer past a break, the far end
of a run whose anchor is missing, and the k-mers that overlap a mismatching base); every other k-mer is *settled*: its
k bases equal the unitig sequence store at the position the anchor's slot implies, and no break bit lies between the
anchor's position and its own.  This file restates

  * the build side: slot = (id | MULTI, store position of the first occurrence, strand bit); break bits set at the first
    k-mer of every sequence, at positions without a k-mer, at every later occurrence of a hash (and the position after
    it), and at the slot's position when a hash becomes a multicount (sgc_table.cuh:table_insert, useq_brk_kernel);
  * the label side: runs (anchor, followers, direction, turned-around flag), the window comparison with first / last
    mismatch, the inheritance limit from the break bits, follow-up runs and ranges (labelseq_kernel's trip loop)

in plain Python and compares, on small cDBGs full of repeats, multicounts, one-k-mer unitigs, reverse-complement reads
and sequencing errors, with what per-k-mer lookups of the reference's table give (cdbg/index_reads.py:143-152): the
same set of unitig ids per read and the same number of matching k-mers.  Test infrastructure only.
"""
import numpy as np
import pytest

GK = 16
MULTI = -2


def _murmur():
    from oracle import oracle as O

    return O


class Model:
    def __init__(self, O, useqs, ids, k):
        self.O, self.k = O, k
        self.store = "".join(useqs)
        n = len(self.store)
        self.brk = np.ones(n + 2, bool)
        self.slot = {}  # hash -> [id or MULTI, position, strand bit]
        self._h = {}
        pos = 0
        for s, uid in zip(useqs, ids):
            nk = len(s) - k + 1
            for r in range(max(nk, 0)):
                if r >= 1:
                    self.brk[pos + r] = False
            pos += len(s)
        pos = 0
        for s, uid in zip(useqs, ids):
            for r in range(max(len(s) - k + 1, 0)):
                hf, hr = self.hashes(s[r : r + k])
                h, p = hf ^ hr, pos + r
                e = self.slot.get(h)
                if e is None:
                    self.slot[h] = [uid, p, hf < hr]
                else:
                    if p != e[1]:
                        self.mark(p)
                    if e[0] != uid:
                        e[0] = MULTI
                        self.mark(e[1])
            pos += len(s)

    def mark(self, p):
        self.brk[p] = self.brk[p + 1] = True

    def hashes(self, kmer):
        v = self._h.get(kmer)
        if v is None:
            b = kmer.encode()
            v = (self.O.murmur3_x64_128(b, 0)[0], self.O.murmur3_x64_128(self.O.revcomp(b), 0)[0])
            self._h[kmer] = v
        return v

    def probe(self, kmer):
        hf, hr = self.hashes(kmer)
        e = self.slot.get(hf ^ hr)
        if e is None or e[0] == MULTI:
            return None, None, None, hf < hr
        return e[0], e[1], e[2], hf < hr

    # ---- what per-k-mer probing gives (the reference's semantics)
    def per_kmer(self, read):
        k = self.k
        ids, hits = set(), 0
        for i in range(len(read) - k + 1):
            v = self.probe(read[i : i + k])[0]
            if v is not None:
                ids.add(v)
                hits += 1
        return ids, hits

    # ---- the kernel's algorithm
    def label(self, read, stats, intervals=None, misses=None):
        """intervals / misses (lists, optional): what the kernel's QUERY mode writes -- per settled run of window indices
        the (first store position, length, id) interval, per k-mer the table does not hold its hash"""
        k, S = self.k, self.store
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        nk = len(read) - k + 1
        ids, hits = set(), 0
        items = [(q, min(GK, nk - q), False, False) for q in range(0, max(nk, 0), GK)]
        probed = set()
        while items:
            qa, n, back, flagged = items.pop(0)
            assert qa not in probed, "a k-mer is the anchor of one item only"
            probed.add(qa)
            stats["probes"] += 1
            v, gp, ori, read_ori = self.probe(read[qa : qa + k])
            if v is not None:
                ids.add(v)
            elif misses is not None:
                hf, hr = self.hashes(read[qa : qa + k])
                misses.append(hf ^ hr)
            if n == 1:
                hits += v is not None
                if v is not None and intervals is not None:
                    intervals.append((gp, 1, v))
                continue
            if v is None:
                if not flagged and n > 2:
                    items.append((qa - (n - 1) if back else qa + (n - 1), n - 1, not back, True))
                else:
                    items += [(qa - d if back else qa + d, 1, False, True) for d in range(1, n)]
                continue
            along = ori == read_ori
            lo = qa - (n - 1) if back else qa
            ao = n - 1 if back else 0
            wn = n - 1 + k

            def store_base(j):
                p = gp - ao + j if along else gp + ao + k - 1 - j
                if p < 0 or p >= len(S):
                    return "A" if along else "T"  # the store is padded with zeros ('A')
                return S[p] if along else comp[S[p]]

            mism = [j for j in range(wn) if read[lo + j] != store_base(j)]
            f, l = (mism[0], mism[-1]) if mism else (wn + GK, -1)
            clean = {i for i in range(n) if i + k - 1 < f or i > l}
            lim = 0
            if along != back:
                while lim < 15 and gp + lim + 1 < len(self.brk) and not self.brk[gp + lim + 1]:
                    lim += 1
            else:
                while lim < 15 and gp - lim >= 0 and not self.brk[gp - lim]:
                    lim += 1
            inh = {i for i in range(n) if (i >= n - 1 - lim if back else i <= lim)}
            settled = (clean & inh) | {ao}
            hits += len(settled)
            if intervals is not None:  # one interval per run of consecutive settled window indices
                idx = sorted(settled)
                start = prev = idx[0]
                for i in idx[1:] + [None]:
                    if i is None or i != prev + 1:
                        first = gp + (start - ao) if along else gp - (prev - ao)
                        intervals.append((first, prev - start + 1, v))
                        start = i
                    prev = i
            stats["settled"] += len(settled) - 1
            singles = sorted(inh - settled)
            assert not singles or singles == list(range(singles[0], singles[-1] + 1)), "one contiguous range"
            items += [(lo + i, 1, False, True) for i in singles]
            if len(inh) < n:
                n2 = n - 1 - lim
                items.append((qa - (lim + 1) if back else qa + lim + 1, n2, back, flagged))
        assert probed <= set(range(nk))
        return ids, hits


def _cdbg(rng, k, glen, alphabet="ACGT", max_piece=40):
    g = "".join(rng.choice(list(alphabet), glen))
    cuts = [0]
    while cuts[-1] < glen - k + 1:
        cuts.append(min(cuts[-1] + int(rng.integers(1, max_piece)), glen - k + 1))
    useqs = [g[a : b + k - 1] for a, b in zip(cuts[:-1], cuts[1:])]
    return g, useqs


@pytest.mark.parametrize("k,glen,alphabet", [(21, 6000, "ACGT"), (11, 5000, "ACGT"), (9, 4000, "AC"), (31, 5000, "ACGT")])
def test_label_model_equals_per_kmer_probing(k, glen, alphabet):
    O = _murmur()
    rng = np.random.default_rng(k * 7 + glen)
    g, useqs = _cdbg(rng, k, glen, alphabet)
    rc = lambda s: O.revcomp(s.encode()).decode()  # noqa: E731
    # unitigs written on either strand, some of them twice (once under another id: multicounts; once under the same
    # id: later occurrences of a hash), in shuffled order
    useqs = [rc(u) if rng.random() < 0.5 else u for u in useqs]
    ids = list(range(len(useqs)))
    useqs += [useqs[3], rc(useqs[10]), useqs[20]]
    ids += [len(ids), len(ids) + 1, 20]
    order = rng.permutation(len(useqs))
    useqs, ids = [useqs[i] for i in order], [ids[i] for i in order]
    m = Model(O, useqs, ids, k)
    assert any(e[0] == MULTI for e in m.slot.values())
    stats = {"probes": 0, "settled": 0}
    total = 0
    for r in range(220):
        n = int(rng.integers(k, 170))
        a = int(rng.integers(0, glen - n))
        read = list(g[a : a + n])
        for p in np.flatnonzero(rng.random(n) < (0.0 if r % 5 == 0 else 0.012)):
            read[p] = alphabet[(alphabet.index(read[p]) + 1) % len(alphabet)]
        read = "".join(read)
        if r & 1:
            read = rc(read)
        if r % 17 == 0:
            read = "".join(rng.choice(list(alphabet), n))  # nothing to do with the cDBG
        want = m.per_kmer(read)
        got = m.label(read, stats)
        assert got == want, (r, read)
        total += n - k + 1
    # the point of the exercise: most k-mers are settled without a probe (less so when k-mers repeat all over the place)
    assert stats["probes"] <= total
    if alphabet == "ACGT" and k >= 21:
        assert stats["probes"] < 0.5 * total and stats["settled"] > 0.4 * total


@pytest.mark.parametrize("k,glen,alphabet", [(21, 6000, "ACGT"), (11, 5000, "ACGT"), (9, 4000, "AC")])
def test_query_model_union_of_intervals_equals_the_set_of_hashes(k, glen, alphabet):
    """sgc_query_count_batch through the sequence store: a query's records are run through the same algorithm, every
    settled run contributes the interval of store positions it lies at, and the number of DISTINCT matching hashes on
    a unitig is the size of the union of the query's intervals with that id; the distinct missing hashes are counted
    separately.  Equals Query's set semantics (search/query_by_sequence.py:170-189) -- also when k-mers repeat inside
    the query, across its records, on both strands, and inside the cDBG (later occurrences of a hash are fenced off by
    break bits, so only the position the table keeps for a hash is ever reported)."""
    O = _murmur()
    rng = np.random.default_rng(k * 11 + glen)
    g, useqs = _cdbg(rng, k, glen, alphabet)
    rc = lambda s: O.revcomp(s.encode()).decode()  # noqa: E731
    useqs = [rc(u) if rng.random() < 0.5 else u for u in useqs]
    ids = list(range(len(useqs)))
    useqs += [useqs[3], rc(useqs[10]), useqs[20]]
    ids += [len(ids), len(ids) + 1, 20]
    order = rng.permutation(len(useqs))
    useqs, ids = [useqs[i] for i in order], [ids[i] for i in order]
    m = Model(O, useqs, ids, k)
    for q in range(12):
        records = []
        for _ in range(int(rng.integers(1, 5))):
            n = int(rng.integers(k, 900))
            a = int(rng.integers(0, glen - n))
            s = list(g[a : a + n])
            for p in np.flatnonzero(rng.random(n) < 0.01):
                s[p] = alphabet[(alphabet.index(s[p]) + 1) % len(alphabet)]
            s = "".join(s)
            records.append(rc(s) if rng.random() < 0.5 else s)
        records.append(records[0])            # the same record twice
        records.append(rc(records[-2]))       # ... and a reverse complement: the same hashes again
        # the reference: one set of hashes per query, counted per unitig
        hashes = set()
        for s in records:
            for i in range(len(s) - k + 1):
                hf, hr = m.hashes(s[i : i + k])
                hashes.add(hf ^ hr)
        want = {}
        for h in hashes:
            e = m.slot.get(h)
            if e is not None and e[0] != MULTI:
                want[e[0]] = want.get(e[0], 0) + 1
        # the kernel's way
        intervals, misses = [], []
        for s in records:
            m.label(s, {"probes": 0, "settled": 0}, intervals, misses)
        covered = {}
        for first, length, v in intervals:
            covered.setdefault(v, set()).update(range(first, first + length))
        got = {v: len(p) for v, p in covered.items()}
        assert got == want, q
        assert sum(got.values()) + len(set(misses)) == len(hashes)
        # positions of different ids never coincide (an interval never spans two unitigs)
        allpos = [p for ps in covered.values() for p in ps]
        assert len(allpos) == len(set(allpos))
