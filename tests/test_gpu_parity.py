"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle and the
reference's own golden fixtures.  Integer work: everything is compared bit-exact."""
import hashlib
import json
import os
import sqlite3

import numpy as np

from bgzf_writer import write_bgzf
import pytest

from conftest import golden_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    from oracle import oracle

    return oracle


@pytest.fixture(scope="module")
def ctx():
    import __graft_entry__ as g

    g.build()
    from spacegraphcats_b200.device import get_context

    return get_context(0)


@pytest.fixture(scope="module")
def dory(O):
    recs = O.read_unitigs_tsv(golden_path("dory_k21_unitigs.tsv.gz"))
    z = np.load(golden_path("dory_k21_contigs.indices.npz"))
    with open(golden_path("dory_k21_contigs.sizes.json")) as fp:
        sz = json.load(fp)
    sizes = {int(a): b for a, b in sz["sizes"].items()}
    return recs, z["mphf_to_hash"], z["mphf_to_value"], sizes


@pytest.fixture(scope="module")
def dory_index(ctx, dory):
    from spacegraphcats_b200.cdbg.index_cdbg_by_kmer import build_mphf
    from spacegraphcats_b200.cdbg.hashing import MPHF_KmerIndex

    recs = dory[0]
    table, sizes = build_mphf(21, lambda: iter(recs), ctx=ctx, verbose=False)
    return MPHF_KmerIndex(21, table, sizes)


def rand_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


# ---------------------------------------------------------------- K1: hashing
KATS = {
    "GACCAAAGGCTTACGGTGAGC": 10218271035842461694,  # reference tests/test_dory_workflow.py:196
    "TCTCCGTTTGCAGAAATCGTA": 8436068710919520258,
    "AGCAGGTACAATGGACTGGAT": 13994045974119358468,
    "TTCGGCAAGAAAATTTTAATA": 11971930231572094512,
    "AGACCAACCTTTAGCCTATTCAAGACGAGTG": 505484382184005004,
    "GACCAACCTTTAGCCTATTCAAGACGAGTGT": 17407874517182356438,
    "A" * 31: 8676834403036734983,
    "ACGT" * 8: 0,  # even-k palindrome
}


def test_hash_kats(ctx):
    from spacegraphcats_b200.cdbg.hashing import hash_sequence

    for kmer, want in KATS.items():
        got = hash_sequence(kmer, len(kmer))
        assert got == [want] and isinstance(got[0], int)


@pytest.mark.parametrize("k", [21, 31, 1, 2, 7, 8, 9, 15, 16, 17, 24, 32, 33, 47, 48, 49, 63, 64, 65, 127, 255])
def test_hash_sequence_vs_oracle_many_k(ctx, O, k):
    from spacegraphcats_b200.cdbg.hashing import hash_sequence

    rng = np.random.default_rng(k)
    for n in (k, k + 1, k + 3, k + 4, 150, 511, 1000):
        if n < k:
            continue
        seq = rand_seq(rng, n)
        assert hash_sequence(seq, k) == O.hash_sequence(seq, k).tolist(), (k, n)


def test_hash_sequence_short_raises(ctx):
    from spacegraphcats_b200.cdbg.hashing import hash_sequence

    with pytest.raises(ValueError):
        hash_sequence("ACGT", 5)


def test_hash_invalid_base_policy(ctx, O):
    """Non-ACGT bytes: the kernel flags the sequence (host raises, like khmer's 'Invalid base in
    read'); the hash values themselves follow the oracle's documented pass-through rule."""
    from spacegraphcats_b200.cdbg.hashing import hash_sequence
    from spacegraphcats_b200.device import pack_sequences

    seq = "ACGTNNacgtRYACGTACGTTTGACCAGGTTACGATCGGACTAGCATCAGGACGT"
    with pytest.raises(ValueError):
        hash_sequence(seq, 9)
    buf, off = pack_sequences([seq, "ACGTACGTAGCTAGCTAGCATCGAT"])
    h, koff, status = ctx.hash_batch(buf, off, 9)
    assert status.tolist() == [1, 0]
    want, woff = O.hash_batch(buf, off, 9)
    assert np.array_equal(h, want) and np.array_equal(koff, woff)


@pytest.mark.parametrize("k", [21, 31, 12])
def test_hash_batch_ragged_empty_and_long(ctx, O, k):
    """ragged batch: empty sequences, shorter than k, exactly k, multi-segment (> 256 k-mers),
    a 300 kbp sequence, and an empty batch."""
    from spacegraphcats_b200.device import pack_sequences

    rng = np.random.default_rng(100 + k)
    lens = [0, 1, k - 1, k, k + 1, 150, 0, 255 + k, 256 + k - 1, 256 + k, 257 + k, 5000, 300_000, k, 3]
    lens += rng.integers(0, 700, 400).tolist()
    seqs = [rand_seq(rng, n) for n in lens]
    buf, off = pack_sequences(seqs)
    h, koff, status = ctx.hash_batch(buf, off, k)
    want, woff = O.hash_batch(buf, off, k)
    assert np.array_equal(koff, woff)
    assert np.array_equal(h, want)
    assert not status.any()
    h, koff, status = ctx.hash_batch(np.zeros(0, np.uint8), np.zeros(1, np.int64), k)
    assert len(h) == 0 and koff.tolist() == [0]
    # unaligned device views of the same buffer give the same hashes
    import torch

    d = ctx.to_device(np.concatenate([np.zeros(3, np.uint8), buf]))
    h2, _, _ = ctx.hash_batch(d[3:], off, k)
    assert np.array_equal(h2, want)


def test_dory_unitig_hashes_match_reference_table(ctx, dory):
    """Every dory unitig k-mer hash, under its unitig id, is a pair of the reference's own
    contigs.indices (212,475 pairs)."""
    from spacegraphcats_b200.device import pack_sequences

    recs, ref_h, ref_v, sizes = dory
    buf, off = pack_sequences([r.sequence for r in recs])
    h, koff, _ = ctx.hash_batch(buf, off, 21)
    ids = np.repeat(np.array([int(r.name) for r in recs], np.uint32), np.diff(koff))
    got = np.unique(np.stack([h, ids.astype(np.uint64)], 1), axis=0)
    want = np.unique(np.stack([ref_h, ref_v.astype(np.uint64)], 1), axis=0)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------- K2/K3: table
def test_dory_build_mphf_whole_table(dory, dory_index):
    recs, ref_h, ref_v, ref_sizes = dory
    t = dory_index.table
    assert len(dory_index) == 212475 == len(t)
    assert dory_index.sizes == ref_sizes
    live, multi, max_id = t.stats()
    assert (live, multi, max_id) == (212475, 0, 735)
    order = np.argsort(ref_h)
    assert np.array_equal(t.mphf_to_hash, ref_h[order])
    assert np.array_equal(t.mphf_to_value, ref_v[order])
    # reference tests/test_dory_workflow.py:195-199
    assert t[10218271035842461694] == 118
    assert t[8436068710919520258] == 118
    assert t[13994045974119358468] == 118
    assert t[11971930231572094512] == 187
    assert dory_index.get_cdbg_id(12345) is None
    assert max(t.mphf_to_value) == 735


def test_table_from_reference_arrays_and_lookup(ctx, O, dory):
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable

    _, ref_h, ref_v, _ = dory
    t = GpuKmerTable.from_pairs(ref_h, ref_v, ctx=ctx)
    assert len(t) == len(ref_h)
    rng = np.random.default_rng(1)
    q = np.concatenate([ref_h[rng.integers(0, len(ref_h), 50000)], rng.integers(0, 2**63, 50000).astype(np.uint64),
                        np.array([0, 2**64 - 1], np.uint64)])
    got = t.lookup_many(q)
    want = O.COracleTable(ref_h, ref_v).lookup(q)
    assert np.array_equal(got, want)
    assert (got[:50000] != 0xFFFFFFFF).all()


def test_save_and_from_directory_roundtrip(tmp_path, dory, dory_index):
    from spacegraphcats_b200.cdbg.hashing import MPHF_KmerIndex

    _, ref_h, ref_v, ref_sizes = dory
    dory_index.save(str(tmp_path))
    for fn in ("contigs.mphf", "contigs.indices", "contigs.sizes"):
        assert os.path.exists(tmp_path / fn)
    z = np.load(str(tmp_path / "contigs.indices"))
    assert set(z.files) == {"mphf_to_hash", "mphf_to_value"}
    assert z["mphf_to_hash"].dtype == np.uint64 and z["mphf_to_value"].dtype == np.uint32
    idx = MPHF_KmerIndex.from_directory(str(tmp_path))
    assert idx.ksize == 21 and idx.sizes == ref_sizes and len(idx) == 212475
    assert idx.get_cdbg_id(11971930231572094512) == 187


def test_build_mphf_multicounts_vs_oracle(ctx, O):
    """Hashes shared by two DIFFERENT unitigs are dropped; repeats inside one unitig are kept
    (index_cdbg_by_kmer.py:40-57).  Includes reverse-complement duplicates and k-mers repeated
    within a record."""
    from spacegraphcats_b200.cdbg.index_cdbg_by_kmer import build_mphf

    rng = np.random.default_rng(9)
    k = 31
    base = [rand_seq(rng, int(n)) for n in rng.integers(31, 900, 300)]
    seqs = list(base)
    seqs[10] = seqs[3][5:200] + rand_seq(rng, 40)  # shares k-mers with unitig 3
    seqs[20] = O.revcomp(seqs[7].encode()).decode()  # the RC of unitig 7: every hash collides
    seqs[30] = seqs[30] + seqs[30][:60]  # internal repeat: stays
    seqs[299] = rand_seq(rng, 500)  # the max id must keep >= 1 k-mer
    order = rng.permutation(len(seqs))
    recs = [O.Record(str(i), seqs[i]) for i in order]  # ids are not in file order
    table, sizes = build_mphf(k, lambda: iter(recs), ctx=ctx, verbose=False)
    otable, osizes, n_multi = O.build_mphf(k, lambda: iter(recs))
    assert n_multi > 0
    assert sizes == osizes
    live, multi, max_id = table.stats()
    assert (live, multi) == (len(otable), n_multi)
    oh, ov = otable.mphf_to_hash, otable.mphf_to_value
    o = np.argsort(oh)
    assert np.array_equal(table.mphf_to_hash, oh[o]) and np.array_equal(table.mphf_to_value, ov[o])
    # dropped hashes are misses
    dropped = O.hash_sequence(seqs[7], k)
    assert (table.lookup_many(dropped) == 0xFFFFFFFF).all()


def test_hash_value_zero_is_a_legal_key(ctx, O):
    """Even-k palindromes hash to 0 (h ^ h): 0 must be insertable, found, and never confused with
    an empty slot; a hash that is absent next to it stays a miss."""
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import pack_sequences

    pal = "ACGT" * 8  # k = 32: its own reverse complement
    assert O.hash_kmer_py(pal) == 0
    rng = np.random.default_rng(3)
    useqs = [rand_seq(rng, 200), pal + "T", rand_seq(rng, 90)]  # unitig 1 holds the palindrome + one more 32-mer
    ubuf, uoff = pack_sequences(useqs)
    for side in (False, True):
        t = GpuKmerTable.from_sequences(ubuf, uoff, 32, ctx=ctx, side=side)
        assert t[0] == 1 and t[1] is None
        got = t.lookup_many(np.array([0, 1, 2**64 - 1], np.uint64))
        assert got.tolist() == [1, 0xFFFFFFFF, 0xFFFFFFFF]
        rb, ro = pack_sequences([pal, "G" + pal, rand_seq(rng, 64)])
        lab = t.label_reads(rb, ro, 32)
        assert lab["ids"][lab["ids_off"][0] : lab["ids_off"][1]].tolist() == [1]
        assert 1 in lab["ids"][lab["ids_off"][1] : lab["ids_off"][2]].tolist()
    e = GpuKmerTable.from_pairs(np.array([5], np.uint64), np.array([7], np.uint32), ctx=ctx)
    assert e[0] is None and e[5] == 7  # an EMPTY slot has key 0 but is not the key 0


def test_table_full_is_an_error(ctx):
    from spacegraphcats_b200.device import DeviceTable
    from spacegraphcats_b200._lib import SgcCapacityError

    t = DeviceTable(ctx, 64)
    rng = np.random.default_rng(2)
    with pytest.raises(SgcCapacityError):
        t.insert_pairs(rng.integers(0, 2**63, 1000).astype(np.uint64), np.zeros(1000, np.uint32))


# ---------------------------------------------------------------- K4: read labelling
@pytest.fixture(scope="module")
def dory_reads(O):
    return O.read_records(golden_path("data", "dory-subset.fq.gz"))


def test_label_reads_dory_vs_oracle(ctx, O, dory, dory_index, dory_reads):
    from spacegraphcats_b200.device import pack_sequences

    _, ref_h, ref_v, _ = dory
    buf, off = pack_sequences([r.sequence for r in dory_reads])
    res = dory_index.table.label_reads(buf, off, 21)
    o_off, o_ids, o_nk, o_nh = O.COracleTable(ref_h, ref_v).label_reads(buf, off, 21)
    assert np.array_equal(res["ids_off"], o_off)
    assert np.array_equal(res["ids"], o_ids)
    assert (res["n_kmers"], res["n_hits"]) == (o_nk, o_nh) == (240920, 240920)
    # SURVEY 8c derived pin: 537 reads -> 907 (read, id) pairs
    assert len(dory_reads) == 537 and len(res["ids"]) == 907
    pairs = sorted((i, int(c)) for i in range(537) for c in res["ids"][res["ids_off"][i] : res["ids_off"][i + 1]])
    lines = ["%d,%d\n" % p for p in pairs]
    assert hashlib.sha256("".join(lines).encode()).hexdigest() == \
        "933dbab55d9262d91ce6f3fe4c883957480e4f22e6d9ff503a44ee24d3ea7c90"
    assert set(res["ids"].tolist()) == set(range(736))  # test_dory_index_reads: rc == 0
    # the same from a table loaded from the reference's arrays (no unitig-ordered side array: every
    # k-mer probes the table) -- both label kernels must agree
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable

    plain = GpuKmerTable.from_pairs(ref_h, ref_v, ctx=ctx)
    assert plain._dev.side_nbytes == 0
    res2 = plain.label_reads(buf, off, 21)
    assert np.array_equal(res2["ids_off"], o_off) and np.array_equal(res2["ids"], o_ids)
    # ... and from a table with the unitig-ordered side array (anchor probe + 64-bit
    # hash matching against neighbouring unitig positions): identical labels
    recs = dory[0]
    ubuf, uoff = pack_sequences([r.sequence for r in recs])
    sided = GpuKmerTable.from_sequences(ubuf, uoff, 21, ids=np.array([int(r.name) for r in recs], np.uint32), ctx=ctx, side=True)
    assert sided._dev.side_nbytes > 0
    res3 = sided.label_reads(buf, off, 21)
    assert np.array_equal(res3["ids_off"], o_off) and np.array_equal(res3["ids"], o_ids)
    assert (res3["n_kmers"], res3["n_hits"]) == (o_nk, o_nh)
    # 2-bit packed input (what the streaming feeder uploads): same labels from both kinds of table
    for t in (sided, plain):
        res4 = t.label_reads(buf, off, 21, packed=True)
        assert np.array_equal(res4["ids_off"], o_off) and np.array_equal(res4["ids"], o_ids)
        assert (res4["n_kmers"], res4["n_hits"]) == (o_nk, o_nh) and not res4["status"].any()


def test_label_reads_synthetic_with_errors_vs_oracle(ctx, O):
    """Reads cut from a random genome with substitution errors and reverse complements against a
    pseudo-unitig table: hits, misses, several ids per read, reads shorter than k, empty reads."""
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import pack_sequences

    rng = np.random.default_rng(21)
    k = 31
    genome = rng.integers(0, 4, 400_000).astype(np.uint8)
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[genome]
    # unitigs: consecutive pieces overlapping by k-1 so every k-mer lies in exactly one
    cuts = np.concatenate([[0], np.cumsum(rng.integers(1, 200, 6000))])
    cuts = cuts[cuts < len(g) - k]
    useqs = [g[a : b + k - 1].tobytes() for a, b in zip(cuts[:-1], cuts[1:])]
    ubuf, uoff = pack_sequences(useqs)
    table = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx)  # default: side-array label kernel
    assert table._dev.side_nbytes > 0
    table_plain = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx, side=False)  # every k-mer probes the table
    otab, _, _ = O.build_mphf(k, lambda: iter(O.Record(str(i), s.decode()) for i, s in enumerate(useqs)))
    ot = O.COracleTable(otab.mphf_to_hash, otab.mphf_to_value)
    assert len(table) == len(ot)

    reads = []
    for i in range(20000):
        ln = int(rng.integers(0, 400)) if i % 50 == 0 else 150
        p = int(rng.integers(0, len(g) - 400))
        r = g[p : p + ln].copy()
        err = rng.random(ln) < 0.01
        r[err] = lut[rng.integers(0, 4, int(err.sum()))]
        r = r.tobytes()
        reads.append(O.revcomp(r) if i & 1 else r)
    buf, off = pack_sequences(reads)
    res = table.label_reads(buf, off, k)
    o_off, o_ids, o_nk, o_nh = ot.label_reads(buf, off, k)
    assert (res["n_kmers"], res["n_hits"]) == (o_nk, o_nh)
    assert 0 < o_nh < o_nk
    assert np.array_equal(res["ids_off"], o_off) and np.array_equal(res["ids"], o_ids)
    res_p = table_plain.label_reads(buf, off, k)  # per-k-mer probing kernel (no side array)
    assert np.array_equal(res_p["ids_off"], o_off) and np.array_equal(res_p["ids"], o_ids)
    assert (res_p["n_kmers"], res_p["n_hits"]) == (o_nk, o_nh)
    # 2-bit packed input: host packer == device packer, and the same labels
    from spacegraphcats_b200.device import pack_reads_host

    pk, ln, bad = pack_reads_host(buf, off)
    d_pk, d_ln, d_st = ctx.pack_reads(buf, off)
    assert np.array_equal(ctx.to_host(d_pk), pk) and np.array_equal(ctx.to_host(d_ln), ln) and not bad.any()
    assert not ctx.to_host(d_st).any()
    res_k = table._dev.label_reads_packed(pk, len(reads), k, lengths=ln)
    assert np.array_equal(res_k["ids_off"], o_off) and np.array_equal(res_k["ids"], o_ids)
    assert (res_k["n_kmers"], res_k["n_hits"]) == (o_nk, o_nh)
    # uniform-length batch without a length array
    uni = [r for r in reads if len(r) == 150]
    ub, uo = pack_sequences(uni)
    upk, _, _ = pack_reads_host(ub, uo)
    res_u = table._dev.label_reads_packed(upk, len(uni), k, uniform_len=150)
    u_off, u_ids, u_nk, u_nh = ot.label_reads(ub, uo, k)
    assert np.array_equal(res_u["ids_off"], u_off) and np.array_equal(res_u["ids"], u_ids)
    assert (res_u["n_kmers"], res_u["n_hits"]) == (u_nk, u_nh)
    # the same through per-k-mer ids (count_dominator_abundance shape) and the K4 tail alone
    lk = table.lookup_sequences(buf, off, k, n_ids=len(useqs))
    oh2, ooff2 = O.hash_batch(buf, off, k)
    want_ids = ot.lookup(oh2)
    assert np.array_equal(lk["kmer_off"], ooff2) and np.array_equal(lk["kmer_ids"], want_ids)
    cnt, miss = ot.count_matches(oh2, len(useqs))
    assert np.array_equal(lk["count_by_id"], cnt) and lk["n_kmers"] - lk["n_hits"] == miss


def test_label_stream_matches_one_shot_labelling(ctx, O):
    """The pipelined host-batch API (pinned double buffers, packed or ASCII input) returns, batch by batch,
    exactly what one-shot label_reads returns; an undersized label buffer is re-sized and redone."""
    import torch

    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import pack_reads_host, pack_sequences

    rng = np.random.default_rng(5)
    k = 31
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[rng.integers(0, 4, 200_000)]
    cuts = np.concatenate([[0], np.cumsum(rng.integers(1, 150, 4000))])
    cuts = cuts[cuts < len(g) - k]
    ubuf, uoff = pack_sequences([g[a : b + k - 1].tobytes() for a, b in zip(cuts[:-1], cuts[1:])])
    table = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx)
    batches, want = [], []
    for b in range(5):
        n = int(rng.integers(1, 3000))
        reads = []
        for i in range(n):
            ln = 150 if b % 2 == 0 else int(rng.integers(0, 300))
            p = int(rng.integers(0, len(g) - 300))
            r = g[p : p + ln].copy()
            err = rng.random(ln) < 0.01
            r[err] = lut[rng.integers(0, 4, int(err.sum()))]
            reads.append(r.tobytes())
        buf, off = pack_sequences(reads)
        want.append(table.label_reads(buf, off, k))
        if b == 3:  # ASCII batch, pageable numpy
            batches.append({"seq": buf, "off": off, "tag": b})
        else:
            pk, ln, bad = pack_reads_host(buf, off)
            assert not bad.any()
            hp = torch.from_numpy(pk.copy()).pin_memory() if b != 1 else pk
            if b % 2 == 0:
                batches.append({"packed": hp, "nseq": n, "lengths": None, "uniform_len": 150, "tag": b})
            else:
                batches.append({"packed": hp, "nseq": n, "lengths": ln, "tag": b})
    for kw in ({}, {"ids_per_read": 0.01}, {"prefetch": False}):
        got = []
        for batch, res in table.label_reads_stream(iter(batches), k, **kw):
            got.append((batch["tag"], res["ids_off"].copy(), res["ids"].copy(), res["n_kmers"], res["n_hits"]))
        assert [t for t, *_ in got] == [0, 1, 2, 3, 4]
        for (_, ioff, ids, nk, nh), w in zip(got, want):
            assert np.array_equal(ioff, w["ids_off"]) and np.array_equal(ids, w["ids"])
            assert (nk, nh) == (w["n_kmers"], w["n_hits"])
    assert list(table.label_reads_stream(iter([]), k)) == []


def test_label_reads_many_ids_per_read_overflows_staging(ctx, O):
    """Worst case for the candidate staging buffer: every k-mer of a read maps to a different
    unitig (unitigs of exactly one k-mer), far more candidates than the per-round buffer."""
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import pack_sequences

    rng = np.random.default_rng(33)
    k = 21
    g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 60_000)]
    useqs = [g[i : i + k].tobytes() for i in range(len(g) - k + 1)]
    ubuf, uoff = pack_sequences(useqs)
    otab, _, _ = O.build_mphf(k, lambda: iter(O.Record(str(i), s.decode()) for i, s in enumerate(useqs)))
    ot = O.COracleTable(otab.mphf_to_hash, otab.mphf_to_value)
    reads = [g[i * 1000 : i * 1000 + 5000].tobytes() for i in range(50)] + [g.tobytes()]
    buf, off = pack_sequences(reads)
    o_off, o_ids, _, _ = ot.label_reads(buf, off, k)
    for side in (False, True):
        table = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx, side=side)
        for packed in (False, True):
            res = table.label_reads(buf, off, k, packed=packed)
            assert np.array_equal(res["ids_off"], o_off) and np.array_equal(res["ids"], o_ids), (side, packed)
            assert res["n_labels"] > 100_000


# ---------------------------------------------------------------- K5/K6: queries
@pytest.fixture(scope="module")
def dory_catlas():
    from spacegraphcats_b200.search.catlas import CAtlas

    return CAtlas(golden_path(), golden_path("dory_k21_r1"))


def test_count_cdbg_matches_semantics(O, dory, dory_index):
    """A list counts duplicates, a set does not (BBHashTable.get_unique_values)."""
    _, ref_h, ref_v, _ = dory
    ot = O.OracleTable.from_arrays(ref_h, ref_v)
    recs = O.read_records(golden_path("data", "dory-head.fa"))
    hs = []
    for r in recs:
        hs.extend(O.hash_sequence(r.sequence, 21).tolist())
    hs = hs + hs[:100] + [5, 6, 7]
    assert dory_index.count_cdbg_matches(hs) == O.count_cdbg_matches(ot, hs)
    assert dory_index.count_cdbg_matches(set(hs)) == O.count_cdbg_matches(ot, set(hs)) == {118: 834, 187: 797}
    with pytest.raises(ValueError):
        dory_index.count_cdbg_matches(hs, require_exist=True)
    assert dory_index.count_cdbg_matches([]) == {}


def test_query_dory_head_matches_reference_outputs(O, dory, dory_index, dory_catlas):
    import gzip

    from spacegraphcats_b200.search.query_by_sequence import Query

    _, ref_h, ref_v, ref_sizes = dory
    dory_catlas.decorate_with_index_sizes(dory_index)
    q = Query(golden_path("data", "dory-head.fa"), 21, 1000, "dory_k21_r1")
    out = q.execute(dory_catlas, dory_index)
    assert q.cdbg_match_counts == {118: 834, 187: 797}
    assert q.n_kmers == 1631 == len(q.kmers)
    # oracle of Query.execute
    ocat = O.OracleCAtlas(golden_path("dory_k21_r1", "catlas.csv"), golden_path("dory_k21_r1", "first_doms.txt"))
    okm = O.query_kmers(O.read_records(golden_path("data", "dory-head.fa")), 21)
    ocdbg, ocat_counts, odoms, oshadow = O.query_execute(okm, O.OracleTable.from_arrays(ref_h, ref_v), ref_sizes, ocat)
    assert q.kmers == okm
    assert q.catlas_match_counts == ocat_counts
    assert out.doms == odoms == {97, 150}
    assert out.shadow == oshadow == {118, 187}
    # the reference's own result files
    with gzip.open(golden_path("dory_k21_r1_search", "dory-head.fa.cdbg_ids.txt.gz"), "rt") as fp:
        assert [ln.strip() for ln in fp] == out.cdbg_ids_lines() == ["118", "187"]
    with gzip.open(golden_path("dory_k21_r1_search", "dory-head.fa.frontier.txt.gz"), "rt") as fp:
        assert sorted(ln.strip() for ln in fp) == sorted(out.frontier_lines())
    assert dory_catlas.index_sizes == O.build_catlas_node_sizes(ref_sizes, ocat)
    assert q.con_sim_upper_bounds(dory_catlas, dory_index)[0] == 1.0
    # the reference's response curve and the index-derived columns of its results.csv row
    with open(golden_path("dory_k21_r1_search", "dory-head.fa.response.txt")) as fp:
        assert [ln.rstrip("\n") for ln in fp] == out.response_curve_lines()
    with open(golden_path("dory_k21_r1_search", "results.csv")) as fp:
        row = [ln.strip().split(",") for ln in fp][1]
    sm = out.summary()
    assert [str(sm[c]) for c in ("bp", "contigs", "ksize", "num_query_kmers", "best_containment", "cdbg_min_overhead",
                                 "catlas_min_overhead", "catlas_name")] == row[3:]
    # ... and its two sourmash columns: scaled MinHash of the query vs that of the retrieved contigs (GPU hashing)
    by_id = {int(r.name): r.sequence for r in dory[0]}
    out.retrieve_contigs(by_id[c] for c in sorted(out.shadow))
    assert [str(out.containment()), str(out.similarity()), str(out.total_bp), str(out.total_seq)] == row[1:5]
    want_mh = sorted({h for r in O.read_records(golden_path("data", "dory-head.fa")) for h in O.minhash_sequence_py(r.sequence, 21)
                      if h <= int((2**64 - 1) / 1000)})
    assert q.minhash_hashes.tolist() == want_mh and len(want_mh) > 0


def test_query_nomatch(dory_index, dory_catlas):
    """reference test_dory_search_nomatch: zero matching k-mers must not crash."""
    from spacegraphcats_b200.search.query_by_sequence import Query

    q = Query(golden_path("data", "random-query-nomatch.fa"), 21, 1000, "dory_k21_r1")
    out = q.execute(dory_catlas, dory_index)
    assert q.n_kmers == 8 and q.cdbg_match_counts == {} and q.catlas_match_counts == {}
    assert out.doms == set() and out.shadow == set()
    with open(golden_path("dory_k21_r1_search", "random-query-nomatch.fa.response.txt")) as fp:
        assert [ln.rstrip("\n") for ln in fp] == out.response_curve_lines()
    with open(golden_path("dory_k21_r1_search", "results.csv")) as fp:
        row = [ln.strip().split(",") for ln in fp][2]
    sm = out.summary()
    assert [str(sm[c]) for c in ("bp", "contigs", "ksize", "num_query_kmers", "best_containment", "cdbg_min_overhead",
                                 "catlas_min_overhead", "catlas_name")] == row[3:]


def test_dominator_abundance_totals(O, dory, dory_index, dory_catlas):
    """count_dominator_abundance.py:80-97 on dory-subset: reference totals 240920 / 212475
    (tests/test_dory_workflow.py:859-870)."""
    from spacegraphcats_b200.device import pack_sequences

    recs = O.read_records(golden_path("data", "dory-subset.fa.gz"))
    buf, off = pack_sequences([r.sequence for r in recs])
    res = dory_index.table.lookup_sequences(buf, off, 21, n_ids=736, want_ids=False)
    assert res["n_kmers"] == 240920 == res["n_hits"]
    cdbg_counts = dict(enumerate(res["count_by_id"].tolist()))
    ocat = O.OracleCAtlas(golden_path("dory_k21_r1", "catlas.csv"), golden_path("dory_k21_r1", "first_doms.txt"))
    _, ref_h, ref_v, ref_sizes = dory
    want_cdbg, want_dom = O.dominator_abundance(recs, O.OracleTable.from_arrays(ref_h, ref_v), 21, ocat)
    assert {k: v for k, v in cdbg_counts.items() if v} == want_cdbg
    got_dom = dory_index.sum_by_catlas_node(res["count_by_id"], dory_catlas)
    assert {d: got_dom[d] for d in want_dom} == want_dom
    assert got_dom[dory_catlas.root] == 240920
    assert sum(want_dom.values()) == 240920 and sum(ref_sizes.values()) == 212475
    # ... and, row for row, the two CSV files the reference's OWN count_dominator_abundance.main wrote for this sample
    # (run on the CPU shims by tests/golden/make_ref_outputs.py, committed as dory_subset_abund.json)
    with open(golden_path("dory_subset_abund.json")) as fp:
        ref_out = json.load(fp)
    assert [[i, int(c), ref_sizes[i]] for i, c in enumerate(res["count_by_id"].tolist())] == ref_out["cdbg"]
    dory_catlas.decorate_with_index_sizes(dory_index)
    sizes = dory_catlas.index_sizes
    got_rows = [[n, dory_catlas.levels[n], got_dom[n], sizes[n]] for n in sorted(dory_catlas)]
    # count_dominator_abundance.py:138-157: whole components hang directly under the root; their rows are repeated for
    # the levels they skip
    max_level = max(dory_catlas.levels.values())
    for n in set(dory_catlas.children[dory_catlas.root]):
        got_rows += [[n, level, got_dom[n], sizes[n]] for level in range(dory_catlas.levels[n] + 1, max_level)]
    assert sorted(got_rows) == sorted(ref_out["dom"])


# ---------------------------------------------------------------- index_reads main
def _write_index(tmp_path, dory_index):
    d = tmp_path / "dory_k21"
    d.mkdir()
    dory_index.save(str(d))
    return str(d)


def test_index_reads_main_dory(tmp_path, O, dory, dory_index, dory_reads):
    """reference test_dory_index_reads (rc 0 <=> all 736 ids hit) + the rows against the oracle's
    restatement of the index_reads loop."""
    from spacegraphcats_b200.cdbg import index_reads
    from spacegraphcats_b200.utils import bgzf

    cdbg = _write_index(tmp_path, dory_index)
    fq = "".join("@%s\n%s\n+\n%s\n" % (r.name, r.sequence, r.quality) for r in dory_reads).encode()
    reads = str(tmp_path / "reads.bgz")
    write_bgzf(reads, fq)
    db = str(tmp_path / "reads.index")
    assert index_reads.main([cdbg, reads, db, "-k", "21"]) == 0
    rows = sorted(sqlite3.connect(db).execute("SELECT offset, cdbg_id FROM sequences").fetchall())
    src = bgzf.open_reads(reads)
    _, _, _, pos = bgzf.parse_records(src.data)
    offs = src.virtual_offsets(pos).tolist()
    _, ref_h, ref_v, _ = dory
    want, n_paired, total = O.label_reads(zip(dory_reads, offs), O.OracleTable.from_arrays(ref_h, ref_v), 21)
    assert rows == sorted(want) and len(total) == 736
    assert index_reads.main([cdbg, reads, db, "-k", "21", "-P"]) == -1  # no pairs in dory-subset
    assert index_reads.main([cdbg, reads, db, "-P", "-N"]) == -1


@pytest.mark.parametrize("batch_mb", [None, "0.0005"])
@pytest.mark.parametrize("ignore_paired", [False, True])
def test_index_reads_pairing_semantics(tmp_path, ctx, O, ignore_paired, batch_mb, monkeypatch):
    """Pairing rules of index_reads.py:124-141 (/1,/2 names, equal names, short mates) on the
    reference's own pairing fixture plus synthetic corner cases, against the oracle loop -- as one batch and
    streamed in ~0.5 KB batches (bounded memory: mates and the pairing state must survive every batch cut)."""
    if batch_mb:
        monkeypatch.setenv("SGC_INDEX_READS_BATCH_MB", batch_mb)
    from spacegraphcats_b200.cdbg import index_reads
    from spacegraphcats_b200.cdbg.hashing import MPHF_KmerIndex
    from spacegraphcats_b200.cdbg.index_cdbg_by_kmer import build_mphf
    from spacegraphcats_b200.utils import bgzf

    k = 31
    ref = O.read_records(golden_path("data", "twofoo-short-paired-test.fa.gz"))
    rng = np.random.default_rng(4)
    genome = rand_seq(rng, 6000)
    extra = [
        O.Record("p1/1", genome[0:150]), O.Record("p1/2", genome[300:450]),
        O.Record("p2/1", genome[600:750]), O.Record("p3/2", genome[900:1050]),  # names differ: unpaired
        O.Record("same", genome[1200:1350]), O.Record("same", genome[1500:1650]),  # SRR style
        O.Record("q/1", genome[2000:2150]), O.Record("short", "ACGT"), O.Record("q/2", genome[2300:2450]),
        O.Record("r/1", genome[2600:2750]), O.Record("r/2", "ACGTACGT"),  # short mate
        O.Record("/1", genome[3000:3150]), O.Record("/2", genome[3300:3450]),  # len(name) == 2
        O.Record("nohit", rand_seq(rng, 150)),
    ]
    reads = list(ref) + extra
    # pseudo-unitigs: 100-k-mer pieces of everything
    pieces = []
    for r in list(ref) + [O.Record("g", genome)]:
        s = r.sequence
        for a in range(0, max(len(s) - k + 1, 0), 100):
            pieces.append(s[a : a + 100 + k - 1])
    urecs = [O.Record(str(i), s) for i, s in enumerate(pieces)]
    table, sizes = build_mphf(k, lambda: iter(urecs), ctx=ctx, verbose=False)
    d = tmp_path / "cdbg"
    d.mkdir()
    MPHF_KmerIndex(k, table, sizes).save(str(d))
    fa = "".join(">%s\n%s\n" % (r.name, r.sequence) for r in reads).encode()
    rp = str(tmp_path / "reads.bgz")
    write_bgzf(rp, fa, block_size=700)
    db = str(tmp_path / "out.index")
    rc = index_reads.main([str(d), rp, db, "-k", str(k)] + (["-N"] if ignore_paired else []))
    assert (index_reads.TIMINGS["batches"] == 1) if not batch_mb else (index_reads.TIMINGS["batches"] > 2)
    rows = sorted(sqlite3.connect(db).execute("SELECT offset, cdbg_id FROM sequences").fetchall())
    src = bgzf.open_reads(rp)
    names, _, _, pos = bgzf.parse_records(src.data)
    assert names == [r.name for r in reads]
    offs = src.virtual_offsets(pos).tolist()
    otab, _, _ = O.build_mphf(k, lambda: iter(urecs))
    want, n_paired, total = O.label_reads(zip(reads, offs), otab, k, ignore_paired=ignore_paired)
    assert rows == sorted(want)
    assert (n_paired > 0) != ignore_paired
    assert rc == (0 if max(total) + 1 == len(total) else -1)


def test_index_reads_many_batches_fastq_and_invalid_base_policy(tmp_path, ctx, O, monkeypatch):
    """index_reads.main over a FASTQ file cut into many small batches (paired reads straddling the cuts), and the
    two policies for reads with non-ACGT bases: raise (default, khmer) or keep the read in the stream without ids
    (SGC_B200_INVALID_BASES=skip, the behaviour the comment at index_reads.py:149-150 assumes)."""
    from spacegraphcats_b200.cdbg import index_reads
    from spacegraphcats_b200.cdbg.hashing import MPHF_KmerIndex
    from spacegraphcats_b200.cdbg.index_cdbg_by_kmer import build_mphf
    from spacegraphcats_b200.utils import bgzf

    k = 21
    rng = np.random.default_rng(12)
    genome = rand_seq(rng, 20000)
    pieces = [genome[a : a + 80 + k - 1] for a in range(0, len(genome) - k + 1, 80)]
    urecs = [O.Record(str(i), s) for i, s in enumerate(pieces)]
    table, sizes = build_mphf(k, lambda: iter(urecs), ctx=ctx, verbose=False)
    d = tmp_path / "cdbg"
    d.mkdir()
    MPHF_KmerIndex(k, table, sizes).save(str(d))
    reads = []
    for i in range(600):
        p = int(rng.integers(0, len(genome) - 400))
        ln = int(rng.integers(10, 200))
        reads.append(O.Record("r%d/1" % i, genome[p : p + ln], "I" * ln))
        if i % 3:
            reads.append(O.Record("r%d/2" % i, genome[p + 150 : p + 150 + ln], "I" * ln))
    clean = list(reads)
    reads[10] = O.Record(reads[10].name, reads[10].sequence[:40] + "N" + reads[10].sequence[41:], reads[10].quality)
    reads[11] = O.Record(reads[11].name, "N" * len(reads[11].sequence), reads[11].quality)
    otab, _, _ = O.build_mphf(k, lambda: iter(urecs))
    db = str(tmp_path / "out.index")
    monkeypatch.setenv("SGC_INDEX_READS_BATCH_MB", "0.01")
    for recs, policy in ((clean, None), (reads, "skip"), (reads, None)):
        fq = "".join("@%s\n%s\n+\n%s\n" % (r.name, r.sequence, r.quality) for r in recs).encode()
        rp = str(tmp_path / "reads.bgz")
        write_bgzf(rp, fq, block_size=2000)
        if policy:
            monkeypatch.setenv("SGC_B200_INVALID_BASES", policy)
        else:
            monkeypatch.delenv("SGC_B200_INVALID_BASES", raising=False)
        if recs is reads and policy is None:
            with pytest.raises(ValueError):
                index_reads.main([str(d), rp, db, "-k", str(k)])
            continue
        rc = index_reads.main([str(d), rp, db, "-k", str(k)])
        assert index_reads.TIMINGS["batches"] > 5
        rows = sorted(sqlite3.connect(db).execute("SELECT offset, cdbg_id FROM sequences").fetchall())
        src = bgzf.open_reads(rp)
        _, _, _, pos = bgzf.parse_records(src.data)
        offs = src.virtual_offsets(pos).tolist()
        hash_fn = (lambda s, kk: ([] if "N" in (s if isinstance(s, str) else s.decode()) else O.hash_sequence(s, kk)))
        want, n_paired, total = O.label_reads(zip(recs, offs), otab, k, hash_fn=hash_fn)
        assert rows == sorted(want) and n_paired > 100
        assert rc == (0 if max(total) + 1 == len(total) else -1)


def test_from_directory_rebuilds_the_side_array_from_the_unitig_db(tmp_path, ctx, O, dory, dory_index, dory_reads):
    """An index directory that also holds bcalm.unitigs.db (the Snakefile layout) is loaded WITH the
    unitig-ordered side array, so `python -m ...index_reads cdbg_dir ...` runs the fast label kernel; without the
    database, or with one that does not match the stored pairs, the pairs are loaded as they are."""
    from spacegraphcats_b200.cdbg.hashing import MPHF_KmerIndex
    from spacegraphcats_b200.device import pack_sequences

    recs, ref_h, ref_v, _ = dory
    d = tmp_path / "dory_k21"
    d.mkdir()
    dory_index.save(str(d))
    buf, off = pack_sequences([r.sequence for r in dory_reads])
    o_off, o_ids, _, _ = O.COracleTable(ref_h, ref_v).label_reads(buf, off, 21)

    plain = MPHF_KmerIndex.from_directory(str(d))
    assert plain.table._dev.side_nbytes == 0

    def write_db(records):
        p = str(d / "bcalm.unitigs.db")
        if os.path.exists(p):
            os.unlink(p)
        con = sqlite3.connect(p)
        con.execute("CREATE TABLE sequences (id INTEGER PRIMARY KEY, sequence TEXT, offset INTEGER, min_hashval TEXT, abund FLOAT)")
        con.executemany("INSERT INTO sequences (id, sequence) VALUES (?, ?)", [(int(r.name), r.sequence) for r in records])
        con.commit()
        con.close()

    write_db(recs)
    sided = MPHF_KmerIndex.from_directory(str(d))
    assert sided.table._dev.side_nbytes > 0 and len(sided) == len(plain) == 212475
    for idx in (plain, sided):
        res = idx.table.label_reads(buf, off, 21)
        assert np.array_equal(res["ids_off"], o_off) and np.array_equal(res["ids"], o_ids)
    write_db(recs[:-5])  # not the unitigs of this index: fall back to the stored pairs
    assert MPHF_KmerIndex.from_directory(str(d)).table._dev.side_nbytes == 0


# ---------------------------------------------------------------- K7: partitioned pipeline
def test_partition_kernels_single_gpu(ctx, O):
    """sgc_hash_partition / sgc_partition_pairs / sgc_unpermute_u32 / sgc_labels_from_ids on one
    GPU, emulating R=4 owner ranks (the exchange itself is covered by the gloo test on CPU and
    by the multi-GPU run): per-owner buckets hold exactly the hashes whose top 2 bits name the
    owner, the permutation restores k-mer order, and the K4 tail reproduces label_reads."""
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import DeviceTable, pack_sequences
    from spacegraphcats_b200.parallel import DeviceEngine

    rng = np.random.default_rng(77)
    k, R = 31, 4
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[rng.integers(0, 4, 80_000)]
    cuts = np.concatenate([[0], np.cumsum(rng.integers(1, 150, 2000))])
    cuts = cuts[cuts < len(g) - k]
    useqs = [g[a : b + k - 1].tobytes() for a, b in zip(cuts[:-1], cuts[1:])]
    reads = [g[p : p + 150].tobytes() for p in rng.integers(0, len(g) - 150, 3000)] + [b"ACGT", b""]
    ubuf, uoff = pack_sequences(useqs)
    rbuf, roff = pack_sequences(reads)
    eng = DeviceEngine(ctx)

    # build side: bucket (hash, id) pairs by owner, insert each bucket into that owner's table
    d_hash, d_ids = eng.hash_pairs(ubuf, uoff, k, np.arange(len(useqs), dtype=np.uint32))
    bh, bi, boff = eng.partition_pairs(d_hash, d_ids, R)
    hh = ctx.to_host(bh, np.uint64)
    assert boff[0] == 0 and boff[-1] == len(hh)
    for r in range(R):
        assert ((hh[boff[r] : boff[r + 1]] >> np.uint64(62)) == r).all()
    tables = []
    for r in range(R):
        t = DeviceTable(ctx, max(boff[r + 1] - boff[r], 1), home_shift=2)
        if boff[r + 1] > boff[r]:
            t.insert_pairs(bh[boff[r] : boff[r + 1]], bi[boff[r] : boff[r + 1]])
        tables.append(t)

    # lookup side
    qh, perm, qoff, koff, n = eng.hash_partition(rbuf, roff, k, R)
    want_h, want_koff = O.hash_batch(rbuf, roff, k)
    assert n == len(want_h) and np.array_equal(ctx.to_host(koff), want_koff)
    got_h = np.empty(n, np.uint64)
    got_h[ctx.to_host(perm)] = ctx.to_host(qh, np.uint64)
    assert np.array_equal(got_h, want_h)
    import torch

    replies = torch.cat([tables[r].lookup(qh[qoff[r] : qoff[r + 1]], to_host=False) for r in range(R)])
    kmer_ids = eng.unpermute(replies, perm)
    whole = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx)
    ref = whole.lookup_sequences(rbuf, roff, k)
    assert np.array_equal(ctx.to_host(kmer_ids, np.uint32), ref["kmer_ids"])
    ioff, ids, nh = eng.labels_from_ids(kmer_ids, koff, len(reads))
    lab = whole.label_reads(rbuf, roff, k)
    assert np.array_equal(ctx.to_host(ioff), lab["ids_off"]) and np.array_equal(ctx.to_host(ids, np.uint32), lab["ids"])
    assert nh == lab["n_hits"]


# ---------------------------------------------------------------- full-size properties
def test_full_size_properties(ctx, O):
    """BASELINE.json configs[3] at FULL size (500 M unitig k-mers, 2^24 x 150 bp reads = 2.0e9
    k-mers per batch): too big for the oracle, so size-independent properties are checked --
    strand symmetry of the labels, batch-split invariance, sortedness, count identities -- plus an
    oracle cross-check of a sample of reads through an independent path (CPU hashes -> stand-alone
    lookup kernel)."""
    import torch

    from spacegraphcats_b200.synth import SyntheticCdbg

    n_kmers = int(os.environ.get("SGC_TEST_TABLE_KMERS", 500_000_000))
    n_reads = int(os.environ.get("SGC_TEST_READS", 1 << 24))
    k = 31
    cd = SyntheticCdbg(ctx, n_kmers, k, seed=1)
    SIDE_FIRST = True  # the default label kernel (side array) first, the per-k-mer one as the cross-check
    table = cd.build_table(side=SIDE_FIRST)
    live, multi, max_id = table.stats()
    assert live + multi == n_kmers  # every genome k-mer is distinct at this size w.h.p.; duplicated unitigs drop
    assert multi == cd.dup_kmers and max_id < cd.n_unitigs - cd.n_dup  # duplicated unitigs (highest ids) vanish

    seq_f, off = cd.reads(n_reads, 150, seed=5, err_ppm=5000, rc_half=False)
    seq_r, _ = cd.reads(n_reads, 150, seed=5, err_ppm=5000, rc_half=True)
    a = table.label_reads(seq_f, off, k, to_host=False)
    b = table.label_reads(seq_r, off, k, to_host=False)
    assert a["n_kmers"] == n_reads * 120 == b["n_kmers"]
    assert a["n_hits"] == b["n_hits"] and 0.80 < a["n_hits"] / a["n_kmers"] < 0.90  # (1-0.005)^31 = 0.856
    # strand symmetry: a read and its reverse complement get the same label set
    assert torch.equal(a["ids_off"], b["ids_off"]) and torch.equal(a["ids"], b["ids"])
    assert not torch.equal(seq_f, seq_r)
    # ids ascending inside every read, offsets monotone, no MISS stored
    ioff, ids = a["ids_off"], a["ids"].to(torch.int64) & 0xFFFFFFFF
    assert bool((ioff[1:] >= ioff[:-1]).all()) and int(ioff[-1]) == ids.numel() == a["n_labels"]
    inc = ids[1:] > ids[:-1]
    starts = torch.zeros(ids.numel(), dtype=torch.bool, device=ids.device)
    starts[ioff[:-1][ioff[:-1] < ids.numel()]] = True
    assert bool((inc | starts[1:]).all())
    assert int(ids.max()) <= max_id
    # batch-split invariance: two halves labelled separately == the whole batch
    h = n_reads // 2
    a1 = table.label_reads(seq_f[: h * 150], off[: h + 1], k, to_host=False)
    a2 = table.label_reads(seq_f[h * 150 :], (off[h:] - off[h]).contiguous(), k, to_host=False)
    assert torch.equal(torch.cat([a1["ids"], a2["ids"]]), a["ids"])
    assert torch.equal(torch.cat([a1["ids_off"], a2["ids_off"][1:] + a1["ids_off"][-1]]), a["ids_off"])
    # per-k-mer path: histogram total == hits == the label path's hits
    lk = table.lookup_sequences(seq_f, off, k, n_ids=max_id + 1, want_ids=False, to_host=False)
    assert lk["n_hits"] == a["n_hits"] and int(lk["count_by_id"].to(torch.int64).sum()) == a["n_hits"]
    # the other label kernel (side array on <-> off) gives the same labels at full size
    del a1, a2, lk
    table_s = cd.build_table(side=not SIDE_FIRST)
    c = table_s.label_reads(seq_f, off, k, to_host=False)
    assert torch.equal(c["ids_off"], a["ids_off"]) and torch.equal(c["ids"], a["ids"]) and c["n_hits"] == a["n_hits"]
    del table_s, c
    # oracle cross-check on a sample of reads through an independent path
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(n_reads, 3000, replace=False))
    hs = ctx.to_host(seq_f.view(n_reads, 150)[torch.from_numpy(pick).to(seq_f.device)])
    sbuf = np.ascontiguousarray(hs).reshape(-1)
    soff = np.arange(len(pick) + 1, dtype=np.int64) * 150
    want_h, koff = O.hash_batch(sbuf, soff, k)  # CPU oracle hashes
    got_ids = table.lookup(want_h)  # stand-alone lookup kernel
    ioff_h, ids_h = ctx.to_host(a["ids_off"]), ctx.to_host(a["ids"], np.uint32)
    for j, r in enumerate(pick.tolist()):
        v = got_ids[koff[j] : koff[j + 1]]
        assert np.array_equal(np.unique(v[v != 0xFFFFFFFF]), ids_h[ioff_h[r] : ioff_h[r + 1]]), r


def test_fused_partition_and_gather_single_gpu(ctx, O):
    """sgc_partition_hashes + sgc_labels_from_replies (the fused K7 path) with R=4 owners emulated on
    one GPU: every owner region holds exactly its hashes, ridx points at each k-mer's own hash, and
    labels built from the owners' replies equal label_reads on the whole table."""
    import torch

    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import DeviceTable, pack_sequences
    from spacegraphcats_b200.parallel import DeviceEngine

    rng = np.random.default_rng(78)
    k, R = 31, 4
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[rng.integers(0, 4, 90_000)]
    cuts = np.concatenate([[0], np.cumsum(rng.integers(1, 150, 2000))])
    cuts = cuts[cuts < len(g) - k]
    useqs = [g[a : b + k - 1].tobytes() for a, b in zip(cuts[:-1], cuts[1:])]
    reads = [g[p : p + int(n)].tobytes() for p, n in zip(rng.integers(0, len(g) - 700, 4000), rng.integers(0, 700, 4000))]
    ubuf, uoff = pack_sequences(useqs)
    rbuf, roff = pack_sequences(reads)
    eng = DeviceEngine(ctx)
    whole = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx)

    d_send, cap, ridx, counts, d_off, n_bytes = eng.partition_hashes(rbuf, roff, k, R)
    want_h, want_koff = O.hash_batch(rbuf, roff, k)
    assert sum(counts) == len(want_h)
    send = ctx.to_host(d_send, np.uint64).reshape(R, cap)
    for r in range(R):
        assert ((send[r, : counts[r]] >> np.uint64(62)) == r).all()
        assert counts[r] == int(((want_h >> np.uint64(62)) == r).sum())
    ri = ctx.to_host(ridx, np.uint32)[: len(want_h)]
    assert np.array_equal(send[ri >> 30, ri & ((1 << 30) - 1)], want_h)  # every k-mer finds its own hash
    # owners answer (here: the whole table answers every region)
    back = ctx.empty(R * cap, torch.int32)
    for r in range(R):
        back[r * cap : r * cap + counts[r]] = whole._dev.lookup(d_send[r * cap : r * cap + counts[r]], to_host=False)
    ioff, ids, nh, kids = eng.labels_from_replies(d_off, n_bytes, k, R, ridx, back, cap, want_kmer_ids=True)
    lab = whole.label_reads(rbuf, roff, k)
    assert np.array_equal(ctx.to_host(ioff), lab["ids_off"]) and np.array_equal(ctx.to_host(ids, np.uint32), lab["ids"])
    assert nh == lab["n_hits"]
    assert np.array_equal(ctx.to_host(kids, np.uint32), whole.lookup_sequences(rbuf, roff, k)["kmer_ids"])


def test_peer_probed_partitioned_table_single_gpu(ctx, O):
    """The peer-probe path (sgc_hash_positions / sgc_table_insert_triples / sgc_store_build / sgc_table_set_peers)
    with R=4 owners emulated on one GPU: four unitig shards with ascending id ranges, four owner tables, the
    store replicated; labels through every owner's handle == the oracle's.  Unitigs repeat across shards
    (multicounts: marks must reach the replicated break bits) and inside one shard (same id)."""
    import ctypes

    import torch

    from spacegraphcats_b200._lib import BRK_PAD_BITS, STORE_PAD_WORDS, check
    from spacegraphcats_b200.device import DeviceTable, pack_sequences

    rng = np.random.default_rng(81)
    k, R = 31, 4
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[rng.integers(0, 4, 120_000)]
    cuts = np.concatenate([[0], np.cumsum(rng.integers(1, 150, 3000))])
    cuts = cuts[cuts < len(g) - k]
    useqs = [g[a : b + k - 1].tobytes() for a, b in zip(cuts[:-1], cuts[1:])]
    n = len(useqs)
    useqs += [useqs[5], useqs[n // 2], useqs[n - 3]]  # the same unitigs under other ids: multicounts across shards
    ids = np.arange(len(useqs), dtype=np.uint32)
    reads = [g[p : p + int(m)].tobytes() for p, m in zip(rng.integers(0, len(g) - 700, 5000), rng.integers(0, 700, 5000))]
    reads = [bytes(r[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))) if i % 2 else r for i, r in enumerate(reads)]
    rbuf, roff = pack_sequences(reads)
    L = ctx._L
    # the shards (ascending id ranges), their triples bucketed by owner
    sh_lo = [len(useqs) * r // R for r in range(R + 1)]
    bounds = [0] + sh_lo[1:R] + [0xFFFFFFFF]
    per_owner = [[] for _ in range(R)]
    shards = []
    for r in range(R):
        sbuf, soff = pack_sequences(useqs[sh_lo[r] : sh_lo[r + 1]])
        d_seq, d_off = ctx.to_device(sbuf, np.uint8), ctx.to_device(soff, np.int64)
        h, aux, koff, stat = ctx.hash_positions(d_seq, d_off, k)
        assert not bool(stat.any())
        with torch.cuda.stream(ctx.stream):
            d_ids = ctx.to_device(ids[sh_lo[r] : sh_lo[r + 1]], np.uint32)
            per = torch.repeat_interleave(d_ids, koff[1:] - koff[:-1])
            owner = (h >> 62) & 3
            for o in range(R):
                sel = owner == o
                per_owner[o].append((h[sel], per[sel], aux[sel]))
        shards.append((d_seq, d_off, len(sbuf)))
    max_bytes = max(s[2] for s in shards)
    ustride = (max_bytes + 15) // 16 + 2 * STORE_PAD_WORDS
    bstride = (max_bytes + BRK_PAD_BITS + 31) // 32 + 2
    with torch.cuda.stream(ctx.stream):
        useq_all = torch.zeros(R * ustride, dtype=torch.int32, device=ctx.device)
        brk_all = torch.full((R * bstride,), -1, dtype=torch.int32, device=ctx.device)
    for r, (d_seq, d_off, nb) in enumerate(shards):
        check(L.sgc_store_build(ctx._h, ctx.ptr(d_seq), nb, ctx.ptr(d_off), d_off.numel() - 1, k,
                                ctypes.c_void_p(useq_all.data_ptr() + 4 * (r * ustride + STORE_PAD_WORDS)),
                                ctypes.c_void_p(brk_all.data_ptr() + 4 * r * bstride)))
    n_max = max(sum(int(t[0].numel()) for t in per_owner[o]) for o in range(R))
    tables, marks = [], []
    for o in range(R):
        t = DeviceTable(ctx, n_max, home_shift=2, shared=(o % 2 == 0))  # both kinds of bucket memory
        with torch.cuda.stream(ctx.stream):
            hh, vv, aa = (torch.cat([x[j] for x in per_owner[o]]) for j in range(3))
        marks.append(t.insert_triples(hh, vv, aa, bounds))
        tables.append(t)
    with torch.cuda.stream(ctx.stream):
        every = torch.cat(marks).contiguous()
    assert every.numel() >= 6  # three repeated unitigs: both copies of each of their k-mers are fenced off
    check(L.sgc_store_apply_marks(ctx._h, ctx.ptr(every), every.numel(), ctx.ptr(brk_all), bstride))
    ptrs = [t.bucket_ptr()[0] for t in tables]
    for t in tables:
        t.set_peers(ptrs, bounds, useq_all.data_ptr() + 4 * STORE_PAD_WORDS, ustride, brk_all.data_ptr(), bstride)
    # the oracle's answer
    ubuf, uoff = pack_sequences(useqs)
    hs, koff = O.hash_batch(ubuf, uoff, k)
    pid = np.repeat(ids, np.diff(koff))
    order = np.argsort(hs, kind="stable")
    hs, pid = hs[order], pid[order]
    first = np.flatnonzero(np.concatenate([[True], hs[1:] != hs[:-1]]))
    lo_id, hi_id = np.minimum.reduceat(pid, first), np.maximum.reduceat(pid, first)
    keep = lo_id == hi_id  # build_mphf: a hash seen under two different ids is dropped
    assert (~keep).sum() >= 3
    otab = O.COracleTable(hs[first][keep], lo_id[keep])
    o_off, o_ids, o_nk, o_nh = otab.label_reads(rbuf, roff, k)
    for packed in (False, True):
        for t in (tables[0], tables[3]):
            res = t.label_reads(rbuf, roff, k, packed=packed)
            assert np.array_equal(res["ids_off"], o_off) and np.array_equal(res["ids"], o_ids), packed
            assert (res["n_kmers"], res["n_hits"]) == (o_nk, o_nh)
    # the replicated miss filter (k-mers probed on their own whose owner is another rank are looked up in it first):
    # same labels with a filter of ~16 bits per hash, and with a deliberately tiny one (nearly every bit set: it
    # filters nothing and must not change anything either)
    for fbits in (max(3, (len(hs) // 4).bit_length()), 3):
        with torch.cuda.stream(ctx.stream):
            filt = torch.zeros(1 << fbits, dtype=torch.int64, device=ctx.device)
        for o in range(R):
            with torch.cuda.stream(ctx.stream):
                hh = torch.cat([x[0] for x in per_owner[o]]).contiguous()
            check(L.sgc_filter_insert(ctx._h, ctx.ptr(hh), hh.numel(), ctx.ptr(filt), fbits))
        ctx.sync()
        if fbits > 3:  # an owner's hashes fill its own quarter of the words, and the filter is not saturated
            set_bits = int(sum(bin(int(x) & 0xFFFFFFFFFFFFFFFF).count("1") for x in filt.cpu().numpy().view(np.uint64)[:256]))
            assert 0 < set_bits < 256 * 40
        for me in (0, 3):
            tables[me].set_filter(filt.data_ptr(), fbits, me)
            for packed in (False, True):
                res = tables[me].label_reads(rbuf, roff, k, packed=packed)
                assert np.array_equal(res["ids_off"], o_off) and np.array_equal(res["ids"], o_ids), (fbits, me, packed)
                assert (res["n_kmers"], res["n_hits"]) == (o_nk, o_nh)
            tables[me].set_filter(0, 0, 0)
    for t in tables:
        t.close()


def test_partitioned_two_gpus_p2p_and_nccl_match_replicated():
    """2 ranks on 2 GPUs (skipped on a 1-GPU box): hash-range partitioned table, exchange fused into
    the kernels over NVLink peer memory (--p2p) and through NCCL all-to-all; both must give the
    labels of the replicated table on every rank."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import ROOT

    # flag protocol (default: one call per step, ranks synchronised through peer-memory stamps), the host-mediated
    # P2P path (counts exchange + barrier) and the NCCL all-to-all baseline
    # ... and first the product path: the owners' tables probed over peer loads + the replicated sequence store
    for extra, env in ((["--peer"], {}), (["--p2p"], {}), (["--p2p"], {"SGC_PART_FLAGS": "0"}), ([], {})):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "scripts", "partitioned_run.py"),
               "--table-kmers", "2e7", "--reads", "3e5", "--steps", "3", "--check"] + extra
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, **env})
        assert out.returncode == 0, out.stderr[-2000:]
        assert "labels identical to the replicated table on every rank: True" in out.stdout


def test_sharded_queries_and_abundance_nccl(tmp_path):
    """SURVEY 8e, replicated table: query_batch(distributed=True) shards whole queries over the ranks and all-gathers
    the sparse results; abundance_sharded all-reduces the per-unitig count vectors.  One rank per visible GPU (up to
    2; a 1-GPU box runs the same NCCL code path with world size 1)."""
    import subprocess
    import sys

    import torch

    from conftest import ROOT

    n = min(torch.cuda.device_count(), 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29549", os.path.join(ROOT, "scripts", "sharded_query_run.py"), str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "sharded queries identical to the single-process batch on every rank: True" in out.stdout


# ---------------------------------------------------------------- next rows f3 / f4: sourmash-style hashing
def test_unitig_min_hashvals_match_reference_sqlite_column(ctx, dory):
    """All 736 ``min_hashval`` values the reference stored for the dory unitigs
    (cdbg/sort_bcalm_unitigs.py:56-128) from the GPU MinHash mode."""
    from spacegraphcats_b200.cdbg.minhash import unitig_min_hashvals

    with open(golden_path("dory_k21_min_hashval.json")) as fp:
        want = {int(a): b for a, b in json.load(fp).items()}
    recs = dory[0]
    got = unitig_min_hashvals([r.sequence for r in recs], 21)
    assert got.tolist() == [want[int(r.name)] for r in recs]


def test_unitig_ids_follow_the_reference_remapping(ctx, dory):
    """cdbg/sort_bcalm_unitigs.py:101-118: ids = rank of the unitig in the TEXT order of min_hashval.  The dory fixture's
    unitigs, shuffled (BCALM's order is arbitrary), get back exactly the ids the reference gave them."""
    from spacegraphcats_b200.cdbg.minhash import assign_unitig_ids

    recs = dory[0]
    perm = np.random.default_rng(3).permutation(len(recs))
    new_id, mins = assign_unitig_ids([recs[i].sequence for i in perm], 21)
    assert new_id.tolist() == [int(recs[i].name) for i in perm]
    assert sorted(str(int(m)) for m in mins) != [str(v) for v in sorted(int(m) for m in mins)]  # TEXT order != numeric order
    with pytest.raises(ValueError):
        assign_unitig_ids([recs[0].sequence, recs[0].sequence], 21)


def test_unitig_ids_from_raw_bcalm_output(ctx, O, dory):
    """f4 from BCALM's own file (data/bcalm.dory.k21.unitigs.fa: arbitrary order, BCALM's numbering, either strand):
    assign_unitig_ids gives every unitig the id it has in the reference's dory cDBG (the reference's own
    sort_bcalm_unitigs.main does the same on the CPU shims: tests/test_refshim_cpu.py)."""
    from spacegraphcats_b200.cdbg.minhash import assign_unitig_ids

    recs = O.read_records(golden_path("data", "bcalm.dory.k21.unitigs.fa.gz"))
    assert len(recs) == 736
    canon = lambda s: min(s, O.revcomp(s.encode()).decode())  # noqa: E731
    want = {canon(r.sequence): int(r.name) for r in dory[0]}
    new_id, _ = assign_unitig_ids([r.sequence for r in recs], 21)
    assert new_id.tolist() == [want[canon(r.sequence)] for r in recs]
    assert new_id.tolist() != [int(r.name.split()[0]) for r in recs]  # BCALM's own numbering is another one


@pytest.mark.parametrize("k", [21, 31, 12, 33])
def test_sourmash_hashes_vs_oracle(ctx, O, k):
    from spacegraphcats_b200.cdbg.minhash import scaled_minhash, sourmash_hashes, unitig_min_hashvals

    rng = np.random.default_rng(k)
    seqs = [rand_seq(rng, int(n)) for n in [k, k + 1, 0, 5, 150, 700, 3000] + rng.integers(0, 400, 60).tolist()]
    seqs.append("ACGT" * 50)  # k-mer == reverse complement ties for even k
    h, koff = sourmash_hashes(seqs, k)
    want = [O.minhash_sequence_py(s, k) if len(s) >= k else [] for s in seqs]
    assert koff.tolist() == np.concatenate([[0], np.cumsum([len(w) for w in want])]).tolist()
    assert h.tolist() == [x for w in want for x in w]
    mins = unitig_min_hashvals(seqs, k)
    assert mins.tolist() == [min(w) if w else 2**64 - 1 for w in want]
    sk = scaled_minhash(seqs, k, 10)
    assert sk.tolist() == sorted({x for w in want for x in w if x <= (2**64 - 1) // 10})
    with pytest.raises(ValueError):
        sourmash_hashes(["ACGTNACGTACGTACGTACGTAGCTAGCTAGCTAGCTAGCATCGATCGATCGATG"], k)


# ---------------------------------------------------------------- configs[1]/[2]: twofoo-short and shew/acido proxies
def test_batched_queries_k31_pseudo_unitigs_vs_oracle(ctx, O):
    """BASELINE.json configs[1] and [2] with the offline proxies of SURVEY.md 8d: a k=31 table from
    pseudo-unitigs of the twofoo-short genomes + acido chunk 1, queried in one batch by
    shew-os223-200k, acido-chunk1/2, {2,47,63}.short and twofoo-genes; per-query {unitig: distinct
    matching k-mers} against the oracle's Query restatement; reads of twofoo-short labelled too."""
    from spacegraphcats_b200.cdbg.hashing import MPHF_KmerIndex
    from spacegraphcats_b200.cdbg.index_cdbg_by_kmer import build_mphf
    from spacegraphcats_b200.device import pack_sequences
    from spacegraphcats_b200.search.query_by_sequence import pseudo_unitigs, query_batch

    k = 31
    data = lambda f: golden_path("data", f)  # noqa: E731
    genome_files = ["63.short.fa.gz", "47.short.fa.gz", "acido-chunk1.fa.gz"]
    gseqs = [r.sequence for f in genome_files for r in O.read_records(data(f))]
    pieces = pseudo_unitigs(gseqs, k)
    urecs = [O.Record(str(i), s) for i, s in enumerate(pieces)]
    table, sizes = build_mphf(k, lambda: iter(urecs), ctx=ctx, verbose=False)
    idx = MPHF_KmerIndex(k, table, sizes)
    otab, osizes, n_multi = O.build_mphf(k, lambda: iter(urecs))
    assert sizes == osizes and len(table) == len(otab)
    assert table.stats()[1] == n_multi

    # configs[2] in full: shew-os223-200k + all eight acido chunks, plus the twofoo queries, as ONE batch
    qfiles = [data(f) for f in ("shew-os223-200k.fa.gz",) + tuple("acido-chunk%d.fa.gz" % i for i in range(1, 9)) +
              ("2.short.fa.gz", "47.short.fa.gz", "63.short.fa.gz", "twofoo-genes.fa.gz")]
    res = query_batch(qfiles, k, None, idx)
    containment = {}
    for (q, _), qf in zip(res, qfiles):
        okm = O.query_kmers(O.read_records(qf), k)
        want = O.count_cdbg_matches(otab, okm)
        assert q.cdbg_match_counts == want, qf
        assert q.n_kmers == len(okm)
        containment[os.path.basename(qf)] = sum(want.values()) / len(okm)
    # sanity of the proxy: an indexed genome is (almost) fully contained, an unrelated one is not; the
    # closely related 47/63 genomes share k-mers, which build_mphf drops as multicounts (~half of 63)
    assert containment["acido-chunk1.fa.gz"] > 0.95 and containment["shew-os223-200k.fa.gz"] < 0.05
    assert n_multi > 1000 and 0.2 < containment["63.short.fa.gz"] < 0.8

    reads = O.read_records(data("twofoo-short-reads.fa.gz"))
    rbuf, roff = pack_sequences([r.sequence for r in reads])
    lab = table.label_reads(rbuf, roff, k)
    o_off, o_ids, o_nk, o_nh = O.COracleTable(otab.mphf_to_hash, otab.mphf_to_value).label_reads(rbuf, roff, k)
    assert np.array_equal(lab["ids_off"], o_off) and np.array_equal(lab["ids"], o_ids)
    assert (lab["n_kmers"], lab["n_hits"]) == (o_nk, o_nh)


def test_count_matches_batch_sets_per_query(ctx, O):
    """sgc_query_count_batch: every query is a SET of hashes (duplicates inside a record, across the records of a
    query and palindromic / reverse-complement repeats count once), the same k-mer in two queries counts in
    both, empty queries and queries of too-short records give empty dicts; against the oracle's Query."""
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import pack_sequences

    rng = np.random.default_rng(77)
    for k in (21, 31, 12):
        lut = np.frombuffer(b"ACGT", np.uint8)
        g = lut[rng.integers(0, 4, 30_000)].tobytes().decode()
        cuts = list(range(0, len(g) - k, 97))
        useqs = [g[a : a + 97 + k - 1] for a in cuts]
        ubuf, uoff = pack_sequences(useqs)
        table = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx)
        otab, _, _ = O.build_mphf(k, lambda: iter(O.Record(str(i), s) for i, s in enumerate(useqs)))
        rc = O.revcomp(g[500:900].encode()).decode()
        queries = [
            [g[100:400], g[100:400], g[350:600]],            # duplicates inside and across records
            [],                                              # no records at all
            [g[100:400], rc, g[500:900]],                    # a region and its reverse complement: same hashes
            ["ACGT"[: k - 1], "A"],                          # only records shorter than k
            ["A" * (k + 40), "ACGT" * 40],                   # low complexity: many identical hashes
            [g[5000:25000]],                                 # one long record
            [lut[rng.integers(0, 4, 3000)].tobytes().decode()],  # unrelated sequence: no (or chance) matches
        ]
        got = table.count_matches_batch(queries, k)
        assert len(got) == len(queries)
        for qi, (q, (counts, n_distinct)) in enumerate(zip(queries, got)):
            okm = O.query_kmers([O.Record("r", s) for s in q], k)
            want = O.count_cdbg_matches(otab, okm)
            assert n_distinct == len(okm), (k, qi)
            assert counts == want, (k, qi)
    assert table.count_matches_batch([], 21) == []


def test_count_matches_batch_many_queries_retry_and_fallback(ctx, O, monkeypatch):
    """The store-position form of sgc_query_count_batch off its beaten path: thousands of queries (32-bit miss keys,
    per-query counters in global memory), record buffers that overflow and are re-sized, and a table that also holds
    hashes without a store position (filled hash by hash), which must fall back to the hash-sort form."""
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import pack_sequences

    rng = np.random.default_rng(91)
    k = 31
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[rng.integers(0, 4, 400_000)].tobytes().decode()
    useqs = [g[a : a + 150 + k - 1] for a in range(0, len(g) - 200, 150)]
    ubuf, uoff = pack_sequences(useqs)
    table = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx)
    otab, _, _ = O.build_mphf(k, lambda: iter(O.Record(str(i), s) for i, s in enumerate(useqs)))

    def check(queries, tab=table, ot=otab):
        got = tab.count_matches_batch(queries, k)
        assert len(got) == len(queries)
        for qi, (q, (counts, n_distinct)) in enumerate(zip(queries, got)):
            okm = O.query_kmers([O.Record("r", s) for s in q], k)
            assert n_distinct == len(okm), qi
            assert counts == O.count_cdbg_matches(ot, okm), qi

    def mutate(s):
        b = bytearray(s.encode())
        for p in rng.integers(0, len(b), max(1, len(b) // 150)):
            b[p] = ord("ACGT"[(("ACGT".index(chr(b[p]))) + 1) % 4])
        return b.decode()

    for nq in (5000, 9000):  # > 4096: 32-bit keys for the missing k-mers; > 8192: no shared-memory counters
        starts = rng.integers(0, len(g) - 400, nq)
        check([[mutate(g[a : a + int(n)])] for a, n in zip(starts, rng.integers(k, 400, nq))])
    few = [[mutate(g[a : a + 30_000]), g[a + 100 : a + 900]] for a in rng.integers(0, len(g) - 31_000, 6)]
    monkeypatch.setenv("SGC_QUERY_SMALL_CAPS", "1")
    check(few)  # the first attempt overflows its interval and miss buffers
    monkeypatch.delenv("SGC_QUERY_SMALL_CAPS")
    # a hash set by hand has no place in the sequence store: the batch falls back to sorting hashes
    extra = "".join(rng.choice(list("ACGT"), 60))
    eh = O.hash_sequence(extra[:k], k)[0]
    table[eh] = 7
    o2 = dict(zip(otab.mphf_to_hash.tolist(), otab.mphf_to_value.tolist()))
    o2[eh] = 7
    got = table.count_matches_batch(few + [[extra]], k)
    okm = O.query_kmers([O.Record("r", extra)], k)
    want = {}
    for h in okm:
        if h in o2:
            want[o2[h]] = want.get(o2[h], 0) + 1
    assert got[-1] == (want, len(okm)) and want.get(7) == 1
    for q, (counts, n_distinct) in zip(few, got):
        okm = O.query_kmers([O.Record("r", s) for s in q], k)
        assert n_distinct == len(okm) and counts == O.count_cdbg_matches(otab, okm)


# ---------------------------------------------------------------- the reference-side binding
def test_reference_side_binding_drives_the_reference_loops(tmp_path, O, dory):
    """tests/refshim/sgc_b200_binding.py -- the raw ctypes stand-in for ``khmer`` / ``bbhash_table`` a reference
    maintainer would add -- driven through the reference's build_mphf loop
    (index_cdbg_by_kmer.py:27-115, restated in oracle.build_mphf; the reference's own file runs on the same shim
    surface in tests/test_refshim_cpu.py where /root/reference exists): KATs of test_dory_workflow.py:195-199,
    the whole table, get_unique_values semantics, save / load, and the reference-format files it writes."""
    import sys

    from conftest import ROOT

    sys.path.insert(0, os.path.join(ROOT, "tests", "refshim"))
    try:
        import sgc_b200_binding as B
    finally:
        sys.path.pop(0)
    recs, ref_h, ref_v, ref_sizes = dory
    nt = B.Nodetable(21, 1, 1)
    assert nt.get_kmer_hashes("GACCAAAGGCTTACGGTGAGC") == [10218271035842461694]
    assert nt.get_kmer_hashes(recs[5].sequence) == O.hash_sequence(recs[5].sequence, 21).tolist()
    with pytest.raises(ValueError):
        nt.get_kmer_hashes("ACGT")
    with pytest.raises(ValueError):
        nt.get_kmer_hashes("ACGTN" * 10)
    hash_fn = lambda s, k: B.Nodetable(k, 1, 1).get_kmer_hashes(s)  # noqa: E731
    table, sizes, n_multi = O.build_mphf(21, lambda: iter(recs), hash_fn=hash_fn, table_cls=B.BBHashTable)
    assert sizes == ref_sizes and n_multi == 0 and len(table) == 212475
    assert table[10218271035842461694] == 118 and table[8436068710919520258] == 118
    assert table[13994045974119358468] == 118 and table[11971930231572094512] == 187 and table[12345] is None
    order = np.argsort(ref_h)
    assert np.array_equal(table.mphf_to_hash, ref_h[order]) and np.array_equal(table.mphf_to_value, ref_v[order])
    assert int(max(table.mphf_to_value)) == 735  # index_cdbg_by_kmer.py:106
    q = nt.get_kmer_hashes(recs[118].sequence[:60])
    assert table.get_unique_values(q + q + [7]) == {118: 2 * len(q)}
    assert table.get_unique_values(set(q)) == {118: len(q)} and table.get_unique_values([]) == {}
    with pytest.raises(ValueError):
        table.get_unique_values(q + [7], require_exist=True)
    mphf, arr = str(tmp_path / "contigs.mphf"), str(tmp_path / "contigs.indices")
    table.save(mphf, arr)
    assert open(mphf, "rb").read() == open(golden_path("dory_k21_contigs.mphf"), "rb").read()  # byte-identical boomphf dump
    again = B.BBHashTable.load(mphf, arr)
    assert len(again) == 212475 and again[11971930231572094512] == 187
    assert again.get_unique_values(q) == {118: len(q)}


# ---------------------------------------------------------------- C-ABI error behaviour
def test_abi_error_paths(ctx):
    """Every failure is a negative status + message, never an exception from the library or a crash."""
    import ctypes

    import torch

    from spacegraphcats_b200 import _lib
    from spacegraphcats_b200.device import DeviceTable, pack_sequences

    L = ctx._L
    buf, off = pack_sequences(["ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"] * 3)
    d_seq, d_off = ctx.to_device(buf), ctx.to_device(off)
    d_out = ctx.empty(200, torch.int64)
    d_koff = ctx.empty(4, torch.int64)
    tot = ctypes.c_int64(0)

    def msg():
        return L.sgc_last_error().decode()

    # k out of range
    for k in (0, 256, -3):
        rc = L.sgc_hash_batch(ctx._h, ctx.ptr(d_seq), d_seq.numel(), ctx.ptr(d_off), 3, k, ctx.ptr(d_out), 200,
                              ctx.ptr(d_koff), None, ctypes.byref(tot))
        assert rc == _lib.SGC_ERR_ARG and "k must be" in msg()
    # output too small: reports the needed size, writes nothing
    rc = L.sgc_hash_batch(ctx._h, ctx.ptr(d_seq), d_seq.numel(), ctx.ptr(d_off), 3, 21, ctx.ptr(d_out), 10,
                          ctx.ptr(d_koff), None, ctypes.byref(tot))
    assert rc == _lib.SGC_ERR_CAPACITY and tot.value == 60
    # NULL handles
    assert L.sgc_hash_batch(None, None, 0, None, 0, 21, None, 0, None, None, None) == _lib.SGC_ERR_ARG
    assert L.sgc_lookup(None, None, 0, None) == _lib.SGC_ERR_ARG
    assert L.sgc_label_reads_finish(None, None, None, None) == _lib.SGC_ERR_ARG
    # device index out of range
    h = ctypes.c_void_p()
    assert L.sgc_ctx_create(99, None, ctypes.byref(h)) == _lib.SGC_ERR_ARG and not h.value
    # finish without launch
    t = DeviceTable(ctx, 1000)
    assert L.sgc_label_reads_finish(t._h, None, None, None) == _lib.SGC_ERR_ARG and "no label call pending" in msg()
    # ranks must be a power of two
    d_boff = ctx.empty(8, torch.int64)
    rc = L.sgc_partition_pairs(ctx._h, ctx.ptr(d_out), None, 10, 3, ctx.ptr(d_out), None, ctx.ptr(d_out), ctx.ptr(d_boff), None)
    assert rc == _lib.SGC_ERR_ARG and "power of two" in msg()
    # table load out of range
    assert L.sgc_table_create(ctx._h, 1000, ctypes.c_float(5.0), ctypes.byref(h)) == _lib.SGC_ERR_ARG
    # ids above SGC_MAX_ID are rejected
    with pytest.raises(_lib.SgcError):
        t.insert_pairs(np.array([1, 2], np.uint64), np.array([5, 0xFFFFFFFE], np.uint32))
    # label buffer too small: the Python layer retries with the size the library reports
    big = DeviceTable(ctx, 10000)
    rng = np.random.default_rng(12)
    seqs = [rand_seq(rng, 200) for _ in range(50)]
    ub, uo = pack_sequences([s[a : a + 31] for s in seqs for a in range(0, 170)])  # one unitig per k-mer
    big.insert_sequences(ub, uo, 31)
    rb, ro = pack_sequences(seqs)
    res = big.label_reads(rb, ro, 31, ids_cap=4)
    assert res["n_labels"] > 4 and len(res["ids"]) == res["n_labels"]
    # offsets that run past the buffer (or go backwards) are reported, and nothing outside it is read
    bad_off = ctx.to_device(np.array([0, 40, 10_000_000, 10_000_040], np.int64))
    rc = L.sgc_hash_batch(ctx._h, ctx.ptr(d_seq), d_seq.numel(), ctx.ptr(bad_off), 3, 21, ctx.ptr(d_out), 200,
                          ctx.ptr(d_koff), None, ctypes.byref(tot))
    assert rc == _lib.SGC_ERR_ARG and "offsets" in msg()
    with pytest.raises(_lib.SgcError):
        big.label_reads(rb, np.array([0, 100, 50, len(rb)], np.int64), 31)
    # two contexts on one device coexist
    from spacegraphcats_b200.device import Context

    c2 = Context(ctx.device_index)
    h2, _, _ = c2.hash_batch(buf, off, 21)
    h1, _, _ = ctx.hash_batch(buf, off, 21)
    assert np.array_equal(h1, h2)
    c2.close()


# ---------------------------------------------------------------- randomized properties (hypothesis)
def test_hypothesis_hash_and_label_vs_oracle(ctx, O):
    """Random batches (random k in 1..64, ragged lengths incl. empty, occasional non-ACGT bytes):
    hashes, status flags and per-read labels (both label kernels) against the oracle."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import pack_sequences

    base = st.text(alphabet="ACGT", min_size=0, max_size=300)
    dirty = st.text(alphabet="ACGTNacgt", min_size=0, max_size=80)
    seqs_st = st.lists(st.one_of(base, base, base, dirty), min_size=0, max_size=40)

    @settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
    @given(seqs=seqs_st, k=st.integers(min_value=1, max_value=64), side=st.booleans())
    def run(seqs, k, side):
        buf, off = pack_sequences(seqs)
        h, koff, status = ctx.hash_batch(buf, off, k)
        want, woff = O.hash_batch(buf, off, k)
        assert np.array_equal(koff, woff) and np.array_equal(h, want)
        # flagged <=> the sequence has k-mers and some base is not A/C/G/T
        assert status.tolist() == [int(len(s) >= k and any(c not in "ACGT" for c in s)) for s in seqs]
        clean = [s for s in seqs if all(c in "ACGT" for c in s)]
        unitigs = [s for s in clean[: max(1, len(clean) // 2)] if len(s) >= k]
        if not unitigs:
            return
        ubuf, uoff = pack_sequences(unitigs)
        table = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx, side=side)
        # build_mphf's table without its closing asserts (they reject inputs whose last unitig loses all
        # of its k-mers, which random tiny inputs do produce): hashes under two ids are dropped
        uh, ukoff = O.hash_batch(ubuf, uoff, k)
        uid = np.repeat(np.arange(len(unitigs), dtype=np.uint32), np.diff(ukoff))
        pairs = np.unique(np.stack([uh, uid.astype(np.uint64)], 1), axis=0)
        hs, cnt = np.unique(pairs[:, 0], return_counts=True)
        keep = np.isin(pairs[:, 0], hs[cnt == 1])
        oh, ov = pairs[keep, 0], pairs[keep, 1].astype(np.uint32)
        assert len(table) == len(oh) and table.stats()[1] == int((cnt > 1).sum())
        cbuf, coff = pack_sequences(clean)
        lab = table.label_reads(cbuf, coff, k)
        if len(oh) == 0:
            assert lab["n_hits"] == 0 and lab["n_labels"] == 0
            return
        o_off, o_ids, o_nk, o_nh = O.COracleTable(oh, ov).label_reads(cbuf, coff, k)
        assert (lab["n_kmers"], lab["n_hits"]) == (o_nk, o_nh)
        assert np.array_equal(lab["ids_off"], o_off) and np.array_equal(lab["ids"], o_ids)
        labp = table.label_reads(cbuf, coff, k, packed=True)
        assert (labp["n_kmers"], labp["n_hits"]) == (o_nk, o_nh)
        assert np.array_equal(labp["ids_off"], o_off) and np.array_equal(labp["ids"], o_ids)

    run()


# ---------------------------------------------------------------- K8: BBHash (boomphf) writer
@pytest.fixture(scope="module")
def dory_dump():
    return open(golden_path("dory_k21_contigs.mphf"), "rb").read()


def test_mphf_device_build_is_byte_identical_to_reference_dump(ctx, dory, dory_dump):
    """The reference's contigs.mphf and contigs.indices for dory, rebuilt on the GPU from the keys."""
    from spacegraphcats_b200.device import DeviceMphf

    _, ref_h, ref_v, _ = dory
    perm = np.random.default_rng(11).permutation(len(ref_h))
    m = DeviceMphf.build(ctx, ref_h[perm])
    assert m.info() == {"nelem": 212475, "nb_levels": 25, "lastbitsetrank": 212475, "n_final": 0, "gamma": 1.0}
    assert m.dumps() == dory_dump
    oh, ov = m.order(ref_h[perm], ref_v[perm])
    assert np.array_equal(oh, ref_h) and np.array_equal(ov, ref_v)
    m.close()


def test_mphf_device_lookup_through_reference_dump(ctx, dory, dory_dump):
    """mphf_to_hash[i] must look up to i through the bitsets the reference wrote."""
    from spacegraphcats_b200.device import DeviceMphf

    _, ref_h, _, _ = dory
    m = DeviceMphf.loads(ctx, dory_dump)
    assert np.array_equal(m.lookup(ref_h), np.arange(len(ref_h), dtype=np.uint64))
    assert m.dumps() == dory_dump
    m.close()


def test_index_save_writes_the_reference_files(ctx, dory, dory_index, dory_dump, tmp_path):
    """build_mphf from the unitig sequences + MPHF_KmerIndex.save == the files the reference wrote
    (contigs.mphf byte for byte, contigs.indices array for array, contigs.sizes)."""
    import pickle
    from spacegraphcats_b200.cdbg.hashing import MPHF_KmerIndex

    _, ref_h, ref_v, ref_sizes = dory
    dory_index.save(str(tmp_path))
    assert open(tmp_path / "contigs.mphf", "rb").read() == dory_dump
    with np.load(tmp_path / "contigs.indices") as z:
        assert sorted(z.files) == ["mphf_to_hash", "mphf_to_value"]
        assert z["mphf_to_hash"].dtype == np.uint64 and z["mphf_to_value"].dtype == np.uint32
        assert np.array_equal(z["mphf_to_hash"], ref_h) and np.array_equal(z["mphf_to_value"], ref_v)
    with open(tmp_path / "contigs.sizes", "rb") as fp:
        k, sizes = pickle.load(fp)
    assert k == 21 and sizes == ref_sizes
    again = MPHF_KmerIndex.from_directory(str(tmp_path))
    assert len(again.table) == len(ref_h) and again.table[int(ref_h[5])] == int(ref_v[5])


@pytest.mark.parametrize("n,levels,gamma", [(0, 25, 1.0), (1, 25, 1.0), (2, 25, 1.0), (65, 25, 1.0), (5000, 4, 1.0),
                                            (5000, 2, 1.0), (30000, 25, 2.0), (100003, 25, 1.0)])
def test_mphf_device_vs_oracle(ctx, n, levels, gamma):
    from oracle import bbhash
    from spacegraphcats_b200.device import DeviceMphf

    rng = np.random.default_rng(n + levels)
    keys = np.unique(rng.integers(0, 2**64, size=n, dtype=np.uint64))
    want = bbhash.Mphf.build(keys, gamma=gamma, nb_levels=levels)
    m = DeviceMphf.build(ctx, keys, gamma=gamma, nb_levels=levels)
    assert m.info()["n_final"] == len(want.final)
    assert m.dumps() == want.dumps()
    idx = m.lookup(keys)
    assert np.array_equal(idx, want.lookup(keys))
    assert sorted(idx.tolist()) == list(range(len(keys)))
    others = rng.integers(0, 2**64, size=1000, dtype=np.uint64)
    assert np.array_equal(m.lookup(others), want.lookup(others))  # an MPHF answers for foreign keys too
    m2 = DeviceMphf.loads(ctx, m.dumps())
    assert np.array_equal(m2.lookup(keys), idx)
    m.close(), m2.close()


def test_mphf_large_is_a_bijection(ctx):
    """Size-independent property at 2^26 keys: indices are exactly a permutation of 0..n-1."""
    import torch
    from spacegraphcats_b200.device import DeviceMphf

    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(5)
    keys = torch.randint(-(2**63), 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    keys = torch.unique(keys)
    m = DeviceMphf.build(ctx, keys)
    info = m.info()
    assert info["nelem"] == keys.numel() and info["lastbitsetrank"] + info["n_final"] == keys.numel()
    idx = m.lookup(keys, to_host=False)
    ctx.sync()
    srt = torch.sort(idx).values
    assert torch.equal(srt, torch.arange(keys.numel(), dtype=torch.int64, device="cuda"))
    m.close()


def test_mphf_error_paths(ctx, dory_dump):
    from spacegraphcats_b200._lib import SgcError
    from spacegraphcats_b200.device import DeviceMphf

    keys = np.arange(1, 1001, dtype=np.uint64)
    with pytest.raises(SgcError):
        DeviceMphf.build(ctx, keys, gamma=0.5)
    with pytest.raises(SgcError):
        DeviceMphf.build(ctx, keys, nb_levels=1)
    with pytest.raises(SgcError):
        DeviceMphf.loads(ctx, dory_dump[:-1])
    with pytest.raises(SgcError):
        DeviceMphf.loads(ctx, dory_dump[:1000])
    with pytest.raises(SgcError):
        DeviceMphf.loads(ctx, b"SGC_B200_NO_MPHF\n1\n")
    m = DeviceMphf.build(ctx, keys)
    with pytest.raises(SgcError):
        m.order(keys[:10], np.zeros(10, np.uint32))  # n != nelem
    # duplicate keys never settle: they end in the final map, the build still terminates
    d = DeviceMphf.build(ctx, np.array([7, 7, 9], np.uint64))
    assert d.info()["n_final"] == 2
    m.close(), d.close()


def test_label_tail_small_runs_and_sorted_fallback(ctx, O, monkeypatch):
    """The label tail groups candidate ids per sequence without a global sort when every sequence has
    few of them, and falls back to the radix-sort path as soon as one sequence has many (long contigs):
    both must give the oracle's CSR, alone and mixed in one batch."""
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import pack_sequences

    rng = np.random.default_rng(33)
    k = 21
    g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 60_000)]
    cuts = np.concatenate([[0], np.cumsum(rng.integers(1, 40, 5000))])
    cuts = cuts[cuts < len(g) - k]
    useqs = [g[a : b + k - 1].tobytes() for a, b in zip(cuts[:-1], cuts[1:])]
    ubuf, uoff = pack_sequences(useqs)
    otab, _, _ = O.build_mphf(k, lambda: iter(O.Record(str(i), s.decode()) for i, s in enumerate(useqs)))
    ot = O.COracleTable(otab.mphf_to_hash, otab.mphf_to_value)
    short = [g[p : p + 100].tobytes() for p in rng.integers(0, len(g) - 6000, 3000)]       # <= ~6 ids each
    longs = [g[p : p + 5000].tobytes() for p in rng.integers(0, len(g) - 6000, 5)]         # hundreds of ids each
    edge = [g[p : p + 160 * 20].tobytes() for p in rng.integers(0, len(g) - 6000, 50)]     # around the 160-id limit
    for side in (True, False):
        table = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx, side=side)
        for batch in (short, short[:500] + longs + short[500:1000], edge + short[:100], longs):
            buf, off = pack_sequences(batch)
            o_off, o_ids, o_nk, o_nh = ot.label_reads(buf, off, k)
            res = table.label_reads(buf, off, k)
            assert (res["n_kmers"], res["n_hits"]) == (o_nk, o_nh)
            assert np.array_equal(res["ids_off"], o_off) and np.array_equal(res["ids"], o_ids)
            monkeypatch.setenv("SGC_CSR_SORT", "1")  # force the general path: same answer
            res2 = table.label_reads(buf, off, k)
            monkeypatch.delenv("SGC_CSR_SORT")
            assert np.array_equal(res2["ids_off"], o_off) and np.array_equal(res2["ids"], o_ids)
    assert max(np.diff(o_off)) > 24  # the last batch (long contigs) really exceeded the per-thread limit
