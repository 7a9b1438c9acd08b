"""Pins the CPU oracle against the reference's own golden fixtures.

CPU-only.  Everything the GPU parity tests trust is first proven here.
"""
import gzip
import json

import numpy as np
import pytest

from oracle import oracle as O
from conftest import golden_path


# ---- Murmur3 itself ---------------------------------------------------------
def test_murmur3_public_vector():
    # mmh3.hash64("foo") == (-2129773440516405919, 9128664383759220103)
    want = ((-2129773440516405919) & (2**64 - 1), 9128664383759220103)
    assert O.murmur3_x64_128(b"foo") == want
    assert O.murmur3_c(b"foo") == want


def test_murmur3_py_vs_c_all_tail_lengths():
    rng = np.random.default_rng(7)
    for n in list(range(0, 70)) + [127, 128, 255]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for seed in (0, 42):
            assert O.murmur3_x64_128(data, seed) == O.murmur3_c(data, seed), n


def test_sourmash_min_hashval_pins_tail_path():
    """sqlite min_hashval = sourmash Murmur3 seed 42 over min(kmer, rc), k=21 -- an
    independent check of the restated Murmur3 (we only check a handful here; cheap)."""
    # value for unitig 0 from tests/test-data/catlas.dory_k21/bcalm.unitigs.db
    recs = O.read_unitigs_tsv(golden_path("dory_k21_unitigs.tsv.gz"))
    seq = recs[0].sequence.encode()
    vals = []
    for i in range(len(seq) - 20):
        km = seq[i : i + 21]
        rc = O.revcomp(km)
        vals.append(O.murmur3_x64_128(min(km, rc), 42)[0])
    assert min(vals) == 10056084056166678


# ---- hash KATs ---------------------------------------------------------------
KATS_K21 = {
    "GACCAAAGGCTTACGGTGAGC": 10218271035842461694,  # tests/test_dory_workflow.py:196
    "TCTCCGTTTGCAGAAATCGTA": 8436068710919520258,  # :197
    "AGCAGGTACAATGGACTGGAT": 13994045974119358468,  # :198
    "TTCGGCAAGAAAATTTTAATA": 11971930231572094512,  # :199
    "AGACCAACCTTTAGCCTATTC": 13216357780530054207,  # SURVEY 8c derived
    "ACGTACGTACGTACGTACGTA": 1083786972790998927,
}
KATS_K31 = {
    "AGACCAACCTTTAGCCTATTCAAGACGAGTG": 505484382184005004,
    "GACCAACCTTTAGCCTATTCAAGACGAGTGT": 17407874517182356438,
    "A" * 31: 8676834403036734983,
}


@pytest.mark.parametrize("kats", [KATS_K21, KATS_K31])
def test_hash_kats(kats):
    for kmer, want in kats.items():
        assert O.hash_kmer_py(kmer) == want
        assert int(O.hash_sequence(kmer, len(kmer))[0]) == want


def test_even_k_palindrome_hashes_to_zero():
    assert O.hash_kmer_py("ACGT" * 8) == 0
    assert int(O.hash_sequence("ACGT" * 8, 32)[0]) == 0


def test_fwd_rc_components_of_first_kat():
    assert O.murmur3_x64_128(b"GACCAAAGGCTTACGGTGAGC")[0] == 11421216072548113777
    assert O.murmur3_x64_128(O.revcomp(b"GACCAAAGGCTTACGGTGAGC"))[0] == 1391287788086436495


def test_hash_sequence_py_vs_c_many_k():
    rng = np.random.default_rng(11)
    seq = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
    for k in (1, 2, 7, 8, 9, 15, 16, 17, 21, 24, 25, 31, 32, 33, 47, 48, 49, 63, 64, 65, 255):
        assert [int(x) for x in O.hash_sequence(seq, k)] == O.hash_sequence_py(seq, k), k
    # reverse-complement invariance (canonical hash)
    rc = O.revcomp(seq.encode()).decode()
    assert [int(x) for x in O.hash_sequence(rc, 31)][::-1] == O.hash_sequence_py(seq, 31)


def test_hash_sequence_short_raises():
    with pytest.raises(ValueError):
        O.hash_sequence("ACGT", 5)
    with pytest.raises(ValueError):
        O.hash_sequence_py("ACGT", 5)


def test_non_acgt_passthrough_consistent():
    seq = "ACGTNNacgtRYACGTACGTTTGACCA"
    assert [int(x) for x in O.hash_sequence(seq, 9)] == O.hash_sequence_py(seq, 9)


# ---- the dory fixture: whole-table parity ------------------------------------
@pytest.fixture(scope="module")
def dory():
    recs = O.read_unitigs_tsv(golden_path("dory_k21_unitigs.tsv.gz"))
    z = np.load(golden_path("dory_k21_contigs.indices.npz"))
    with open(golden_path("dory_k21_contigs.sizes.json")) as fp:
        sz = json.load(fp)
    sizes = {int(a): b for a, b in sz["sizes"].items()}
    return recs, z["mphf_to_hash"], z["mphf_to_value"], sz["ksize"], sizes


def test_dory_whole_table(dory):
    """All 212,475 (hash -> unitig id) pairs the reference wrote to contigs.indices."""
    recs, ref_h, ref_v, k, ref_sizes = dory
    assert k == 21 and len(recs) == 736 and len(ref_h) == 212475
    table, sizes, n_multi = O.build_mphf(k, lambda: iter(recs))
    assert n_multi == 0
    assert len(table) == 212475
    assert sizes == ref_sizes and sum(sizes.values()) == 212475
    want = dict(zip(ref_h.tolist(), ref_v.tolist()))
    got = dict(zip(table.mphf_to_hash.tolist(), table.mphf_to_value.tolist()))
    assert got == want
    # KAT lookups through the table (tests/test_dory_workflow.py:195-199)
    assert table[10218271035842461694] == 118
    assert table[8436068710919520258] == 118
    assert table[13994045974119358468] == 118
    assert table[11971930231572094512] == 187
    assert table[12345] is None


def test_c_table_matches_dict(dory):
    recs, ref_h, ref_v, k, _ = dory
    ct = O.COracleTable(ref_h, ref_v)
    assert len(ct) == len(ref_h)
    rng = np.random.default_rng(3)
    probe = np.concatenate([ref_h[rng.integers(0, len(ref_h), 5000)], rng.integers(0, 2**63, 5000).astype(np.uint64)])
    want = dict(zip(ref_h.tolist(), ref_v.tolist()))
    got = ct.lookup(probe)
    for h, v in zip(probe.tolist(), got.tolist()):
        assert want.get(h, O.MISS) == v


def test_build_mphf_multicounts():
    """index_cdbg_by_kmer.py:43-57: hashes in >= 2 different unitigs vanish; repeats inside one
    unitig are kept once; sizes count every position."""
    a = "ACGTTGCATGGACCATTGACAGT"  # 23 bp
    recs = [O.Record("0", a + "ACCA" + a[:10]), O.Record("1", "TTTTTTTTTTGGGGGGGGGGCCCCCCCCCAAAAAAAAATT" + a), O.Record("2", "GATTACAGATTACAGATTACAGATTACA")]
    k = 21
    table, sizes, n_multi = O.build_mphf(k, lambda: iter(recs))
    h_a = set(O.hash_sequence_py(a, k))
    assert n_multi == len(h_a) == 3
    for h in h_a:
        assert table[h] is None
    assert sizes == {0: len(recs[0].sequence) - 20, 1: len(recs[1].sequence) - 20, 2: len(recs[2].sequence) - 20}
    # GATTACA repeat: 28 bp -> 8 positions but period 7 => 7 distinct k-mers, all kept once
    h2 = O.hash_sequence_py(recs[2].sequence, k)
    assert len(h2) == 8 and len(set(h2)) == 7
    assert all(table[h] == 2 for h in h2)


# ---- query: dory-head against the dory catlas ---------------------------------
@pytest.fixture(scope="module")
def dory_index(dory):
    recs, ref_h, ref_v, k, sizes = dory
    table = O.OracleTable.from_arrays(ref_h, ref_v)
    catlas = O.OracleCAtlas(golden_path("dory_k21_r1", "catlas.csv"), golden_path("dory_k21_r1", "first_doms.txt"))
    return table, sizes, catlas, k


def test_catlas_shape(dory_index):
    _, sizes, catlas, _ = dory_index
    # tests/test_catlas.py:33-57 of the reference
    assert catlas.root == 793 and len(catlas.levels) == 794
    assert len(catlas.layer1_to_cdbg) == 580
    assert len(catlas.shadow([catlas.root])) == 736
    node_sizes = O.build_catlas_node_sizes(sizes, catlas)
    assert node_sizes[catlas.root] == 212475


def test_query_dory_head(dory_index):
    table, sizes, catlas, k = dory_index
    kmers = O.query_kmers(O.read_records(golden_path("data", "dory-head.fa")), k)
    assert len(kmers) == 1631
    cdbg, cat, doms, shadow = O.query_execute(kmers, table, sizes, catlas)
    assert cdbg == {118: 834, 187: 797}
    assert doms == {97, 150}
    assert shadow == {118, 187}
    assert cat[catlas.root] == 1631
    with gzip.open(golden_path("dory_k21_r1_search", "dory-head.fa.cdbg_ids.txt.gz"), "rt") as fp:
        assert sorted(int(x) for x in fp.read().split()) == sorted(shadow)
    with gzip.open(golden_path("dory_k21_r1_search", "dory-head.fa.frontier.txt.gz"), "rt") as fp:
        assert sorted(int(x) for x in fp.read().split()) == sorted(doms)
    # response curve rows: "834 ... 97" and "1631 ... 797 0 150"
    with open(golden_path("dory_k21_r1_search", "dory-head.fa.response.txt")) as fp:
        rows = [ln.split() for ln in fp.read().strip().split("\n")[1:]]
    assert [(int(r[3]), int(r[5])) for r in rows] == [(cat[97], 97), (cat[150], 150)]


def test_query_nomatch(dory_index):
    table, sizes, catlas, k = dory_index
    kmers = O.query_kmers(O.read_records(golden_path("data", "random-query-nomatch.fa")), k)
    assert len(kmers) == 8
    cdbg, cat, doms, shadow = O.query_execute(kmers, table, sizes, catlas)
    assert cdbg == {} and cat == {} and doms == set() and shadow == set()


# ---- index_reads labelling ----------------------------------------------------
def test_label_reads_dory_subset(dory_index):
    import hashlib

    table, sizes, catlas, k = dory_index
    recs = O.read_records(golden_path("data", "dory-subset.fq.gz"))
    assert len(recs) == 537
    rows, n_paired, total = O.label_reads(((r, i) for i, r in enumerate(recs)), table, k)
    assert total == set(range(736))  # tests/test_dory_workflow.py:514-536 (rc == 0)
    assert len(rows) == 907
    lines = ["%d,%d\n" % r for r in sorted(rows)]  # numeric order
    # SURVEY 8c derived digest
    assert hashlib.sha256("".join(lines).encode()).hexdigest() == "933dbab55d9262d91ce6f3fe4c883957480e4f22e6d9ff503a44ee24d3ea7c90"
    # C label_reads agrees
    ct = O.COracleTable(table.mphf_to_hash, table.mphf_to_value)
    buf, off = O.pack_sequences([r.sequence for r in recs])
    ids_off, ids, nk, nh = ct.label_reads(buf, off, k)
    assert nk == 240920 and nh == 240920  # every dory-subset k-mer hits
    crow = sorted((i, int(c)) for i in range(len(recs)) for c in ids[ids_off[i] : ids_off[i + 1]])
    assert crow == sorted(set(rows))
    nk2, nh2, nl2, cs = ct.label_reads_checksum(buf, off, k)
    assert (nk2, nh2, nl2) == (nk, nh, len(ids))
    assert cs == O.labels_checksum(ids_off, ids)


def test_label_reads_pairing_semantics(dory_index):
    """tests/test_twofoo_short.py:86-172: /1,/2 mates share the union of ids; -N disables it;
    records shorter than k never become `last_record` (index_reads.py:143-144,165-167)."""
    table, sizes, catlas, k = dory_index
    u = O.read_unitigs_tsv(golden_path("dory_k21_unitigs.tsv.gz"))
    s118, s187 = u[118].sequence[:60], u[187].sequence[:60]
    R = O.Record
    recs = [R("x/1", s118), R("x/2", s187), R("y/1", s118), R("z/2", s187), R("bare", s118), R("bare", s187), R("w/1", s118), R("short", "ACGT"), R("w/2", s187)]
    rows, n_paired, _ = O.label_reads(((r, 100 + i) for i, r in enumerate(recs)), table, k)
    by = {}
    for o, c in rows:
        by.setdefault(o, set()).add(c)
    assert n_paired == 3
    assert by[100] == {118, 187} and by[101] == {118, 187}  # x/1 + x/2
    assert by[102] == {118} and by[103] == {187}  # y/1 then z/2: not mates
    assert by[104] == {118, 187} and by[105] == {118, 187}  # equal bare names
    assert by[106] == {118, 187} and by[108] == {187} and 107 not in by  # short read skipped, w/1+w/2 still mates
    rows_n, n_paired_n, _ = O.label_reads(((r, 100 + i) for i, r in enumerate(recs)), table, k, ignore_paired=True)
    assert n_paired_n == 0
    byn = {}
    for o, c in rows_n:
        byn.setdefault(o, set()).add(c)
    assert byn[100] == {118} and byn[101] == {187}


def test_dominator_abundance_totals(dory_index):
    table, sizes, catlas, k = dory_index
    recs = O.read_records(golden_path("data", "dory-subset.fa.gz"))
    cdbg_counts, dom_counts = O.dominator_abundance(recs, table, k, catlas)
    # tests/test_dory_workflow.py:859-860, 869-870
    assert sum(cdbg_counts.values()) == 240920 and None not in cdbg_counts
    assert sum(dom_counts.values()) == 240920


def test_sourmash_min_hashval_all_unitigs():
    """Every unitig's ``min_hashval`` written by the reference (cdbg/sort_bcalm_unitigs.py:56-128,
    sourmash MinHash n=1 seed 42) equals the minimum of the restated sourmash k-mer hash: pins
    oracle.minhash_kmer_py (the checker of the GPU MinHash mode) on 736 reference values."""
    with open(golden_path("dory_k21_min_hashval.json")) as fp:
        want = {int(a): b for a, b in json.load(fp).items()}
    recs = O.read_unitigs_tsv(golden_path("dory_k21_unitigs.tsv.gz"))
    assert len(want) == 736
    for r in recs[::7] + recs[-3:]:
        assert min(O.minhash_sequence_py(r.sequence, 21)) == want[int(r.name)], r.name


# ---- BBHash (boomphf) restatement, pinned on the reference's own contigs.mphf + contigs.indices ----
@pytest.fixture(scope="module")
def dory_mphf():
    from oracle import bbhash

    raw = open(golden_path("dory_k21_contigs.mphf"), "rb").read()
    with np.load(golden_path("dory_k21_contigs.indices.npz")) as z:
        keys, vals = z["mphf_to_hash"].astype(np.uint64), z["mphf_to_value"]
    return bbhash, raw, keys, vals


def test_bbhash_lookup_reproduces_reference_order(dory_mphf):
    # mphf_to_hash[i] is the key whose MPHF index is i (BBHashTable.__setitem__): looking every key up
    # through the level bitsets of the reference's dump must give 0..n-1 in order.
    bbhash, raw, keys, _ = dory_mphf
    m = bbhash.Mphf.loads(raw)
    assert (m.gamma, m.nb_levels, m.nelem, m.lastbitsetrank) == (1.0, 25, 212475, 212475)
    assert np.array_equal(m.lookup(keys), np.arange(len(keys), dtype=np.uint64))
    assert m.dumps() == raw


def test_bbhash_build_is_byte_identical_to_reference_dump(dory_mphf):
    bbhash, raw, keys, _ = dory_mphf
    shuffled = np.random.default_rng(7).permutation(keys)  # construction is order independent
    assert bbhash.Mphf.build(shuffled, gamma=1.0, nb_levels=25).dumps() == raw


def test_bbhash_small_sets_and_final_map():
    from oracle import bbhash

    rng = np.random.default_rng(3)
    for n, levels in ((0, 25), (1, 25), (2, 25), (65, 25), (5000, 4), (5000, 2)):
        keys = np.unique(rng.integers(0, 2**63, size=n, dtype=np.uint64))
        m = bbhash.Mphf.build(keys, nb_levels=levels)
        idx = m.lookup(keys)
        assert sorted(idx.tolist()) == list(range(len(keys)))
        assert (len(m.final) > 0) == (levels < 25 and n == 5000)
        again = bbhash.Mphf.loads(m.dumps())
        assert np.array_equal(again.lookup(keys), idx)
