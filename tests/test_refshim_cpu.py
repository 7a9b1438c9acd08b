"""The reference's OWN Python (spacegraphcats/cdbg/{hashing,index_cdbg_by_kmer,index_reads}.py and
search/{query_by_sequence,catlas,search_utils}.py, imported UNMODIFIED from /root/reference) running on the test
shims of its third-party natives -- ``khmer``, ``bbhash_table``, ``screed``, ``sourmash`` (tests/refshim/) -- backed
here by the CPU oracle.  This pins, in a container without a GPU, that

  * the shim surface is exactly what the reference calls (the same shim modules, switched to the raw ctypes
    binding of libsgc_b200.so, are exercised on the GPU by tests/test_gpu_parity.py::test_reference_side_binding_*),
  * the oracle reproduces the reference's workflow-level pins when driven by the reference's own loops:
    tests/test_dory_workflow.py:190-199 (build_mphf + the four KATs) and :514-536 (index_reads rc == 0).

Skipped where /root/reference does not exist (the GPU box)."""
import os
import sqlite3
import sys
import types

import numpy as np
import pytest

from bgzf_writer import write_bgzf
from conftest import ROOT, golden_path

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "spacegraphcats")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    """import the reference package without its version check (spacegraphcats/__init__.py:1-15 needs an installed
    dist): a bare package object whose __path__ is the reference tree, so every submodule is the reference's file"""
    shim = os.path.join(ROOT, "tests", "refshim")
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items()}
    os.environ["SGC_REFSHIM_BACKEND"] = "oracle"
    sys.path.insert(0, shim)
    for name in ("khmer", "bbhash_table", "screed", "sourmash", "_backend"):
        sys.modules.pop(name, None)
    pkg = types.ModuleType("spacegraphcats")
    pkg.__path__ = [os.path.join(REF, "spacegraphcats")]
    sys.modules["spacegraphcats"] = pkg
    import spacegraphcats.cdbg.hashing as hashing
    import spacegraphcats.cdbg.index_cdbg_by_kmer as index_cdbg_by_kmer
    import spacegraphcats.cdbg.index_reads as index_reads

    assert hashing.__file__.startswith(REF) and index_cdbg_by_kmer.__file__.startswith(REF)
    yield types.SimpleNamespace(hashing=hashing, index_cdbg_by_kmer=index_cdbg_by_kmer, index_reads=index_reads)
    sys.path[:] = saved_path
    for k in list(sys.modules):
        if k not in saved_mods:
            del sys.modules[k]


@pytest.fixture(scope="module")
def dory_records():
    from oracle import oracle as O

    return O.read_unitigs_tsv(golden_path("dory_k21_unitigs.tsv.gz"))


def test_reference_hash_sequence_kats(ref):
    hs = ref.hashing.hash_sequence
    assert hs("GACCAAAGGCTTACGGTGAGC", 21) == [10218271035842461694]
    assert hs("AGACCAACCTTTAGCCTATTCAAGACGAGTG", 31) == [505484382184005004]
    assert hs("ACGT" * 8, 32) == [0]
    with pytest.raises(ValueError):
        hs("ACGT", 21)


def test_reference_build_mphf_on_dory(ref, dory_records, tmp_path):
    """tests/test_dory_workflow.py:190-199 through the reference's own build_mphf + MPHF_KmerIndex"""
    table, sizes = ref.index_cdbg_by_kmer.build_mphf(21, lambda: iter(dory_records))
    kmer_idx = ref.hashing.MPHF_KmerIndex(21, table, sizes)
    assert kmer_idx.table[10218271035842461694] == 118
    assert kmer_idx.table[8436068710919520258] == 118
    assert kmer_idx.table[13994045974119358468] == 118
    assert kmer_idx.table[11971930231572094512] == 187
    assert kmer_idx.get_cdbg_id(1) is None and len(kmer_idx) == 212475
    z = np.load(golden_path("dory_k21_contigs.indices.npz"))
    got = dict(zip(table.mphf_to_hash.tolist(), table.mphf_to_value.tolist()))
    assert got == dict(zip(z["mphf_to_hash"].tolist(), z["mphf_to_value"].tolist()))
    assert sum(sizes.values()) == 212475 and len(sizes) == 736
    # count_cdbg_matches (hashing.py:67-69): list counts duplicates, set does not
    q = ref.hashing.hash_sequence(dory_records[118].sequence[:60], 21)
    assert kmer_idx.count_cdbg_matches(q + q) == {118: 2 * len(q)}
    assert kmer_idx.count_cdbg_matches(set(q + q)) == {118: len(q)}
    # save / from_directory round trip through the shim's on-disk arrays (hashing.py:109-130)
    kmer_idx.save(str(tmp_path))
    again = ref.hashing.MPHF_KmerIndex.from_directory(str(tmp_path))
    assert again.ksize == 21 and len(again) == 212475 and again.table[11971930231572094512] == 187


def test_reference_index_reads_main_on_dory(ref, dory_records, tmp_path):
    """tests/test_dory_workflow.py:514-536: the reference's index_reads.main (its BgzfReader, its iterators, its
    pairing loop) over dory-subset.fq on the shims: rc == 0 <=> every unitig id 0..735 received reads; rows equal
    the oracle's restatement of the loop."""
    from oracle import oracle as O

    z = np.load(golden_path("dory_k21_contigs.indices.npz"))
    d = tmp_path / "dory_k21"
    d.mkdir()
    table, sizes = ref.index_cdbg_by_kmer.build_mphf(21, lambda: iter(dory_records))
    ref.hashing.MPHF_KmerIndex(21, table, sizes).save(str(d))
    reads = O.read_records(golden_path("data", "dory-subset.fq.gz"))
    fq = "".join("@%s\n%s\n+\n%s\n" % (r.name, r.sequence, r.quality) for r in reads).encode()
    rp = str(tmp_path / "reads.bgz")
    write_bgzf(rp, fq)
    db = str(tmp_path / "reads.index")
    assert ref.index_reads.main([str(d), rp, db, "-k", "21"]) == 0
    rows = sorted(sqlite3.connect(db).execute("SELECT offset, cdbg_id FROM sequences").fetchall())
    assert len({c for _, c in rows}) == 736
    # the oracle's loop over the same (record, virtual offset) stream
    from spacegraphcats_b200.utils import bgzf

    src = bgzf.open_reads(rp)
    _, _, _, pos = bgzf.parse_records(src.data)
    offs = src.virtual_offsets(pos).tolist()
    want, _, total = O.label_reads(zip(reads, offs), O.OracleTable.from_arrays(z["mphf_to_hash"], z["mphf_to_value"]), 21)
    assert rows == sorted(want) and len(total) == 736


def test_reference_query_on_dory_head(ref, dory_records, tmp_path):
    """The reference's own ``Query`` / ``QueryOutput`` (search/query_by_sequence.py:27-267) and ``CAtlas``
    (search/catlas.py) on the shims: dory-head against the dory cDBG reproduces tests/test_dory_workflow.py:217-228
    and the reference's result files -- match counts {118: 834, 187: 797}, frontier {97, 150}, the results.csv row
    (containment, similarity, bp, contigs, k, query k-mers, upper bounds), cdbg ids, response curve -- and equals
    the oracle's restatement of Query.execute."""
    import csv
    import gzip
    import io

    from oracle import oracle as O

    import spacegraphcats.search.query_by_sequence as qbs
    from spacegraphcats.search.catlas import CAtlas

    assert qbs.__file__.startswith(REF)
    table, sizes = ref.index_cdbg_by_kmer.build_mphf(21, lambda: iter(dory_records))
    kmer_idx = ref.hashing.MPHF_KmerIndex(21, table, sizes)
    catlas = CAtlas(str(tmp_path), os.path.dirname(golden_path("dory_k21_r1", "catlas.csv")), load_sizefile=None)
    catlas.decorate_with_index_sizes(kmer_idx)
    qfile = golden_path("data", "dory-head.fa")
    q = qbs.Query(qfile, 21, 1000, "dory_k21_r1")
    out = q.execute(catlas, kmer_idx)
    assert len(q.kmers) == 1631
    assert q.cdbg_match_counts == {118: 834, 187: 797}
    assert out.doms == {97, 150} and out.shadow == {118, 187}
    # the oracle's restatement of the same steps
    ocat = O.OracleCAtlas(golden_path("dory_k21_r1", "catlas.csv"), golden_path("dory_k21_r1", "first_doms.txt"))
    z = np.load(golden_path("dory_k21_contigs.indices.npz"))
    ocdbg, ocat_counts, odoms, oshadow = O.query_execute(O.query_kmers(O.read_records(qfile), 21),
                                                         O.OracleTable.from_arrays(z["mphf_to_hash"], z["mphf_to_value"]),
                                                         sizes, ocat)
    assert (q.cdbg_match_counts, q.catlas_match_counts, out.doms, out.shadow) == (ocdbg, ocat_counts, odoms, oshadow)
    # QueryOutput: contigs from a sqlite store like sort_bcalm_unitigs writes, then the result files
    db = sqlite3.connect(str(tmp_path / "bcalm.unitigs.db"))
    db.execute("CREATE TABLE sequences (id INTEGER, sequence TEXT, offset INTEGER PRIMARY KEY, min_hashval TEXT, abund FLOAT)")
    db.executemany("INSERT INTO sequences (id, sequence, offset) VALUES (?, ?, ?)",
                   [(int(r.name), r.sequence, i) for i, r in enumerate(dory_records)])
    db.commit()
    out.retrieve_contigs(db)
    buf = io.StringIO()
    outdir = tmp_path / "search"
    outdir.mkdir()
    out.write(csv.writer(buf), buf, str(outdir), "dory_k21_r1")
    got_row = buf.getvalue().strip().split(",")
    with open(golden_path("dory_k21_r1_search", "results.csv")) as fp:
        want_row = [ln.strip().split(",") for ln in fp][1]
    assert got_row[1:] == want_row[1:]  # (column 0 is the query's path on the reference authors' machine)
    with gzip.open(golden_path("dory_k21_r1_search", "dory-head.fa.cdbg_ids.txt.gz"), "rt") as a, \
            gzip.open(str(outdir / "dory-head.fa.cdbg_ids.txt.gz"), "rt") as b:
        assert a.read() == b.read()
    with gzip.open(golden_path("dory_k21_r1_search", "dory-head.fa.frontier.txt.gz"), "rt") as a, \
            gzip.open(str(outdir / "dory-head.fa.frontier.txt.gz"), "rt") as b:
        assert sorted(a.read().split()) == sorted(b.read().split())
    with open(golden_path("dory_k21_r1_search", "dory-head.fa.response.txt")) as a, \
            open(str(outdir / "dory-head.fa.response.txt")) as b:
        assert a.read() == b.read()


def test_reference_sort_bcalm_unitigs_on_dory(ref, dory_records, tmp_path):
    """The reference's own ``sort_bcalm_unitigs.main`` (cdbg/sort_bcalm_unitigs.py:19-196) on raw BCALM output
    (data/bcalm.dory.k21.unitigs.fa) with the sourmash shim: the ids it assigns (rank in the TEXT order of
    min_hashval, :101-118) and the min_hashvals are those of the reference's dory fixture -- which pins the shim's
    hash (Murmur3 seed 42 of the canonical k-mer) and the id rule spacegraphcats_b200.cdbg.minhash.assign_unitig_ids
    follows, starting from BCALM's own file."""
    import json
    import shutil

    from oracle import oracle as O

    import spacegraphcats.cdbg.sort_bcalm_unitigs as sbu

    assert sbu.__file__.startswith(REF)
    fa = str(tmp_path / "bcalm.dory.k21.unitigs.fa")  # (main writes a .sig next to its input)
    shutil.copy(os.path.join(REF, "data", "bcalm.dory.k21.unitigs.fa"), fa)
    db_out, pkl = str(tmp_path / "bcalm.unitigs.db"), str(tmp_path / "bcalm.unitigs.pickle")
    assert sbu.main([fa, db_out, pkl, "-k", "21"]) == 0
    rows = sqlite3.connect(db_out).execute("SELECT id, sequence, min_hashval FROM sequences ORDER BY id").fetchall()
    assert len(rows) == len(dory_records) == 736
    with open(golden_path("dory_k21_min_hashval.json")) as fp:
        want_min = {int(a): b for a, b in json.load(fp).items()}
    canon = lambda s: min(s, O.revcomp(s.encode()).decode())  # noqa: E731  (BCALM may write a unitig on either strand)
    by_id = {int(r.name): r.sequence for r in dory_records}
    for cdbg_id, seq, mh in rows:
        assert int(mh) == want_min[cdbg_id]
        assert canon(seq) == canon(by_id[cdbg_id])
    assert [r[0] for r in sorted(rows, key=lambda r: r[2])] == list(range(736))  # TEXT order of min_hashval = id order


def test_reference_count_dominator_abundance_on_dory(ref, tmp_path):
    """The reference's own ``count_dominator_abundance.main`` (search/count_dominator_abundance.py:24-170) on the shims,
    over dory-subset.fa: the pins of tests/test_dory_workflow.py:849-870 (737 + 2836 CSV lines, totals 240920 / 212475,
    the top dominator row) and, row for row, the committed golden vectors the GPU test compares with
    (tests/golden/dory_subset_abund.json, written by tests/golden/make_ref_outputs.py from this very run)."""
    import json

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    try:
        import make_ref_outputs
    finally:
        sys.path.pop(0)
    res = make_ref_outputs.run_abundance(str(tmp_path))
    assert len(res["cdbg"]) + 1 == 737 and len(res["dom"]) + 1 == 2836
    assert sum(r[1] for r in res["cdbg"]) == 240920 and sum(r[2] for r in res["cdbg"]) == 212475
    top = [r for r in res["dom"] if r[1] == 7]
    assert top and all(r[2] == 240920 and r[3] == 212475 for r in top)
    with open(golden_path("dory_subset_abund.json")) as fp:
        assert json.load(fp) == res


@pytest.mark.parametrize("flags", [[], ["-N"], ["-P"]])
def test_reference_index_reads_pairing_vs_oracle_loop(ref, tmp_path, flags):
    """The reference's own index_reads.main on mate pairs (its pairing fixture tests/test-data/twofoo-short-paired-test
    plus corner cases: /1,/2 names, equal names, a short read between mates, a short mate, names of length 2, a read
    that hits nothing) in its three modes: the rows equal the oracle's restatement of the loop (oracle.label_reads,
    which the GPU tests compare the CUDA path with), so that restatement is pinned on the reference's code itself."""
    from oracle import oracle as O
    from spacegraphcats_b200.utils import bgzf

    k = 31
    fixture = O.read_records(golden_path("data", "twofoo-short-paired-test.fa.gz"))
    rng = np.random.default_rng(4)
    genome = "".join(rng.choice(list("ACGT"), 6000))
    extra = [
        O.Record("p1/1", genome[0:150]), O.Record("p1/2", genome[300:450]),
        O.Record("p2/1", genome[600:750]), O.Record("p3/2", genome[900:1050]),
        O.Record("same", genome[1200:1350]), O.Record("same", genome[1500:1650]),
        O.Record("q/1", genome[2000:2150]), O.Record("short", "ACGT"), O.Record("q/2", genome[2300:2450]),
        O.Record("r/1", genome[2600:2750]), O.Record("r/2", "ACGTACGT"),
        O.Record("/1", genome[3000:3150]), O.Record("/2", genome[3300:3450]),
        O.Record("nohit", "".join(rng.choice(list("ACGT"), 150))),
    ]
    reads = list(fixture) + extra
    pieces = []
    for r in list(fixture) + [O.Record("g", genome)]:
        for a in range(0, max(len(r.sequence) - k + 1, 0), 100):
            pieces.append(r.sequence[a : a + 100 + k - 1])
    urecs = [O.Record(str(i), s) for i, s in enumerate(pieces)]
    table, sizes = ref.index_cdbg_by_kmer.build_mphf(k, lambda: iter(urecs))
    d = tmp_path / "cdbg"
    d.mkdir()
    ref.hashing.MPHF_KmerIndex(k, table, sizes).save(str(d))
    rp = str(tmp_path / "reads.bgz")
    write_bgzf(rp, "".join(">%s\n%s\n" % (r.name, r.sequence) for r in reads).encode(), block_size=700)
    db = str(tmp_path / "out.index")
    rc = ref.index_reads.main([str(d), rp, db, "-k", str(k)] + flags)
    rows = sorted(sqlite3.connect(db).execute("SELECT offset, cdbg_id FROM sequences").fetchall())
    src = bgzf.open_reads(rp)
    names, _, _, pos = bgzf.parse_records(src.data)
    assert names == [r.name for r in reads]
    offs = src.virtual_offsets(pos).tolist()
    otab, _, _ = O.build_mphf(k, lambda: iter(urecs))
    want, n_paired, total = O.label_reads(zip(reads, offs), otab, k, ignore_paired="-N" in flags)
    assert rows == sorted(want)
    assert (n_paired > 0) == ("-N" not in flags)
    ids_complete = max(total) + 1 == len(total)
    assert rc == (0 if ids_complete and not ("-P" in flags and not n_paired) else -1)
