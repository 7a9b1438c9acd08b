"""The reference's OWN Python (spacegraphcats/cdbg/{hashing,index_cdbg_by_kmer,index_reads}.py, imported UNMODIFIED
from /root/reference) running on the test shims of its two third-party natives -- ``khmer`` and ``bbhash_table``
(tests/refshim/) -- backed here by the CPU oracle.  This pins, in a container without a GPU, that

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
