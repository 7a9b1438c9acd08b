"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/sgc_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    from spacegraphcats_b200 import _lib

    return _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sgc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(built):
    L = built.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libsgc_b200.so does not export %s" % n
        assert n in built.PROTOTYPES, "no ctypes prototype for %s" % n
    assert L.sgc_abi_version() == 1


def test_no_cpu_fallback_without_gpu(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = built.lib()
    h = ctypes.c_void_p()
    rc = L.sgc_ctx_create(0, None, ctypes.byref(h))
    assert rc == built.SGC_ERR_CUDA and not h.value
    assert b"no CPU fallback" in L.sgc_last_error()
    from spacegraphcats_b200.cdbg.hashing import hash_sequence

    with pytest.raises(RuntimeError):
        hash_sequence("ACGTACGTACGTACGTACGTACGTA", 21)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under spacegraphcats_b200/ may reach it."""
    pkg = os.path.join(ROOT, "spacegraphcats_b200")
    pat = re.compile(r"(^|\n)\s*(from|import)\s+oracle|libsgc_oracle|oracle\.oracle|sgc_oracle_")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dp, fn)).read()), fn


def test_host_emulation_of_tile_geometry_matches_oracle():
    """The staging / window-alignment arithmetic of the tile kernel (sgc_kmer.cuh), compiled
    for the host and driven the way one GPU thread drives it, against the oracle."""
    import subprocess

    from oracle import oracle as O

    src = os.path.join(ROOT, "tests", "host", "emul_tile.cpp")
    so = os.path.join(ROOT, "tests", "host", "libemul_tile.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    L = ctypes.CDLL(so)
    rng = np.random.default_rng(3)
    for k in (21, 31, 32, 5, 1, 7, 8, 9, 16, 17, 33, 47, 48, 64, 100, 255):
        for ln in (k, k + 1, k + 2, k + 3, k + 4, k + 5, 150, 300, 700):
            if ln < k:
                continue
            for seg_k in (256, 8, 5):
                seq = rng.choice(list(b"ACGT"), ln).astype(np.uint8)
                if ln == 300:
                    seq[17], seq[100] = ord("N"), ord("a")
                out = np.zeros(ln - k + 1, np.uint64)
                inv = ctypes.c_int(0)
                rc = L.emul_hash_sequence(seq.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(ln), k, seg_k,
                                          out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(inv))
                assert rc == 0
                assert np.array_equal(out, O.hash_sequence(seq.tobytes(), k)), (k, ln, seg_k)
                assert inv.value == (1 if ln == 300 else 0)


def test_host_emulation_of_minhash_geometry_matches_oracle():
    """sourmash-style hash (lexicographic min of k-mer / reverse complement, seeded Murmur3) through
    the tile kernel's staged windows, on the host."""
    from oracle import oracle as O

    L = ctypes.CDLL(os.path.join(ROOT, "tests", "host", "libemul_tile.so"))
    rng = np.random.default_rng(5)
    for k in (21, 31, 5, 8, 9, 16, 33, 64, 255):
        for ln in (k, k + 1, k + 5, 150, 600):
            if ln < k:
                continue
            seq = rng.choice(list(b"ACGT"), ln).astype(np.uint8)
            if ln == 150:
                seq[:] = np.frombuffer((b"ACGT" * 40)[:ln], np.uint8)  # many k-mer == revcomp ties
            out = np.zeros(ln - k + 1, np.uint64)
            rc = L.emul_minhash_sequence(seq.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(ln), k, 7,
                                         ctypes.c_uint64(42), out.ctypes.data_as(ctypes.c_void_p))
            assert rc == 0
            assert out.tolist() == O.minhash_sequence_py(seq.tobytes(), k), (k, ln)


def test_bgzf_roundtrip_and_record_offsets(tmp_path):
    from spacegraphcats_b200.utils import bgzf

    rng = np.random.default_rng(5)
    recs = []
    for i in range(300):
        s = "".join("ACGT"[j] for j in rng.integers(0, 4, int(rng.integers(30, 400))))
        recs.append(("read%d/%d" % (i // 2, i % 2 + 1), s))
    fq = "".join("@%s\n%s\n+\n%s\n" % (n, s, "I" * len(s)) for n, s in recs).encode()
    p = str(tmp_path / "r.fq.bgz")
    bgzf.write_bgzf(p, fq, block_size=5000)
    import gzip

    assert gzip.open(p, "rb").read() == fq  # BGZF is valid multi-member gzip
    src = bgzf.open_reads(p)
    assert isinstance(src, bgzf.BgzfData)
    names, seq, off, pos = bgzf.parse_records(src.data)
    assert names == [n for n, _ in recs]
    assert [seq[off[i] : off[i + 1]].tobytes().decode() for i in range(len(recs))] == [s for _, s in recs]
    vo = src.virtual_offsets(pos)
    # a virtual offset must address the record's '@' inside its block
    raw = open(p, "rb").read()
    import zlib, struct

    for v, (n, _) in list(zip(vo.tolist(), recs))[::37]:
        cs, within = v >> 16, v & 0xFFFF
        xlen = struct.unpack_from("<H", raw, cs + 10)[0]
        bsize = struct.unpack_from("<H", raw, cs + 16)[0] + 1
        block = zlib.decompress(raw[cs + 12 + xlen : cs + bsize - 8], -15)
        assert block[within : within + 1 + len(n)] == b"@" + n.encode()
    # FASTA, multi-line
    fa = "".join(">%s\n%s\n" % (n, "\n".join(s[i : i + 60] for i in range(0, len(s), 60))) for n, s in recs).encode()
    names2, seq2, off2, pos2 = bgzf.parse_records(np.frombuffer(fa, np.uint8))
    assert names2 == names and np.array_equal(seq2, seq) and np.array_equal(off2, off)
    assert all(fa[p : p + 1] == b">" for p in pos2.tolist())


def test_index_reads_host_rules_vectorised_vs_oracle_loop():
    """rows_from_labels (the vectorised pairing / carry-over rules of index_reads.py:105-167) against
    the oracle's literal restatement of the reference loop, on names that exercise every branch:
    /1-/2 mates, equal names, chains of equal names, short reads between mates, empty names."""
    from oracle import oracle as O
    from spacegraphcats_b200.cdbg.index_reads import rows_from_labels

    rng = np.random.default_rng(8)
    k = 21
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, 30000))
    pieces = [genome[a : a + 60 + k - 1] for a in range(0, len(genome) - k + 1, 60)]
    table, _, _ = O.build_mphf(k, lambda: iter(O.Record(str(i), s) for i, s in enumerate(pieces)))
    ctab = O.COracleTable(table.mphf_to_hash, table.mphf_to_value)
    names_pool = ["a/1", "a/2", "b/1", "c/2", "same", "same", "same", "/1", "/2", "", "", "x", "q/1", "short", "q/2",
                  "r/1", "r/2", "r/2", "t/1", "t/1"]
    recs = []
    for rep in range(40):
        for nm in names_pool:
            ln = 5 if nm == "short" or rng.random() < 0.1 else int(rng.integers(k, 400))
            p = int(rng.integers(0, len(genome) - 400))
            recs.append(O.Record(nm if rng.random() < 0.8 else nm + str(rep), genome[p : p + ln]))
    offs = (np.arange(len(recs), dtype=np.int64) * 1000 + 17).tolist()
    buf, off = O.pack_sequences([r.sequence for r in recs])
    ioff, ids, _, _ = ctab.label_reads(buf, off, k)
    for ignore in (False, True):
        want, n_paired, total = O.label_reads(zip(recs, offs), table, k, ignore_paired=ignore)
        ro, ri, got_paired, first = rows_from_labels([r.name for r in recs], np.diff(off), offs, ioff, ids, k, ignore)
        assert sorted(zip(ro.tolist(), ri.tolist())) == sorted(want)
        assert got_paired == n_paired and (first >= 0) == (n_paired > 0)
        assert n_paired > 0 or ignore
