"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/sgc_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np

from bgzf_writer import write_bgzf
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
    write_bgzf(p, fq, block_size=5000)
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


# ---- native host feeder (sgc_bgzf_* / sgc_reads_*): no GPU needed -----------------------------
def _messy_fastq(rng, n):
    out = []
    for i in range(n):
        ln = int(rng.integers(0, 300))
        s = "".join("ACGTN"[j] for j in rng.integers(0, 5, ln))
        eol = "\r\n" if rng.random() < 0.3 else "\n"
        name = "r%d/%d" % (i // 2, i % 2 + 1) + (" extra words\t" if rng.random() < 0.2 else "")
        pad = " " if rng.random() < 0.1 else ""
        out.append("@%s%s%s%s%s+%s%s%s" % (name, eol, pad, s, pad + eol, pad, eol, "I" * ln + eol))
    return "".join(out)


def _messy_fasta(rng, n):
    out = []
    for i in range(n):
        ln = int(rng.integers(0, 900))
        s = "".join("ACGT"[j] for j in rng.integers(0, 4, ln))
        eol = "\r\n" if rng.random() < 0.3 else "\n"
        w = int(rng.integers(20, 90))
        body = "".join(("  " if rng.random() < 0.05 else "") + s[a : a + w] + ("\t" if rng.random() < 0.05 else "") + eol
                       for a in range(0, ln, w))
        if rng.random() < 0.1:
            body += eol  # blank line inside the file
        out.append(">%s seq%d %s%s%s" % (" " if rng.random() < 0.1 else "", i, "desc " * int(rng.integers(0, 3)), eol, body))
    return "".join(out)


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_native_record_scan_vs_reference_iterators(fmt, threads, monkeypatch):
    """sgc_reads_scan against the oracle's restatement of my_fasta_iter / my_fastq_iter: names,
    sequences and header positions, over CRLF, padded lines, blank lines, descriptions, empty
    sequences and a last line without newline; buffers large enough that every thread owns chunks."""
    from oracle import oracle as O
    from spacegraphcats_b200.utils import bgzf

    monkeypatch.setenv("SGC_FEED_THREADS", threads)
    rng = np.random.default_rng(len(fmt) + int(threads))
    for trial in range(3):
        text = (_messy_fastq if fmt == "fastq" else _messy_fasta)(rng, 6000 if trial == 0 else int(rng.integers(1, 40)))
        if trial == 1:
            text = text.rstrip("\r\n")  # no trailing newline
        data = text.encode()
        want = list(O.iterate_records(data))
        names, seq, off, pos = bgzf.parse_records(np.frombuffer(data, np.uint8))
        assert len(names) == len(want)
        assert names == [r.name for r, _ in want]
        assert pos.tolist() == [p for _, p in want]
        raw = seq.tobytes().decode()
        assert [raw[off[i] : off[i + 1]] for i in range(len(want))] == [r.sequence for r, _ in want]


def test_native_record_scan_error_paths():
    from spacegraphcats_b200.utils import bgzf

    with pytest.raises(Exception, match="unknown start chr"):
        bgzf.parse_records(np.frombuffer(b"ACGT\n", np.uint8))
    with pytest.raises(ValueError, match="4 lines"):
        bgzf.parse_records(np.frombuffer(b"@a\nACGT\n+\nIIII\n@b\nACGT\nIIII\n@c\n", np.uint8))
    n, s, o, p = bgzf.parse_records(np.zeros(0, np.uint8))
    assert len(n) == 0 and o.tolist() == [0] and len(p) == 0
    # an incomplete trailing record is ignored
    n, s, o, p = bgzf.parse_records(np.frombuffer(b"@a\nACGT\n+\nIIII\n@b\nAC", np.uint8))
    assert n == ["a"] and s.tobytes() == b"ACGT"


def test_native_bgzf_inflate_and_corruption(tmp_path):
    import gzip
    from spacegraphcats_b200._lib import SgcError
    from spacegraphcats_b200.utils import bgzf

    rng = np.random.default_rng(9)
    payload = bytes(rng.integers(65, 91, 1_500_000, dtype=np.uint8))
    p = str(tmp_path / "x.bgz")
    write_bgzf(p, payload, block_size=40000)
    src = bgzf.BgzfData(p)
    assert src.data.tobytes() == payload == gzip.open(p, "rb").read()
    pos = np.array([0, 39999, 40000, 40001, len(payload) - 1], np.int64)
    vo = src.virtual_offsets(pos)
    assert (vo & 0xFFFF).tolist() == [0, 39999, 0, 1, (len(payload) - 1) % 40000]
    assert (vo >> 16)[1] == 0 and (vo >> 16)[2] > 0
    raw = bytearray(open(p, "rb").read())
    raw[len(raw) // 2] ^= 0x55  # flip bits inside some block's deflate stream
    bad = str(tmp_path / "bad.bgz")
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(SgcError):
        bgzf.BgzfData(bad)
    open(bad, "wb").write(bytes(raw[: len(raw) // 3]))  # truncated
    with pytest.raises(SgcError):
        bgzf.BgzfData(bad)


def test_native_pair_flags_vs_python_rule():
    from spacegraphcats_b200.cdbg.index_reads import pair_flags

    rng = np.random.default_rng(4)
    pool = ["a/1", "a/2", "b/1", "b/2", "c/2", "same", "", "/1", "/2", "x", "ab/1", "b/2", "a/12", "/2/2"]
    names = [pool[i] for i in rng.integers(0, len(pool), 5000)]
    elig = rng.random(5000) < 0.8
    got_p, got_l = pair_flags(names, elig)
    last = None
    for i, nm in enumerate(names):  # index_reads.py:124-141, literally
        is_paired = False
        if last is not None:
            ln = names[last]
            if ln.endswith("/1") and nm.endswith("/2"):
                if ln[:-2] == nm[:-2] and len(ln) > 2:
                    is_paired = True
            elif ln == nm and ln:
                is_paired = True
        assert bool(got_p[i]) == is_paired and got_l[i] == (-1 if last is None else last), i
        if elig[i]:
            last = i
    p2, l2 = pair_flags(names, elig, ignore_paired=True)
    assert not p2.any() and (l2 == -1).all()


def test_bgzf_reader_on_python_written_blocks(tmp_path):
    import gzip
    from spacegraphcats_b200.utils import bgzf

    p = str(tmp_path / "w.bgz")
    write_bgzf(p, b"")
    assert open(p, "rb").read() == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    rng = np.random.default_rng(2)
    noise = bytes(rng.integers(0, 256, 300_000, dtype=np.uint8))  # incompressible: blocks grow, must still fit 64 KiB
    for level in (0, 1, 6, 9):
        write_bgzf(p, noise, level=level)
        assert bgzf.is_bgzf(p) and gzip.open(p, "rb").read() == noise
        assert bgzf.BgzfData(p).data.tobytes() == noise


def test_host_packer_layout_and_bad_base_flags():
    """sgc_pack_reads: 2 bits per base, A=0 C=1 G=2 T=3, records padded to 4 bytes; non-ACGT flagged."""
    from spacegraphcats_b200.device import pack_reads_host, pack_sequences

    rng = np.random.default_rng(4)
    seqs = ["", "A", "ACGT" * 16, "ACGT" * 16 + "T", "ACGNT", "acgt"] + [
        "".join("ACGT"[c] for c in rng.integers(0, 4, int(n))) for n in rng.integers(0, 400, 200)
    ]
    buf, off = pack_sequences(seqs)
    for threads in (1, 3):
        pk, ln, bad = pack_reads_host(buf, off, n_threads=threads)
        assert ln.tolist() == [len(s) for s in seqs]
        assert bad.tolist() == [any(c not in "ACGT" for c in s) for s in seqs]
        pos = 0
        for s in seqs:
            nb = (len(s) + 15) // 16 * 4
            rec = pk[pos:pos + nb]
            pos += nb
            dec = "".join("ACGT"[(int(rec[b // 4]) >> (2 * (b % 4))) & 3] for b in range(len(s)))
            assert dec == "".join(c if c in "ACGT" else "A" for c in s)
            assert not rec[(len(s) + 3) // 4:].any()  # padding is zero
        assert pos == len(pk)
    from spacegraphcats_b200._lib import lib

    L = lib()
    assert L.sgc_packed_bytes(None, 150, 10) == 10 * 40
    lens = np.array([0, 1, 64, 65], np.int32)
    assert L.sgc_packed_bytes(lens.ctypes.data_as(ctypes.c_void_p), -1, 4) == (0 + 1 + 4 + 5) * 4


def test_counts8_to_offsets_host():
    """sgc_counts8_to_offsets (host, threaded): the CSR offsets of a labelled batch rebuilt from one byte per read."""
    from spacegraphcats_b200._lib import lib

    L = lib()
    rng = np.random.default_rng(8)
    for n in (0, 1, 7, 1000, 3_000_001):
        cnt = rng.integers(0, 256, n + 1).astype(np.uint8)  # (the byte after the last count is the overflow flag)
        for threads in (1, 5, 0):
            off = np.full(n + 1, -1, np.int64)
            rc = L.sgc_counts8_to_offsets(cnt.ctypes.data_as(ctypes.c_void_p), n, threads, off.ctypes.data_as(ctypes.c_void_p))
            assert rc == 0
            assert np.array_equal(off, np.concatenate([[0], np.cumsum(cnt[:n].astype(np.int64))]))
    assert L.sgc_counts8_to_offsets(None, 5, 1, None) != 0
