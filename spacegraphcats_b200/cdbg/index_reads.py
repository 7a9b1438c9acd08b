"""Label every read with the cDBG unitigs its k-mers hit and write the ``(offset, cdbg_id)``
sqlite table (reference: ``spacegraphcats/cdbg/index_reads.py`` ``main`` :39-201).

The per-read arithmetic -- ``hash_sequence`` + ``count_cdbg_matches`` (:147-152) -- runs as ONE
fused GPU pass per batch of reads (hash, table probe, per-read distinct ids as CSR).  The
pairing rules (:124-141), the too-short skip (:143-144), the sqlite schema (:72, :187-188) and
the exit codes are reproduced on the host in record order.
"""
import argparse
import ctypes
import os
import sqlite3
import sys
import time

import numpy as np

from .._lib import check, lib
from ..utils import bgzf
from .hashing import MPHF_KmerIndex

DEFAULT_KSIZE = 31
TIMINGS = {}  # seconds per stage of the latest main() run (scripts/index_reads_timing.py)
BATCH_READS = 1 << 22


def pair_flags(names, elig, ignore_paired=False):
    """The pairing test of index_reads.py:124-141 for every record: is it the mate of ``last_record``
    (the latest earlier record that was long enough)?  -> (paired bool[n], last int64[n]).
    ``names``: a ``bgzf.NameTable`` (byte ranges of the read file) or a list of str; the comparison
    itself is native (sgc_reads_pair_flags)."""
    if not isinstance(names, bgzf.NameTable):
        names = bgzf.NameTable.from_strings(names)
    n = len(names)
    paired = np.zeros(n, np.uint8)
    last = np.full(n, -1, np.int64)
    if n:
        data = np.ascontiguousarray(names.data, np.uint8)
        lo = np.ascontiguousarray(names.lo, np.int64)
        hi = np.ascontiguousarray(names.hi, np.int64)
        el = np.ascontiguousarray(elig, np.uint8)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        check(lib().sgc_reads_pair_flags(vp(data), vp(lo), vp(hi), vp(el), n, int(bool(ignore_paired)), vp(paired),
                                         vp(last)))
    return paired.astype(bool), last


def rows_from_labels(names, lens, offsets, ids_off, ids, ksize, ignore_paired=False):
    """The host half of the index_reads loop (index_reads.py:105-167): given every read's distinct
    unitig ids (CSR) produce the rows the reference INSERTs, in its order (native: sgc_reads_rows).

    State of the reference loop after record i: S_i = (paired_i ? S_{i-1} : {}) | (long_i ? ids_i : {}),
    i.e. the union of the ids of the records since the last record that was NOT paired; a record long
    enough emits S_i under its mate's offset (when paired) and under its own.
    -> (rows_offset, rows_id, n_paired, first_paired)"""
    n = len(names)
    elig = np.ascontiguousarray(np.asarray(lens, np.int64) >= ksize, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    ids_off = np.ascontiguousarray(ids_off, np.int64)
    ids = np.ascontiguousarray(ids, np.uint32)
    paired, last = pair_flags(names, elig, ignore_paired)
    n_paired = int(paired.sum())
    first_paired = int(np.flatnonzero(paired)[0]) if n_paired else -1
    pflag = np.ascontiguousarray(paired, np.uint8)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    L = lib()
    n_rows = ctypes.c_int64(0)
    args = (vp(elig), vp(pflag), vp(last), vp(offsets), vp(ids_off), vp(ids), n)
    check(L.sgc_reads_rows(*args, None, None, 0, ctypes.byref(n_rows)))
    rows_o = np.empty(n_rows.value, np.int64)
    rows_i = np.empty(n_rows.value, np.int64)
    if n_rows.value:
        check(L.sgc_reads_rows(*args, vp(rows_o), vp(rows_i), n_rows.value, ctypes.byref(n_rows)))
    return rows_o, rows_i, n_paired, first_paired


def label_records(names, seq, seq_off, offsets, kmer_idx, ksize, ignore_paired=False, timings=None):
    """The loop of index_reads.py:105-167 over an in-memory batch: ONE fused GPU pass for the per-read
    ids, then the vectorised host rules.

    -> (rows_offset int64[m], rows_id int64[m], n_paired, set_of_all_ids, first_paired).
    ``rows_*`` are the rows the reference INSERTs (as a multiset; duplicates included, cf.
    :156-157); ``first_paired`` is the index of the first read found paired (-1: none)."""
    t0 = time.time()
    res = kmer_idx.table.label_reads(seq, seq_off, ksize)
    if res["status"].any():
        bad = int(np.flatnonzero(res["status"])[0])
        raise ValueError("Invalid base in read %r" % names[bad])
    t1 = time.time()
    ro, ri, n_paired, first_paired = rows_from_labels(names, np.diff(seq_off), offsets, res["ids_off"], res["ids"], ksize,
                                                      ignore_paired)
    if timings is not None:
        timings["label_gpu"] = t1 - t0
        timings["pairing_rows"] = time.time() - t1
    return ro, ri, n_paired, set(np.unique(ri).tolist()), first_paired


def main(argv=sys.argv[1:]):
    p = argparse.ArgumentParser()
    p.add_argument("cdbg_prefix", help="cDBG prefix")
    p.add_argument("reads")
    p.add_argument("savename")
    p.add_argument("-k", "--ksize", default=DEFAULT_KSIZE, type=int)
    p.add_argument("-P", "--expect-paired", action="store_true")
    p.add_argument("-N", "--ignore-paired", action="store_true")
    args = p.parse_args(argv)

    if args.expect_paired and args.ignore_paired:
        print("cannot set both -P/--expect-paired and -N/--ignore-paired")
        return -1
    if args.expect_paired:
        print("We will REQUIRE that some of the reads are in pairs (-P/--expect-paired)")
    else:
        print("We will NOT require that some of the reads be in pairs (default).")
    if args.ignore_paired:
        print("Ignoring paired reads (-N/--ignore-paired)")
    else:
        print("Paired reads will be indexed together (default).")

    dbfilename = args.savename
    if os.path.exists(dbfilename):
        print(f"removing existing db '{dbfilename}'")
        os.unlink(dbfilename)
    db = sqlite3.connect(dbfilename)
    cursor = db.cursor()
    cursor.execute("CREATE TABLE sequences (offset INTEGER, cdbg_id INTEGER)")
    db.commit()

    TIMINGS.clear()
    t0 = time.time()
    kmer_idx = MPHF_KmerIndex.from_directory(args.cdbg_prefix)
    TIMINGS["load_index"] = time.time() - t0
    assert args.ksize == kmer_idx.ksize
    print("loaded {} k-mers in index ({:.1f}s)".format(len(kmer_idx), time.time() - t0), file=sys.stderr)

    cursor.execute("PRAGMA cache_size=1000000")
    cursor.execute("PRAGMA synchronous = OFF")
    cursor.execute("PRAGMA journal_mode = MEMORY")

    print(f"walking read file: '{args.reads}'")
    t0 = time.time()
    src = bgzf.open_reads(args.reads)
    TIMINGS["inflate"] = time.time() - t0
    t1 = time.time()
    names, seq, seq_off, rec_pos = bgzf.parse_records(src.data)
    offsets = src.virtual_offsets(rec_pos)
    TIMINGS["scan"] = time.time() - t1
    n = len(names)
    total_bp = int(seq_off[-1])

    # NB: pairing links a read to the previous one, so the whole file is labelled as one stream;
    # the GPU pass itself is batched inside the table object.
    ro, ri, n_paired, total_ids, first_paired = label_records(
        names, seq, seq_off, offsets, kmer_idx, args.ksize, ignore_paired=args.ignore_paired, timings=TIMINGS
    )
    if args.expect_paired:
        # the reference checks at every 5e5-bp watermark that a pair has been seen (:108-122)
        before = np.cumsum(np.diff(seq_off)) - np.diff(seq_off)
        crossed = np.flatnonzero(before >= 5e5)
        if len(crossed) and (first_paired < 0 or first_paired >= crossed[0]):
            print("ERROR: no paired reads!? but -P/--expect-paired set. Quitting.")
            return -1
    t1 = time.time()
    cursor.executemany(
        "INSERT INTO sequences (offset, cdbg_id) VALUES (?, ?)", zip(ro.tolist(), ri.tolist())
    )
    db.commit()
    TIMINGS["sqlite_insert"] = time.time() - t1
    TIMINGS["rows"] = len(ro)

    err = lambda s: print(s, file=sys.stderr)  # noqa: E731
    err(f"{total_bp:5.2e} bp in {n} reads")
    err(f"{2*n_paired} paired of {n} reads; {n-2*n_paired} singletons")
    err(f"found reads for {len(total_ids)} cDBG IDs")
    err(f"{time.time() - t0:.1f}s seconds total")

    fail_exit = False
    if not total_ids or max(total_ids) + 1 != len(total_ids):
        fail_exit = True
        err("ERROR: max cDBG ID found is {}, different from:".format(max(total_ids) if total_ids else None))
        err(f"ERROR: length of distinct cDBG IDs found is {len(total_ids)}")

    err("creating indices in database.")
    t1 = time.time()
    cursor.execute("CREATE INDEX offset_idx ON sequences (offset)")
    cursor.execute("CREATE INDEX cdbg_id_idx ON sequences (cdbg_id)")
    db.close()
    TIMINGS["sqlite_index"] = time.time() - t1
    print(f"done; closing {dbfilename}!")

    if args.expect_paired and not n_paired:
        fail_exit = True
        print("ERROR: no paired reads!? but -P/--expect-paired set. Failing.")
    return -1 if fail_exit else 0


if __name__ == "__main__":
    sys.exit(main())
