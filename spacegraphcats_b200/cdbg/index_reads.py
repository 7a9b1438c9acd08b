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


def invalid_base_policy():
    """What to do with a read that holds a byte other than A/C/G/T.  ``raise`` (default) follows khmer, which
    throws on such reads; ``skip`` (SGC_B200_INVALID_BASES=skip) follows the comment at index_reads.py:149-150
    ("if reads have 'N' in them, they will not appear in cDBG"): the read is kept in the stream -- pairing and
    offsets are unchanged -- but contributes no unitig ids."""
    v = os.environ.get("SGC_B200_INVALID_BASES", "raise").lower()
    if v not in ("raise", "skip"):
        raise ValueError("SGC_B200_INVALID_BASES must be 'raise' or 'skip'")
    return v


def record_batches(reads_path, ksize, ignore_paired=False, chunk_bytes=None, timings=None, policy="raise"):
    """Walk a read file in bounded memory and yield batches of records ready for ``label_reads_stream``:
    dicts with the 2-bit packed reads (``packed``, ``nseq``, ``lengths``) plus the host-side columns the row
    rules need (``names``, ``lens``, ``offsets``) and ``bad`` (records with a non-ACGT byte).

    A batch always ends just BEFORE a record that is long enough and is not the mate of its predecessor: the
    pairing state of index_reads.py:124-141 (``last_record``, the carried id set) is empty at that point, so
    batches can be labelled and turned into rows independently.  The held-back records open the next batch."""
    from ..device import pack_reads_host

    if chunk_bytes is None:
        chunk_bytes = int(float(os.environ.get("SGC_INDEX_READS_BATCH_MB", "512")) * (1 << 20))
    stream = bgzf.ReadStream(reads_path, chunk_bytes)
    tm = timings if timings is not None else {}
    carry, carry_g = np.zeros(0, np.uint8), 0
    chunks = stream.chunks()
    nxt = next(chunks, None)
    while nxt is not None:
        t0 = time.time()
        data, gstart = nxt
        nxt = next(chunks, None)
        tm["inflate"] = tm.get("inflate", 0.0) + time.time() - t0
        last = nxt is None
        if len(carry):
            data = np.concatenate([carry, data])
            gstart = carry_g
        t1 = time.time()
        names, seq, seq_off, rec_pos = bgzf.parse_records(data)
        n = len(names)
        lens = np.diff(seq_off)
        cut = n
        if not last:
            # hold back from the last record that is long enough and unpaired (it resets the pairing state)
            elig = lens >= ksize
            paired, _ = pair_flags(names, elig, ignore_paired)
            safe = np.flatnonzero(elig & ~paired)
            safe = safe[safe > 0]
            if len(safe) == 0:  # no safe boundary in this chunk (pathological): grow it with the next one
                carry, carry_g = data, gstart
                continue
            cut = int(safe[-1])
            carry, carry_g = data[rec_pos[cut]:].copy(), gstart + int(rec_pos[cut])
        else:
            carry = np.zeros(0, np.uint8)
        tm["scan"] = tm.get("scan", 0.0) + time.time() - t1
        t2 = time.time()
        sub_off = seq_off[: cut + 1]
        packed, plen, bad = pack_reads_host(seq[: int(sub_off[-1])], sub_off, n_threads=bgzf.n_threads())
        if bad.any():
            if policy == "raise":
                raise ValueError("Invalid base in read %r" % names[int(np.flatnonzero(bad)[0])])
            # skip: the record stays in the stream (pairing, offsets) but is packed with length 0: no k-mers
            sub_lens = lens[:cut]
            keep = np.repeat(~bad, sub_lens)
            off2 = np.zeros(cut + 1, np.int64)
            np.cumsum(np.where(bad, 0, sub_lens), out=off2[1:])
            packed, plen, _ = pack_reads_host(seq[: int(sub_off[-1])][keep], off2, n_threads=bgzf.n_threads())
        tm["pack"] = tm.get("pack", 0.0) + time.time() - t2
        sub_names = bgzf.NameTable(names.data, names.lo[:cut], names.hi[:cut])
        yield {"packed": packed, "nseq": cut, "lengths": plen, "names": sub_names, "lens": lens[:cut], "bad": bad,
               "offsets": stream.virtual_offsets(gstart + rec_pos[:cut])}


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
    policy = invalid_base_policy()
    watermark = 5e5  # the reference checks for pairs every 5e5 bp; only the first crossing can fail (:107-122)
    n = total_bp = n_paired = n_rows = n_batches = n_skipped = 0
    first_paired_seen = False
    seen = np.zeros(0, bool)  # unitig ids that received reads

    batches = lambda: record_batches(args.reads, args.ksize, args.ignore_paired, timings=TIMINGS, policy=policy)  # noqa: E731

    # the GPU labels batch i while the helper thread inflates / scans / packs batch i+1 and this thread turns
    # the labels of batch i-1 into rows: memory is bounded by a few batches whatever the size of the read file
    for batch, res in kmer_idx.table.label_reads_stream(batches(), args.ksize):
        t1 = time.time()
        lens = batch["lens"]
        n_skipped += int(batch["bad"].sum())  # such a read keeps its place in the stream (it can be a mate) but has no ids
        ro, ri, npair, first_paired = rows_from_labels(batch["names"], lens, batch["offsets"], res["ids_off"], res["ids"],
                                                        args.ksize, args.ignore_paired)
        if args.expect_paired and not first_paired_seen:
            before = total_bp + np.cumsum(lens) - lens
            crossed = np.flatnonzero(before >= watermark)
            if len(crossed) and (first_paired < 0 or first_paired >= crossed[0]):
                print("ERROR: no paired reads!? but -P/--expect-paired set. Quitting.")
                return -1
            if first_paired >= 0:
                first_paired_seen = True
        TIMINGS["pairing_rows"] = TIMINGS.get("pairing_rows", 0.0) + time.time() - t1
        t1 = time.time()
        cursor.executemany("INSERT INTO sequences (offset, cdbg_id) VALUES (?, ?)", zip(ro.tolist(), ri.tolist()))
        TIMINGS["sqlite_insert"] = TIMINGS.get("sqlite_insert", 0.0) + time.time() - t1
        if len(ri):
            top = int(ri.max()) + 1
            if top > len(seen):
                seen = np.concatenate([seen, np.zeros(top - len(seen), bool)])
            seen[np.unique(ri)] = True
        n += len(lens)
        total_bp += int(lens.sum())
        n_paired += npair
        n_rows += len(ro)
        n_batches += 1
    db.commit()
    TIMINGS["rows"] = n_rows
    TIMINGS["batches"] = n_batches

    err = lambda s: print(s, file=sys.stderr)  # noqa: E731
    n_ids = int(seen.sum())
    err(f"{total_bp:5.2e} bp in {n} reads")
    err(f"{2*n_paired} paired of {n} reads; {n-2*n_paired} singletons")
    if n_skipped:
        err(f"{n_skipped} reads with non-ACGT bases contributed no cDBG IDs (SGC_B200_INVALID_BASES=skip)")
    err(f"found reads for {n_ids} cDBG IDs")
    err(f"{time.time() - t0:.1f}s seconds total")

    fail_exit = False
    if not n_ids or len(seen) != n_ids:
        fail_exit = True
        err("ERROR: max cDBG ID found is {}, different from:".format(len(seen) - 1 if len(seen) else None))
        err(f"ERROR: length of distinct cDBG IDs found is {n_ids}")

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
