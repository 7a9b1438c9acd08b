"""Label every read with the cDBG unitigs its k-mers hit and write the ``(offset, cdbg_id)``
sqlite table (reference: ``spacegraphcats/cdbg/index_reads.py`` ``main`` :39-201).

The per-read arithmetic -- ``hash_sequence`` + ``count_cdbg_matches`` (:147-152) -- runs as ONE
fused GPU pass per batch of reads (hash, table probe, per-read distinct ids as CSR).  The
pairing rules (:124-141), the too-short skip (:143-144), the sqlite schema (:72, :187-188) and
the exit codes are reproduced on the host in record order.
"""
import argparse
import os
import sqlite3
import sys
import time

import numpy as np

from ..utils import bgzf
from .hashing import MPHF_KmerIndex

DEFAULT_KSIZE = 31
BATCH_READS = 1 << 22


def label_records(names, seq, seq_off, offsets, kmer_idx, ksize, ignore_paired=False):
    """The loop of index_reads.py:105-167 over an in-memory batch.

    -> (rows_offset int64[m], rows_id int64[m], n_paired, set_of_all_ids, first_paired).
    ``rows_*`` are the rows the reference INSERTs (as a multiset; duplicates included, cf.
    :156-157); ``first_paired`` is the index of the first read found paired (-1: none)."""
    n = len(names)
    res = kmer_idx.table.label_reads(seq, seq_off, ksize)
    if res["status"].any():
        bad = int(np.flatnonzero(res["status"])[0])
        raise ValueError("Invalid base in read %r" % names[bad])
    ids_off, ids = res["ids_off"], res["ids"].astype(np.int64)
    lens = np.diff(seq_off)

    rows_o, rows_i = [], []
    n_paired = 0
    first_paired = -1
    last = -1  # index of last_record
    cur = ()  # cdbg_ids carried from the mate
    for i in range(n):
        paired = False
        if last >= 0:
            a, b = names[last], names[i]
            if a.endswith("/1") and b.endswith("/2"):
                paired = a[:-2] == b[:-2] and len(a) > 2
            elif a == b and a:
                paired = True
        if paired:
            n_paired += 1
            if first_paired < 0:
                first_paired = i
        else:
            cur = ()
        if lens[i] < ksize:
            continue
        mine = ids[ids_off[i] : ids_off[i + 1]]
        cur = np.union1d(cur, mine) if len(cur) else mine
        if len(cur):
            if paired:
                rows_o.append(np.repeat(np.array([offsets[last], offsets[i]], np.int64), len(cur)))
                rows_i.append(np.tile(cur, 2))
            else:
                rows_o.append(np.full(len(cur), offsets[i], np.int64))
                rows_i.append(cur)
        if not ignore_paired:
            last = i
    ro = np.concatenate(rows_o) if rows_o else np.zeros(0, np.int64)
    ri = np.concatenate(rows_i) if rows_i else np.zeros(0, np.int64)
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

    t0 = time.time()
    kmer_idx = MPHF_KmerIndex.from_directory(args.cdbg_prefix)
    assert args.ksize == kmer_idx.ksize
    print("loaded {} k-mers in index ({:.1f}s)".format(len(kmer_idx), time.time() - t0), file=sys.stderr)

    cursor.execute("PRAGMA cache_size=1000000")
    cursor.execute("PRAGMA synchronous = OFF")
    cursor.execute("PRAGMA journal_mode = MEMORY")

    print(f"walking read file: '{args.reads}'")
    t0 = time.time()
    src = bgzf.open_reads(args.reads)
    names, seq, seq_off, rec_pos = bgzf.parse_records(src.data)
    offsets = src.virtual_offsets(rec_pos)
    n = len(names)
    total_bp = int(seq_off[-1])

    # NB: pairing links a read to the previous one, so the whole file is labelled as one stream;
    # the GPU pass itself is batched inside the table object.
    ro, ri, n_paired, total_ids, first_paired = label_records(
        names, seq, seq_off, offsets, kmer_idx, args.ksize, ignore_paired=args.ignore_paired
    )
    if args.expect_paired:
        # the reference checks at every 5e5-bp watermark that a pair has been seen (:108-122)
        before = np.cumsum(np.diff(seq_off)) - np.diff(seq_off)
        crossed = np.flatnonzero(before >= 5e5)
        if len(crossed) and (first_paired < 0 or first_paired >= crossed[0]):
            print("ERROR: no paired reads!? but -P/--expect-paired set. Quitting.")
            return -1
    cursor.executemany(
        "INSERT INTO sequences (offset, cdbg_id) VALUES (?, ?)", zip(ro.tolist(), ri.tolist())
    )
    db.commit()

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
    cursor.execute("CREATE INDEX offset_idx ON sequences (offset)")
    cursor.execute("CREATE INDEX cdbg_id_idx ON sequences (cdbg_id)")
    db.close()
    print(f"done; closing {dbfilename}!")

    if args.expect_paired and not n_paired:
        fail_exit = True
        print("ERROR: no paired reads!? but -P/--expect-paired set. Failing.")
    return -1 if fail_exit else 0


if __name__ == "__main__":
    sys.exit(main())
