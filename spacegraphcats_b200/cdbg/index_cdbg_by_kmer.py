"""Build the k-mer -> cDBG-unitig table on the GPU (reference:
``spacegraphcats/cdbg/index_cdbg_by_kmer.py``: ``build_mphf`` :27-115, ``main`` :118-139).

The reference makes two passes over the unitigs with Python sets and then builds a BBHash MPHF.
Here both passes are ONE fused kernel launch per batch of unitigs (hash every k-mer, insert it
under the unitig's id; a hash that arrives under two different ids is dropped, which is exactly
the reference's ``multicounts`` rule), and the observable result -- the exact map, ``sizes``, the
consistency asserts -- is the same.
"""
import argparse
import sqlite3
import sys

import numpy as np

from ..device import get_context
from .hashing import GpuKmerTable, MPHF_KmerIndex, UNSET_VALUE, side_default

BATCH_BASES = 1 << 28  # unitig bases per insert launch


def build_mphf(ksize, records_iter_fn, ctx=None, verbose=True):
    """-> (table, sizes) like the reference (index_cdbg_by_kmer.py:27-115).

    ``records_iter_fn()`` yields records with ``.name`` (decimal unitig id) and ``.sequence``.
    """
    ctx = ctx or get_context()
    say = print if verbose else (lambda *a, **k: None)
    ksize = int(ksize)

    # one pass on the host just to size the table and collect ids / lengths
    names, lengths, chunks = [], [], []
    for record in records_iter_fn():
        seq = record.sequence
        seq = seq.encode("ascii") if isinstance(seq, str) else bytes(seq)
        if len(seq) < ksize:  # khmer raises for these (cf. hashing.hash_sequence)
            raise ValueError(
                "sequence length ({}) must >= the hashtable k-mer size ({})".format(len(seq), ksize)
            )
        names.append(int(record.name))
        lengths.append(len(seq))
        chunks.append(seq)
    n_contigs = len(names)
    if n_contigs == 0:
        raise ValueError("no contigs to index")
    say(f"loaded {n_contigs} contigs.\n")

    ids = np.array(names, np.int64)
    if ids.min() < 0 or ids.max() >= UNSET_VALUE - 1:
        raise ValueError("cDBG ids must be in 0..2^32-3")
    lengths = np.array(lengths, np.int64)
    nk = lengths - ksize + 1
    total_kmers = int(nk.sum())

    table = GpuKmerTable(ctx)
    from ..device import DeviceTable

    table._dev = DeviceTable(ctx, max(total_kmers, 1))
    total_bases = int(lengths.sum())
    # unitig sequence store for read labelling (SGC_SIDE=0: off); every batch starts at a
    # 64-base boundary of the store
    if side_default(total_bases + 64 * (total_bases // BATCH_BASES + n_contigs // 1024 + 2)):
        table._dev.enable_side()
    lo = 0
    while lo < n_contigs:
        hi, nb = lo, 0
        while hi < n_contigs and (hi == lo or nb + lengths[hi] <= BATCH_BASES):
            nb += int(lengths[hi])
            hi += 1
        buf = np.frombuffer(b"".join(chunks[lo:hi]), np.uint8)
        off = np.zeros(hi - lo + 1, np.int64)
        np.cumsum(lengths[lo:hi], out=off[1:])
        status = ctx.to_host(table._dev.insert_sequences(buf, off, ksize, ids=ids[lo:hi].astype(np.uint32)))
        if status.any():
            raise ValueError("Invalid base in read (contig %s)" % names[lo + int(np.flatnonzero(status)[0])])
        lo = hi
    del chunks

    live, n_multi, max_table_value = table.stats()
    if n_multi:
        say("NOTE: likely hash collisions (or duplicate k-mers?) in input cDBG")
        say(f"NOTE: {n_multi} k-mer hash values are present more than once.")
        say("NOTE: these k-mers are being removed from consideration.")
    else:
        say("NOTE: no multicount hashvals detected.")
    say(f"built GPU table for {live} k-mers in {n_contigs} nodes.")

    sizes = dict(zip(ids.tolist(), nk.tolist()))  # :91 -- L-k+1, duplicates and multicounts included

    # the reference's consistency asserts (:103-113)
    n = n_contigs - 1
    assert n <= UNSET_VALUE
    assert live > 0 and max_table_value != UNSET_VALUE
    assert n == int(ids.max())
    assert n == max_table_value, (n, max_table_value)
    return table, sizes


def contigs_iter_sqlite(contigs_db):
    """Records (.name = str(id), .sequence) from the unitig sqlite DB, like
    search_utils.contigs_iter_sqlite (search/search_utils.py:200-208)."""

    class Record:
        __slots__ = ("name", "sequence")

        def __init__(self, name, sequence):
            self.name, self.sequence = name, sequence

    cursor = contigs_db.cursor()
    cursor.execute("SELECT id, sequence FROM sequences")
    for ident, sequence in cursor:
        yield Record(str(ident), sequence)


def main(argv):
    p = argparse.ArgumentParser()
    p.add_argument("catlas_prefix")
    p.add_argument("--contigs-db", required=True)
    p.add_argument("-k", "--ksize", default=31, type=int)
    args = p.parse_args(argv)

    sqlite_db = sqlite3.connect(args.contigs_db)

    def create_records_iter():
        print(f"reading cDBG nodes from sqlite DB {args.contigs_db}")
        return contigs_iter_sqlite(sqlite_db)

    table, sizes = build_mphf(args.ksize, create_records_iter)
    print(f"done! saving to {args.catlas_prefix}")
    MPHF_KmerIndex(args.ksize, table, sizes).save(args.catlas_prefix)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
