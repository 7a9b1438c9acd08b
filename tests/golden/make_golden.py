#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference checkout's DATA files.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py [/root/reference]``.

Nothing here imports or executes reference *code* (it cannot run: khmer, bbhash
and screed are absent).  The script only copies / re-encodes
the reference's own golden data so the oracle and the CUDA path can be pinned
against it:

  dory_k21_unitigs.tsv.gz      <- tests/test-data/catlas.dory_k21/bcalm.unitigs.db
                                  (sqlite ``sequences(id, sequence)``; "id<TAB>sequence")
  dory_k21_min_hashval.json    <- same DB, column ``min_hashval`` (sourmash MinHash n=1, seed 42, written by
                                  cdbg/sort_bcalm_unitigs.py:56-128)
  dory_k21_contigs.indices.npz <- .../catlas.dory_k21/contigs.indices  (verbatim; written
                                  by the reference's BBHashTable.save: mphf_to_hash u64[n],
                                  mphf_to_value u32[n] -- 212,475 (hash -> unitig id) pairs)
  dory_k21_contigs.mphf        <- .../catlas.dory_k21/contigs.mphf  (verbatim; the boomphf dump written by the
                                  reference's BBHashTable.save: gamma 1.0, 25 levels, 212,475 keys)
  dory_k21_contigs.sizes.json  <- .../catlas.dory_k21/contigs.sizes (pickle (k, {id: n_kmers}))
  dory_k21_r1/catlas.csv, first_doms.txt  <- .../catlas.dory_k21_r1/  (verbatim)
  dory_k21_r1_search/*         <- .../catlas.dory_k21_r1_search_oh0/  (expected query outputs)
  data/*                       <- data/ and tests/test-data/ sequence files (verbatim, gz'd)
"""
import gzip
import json
import os
import pickle
import shutil
import sqlite3
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ref(*p):
    return os.path.join(REF, *p)


def out(*p):
    path = os.path.join(HERE, *p)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    return path


def gz_copy(src, dst):
    with open(src, "rb") as fi, open(dst, "wb") as raw:
        # mtime=0 / no filename => byte-reproducible output
        with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fo:
            shutil.copyfileobj(fi, fo)


def main():
    # unitigs
    db = sqlite3.connect("file:%s?mode=ro" % ref("tests/test-data/catlas.dory_k21/bcalm.unitigs.db"), uri=True)
    rows = db.execute("SELECT id, sequence FROM sequences ORDER BY id").fetchall()
    with open(out("dory_k21_unitigs.tsv.gz"), "wb") as raw:
        with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fo:
            for i, s in rows:
                fo.write(("%d\t%s\n" % (i, s)).encode())
    mins = db.execute("SELECT id, min_hashval FROM sequences ORDER BY id").fetchall()
    with open(out("dory_k21_min_hashval.json"), "w") as fp:  # sourmash n=1 MinHash (seed 42) per unitig
        json.dump({str(i): int(v) for i, v in mins}, fp)
    shutil.copyfile(ref("tests/test-data/catlas.dory_k21/contigs.indices"), out("dory_k21_contigs.indices.npz"))
    shutil.copyfile(ref("tests/test-data/catlas.dory_k21/contigs.mphf"), out("dory_k21_contigs.mphf"))
    with open(ref("tests/test-data/catlas.dory_k21/contigs.sizes"), "rb") as fp:
        k, sizes = pickle.load(fp)
    with open(out("dory_k21_contigs.sizes.json"), "w") as fp:
        json.dump({"ksize": k, "sizes": {str(a): b for a, b in sorted(sizes.items())}}, fp)
    shutil.copyfile(ref("tests/test-data/catlas.dory_k21/contigs.info.csv"), out("dory_k21_contigs.info.csv"))
    for f in ("catlas.csv", "first_doms.txt"):
        shutil.copyfile(ref("tests/test-data/catlas.dory_k21_r1", f), out("dory_k21_r1", f))
    for f in (
        "results.csv",
        "dory-head.fa.response.txt",
        "random-query-nomatch.fa.response.txt",
        "dory-head.fa.cdbg_ids.txt.gz",
        "dory-head.fa.frontier.txt.gz",
        "dory-head.fa.cdbg_ids.reads.fa.gz",
    ):
        shutil.copyfile(ref("tests/test-data/catlas.dory_k21_r1_search_oh0", f), out("dory_k21_r1_search", f))
    # sequence data
    for f in ("dory-head.fa", "random-query-nomatch.fa", "dory-k21-hashval-queries.txt", "dory-k31-hashval-queries.txt"):
        shutil.copyfile(ref("data", f), out("data", f))
    for f in ("dory-subset.fq", "dory-subset.fa", "shew-os223-200k.fa", "bcalm.dory.k21.unitigs.fa"):
        gz_copy(ref("data", f), out("data", f + ".gz"))
    for f in ("twofoo-genes.fa.gz",) + tuple("acido-chunk%d.fa.gz" % i for i in range(1, 9)):
        shutil.copyfile(ref("data", f), out("data", f))
    for f in ("twofoo-short-paired-test.fa.gz", "twofoo-short-reads.fa.gz", "63.short.fa.gz", "2.short.fa.gz", "47.short.fa.gz"):
        shutil.copyfile(ref("tests/test-data", f), out("data", f))
    print("golden fixtures written under", HERE)


if __name__ == "__main__":
    main()
