#!/usr/bin/env python
"""End-to-end index_reads on a synthetic cDBG + paired FASTQ in BGZF: stage
breakdown of spacegraphcats_b200.cdbg.index_reads.main.  Not a bench line; numbers go to DESIGN.md.

usage: index_reads_timing.py [n_reads=4000000] [table_kmers=50000000]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import __graft_entry__ as g  # noqa: E402

g.build()
from spacegraphcats_b200.cdbg import index_reads  # noqa: E402
from spacegraphcats_b200.cdbg.hashing import GpuKmerTable, MPHF_KmerIndex  # noqa: E402
from spacegraphcats_b200.device import get_context  # noqa: E402
from spacegraphcats_b200.synth import SyntheticCdbg  # noqa: E402
from bgzf_writer import write_bgzf  # noqa: E402  (test-only BGZF writer)

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
n_kmers = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
K, RL = 31, 150
ctx = get_context(0)
tmp = tempfile.mkdtemp(prefix="sgc_ir_")

t0 = time.time()
cdbg = SyntheticCdbg(ctx, n_kmers, K, seed=1, dup_fraction=0.0)
table = GpuKmerTable.from_sequences(cdbg.useq, cdbg.uoff, K, ctx=ctx)
sizes = dict(enumerate((np.diff(ctx.to_host(cdbg.uoff)) - K + 1).tolist()))
t1 = time.time()
MPHF_KmerIndex(K, table, sizes).save(tmp)
t2 = time.time()
print("cDBG: %d unitigs, %d k-mers; table build %.2fs, save (BBHash on GPU + npz + pickle) %.2fs"
      % (cdbg.n_unitigs, len(table), t1 - t0, t2 - t1), flush=True)
del table

# paired FASTQ text, fixed-width records, built with numpy
seq, _ = cdbg.reads(n_reads, RL, seed=5, err_ppm=5000)
seq = ctx.to_host(seq).reshape(n_reads, RL)
name_w = 10
W = 1 + name_w + 2 + 1 + RL + 3 + RL + 1
rec = np.full((n_reads, W), ord("I"), np.uint8)
rec[:, 0] = ord("@")
pair = np.arange(n_reads) // 2
for d in range(name_w):
    rec[:, 1 + d] = ord("0") + (pair // 10 ** (name_w - 1 - d)) % 10
rec[:, 1 + name_w] = ord("/")
rec[:, 2 + name_w] = ord("1") + (np.arange(n_reads) & 1)
rec[:, 3 + name_w] = 10
rec[:, 4 + name_w : 4 + name_w + RL] = seq
rec[:, 4 + name_w + RL : 7 + name_w + RL] = np.frombuffer(b"\n+\n", np.uint8)
rec[:, -1] = 10
t0 = time.time()
reads_path = os.path.join(tmp, "reads.fq.bgz")
write_bgzf(reads_path, rec.reshape(-1), level=1)
print("reads: %d x %d bp, %.2f GB FASTQ -> %.2f GB BGZF in %.1fs" % (n_reads, RL, rec.size / 1e9,
      os.path.getsize(reads_path) / 1e9, time.time() - t0), flush=True)
fastq_bytes = rec.size
del rec, seq
torch.cuda.empty_cache()

t0 = time.time()
rc = index_reads.main([tmp, reads_path, os.path.join(tmp, "reads.index"), "-k", str(K)])
wall = time.time() - t0
T = dict(index_reads.TIMINGS)
T.update(rc=rc, wall=wall, n_reads=n_reads, fastq_gb=fastq_bytes / 1e9, kmers=n_reads * (RL - K + 1),
         host_threads=len(os.sched_getaffinity(0)))
print(json.dumps(T))
front = T["inflate"] + T["scan"] + T["label_gpu"] + T["pairing_rows"]
print("front end (inflate+scan+GPU label+pairing rows): %.2fs = %.2f M reads/s, %.2f G k-mers/s; sqlite insert+index: %.1fs"
      % (front, n_reads / front / 1e6, T["kmers"] / front / 1e9, T["sqlite_insert"] + T["sqlite_index"]))
