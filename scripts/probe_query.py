#!/usr/bin/env python
"""One batch of the bench's query workload (for ncu launch lists): probe_query.py [queries=512] [kmers=500000]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import __graft_entry__ as g
g.build()
from spacegraphcats_b200.device import Context
from spacegraphcats_b200.synth import SyntheticCdbg
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 512
qk = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
ctx = Context(0)
cd = SyntheticCdbg(ctx, 500_000_000, 31, seed=1)
t = cd.build_table()
seq, off = cd.reads(nq, qk + 30, seed=7000, err_ppm=5000)
for i in range(2):
    res = t.count_matches_batch(seq, off, np.arange(nq, dtype=np.int32), nq, 31)
print(len(res["keys"]), res["n_kmers"], res["n_hits"])
