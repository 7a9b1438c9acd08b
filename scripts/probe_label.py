"""Small label_reads run for ncu captures: 1e8-k-mer table, 4M reads, 3 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
from spacegraphcats_b200.device import get_context
from spacegraphcats_b200.synth import SyntheticCdbg
n_kmers = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
n_reads = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4_000_000
ctx = get_context(0)
cd = SyntheticCdbg(ctx, n_kmers)
table = cd.build_table()
seq, off = cd.reads(n_reads)
for it in range(3):
    res = table.label_reads(seq, off, 31, to_host=False)
print("ok", res["n_kmers"], res["n_hits"], res["n_labels"])
