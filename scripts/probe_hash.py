"""hash-only run for ncu captures: 4M reads, 3 launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
from spacegraphcats_b200.device import get_context
from spacegraphcats_b200.synth import SyntheticCdbg
ctx = get_context(0)
cd = SyntheticCdbg(ctx, 20_000_000)
seq, off = cd.reads(4_000_000)
for it in range(3):
    h, koff, st = ctx.hash_batch(seq, off, 31, to_host=False)
print("ok", h.numel())
