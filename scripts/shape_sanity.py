"""Timing sanity on awkward input shapes (one huge sequence, millions of tiny ones, unitig-like ragged)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from spacegraphcats_b200.device import get_context
from spacegraphcats_b200.synth import SyntheticCdbg
ctx = get_context(0)
cd = SyntheticCdbg(ctx, 50_000_000)
table = cd.build_table()
def timed(name, fn, n):
    fn(); ctx.sync(); t0 = time.perf_counter(); fn(); ctx.sync(); dt = time.perf_counter() - t0
    print("%-52s %8.2f ms  %7.2f G k-mers/s" % (name, dt * 1e3, n / dt / 1e9), flush=True)
# 1. one 50 Mbp sequence (a genome query)
G = cd.genome; off1 = torch.tensor([0, G.numel()], dtype=torch.int64, device=ctx.device)
timed("hash: one 50 Mbp sequence", lambda: ctx.hash_batch(G, off1, 31, to_host=False), G.numel() - 30)
timed("label: one 50 Mbp sequence", lambda: table.label_reads(G, off1, 31, to_host=False), G.numel() - 30)
timed("lookup_sequences(count): one 50 Mbp sequence", lambda: table.lookup_sequences(G, off1, 31, n_ids=cd.n_unitigs, want_ids=False, to_host=False), G.numel() - 30)
# 2. 2M sequences of exactly k bases (1 k-mer each) + empties
n = 2_000_000
lens = torch.full((n,), 31, dtype=torch.int64, device=ctx.device); lens[::5] = 0; lens[1::7] = 12
off2 = torch.zeros(n + 1, dtype=torch.int64, device=ctx.device); off2[1:] = torch.cumsum(lens, 0)
assert int(off2[-1].item()) <= G.numel()
seq2 = G[: int(off2[-1].item())].contiguous()
nk2 = int((lens >= 31).sum().item())
timed("label: 2M sequences of 0 / 12 / 31 bases", lambda: table.label_reads(seq2, off2, 31, to_host=False), nk2)
# 3. the unitigs themselves (ragged 31..4126 bases)
timed("label: 780k ragged unitigs", lambda: table.label_reads(cd.useq, cd.uoff, 31, to_host=False), cd.total_unitig_kmers)
# 4. k = 21 and a runtime-k (k = 25) pass over reads
seq, off = cd.reads(2_000_000)
for k in (21, 25, 31):
    timed("hash: 2M reads, k=%d" % k, lambda k=k: ctx.hash_batch(seq, off, k, to_host=False), 2_000_000 * (150 - k + 1))
