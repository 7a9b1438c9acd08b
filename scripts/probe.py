"""First-contact GPU probe: stage-by-stage throughput on a reduced synthetic workload."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from spacegraphcats_b200.device import get_context
from spacegraphcats_b200.synth import SyntheticCdbg

n_kmers = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
n_reads = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4_000_000
ctx = get_context(0)
def ev(): 
    e = torch.cuda.Event(enable_timing=True); e.record(ctx.stream); return e
t0 = time.time()
cd = SyntheticCdbg(ctx, n_kmers)
print("synth cdbg: %d unitigs, %.2fs" % (cd.n_unitigs, time.time() - t0), flush=True)
a = ev(); table = cd.build_table(); b = ev(); ctx.sync()
print("table build: %.1f ms -> %.2f G k-mers/s; table %.2f GB; stats %s" % (a.elapsed_time(b), cd.total_unitig_kmers / a.elapsed_time(b) / 1e6, table.nbytes / 1e9, table.stats()), flush=True)
seq, off = cd.reads(n_reads)
ctx.sync()
nk = n_reads * 120
for it in range(3):
    a = ev(); h, koff, st = ctx.hash_batch(seq, off, 31, to_host=False); b = ev(); ctx.sync()
    print("hash_batch: %.2f ms -> %.1f G k-mers/s" % (a.elapsed_time(b), nk / a.elapsed_time(b) / 1e6), flush=True)
for it in range(3):
    a = ev(); ids = table.lookup(h, to_host=False); b = ev(); ctx.sync()
    print("lookup (precomputed hashes): %.2f ms -> %.1f G/s" % (a.elapsed_time(b), nk / a.elapsed_time(b) / 1e6), flush=True)
rnd = torch.randint(-2**62, 2**62, (nk,), device=ctx.device, dtype=torch.int64)
for it in range(2):
    a = ev(); ids = table.lookup(rnd, to_host=False); b = ev(); ctx.sync()
    print("lookup (random misses): %.2f ms -> %.1f G/s" % (a.elapsed_time(b), nk / a.elapsed_time(b) / 1e6), flush=True)
ctx._L.sgc_ctx_set_timing(ctx._h, 1)
import ctypes
for it in range(3):
    a = ev(); res = table.lookup_sequences(seq, off, 31, want_ids=False, to_host=False); b = ev(); ctx.sync()
    ms = ctypes.c_double(); n = ctypes.c_int64(); ctx._L.sgc_ctx_tile_ms(ctx._h, ctypes.byref(ms), ctypes.byref(n))
    print("lookup_sequences (fused, no out): %.2f ms (tile %.2f ms) -> %.1f G/s; hits %.3f" % (a.elapsed_time(b), ms.value, nk / ms.value / 1e6, res["n_hits"] / res["n_kmers"]), flush=True)
for it in range(3):
    a = ev(); res = table.label_reads(seq, off, 31, to_host=False); b = ev(); ctx.sync()
    ms = ctypes.c_double(); n = ctypes.c_int64(); ctx._L.sgc_ctx_tile_ms(ctx._h, ctypes.byref(ms), ctypes.byref(n))
    print("label_reads: %.2f ms total, tile kernel %.2f ms -> %.1f G k-mers/s (tile) %.1f (step); labels/read %.2f hits %.3f" % (
        a.elapsed_time(b), ms.value, nk / ms.value / 1e6, nk / a.elapsed_time(b) / 1e6, res["n_labels"] / n_reads, res["n_hits"] / res["n_kmers"]), flush=True)
