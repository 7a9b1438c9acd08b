"""Hash-range partitioned table over N GPUs (BASELINE.json configs[4] shape, scaled by flags):
torchrun --nproc-per-node N scripts/partitioned_run.py [--table-kmers 1e9] [--reads 4e6]
Every rank builds 1/N of the unitigs, the table is partitioned by hash range with an NCCL
all-to-all, each rank labels its own reads through request/reply all-to-alls.  Rank 0 checks its
labels against a replicated table built on the same GPU (when --check) and prints k-mers/s."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--table-kmers", type=float, default=4e8)
ap.add_argument("--reads", type=float, default=4e6)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--check", action="store_true")
ap.add_argument("--p2p", action="store_true", help="fuse the exchange into the kernels (NVLink peer stores)")
ap.add_argument("--peer", action="store_true", help="probe the owners' tables over NVLink peer loads + replicated sequence store")
ap.add_argument("--pipeline", type=int, default=0, help="with --p2p: split each step into this many sub-batches and overlap their phases")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
import __graft_entry__ as g
if rank == 0: g.build()
dist.barrier()
from spacegraphcats_b200.device import Context
from spacegraphcats_b200.synth import SyntheticCdbg
from spacegraphcats_b200.parallel import DeviceEngine, PartitionedIndex, shard_range

ctx = Context(local)
K = 31
n_kmers, n_reads = int(args.table_kmers), int(args.reads)
cd = SyntheticCdbg(ctx, n_kmers, K, seed=1)          # same genome/unitigs on every rank (same seeds)
ulo, uhi = shard_range(cd.n_unitigs, rank, world)
uoff_all = cd.uoff
with torch.cuda.stream(ctx.stream):
    b0, b1 = int(uoff_all[ulo].item()), int(uoff_all[uhi].item())
    useq = cd.useq[b0:b1]
    uoff = (uoff_all[ulo:uhi + 1] - b0).contiguous()
    ids = torch.arange(ulo, uhi, dtype=torch.int32, device=ctx.device)
eng = DeviceEngine(ctx)
t0 = time.time()
idx = PartitionedIndex(eng, p2p=args.p2p, peer_probe=args.peer).build(useq, uoff, K, ids)
ctx.sync(); dist.barrier()
if rank == 0:
    print("partitioned build: %.2fs, rank0 holds %d hashes (%.2f GB)" % (time.time() - t0, idx.n_local, idx.table.nbytes / 1e9), flush=True)
seq, off = cd.reads(n_reads, 150, seed=100 + rank)
ctx.sync()
res = None
subs = None
if args.pipeline > 1:
    assert args.p2p
    per = n_reads // args.pipeline
    with torch.cuda.stream(ctx.stream):
        subs = [(seq[j * per * 150:(j + 1) * per * 150], (off[j * per:(j + 1) * per + 1] - off[j * per]).contiguous())
                for j in range(args.pipeline)]
    ctx.sync()
import ctypes
trace = bool(os.environ.get("SGC_TRACE"))
for it in range(args.steps + 1):
    if it == 1:
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        if trace: ctx._L.sgc_ctx_set_timing(ctx._h, 1)
    if subs is not None:
        parts = idx.label_reads_pipelined(subs, K)
        continue
    res = idx.label_reads(seq, off, K)
if subs is not None:
    torch.cuda.synchronize()
    ids_all = torch.cat([p[1] for p in parts])
    offs, base = [parts[0][0]], int(parts[0][0][-1])
    for p in parts[1:]:
        offs.append(p[0][1:] + base); base += int(p[0][-1])
    res = (torch.cat(offs), ids_all, sum(p[2] for p in parts))
ctx.sync()
t_mine = time.perf_counter() - t0
dist.barrier(); torch.cuda.synchronize()
dt = time.perf_counter() - t0
if trace:
    ms, n = ctypes.c_double(0), ctypes.c_int64(0)
    ctx._L.sgc_ctx_tile_ms(ctx._h, ctypes.byref(ms), ctypes.byref(n))
    print("[rank %d] my loop %.1f ms; label kernel: %d launches, %.2f ms each" % (rank, t_mine * 1e3, n.value, ms.value / max(n.value, 1)), flush=True)
t = torch.tensor([dt], dtype=torch.float64, device=ctx.device); dist.all_reduce(t, op=dist.ReduceOp.MAX)
rate = world * n_reads * 120 * args.steps / float(t.item())
if args.check:
    table = cd.build_table()
    ref = table.label_reads(seq, off, K, to_host=False)
    ok = torch.equal(ref["ids_off"], res[0]) and torch.equal(ref["ids"], res[1])
    okt = torch.tensor([1 if ok else 0], device=ctx.device); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0: print("labels identical to the replicated table on every rank:", bool(okt.item()), flush=True)
idx.close()
if rank == 0:
    print(json.dumps({"mode": "partitioned-peer-probe" if args.peer else ("partitioned-p2p-pipelined%d" % args.pipeline if args.pipeline > 1 else "partitioned-p2p") if args.p2p else "partitioned-nccl", "n_gpus": world, "table_kmers": n_kmers, "reads_per_rank_step": n_reads,
                      "value": rate, "unit": "k-mers/s", "ms_per_step": float(t.item()) / args.steps * 1e3}), flush=True)
dist.destroy_process_group()
