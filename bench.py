#!/usr/bin/env python
"""bench.py -- k-mers hashed + looked-up per second on the index_reads hot path.

Workload (BASELINE.json configs[3]): synthetic metagenome cDBG with
500 M unitig k-mers (k=31) in a GPU-resident table replicated per GPU; 150-bp reads with 0.5 %
substitution errors, half reverse-complemented, labelled with the distinct unitig ids their
k-mers hit (cdbg/index_reads.py:143-152).  One STEP = one batch of 2^24 reads (2.01e9 k-mers,
2.5 GB of ASCII, far larger than L2); the default 12 steps are the config's 200 M reads.

  value  k-mers/s, whole job, inputs resident in HBM, timed with CUDA events on the launch stream
  e2e    the same through the public host API: pinned host buffers -> H2D -> kernels -> D2H of the
         CSR labels, double-buffered; H2D / D2H bytes counted from the tensors copied
  roofline   the tile kernel (hash + probe + label) alone, CUDA events recorded around its launches
  cpu_baseline / --impl reference   the CPU oracle (oracle/sgc_oracle.c, OpenMP over all host
         cores) on a bounded, scaled-down sample of the same workload

N > 1 (torchrun): every rank holds the full table and labels its own read batches (weak scaling,
no data-path collective); the time is the max over ranks.  ``--table partitioned`` instead runs the
hash-range partitioned table of configs[4] (exchange fused into the kernels over NVLink peer memory;
``partitioned-nccl`` = the same regions through NCCL all-to-all) and reports the NVLink roofline.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KSIZE = 31
READ_LEN = 150
ERR_PPM = 5000
SECTOR = 32


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi SM clock / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def cpu_workload(n_kmers, n_reads, seed=1):
    """Scaled-down copy of the workload for the CPU oracle: (table pairs, reads).  Same generator
    rules as spacegraphcats_b200/synth.py, done with numpy on the host."""
    from oracle import oracle as O
    from spacegraphcats_b200.synth import unitig_cuts

    rng = np.random.default_rng(seed)
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[rng.integers(0, 4, n_kmers + KSIZE - 1)]
    cuts = unitig_cuts(n_kmers, seed + 1)
    # unitigs overlap by k-1: hash the genome once, id of k-mer q = its unitig
    h = O.hash_sequence(g.tobytes(), KSIZE)
    ids = (np.searchsorted(cuts, np.arange(n_kmers), side="right") - 1).astype(np.uint32)
    # drop hashes that occur under two ids (multicounts), as build_mphf does
    order = np.argsort(h, kind="stable")
    hs, vs = h[order], ids[order]
    dup = np.zeros(len(hs), bool)
    same = hs[1:] == hs[:-1]
    diff = same & (vs[1:] != vs[:-1])
    dup[1:] |= diff
    dup[:-1] |= diff
    bad = np.unique(hs[dup])
    keep = ~np.isin(hs, bad)
    first = np.ones(len(hs), bool)
    first[1:] = ~same
    keep &= first
    table = O.COracleTable(hs[keep], vs[keep])
    pos = rng.integers(0, len(g) - READ_LEN, n_reads)
    idx = pos[:, None] + np.arange(READ_LEN)[None, :]
    reads = g[idx]
    err = rng.random(reads.shape) < ERR_PPM / 1e6
    sub = lut[rng.integers(0, 4, int(err.sum()))]
    reads[err] = sub
    rc = rng.random(n_reads) < 0.5
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    reads[rc] = comp[reads[rc]][:, ::-1]
    buf = np.ascontiguousarray(reads).reshape(-1)
    off = np.arange(n_reads + 1, dtype=np.int64) * READ_LEN
    return table, buf, off


def cpu_time_sample(table, buf, off, threads, min_seconds=10.0):
    """Run the oracle's fused hash+lookup+per-read-distinct pass until >= min_seconds."""
    nk_total, t_total, reps = 0, 0.0, 0
    while t_total < min_seconds:
        t0 = time.perf_counter()
        nk, nh, nl, cs = table.label_reads_checksum(buf, off, KSIZE, threads=threads)
        t_total += time.perf_counter() - t0
        nk_total += nk
        reps += 1
    return nk_total / t_total, nk_total, t_total, reps


def run_reference(args):
    """--impl reference: the reference's CPU path for this metric.  The reference's own natives
    (khmer, bbhash) are absent, so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g

    g.build()
    from oracle import oracle as O

    # all host cores, explicitly: torchrun exports OMP_NUM_THREADS=1 to every rank
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n_kmers, n_reads = args.cpu_kmers, args.cpu_reads
    log("[reference] building CPU table (%d unitig k-mers) and %d reads ..." % (n_kmers, n_reads))
    table, buf, off = cpu_workload(n_kmers, n_reads)
    step_kmers = n_reads * (READ_LEN - KSIZE + 1)
    for _ in range(args.warmup):
        table.label_reads_checksum(buf, off, KSIZE, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        table.label_reads_checksum(buf, off, KSIZE, threads=threads)
    dt = time.perf_counter() - t0
    value = step_kmers * args.steps / dt
    sample = ("oracle port (oracle/sgc_oracle.c, OpenMP): %d x 150-bp reads per step against a %d-k-mer "
              "synthetic cDBG table (scaled-down copy of the GPU workload)" % (n_reads, n_kmers))
    line = {
        "metric": "k-mers hashed+looked-up/sec (index_reads)", "value": value, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "impl": "reference",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cpu=False):
    return {
        "workload": "synthetic metagenome cDBG, %d unitig k-mers k=%d, 150bp reads labelled by index_reads, "
                    "1 B200 replicated table (BASELINE.json configs[3])" % (args.table_kmers, KSIZE),
        "ksize": KSIZE, "read_len": READ_LEN, "reads_per_step": args.batch_reads,
        "kmers_per_step": args.batch_reads * (READ_LEN - KSIZE + 1), "error_rate": ERR_PPM / 1e6,
        "table": "replicated",
        "label_kernel": "per-k-mer probes (SGC_SIDE=0)" if os.environ.get("SGC_SIDE", "1") == "0"
                        else "anchor probe + unitig-ordered side array (default)",
        "l2": "inputs (2.5 GB reads/step, 21 GB table) far larger than the 126 MB L2; no flush needed",
        **({"cpu_sample": {"table_kmers": args.cpu_kmers, "reads_per_step": args.cpu_reads}} if cpu else {}),
    }


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores local to GPU `index` (NVML affinity) BEFORE any pinned host
    memory is allocated, so that the staging buffers are first-touched on the GPU's own NUMA node.
    Matters for the end-to-end arm when 8 ranks share one host; harmless otherwise."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_partitioned(args, rank, world, local):
    """BASELINE.json configs[4] shape: --table-kmers unitig k-mers in TOTAL, hash-range partitioned over
    the ranks; every rank labels its own read batches through request/reply exchanges."""
    import torch
    import torch.distributed as dist

    from spacegraphcats_b200.device import Context
    from spacegraphcats_b200.parallel import DeviceEngine, PartitionedIndex, shard_range
    from spacegraphcats_b200.synth import SyntheticCdbg

    if world < 2:
        raise SystemExit("--table partitioned needs torchrun with >= 2 ranks")
    ctx = Context(local)
    t0 = time.time()
    cdbg = SyntheticCdbg(ctx, args.table_kmers, KSIZE, seed=1)  # same genome / unitigs on every rank
    ulo, uhi = shard_range(cdbg.n_unitigs, rank, world)
    with torch.cuda.stream(ctx.stream):
        b0, b1 = int(cdbg.uoff[ulo].item()), int(cdbg.uoff[uhi].item())
        useq = cdbg.useq[b0:b1]
        uoff = (cdbg.uoff[ulo:uhi + 1] - b0).contiguous()
        ids = torch.arange(ulo, uhi, dtype=torch.int32, device=ctx.device)
    p2p = args.table == "partitioned"
    idx = PartitionedIndex(DeviceEngine(ctx), p2p=p2p).build(useq, uoff, KSIZE, ids)
    ctx.sync()
    log("[rank %d] partitioned build %.1fs: %d hashes (%.2f GB) on this rank" % (rank, time.time() - t0, idx.n_local, idx.table.nbytes / 1e9))
    n_reads = args.batch_reads
    step_kmers = n_reads * (READ_LEN - KSIZE + 1)
    nb = max(1, min(2, args.steps + args.warmup))
    batches = [cdbg.reads(n_reads, READ_LEN, seed=1000 * (rank + 1) + 3 + b, err_ppm=ERR_PPM) for b in range(nb)]
    ctx.sync()
    for i in range(args.warmup):
        res = idx.label_reads(*batches[i % nb], KSIZE)
    launches0 = ctx.launches
    dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for i in range(args.steps):
            res = idx.label_reads(*batches[(args.warmup + i) % nb], KSIZE)
        ctx.sync()
        dist.barrier()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        value = world * step_kmers * args.steps / dt
        link_bytes = (world - 1) / world * 12.0 * step_kmers  # per GPU and direction (8 B out + 4 B back)
        achieved = link_bytes * args.steps / dt / 1e9
        line = {
            "metric": "k-mers hashed+looked-up/sec (index_reads)", "value": value, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "synthetic cDBG, %d unitig k-mers k=%d hash-range partitioned across %d B200 "
                                   "(BASELINE.json configs[4] shape), 150bp reads labelled by index_reads" % (args.table_kmers, KSIZE, world),
                       "ksize": KSIZE, "read_len": READ_LEN, "reads_per_rank_step": n_reads, "table": args.table,
                       "exchange": "P2P stores into peer inboxes over NVLink (CUDA IPC)" if p2p else "NCCL all-to-all"},
            "labels_per_read": int(res[1].numel()) / n_reads, "clocks": clocks.summary(), "e2e": None,
            "gpu_launches": int(ctx.launches - launches0),
            "roofline": {"bound": "nvlink", "achieved": achieved, "peak": 770.0, "unit": "GB/s", "frac": achieved / 770.0,
                         "traffic": None, "peak_source": "measured peer copy, B200_PROFILING.md",
                         "note": "(R-1)/R x 12 B per k-mer per GPU and direction over the whole step; the step is bound by "
                                 "the un-overlapped hash / probe / gather kernels, not by the link (DESIGN.md section 6)"},
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


# --------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--table-kmers", type=int, default=500_000_000)
    ap.add_argument("--batch-reads", type=int, default=1 << 24)
    ap.add_argument("--distinct-batches", type=int, default=4, help="device-resident read batches cycled through")
    ap.add_argument("--cpu-kmers", type=int, default=20_000_000)
    ap.add_argument("--cpu-reads", type=int, default=400_000)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--table", default="replicated", choices=["replicated", "partitioned", "partitioned-nccl"],
                    help="replicated (configs[3], default) or hash-range partitioned across the ranks (configs[4]; "
                         "exchange fused into the kernels over NVLink peer memory, or through NCCL all-to-all)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)

    import __graft_entry__ as g

    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from spacegraphcats_b200.device import Context
    from spacegraphcats_b200.synth import SyntheticCdbg
    from spacegraphcats_b200._lib import check

    if args.table != "replicated":
        return run_partitioned(args, rank, world, local)

    ctx = Context(local)
    L = ctx._L
    t0 = time.time()
    cdbg = SyntheticCdbg(ctx, args.table_kmers, KSIZE, seed=1)
    table = cdbg.build_table()
    live, multi, max_id = table.stats()
    log("[rank %d] cDBG: %d unitigs, table %d live hashes (%d multicount) in %.2f GB, built in %.1fs"
        % (rank, cdbg.n_unitigs, live, multi, table.nbytes / 1e9, time.time() - t0))
    nb = max(1, min(args.distinct_batches, args.steps + args.warmup))
    batches = [cdbg.reads(args.batch_reads, READ_LEN, seed=1000 * (rank + 1) + 3 + b, err_ppm=ERR_PPM) for b in range(nb)]
    ctx.sync()
    n_reads = args.batch_reads
    step_kmers = n_reads * (READ_LEN - KSIZE + 1)
    ids_cap = 6 * n_reads
    d_ids_off = ctx.empty(n_reads + 1, torch.int64)
    d_ids = ctx.empty(ids_cap, torch.int32)
    d_stat = ctx.empty(n_reads, torch.int32)

    def label(seq, off):
        nk, nh, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        check(L.sgc_label_reads(table._h, ctx.ptr(seq), seq.numel(), ctx.ptr(off), n_reads, KSIZE, ctx.ptr(d_ids_off),
                                ctx.ptr(d_ids), ids_cap, ctx.ptr(d_stat), ctypes.byref(nk), ctypes.byref(nh),
                                ctypes.byref(nl)))
        return nk.value, nh.value, nl.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs ------------------------------------------------
    args.warmup = max(args.warmup, 1)  # at least one untimed step (context / allocator / attribute set-up)
    for i in range(args.warmup):
        stats = label(*batches[i % nb])
    assert stats[0] == step_kmers, stats
    launches0 = ctx.launches
    check(L.sgc_ctx_set_timing(ctx._h, 1))
    ms_dummy, n_dummy = ctypes.c_double(0), ctypes.c_int64(0)
    check(L.sgc_ctx_tile_ms(ctx._h, ctypes.byref(ms_dummy), ctypes.byref(n_dummy)))
    barrier()
    with ClockSampler(local) as clocks:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ctx.stream)
        for i in range(args.steps):
            stats = label(*batches[(args.warmup + i) % nb])
        ev1.record(ctx.stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    gpu_launches = ctx.launches - launches0
    tile_ms, tile_n = ctypes.c_double(0), ctypes.c_int64(0)
    check(L.sgc_ctx_tile_ms(ctx._h, ctypes.byref(tile_ms), ctypes.byref(tile_n)))
    check(L.sgc_ctx_set_timing(ctx._h, 0))
    t = torch.tensor([ms], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * step_kmers * args.steps / (ms_max / 1e3)
    hit_frac = stats[1] / stats[0]
    labels_per_read = stats[2] / n_reads

    # ---- e2e: host buffers, H2D + kernels + D2H inside the timed region -----------------
    e2e = None
    if not args.no_e2e:
        nh_b = 2
        h_seq = [torch.empty(n_reads * READ_LEN, dtype=torch.uint8).pin_memory() for _ in range(nh_b)]
        h_off = [torch.empty(n_reads + 1, dtype=torch.int64).pin_memory() for _ in range(nh_b)]
        for b in range(nh_b):
            h_seq[b].copy_(batches[b % nb][0])
            h_off[b].copy_(batches[b % nb][1])
        h_ids_off = [torch.empty(n_reads + 1, dtype=torch.int64).pin_memory() for _ in range(2)]
        h_ids = [torch.empty(ids_cap, dtype=torch.int32).pin_memory() for _ in range(2)]
        d_seq = [ctx.empty(n_reads * READ_LEN, torch.uint8) for _ in range(2)]
        d_off = [ctx.empty(n_reads + 1, torch.int64) for _ in range(2)]
        o_ids_off = [ctx.empty(n_reads + 1, torch.int64) for _ in range(2)]
        o_ids = [ctx.empty(ids_cap, torch.int32) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=ctx.device)
        back_stream = torch.cuda.Stream(device=ctx.device)
        copied = [torch.cuda.Event() for _ in range(2)]
        returned = [torch.cuda.Event() for _ in range(2)]
        torch.cuda.synchronize()

        def h2d(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                d_seq[s].copy_(h_seq[i % nh_b], non_blocking=True)
                d_off[s].copy_(h_off[i % nh_b], non_blocking=True)
                copied[s].record(copy_stream)

        def label_into(s):
            nk, nh, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            check(L.sgc_label_reads(table._h, ctx.ptr(d_seq[s]), d_seq[s].numel(), ctx.ptr(d_off[s]), n_reads, KSIZE,
                                    ctx.ptr(o_ids_off[s]), ctx.ptr(o_ids[s]), ids_cap, ctx.ptr(d_stat), ctypes.byref(nk),
                                    ctypes.byref(nh), ctypes.byref(nl)))
            return nl.value

        def e2e_steps(n):
            """H2D of step i+1 and D2H of step i-1 overlap the kernels of step i."""
            h2d(0)
            tot_labels = 0
            for i in range(n):
                s = i % 2
                ctx.stream.wait_event(copied[s])
                ctx.stream.wait_event(returned[s])  # outputs of step i-2 have left this buffer
                if i + 1 < n:
                    h2d(i + 1)
                nl = label_into(s)  # returns when the kernels of step i are done (host sync inside)
                with torch.cuda.stream(back_stream):
                    h_ids_off[s].copy_(o_ids_off[s], non_blocking=True)
                    h_ids[s][:nl].copy_(o_ids[s][:nl], non_blocking=True)
                    returned[s].record(back_stream)
                tot_labels += nl
            back_stream.synchronize()
            return tot_labels

        e2e_steps(min(2, args.warmup))
        barrier()
        t0 = time.perf_counter()
        tot_labels = e2e_steps(args.steps)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {
            "value": world * step_kmers * args.steps / dt, "unit": "k-mers/s",
            "h2d_bytes_per_step": int(n_reads * READ_LEN + (n_reads + 1) * 8),
            "d2h_bytes_per_step": int((n_reads + 1) * 8 + tot_labels // args.steps * 4),
            "ms_per_step": dt / args.steps * 1e3, "pipeline": "double-buffered H2D on a copy stream",
        }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bytes_per_kmer = READ_LEN / (READ_LEN - KSIZE + 1) + SECTOR + 4.0 * labels_per_read / (READ_LEN - KSIZE + 1)
    tile_s = tile_ms.value / 1e3 / max(tile_n.value, 1)
    achieved = bytes_per_kmer * step_kmers / tile_s / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes per k-mer of this kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = float(tj["dram_bytes_per_kmer"]) * step_kmers / 1e9
        traffic_src = tj.get("source")
    except Exception:
        pass
    roofline = {
        "bound": "hbm",
        "kernel": "tile_kernel<31, MODE_LABEL>" if os.environ.get("SGC_SIDE", "1") == "0" else "tile_kernel<31, MODE_LABEL_SIDE>",
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_unit": "GB per launch", "traffic_source": traffic_src,
        "peak_source": peak_src,
        "bytes_per_kmer": bytes_per_kmer, "kernel_ms": tile_s * 1e3, "kernel_launches": int(tile_n.value),
        "kernel_share_of_step": tile_ms.value / ms,
        "note": "algorithmic bytes = read bases + ONE 32-byte table sector per k-mer + label output (SURVEY 8d); the "
                "measured random-access ceiling of HBM on this part is 36-46 G accesses/s and every access moves a "
                "128-byte line (scripts/sector_probe*.cu, profiles/); the default label kernel probes the table "
                "once per four k-mers and matches the rest in a unitig-ordered side array: 64 B of HBM per k-mer "
                "instead of 129",
    }

    # ---- CPU baseline on the host cores ---------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle as O

        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        log("[cpu_baseline] building the scaled-down CPU workload ...")
        ctab, cbuf, coff = cpu_workload(args.cpu_kmers, args.cpu_reads)
        rate, nk_tot, secs, reps = cpu_time_sample(ctab, cbuf, coff, threads, args.cpu_seconds)
        # the reference itself is single-threaded (one Python process per snakemake rule): same pass on 1 core
        n1 = max(1, args.cpu_reads // 8)
        rate1, _, _, _ = cpu_time_sample(ctab, cbuf[: n1 * READ_LEN], coff[: n1 + 1], 1, 2.0)
        cpu = {"value": rate, "unit": "k-mers/s", "cores": threads, "kind": "port", "single_thread_value": rate1,
               "sample": "%d passes over %d 150-bp reads (%.2e k-mers, %.1f s) against a %d-k-mer synthetic cDBG table; "
                         "oracle/sgc_oracle.c fused hash+lookup+per-read distinct, OpenMP" % (reps, args.cpu_reads, nk_tot, secs, args.cpu_kmers)}

    line = {
        "metric": "k-mers hashed+looked-up/sec (index_reads)", "value": value, "unit": "k-mers/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args),
        "hit_fraction": hit_frac, "labels_per_read": labels_per_read,
        "table_bytes": table.nbytes, "table_entries": live,
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(gpu_launches),
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
