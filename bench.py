#!/usr/bin/env python
"""bench.py -- k-mers hashed + looked-up per second on the index_reads hot path (+ the query half of the metric).

Workload (BASELINE.json configs[3]): synthetic metagenome cDBG with
500 M unitig k-mers (k=31) in a GPU-resident table replicated per GPU; 150-bp reads with 0.5 %
substitution errors, half reverse-complemented, labelled with the distinct unitig ids their
k-mers hit (cdbg/index_reads.py:143-152).  One STEP = one batch of 2^24 reads (2.01e9 k-mers);
the reads are held 2 bits per base (40 bytes per 150-bp read, the layout the host feeder emits) and are
expanded to ASCII inside the kernel.  The default 12 steps are the config's 200 M reads.

  value  k-mers/s, whole job, inputs resident in HBM, timed with CUDA events on the launch stream
  e2e    the same through the product's own streaming API (GpuKmerTable.label_reads_stream, what
         index_reads.main runs): pinned host batches -> H2D -> kernels -> D2H of the CSR labels,
         double-buffered; H2D / D2H bytes counted from the tensors copied
  roofline   the label kernel (hash + probe + label) alone, CUDA events recorded around its launches
  query  the other half of the metric: batched count_cdbg_matches over many query sequences
         (search/query_by_sequence.py:153-189): per-query distinct hashes -> {unitig: count}
  cpu_baseline / --impl reference   the CPU oracle (oracle/sgc_oracle.c, OpenMP over all host
         cores) on a bounded, scaled-down sample of the same workload

N > 1 (torchrun): every rank holds the full table and labels its own read batches (weak scaling,
no data-path collective); the time is the max over ranks.  The same run then ALSO builds the
hash-range partitioned table of configs[4] (--part-kmers-per-gpu unitig k-mers per GPU; requests
and replies stored straight into peer memory over NVLink), checks its labels against a replicated
table and reports it as the "partitioned" sub-object with the NVLink roofline.
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
def unitig_cuts(n_kmers, seed=2, mean=64, cap=4096):
    """Cut points (k-mer coordinates) of the synthetic unitigs: 1 + Geometric(1/64) k-mers each, capped at
    4096 (the same rule as spacegraphcats_b200/synth.py, restated here so that the CPU
    arm imports nothing of the product)."""
    rng = np.random.default_rng(seed)
    est = int(n_kmers / mean * 1.05) + 1024
    cuts = [np.zeros(1, np.int64)]
    total = 0
    while total < n_kmers:
        sizes = np.minimum(rng.geometric(1.0 / mean, est), cap).astype(np.int64)
        c = np.cumsum(sizes) + total
        cuts.append(c)
        total = int(c[-1])
    cuts = np.concatenate(cuts)
    cuts = cuts[cuts < n_kmers]
    return np.concatenate([cuts, [n_kmers]]).astype(np.int64)


def build_oracle():
    """Compile the CPU oracle only (gcc); the product library is neither built nor loaded by the CPU arm."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def cpu_workload(n_kmers, n_reads, seed=1):
    """Scaled-down copy of the workload for the CPU oracle: (table, reads, offsets, genome, cuts)."""
    from oracle import oracle as O

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
    return table, buf, off, g, len(cuts) - 1


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


def cpu_python_granularity(table, buf, off, seconds=3.0):
    """BASELINE.md's B1: the reference's own loop shape (cdbg/index_reads.py:143-152) -- one Python call per
    read to a native hash_sequence that returns a list of Python ints, one native get_unique_values over that
    list, a Python set update -- on ONE thread, the reference's execution model.  -> k-mers/s"""
    from oracle import oracle as O

    ot = table
    n = len(off) - 1
    done, nk, t0 = 0, 0, time.perf_counter()
    raw = buf.tobytes()
    while time.perf_counter() - t0 < seconds and done < n:
        seq = raw[off[done]:off[done + 1]]
        kmers = O.hash_sequence(seq, KSIZE).tolist()          # khmer get_kmer_hashes -> list of PyLongs
        counts = ot.get_unique_values_list(kmers)             # BBHashTable.get_unique_values -> dict
        ids = set()
        ids.update(counts)
        nk += len(kmers)
        done += 1
    return nk / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's CPU path for this metric.  The reference's own natives
    (khmer, bbhash) are absent, so this is the oracle port with all host threads.
    Nothing of the product is imported, built or loaded here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    build_oracle()
    # all host cores, explicitly: torchrun exports OMP_NUM_THREADS=1 to every rank
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n_kmers, n_reads = args.cpu_kmers, args.cpu_reads
    log("[reference] building CPU table (%d unitig k-mers) and %d reads ..." % (n_kmers, n_reads))
    table, buf, off, _, _ = cpu_workload(n_kmers, n_reads)
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
    cfg = workload_config(args, cpu=True)
    line = {
        "metric": "k-mers hashed+looked-up/sec (index_reads)", "value": value, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "impl": "reference",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cpu=False):
    if cpu:
        # the CPU arm runs a bounded, scaled-down sample of the same workload: say so where the driver reads it
        return {
            "workload": "synthetic metagenome cDBG k=%d, 150bp reads labelled by index_reads (BASELINE.json configs[3] "
                        "shape), CPU SAMPLE: %d unitig k-mers in the table, %d reads (%d k-mers) per step"
                        % (KSIZE, args.cpu_kmers, args.cpu_reads, args.cpu_reads * (READ_LEN - KSIZE + 1)),
            "ksize": KSIZE, "read_len": READ_LEN, "reads_per_step": args.cpu_reads,
            "kmers_per_step": args.cpu_reads * (READ_LEN - KSIZE + 1), "error_rate": ERR_PPM / 1e6,
            "table": "replicated", "table_kmers": args.cpu_kmers,
            "gpu_arm_shape": {"table_kmers": args.table_kmers, "reads_per_step": args.batch_reads},
        }
    return {
        "workload": "synthetic metagenome cDBG, %d unitig k-mers k=%d, 150bp reads labelled by index_reads, "
                    "1 B200 replicated table (BASELINE.json configs[3])" % (args.table_kmers, KSIZE),
        "ksize": KSIZE, "read_len": READ_LEN, "reads_per_step": args.batch_reads,
        "kmers_per_step": args.batch_reads * (READ_LEN - KSIZE + 1), "error_rate": ERR_PPM / 1e6,
        "table": "replicated", "table_kmers": args.table_kmers,
        "read_format": "ASCII (1 byte per base + int64 offsets)" if args.ascii else
                       "2-bit packed, 40 bytes per 150-bp read (the host feeder's output; expanded to ASCII in the kernel)",
        "label_kernel": "per-k-mer probes (SGC_NO_SIDE)" if os.environ.get("SGC_NO_SIDE") else
                        "labelseq: one hashed+probed anchor per 16 k-mers, followers settled by comparing the read's 2-bit bases with the table's unitig sequence store; k-mers past a unitig end re-anchored as a run, k-mers overlapping a mismatch hashed+probed one per lane",
        "l2": "inputs (0.67 GB packed reads/step, 27 GB table + 0.3 GB unitig sequence store) far larger than the 126 MB L2; no flush needed",
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


def partitioned_arm(args, ctx, rank, world):
    """BASELINE.json configs[4]: a cDBG of world x --part-kmers-per-gpu unitig k-mers, hash-range partitioned over
    the ranks (every rank owns one range); every rank labels its own read batches, requests and replies
    stored straight into the owners' / requesters' memory over NVLink.  The labels of one batch are checked
    against a REPLICATED table of the same cDBG built on every rank.  -> dict (rank 0) / None"""
    import torch
    import torch.distributed as dist

    from spacegraphcats_b200.device import DeviceTable
    from spacegraphcats_b200.parallel import DeviceEngine, PartitionedIndex, shard_range
    from spacegraphcats_b200.synth import SyntheticCdbg

    total_kmers = args.part_kmers_per_gpu * world
    t0 = time.time()
    cdbg = SyntheticCdbg(ctx, total_kmers, KSIZE, seed=1)  # same genome / unitigs on every rank
    ulo, uhi = shard_range(cdbg.n_unitigs, rank, world)
    with torch.cuda.stream(ctx.stream):
        b0, b1 = int(cdbg.uoff[ulo].item()), int(cdbg.uoff[uhi].item())
        useq = cdbg.useq[b0:b1]
        uoff = (cdbg.uoff[ulo:uhi + 1] - b0).contiguous()
        ids = torch.arange(ulo, uhi, dtype=torch.int32, device=ctx.device)
    idx = PartitionedIndex(DeviceEngine(ctx), p2p=True, peer_probe=not args.part_exchange).build(useq, uoff, KSIZE, ids)
    ctx.sync()
    del useq, uoff, ids
    torch.cuda.empty_cache()
    log("[rank %d] partitioned build %.1fs: %d of %d hashes (%.2f GB) on this rank"
        % (rank, time.time() - t0, idx.n_local, total_kmers, idx.table.nbytes / 1e9))
    n_reads = args.part_reads
    step_kmers = n_reads * (READ_LEN - KSIZE + 1)
    nb = 2
    batches = [cdbg.reads(n_reads, READ_LEN, seed=1000 * (rank + 1) + 3 + b, err_ppm=ERR_PPM) for b in range(nb)]
    ctx.sync()
    warm = max(args.warmup, 1)
    res = None
    for i in range(warm):
        res = idx.label_reads(*batches[i % nb], KSIZE)
    check_batch = (warm - 1) % nb
    launches0 = ctx.launches
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.part_steps):
        idx.label_reads(*batches[(warm + i) % nb], KSIZE)
    ctx.sync()
    dist.barrier()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = ctx.launches - launches0
    t = torch.tensor([dt], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    # parity: the same batch against a replicated table of the whole cDBG on this GPU (per-k-mer probing; packed
    # denser than the working tables so that the 8-GPU case, 4e9 hashes, fits next to the shard)
    ok = None
    if not args.no_part_check:
        t1 = time.time()
        torch.cuda.empty_cache()
        rep = DeviceTable(ctx, cdbg.total_unitig_kmers, load=1.3)
        rep.insert_sequences(cdbg.useq, cdbg.uoff, KSIZE)
        ref = rep.label_reads(*batches[check_batch], KSIZE, to_host=False)
        same = torch.equal(ref["ids_off"], res[0]) and torch.equal(ref["ids"], res[1])
        okt = torch.tensor([1 if same else 0], device=ctx.device)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item())
        rep.close()
        log("[rank %d] partitioned labels == replicated labels: %s (check %.1fs)" % (rank, same, time.time() - t1))
    labels_per_read = int(res[1].numel()) / n_reads
    table_bytes = idx.table.nbytes
    idx.close()
    del cdbg
    if rank != 0:
        return None
    value = world * step_kmers * args.part_steps / dt
    link_bytes = (world - 1) / world * 12.0 * step_kmers  # per GPU and direction (8 B request + 4 B reply)
    achieved = link_bytes * args.part_steps / dt / 1e9
    return {
        "value": value, "unit": "k-mers/s", "per_gpu": value / world, "ms_per_step": dt / args.part_steps * 1e3,
        "steps": args.part_steps, "table_kmers": total_kmers, "table_kmers_per_gpu": args.part_kmers_per_gpu,
        "table_bytes_per_gpu": table_bytes, "reads_per_rank_step": n_reads, "labels_per_read": labels_per_read,
        "labels_identical_to_replicated": ok, "gpu_launches": int(launches),
        "exchange": ("requests / replies stored straight into peer memory over NVLink (CUDA IPC), no NCCL on the data path"
                     if args.part_exchange else
                     "none: the label kernel loads each home bucket (32 B) straight from its owner GPU's table over NVLink "
                     "(bucket arrays shared with cuMemCreate + POSIX fds); one k-mer in sixteen + the unsettled ones are "
                     "probed, the rest is compared with the replicated unitig sequence store; singly probed k-mers owned by another "
                     "GPU are first looked up in a replicated miss filter in local HBM (16 bits per hash; no false negatives); no "
                     "collective in the label path"),
        "workload": "synthetic cDBG, %d unitig k-mers k=%d hash-range partitioned across %d B200 (BASELINE.json configs[4]), "
                    "150bp reads labelled by index_reads" % (total_kmers, KSIZE, world),
        "roofline": {"bound": "nvlink", "achieved": achieved, "peak": 770.0, "unit": "GB/s", "frac": achieved / 770.0,
                     "peak_source": "measured peer copy, B200_PROFILING.md",
                     "note": "achieved = SURVEY 8d's algorithmic figure, (R-1)/R x 12 B per k-mer per GPU and direction, over "
                             "the whole step (73 G k-mers/s per GPU at R=8).  What the link really carries here is one 32-byte "
                             "load per k-mer that is probed remotely (~0.1 per k-mer with the miss filter); a B200 serves 6.8 G random 32-byte loads/s from "
                             "another process's memory over NVLink (scripts/peer_probe.cu, profiles/r02_peer_probe_microbench.txt)"},
    }


def query_arm(args, ctx, table, cdbg, rank, world):
    """The count_cdbg_matches half of the metric (BASELINE.json configs[2] shape, scaled up): --queries query
    sequences of --query-kmers k-mers each (the acido-chunk size), cut from the cDBG's genome with the read error
    model, answered as ONE batch per step by sgc_query_count_batch: per query its SET of k-mer hashes ->
    {unitig: count} (search/query_by_sequence.py:153-189).  Every rank runs its own queries (weak scaling)."""
    import torch
    import torch.distributed as dist

    L = ctx._L
    dev = table._dev
    nq, qlen = args.queries, args.query_kmers + KSIZE - 1
    step_kmers = nq * args.query_kmers
    nb = 2
    batches = [cdbg.reads(nq, qlen, seed=7000 * (rank + 1) + b, err_ppm=ERR_PPM) for b in range(nb)]
    d_sq = ctx.to_device(np.arange(nq, dtype=np.int32))
    d_qd = ctx.empty(nq, torch.int64)
    cap = max(step_kmers // 16, 1 << 16)
    d_keys, d_cnt = ctx.empty(cap, torch.int64), ctx.empty(cap, torch.int32)
    ctx.sync()

    def run(seq, off):
        npairs, nk, nh = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        from spacegraphcats_b200._lib import check

        check(L.sgc_query_count_batch(dev._h, ctx.ptr(seq), seq.numel(), ctx.ptr(off), nq, ctx.ptr(d_sq), nq, KSIZE,
                                      ctx.ptr(d_keys), ctx.ptr(d_cnt), cap, ctx.ptr(d_qd), None, ctypes.byref(npairs),
                                      ctypes.byref(nk), ctypes.byref(nh)))
        return npairs.value, nk.value, nh.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(2):
        stats = run(*batches[i % nb])
    assert stats[1] == step_kmers, stats
    launches0 = ctx.launches
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ctx.stream)
    for i in range(args.query_steps):
        stats = run(*batches[i % nb])
    ev1.record(ctx.stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * step_kmers * args.query_steps / (ms / 1e3)
    distinct = int(d_qd.sum().item())

    # N >= 2: the same steps with the sparse results all-gathered, so that every rank ends up with the answers of
    # ALL ranks' queries (query batches shard across the ranks; parallel.count_matches_batch_sharded is
    # the host-level form of this step)
    sharded = None
    if world > 1:
        from spacegraphcats_b200.parallel import allgather_var

        def sharded_step(i):
            npairs, _, _ = run(*batches[i % nb])
            with torch.cuda.stream(ctx.stream):
                ks = allgather_var(d_keys[:npairs], None)
                cs = allgather_var(d_cnt[:npairs], None)
                qd = [torch.empty_like(d_qd) for _ in range(world)]
                dist.all_gather(qd, d_qd)
            return sum(int(k.numel()) for k in ks), cs, qd

        sharded_step(0)
        barrier()
        ev0.record(ctx.stream)
        for i in range(args.query_steps):
            tot_pairs_all, _, _ = sharded_step(i)
        ev1.record(ctx.stream)
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sms = float(t.item())
        sharded = {"value": world * step_kmers * args.query_steps / (sms / 1e3), "unit": "k-mers/s",
                   "ms_per_step": sms / args.query_steps, "pairs_gathered_per_step": int(tot_pairs_all),
                   "collective": "all_gather (NCCL) of every rank's (query << 32 | unitig, count) pairs and per-query "
                                 "distinct counts; no collective on the lookup path (replicated table)"}

    # e2e: pinned host sequences -> H2D -> the batch call -> D2H of the (query, unitig, count) triples
    h_seq = [torch.empty(nq * qlen, dtype=torch.uint8).pin_memory() for _ in range(nb)]
    h_off = torch.empty(nq + 1, dtype=torch.int64).pin_memory()
    for b in range(nb):
        h_seq[b].copy_(batches[b][0])
    h_off.copy_(batches[0][1])
    torch.cuda.synchronize()

    def e2e_steps(n):
        tot = 0
        for i in range(n):
            res = dev.count_matches_batch(h_seq[i % nb], h_off, np.arange(nq, dtype=np.int32), nq, KSIZE)
            tot += len(res["keys"])
        return tot

    e2e_steps(1)
    barrier()
    t0 = time.perf_counter()
    tot_pairs = e2e_steps(args.query_steps)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank != 0:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_kmer = qlen / args.query_kmers + SECTOR + 4.0  # sequence in + one table sector + the id
    achieved = bytes_per_kmer * step_kmers * args.query_steps / (ms / 1e3) / 1e9
    out = {
        "metric": "k-mers hashed+looked-up/sec (count_cdbg_matches)", "value": value, "unit": "k-mers/s",
        "ms_per_step": ms / args.query_steps, "steps": args.query_steps,
        "workload": "%d query sequences x %d k-mers (k=%d, acido-chunk shape of BASELINE.json configs[2], cut from the "
                    "synthetic cDBG's genome with 0.5 %% errors) per step against the %d-k-mer table; per query the SET "
                    "of its k-mer hashes -> {unitig: count}" % (nq, args.query_kmers, KSIZE, args.table_kmers),
        "queries_per_step": nq, "kmers_per_step": step_kmers, "hit_fraction": stats[2] / max(distinct, 1),
        "distinct_fraction": distinct / step_kmers, "pairs_per_query": stats[0] / nq, "gpu_launches": int(launches),
        "sharded_with_gather": sharded,
        "e2e": {"value": world * step_kmers * args.query_steps / dt, "unit": "k-mers/s", "ms_per_step": dt / args.query_steps * 1e3,
                "h2d_bytes_per_step": int(nq * qlen + (nq + 1) * 8 + nq * 4),
                "d2h_bytes_per_step": int(tot_pairs // args.query_steps * 12 + nq * 8),
                "api": "DeviceTable.count_matches_batch (what search.query_by_sequence.query_batch calls)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "bytes_per_kmer": bytes_per_kmer,
                     "note": "whole batch call against sequence-in + one 32-byte sector + 4 B id per k-mer.  The label "
                             "kernel's runs give per-query INTERVALS of unitig-store positions (16 B per settled run, not "
                             "8 B per k-mer) and the hashes of the missing k-mers; distinct matches = union of the "
                             "intervals (sort, running max, reduce by (query, unitig)), missing k-mers de-duplicated by one "
                             "32-bit-key sort.  Kernel ~2.2 ms of the ~6 ms step; the rest is the two sorts and the plan"},
    }
    if not args.no_cpu and world == 1:
        # CPU arm of the query half: the oracle's batched Query restatement (C, one query per thread) on a
        # bounded sample against the scaled-down CPU table; also a result check of that sample's shape
        build_oracle()
        from oracle import oracle as O

        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        ctab, _, _, g, _ = cpu_workload(args.cpu_kmers, 1000)
        rng = np.random.default_rng(11)
        cq = max(threads, 16)
        pos = rng.integers(0, len(g) - qlen, cq)
        seqs = np.stack([g[p:p + qlen] for p in pos])
        err = rng.random(seqs.shape) < ERR_PPM / 1e6
        seqs[err] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(err.sum()))]
        buf = np.ascontiguousarray(seqs).reshape(-1)
        off = np.arange(cq + 1, dtype=np.int64) * qlen
        qseq = np.arange(cq + 1, dtype=np.int64)
        t0 = time.perf_counter()
        reps, nk_tot = 0, 0
        while time.perf_counter() - t0 < min(args.cpu_seconds, 8.0):
            _, _, nk = ctab.query_counts(buf, off, qseq, KSIZE, threads=threads)
            nk_tot += nk
            reps += 1
        secs = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nk_tot / secs, "unit": "k-mers/s", "cores": threads, "kind": "port",
                               "sample": "%d passes over %d queries x %d k-mers against a %d-k-mer table; "
                                         "oracle/sgc_oracle.c sgc_oracle_query_counts (hash, sort, unique, lookup, count; "
                                         "one query per thread)" % (reps, cq, args.query_kmers, args.cpu_kmers)}
    return out


def source_digest():
    """sha256 over the kernel sources: ties profiles/ncu_traffic.json to the code it was captured from."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "spacegraphcats_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".hpp")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


# --------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--table-kmers", type=int, default=500_000_000)
    ap.add_argument("--batch-reads", type=int, default=1 << 24)
    ap.add_argument("--distinct-batches", type=int, default=4, help="read batches cycled through")
    ap.add_argument("--cpu-kmers", type=int, default=20_000_000)
    ap.add_argument("--cpu-reads", type=int, default=400_000)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ascii", action="store_true", help="feed ASCII reads + int64 offsets instead of 2-bit packed records")
    ap.add_argument("--part-kmers-per-gpu", type=int, default=500_000_000,
                    help="N >= 2: unitig k-mers per GPU of the hash-range partitioned table (configs[4]: 4e9 over 8 GPUs)")
    ap.add_argument("--part-reads", type=int, default=1 << 22, help="reads per rank and step of the partitioned arm")
    ap.add_argument("--part-steps", type=int, default=10)
    ap.add_argument("--no-partitioned", action="store_true")
    ap.add_argument("--part-exchange", action="store_true",
                    help="partitioned arm through the request / reply exchange of round 1 instead of peer probing")
    ap.add_argument("--no-part-check", action="store_true")
    ap.add_argument("--query-kmers", type=int, default=500_000, help="k-mers per query sequence (acido-chunk shape)")
    ap.add_argument("--queries", type=int, default=512, help="query sequences per step of the query arm")
    ap.add_argument("--query-steps", type=int, default=10)
    ap.add_argument("--no-query", action="store_true")
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
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable
    from spacegraphcats_b200.device import Context
    from spacegraphcats_b200.synth import SyntheticCdbg
    from spacegraphcats_b200._lib import check

    ctx = Context(local)
    L = ctx._L
    t0 = time.time()
    cdbg = SyntheticCdbg(ctx, args.table_kmers, KSIZE, seed=1)
    # build_mphf (cdbg/index_cdbg_by_kmer.py:27-115) on the device: hash every unitig k-mer, insert with the multicount
    # rule, record store positions + break bits; timed once (table memory is cleared inside)
    ctx.sync()
    tb0 = time.perf_counter()
    dev_table = cdbg.build_table()
    ctx.sync()
    build_s = time.perf_counter() - tb0
    build = {"metric": "unitig k-mers hashed+inserted/sec (build_mphf: index_cdbg_by_kmer)",
             "value": cdbg.total_unitig_kmers / build_s, "unit": "k-mers/s", "seconds": build_s,
             "kmers": int(cdbg.total_unitig_kmers), "unitigs": int(cdbg.n_unitigs),
             "note": "one sgc_table_insert_sequences call (both reference passes fused: hash, 128-bit CAS insert, "
                     "multicount rule, sequence store + break bits), table allocation and clearing included; "
                     "wall clock around a synchronised call, single run"}
    table = GpuKmerTable(ctx)
    table._dev = dev_table  # the product object whose streaming API the e2e arm calls
    live, multi, max_id = dev_table.stats()
    TABLE_BYTES.append(dev_table.nbytes + dev_table.side_nbytes)
    HAS_STORE[0] = dev_table.side_nbytes > 0 and not os.environ.get("SGC_NO_SIDE")
    log("[rank %d] cDBG: %d unitigs, table %d live hashes (%d multicount) in %.2f GB (+%.2f GB unitig sequence store), built in %.1fs"
        % (rank, cdbg.n_unitigs, live, multi, dev_table.nbytes / 1e9, dev_table.side_nbytes / 1e9, time.time() - t0))
    nb = max(1, min(args.distinct_batches, args.steps + args.warmup))
    n_reads = args.batch_reads
    step_kmers = n_reads * (READ_LEN - KSIZE + 1)
    ids_cap = 6 * n_reads
    # device-resident batches: ASCII reads from the generator, packed 2 bits per base on the device
    batches = []
    for b in range(nb):
        seq, off = cdbg.reads(n_reads, READ_LEN, seed=1000 * (rank + 1) + 3 + b, err_ppm=ERR_PPM)
        if args.ascii:
            batches.append((seq, off))
        else:
            pk, ln, st = ctx.pack_reads(seq, off)
            assert not bool(st.any().item())
            batches.append((pk.clone(), None))
            del seq, off, pk, ln, st
    ctx.sync()
    torch.cuda.empty_cache()
    d_ids_off = ctx.empty(n_reads + 1, torch.int64)
    d_ids = ctx.empty(ids_cap, torch.int32)
    d_stat = ctx.empty(n_reads, torch.int32)

    def label(batch):
        nk, nh, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        if args.ascii:
            seq, off = batch
            check(L.sgc_label_reads(dev_table._h, ctx.ptr(seq), seq.numel(), ctx.ptr(off), n_reads, KSIZE, ctx.ptr(d_ids_off),
                                    ctx.ptr(d_ids), ids_cap, ctx.ptr(d_stat), ctypes.byref(nk), ctypes.byref(nh),
                                    ctypes.byref(nl)))
        else:
            pk = batch[0]
            check(L.sgc_label_reads_packed(dev_table._h, ctx.ptr(pk), pk.numel(), None, READ_LEN, n_reads, KSIZE,
                                           ctx.ptr(d_ids_off), ctx.ptr(d_ids), ids_cap, ctypes.byref(nk), ctypes.byref(nh),
                                           ctypes.byref(nl)))
        return nk.value, nh.value, nl.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs ------------------------------------------------
    args.warmup = max(args.warmup, 1)  # at least one untimed step (context / allocator / attribute set-up)
    for i in range(args.warmup):
        stats = label(batches[i % nb])
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
            stats = label(batches[(args.warmup + i) % nb])
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

    # ---- e2e: pinned host batches through the product's streaming API ---------------------------
    e2e = None
    if not args.no_e2e:
        nh_b = min(2, nb)
        host_batches = []
        for b in range(nh_b):
            if args.ascii:
                hs = torch.empty(n_reads * READ_LEN, dtype=torch.uint8).pin_memory()
                ho = torch.empty(n_reads + 1, dtype=torch.int64).pin_memory()
                hs.copy_(batches[b][0])
                ho.copy_(batches[b][1])
                host_batches.append({"seq": hs, "off": ho})
            else:
                hp = torch.empty(batches[b][0].numel(), dtype=torch.uint8).pin_memory()
                hp.copy_(batches[b][0])
                host_batches.append({"packed": hp, "nseq": n_reads, "lengths": None, "uniform_len": READ_LEN})
        torch.cuda.synchronize()
        h2d = (n_reads * READ_LEN + (n_reads + 1) * 8) if args.ascii else int(host_batches[0]["packed"].numel())

        def e2e_steps(n):
            tot_labels, tot_kmers, acc = 0, 0, 0
            gen = (host_batches[i % nh_b] for i in range(n))
            for _, res in table.label_reads_stream(gen, KSIZE, ids_per_read=6, prefetch=False):
                tot_labels += res["n_labels"]
                tot_kmers += res["n_kmers"]
                acc ^= int(res["ids_off"][-1]) ^ int(res["ids"][-1])  # the host reads the result it was given
            return tot_labels, tot_kmers

        e2e_steps(min(3, args.warmup + 1))
        barrier()
        t0 = time.perf_counter()
        tot_labels, tot_kmers = e2e_steps(args.steps)
        barrier()
        dt = time.perf_counter() - t0
        assert tot_kmers == step_kmers * args.steps
        t = torch.tensor([dt], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {
            "value": world * step_kmers * args.steps / dt, "unit": "k-mers/s",
            "h2d_bytes_per_step": int(h2d),
            # the CSR offsets return as one byte per read (rebuilt on the host by sgc_counts8_to_offsets) + the ids
            "d2h_bytes_per_step": int((n_reads + 1) + tot_labels // args.steps * 4),
            "ms_per_step": dt / args.steps * 1e3,
            "api": "GpuKmerTable.label_reads_stream (pinned double buffers, H2D / kernels / D2H of consecutive batches "
                   "overlapped; the loop index_reads.main runs)",
        }
        del host_batches

    # ---- the other half of the metric: batched count_cdbg_matches --------------------------------
    query = None
    if not args.no_query:
        try:
            query = query_arm(args, ctx, table, cdbg, rank, world)
        except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
            log("[query arm failed] %r" % (e,))
            query = {"error": repr(e)}

    # ---- configs[4]: hash-range partitioned table (N >= 2), after the replicated table has been freed ----
    partitioned = None
    if world > 1 and not args.no_partitioned:
        del batches, d_ids, d_ids_off, d_stat
        table._dev = None
        dev_table.close()
        del cdbg
        torch.cuda.empty_cache()
        try:
            partitioned = partitioned_arm(args, ctx, rank, world)
        except Exception as e:  # noqa: BLE001
            log("[rank %d] partitioned arm failed: %r" % (rank, e))
            partitioned = {"error": repr(e)}

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
    nkr = READ_LEN - KSIZE + 1
    # SURVEY 8(d)'s per-unit figure: L/(L-k+1) B of sequence + ONE 32-byte table sector + 4 B of id per k-mer
    survey_bytes = READ_LEN / nkr + SECTOR + 4.0
    # what this launch strictly needs: the read as it is stored + one sector + the labels it writes
    in_bytes = READ_LEN / nkr if args.ascii else 40.0 / nkr
    strict_bytes = in_bytes + SECTOR + 4.0 * labels_per_read / nkr
    tile_s = tile_ms.value / 1e3 / max(tile_n.value, 1)
    achieved = survey_bytes * step_kmers / tile_s / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes per k-mer of this kernel from the committed ncu --set full capture OF THESE SOURCES
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tj.get("source_digest") == source_digest():
            traffic = float(tj["dram_bytes_per_kmer"]) * step_kmers / 1e9
            traffic_src = tj.get("source")
        else:
            traffic_src = "profiles/ncu_traffic.json was captured from other kernel sources (digest %s != %s): not quoted" % (
                tj.get("source_digest"), source_digest())
    except Exception:
        pass
    roofline = {
        "bound": "hbm",
        "kernel": "tile_kernel<31, MODE_LABEL>" if os.environ.get("SGC_NO_SIDE") else
                  ("labelseq_kernel<31, ascii>" if args.ascii else "labelseq_kernel<31, packed>"),
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_unit": "GB per launch", "traffic_source": traffic_src,
        "peak_source": peak_src,
        "bytes_per_kmer": survey_bytes, "kernel_ms": tile_s * 1e3, "kernel_launches": int(tile_n.value),
        "kernel_kmers_per_s": step_kmers / tile_s,
        "kernel_share_of_step": tile_ms.value / ms,
        "frac_strict": strict_bytes * step_kmers / tile_s / 1e9 / peak, "bytes_per_kmer_strict": strict_bytes,
        "note": "achieved = SURVEY 8(d)'s algorithmic figure (150/120 B of sequence + ONE 32-byte table sector + 4 B of "
                "id per k-mer = 37.25 B) x k-mers per launch / measured kernel time; frac_strict counts only what this "
                "launch needs (packed read bytes + one sector + the labels written).  HBM on this part serves 36-46 G "
                "random lines/s and moves 128 bytes per access (scripts/sector_probe*.cu, profiles/); the kernel hashes "
                "and probes one k-mer in sixteen, the first k-mer past a unitig end, and the k-mers that overlap a "
                "sequencing error; everything else is settled against the unitig sequence store: ~0.2 table lines per k-mer",
    }

    # ---- CPU baseline on the host cores ---------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        log("[cpu_baseline] building the scaled-down CPU workload ...")
        build_oracle()
        ctab, cbuf, coff, _, _ = cpu_workload(args.cpu_kmers, args.cpu_reads)
        rate, nk_tot, secs, reps = cpu_time_sample(ctab, cbuf, coff, threads, args.cpu_seconds)
        # the reference itself is single-threaded (one Python process per snakemake rule): same pass on 1 core
        n1 = max(1, args.cpu_reads // 8)
        rate1, _, _, _ = cpu_time_sample(ctab, cbuf[: n1 * READ_LEN], coff[: n1 + 1], 1, 2.0)
        rate_py = cpu_python_granularity(ctab, cbuf, coff, 3.0)
        cpu = {"value": rate, "unit": "k-mers/s", "cores": threads, "kind": "port", "single_thread_value": rate1,
               "reference_python_value": rate_py,
               "reference_python_note": "1 thread, the reference's own loop shape (cdbg/index_reads.py:143-152): per read "
                                        "one native hash call returning a list of Python ints + one native "
                                        "get_unique_values over it + a set update",
               "sample": "%d passes over %d 150-bp reads (%.2e k-mers, %.1f s) against a %d-k-mer synthetic cDBG table; "
                         "oracle/sgc_oracle.c fused hash+lookup+per-read distinct, OpenMP" % (reps, args.cpu_reads, nk_tot, secs, args.cpu_kmers)}

    line = {
        "metric": "k-mers hashed+looked-up/sec (index_reads)", "value": value, "unit": "k-mers/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args),
        "hit_fraction": hit_frac, "labels_per_read": labels_per_read,
        "table_bytes": dev_table_bytes(live), "table_entries": live,
        "build": build,
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(gpu_launches),
        "roofline": roofline, "cpu_baseline": cpu, "query": query, "partitioned": partitioned,
    }
    if not HAS_STORE[0]:  # e.g. more than 2^30 - 2 store positions per table: the per-k-mer kernel ran, say so
        line["config"]["label_kernel"] = ("per-k-mer probes: this table carries no unitig sequence store (more than 2^30 "
                                          "positions, SGC_NO_SIDE, or a table filled hash by hash)")
        line["roofline"]["kernel"] = "tile_kernel<31, MODE_LABEL>"
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


HAS_STORE = [True]


def dev_table_bytes(live):
    return int(TABLE_BYTES[0]) if TABLE_BYTES else None


TABLE_BYTES = []


if __name__ == "__main__":
    main()
