"""torchrun helper (tests/test_gpu_parity.py): query_batch(distributed=True) on R ranks with replicated tables --
every rank must end up with the answers of ALL queries, equal to the single-process batch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from spacegraphcats_b200.cdbg.hashing import GpuKmerTable, MPHF_KmerIndex
    from spacegraphcats_b200.device import get_context, pack_sequences
    from spacegraphcats_b200.parallel import abundance_sharded
    from spacegraphcats_b200.search.query_by_sequence import query_batch

    ctx = get_context(local)
    k = 31
    rng = np.random.default_rng(5)
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[rng.integers(0, 4, 200_000)].tobytes().decode()
    useqs = [g[a: a + 120 + k - 1] for a in range(0, len(g) - 160, 120)]
    ubuf, uoff = pack_sequences(useqs)
    table = GpuKmerTable.from_sequences(ubuf, uoff, k, ctx=ctx)
    idx = MPHF_KmerIndex(k, table, {i: len(s) - k + 1 for i, s in enumerate(useqs)})
    tmp = sys.argv[1]
    qfiles = []
    for i in range(7):
        p = os.path.join(tmp, "q%d.rank%d.fa" % (i, dist.get_rank()))
        a = int(rng.integers(0, len(g) - 20_000))
        with open(p, "w") as fp:
            fp.write(">q%d\n%s\n>q%db\n%s\n" % (i, g[a: a + 5000 * (i + 1) // 2], i, g[a + 100: a + 900]))
        qfiles.append(p)
    one = query_batch(qfiles, k, None, idx)
    many = query_batch(qfiles, k, None, idx, distributed=True)
    same = all(a[0].cdbg_match_counts == b[0].cdbg_match_counts and a[0].n_kmers == b[0].n_kmers for a, b in zip(one, many))
    same = same and len(one) == len(many) == 7 and any(len(a[0].cdbg_match_counts) > 5 for a in one)

    # per-unitig abundance of a read set: shards looked up per rank, count vectors all-reduced
    reads = [g[int(p): int(p) + 150] for p in rng.integers(0, len(g) - 150, 2000)]
    rbuf, roff = pack_sequences(reads)

    def lookup_counts(buf, off, ksize, n_ids):
        r = table.lookup_sequences(buf, off, ksize, n_ids=n_ids, want_ids=False)
        return r["count_by_id"], r["n_kmers"], r["n_hits"]

    cnt, nk, nh = abundance_sharded(lookup_counts, rbuf, roff, k, len(useqs), device=ctx.device)
    w = lookup_counts(rbuf, roff, k, len(useqs))
    same = same and np.array_equal(cnt, w[0]) and (nk, nh) == (w[1], w[2]) and nh > 0
    ok = torch.tensor([1 if same else 0], device=ctx.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if dist.get_rank() == 0:
        print("sharded queries identical to the single-process batch on every rank: %s" % bool(ok.item()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
