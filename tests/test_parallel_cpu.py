"""world_size-2 gloo tests of the multi-GPU host logic (spacegraphcats_b200/parallel.py): the
hash-range partition / all-to-all / un-permute protocol and the read sharding, run on CPU with a
host engine built on the oracle (test infrastructure), compared with the single-process oracle."""
import contextlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class HostEngine:
    """Same interface as parallel.DeviceEngine, computed with numpy + the oracle on CPU tensors."""

    def __init__(self):
        from oracle import oracle as O

        self.O = O

    def _owner(self, h, world):
        bits = world.bit_length() - 1
        return (h >> np.uint64(64 - bits)).astype(np.int64) if bits else np.zeros(len(h), np.int64)

    def hash_pairs(self, seq, off, ksize, ids):
        h, koff = self.O.hash_batch(np.asarray(seq), np.asarray(off), ksize)
        per = np.repeat(np.asarray(ids, np.uint32), np.diff(koff))
        return torch.from_numpy(h.view(np.int64)), torch.from_numpy(per.view(np.int32))

    def partition_pairs(self, d_hash, d_ids, world):
        h = d_hash.numpy().view(np.uint64)
        own = self._owner(h, world)
        order = np.argsort(own, kind="stable")
        boff = np.searchsorted(own[order], np.arange(world + 1)).tolist()
        return d_hash[torch.from_numpy(order)], d_ids[torch.from_numpy(order)], boff

    def new_table(self, n, home_shift):
        return {"pairs": None}

    def insert_pairs(self, table, d_hash, d_ids):
        h, v = d_hash.numpy().view(np.uint64), d_ids.numpy().view(np.uint32)
        # build_mphf semantics: a hash under two different ids is dropped
        o = np.lexsort((v, h))
        h, v = h[o], v[o]
        keep = np.ones(len(h), bool)
        same = h[1:] == h[:-1]
        diff = same & (v[1:] != v[:-1])
        bad = np.unique(np.concatenate([h[1:][diff], h[:-1][diff]]))
        keep &= ~np.isin(h, bad)
        keep[1:] &= ~same
        table["t"] = self.O.COracleTable(h[keep], v[keep])

    def hash_partition(self, seq, off, ksize, world):
        h, koff = self.O.hash_batch(np.asarray(seq), np.asarray(off), ksize)
        own = self._owner(h, world)
        order = np.argsort(own, kind="stable")
        boff = np.searchsorted(own[order], np.arange(world + 1)).tolist()
        return (torch.from_numpy(h[order].view(np.int64)), torch.from_numpy(order.astype(np.int64)), boff,
                torch.from_numpy(koff), len(h))

    def lookup(self, table, d_hash):
        if "t" not in table:
            return torch.full((d_hash.numel(),), -1, dtype=torch.int32)
        return torch.from_numpy(table["t"].lookup(d_hash.numpy().view(np.uint64)).view(np.int32))

    def unpermute(self, d_ids, d_perm):
        out = torch.empty_like(d_ids)
        out[d_perm] = d_ids
        return out

    def labels_from_ids(self, d_kmer_ids, d_koff, nseq):
        ids = d_kmer_ids.numpy().view(np.uint32)
        koff = d_koff.numpy()
        ioff, out = [0], []
        for s in range(nseq):
            v = ids[koff[s] : koff[s + 1]]
            u = np.unique(v[v != 0xFFFFFFFF])
            out.append(u)
            ioff.append(ioff[-1] + len(u))
        return (torch.tensor(ioff), torch.from_numpy(np.concatenate(out).astype(np.uint32).view(np.int32)) if out else torch.zeros(0, dtype=torch.int32),
                int((ids != 0xFFFFFFFF).sum()))

    def sync(self):
        pass

    def stream(self):
        return contextlib.nullcontext()


def _workload():
    from oracle import oracle as O

    rng = np.random.default_rng(17)
    k = 31
    lut = np.frombuffer(b"ACGT", np.uint8)
    g = lut[rng.integers(0, 4, 60_000)]
    cuts = np.concatenate([[0], np.cumsum(rng.integers(1, 150, 1500))])
    cuts = cuts[cuts < len(g) - k]
    useqs = [g[a : b + k - 1].tobytes() for a, b in zip(cuts[:-1], cuts[1:])]
    useqs[7] = useqs[3]  # a duplicated unitig: its k-mers are multicounts and must vanish
    reads = []
    for i in range(600):
        p = int(rng.integers(0, len(g) - 200))
        r = g[p : p + int(rng.integers(20, 180))].copy()
        e = rng.random(len(r)) < 0.01
        r[e] = lut[rng.integers(0, 4, int(e.sum()))]
        reads.append(O.revcomp(r.tobytes()) if i & 1 else r.tobytes())
    return k, useqs, reads


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from spacegraphcats_b200.parallel import PartitionedIndex, shard_range

        k, useqs, reads = _workload()
        ulo, uhi = shard_range(len(useqs), rank, world)
        ubuf, uoff = O.pack_sequences(useqs[ulo:uhi])
        idx = PartitionedIndex(HostEngine()).build(ubuf, uoff, k, np.arange(ulo, uhi, dtype=np.uint32))
        rlo, rhi = shard_range(len(reads), rank, world)
        rbuf, roff = O.pack_sequences(reads[rlo:rhi])
        ioff, ids, nh = idx.label_reads(rbuf, roff, k)
        q.put((rank, rlo, rhi, ioff.numpy().tolist(), ids.numpy().view(np.uint32).tolist(), nh, idx.n_local))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(180)
def test_partitioned_index_world2_gloo_matches_single_process_oracle():
    from oracle import oracle as O

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0

    k, useqs, reads = _workload()
    otab, _, n_multi = O.build_mphf(k, lambda: iter(O.Record(str(i), s.decode()) for i, s in enumerate(useqs)))
    assert n_multi > 0
    ot = O.COracleTable(otab.mphf_to_hash, otab.mphf_to_value)
    assert sum(r[6] for r in results) >= len(otab)  # every live hash landed on exactly one owner (+ dropped ones)
    buf, off = O.pack_sequences(reads)
    w_off, w_ids, w_nk, w_nh = ot.label_reads(buf, off, k)
    got_ids, got_counts, got_hits = [], [], 0
    for rank, rlo, rhi, ioff, ids, nh, _ in results:
        got_counts.extend(np.diff(ioff).tolist())
        got_ids.extend(ids)
        got_hits += nh
    assert got_counts == np.diff(w_off).tolist()
    assert got_ids == w_ids.tolist()
    assert got_hits == w_nh


def test_shard_range_covers_everything():
    from spacegraphcats_b200.parallel import shard_range, log2_ranks

    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert [log2_ranks(w) for w in (1, 2, 4, 8)] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        log2_ranks(3)


# ----------------------------------------------------------------------------------------------
# replicated table, sharded queries / abundance: whole queries per rank, sparse results
# all-gathered; per-unitig count vectors all-reduced
# ----------------------------------------------------------------------------------------------
def _oracle_table():
    from oracle import oracle as O

    k, useqs, reads = _workload()
    otab, _, _ = O.build_mphf(k, lambda: iter(O.Record(str(i), s.decode()) for i, s in enumerate(useqs)))
    return O, k, len(useqs), O.COracleTable(otab.mphf_to_hash, otab.mphf_to_value), reads


def _queries(reads):
    # 7 queries of several records each (one shorter than k, one empty query): uneven over 2 ranks
    qs = [reads[0:40], reads[40:41], [], reads[41:200], [b"ACGT"], reads[200:420] + reads[0:10], reads[420:600]]
    return [list(q) for q in qs]


def _count_batch_oracle(O, ot):
    def count_batch(queries, ksize):
        seqs = [s for q in queries for s in q]
        qseq = np.concatenate([[0], np.cumsum([len(q) for q in queries])]).astype(np.int64)
        buf, off = O.pack_sequences(seqs) if seqs else (np.zeros(0, np.uint8), np.zeros(1, np.int64))
        dicts, nd, _ = ot.query_counts(buf, off, qseq, ksize)
        return [(d, int(n)) for d, n in zip(dicts, nd)]

    return count_batch


def _worker_queries(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spacegraphcats_b200.parallel import abundance_sharded, count_matches_batch_sharded

        O, k, n_ids, ot, reads = _oracle_table()
        res = count_matches_batch_sharded(_count_batch_oracle(O, ot), _queries(reads), k)

        def lookup_counts(buf, off, ksize, n):
            h, _ = O.hash_batch(np.asarray(buf), np.asarray(off), ksize)
            cnt, miss = ot.count_matches(h, n)
            return cnt, len(h), len(h) - int(miss)

        buf, off = O.pack_sequences(reads)
        cnt, nk, nh = abundance_sharded(lookup_counts, buf, off, k, n_ids)
        q.put((rank, res, cnt.tolist(), nk, nh))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_queries_and_abundance_world2_gloo_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_queries, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0

    O, k, n_ids, ot, reads = _oracle_table()
    want = _count_batch_oracle(O, ot)(_queries(reads), k)
    assert any(len(d) > 3 for d, _ in want) and want[2] == ({}, 0) and want[4] == ({}, 0)
    buf, off = O.pack_sequences(reads)
    h, _ = O.hash_batch(buf, off, k)
    w_cnt, w_miss = ot.count_matches(h, n_ids)
    for rank, res, cnt, nk, nh in results:  # every rank holds the complete answer
        assert res == want
        assert cnt == w_cnt.tolist() and nk == len(h) and nh == len(h) - int(w_miss)
