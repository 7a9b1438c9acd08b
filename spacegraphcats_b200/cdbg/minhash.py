"""sourmash-compatible k-mer hashing on the GPU -- "next" rows f3 / f4 of the scope table.

The reference gives every unitig a deterministic id from its n=1 MinHash (``min_hashval``,
cdbg/sort_bcalm_unitigs.py:56-128) and sketches retrieved contigs with a scaled MinHash
(search/query_by_sequence.py:52-89, 176).  Both are ``MinHash.add_sequence`` of the third-party
sourmash: MurmurHash3_x64_128(min(kmer, revcomp(kmer)), seed 42).lo64 per k-mer -- the same Murmur3
core as the k-mer index hash, so it runs in the same tile kernel (``MODE_MINHASH``).
"""
import ctypes

import numpy as np

from ..device import get_context, pack_sequences
from .._lib import check

SOURMASH_SEED = 42


def _run(seqs, ksize, seed, want_hashes, want_min, ctx):
    import torch

    ctx = ctx or get_context()
    buf, off = pack_sequences(seqs)
    d_seq = ctx.to_device(buf, np.uint8)
    d_off = ctx.to_device(off, np.int64)
    nseq = len(off) - 1
    cap = max(d_seq.numel(), 1)
    d_hash = ctx.empty(cap, torch.int64) if want_hashes else None
    d_koff = ctx.empty(nseq + 1, torch.int64)
    d_min = ctx.empty(nseq, torch.int64) if want_min else None
    d_stat = ctx.empty(nseq, torch.int32)
    total = ctypes.c_int64(0)
    check(ctx._L.sgc_minhash_batch(ctx._h, ctx.ptr(d_seq), d_seq.numel(), ctx.ptr(d_off), nseq, int(ksize), int(seed),
                                   ctx.ptr(d_hash) if d_hash is not None else None, cap, ctx.ptr(d_koff),
                                   ctx.ptr(d_min) if d_min is not None else None, ctx.ptr(d_stat), ctypes.byref(total)))
    status = ctx.to_host(d_stat)
    if status.any():  # sourmash add_sequence(force=False) raises on non-ACGT
        raise ValueError("invalid DNA character in input k-mer (sequence %d)" % int(np.flatnonzero(status)[0]))
    hashes = ctx.to_host(d_hash[: total.value], np.uint64) if want_hashes else None
    koff = ctx.to_host(d_koff)
    mins = ctx.to_host(d_min, np.uint64) if want_min else None
    return hashes, koff, mins


def unitig_min_hashvals(seqs, ksize, seed=SOURMASH_SEED, ctx=None):
    """Per-sequence minimum sourmash hash (MinHash n=1) -> uint64[n]; 2^64-1 for a sequence with no
    k-mer.  The ``min_hashval`` sort key of sort_bcalm_unitigs.py:56-128."""
    return _run(seqs, ksize, seed, False, True, ctx)[2]


def sourmash_hashes(seqs, ksize, seed=SOURMASH_SEED, ctx=None):
    """All sourmash k-mer hashes -> (uint64[total], kmer_off[n+1])."""
    h, koff, _ = _run(seqs, ksize, seed, True, False, ctx)
    return h, koff


def scaled_minhash(seqs, ksize, scaled, seed=SOURMASH_SEED, ctx=None):
    """Sorted distinct hashes below 2^64/scaled over all sequences (a scaled MinHash sketch)."""
    h, _ = sourmash_hashes(seqs, ksize, seed, ctx)
    max_hash = (1 << 64) - 1 if scaled <= 1 else int(((1 << 64) - 1) / scaled)
    return np.unique(h[h <= np.uint64(max_hash)])


def assign_unitig_ids(seqs, ksize, ctx=None):
    """The deterministic unitig ids of cdbg/sort_bcalm_unitigs.py:56-128: BCALM's unitigs are renumbered in the
    order of their ``min_hashval`` -- which the reference stores as TEXT, so ``ORDER BY min_hashval`` (:108-118) is
    the lexicographic order of the decimal strings, not the numeric one.  The reference's UNIQUE index on the column
    (:101) makes ties an error there; here they raise ``ValueError``.
    -> (new_id int64[n]: id of input unitig i, min_hashval uint64[n])"""
    mins = unitig_min_hashvals(seqs, ksize, ctx=ctx)
    if len(seqs) and int(mins.max()) == (1 << 64) - 1 and any(len(s) < int(ksize) for s in seqs):
        raise ValueError("a unitig shorter than k has no min_hashval")
    keys = np.array([str(int(m)) for m in mins])
    order = np.argsort(keys, kind="stable")
    if len(keys) > 1 and (keys[order][1:] == keys[order][:-1]).any():
        raise ValueError("UNIQUE constraint failed: sequences.min_hashval")
    new_id = np.empty(len(seqs), np.int64)
    new_id[order] = np.arange(len(seqs), dtype=np.int64)
    return new_id, mins
