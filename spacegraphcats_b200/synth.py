"""Synthetic metagenome workload of SURVEY.md 8(d) config 4, generated on the device (bench and
large-size property tests; not part of the reference surface).

genome G: iid uniform ACGT (seed 1).  unitigs: consecutive pieces of G overlapping by k-1 bases so
every k-mer of G lies in exactly one unitig; k-mers per unitig = 1 + Geometric(1/64), truncated at
4096 (seed 2); ids = order of appearance; 0.01 % of the unitigs are duplicated under a second id
(their k-mers become multicounts and are dropped, exercising that path).  reads: ``read_len``-base
windows of G at uniform positions, half of them reverse-complemented, iid substitution errors
(seed 3 + batch index).
"""

import numpy as np

from ._lib import check


def _torch():
    import torch

    return torch


def unitig_cuts(n_kmers, seed=2, mean=64, cap=4096):
    """Host: cut points (k-mer coordinates) 0 = c0 < c1 < ... = n_kmers."""
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


class SyntheticCdbg:
    """genome + unitigs (device resident) for n_kmers unitig k-mers."""

    def __init__(self, ctx, n_kmers, ksize=31, seed=1, dup_fraction=1e-4):
        torch = _torch()
        L = ctx._L
        self.ctx, self.ksize, self.n_kmers = ctx, int(ksize), int(n_kmers)
        glen = self.n_kmers + self.ksize - 1
        self.genome = ctx.empty(glen, torch.uint8)
        check(L.sgc_synth_genome(ctx._h, ctx.ptr(self.genome), glen, seed))
        cuts = unitig_cuts(self.n_kmers, seed + 1)
        n_main = len(cuts) - 1
        # duplicates: re-emit a few unitigs under new ids at the end
        rng = np.random.default_rng(seed + 7)
        n_dup = int(round(n_main * dup_fraction))
        dup = np.sort(rng.choice(n_main, n_dup, replace=False)) if n_dup else np.zeros(0, np.int64)
        self.n_unitigs = n_main + n_dup
        self.n_dup = n_dup
        self.dup_kmers = int((cuts[dup + 1] - cuts[dup]).sum()) if n_dup else 0
        # main unitigs
        out_bytes = int(cuts[-1] - cuts[0] + n_main * (self.ksize - 1))
        dup_lens = (cuts[dup + 1] - cuts[dup] + self.ksize - 1) if n_dup else np.zeros(0, np.int64)
        total_bytes = out_bytes + int(dup_lens.sum())
        self.useq = ctx.empty(total_bytes, torch.uint8)
        self.uoff = ctx.empty(self.n_unitigs + 1, torch.int64)
        d_cuts = ctx.to_device(cuts)
        check(L.sgc_synth_unitigs(ctx._h, ctx.ptr(self.genome), ctx.ptr(d_cuts), n_main, self.ksize, out_bytes,
                                  ctx.ptr(self.useq), ctx.ptr(self.uoff)))
        if n_dup:
            with torch.cuda.stream(ctx.stream):
                pos = out_bytes
                offs = [pos]
                for a, ln in zip(cuts[dup].tolist(), dup_lens.tolist()):
                    self.useq[pos : pos + ln] = self.genome[a : a + ln]
                    pos += ln
                    offs.append(pos)
                self.uoff[n_main:] = torch.tensor(offs, dtype=torch.int64, device=ctx.device)
        ctx.sync()
        self.total_unitig_kmers = self.n_kmers + self.dup_kmers

    def build_table(self, side=None):
        from .device import DeviceTable

        if side is None:
            from .cdbg.hashing import side_default

            side = side_default(int(self.useq.numel()))

        t = DeviceTable(self.ctx, self.total_unitig_kmers)
        if side:
            t.enable_side()
        t.insert_sequences(self.useq, self.uoff, self.ksize)
        return t

    def reads(self, n_reads, read_len=150, seed=3, err_ppm=5000, rc_half=True):
        """-> (seq uint8[n*len], off int64[n+1]) on the device."""
        torch = _torch()
        ctx = self.ctx
        seq = ctx.empty(n_reads * read_len, torch.uint8)
        off = ctx.empty(n_reads + 1, torch.int64)
        check(ctx._L.sgc_synth_reads(ctx._h, ctx.ptr(self.genome), self.genome.numel(), n_reads, read_len, seed,
                                     err_ppm, 1 if rc_half else 0, ctx.ptr(seq), ctx.ptr(off)))
        return seq, off
