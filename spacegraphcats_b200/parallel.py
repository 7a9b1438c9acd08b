"""Multi-GPU host layer: one process per GPU over ``torch.distributed``.

* **Replicated table** needs nothing from here: every rank builds / loads the full table and labels
  its own contiguous range of reads (:func:`shard_range`); results are concatenated by the caller.
* **Hash-range partitioned table** (:class:`PartitionedIndex`): owner(hash) = top log2(R) bits.
  Build: every rank hashes its shard of the unitigs, buckets the (hash, id) pairs by owner, one
  all-to-all, local insert -- all copies of a hash meet on one rank, so build_mphf's multicount rule
  stays exact.  Lookup: hash local reads, bucket hashes by owner (remembering the permutation),
  all-to-all (8 B per k-mer), owner-side probe, all-to-all back (4 B per k-mer), un-permute,
  per-read distinct ids.

The compute steps go through an *engine* object; :class:`DeviceEngine` is the product engine (CUDA
kernels through the C-ABI, NCCL for the exchange).  The exchange protocol itself is backend
agnostic, which is what the world_size-2 gloo tests exercise with a host engine of their own.
"""
import ctypes

import numpy as np

from ._lib import check


def _torch():
    import torch

    return torch


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of n records for this rank (keeps mates together when n is even-split
    by the caller on pair boundaries)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def log2_ranks(world):
    b = world.bit_length() - 1
    if world < 1 or (1 << b) != world:
        raise ValueError("hash-range partitioning needs a power-of-two number of ranks (got %d)" % world)
    return b


def exchange(send, send_counts, group=None):
    """Variable all-to-all of a 1-D tensor: ``send`` holds the rows for rank 0, 1, ... back to
    back (``send_counts[r]`` rows for rank r).  -> (recv tensor, recv_counts list)."""
    torch = _torch()
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=send.device)
    rc = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.tolist()
    recv = torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    return recv, recv_counts


class _RawPtr:
    """Minimal stand-in for a tensor where only ``data_ptr()`` is needed (library-owned memory)."""

    def __init__(self, ptr):
        self._p = int(ptr)

    def data_ptr(self):
        return self._p


class DeviceEngine:
    """CUDA engine: every step is a kernel of libsgc_b200.so on the context's stream."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._L = ctx._L

    # -- build side -------------------------------------------------------
    def hash_pairs(self, seq, off, ksize, ids):
        """(hash per k-mer, id per k-mer) for a shard of unitigs, on the device."""
        torch = _torch()
        c = self.ctx
        d_off = c.to_device(off, np.int64)
        d_hash, d_koff, d_stat = c.hash_batch(seq, d_off, ksize, to_host=False)
        if d_stat is not None and d_stat.numel() and bool(d_stat.any().item()):
            raise ValueError("Invalid base in read")
        d_ids = c.to_device(ids, np.uint32)
        with torch.cuda.stream(c.stream):
            per_kmer = torch.repeat_interleave(d_ids, (d_koff[1:] - d_koff[:-1]))
        return d_hash, per_kmer

    def partition_pairs(self, d_hash, d_ids, world, want_perm=False):
        torch = _torch()
        c = self.ctx
        n = d_hash.numel()
        oh = c.empty(n, torch.int64)
        oi = c.empty(n, torch.int32)
        d_boff = c.empty(world + 1, torch.int64)
        d_perm = c.empty(max(n, 1), torch.int64) if want_perm else None
        h_boff = (ctypes.c_int64 * (world + 1))()
        check(self._L.sgc_partition_pairs(c._h, c.ptr(d_hash), c.ptr(d_ids), n, world, c.ptr(oh), c.ptr(oi),
                                          c.ptr(d_perm) if want_perm else None, c.ptr(d_boff), h_boff))
        if want_perm:
            return oh, oi, list(h_boff), d_perm[:n]  # out[j] = in[perm[j]]
        return oh, oi, list(h_boff)

    def new_table(self, n, home_shift, shared=False):
        from .device import DeviceTable

        return DeviceTable(self.ctx, max(n, 1), home_shift=home_shift, shared=shared)

    def insert_pairs(self, table, d_hash, d_ids):
        table.insert_pairs(d_hash, d_ids)

    # -- lookup side ------------------------------------------------------
    def hash_partition(self, seq, off, ksize, world):
        """-> (bucketed hashes, perm, bucket_off list, kmer_off tensor, n_kmers)"""
        torch = _torch()
        c = self.ctx
        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        nseq = d_off.numel() - 1
        cap = max(d_seq.numel(), 1)
        d_b = c.empty(cap, torch.int64)
        d_perm = c.empty(cap, torch.int64)
        d_boff = c.empty(world + 1, torch.int64)
        d_koff = c.empty(nseq + 1, torch.int64)
        d_stat = c.empty(nseq, torch.int32)
        h_boff = (ctypes.c_int64 * (world + 1))()
        check(self._L.sgc_hash_partition(c._h, c.ptr(d_seq), d_seq.numel(), c.ptr(d_off), nseq, int(ksize), world,
                                         c.ptr(d_b), c.ptr(d_perm), cap, c.ptr(d_boff), c.ptr(d_koff), c.ptr(d_stat),
                                         h_boff))
        boff = list(h_boff)
        if nseq and bool(d_stat.any().item()):
            raise ValueError("Invalid base in read")
        n = boff[-1]
        return d_b[:n], d_perm[:n], boff, d_koff, n

    def lookup(self, table, d_hash):
        return table.lookup(d_hash, to_host=False)

    def unpermute(self, d_ids, d_perm):
        torch = _torch()
        c = self.ctx
        out = c.empty(d_ids.numel(), torch.int32)
        check(self._L.sgc_unpermute_u32(c._h, c.ptr(d_ids), c.ptr(d_perm), d_ids.numel(), c.ptr(out)))
        return out

    def labels_from_ids(self, d_kmer_ids, d_koff, nseq):
        torch = _torch()
        c = self.ctx
        cap = max(4 * nseq, 1024)
        d_ioff = c.empty(nseq + 1, torch.int64)
        for _ in range(8):
            d_out = c.empty(cap, torch.int32)
            nh, nl = ctypes.c_int64(0), ctypes.c_int64(0)
            from ._lib import SgcCapacityError

            try:
                check(self._L.sgc_labels_from_ids(c._h, c.ptr(d_kmer_ids), c.ptr(d_koff), nseq, d_kmer_ids.numel(),
                                                  c.ptr(d_ioff), c.ptr(d_out), cap, ctypes.byref(nh), ctypes.byref(nl)))
                break
            except SgcCapacityError:
                cap = max(2 * cap, int(nl.value) + 1024)
        return d_ioff, d_out[: nl.value], nh.value

    # -- fused lookup side (no sort, no permutation array) ---------------------------
    def partition_hashes(self, seq, off, ksize, world, slack=1.08):
        """hash + owner bucketing in one kernel.  -> (send regions [world*cap] int64, cap, ridx int32,
        counts list, d_off, n_bytes)"""
        torch = _torch()
        c = self.ctx
        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        nseq = d_off.numel() - 1
        n_max = max(d_seq.numel(), 1)  # k-mers <= bases
        cap = int(n_max / world * slack) + 4096
        from ._lib import SgcCapacityError

        for _ in range(4):
            d_send = c.empty(world * cap, torch.int64)
            d_ridx = c.empty(n_max, torch.int32)
            d_stat = c.empty(nseq, torch.int32)
            h_counts = (ctypes.c_int64 * world)()
            try:
                check(self._L.sgc_partition_hashes(c._h, c.ptr(d_seq), d_seq.numel(), c.ptr(d_off), nseq, int(ksize),
                                                   world, c.ptr(d_send), cap, c.ptr(d_ridx), n_max, c.ptr(d_stat),
                                                   h_counts))
                break
            except SgcCapacityError:
                cap = max(list(h_counts)) + 4096  # skewed owners: size for the real maximum and redo
        else:
            raise RuntimeError("partition_hashes: could not size the send regions")
        if nseq and bool(d_stat.any().item()):
            raise ValueError("Invalid base in read")
        return d_send, cap, d_ridx, list(h_counts), d_off, d_seq.numel()

    def labels_from_replies(self, d_off, n_bytes, ksize, world, d_ridx, d_replies, cap, want_kmer_ids=False):
        torch = _torch()
        c = self.ctx
        nseq = d_off.numel() - 1
        ids_cap = max(4 * nseq, 1024)
        d_ioff = c.empty(nseq + 1, torch.int64)
        d_kids = c.empty(max(n_bytes, 1), torch.int32) if want_kmer_ids else None
        from ._lib import SgcCapacityError

        for _ in range(8):
            d_out = c.empty(ids_cap, torch.int32)
            nk, nh, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            try:
                check(self._L.sgc_labels_from_replies(c._h, c.ptr(d_off), nseq, n_bytes, int(ksize), world,
                                                      c.ptr(d_ridx), c.ptr(d_replies), cap, c.ptr(d_ioff), c.ptr(d_out),
                                                      ids_cap, c.ptr(d_kids) if d_kids is not None else None,
                                                      ctypes.byref(nk), ctypes.byref(nh), ctypes.byref(nl)))
                break
            except SgcCapacityError:
                ids_cap = max(2 * ids_cap, int(nl.value) + 1024)
        if d_kids is not None:
            d_kids = d_kids[: nk.value]
        return d_ioff, d_out[: nl.value], nh.value, d_kids

    def sync(self):
        self.ctx.sync()

    def stream(self):
        """Context in which collectives must be issued so that they are ordered with the kernels."""
        return _torch().cuda.stream(self.ctx.stream)

    def to_host(self, t, dtype=None):
        return self.ctx.to_host(t, dtype)


def exchange_fds(my_fd, rank, world, group=None):
    """Every rank of ONE node offers a file descriptor to every other rank (SCM_RIGHTS over unix sockets).
    -> list of world fds (None at ``rank``); the caller closes them."""
    import os
    import socket
    import tempfile
    import threading

    import torch.distributed as dist

    path = os.path.join(tempfile.gettempdir(), "sgc_b200_%d_%d.sock" % (os.getpid(), rank))
    if os.path.exists(path):
        os.unlink(path)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(world)

    def serve():
        for _ in range(world - 1):
            conn, _ = srv.accept()
            with conn:
                socket.send_fds(conn, [b"f"], [my_fd])
                conn.recv(1)  # the receiver holds its copy

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    paths = [None] * world
    dist.all_gather_object(paths, path, group=group)
    fds = [None] * world
    for r in range(world):
        if r == rank:
            continue
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
            s.connect(paths[r])
            _, got, _, _ = socket.recv_fds(s, 1, 1)
            fds[r] = got[0]
            s.send(b"k")
    th.join()
    srv.close()
    os.unlink(path)
    return fds


class PeerBuffers:
    """This rank's inbox (R regions of region_cap hashes), reply buffer (R regions of region_cap ids) and
    mailbox (per source rank: request count, request stamp, reply stamp), allocated by the library and mapped
    into every other rank's process with CUDA IPC, so kernels store requests / replies / flags straight into
    the destination GPU over NVLink."""

    KINDS = ("inbox", "replies", "mail_count", "mail_stamp", "reply_stamp")

    def __init__(self, ctx, group, region_cap):
        import torch.distributed as dist

        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.region_cap = int(region_cap)
        L = ctx._L
        self._own = []
        handles = []
        sizes = (self.world * self.region_cap * 8, self.world * self.region_cap * 4, 256, 256, 256)
        for nbytes in sizes:
            ptr = ctypes.c_void_p()
            hnd = (ctypes.c_uint8 * 64)()
            check(L.sgc_peer_alloc(ctx._h, nbytes, ctypes.byref(ptr), hnd))
            self._own.append(ptr)
            handles.append(bytes(hnd))
        ctx.sync()  # sgc_peer_alloc zeroes what it allocates: mailboxes start at zero, epochs count from 1
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=group)
        self._opened = []
        self.ptrs = {kind: (ctypes.c_void_p * self.world)() for kind in self.KINDS}
        for r in range(self.world):
            for which, kind in enumerate(self.KINDS):
                if r == self.rank:
                    self.ptrs[kind][r] = self._own[which].value
                else:
                    ptr = ctypes.c_void_p()
                    hnd = (ctypes.c_uint8 * 64).from_buffer_copy(everyone[r][which])
                    check(L.sgc_peer_open(ctx._h, hnd, ctypes.byref(ptr)))
                    self._opened.append(ptr)
                    self.ptrs[kind][r] = ptr.value
        self.inbox_ptrs = self.ptrs["inbox"]
        self.reply_ptrs = self.ptrs["replies"]
        self.inbox = self._own[0]
        self.replies = self._own[1]
        self.epoch = 0
        dist.barrier(group=group)  # every mailbox is zeroed and mapped before anybody stamps one

    def close(self):
        import torch.distributed as dist

        L = self.ctx._L
        self.ctx.sync()
        for ptr in self._opened:
            L.sgc_peer_close(self.ctx._h, ptr)
        self._opened = []
        dist.barrier(group=self.group)  # nobody still maps our memory
        for ptr in self._own:
            L.sgc_peer_free(self.ctx._h, ptr)
        self._own = []


class PartitionedIndex:
    """k-mer -> unitig table hash-range partitioned over the ranks of a process group."""

    def __init__(self, engine, group=None, p2p=False, peer_probe=False, use_filter=True):
        import torch.distributed as dist

        self.engine = engine
        self.group = group
        self.p2p = bool(p2p)  # fuse the exchange into the kernels over NVLink peer memory
        # probe the owners' tables where they lie (peer loads over NVLink) and settle most k-mers against a
        # replicated unitig sequence store: no exchange buffers, no collective in the label path
        self.peer_probe = bool(peer_probe)
        self._store = None  # (useq_all, brk_all) tensors of the replicated store
        # peer probing: a replicated miss filter (~2 bytes per hash of the whole cDBG on every GPU) answers most
        # probes of k-mers that are in nobody's table from local HBM instead of over NVLink
        self.use_filter = bool(use_filter)
        self._filter = None
        self._peer_tables = []  # peer-mapped bucket arrays (to be closed)
        self._peers = None
        self._need_seen = 0
        self._peers2 = None  # double-buffered peer buffers of the pipelined path
        self._aux = None  # (partition context, gather context) of the pipelined path
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.bits = log2_ranks(self.world)
        self.table = None
        self.n_local = 0

    def close(self):
        """Unmap / free the peer buffers (collective: every rank must call it) and the local table."""
        for pb in ([self._peers] if self._peers is not None else []) + list(self._peers2 or ()):
            pb.close()
        self._peers, self._peers2 = None, None
        if self._peer_tables or self._store is not None:
            import torch.distributed as dist

            c = self.engine.ctx
            c.sync()
            dist.barrier(group=self.group)  # nobody is still probing this rank's table
            for ptr, nbytes in self._peer_tables:
                c._L.sgc_peer_unmap(c._h, ctypes.c_void_p(ptr), nbytes)
            self._peer_tables, self._store, self._filter = [], None, None
            dist.barrier(group=self.group)  # nobody still maps our table
        if self.table is not None and hasattr(self.table, "close"):
            self.table.close()
        self.table = None

    def _build_peer(self, seq, off, ksize, ids):
        """Build for peer probing: every slot also records where its k-mer lies in the shard of the unitig
        sequence store of the rank that hashed it; the store is replicated, the owners' tables are mapped
        into every rank.  The unitig ids must be dealt to the ranks in ascending ranges."""
        import torch.distributed as dist

        from ._lib import BRK_PAD_BITS, STORE_PAD_WORDS

        torch = _torch()
        e, c, R = self.engine, self.engine.ctx, self.world
        L = c._L
        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        d_ids = c.to_device(ids, np.uint32)
        nseq, n_bytes = d_off.numel() - 1, d_seq.numel()
        lo = hi = None
        if nseq:
            with e.stream():
                u = d_ids.to(torch.int64) & 0xFFFFFFFF
                lo, hi = int(u.min().item()), int(u.max().item())
        d_hash, d_aux, d_koff, d_stat = c.hash_positions(d_seq, d_off, ksize)
        bad = 1 if (nseq and bool(d_stat.any().item())) else 0
        info = [None] * R
        dist.all_gather_object(info, (lo, hi, n_bytes, bad), group=self.group)
        if any(x[3] for x in info):
            raise ValueError("Invalid base in read")
        bounds, prev_hi = [], -1
        for lo_r, hi_r, _, _ in info:
            if lo_r is None:
                lo_r, hi_r = prev_hi + 1, prev_hi
            if lo_r <= prev_hi:
                raise ValueError("a peer-probed partitioned table needs the unitig ids dealt to the ranks in ascending ranges")
            bounds.append(lo_r)
            prev_hi = max(prev_hi, hi_r)
        bounds[0] = 0
        bounds.append(0xFFFFFFFF)
        max_bytes = max(x[2] for x in info)
        with e.stream():
            per_kmer = torch.repeat_interleave(d_ids, (d_koff[1:] - d_koff[:-1]))
        bh, bi, boff, perm = e.partition_pairs(d_hash, per_kmer, R, want_perm=True)
        with e.stream():
            ba = d_aux[perm]
        del d_hash, d_aux, per_kmer, perm
        counts = [boff[r + 1] - boff[r] for r in range(R)]
        e.sync()
        with e.stream():
            rh, rcounts = exchange(bh, counts, self.group)
            ri, _ = exchange(bi, counts, self.group)
            ra, _ = exchange(ba, counts, self.group)
        del bh, bi, ba
        self.n_local = int(sum(rcounts))
        n_table, = self._collective_flags(self.n_local)  # the same geometry everywhere: one nbuckets for every owner
        self.table = e.new_table(max(n_table, 1), self.bits, shared=True)
        marks = self.table.insert_triples(rh, ri, ra, bounds)
        # the miss filter: ~16 bits per hash, one 64-bit word per hash chosen by its top bits -- the owner's hashes fill
        # one contiguous slice of it, so the replicated filter is simply the all-gather of the owners' slices
        tot = torch.tensor([self.n_local], dtype=torch.int64, device=c.device)
        dist.all_reduce(tot, group=self.group)
        fbits = max(self.bits + 4, int(max(int(tot.item()), 1) // 4).bit_length())
        self._filter = None
        if self.use_filter:
            with e.stream():
                filt = torch.zeros(1 << fbits, dtype=torch.int64, device=c.device)
            check(L.sgc_filter_insert(c._h, c.ptr(rh), rh.numel(), c.ptr(filt), fbits))
            with e.stream():
                per = filt.numel() // R
                mine = filt[self.rank * per:(self.rank + 1) * per].clone()
                dist.all_gather_into_tensor(filt, mine, group=self.group)
                del mine
            self._filter = (filt, fbits)
        del rh, ri, ra
        # my shard of the store, then everybody's
        ustride = (max_bytes + 15) // 16 + 2 * STORE_PAD_WORDS
        bstride = (max_bytes + BRK_PAD_BITS + 31) // 32 + 2
        with e.stream():
            my_useq = torch.zeros(ustride, dtype=torch.int32, device=c.device)
            my_brk = torch.full((bstride,), -1, dtype=torch.int32, device=c.device)
        check(L.sgc_store_build(c._h, c.ptr(d_seq), n_bytes, c.ptr(d_off), nseq, int(ksize),
                                ctypes.c_void_p(my_useq.data_ptr() + 4 * STORE_PAD_WORDS), c.ptr(my_brk)))
        n_marks, = self._collective_flags(int(marks.numel()))
        with e.stream():
            useq_all = torch.empty(R * ustride, dtype=torch.int32, device=c.device)
            brk_all = torch.empty(R * bstride, dtype=torch.int32, device=c.device)
            dist.all_gather_into_tensor(useq_all, my_useq, group=self.group)
            dist.all_gather_into_tensor(brk_all, my_brk, group=self.group)
            if n_marks:
                mine = torch.full((n_marks,), -1, dtype=torch.int64, device=c.device)  # -1 = no mark
                mine[: marks.numel()] = marks
                every = torch.empty(R * n_marks, dtype=torch.int64, device=c.device)
                dist.all_gather_into_tensor(every, mine, group=self.group)
                every = every[every >= 0].contiguous()
        if n_marks and every.numel():
            check(L.sgc_store_apply_marks(c._h, c.ptr(every), every.numel(), c.ptr(brk_all), bstride))
        # everybody's bucket array, mapped with 2 MB pages (a file descriptor per table, passed between the ranks)
        import os

        ptr, nbytes, fd = self.table.export_fd()
        fds = exchange_fds(fd, self.rank, R, self.group)
        os.close(fd)
        ptrs = []
        for r in range(R):
            if r == self.rank:
                ptrs.append(ptr)
            else:
                p = ctypes.c_void_p()
                check(L.sgc_peer_import_fd(c._h, fds[r], nbytes, ctypes.byref(p)))
                os.close(fds[r])
                self._peer_tables.append((p.value, nbytes))
                ptrs.append(p.value)
        self._store = (useq_all, brk_all)
        self.table.set_peers(ptrs, bounds, useq_all.data_ptr() + 4 * STORE_PAD_WORDS, ustride, brk_all.data_ptr(), bstride)
        if self._filter is not None:
            self.table.set_filter(self._filter[0].data_ptr(), self._filter[1], self.rank)
        c.sync()
        del d_seq, d_off, d_ids, d_koff, d_stat, my_useq, my_brk
        torch.cuda.empty_cache()  # the build's temporaries go back to the driver: the library allocates its own scratch
        dist.barrier(group=self.group)  # every table is complete before anybody probes it
        return self

    def build(self, seq, off, ksize, ids):
        """Every rank passes ITS shard of the unitigs (ids are global cDBG ids)."""
        if self.peer_probe:
            return self._build_peer(seq, off, ksize, ids)
        e = self.engine
        d_hash, d_ids = e.hash_pairs(seq, off, ksize, ids)
        bh, bi, boff = e.partition_pairs(d_hash, d_ids, self.world)
        counts = [boff[r + 1] - boff[r] for r in range(self.world)]
        e.sync()
        with e.stream():
            rh, rcounts = exchange(bh, counts, self.group)
            ri, _ = exchange(bi, counts, self.group)
        self.n_local = int(sum(rcounts))
        self.table = e.new_table(self.n_local, self.bits)
        if self.n_local:
            e.insert_pairs(self.table, rh, ri)
        return self

    def lookup_sequences(self, seq, off, ksize):
        """Per-k-mer unitig ids (MISS = 0xFFFFFFFF) of this rank's sequences, original order.
        -> (kmer_ids tensor, kmer_off tensor)"""
        e = self.engine
        bh, perm, boff, koff, n = e.hash_partition(seq, off, ksize, self.world)
        counts = [boff[r + 1] - boff[r] for r in range(self.world)]
        e.sync()
        with e.stream():
            rh, rcounts = exchange(bh, counts, self.group)  # requests: 8 B per k-mer
        rid = e.lookup(self.table, rh)
        e.sync()
        with e.stream():
            back, bcounts = exchange(rid, rcounts, self.group)  # replies: 4 B per k-mer
        assert bcounts == counts
        return e.unpermute(back, perm), koff

    def _label_reads_fused(self, seq, off, ksize):
        """Fused request/reply path of the CUDA engine: one kernel hashes and appends each hash to
        its owner's send region; the owners probe; one kernel turns the replies into labels."""
        import torch.distributed as dist

        import os
        import time

        torch = _torch()
        e, R = self.engine, self.world
        trace = os.environ.get("SGC_TRACE") and self.rank == 0
        marks = []

        def mark(name):
            if trace:
                e.sync()
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))

        mark("start")
        d_send, cap, ridx, counts, d_off, n_bytes = e.partition_hashes(seq, off, ksize, R)
        mark("hash+partition")
        with e.stream():
            sc = torch.tensor(counts, dtype=torch.int64, device=d_send.device)
            rc = torch.empty(R, dtype=torch.int64, device=d_send.device)
            dist.all_to_all_single(rc, sc, group=self.group)
            rcounts = rc.tolist()
            recv = torch.empty(int(sum(rcounts)), dtype=torch.int64, device=d_send.device)
            dist.all_to_all(list(recv.split(rcounts)), [d_send[r * cap : r * cap + counts[r]] for r in range(R)],
                            group=self.group)  # requests: 8 B per k-mer
        mark("all-to-all requests")
        rid = e.lookup(self.table, recv)
        e.sync()
        mark("owner probe")
        with e.stream():
            back = torch.empty(R * cap, dtype=torch.int32, device=d_send.device)
            dist.all_to_all([back[r * cap : r * cap + counts[r]] for r in range(R)], list(rid.split(rcounts)),
                            group=self.group)  # replies: 4 B per k-mer, same positions as the requests
        mark("all-to-all replies")
        ioff, ids, nh, _ = e.labels_from_replies(d_off, n_bytes, ksize, R, ridx, back, cap)
        mark("gather+labels")
        if trace:
            print("[K7 trace] " + ", ".join("%s %.2f ms" % (b[0], (b[1] - a[1]) * 1e3) for a, b in zip(marks, marks[1:])),
                  flush=True)
        return ioff, ids, nh

    def _label_reads_p2p(self, seq, off, ksize):
        """Exchange fused into the kernels: the hash kernel stores each hash straight into its owner
        GPU's inbox, the owner's probe kernel stores each id straight into the requester's reply
        buffer (peer memory over NVLink).  Per step the ranks exchange only the region counts (which
        also orders the stores) and one barrier."""
        import os
        import time

        import torch.distributed as dist

        torch = _torch()
        e, R, c = self.engine, self.world, self.engine.ctx
        L = c._L
        trace = os.environ.get("SGC_TRACE") and self.rank == 0
        marks = [("start", time.perf_counter())] if trace else []

        def mark(name):
            if trace:
                c.sync()
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))

        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        nseq, n_bytes = d_off.numel() - 1, d_seq.numel()
        need = int(max(n_bytes, 1) / R * 1.08) + 4096
        with e.stream():
            t = torch.tensor([need], dtype=torch.int64, device=c.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            need = int(t.item())  # read on the stream the collective was issued on
        if self._peers is None or self._peers.region_cap < need:
            if self._peers is not None:
                self._peers.close()
            self._peers = PeerBuffers(c, self.group, need)
        pb = self._peers
        cap = pb.region_cap
        d_ridx = c.empty(max(n_bytes, 1), torch.int32)
        d_stat = c.empty(nseq, torch.int32)
        h_counts = (ctypes.c_int64 * R)()
        mark("setup")
        check(L.sgc_partition_hashes_p2p(c._h, c.ptr(d_seq), n_bytes, c.ptr(d_off), nseq, int(ksize), R, self.rank,
                                         pb.inbox_ptrs, cap, c.ptr(d_ridx), max(n_bytes, 1), c.ptr(d_stat), h_counts))
        if nseq and bool(d_stat.any().item()):
            raise ValueError("Invalid base in read")
        mark("hash + P2P stores into the owners' inboxes")
        with e.stream():
            sc = torch.tensor(list(h_counts), dtype=torch.int64, device=c.device)
            rc = torch.empty(R, dtype=torch.int64, device=c.device)
            dist.all_to_all_single(rc, sc, group=self.group)  # also orders every rank's stores before the probes
            counts_from = (ctypes.c_int64 * R)(*rc.tolist())
        mark("counts")
        check(L.sgc_lookup_regions_p2p(self.table._h, pb.inbox, counts_from, cap, R, self.rank, pb.reply_ptrs, 0))
        c.sync()
        mark("owner probe + P2P stores of the replies")
        with e.stream():
            dist.barrier(group=self.group)  # all replies have landed; all inboxes may be reused
        mark("barrier")
        ioff, ids, nh, _ = e.labels_from_replies(d_off, n_bytes, ksize, R, d_ridx, _RawPtr(pb.replies.value), cap)
        mark("gather+labels")
        if trace:
            print("[K7 p2p trace] " + ", ".join("%s %.2f ms" % (b[0], (b[1] - a[1]) * 1e3) for a, b in zip(marks, marks[1:])),
                  flush=True)
        return ioff, ids, nh

    def _collective_flags(self, *flags):
        """MAX over the ranks of a few small ints, so that every rank takes the same error / retry decision
        (a rank-local raise before the next collective would leave the others waiting for ever)."""
        import torch.distributed as dist

        torch = _torch()
        c = self.engine.ctx
        with self.engine.stream():
            t = torch.tensor([int(f) for f in flags], dtype=torch.int64, device=c.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            return t.tolist()

    def _label_reads_flags(self, seq, off, ksize):
        """The fused path with the ranks synchronised by flags in peer memory: ONE library call per step
        (request, serve and gather kernels back to back on the stream, no count exchange, no barrier), then
        one tiny all-reduce that makes the error / retry decision collective."""
        from ._lib import SgcCapacityError, SgcError

        torch = _torch()
        e, R, c = self.engine, self.world, self.engine.ctx
        L = c._L
        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        nseq, n_bytes = d_off.numel() - 1, d_seq.numel()
        need = int(max(n_bytes, 1) / R * 1.08) + 4096
        if self._peers is None or self._need_seen < need:
            need, = self._collective_flags(need)  # sized once for the largest batch seen so far (collective)
            self._need_seen = need
        else:
            need = self._peers.region_cap
        d_ridx = c.empty(max(n_bytes, 1), torch.int32)
        d_stat = c.empty(max(nseq, 1), torch.int32)
        d_ioff = c.empty(nseq + 1, torch.int64)
        ids_cap = max(4 * nseq, 1024)
        for _ in range(6):
            if self._peers is None or self._peers.region_cap < need:
                if self._peers is not None:
                    self._peers.close()
                self._peers = PeerBuffers(c, self.group, need)
            pb = self._peers
            pb.epoch += 1
            d_ids = c.empty(ids_cap, torch.int32)
            nk, nh, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            flags = ctypes.c_int32(0)
            err, cap_short = None, 0
            try:
                check(L.sgc_label_reads_partitioned(
                    self.table._h, c.ptr(d_seq), n_bytes, c.ptr(d_off), nseq, int(ksize), R, self.rank, pb.epoch,
                    pb.ptrs["inbox"], pb.ptrs["replies"], pb.ptrs["mail_count"], pb.ptrs["mail_stamp"],
                    pb.ptrs["reply_stamp"], pb.region_cap, c.ptr(d_ridx), max(n_bytes, 1), c.ptr(d_ioff), c.ptr(d_ids),
                    ids_cap, c.ptr(d_stat), ctypes.byref(nk), ctypes.byref(nh), ctypes.byref(nl), ctypes.byref(flags)))
            except SgcCapacityError as ex:
                err = ex
                cap_short = 1 if (flags.value & 4) else 2  # 1: request regions, 2: label buffer (rank-local)
            except SgcError as ex:
                err = ex
            bad_base = 1 if (nseq and err is None and bool(d_stat[:nseq].any().item())) else 0
            fatal = 1 if (err is not None and not cap_short) else 0
            g_fatal, g_region, g_bad = self._collective_flags(fatal, 1 if cap_short == 1 else 0, bad_base)
            if g_fatal:
                raise err if err is not None else RuntimeError("partitioned step failed on another rank")
            if g_bad:
                raise ValueError("Invalid base in read")
            if g_region:  # some rank's owner ranges are skewed: all ranks grow their buffers and redo the step
                need = int(self._peers.region_cap * 1.5) + 4096
                self._need_seen = need
                continue
            if cap_short == 2:  # only my label buffer was short; the redo is collective all the same (flags)
                ids_cap = max(2 * ids_cap, int(nl.value) + 1024)
            redo, = self._collective_flags(1 if cap_short == 2 else 0)
            if redo:
                continue
            return d_ioff, d_ids[: nl.value], nh.value
        raise RuntimeError("partitioned step: could not size the exchange buffers")

    def label_reads_pipelined(self, batches, ksize):
        """Label a list of sub-batches [(seq, off), ...] with the phases of consecutive sub-batches
        overlapped on three streams: while the owners probe sub-batch j (HBM-bound, table context),
        this rank turns the replies of sub-batch j-1 into labels (gather context) and hashes +
        stores sub-batch j+1 into the owners' other inbox (partition context, integer-bound).
        Inboxes and reply buffers are double-buffered.  -> list of (ids_off, ids, n_hits)."""
        import torch.distributed as dist

        from .device import Context

        torch = _torch()
        e, R, cT = self.engine, self.world, self.engine.ctx
        L = cT._L
        if self._aux is None:
            self._aux = (Context(cT.device_index), Context(cT.device_index))
        cP, cG = self._aux
        eG = DeviceEngine(cG)
        S = len(batches)
        if S == 0:
            return []
        d_in = [(cP.to_device(s, np.uint8), cP.to_device(o, np.int64)) for s, o in batches]
        need = max(int(max(s.numel(), 1) / R * 1.08) + 4096 for s, _ in d_in)
        with torch.cuda.stream(cP.stream):
            t = torch.tensor([need], dtype=torch.int64, device=cP.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            need = int(t.item())
        if self._peers2 is None or self._peers2[0].region_cap < need:
            if self._peers2 is not None:
                for pb in self._peers2:
                    pb.close()
            self._peers2 = (PeerBuffers(cT, self.group, need), PeerBuffers(cT, self.group, need))
        cap = self._peers2[0].region_cap
        ridx = [None, None]

        def part(j):
            d_seq, d_off = d_in[j]
            nseq, n_bytes = d_off.numel() - 1, d_seq.numel()
            pb = self._peers2[j % 2]
            ridx[j % 2] = cP.empty(max(n_bytes, 1), torch.int32)
            d_stat = cP.empty(nseq, torch.int32)
            h_counts = (ctypes.c_int64 * R)()
            check(L.sgc_partition_hashes_p2p(cP._h, cP.ptr(d_seq), n_bytes, cP.ptr(d_off), nseq, int(ksize), R, self.rank,
                                             pb.inbox_ptrs, cap, cP.ptr(ridx[j % 2]), max(n_bytes, 1), cP.ptr(d_stat),
                                             h_counts))
            if nseq and bool(d_stat.any().item()):
                raise ValueError("Invalid base in read")
            with torch.cuda.stream(cP.stream):
                sc = torch.tensor(list(h_counts), dtype=torch.int64, device=cP.device)
                rc = torch.empty(R, dtype=torch.int64, device=cP.device)
                dist.all_to_all_single(rc, sc, group=self.group)
                return (ctypes.c_int64 * R)(*rc.tolist())

        def gather(j):
            d_seq, d_off = d_in[j]
            pb = self._peers2[j % 2]
            ioff, ids, nh, _ = eG.labels_from_replies(d_off, d_seq.numel(), ksize, R, ridx[j % 2], _RawPtr(pb.replies.value), cap)
            return ioff, ids, nh

        import os
        import time

        trace = os.environ.get("SGC_TRACE") and self.rank == 0
        marks = [("start", time.perf_counter())]

        def mark(name):
            if trace:
                marks.append((name, time.perf_counter()))

        out = [None] * S
        cf = part(0)
        mark("part0")
        for j in range(S):
            pb = self._peers2[j % 2]
            check(L.sgc_lookup_regions_p2p(self.table._h, pb.inbox, cf, cap, R, self.rank, pb.reply_ptrs, 2))  # async
            mark("launch probe%d" % j)
            if j >= 1:
                out[j - 1] = gather(j - 1)
                mark("gather%d" % (j - 1))
            if j + 1 < S:
                cf = part(j + 1)
                mark("part%d" % (j + 1))
            cT.sync()
            mark("probe%d done" % j)
            with torch.cuda.stream(cT.stream):
                dist.barrier(group=self.group)
            cT.sync()  # host-blocking: the other two streams are not ordered after the barrier by themselves
            mark("barrier%d" % j)
        out[S - 1] = gather(S - 1)
        mark("gather%d" % (S - 1))
        if trace:
            print("[K7 pipeline trace] " + ", ".join("%s %.2f" % (b[0], (b[1] - a[1]) * 1e3) for a, b in zip(marks, marks[1:])),
                  flush=True)
        return out

    def label_reads(self, seq, off, ksize):
        """index_reads' per-read step against the partitioned table.
        -> (ids_off, ids, n_hits) as engine tensors"""
        if self.peer_probe:  # the ordinary label call: the kernel itself goes to the owners' memory
            res = self.table.label_reads(seq, off, ksize, to_host=False)
            if res["status"].numel() and bool(res["status"].any().item()):
                raise ValueError("Invalid base in read")
            return res["ids_off"], res["ids"], res["n_hits"]
        if self.p2p and hasattr(self.engine, "partition_hashes"):
            import os

            if os.environ.get("SGC_PART_FLAGS", "1") != "0":
                return self._label_reads_flags(seq, off, ksize)
            return self._label_reads_p2p(seq, off, ksize)
        if hasattr(self.engine, "partition_hashes"):
            return self._label_reads_fused(seq, off, ksize)
        kmer_ids, koff = self.lookup_sequences(seq, off, ksize)
        nseq = koff.numel() - 1 if hasattr(koff, "numel") else len(koff) - 1
        return self.engine.labels_from_ids(kmer_ids, koff, nseq)


# ----------------------------------------------------------------------------------------------
# Replicated table, sharded queries ("query batches shard naturally")
# ----------------------------------------------------------------------------------------------
def allgather_var(t, group=None):
    """All-gather of 1-D tensors of different lengths (same dtype / device on every rank).
    -> list of world tensors (rank order).  Two collectives: the lengths, then the rows padded to the longest."""
    torch = _torch()
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    ns = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(x.item()) for x in ns]
    m = max(ns + [1])
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[:c] for o, c in zip(out, ns)]


def allreduce_counts(counts, group=None):
    """Per-unitig count vectors of the ranks' shards -> their sum on every rank, in place
    (``ncclAllReduce(sum, u32[n_ids])`` of SURVEY.md 8e; u32 travels as i32, the sum wraps the same way)."""
    torch = _torch()
    import torch.distributed as dist

    v = counts.view(torch.int32) if counts.dtype != torch.int32 else counts
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return counts


def count_matches_batch_sharded(count_batch, queries, ksize, device=None, group=None):
    """``count_cdbg_matches`` of many queries on R ranks that all hold the (replicated) table: rank r answers
    the contiguous range ``shard_range(len(queries), r, R)`` -- a query's k-mer SET must be formed on one rank,
    so whole queries are the unit -- and the sparse per-query results are all-gathered, so every rank returns
    the complete list.  search/query_by_sequence.py:153-189 for every query; SURVEY.md 8e.

    ``count_batch(list of queries, ksize)`` -> list of (``{unitig: count}``, n_distinct) is the per-rank
    engine: ``GpuKmerTable.count_matches_batch`` in the product, an oracle-backed callable in the gloo tests.
    ``device``: where the exchanged arrays live (the rank's GPU under NCCL, None = CPU under gloo)."""
    torch = _torch()
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(len(queries), rank, world)
    mine = count_batch(queries[lo:hi], ksize) if hi > lo else []
    npairs = np.array([len(c) for c, _ in mine], np.int64)
    ids = np.fromiter((i for c, _ in mine for i in c.keys()), np.int64, int(npairs.sum()))
    cnt = np.fromiter((v for c, _ in mine for v in c.values()), np.int64, int(npairs.sum()))
    nd = np.array([d for _, d in mine], np.int64)

    def dev(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device) if device is not None else t

    g_np, g_ids, g_cnt, g_nd = (allgather_var(dev(a), group) for a in (npairs, ids, cnt, nd))
    out = []
    for r in range(world):
        rnp, rids, rcnt, rnd = (x[r].cpu().numpy() for x in (g_np, g_ids, g_cnt, g_nd))
        o = np.concatenate([[0], np.cumsum(rnp)])
        for j in range(len(rnp)):
            a, b = int(o[j]), int(o[j + 1])
            out.append((dict(zip(rids[a:b].tolist(), rcnt[a:b].tolist())), int(rnd[j])))
    assert len(out) == len(queries)
    return out


def abundance_sharded(lookup_counts, buf, off, ksize, n_ids, device=None, group=None):
    """Per-unitig k-mer abundance of a read set on R ranks with replicated tables (the loop of
    search/count_dominator_abundance.py:80-97, no de-duplication): rank r looks up the k-mers of its contiguous
    range of sequences, the u32[n_ids] count vectors are summed with one all-reduce.

    ``lookup_counts(buf, off, ksize, n_ids)`` -> (count_by_id u32[n_ids] numpy, n_kmers, n_hits) for a shard.
    -> (count_by_id numpy u32[n_ids], n_kmers, n_hits) of the whole set, on every rank."""
    torch = _torch()
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    off = np.asarray(off, np.int64)
    lo, hi = shard_range(len(off) - 1, rank, world)
    sub = np.asarray(buf)[off[lo]: off[hi]]
    cnt, nk, nh = lookup_counts(sub, off[lo: hi + 1] - off[lo], ksize, n_ids)
    t = torch.from_numpy(np.ascontiguousarray(cnt, np.uint32).view(np.int32).copy())
    tot = torch.tensor([nk, nh], dtype=torch.int64)
    if device is not None:
        t, tot = t.to(device), tot.to(device)
    allreduce_counts(t, group)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().view(np.uint32), int(tot[0].item()), int(tot[1].item())
