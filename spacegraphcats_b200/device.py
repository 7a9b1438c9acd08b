"""Host-side handles over the C-ABI: one :class:`Context` per GPU, :class:`DeviceTable`
for the GPU-resident hash -> unitig-id map.  PyTorch is used only to own device / pinned
memory and the CUDA stream; every computation is a kernel of ``libsgc_b200.so``.

Inputs may be numpy arrays (copied host->device through pinned memory on the context's
stream) or CUDA tensors already resident on the context's device (used as they are).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import SGC_MISS, SgcCapacityError, check

_contexts = {}


def _torch():
    import torch

    return torch


def pack_sequences(seqs):
    """list of str/bytes -> (uint8 concatenation, int64 offsets[n+1])."""
    bs = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8) if off[-1] else np.zeros(0, np.uint8)
    return buf, off


def pack_reads_host(buf, off, n_threads=0):
    """ASCII (uint8 concatenation, int64 offsets) -> 2-bit packed records for ``label_reads_packed``
    (sgc_pack_reads).  -> (packed uint8[...], lengths int32[n], bad bool[n]); ``bad`` marks records
    holding a byte other than A/C/G/T (packed as A; the caller applies the reference's policy)."""
    L = _lib.lib()
    buf = np.ascontiguousarray(buf, np.uint8)
    off = np.ascontiguousarray(off, np.int64)
    n = len(off) - 1
    lens = np.diff(off)
    cap = int(((lens + 15) // 16).sum()) * 4
    packed = np.empty(max(cap, 16), np.uint8)
    h_len = np.zeros(max(n, 1), np.int32)
    bad = np.zeros(max(n, 1), np.uint8)
    nb = ctypes.c_int64(0)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    check(L.sgc_pack_reads(vp(buf), vp(off), n, int(n_threads), vp(packed), cap, vp(h_len), vp(bad), ctypes.byref(nb)))
    return packed[: nb.value], h_len[:n], bad[:n].astype(bool)


def kmer_counts(off, ksize):
    """per-sequence k-mer counts max(0, L-k+1) from offsets (host)."""
    return np.maximum(np.diff(np.asarray(off, np.int64)) - int(ksize) + 1, 0)


class Context:
    """One GPU, one CUDA stream, scratch memory (``sgc_ctx``)."""

    def __init__(self, device=0):
        torch = _torch()
        L = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("spacegraphcats_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self.stream = torch.cuda.Stream(device=self.device)
        h = ctypes.c_void_p()
        check(L.sgc_ctx_create(self.device_index, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(h)))
        self._h = h
        self._L = L

    def close(self):
        if getattr(self, "_h", None):
            self._L.sgc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- memory helpers -------------------------------------------------
    def empty(self, n, dtype):
        torch = _torch()
        with torch.cuda.stream(self.stream):  # block belongs to the context stream's pool
            return torch.empty(max(int(n), 0), dtype=dtype, device=self.device)

    def to_device(self, a, dtype=None):
        """numpy array / CUDA tensor -> CUDA tensor on this context's device (stream-ordered)."""
        torch = _torch()
        if isinstance(a, torch.Tensor):
            if a.device.type == "cpu":  # host tensor (pinned memory copies asynchronously on the context's stream)
                with torch.cuda.stream(self.stream):
                    return a.contiguous().to(self.device, non_blocking=True)
            if a.device != self.device:
                raise ValueError("tensor lives on %s, context on %s" % (a.device, self.device))
            # the caller's stream produced it: order the context's stream after that work (ADVICE: stream safety)
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            a.record_stream(self.stream)
            return a.contiguous()
        a = np.ascontiguousarray(a, dtype=dtype)
        if a.dtype == np.uint64:
            a = a.view(np.int64)
        elif a.dtype == np.uint32:
            a = a.view(np.int32)
        if not a.flags.writeable:
            a = a.copy()
        with torch.cuda.stream(self.stream):
            if a.size == 0:
                return torch.empty(0, dtype=torch.from_numpy(a).dtype, device=self.device)
            t = torch.from_numpy(a)
            return t.to(self.device, non_blocking=False)

    def to_host(self, t, dtype=None):
        """CUDA tensor -> numpy array.  Large results travel through pinned memory (torch's caching host
        allocator recycles the blocks; the array keeps its block alive): ~50 GB/s instead of a pageable copy."""
        torch = _torch()
        with torch.cuda.stream(self.stream):
            if t.numel() * t.element_size() >= (1 << 16):
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t, non_blocking=True)
                self.stream.synchronize()
                a = h.numpy()
            else:
                a = t.cpu().numpy()
        return a.view(dtype) if dtype is not None else a

    @staticmethod
    def ptr(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def hand_over(self, *tensors):
        """Device tensors produced on the context's stream are about to be returned to the caller: order the caller's
        current stream after the work that fills them and keep the allocator from recycling them early (ADVICE r1:
        the context stream does not synchronise with the default stream).  -> the tensors"""
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        if cur != self.stream:
            cur.wait_stream(self.stream)
            for t in tensors:
                if t is not None and hasattr(t, "record_stream"):
                    t.record_stream(cur)
        return tensors[0] if len(tensors) == 1 else tensors

    def sync(self):
        check(self._L.sgc_ctx_sync(self._h))

    @property
    def launches(self):
        return int(self._L.sgc_ctx_launches(self._h))

    # ---- K1 --------------------------------------------------------------
    def hash_batch(self, seq, off, ksize, want_status=True, to_host=True):
        """Hash every k-mer of every sequence.  Returns (hashes u64, kmer_off i64, status i32)."""
        torch = _torch()
        d_seq = self.to_device(seq, np.uint8)
        d_off = self.to_device(off, np.int64)
        nseq = d_off.numel() - 1
        nbytes = d_seq.numel()
        cap = max(nbytes, 1)  # total k-mers <= total bases
        d_hash = self.empty(cap, torch.int64)
        d_koff = self.empty(nseq + 1, torch.int64)
        d_stat = self.empty(nseq, torch.int32) if want_status else None
        total = ctypes.c_int64(0)
        check(
            self._L.sgc_hash_batch(
                self._h, self.ptr(d_seq), nbytes, self.ptr(d_off), nseq, int(ksize), self.ptr(d_hash), cap,
                self.ptr(d_koff), self.ptr(d_stat) if d_stat is not None else None, ctypes.byref(total),
            )
        )
        d_hash = d_hash[: total.value]
        if not to_host:
            self.sync()
            return d_hash, d_koff, d_stat
        return (
            self.to_host(d_hash, np.uint64),
            self.to_host(d_koff),
            self.to_host(d_stat) if d_stat is not None else None,
        )

    def hash_positions(self, seq, off, ksize):
        """Per k-mer of every sequence: hash and (byte offset of its first base in ``seq`` | strand bit << 30), on
        the device (partitioned build with a sequence store).  -> (hashes i64, aux i32, kmer_off i64, status i32)"""
        torch = _torch()
        d_seq = self.to_device(seq, np.uint8)
        d_off = self.to_device(off, np.int64)
        nseq = d_off.numel() - 1
        nbytes = d_seq.numel()
        cap = max(nbytes, 1)
        d_hash = self.empty(cap, torch.int64)
        d_aux = self.empty(cap, torch.int32)
        d_koff = self.empty(nseq + 1, torch.int64)
        d_stat = self.empty(max(nseq, 1), torch.int32)
        total = ctypes.c_int64(0)
        check(self._L.sgc_hash_positions(self._h, self.ptr(d_seq), nbytes, self.ptr(d_off), nseq, int(ksize),
                                         self.ptr(d_hash), self.ptr(d_aux), cap, self.ptr(d_koff), self.ptr(d_stat),
                                         ctypes.byref(total)))
        self.sync()
        return d_hash[: total.value], d_aux[: total.value], d_koff, d_stat[:nseq]

    def pack_reads(self, seq, off):
        """ASCII reads on the device (or host arrays, uploaded) -> 2-bit packed records on the device.
        -> (packed uint8 tensor, lengths int32 tensor, status int32 tensor)"""
        torch = _torch()
        d_seq = self.to_device(seq, np.uint8)
        d_off = self.to_device(off, np.int64)
        nseq = d_off.numel() - 1
        cap = (d_seq.numel() // 4 + 4 * nseq + 16 + 15) // 16 * 16  # every record rounds up to 4 bytes
        d_packed = self.empty(cap, torch.uint8)
        d_len = self.empty(max(nseq, 1), torch.int32)
        d_stat = self.empty(max(nseq, 1), torch.int32)
        nb = ctypes.c_int64(0)
        check(self._L.sgc_pack_reads_device(self._h, self.ptr(d_seq), d_seq.numel(), self.ptr(d_off), nseq,
                                            self.ptr(d_packed), cap, self.ptr(d_len), self.ptr(d_stat), ctypes.byref(nb)))
        return d_packed[: nb.value], d_len[:nseq], d_stat[:nseq]

    # ---- K6 --------------------------------------------------------------
    def count_doms(self, d_count_by_id, member_off, member, level_nodes, level_off, n_nodes):
        torch = _torch()
        d_moff = self.to_device(member_off, np.int64)
        d_mem = self.to_device(member, np.uint32)
        d_lvl = self.to_device(level_nodes, np.uint32)
        h_loff = np.ascontiguousarray(level_off, np.int64)
        d_out = self.empty(n_nodes, torch.int32)
        check(
            self._L.sgc_count_doms(
                self._h, self.ptr(d_count_by_id), self.ptr(d_moff), self.ptr(d_mem), self.ptr(d_lvl),
                h_loff.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(h_loff) - 1, self.ptr(d_out), int(n_nodes),
            )
        )
        return d_out


def get_context(device=None):
    """Process-wide context for a device (default: LOCAL_RANK or 0)."""
    import os

    if device is None:
        device = int(os.environ.get("SGC_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


class DeviceTable:
    """GPU-resident exact map u64 k-mer hash -> u32 unitig id (``sgc_table``)."""

    def __init__(self, ctx, max_entries, home_shift=0, load=0.0, shared=False):
        self.ctx = ctx
        self._L = ctx._L
        h = ctypes.c_void_p()
        create = self._L.sgc_table_create_shared if shared else self._L.sgc_table_create  # shared: mappable by peers
        check(create(ctx._h, int(max_entries), float(load), ctypes.byref(h)))
        self._h = h
        self.max_entries = int(max_entries)
        if home_shift:
            check(self._L.sgc_table_set_home_shift(self._h, int(home_shift)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.sgc_table_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nbytes(self):
        return int(self._L.sgc_table_bytes(self._h))

    # ---- K2 --------------------------------------------------------------
    def insert_pairs(self, hashes, ids):
        c = self.ctx
        d_h = c.to_device(hashes, np.uint64)
        d_v = c.to_device(ids, np.uint32)
        if d_h.numel() != d_v.numel():
            raise ValueError("hashes and ids differ in length")
        check(self._L.sgc_table_insert_pairs(self._h, c.ptr(d_h), c.ptr(d_v), d_h.numel()))

    def insert_triples(self, hashes, ids, aux, id_bounds, marks_cap=None):
        """Owner side of a partitioned build with a sequence store: insert (hash, id, position | strand) triples.
        -> int64 CUDA tensor of break marks (shard << 32 | position) for the replicated store."""
        torch = _torch()
        c = self.ctx
        d_h = c.to_device(hashes, np.uint64)
        d_v = c.to_device(ids, np.uint32)
        d_a = c.to_device(aux, np.uint32)
        n = d_h.numel()
        if d_v.numel() != n or d_a.numel() != n:
            raise ValueError("hashes, ids and aux differ in length")
        # repeated hashes are rare in a cDBG (multicounts); a list that overflows cannot be redone in place (the
        # table has changed), so it is sized generously and an overflow is an error
        cap = int(marks_cap) if marks_cap is not None else max(1 << 20, n // 8)
        bounds = (ctypes.c_uint32 * len(id_bounds))(*[int(b) for b in id_bounds])
        d_m = c.empty(cap, torch.int64)
        nm = ctypes.c_int64(0)
        check(self._L.sgc_table_insert_triples(self._h, c.ptr(d_h), c.ptr(d_v), c.ptr(d_a), n, bounds,
                                               len(id_bounds) - 1, c.ptr(d_m), cap, ctypes.byref(nm)))
        return d_m[: nm.value]

    def bucket_ptr(self):
        """-> (device pointer of the bucket array, mapped bytes)."""
        ptr, nb = ctypes.c_void_p(), ctypes.c_int64(0)
        check(self._L.sgc_table_export_fd(self._h, ctypes.byref(ptr), ctypes.byref(nb), None))
        return ptr.value, nb.value

    def export_fd(self):
        """-> (pointer, mapped bytes, a NEW file descriptor of the bucket array) of a ``shared=True`` table; pass
        the fd to the peer processes (``socket.send_fds``) and close it."""
        ptr, nb, fd = ctypes.c_void_p(), ctypes.c_int64(0), ctypes.c_int(-1)
        check(self._L.sgc_table_export_fd(self._h, ctypes.byref(ptr), ctypes.byref(nb), ctypes.byref(fd)))
        return ptr.value, nb.value, fd.value

    def set_peers(self, bucket_ptrs, id_bounds, useq_ptr, useq_stride_words, brk_ptr, brk_stride_words):
        """Probe ``bucket_ptrs[owner(hash)]`` (peer-mapped bucket arrays of every rank, this one included) and match
        against the replicated sequence store from now on (sgc_table_set_peers); the caller keeps the store alive."""
        n = len(bucket_ptrs)
        ptrs = (ctypes.c_void_p * n)(*[int(p) for p in bucket_ptrs])
        bounds = (ctypes.c_uint32 * (n + 1))(*[int(b) for b in id_bounds])
        check(self._L.sgc_table_set_peers(self._h, n, ptrs, bounds, ctypes.c_void_p(int(useq_ptr)), int(useq_stride_words),
                                          ctypes.c_void_p(int(brk_ptr)), int(brk_stride_words)))

    def set_filter(self, words_ptr, filt_bits, my_rank):
        """Give this peer-probed table the replicated miss filter (sgc_table_set_filter); words_ptr = 0 turns it off."""
        check(self._L.sgc_table_set_filter(self._h, ctypes.c_void_p(int(words_ptr)) if words_ptr else None, int(filt_bits),
                                           int(my_rank)))

    def enable_side(self):
        """Allocate the unitig-ordered side array (8 B + 1 bit per k-mer); must precede insert_sequences."""
        check(self._L.sgc_table_enable_side(self._h))

    @property
    def side_nbytes(self):
        return int(self._L.sgc_table_side_bytes(self._h))

    def insert_sequences(self, seq, off, ksize, ids=None, id_base=0):
        """Fused hash + insert of every k-mer of every sequence under the sequence's id.
        Returns the per-sequence status array (bit 0: non-ACGT byte seen)."""
        torch = _torch()
        c = self.ctx
        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        nseq = d_off.numel() - 1
        d_ids = c.to_device(ids, np.uint32) if ids is not None else None
        d_stat = c.empty(nseq, torch.int32)
        check(
            self._L.sgc_table_insert_sequences(
                self._h, c.ptr(d_seq), d_seq.numel(), c.ptr(d_off), nseq, int(ksize),
                c.ptr(d_ids) if d_ids is not None else None, int(id_base), c.ptr(d_stat),
            )
        )
        return d_stat

    def stats(self):
        live, multi, mx = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_uint32(0)
        check(self._L.sgc_table_stats(self._h, ctypes.byref(live), ctypes.byref(multi), ctypes.byref(mx)))
        return live.value, multi.value, mx.value

    def export(self, to_host=True):
        """All live (hash, id) pairs sorted by hash -> (uint64[n], uint32[n]) on the host
        (``to_host=False``: int64 / int32 CUDA tensors holding the same bits)."""
        torch = _torch()
        c = self.ctx
        live, _, _ = self.stats()
        d_h = c.empty(max(live, 1), torch.int64)
        d_v = c.empty(max(live, 1), torch.int32)
        n = ctypes.c_int64(0)
        check(self._L.sgc_table_export(self._h, c.ptr(d_h), c.ptr(d_v), live, ctypes.byref(n)))
        if not to_host:
            return c.hand_over(d_h[: n.value], d_v[: n.value])
        return c.to_host(d_h[: n.value], np.uint64), c.to_host(d_v[: n.value], np.uint32)

    # ---- K3 --------------------------------------------------------------
    def lookup(self, hashes, to_host=True):
        torch = _torch()
        c = self.ctx
        d_h = c.to_device(hashes, np.uint64)
        d_o = c.empty(d_h.numel(), torch.int32)
        check(self._L.sgc_lookup(self._h, c.ptr(d_h), d_h.numel(), c.ptr(d_o)))
        if not to_host:
            return c.hand_over(d_o)
        return c.to_host(d_o, np.uint32)

    # ---- K5 --------------------------------------------------------------
    def count_matches(self, hashes, n_ids, distinct=False, to_host=True):
        """-> (count_by_id u32[n_ids], n_miss, n_distinct)"""
        torch = _torch()
        c = self.ctx
        d_h = c.to_device(hashes, np.uint64)
        d_cnt = c.empty(n_ids, torch.int32)
        miss, nd = ctypes.c_int64(0), ctypes.c_int64(0)
        check(
            self._L.sgc_count_matches(
                self._h, c.ptr(d_h), d_h.numel(), 1 if distinct else 0, c.ptr(d_cnt), int(n_ids),
                ctypes.byref(miss), ctypes.byref(nd),
            )
        )
        cnt = c.to_host(d_cnt, np.uint32) if to_host else d_cnt
        return cnt, miss.value, nd.value

    def count_matches_batch(self, seq, off, seq_query, n_queries, ksize):
        """Batched ``Query`` + ``count_cdbg_matches``: sequence ``s`` belongs to query ``seq_query[s]`` (grouped).
        -> dict(keys u64 (query << 32 | unitig id, ascending), counts u32, q_distinct i64[n_queries],
        status i32[nseq], n_kmers, n_hits): per query the number of its DISTINCT k-mer hashes on each unitig
        (search/query_by_sequence.py:170-176, :189)."""
        torch = _torch()
        c = self.ctx
        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        d_sq = c.to_device(seq_query, np.int32)
        nseq = d_off.numel() - 1
        d_qd = c.empty(max(int(n_queries), 1), torch.int64)
        d_stat = c.empty(max(nseq, 1), torch.int32)
        cap = max(d_seq.numel() // 16, 1 << 16)
        for _ in range(6):
            d_keys = c.empty(cap, torch.int64)
            d_cnt = c.empty(cap, torch.int32)
            npairs, nk, nh = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            try:
                check(self._L.sgc_query_count_batch(self._h, c.ptr(d_seq), d_seq.numel(), c.ptr(d_off), nseq, c.ptr(d_sq),
                                                    int(n_queries), int(ksize), c.ptr(d_keys), c.ptr(d_cnt), cap, c.ptr(d_qd),
                                                    c.ptr(d_stat), ctypes.byref(npairs), ctypes.byref(nk), ctypes.byref(nh)))
                break
            except SgcCapacityError:
                cap = int(npairs.value) + 1024
        else:
            raise RuntimeError("count_matches_batch: could not size the output")
        n = npairs.value
        return {"keys": c.to_host(d_keys[:n], np.uint64), "counts": c.to_host(d_cnt[:n], np.uint32),
                "q_distinct": c.to_host(d_qd[: int(n_queries)]), "status": c.to_host(d_stat[:nseq]),
                "n_kmers": nk.value, "n_hits": nh.value}

    def lookup_sequences(self, seq, off, ksize, n_ids=0, want_ids=True, to_host=True):
        """Fused hash + lookup of every k-mer.  -> dict(kmer_ids, kmer_off, count_by_id, status, n_kmers, n_hits)"""
        torch = _torch()
        c = self.ctx
        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        nseq = d_off.numel() - 1
        nbytes = d_seq.numel()
        cap = max(nbytes, 1)
        d_ids = c.empty(cap, torch.int32) if want_ids else None
        d_koff = c.empty(nseq + 1, torch.int64)
        d_cnt = c.empty(n_ids, torch.int32) if n_ids else None
        d_stat = c.empty(nseq, torch.int32)
        nk, nh = ctypes.c_int64(0), ctypes.c_int64(0)
        check(
            self._L.sgc_lookup_sequences(
                self._h, c.ptr(d_seq), nbytes, c.ptr(d_off), nseq, int(ksize),
                c.ptr(d_ids) if d_ids is not None else None, cap, c.ptr(d_koff),
                c.ptr(d_cnt) if d_cnt is not None else None, int(n_ids), c.ptr(d_stat),
                ctypes.byref(nk), ctypes.byref(nh),
            )
        )
        out = {"n_kmers": nk.value, "n_hits": nh.value}
        if d_ids is not None:
            d_ids = d_ids[: nk.value]
        if to_host:
            out["kmer_ids"] = c.to_host(d_ids, np.uint32) if d_ids is not None else None
            out["kmer_off"] = c.to_host(d_koff)
            out["count_by_id"] = c.to_host(d_cnt, np.uint32) if d_cnt is not None else None
            out["status"] = c.to_host(d_stat)
        else:
            c.hand_over(d_ids, d_koff, d_cnt, d_stat)
            out.update(kmer_ids=d_ids, kmer_off=d_koff, count_by_id=d_cnt, status=d_stat)
        return out

    # ---- K4 --------------------------------------------------------------
    def label_reads(self, seq, off, ksize, ids_cap=None, to_host=True, packed=False):
        """index_reads' per-read step: distinct unitig ids per sequence as CSR.
        -> dict(ids_off i64[nseq+1], ids u32[n_labels], status, n_kmers, n_hits)
        ``packed=True`` first packs the reads 2 bits per base on the device and labels the packed
        records (the path the streaming feeder uses; same labels)."""
        torch = _torch()
        c = self.ctx
        d_seq = c.to_device(seq, np.uint8)
        d_off = c.to_device(off, np.int64)
        nseq = d_off.numel() - 1
        if packed:
            d_packed, d_len, d_stat = c.pack_reads(d_seq, d_off)
            out = self.label_reads_packed(d_packed, nseq, ksize, lengths=d_len, ids_cap=ids_cap, to_host=to_host)
            out["status"] = c.to_host(d_stat) if to_host else d_stat
            return out
        nbytes = d_seq.numel()
        cap = int(ids_cap) if ids_cap is not None else max(4 * nseq, 1024)
        d_stat = c.empty(nseq, torch.int32)
        d_ioff = c.empty(nseq + 1, torch.int64)
        for _ in range(8):
            d_ids = c.empty(cap, torch.int32)
            nk, nh, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            try:
                check(
                    self._L.sgc_label_reads(
                        self._h, c.ptr(d_seq), nbytes, c.ptr(d_off), nseq, int(ksize), c.ptr(d_ioff), c.ptr(d_ids),
                        cap, c.ptr(d_stat), ctypes.byref(nk), ctypes.byref(nh), ctypes.byref(nl),
                    )
                )
                break
            except SgcCapacityError:
                cap = max(2 * cap, int(nl.value) + 1024)
        else:
            raise RuntimeError("label_reads: could not size the label buffer")
        d_ids = d_ids[: nl.value]
        out = {"n_kmers": nk.value, "n_hits": nh.value, "n_labels": nl.value}
        if to_host:
            out.update(ids_off=c.to_host(d_ioff), ids=c.to_host(d_ids, np.uint32), status=c.to_host(d_stat))
        else:
            c.hand_over(d_ioff, d_ids, d_stat)
            out.update(ids_off=d_ioff, ids=d_ids, status=d_stat)
        return out

    def label_reads_packed(self, packed, nseq, ksize, lengths=None, uniform_len=-1, ids_cap=None, to_host=True):
        """The same step on 2-bit packed records (``pack_reads_host`` / ``Context.pack_reads`` layout).
        ``lengths``: int32 per record, or None with ``uniform_len`` when all records have that many bases."""
        torch = _torch()
        c = self.ctx
        d_packed = c.to_device(packed, np.uint8)
        d_len = c.to_device(lengths, np.int32) if lengths is not None else None
        nseq = int(nseq)
        cap = int(ids_cap) if ids_cap is not None else max(4 * nseq, 1024)
        d_ioff = c.empty(nseq + 1, torch.int64)
        for _ in range(8):
            d_ids = c.empty(cap, torch.int32)
            nk, nh, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            try:
                check(
                    self._L.sgc_label_reads_packed(
                        self._h, c.ptr(d_packed), d_packed.numel(), c.ptr(d_len) if d_len is not None else None,
                        int(uniform_len), nseq, int(ksize), c.ptr(d_ioff), c.ptr(d_ids), cap,
                        ctypes.byref(nk), ctypes.byref(nh), ctypes.byref(nl),
                    )
                )
                break
            except SgcCapacityError:
                cap = max(2 * cap, int(nl.value) + 1024)
        else:
            raise RuntimeError("label_reads_packed: could not size the label buffer")
        d_ids = d_ids[: nl.value]
        out = {"n_kmers": nk.value, "n_hits": nh.value, "n_labels": nl.value}
        if to_host:
            out.update(ids_off=c.to_host(d_ioff), ids=c.to_host(d_ids, np.uint32))
        else:
            c.hand_over(d_ioff, d_ids)
            out.update(ids_off=d_ioff, ids=d_ids)
        return out

    # ---- K4, streaming ------------------------------------------------------
    def label_stream(self, batches, ksize, ids_per_read=6, prefetch=True):
        """Label a stream of HOST batches with the uploads, the kernels and the downloads of consecutive
        batches overlapped (the loop of cdbg/index_reads.py:105-167 with the per-read arithmetic on the GPU):
        while the kernels of batch i run, batch i+1 is copied to the device on a copy stream and the labels
        of batch i-1 return to pinned host memory on a third stream.  Memory is bounded by two batches.

        ``batches`` yields dicts, either ``{"packed": uint8, "nseq": n, "lengths": int32 or None,
        "uniform_len": L}`` (2-bit packed records, :func:`pack_reads_host` layout; the fast path) or
        ``{"seq": uint8, "off": int64}`` (ASCII).  Arrays may be numpy arrays or torch CPU tensors; pinned
        tensors are copied asynchronously.  Extra keys are passed through.
        Yields ``(batch, result)`` in order, ``result = dict(ids_off, ids, n_kmers, n_hits, n_labels[,
        status])`` with numpy views of pinned buffers that stay valid until the NEXT-BUT-ONE batch is
        yielded (copy what must live longer).  With ``prefetch`` the batches are pulled from the iterable on
        a helper thread, so the host-side preparation of batch i+1 (inflate, scan, pack: native code that
        releases the GIL) overlaps the GPU work of batch i."""
        torch = _torch()
        c = self.ctx
        L = self._L
        if getattr(self, "_stream_state", None) is None:
            self._stream_state = ([_StreamSlot(), _StreamSlot()], torch.cuda.Stream(device=c.device),
                                  torch.cuda.Stream(device=c.device))
        slots, copy_stream, back_stream = self._stream_state
        import os as _os

        if _os.environ.get("SGC_STREAM_SERIAL"):  # experiment knob: uploads and downloads on ONE copy stream
            back_stream = copy_stream
        no_d2h = bool(_os.environ.get("SGC_STREAM_NO_D2H"))  # experiment knob: results stay on the device (timing only)
        for s in slots:
            s.copied, s.returned = torch.cuda.Event(), torch.cuda.Event()
            s.returned.record(back_stream)

        def as_tensor(a, dtype):
            if isinstance(a, torch.Tensor):
                return a
            a = np.ascontiguousarray(a)
            if a.dtype == np.uint32:
                a = a.view(np.int32)
            t = torch.from_numpy(a)
            assert t.dtype == dtype, (t.dtype, dtype)
            return t

        def upload(slot, b):
            with torch.cuda.stream(copy_stream):
                if "packed" in b:
                    hp = as_tensor(b["packed"], torch.uint8)
                    slot.dev(c, "packed", hp.numel(), torch.uint8)[: hp.numel()].copy_(hp, non_blocking=True)
                    if b.get("lengths") is not None:
                        hl = as_tensor(b["lengths"], torch.int32)
                        slot.dev(c, "len", hl.numel(), torch.int32)[: hl.numel()].copy_(hl, non_blocking=True)
                else:
                    hs, ho = as_tensor(b["seq"], torch.uint8), as_tensor(b["off"], torch.int64)
                    slot.dev(c, "seq", hs.numel(), torch.uint8)[: hs.numel()].copy_(hs, non_blocking=True)
                    slot.dev(c, "off", ho.numel(), torch.int64)[: ho.numel()].copy_(ho, non_blocking=True)
                slot.copied.record(copy_stream)

        def nseq_of(b):
            return int(b["nseq"]) if "packed" in b else int(len(b["off"]) - 1)

        def launch(slot, b, cap):
            n = nseq_of(b)
            d_ioff = slot.dev(c, "ids_off", n + 1, torch.int64)
            d_ids = slot.dev(c, "ids", cap, torch.int32)
            if "packed" in b:
                nb = int(b["packed"].numel() if isinstance(b["packed"], torch.Tensor) else len(b["packed"]))
                d_len = slot.d_in["len"] if b.get("lengths") is not None else None
                check(L.sgc_label_reads_packed_launch(self._h, c.ptr(slot.d_in["packed"]), nb,
                                                      c.ptr(d_len) if d_len is not None else None,
                                                      int(b.get("uniform_len", -1)), n, int(ksize), c.ptr(d_ioff),
                                                      c.ptr(d_ids), d_ids.numel()))
            else:
                nb = int(b["seq"].numel() if isinstance(b["seq"], torch.Tensor) else len(b["seq"]))
                d_stat = slot.dev(c, "status", max(n, 1), torch.int32)
                check(L.sgc_label_reads_launch(self._h, c.ptr(slot.d_in["seq"]), nb, c.ptr(slot.d_in["off"]), n,
                                               int(ksize), c.ptr(d_ioff), c.ptr(d_ids), d_ids.numel(), c.ptr(d_stat)))

        def finish(slot, b):
            nk, nh, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
            try:
                check(L.sgc_label_reads_finish(self._h, ctypes.byref(nk), ctypes.byref(nh), ctypes.byref(nl)))
            except SgcCapacityError:  # more labels than the output holds: size it for the real count and redo
                launch(slot, b, int(nl.value) + 1024)
                check(L.sgc_label_reads_finish(self._h, ctypes.byref(nk), ctypes.byref(nh), ctypes.byref(nl)))
            return nk.value, nh.value, nl.value

        def download(slot, b, stats):
            n, nl = nseq_of(b), stats[2]
            # the offsets travel as one byte per read (its number of labels) and are rebuilt on the host by
            # sgc_counts8_to_offsets (threads): 1 instead of 8 bytes per read back over PCIe
            d_cnt8 = slot.dev(c, "cnt8", n + 1, torch.uint8)
            check(L.sgc_offsets_to_counts8(c._h, c.ptr(slot.d_in["ids_off"]), n, c.ptr(d_cnt8)))
            done = torch.cuda.Event()
            done.record(c.stream)
            h_ioff = slot.host(torch, "ids_off", n + 1, torch.int64)
            h_cnt8 = slot.host(torch, "cnt8", n + 1, torch.uint8)
            h_ids = slot.host(torch, "ids", max(nl, 1), torch.int32)
            h_stat = slot.host(torch, "status", max(n, 1), torch.int32) if "packed" not in b else None
            with torch.cuda.stream(back_stream):
                back_stream.wait_event(done)
                h_cnt8[: n + 1].copy_(d_cnt8[: n + 1], non_blocking=True)
                if nl and not no_d2h:
                    h_ids[:nl].copy_(slot.d_in["ids"][:nl], non_blocking=True)
                if h_stat is not None and n:
                    h_stat[:n].copy_(slot.d_in["status"][:n], non_blocking=True)
                slot.returned.record(back_stream)
            res = {"n_kmers": stats[0], "n_hits": stats[1], "n_labels": nl,
                   "ids_off": h_ioff[: n + 1].numpy(), "ids": h_ids[:nl].numpy().view(np.uint32)}
            if h_stat is not None:
                res["status"] = h_stat[:n].numpy()

            def rebuild():  # after slot.returned: bytes -> offsets (or, a read with > 255 labels: the offsets themselves)
                if int(h_cnt8[n]):
                    with torch.cuda.stream(back_stream):
                        h_ioff[: n + 1].copy_(slot.d_in["ids_off"][: n + 1], non_blocking=True)
                    back_stream.synchronize()
                else:
                    check(L.sgc_counts8_to_offsets(ctypes.c_void_p(h_cnt8.data_ptr()), n, 0, ctypes.c_void_p(h_ioff.data_ptr())))

            res["_rebuild"] = rebuild
            return res

        it = _prefetch_iter(batches) if prefetch else iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        upload(slots[0], cur)
        prev, idx = None, 0
        while cur is not None:
            slot = slots[idx % 2]
            c.stream.wait_event(slot.copied)
            c.stream.wait_event(slot.returned)  # the labels of batch idx-2 have left this slot's output buffers
            launch(slot, cur, max(int(ids_per_read * nseq_of(cur)), 1024))
            nxt = next(it, None)
            if nxt is not None:
                upload(slots[(idx + 1) % 2], nxt)  # that slot's kernels (batch idx-1) have finished: finish() synchronised
            if prev is not None:
                prev[0].returned.synchronize()
                prev[2].pop("_rebuild")()
                yield prev[1], prev[2]  # the consumer works on batch idx-1 while the GPU labels batch idx
            stats = finish(slot, cur)
            prev = (slot, cur, download(slot, cur, stats))
            cur, idx = nxt, idx + 1
        prev[0].returned.synchronize()
        prev[2].pop("_rebuild")()
        yield prev[1], prev[2]


def _prefetch_iter(iterable, depth=2):
    """Pull items from ``iterable`` on a helper thread, ``depth`` ahead (the producer's native calls release the GIL)."""
    import queue
    import threading

    q = queue.Queue(maxsize=depth)
    end = object()

    def run():
        try:
            for x in iterable:
                q.put(x)
            q.put(end)
        except BaseException as e:  # noqa: BLE001  (re-raised in the consumer)
            q.put(e)

    threading.Thread(target=run, daemon=True).start()
    while True:
        x = q.get()
        if x is end:
            return
        if isinstance(x, BaseException):
            raise x
        yield x


class _StreamSlot:
    """Device input/output buffers of one in-flight batch of :meth:`DeviceTable.label_stream`."""

    def __init__(self):
        self.d_in = {}      # name -> CUDA tensor (grown on demand)
        self.h_out = {}     # name -> pinned host tensor
        self.copied = None  # H2D of this slot's inputs has finished
        self.returned = None  # D2H of this slot's previous outputs has finished

    def dev(self, ctx, name, n, dtype):
        t = self.d_in.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = self.d_in[name] = ctx.empty(max(int(n * 1.25) + 16, 16), dtype)
        return t

    def host(self, torch, name, n, dtype):
        t = self.h_out.get(name)
        if t is None or t.numel() < n:
            t = self.h_out[name] = torch.empty(max(int(n * 1.25) + 16, 16), dtype=dtype).pin_memory()
        return t


class DeviceMphf:
    """BBHash (boomphf) minimal perfect hash built on the GPU (``sgc_mphf``): the writer of the
    reference's ``contigs.mphf`` (bbhash.PyMPHF + save; reference cdbg/hashing.py:8, the
    BBHashTable it instantiates at index_cdbg_by_kmer.py:63-64)."""

    NOT_FOUND = 0xFFFFFFFFFFFFFFFF

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self._L = ctx._L
        self._h = handle

    @classmethod
    def build(cls, ctx, keys, gamma=1.0, nb_levels=25):
        """keys: distinct uint64 hashes (numpy array or CUDA int64 tensor)."""
        d_k = ctx.to_device(keys, np.uint64)
        h = ctypes.c_void_p()
        check(ctx._L.sgc_mphf_build(ctx._h, ctx.ptr(d_k), d_k.numel(), float(gamma), int(nb_levels), ctypes.byref(h)))
        return cls(ctx, h)

    @classmethod
    def loads(cls, ctx, raw):
        """Parse the bytes of a boomphf dump (a reference ``contigs.mphf``)."""
        raw = bytes(raw)
        h = ctypes.c_void_p()
        check(ctx._L.sgc_mphf_load(ctx._h, ctypes.c_char_p(raw), len(raw), ctypes.byref(h)))
        return cls(ctx, h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.sgc_mphf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        a = (ctypes.c_int64 * 4)()
        g = ctypes.c_double(0)
        check(self._L.sgc_mphf_info(self._h, a, ctypes.byref(g)))
        return {"nelem": a[0], "nb_levels": a[1], "lastbitsetrank": a[2], "n_final": a[3], "gamma": g.value}

    def __len__(self):
        return self.info()["nelem"]

    def lookup(self, keys, to_host=True):
        """MPHF index of every key (NOT_FOUND, or an arbitrary index, for keys outside the build set)."""
        torch = _torch()
        c = self.ctx
        d_k = c.to_device(keys, np.uint64)
        d_o = c.empty(d_k.numel(), torch.int64)
        check(self._L.sgc_mphf_lookup(self._h, c.ptr(d_k), d_k.numel(), c.ptr(d_o)))
        return c.to_host(d_o, np.uint64) if to_host else d_o

    def order(self, keys, values):
        """(mphf_to_hash uint64[n], mphf_to_value uint32[n]) for the build keys and their values."""
        torch = _torch()
        c = self.ctx
        d_k = c.to_device(keys, np.uint64)
        d_v = c.to_device(values, np.uint32)
        if d_k.numel() != d_v.numel():
            raise ValueError("keys and values differ in length")
        d_oh = c.empty(d_k.numel(), torch.int64)
        d_ov = c.empty(d_k.numel(), torch.int32)
        check(self._L.sgc_mphf_order(self._h, c.ptr(d_k), c.ptr(d_v), d_k.numel(), c.ptr(d_oh), c.ptr(d_ov)))
        return c.to_host(d_oh, np.uint64), c.to_host(d_ov, np.uint32)

    def dumps(self):
        """The boomphf dump (what ``mphf.save`` writes)."""
        n = ctypes.c_int64(0)
        check(self._L.sgc_mphf_dump_bytes(self._h, ctypes.byref(n)))
        buf = ctypes.create_string_buffer(n.value)
        check(self._L.sgc_mphf_dump(self._h, buf, n.value))
        return buf.raw


__all__ = ["Context", "DeviceTable", "DeviceMphf", "get_context", "pack_sequences", "pack_reads_host", "kmer_counts", "SGC_MISS"]
