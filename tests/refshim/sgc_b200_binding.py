"""The reference-side binding of libsgc_b200.so: what a spacegraphcats maintainer would drop
next to ``spacegraphcats/cdbg/hashing.py`` so that its two third-party imports (:7-8)

    import khmer                              ->  from . import _sgc_b200 as khmer
    from bbhash_table import BBHashTable      ->  from ._sgc_b200 import BBHashTable

resolve to the GPU library.  Raw ctypes over ``include/sgc_b200.h`` only (torch just owns device memory); nothing
of the ``spacegraphcats_b200`` package is used, so this file is an independent client of the C-ABI.  It is
EXECUTED by tests/test_gpu_parity.py::test_reference_side_binding_* (no documentation-only code).
Scalar granularity -- one launch per call -- correct but slow; the batched entry points are what the
replacement modules use (INTEGRATION.md, table of loops)."""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_L = ctypes.CDLL(os.environ.get("SGC_B200_LIB", os.path.join(_HERE, "..", "..", "spacegraphcats_b200", "libsgc_b200.so")))
_L.sgc_last_error.restype = ctypes.c_char_p
vp, i64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint32
UNSET_VALUE = 2**32 - 1


def _ck(rc):
    if rc:
        raise RuntimeError(_L.sgc_last_error().decode())


_ctx = None


def _context():
    global _ctx
    if _ctx is None:
        c = vp()
        _ck(_L.sgc_ctx_create(0, None, ctypes.byref(c)))  # GPU 0, the library's own stream
        _ctx = c
    return _ctx


def _dev(a):
    """numpy array -> CUDA tensor holding the same bits"""
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    elif a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a.copy()).cuda()


def _ptr(t):
    return vp(t.data_ptr())


class Nodetable:
    """stands in for khmer.Nodetable(k, 1, 1) as used by hash_sequence (cdbg/hashing.py:14-20)"""

    def __init__(self, ksize, *_):
        self.k = int(ksize)

    def get_kmer_hashes(self, seq):
        if isinstance(seq, str):
            seq = seq.encode("ascii")
        if len(seq) < self.k:
            raise ValueError(f"sequence length ({len(seq)}) must >= the hashtable k-mer size ({self.k})")
        s = _dev(np.frombuffer(seq, np.uint8))
        off = torch.tensor([0, len(seq)], dtype=torch.int64).cuda()
        n = len(seq) - self.k + 1
        out = torch.empty(n, dtype=torch.int64, device="cuda")
        koff = torch.empty(2, dtype=torch.int64, device="cuda")
        st = torch.zeros(1, dtype=torch.int32, device="cuda")
        tot = i64()
        torch.cuda.synchronize()
        _ck(_L.sgc_hash_batch(_context(), _ptr(s), i64(len(seq)), _ptr(off), i64(1), i32(self.k), _ptr(out), i64(n),
                              _ptr(koff), _ptr(st), ctypes.byref(tot)))
        _ck(_L.sgc_ctx_sync(_context()))
        if int(st[0]):
            raise ValueError("Invalid base in read")
        return out.cpu().numpy().view(np.uint64).tolist()


class BBHashTable:
    """stands in for bbhash_table.BBHashTable (cdbg/hashing.py:34-37,69,115,126; index_cdbg_by_kmer.py:63-64,89,106)"""

    def __init__(self):
        self._t = None
        self._h, self._v = [], []

    def __del__(self):
        if getattr(self, "_t", None):
            _L.sgc_table_free(self._t)
            self._t = None

    def initialize(self, hashes):  # index_cdbg_by_kmer.py:63-64: declare the key set, values unset
        t = vp()
        _ck(_L.sgc_table_create(_context(), i64(max(len(hashes), 1)), ctypes.c_float(0), ctypes.byref(t)))
        self._t = t
        self._h, self._v = [], []

    def __setitem__(self, h, v):  # :89 -- buffered, flushed by the first read access
        self._h.append(int(h))
        self._v.append(int(v))

    def _flush(self):
        if self._h:
            h, v = _dev(np.array(self._h, np.uint64)), _dev(np.array(self._v, np.uint32))
            torch.cuda.synchronize()
            _ck(_L.sgc_table_insert_pairs(self._t, _ptr(h), _ptr(v), i64(len(self._h))))
            self._h, self._v = [], []

    def _stats(self):
        self._flush()
        live, multi, mx = i64(), i64(), u32()
        _ck(_L.sgc_table_stats(self._t, ctypes.byref(live), ctypes.byref(multi), ctypes.byref(mx)))
        return live.value, multi.value, mx.value

    def __len__(self):  # hashing.py:33-34
        return self._stats()[0]

    def __getitem__(self, h):  # hashing.py:36-37
        self._flush()
        q = _dev(np.array([int(h)], np.uint64))
        o = torch.empty(1, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        _ck(_L.sgc_lookup(self._t, _ptr(q), i64(1), _ptr(o)))
        _ck(_L.sgc_ctx_sync(_context()))
        r = int(o[0]) & 0xFFFFFFFF
        return None if r == 0xFFFFFFFF else r

    def get_unique_values(self, hashes, require_exist=False):  # hashing.py:67-69
        live, _, mx = self._stats()
        h = np.fromiter((int(x) for x in hashes), np.uint64)
        if len(h) == 0:
            return {}
        n_ids = mx + 1 if live else 1
        d_h = _dev(h)
        cnt = torch.empty(n_ids, dtype=torch.int32, device="cuda")
        miss, nd = i64(), i64()
        torch.cuda.synchronize()
        _ck(_L.sgc_count_matches(self._t, _ptr(d_h), i64(len(h)), i32(0), _ptr(cnt), i64(n_ids), ctypes.byref(miss),
                                 ctypes.byref(nd)))
        if require_exist and miss.value:
            raise ValueError("hashval not found.")
        c = cnt.cpu().numpy().view(np.uint32)
        nz = np.flatnonzero(c)
        return dict(zip(nz.tolist(), c[nz].tolist()))

    def _export(self):
        live = self._stats()[0]
        d_h = torch.empty(max(live, 1), dtype=torch.int64, device="cuda")
        d_v = torch.empty(max(live, 1), dtype=torch.int32, device="cuda")
        n = i64()
        torch.cuda.synchronize()
        _ck(_L.sgc_table_export(self._t, _ptr(d_h), _ptr(d_v), i64(live), ctypes.byref(n)))
        return d_h[: n.value], d_v[: n.value]

    @property
    def mphf_to_hash(self):
        return self._export()[0].cpu().numpy().view(np.uint64)

    @property
    def mphf_to_value(self):  # index_cdbg_by_kmer.py:106 takes max() of it
        return self._export()[1].cpu().numpy().view(np.uint32)

    def save(self, mphf_filename, array_filename):  # hashing.py:115: boomphf dump + arrays in MPHF order
        d_h, d_v = self._export()
        n = d_h.numel()
        m = vp()
        _ck(_L.sgc_mphf_build(_context(), _ptr(d_h), i64(n), ctypes.c_double(1.0), i32(25), ctypes.byref(m)))
        try:
            oh = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
            ov = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            if n:
                _ck(_L.sgc_mphf_order(m, _ptr(d_h), _ptr(d_v), i64(n), _ptr(oh), _ptr(ov)))
            nb = i64()
            _ck(_L.sgc_mphf_dump_bytes(m, ctypes.byref(nb)))
            buf = ctypes.create_string_buffer(nb.value)
            _ck(_L.sgc_mphf_dump(m, buf, nb))
        finally:
            _L.sgc_mphf_free(m)
        with open(array_filename, "wb") as fp:
            np.savez_compressed(fp, mphf_to_hash=oh[:n].cpu().numpy().view(np.uint64),
                                mphf_to_value=ov[:n].cpu().numpy().view(np.uint32))
        with open(mphf_filename, "wb") as fp:
            fp.write(buf.raw)

    @classmethod
    def load(cls, mphf_filename, array_filename):  # hashing.py:126
        with np.load(array_filename) as z:
            h, v = z["mphf_to_hash"], z["mphf_to_value"]
        keep = v != UNSET_VALUE
        t = cls()
        t.initialize(h[keep])
        t._h, t._v = h[keep].tolist(), v[keep].tolist()
        t._flush()
        return t
