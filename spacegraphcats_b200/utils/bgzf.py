"""Block-gzip (BGZF) support for read labelling: whole-file block inflate plus the mapping from
uncompressed positions to BGZF *virtual offsets* ``(block_start << 16) | within_block`` -- the
``offset`` column index_reads writes (reference: cdbg/index_reads.py:105, search_utils.py:67-154
on top of utils/bgzf/bgzf.py ``BgzfReader.tell`` :596-611).  A position is always reported inside
the block that holds its byte (a position at a block end belongs to the NEXT block, offset 0),
which is what the reference's reader returns after ``readline``.

The heavy lifting is native and multi-threaded (``csrc/sgc_feed.hpp`` behind ``sgc_bgzf_*`` /
``sgc_reads_*``): parallel block inflate with CRC check, record scan straight into the packed
``(seq, off)`` layout the kernels take, names kept as byte ranges instead of Python strings.
"""
import ctypes
import os

import numpy as np

from .._lib import check, lib



def is_bgzf(path):
    with open(path, "rb") as fp:
        head = fp.read(18)
    return len(head) >= 18 and head[:4] == b"\x1f\x8b\x08\x04" and head[12:14] == b"BC"


def n_threads():
    """Host threads for the feeder: SGC_FEED_THREADS, else the CPUs this process may run on."""
    env = os.environ.get("SGC_FEED_THREADS")
    if env:
        return max(int(env), 1)
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return max(os.cpu_count() or 1, 1)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class BgzfData:
    """A fully inflated BGZF file: ``data`` (uint8) and position -> virtual offset."""

    def __init__(self, path):
        L = lib()
        raw = np.fromfile(path, np.uint8)
        nb, total = ctypes.c_int64(0), ctypes.c_int64(0)
        check(L.sgc_bgzf_index(_p(raw), raw.size, None, None, 0, ctypes.byref(nb), ctypes.byref(total)))
        cstart = np.zeros(nb.value, np.int64)
        usize = np.zeros(nb.value, np.int64)
        check(L.sgc_bgzf_index(_p(raw), raw.size, _p(cstart), _p(usize), nb.value, ctypes.byref(nb), ctypes.byref(total)))
        self.data = np.empty(total.value, np.uint8)
        check(L.sgc_bgzf_inflate(_p(raw), raw.size, _p(cstart), _p(usize), nb.value, n_threads(), _p(self.data),
                                 total.value))
        keep = usize > 0  # empty blocks (the EOF marker) hold no position
        self._cstart = cstart[keep]
        self._ustart = (np.cumsum(usize) - usize)[keep]

    def virtual_offsets(self, positions):
        positions = np.asarray(positions, np.int64)
        b = np.searchsorted(self._ustart, positions, side="right") - 1
        return (self._cstart[b] << 16) | (positions - self._ustart[b])


class PlainData:
    """Uncompressed / plain-gzip input: offsets are plain uncompressed positions."""

    def __init__(self, path):
        import gzip

        with open(path, "rb") as fp:
            raw = fp.read()
        if raw[:2] == b"\x1f\x8b":
            raw = gzip.decompress(raw)
        self.data = np.frombuffer(raw, np.uint8)

    def virtual_offsets(self, positions):
        return np.asarray(positions, np.int64)


def open_reads(path):
    return BgzfData(path) if is_bgzf(path) else PlainData(path)



class ReadStream:
    """Bounded-memory walk over a read file: ``chunks()`` yields ``(data uint8, gstart)`` pieces that each hold
    only COMPLETE records (``gstart`` = uncompressed position of ``data[0]`` in the file), ``virtual_offsets``
    maps uncompressed positions to what the reference's reader reports (BGZF virtual offsets, or plain positions).

    BGZF files are memory-mapped and inflated ``chunk_bytes`` of blocks at a time (native, multi-threaded); the
    unfinished record at the end of a chunk is carried over to the next one.  Plain / gzip files are one chunk."""

    def __init__(self, path, chunk_bytes=512 << 20):
        self.path, self.chunk_bytes = path, int(chunk_bytes)
        self.bgzf = is_bgzf(path)
        if self.bgzf:
            L = lib()
            self._raw = np.memmap(path, np.uint8, mode="r")
            nb, total = ctypes.c_int64(0), ctypes.c_int64(0)
            check(L.sgc_bgzf_index(_p(self._raw), self._raw.size, None, None, 0, ctypes.byref(nb), ctypes.byref(total)))
            self._cstart_all = np.zeros(nb.value, np.int64)
            self._usize_all = np.zeros(nb.value, np.int64)
            check(L.sgc_bgzf_index(_p(self._raw), self._raw.size, _p(self._cstart_all), _p(self._usize_all), nb.value,
                                   ctypes.byref(nb), ctypes.byref(total)))
            self.total_bytes = int(total.value)
            keep = self._usize_all > 0  # empty blocks (the EOF marker) hold no position
            self._cstart = self._cstart_all[keep]
            self._ustart = (np.cumsum(self._usize_all) - self._usize_all)[keep]
        else:
            self._plain = PlainData(path)
            self.total_bytes = int(self._plain.data.size)

    def virtual_offsets(self, positions):
        positions = np.asarray(positions, np.int64)
        if not self.bgzf:
            return positions
        b = np.searchsorted(self._ustart, positions, side="right") - 1
        return (self._cstart[b] << 16) | (positions - self._ustart[b])

    def chunks(self):
        if not self.bgzf:
            yield self._plain.data, 0
            return
        L = lib()
        nblocks = len(self._usize_all)
        carry = np.zeros(0, np.uint8)
        b0, gpos = 0, 0  # next block, its uncompressed position
        while b0 < nblocks:
            b1, size = b0, 0
            while b1 < nblocks and (size < self.chunk_bytes or b1 == b0):
                size += int(self._usize_all[b1])
                b1 += 1
            buf = np.empty(len(carry) + size, np.uint8)
            buf[: len(carry)] = carry
            if size:
                check(L.sgc_bgzf_inflate(_p(self._raw), self._raw.size, _p(self._cstart_all[b0:b1]), _p(self._usize_all[b0:b1]),
                                         b1 - b0, n_threads(), _p(buf[len(carry):]), size))
            gstart = gpos - len(carry)
            gpos += size
            b0 = b1
            if b0 >= nblocks:  # end of the file: whatever is left is the last record(s)
                if buf.size:
                    yield buf, gstart
                return
            cut = ctypes.c_int64(0)
            if buf.size:
                try:
                    check(L.sgc_reads_cut(_p(buf), buf.size, n_threads(), ctypes.byref(cut)))
                except Exception as e:
                    if "unknown start chr" in str(e):
                        raise Exception("unknown start chr {}".format(chr(buf[0])))
                    raise
            if cut.value:
                yield buf[: cut.value], gstart
            carry = buf[cut.value:].copy()

class NameTable:
    """Record names as byte ranges ``[lo[i], hi[i])`` of a uint8 buffer: no per-record Python
    object is made until one is asked for.  Compares equal to a list of str."""

    def __init__(self, data, lo, hi):
        self.data, self.lo, self.hi = data, lo, hi

    @classmethod
    def from_strings(cls, names):
        enc = [s.encode("ascii") for s in names]
        ln = np.fromiter((len(e) for e in enc), np.int64, len(enc))
        hi = np.cumsum(ln)
        return cls(np.frombuffer(b"".join(enc) or b"\0", np.uint8), hi - ln, hi)

    def __len__(self):
        return len(self.lo)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return self.data[self.lo[i] : self.hi[i]].tobytes().decode("ascii")

    def __iter__(self):
        raw = self.data.tobytes()
        return (raw[a:b].decode("ascii") for a, b in zip(self.lo.tolist(), self.hi.tolist()))

    def tolist(self):
        return list(self)

    def __eq__(self, other):
        if isinstance(other, NameTable):
            other = other.tolist()
        return self.tolist() == other


def parse_records(data):
    """Native FASTA/FASTQ scan of an inflated buffer (sgc_reads_count + sgc_reads_scan).

    -> (names NameTable, seq uint8 concatenation, seq_off int64[n+1], rec_pos int64[n])
    ``rec_pos[i]`` is the uncompressed position of record i's header line (the position the
    reference's iterators report: search_utils.py:88/121-124 for FASTA, :136 for FASTQ).  Names and
    sequence lines are stripped the way those iterators strip them."""
    L = lib()
    data = np.ascontiguousarray(data, np.uint8)
    if data.size == 0:
        z = np.zeros(0, np.int64)
        return NameTable(data, z, z), np.zeros(0, np.uint8), np.zeros(1, np.int64), z
    info = np.zeros(3, np.int64)
    nt = n_threads()
    try:
        check(L.sgc_reads_count(_p(data), data.size, nt, _p(info)))
    except Exception as e:  # the reference raises a bare Exception for an unknown first byte
        if "unknown start chr" in str(e):
            raise Exception("unknown start chr {}".format(chr(data[0])))
        raise ValueError(str(e))
    nrec, nbytes = int(info[0]), int(info[1])
    rec_pos = np.empty(nrec, np.int64)
    lo = np.empty(nrec, np.int64)
    hi = np.empty(nrec, np.int64)
    seq_off = np.empty(nrec + 1, np.int64)
    seq = np.empty(nbytes, np.uint8)
    check(L.sgc_reads_scan(_p(data), data.size, nt, nrec, nbytes, _p(rec_pos), _p(lo), _p(hi), _p(seq_off), _p(seq)))
    return NameTable(data, lo, hi), seq, seq_off, rec_pos
