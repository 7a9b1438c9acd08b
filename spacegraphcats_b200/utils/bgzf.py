"""Block-gzip (BGZF) support for read labelling: whole-file block inflate plus the mapping from
uncompressed positions to BGZF *virtual offsets* ``(block_start << 16) | within_block`` -- the
``offset`` column index_reads writes (reference: cdbg/index_reads.py:105, search_utils.py:67-154
on top of utils/bgzf/bgzf.py ``BgzfReader.tell`` :596-611).  A position is always reported inside
the block that holds its byte (a position at a block end belongs to the NEXT block, offset 0),
which is what the reference's reader returns after ``readline``.

BGZF I/O is host-side glue and stays on the CPU (north star); this is a small independent
implementation (zlib per block), not the reference's BioPython module.
"""
import struct
import zlib

import numpy as np

_EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def is_bgzf(path):
    with open(path, "rb") as fp:
        head = fp.read(18)
    return len(head) >= 18 and head[:4] == b"\x1f\x8b\x08\x04" and head[12:14] == b"BC"


def _blocks(raw):
    """Yield (compressed_start, compressed_size, payload_bytes) for each BGZF block."""
    pos, n = 0, len(raw)
    while pos < n:
        if raw[pos : pos + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block at byte %d" % pos)
        xlen = struct.unpack_from("<H", raw, pos + 10)[0]
        extra = raw[pos + 12 : pos + 12 + xlen]
        bsize, i = None, 0
        while i + 4 <= len(extra):
            si1, si2, slen = extra[i], extra[i + 1], struct.unpack_from("<H", extra, i + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", extra, i + 4)[0] + 1
            i += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC subfield at byte %d" % pos)
        cdata = raw[pos + 12 + xlen : pos + bsize - 8]
        payload = zlib.decompress(cdata, -15) if cdata else b""
        yield pos, bsize, payload
        pos += bsize


class BgzfData:
    """A fully inflated BGZF file: ``data`` (uint8) and position -> virtual offset."""

    def __init__(self, path):
        with open(path, "rb") as fp:
            raw = fp.read()
        cstart, ustart, parts, upos = [], [], [], 0
        for cs, _, payload in _blocks(raw):
            if payload:
                cstart.append(cs)
                ustart.append(upos)
                parts.append(payload)
                upos += len(payload)
        self.data = np.frombuffer(b"".join(parts), np.uint8)
        self._cstart = np.array(cstart, np.int64)
        self._ustart = np.array(ustart, np.int64)

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


def write_bgzf(path, payload, block_size=65280):
    """Write ``payload`` (bytes) as BGZF (used by tests and utils.make_bgzf-style conversion)."""
    with open(path, "wb") as fp:
        for lo in range(0, len(payload), block_size):
            chunk = payload[lo : lo + block_size]
            comp = zlib.compressobj(6, zlib.DEFLATED, -15)
            cdata = comp.compress(chunk) + comp.flush()
            bsize = len(cdata) + 25
            fp.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00")
            fp.write(struct.pack("<H", bsize))
            fp.write(cdata)
            fp.write(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        fp.write(_EOF_BLOCK)


def parse_records(data):
    """Vectorised FASTA/FASTQ scan of an inflated buffer.

    -> (names list[str], seq uint8 concatenation, seq_off int64[n+1], rec_pos int64[n])
    ``rec_pos[i]`` is the uncompressed position of record i's header line (the position the
    reference's iterators report: search_utils.py:88/121-124 for FASTA, :136 for FASTQ)."""
    data = np.asarray(data, np.uint8)
    if data.size == 0:
        return [], np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.int64)
    nl = np.flatnonzero(data == 10)
    line_start = np.concatenate([[0], nl + 1])
    line_end = np.concatenate([nl, [data.size]])  # exclusive of '\n'
    if line_start[-1] >= data.size:  # file ends with '\n'
        line_start, line_end = line_start[:-1], line_end[:-1]
    # strip '\r'
    has_cr = (line_end > line_start) & (data[np.maximum(line_end - 1, 0)] == 13)
    line_end = line_end - has_cr
    first = chr(data[0])
    raw = data.tobytes()
    if first == "@":
        nrec = len(line_start) // 4
        hs, he = line_start[0 : 4 * nrec : 4], line_end[0 : 4 * nrec : 4]
        ss, se = line_start[1 : 4 * nrec : 4], line_end[1 : 4 * nrec : 4]
        plus = data[line_start[2 : 4 * nrec : 4]]
        if nrec and (not (data[hs] == ord("@")).all() or not (plus == ord("+")).all()):
            raise ValueError("FASTQ records must be exactly 4 lines")
        lens = se - ss
        seq_off = np.zeros(nrec + 1, np.int64)
        np.cumsum(lens, out=seq_off[1:])
        delta = np.zeros(data.size + 1, np.int32)
        np.add.at(delta, ss, 1)
        np.add.at(delta, se, -1)
        mask = np.cumsum(delta[:-1]) > 0
        seq = data[mask]
        names = [raw[a + 1 : b].decode("ascii").strip() for a, b in zip(hs.tolist(), he.tolist())]
        return names, seq, seq_off, hs.astype(np.int64)
    if first == ">":
        is_hdr = data[line_start] == ord(">")
        hidx = np.flatnonzero(is_hdr)
        hs, he = line_start[hidx], line_end[hidx]
        nrec = len(hidx)
        # bytes of sequence lines, whitespace removed
        delta = np.zeros(data.size + 1, np.int32)
        np.add.at(delta, hs, 1)
        np.add.at(delta, np.minimum(he + 1, data.size), -1)
        in_hdr = np.cumsum(delta[:-1]) > 0
        mask = ~in_hdr & (data != 10) & (data != 13) & (data != 32) & (data != 9)
        seq = data[mask]
        csum = np.concatenate([[0], np.cumsum(mask, dtype=np.int64)])
        bounds = np.concatenate([hs, [data.size]])
        seq_off = csum[bounds] - csum[bounds[0]]
        names = [raw[a + 1 : b].decode("ascii").strip() for a, b in zip(hs.tolist(), he.tolist())]
        return names, seq, seq_off.astype(np.int64), hs.astype(np.int64)
    raise Exception("unknown start chr {}".format(first))
