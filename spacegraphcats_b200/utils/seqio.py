"""Minimal FASTA / FASTQ (optionally gzip'd) record reader.

The reference reads sequences with ``screed`` (absent here and not on the hot path); the
entry points of this package only need ``.name`` / ``.sequence`` / ``.quality`` records.
``name`` is the whole header line without the leading ``>`` / ``@`` (screed's default).
"""
import gzip
import io


class Record:
    __slots__ = ("name", "sequence", "quality")

    def __init__(self, name, sequence, quality=None):
        self.name, self.sequence, self.quality = name, sequence, quality

    def __repr__(self):
        return "Record(%r, <%d bases>)" % (self.name, len(self.sequence))


def _open_text(path):
    raw = open(path, "rb")
    magic = raw.read(2)
    raw.seek(0)
    if magic == b"\x1f\x8b":
        return io.TextIOWrapper(gzip.GzipFile(fileobj=raw), encoding="ascii", newline=None)
    return io.TextIOWrapper(raw, encoding="ascii", newline=None)


def read_records(path):
    """Yield Records from a FASTA or FASTQ file (plain, gzip or BGZF)."""
    with _open_text(path) as fp:
        line = fp.readline()
        while line:
            line = line.rstrip("\r\n")
            if not line:
                line = fp.readline()
                continue
            if line[0] == ">":
                name, parts = line[1:], []
                line = fp.readline()
                while line and not line.startswith(">"):
                    parts.append(line.strip())
                    line = fp.readline()
                yield Record(name, "".join(parts))
                continue
            if line[0] == "@":
                name = line[1:]
                seq = fp.readline().rstrip("\r\n")
                plus = fp.readline()
                qual = fp.readline().rstrip("\r\n")
                if not plus.startswith("+"):
                    raise ValueError("malformed FASTQ record %r in %s" % (name, path))
                yield Record(name, seq, qual)
                line = fp.readline()
                continue
            raise ValueError("unrecognised sequence file line %r in %s" % (line[:40], path))
