"""Test shim for ``screed``: open() over FASTA/FASTQ(.gz) yielding records with .name / .sequence / .quality."""
import builtins
import gzip

from .screedRecord import Record


class _Records(list):
    """what screed.open returns: iterable, and a context manager (search/query_by_sequence.py:170)"""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass


def open(filename):  # noqa: A001
    with builtins.open(filename, "rb") as fp:
        raw = fp.read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    return _Records(_parse(raw.decode("ascii")))


def _parse(text):
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith(">"):
            name = ln[1:].strip()
            i += 1
            seq = []
            while i < len(lines) and not lines[i].startswith(">"):
                seq.append(lines[i].strip())
                i += 1
            yield Record(name=name, sequence="".join(seq))
        elif ln.startswith("@"):
            yield Record(name=ln[1:].strip(), sequence=lines[i + 1].strip(), quality=lines[i + 3].strip())
            i += 4
        else:
            i += 1
