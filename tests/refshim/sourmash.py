"""Test shim for ``sourmash``: the slice of ``MinHash`` / ``SourmashSignature`` the reference's query path calls
(search/query_by_sequence.py:27-49, 153-185; cdbg/sort_bcalm_unitigs.py:56-128), over the CPU oracle's restatement of
sourmash's DNA k-mer hash (Murmur3_x64_128 of min(kmer, revcomp), seed 42 -- pinned by the reference's own
``min_hashval`` column, tests/golden/dory_k21_min_hashval.json).  Scaled sketches (num = 0) and bottom-n sketches."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

_MAX = (1 << 64) - 1


class MinHash:
    def __init__(self, n, ksize, scaled=0, seed=42, **_):
        self.num, self.ksize, self.scaled, self.seed = int(n), int(ksize), int(scaled), int(seed)
        self._max_hash = _MAX if self.scaled <= 1 else int(round(_MAX / self.scaled, 0)) if self.scaled else _MAX
        self._h = set()

    def add_sequence(self, sequence, force=False):
        from oracle import oracle as O

        seq = sequence.upper()
        for i in range(len(seq) - self.ksize + 1):
            kmer = seq[i : i + self.ksize]
            if any(c not in "ACGT" for c in kmer):
                if force:
                    continue
                raise ValueError("invalid DNA character in input k-mer: %s" % kmer)
            self.add_hash(O.minhash_kmer_py(kmer, self.seed))

    def add_hash(self, h):
        if self.num:
            self._h.add(h)
            if len(self._h) > self.num:
                self._h.discard(max(self._h))
        elif h <= self._max_hash:
            self._h.add(h)

    @property
    def hashes(self):
        return {h: 1 for h in sorted(self._h)}

    def __len__(self):
        return len(self._h)

    def copy_and_clear(self):
        return MinHash(self.num, self.ksize, scaled=self.scaled, seed=self.seed)

    def contained_by(self, other):
        return len(self._h & other._h) / len(self._h) if self._h else 0.0

    def similarity(self, other):
        u = len(self._h | other._h)
        return len(self._h & other._h) / u if u else 0.0


class SourmashSignature:
    def __init__(self, minhash, name="", filename=""):
        self.minhash, self.name, self.filename = minhash, name, filename


def save_signatures(sigs, fp):
    import json

    json.dump([{"name": s.name, "filename": s.filename, "ksize": s.minhash.ksize, "mins": sorted(s.minhash._h)} for s in sigs], fp)
