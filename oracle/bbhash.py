"""CPU restatement of the BBHash minimal perfect hash (boomphf) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module; the product
(spacegraphcats_b200/) never does.

What it restates.  The reference stores its k-mer index through bbhash_table.BBHashTable
(reference spacegraphcats/cdbg/hashing.py:8,30-37; index_cdbg_by_kmer.py:63-64,89,105), whose
`contigs.mphf` file is the dump of a boomphf::mphf<uint64_t, SingleHashFunctor<uint64_t>> built by
pybbhash (`bbhash.PyMPHF(keys, n, num_threads=4, gamma=1.0)`).  Neither bbhash_table nor pybbhash
nor BBHash's BooPHF.h is vendored under /root/reference (setup/pip dependencies `bbhash`,
`bbhash-table`, unpinned), so this file restates the published algorithm (Limasset, Rizk, Chikhi,
Peterlongo 2017, "Fast and scalable minimal perfect hashing for massive key sets"; rchikhi/BBHash
BooPHF.h) and pins it on the reference's own fixture:

  * every key of tests/test-data/catlas.dory_k21/contigs.indices (`mphf_to_hash[i]`) must look up
    to index i through the level bitsets read from the fixture's contigs.mphf  (pins the hash
    function, the xorshift level sequence, the fastrange reduction and the rank structure);
  * building from the same 212,475 keys must reproduce the fixture's contigs.mphf byte for byte
    (pins level sizing, collision clearing, rank sampling and the dump layout).

Both pins are exercised in tests/test_oracle.py against the committed golden copy.
Residual ambiguity (stated, not pinned): level sizing truncates `hash_domain * p^level` to an
integer before rounding up to 64 bits; on the fixture truncation and ceiling agree on all 25 levels.

Algorithm.
  hash64(key, seed)      the BBHash integer mixer (below);
  h_0 = hash64(key, 0xAAAAAAAA55555555), h_1 = hash64(key, 0x33333333CCCCCCCC),
  h_l (l >= 2) = xorshift128+ stream seeded with (h_0, h_1);
  level l owns a bitset of size_l bits; a key lands at pos = (h_l * size_l) >> 64 (fastrange64);
  a position hit by exactly one still-unplaced key keeps its bit, positions hit by several are
  cleared and those keys fall through to level l+1; after nb_levels-1 bitset levels the leftovers
  go to an explicit final map;
  index(key) = rank of its bit counted over the concatenated level bitsets (rank samples every
  512 bits, carrying the offset of the earlier levels).
"""
from __future__ import annotations

import math
import struct

import numpy as np

U = np.uint64
SEED0 = U(0xAAAAAAAA55555555)
SEED1 = U(0x33333333CCCCCCCC)
NOT_FOUND = np.iinfo(np.uint64).max


def hash64(key: np.ndarray, seed) -> np.ndarray:
    key = np.asarray(key, dtype=U)
    with np.errstate(over="ignore"):
        h = np.full_like(key, seed)
        h ^= (h << U(7)) ^ (key * (h >> U(3))) ^ (~((h << U(11)) + (key ^ (h >> U(5)))))
        h = (~h) + (h << U(21))
        h ^= h >> U(24)
        h = (h + (h << U(3))) + (h << U(8))
        h ^= h >> U(14)
        h = (h + (h << U(2))) + (h << U(4))
        h ^= h >> U(28)
        h = h + (h << U(31))
    return h


def mulhi64(a: np.ndarray, b: int) -> np.ndarray:
    """(a * b) >> 64 for uint64 arrays."""
    b = U(b)
    m = U(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        a0, a1, b0, b1 = a & m, a >> U(32), b & m, b >> U(32)
        t1 = a1 * b0 + ((a0 * b0) >> U(32))
        t2 = a0 * b1 + (t1 & m)
        return a1 * b1 + (t1 >> U(32)) + (t2 >> U(32))


class LevelHasher:
    """Yields h_0, h_1, then the xorshift stream, one array per level."""

    def __init__(self, keys: np.ndarray):
        self.s0 = hash64(keys, SEED0)
        self.s1 = hash64(keys, SEED1)
        self.level = 0

    def next(self) -> np.ndarray:
        lv = self.level
        self.level += 1
        if lv == 0:
            return self.s0
        if lv == 1:
            return self.s1
        with np.errstate(over="ignore"):
            a, b = self.s0.copy(), self.s1
            self.s0 = b
            a ^= a << U(23)
            self.s1 = a ^ b ^ (a >> U(17)) ^ (b >> U(26))
            return self.s1 + b

    def select(self, mask: np.ndarray) -> None:
        self.s0, self.s1 = self.s0[mask], self.s1[mask]


def level_sizes(nelem: int, gamma: float, nb_levels: int) -> list:
    if nelem == 0:
        return [64] * nb_levels
    gn = gamma * float(nelem)
    proba = 1.0 - ((gn - 1.0) / gn) ** (nelem - 1)
    domain = int(math.ceil(float(nelem) * gamma))
    sizes = []
    for lv in range(nb_levels):
        s = ((int(domain * proba ** lv) + 63) // 64) * 64
        sizes.append(s if s else 64)
    return sizes


def _popcount_words(words: np.ndarray) -> np.ndarray:
    return np.unpackbits(words.view(np.uint8)).reshape(-1, 64).sum(axis=1).astype(np.uint64)


class Mphf:
    def __init__(self, gamma, nb_levels, nelem, sizes, words, ranks, lastbitsetrank, final):
        self.gamma, self.nb_levels, self.nelem = gamma, nb_levels, nelem
        self.sizes, self.words, self.ranks = sizes, words, ranks
        self.lastbitsetrank, self.final = lastbitsetrank, final

    # ---------------------------------------------------------------- build
    @classmethod
    def build(cls, keys, gamma: float = 1.0, nb_levels: int = 25) -> "Mphf":
        keys = np.ascontiguousarray(keys, dtype=U)
        n = len(keys)
        sizes = level_sizes(n, gamma, nb_levels)
        hasher = LevelHasher(keys)
        alive = keys
        words, ranks = [], []
        offset = 0
        for lv in range(nb_levels):
            size = sizes[lv]
            w = np.zeros(1 + size // 64, dtype=U)
            if lv < nb_levels - 1 and len(alive):
                pos = mulhi64(hasher.next(), size)
                upos, cnt = np.unique(pos, return_counts=True)
                single = upos[cnt == 1].astype(np.int64)
                np.bitwise_or.at(w, single >> 6, U(1) << (single & 63).astype(U))
                placed = np.isin(pos, upos[cnt == 1])
                alive = alive[~placed]
                hasher.select(~placed)
            pc = _popcount_words(w)
            cum = np.concatenate([[0], np.cumsum(pc)]).astype(np.uint64)
            ranks.append((cum[:-1][::8] + U(offset)).astype(U))
            offset += int(cum[-1])
            words.append(w)
        # boomphf numbers the leftovers in (thread dependent) arrival order; ascending keys here
        final = {int(k): i for i, k in enumerate(np.sort(alive))}
        return cls(gamma, nb_levels, n, sizes, words, ranks, offset, final)

    # --------------------------------------------------------------- lookup
    def lookup(self, keys) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=U)
        out = np.full(len(keys), NOT_FOUND, dtype=U)
        idx = np.arange(len(keys))
        hasher = LevelHasher(keys)
        for lv in range(self.nb_levels - 1):
            if not len(idx):
                break
            pos = mulhi64(hasher.next(), self.sizes[lv])
            w = self.words[lv]
            wi = (pos >> U(6)).astype(np.int64)
            bit = pos & U(63)
            hit = ((w[wi] >> bit) & U(1)).astype(bool)
            if hit.any():
                hw, hb, hp = wi[hit], bit[hit], pos[hit]
                blk = (hp >> U(9)).astype(np.int64)
                r = self.ranks[lv][blk].copy()
                # whole words between the 512-bit sample and the key's word, then the partial word
                pcw = _popcount_words(w)
                cumw = np.concatenate([[0], np.cumsum(pcw)]).astype(np.uint64)
                r += cumw[hw] - cumw[blk * 8]
                below = w[hw] & ((U(1) << hb) - U(1))
                r += _popcount_words(below)
                out[idx[hit]] = r
            keep = ~hit
            idx = idx[keep]
            hasher.select(keep)
        for j in idx:
            v = self.final.get(int(keys[j]))
            if v is not None:
                out[j] = v + self.lastbitsetrank
        return out

    # ----------------------------------------------------------- dump / load
    def dumps(self) -> bytes:
        parts = [struct.pack("<diQQ", self.gamma, self.nb_levels, self.lastbitsetrank, self.nelem)]
        for lv in range(self.nb_levels):
            w, r = self.words[lv], self.ranks[lv]
            parts.append(struct.pack("<QQ", self.sizes[lv], len(w)))
            parts.append(w.astype("<u8").tobytes())
            parts.append(struct.pack("<Q", len(r)))
            parts.append(r.astype("<u8").tobytes())
        parts.append(struct.pack("<Q", len(self.final)))
        for k, v in self.final.items():
            parts.append(struct.pack("<QQ", k, v))
        return b"".join(parts)

    @classmethod
    def loads(cls, raw: bytes) -> "Mphf":
        gamma, nb_levels, last, nelem = struct.unpack_from("<diQQ", raw, 0)
        off = 28
        sizes, words, ranks = [], [], []
        for _ in range(nb_levels):
            size, nchar = struct.unpack_from("<QQ", raw, off)
            off += 16
            words.append(np.frombuffer(raw, dtype="<u8", count=nchar, offset=off).copy())
            off += 8 * nchar
            (nr,) = struct.unpack_from("<Q", raw, off)
            off += 8
            ranks.append(np.frombuffer(raw, dtype="<u8", count=nr, offset=off).copy())
            off += 8 * nr
            sizes.append(size)
        (nf,) = struct.unpack_from("<Q", raw, off)
        off += 8
        final = {}
        for _ in range(nf):
            k, v = struct.unpack_from("<QQ", raw, off)
            off += 16
            final[k] = v
        if off != len(raw):
            raise ValueError("trailing bytes in mphf dump")
        return cls(gamma, nb_levels, nelem, sizes, words, ranks, last, final)
