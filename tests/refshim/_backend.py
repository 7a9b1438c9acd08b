"""Backend switch of the test shims ``khmer`` / ``bbhash_table``: SGC_REFSHIM_BACKEND = ``b200`` (the raw ctypes
binding of libsgc_b200.so, tests/refshim/sgc_b200_binding.py) or ``oracle`` (the CPU oracle; lets the reference's
own Python run in a container without a GPU)."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

BACKEND = os.environ.get("SGC_REFSHIM_BACKEND", "oracle")

if BACKEND == "b200":
    from sgc_b200_binding import BBHashTable, Nodetable  # noqa: F401
else:
    from oracle import oracle as _O

    class Nodetable:
        def __init__(self, ksize, *_):
            self.k = int(ksize)

        def get_kmer_hashes(self, seq):
            if len(seq) < self.k:
                raise ValueError(f"sequence length ({len(seq)}) must >= the hashtable k-mer size ({self.k})")
            return _O.hash_sequence(seq, self.k).tolist()

    class BBHashTable(_O.OracleTable):
        def save(self, mphf_filename, array_filename):
            with open(array_filename, "wb") as fp:
                np.savez_compressed(fp, mphf_to_hash=self.mphf_to_hash, mphf_to_value=self.mphf_to_value)
            open(mphf_filename, "wb").write(b"")

        @classmethod
        def load(cls, mphf_filename, array_filename):
            with np.load(array_filename) as z:
                return cls.from_arrays(z["mphf_to_hash"], z["mphf_to_value"])
