from .hashing import hash_sequence, hash_sequences, MPHF_KmerIndex, GpuKmerTable  # noqa: F401
