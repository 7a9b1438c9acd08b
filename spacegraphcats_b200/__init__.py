"""spacegraphcats_b200 -- B200 (sm_100a) implementation of spacegraphcats' k-mer -> cDBG-unitig
indexing hot path, behind the reference's Python surface:

    spacegraphcats_b200.cdbg.hashing            hash_sequence, MPHF_KmerIndex   (cdbg/hashing.py)
    spacegraphcats_b200.cdbg.index_cdbg_by_kmer build_mphf, main                (cdbg/index_cdbg_by_kmer.py)
    spacegraphcats_b200.cdbg.index_reads        main                            (cdbg/index_reads.py)
    spacegraphcats_b200.search.catlas           CAtlas                          (search/catlas.py)
    spacegraphcats_b200.search.query_by_sequence Query                          (search/query_by_sequence.py)

All arithmetic runs in hand-written CUDA kernels reached through the C-ABI of
``libsgc_b200.so`` (``include/sgc_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
