/*
 * sgc_b200.h -- C-ABI of libsgc_b200.so: the B200 (sm_100a) implementation of
 * spacegraphcats' k-mer -> cDBG-unitig indexing hot path.
 *
 * This is the drop-in boundary.  The reference has no native code of its own on
 * this path; it binds two third-party natives (khmer, bbhash_table) from Python at
 * spacegraphcats/cdbg/hashing.py:7-8.  Each entry point below names the reference
 * interface it replaces.  Paths are relative to the reference checkout.
 *
 * Conventions
 *   - plain C types only; every pointer named d_* is DEVICE memory on the context's
 *     GPU, every h_* is HOST memory; sizes are explicit; buffers are caller-owned.
 *   - every function returns 0 on success or a negative sgc_status; it never throws
 *     and there is no CPU fallback: without a usable CUDA device sgc_ctx_create
 *     fails with SGC_ERR_CUDA.  sgc_last_error() gives the message (thread-local).
 *   - work is enqueued on the context's CUDA stream.  Functions with host out
 *     parameters (h_*) synchronise that stream before returning; the others are
 *     asynchronous.
 *   - sequences are passed as one concatenated ASCII buffer d_seq[seq_bytes] plus
 *     int64 offsets d_off[nseq+1] (sequence s = d_seq[d_off[s] .. d_off[s+1])).
 *     A sequence shorter than k contributes no k-mers (callers of the reference
 *     guard this case: cdbg/index_reads.py:143, search/query_by_sequence.py:174).
 *   - unitig ids are uint32; SGC_MISS (0xFFFFFFFF, the reference's UNSET_VALUE,
 *     cdbg/index_cdbg_by_kmer.py:24) means "hash not in the table" (Python None).
 */
#ifndef SGC_B200_H
#define SGC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGC_MISS 0xFFFFFFFFu
#define SGC_MAX_ID 0xFFFFFFFDu /* 0xFFFFFFFE marks a dropped (multicount) hash internally */
#define SGC_MAX_K 255          /* khmer WordLength is a byte */

typedef enum {
    SGC_OK = 0,
    SGC_ERR_CUDA = -1,      /* CUDA runtime error (message in sgc_last_error) */
    SGC_ERR_ARG = -2,       /* bad argument */
    SGC_ERR_CAPACITY = -3,  /* an output buffer or the table is too small */
    SGC_ERR_NOMEM = -4
} sgc_status;

typedef struct sgc_ctx sgc_ctx;     /* one GPU + one stream + scratch memory */
typedef struct sgc_table sgc_table; /* GPU-resident exact map u64 hash -> u32 unitig id */

const char* sgc_last_error(void);
int sgc_abi_version(void);

/* device = CUDA ordinal; stream = an existing cudaStream_t to adopt, or NULL to create one. */
int sgc_ctx_create(int device, void* stream, sgc_ctx** out);
void sgc_ctx_destroy(sgc_ctx* ctx);
int sgc_ctx_sync(sgc_ctx* ctx);
void* sgc_ctx_stream(sgc_ctx* ctx);
/* number of kernels of this library launched through ctx so far (bench.py gpu_launches) */
int64_t sgc_ctx_launches(const sgc_ctx* ctx);

/*
 * K1 -- replaces khmer.Nodetable(k,1,1).get_kmer_hashes(seq) as called by
 * hash_sequence(), cdbg/hashing.py:14-20, batched over nseq sequences.
 * Writes d_kmer_off[nseq+1] (exclusive scan of max(0, L-k+1)) and the hashes of
 * sequence s at d_hash_out[d_kmer_off[s] ..], in k-mer position order.
 * d_status (nullable, nseq int32): bit 0 set when the sequence contains a byte other
 * than A/C/G/T (khmer raises on those; the caller decides).  hash_cap = capacity of
 * d_hash_out in k-mers (SGC_ERR_CAPACITY when short; nothing is written then).
 * h_total_kmers (nullable): total k-mers; passing it synchronises.
 */
int sgc_hash_batch(sgc_ctx* ctx, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                   int64_t nseq, int k, uint64_t* d_hash_out, int64_t hash_cap,
                   int64_t* d_kmer_off, int32_t* d_status, int64_t* h_total_kmers);

/*
 * sourmash-style DNA k-mer hash, next rows f3/f4 of the scope table: for every k-mer
 * MurmurHash3_x64_128(min(kmer, revcomp(kmer)), k, seed).lo64 (lexicographic minimum of the ASCII
 * strings; sourmash uses seed 42).  Replaces MinHash.add_sequence as used for the per-unitig
 * n=1 MinHash of cdbg/sort_bcalm_unitigs.py:56-128 (d_seq_min[s] = minimum hash of sequence s,
 * UINT64_MAX when it has no k-mer) and for scaled signatures (search/query_by_sequence.py:52-89,176:
 * d_hash_out holds every hash, the caller keeps those below 2^64/scaled).  Either output may be NULL.
 */
int sgc_minhash_batch(sgc_ctx* ctx, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                      int64_t nseq, int k, uint64_t seed, uint64_t* d_hash_out, int64_t hash_cap,
                      int64_t* d_kmer_off, uint64_t* d_seq_min, int32_t* d_status,
                      int64_t* h_total_kmers);

/*
 * K2 -- replaces BBHashTable() + .initialize(list(all_kmers)) + table[kmer] = cdbg_id,
 * i.e. build_mphf(), cdbg/index_cdbg_by_kmer.py:27-115.
 * sgc_table_create sizes an empty table for up to max_entries distinct hashes; load = expected
 * hashes per 32-byte bucket (0 = default 0.6; memory = 32 B * max_entries / load).
 * Insertion semantics are build_mphf's: a hash inserted with two DIFFERENT ids is
 * dropped from the map for good (multicount, :40-57); the same (hash, id) again is a no-op.
 */
int sgc_table_create(sgc_ctx* ctx, int64_t max_entries, float load, sgc_table** out);
void sgc_table_free(sgc_table* t);
/* insert n (hash, id) pairs; also how a reference-written contigs.indices
 * (mphf_to_hash, mphf_to_value; BBHashTable.load, cdbg/hashing.py:126) is loaded. */
int sgc_table_insert_pairs(sgc_table* t, const uint64_t* d_hash, const uint32_t* d_id, int64_t n);
/* fused K1+K2: hash every k-mer of every sequence and insert it with the sequence's id
 * (d_ids[s], or id_base + s when d_ids is NULL).  Both passes of build_mphf in one. */
int sgc_table_insert_sequences(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes,
                               const int64_t* d_off, int64_t nseq, int k, const uint32_t* d_ids,
                               uint32_t id_base, int32_t* d_status);
/* live = hashes in the map (len(table), cdbg/hashing.py:33-34); multi = dropped hashes;
 * max_id = max stored id (index_cdbg_by_kmer.py:106).  Synchronises. */
int sgc_table_stats(sgc_table* t, int64_t* h_live, int64_t* h_multi, uint32_t* h_max_id);
/* all live (hash, id) pairs, sorted by hash -- what BBHashTable.save writes to
 * contigs.indices (cdbg/hashing.py:115), up to order.  cap in pairs.  Synchronises. */
int sgc_table_export(sgc_table* t, uint64_t* d_hash_out, uint32_t* d_id_out, int64_t cap,
                     int64_t* h_n_out);
int64_t sgc_table_bytes(const sgc_table* t);
/* Optional unitig SEQUENCE STORE: every sequence buffer passed to sgc_table_insert_sequences, 2 bits per base
 * (0.25 B per base) plus one "break" bit per base position; every slot then records where its k-mer lies in
 * the store (30 bits) and on which strand.  Call before sgc_table_insert_sequences.  With it sgc_label_reads
 * hashes and probes only the first k-mer of every group of sixteen; the other fifteen are settled by comparing
 * the read's bases with the store, and inherit the anchor's id when no break bit (= "a table lookup may answer
 * differently for this position": sequence starts, positions without a k-mer, repeated hashes, dropped
 * multicounts) lies in between; everything else is hashed and probed individually: identical labels, a quarter
 * of the hashes and table lines.  SGC_ERR_CAPACITY from the insert when the store would exceed 2^30 - 65 bases
 * (positions are 30 bits of a slot).  Replaces what the reference does per k-mer at cdbg/index_reads.py:143-152. */
int sgc_table_enable_side(sgc_table* t);
int64_t sgc_table_side_bytes(const sgc_table* t);
/* hash-range partitioned tables: the top `shift` bits of every hash stored here are the owner
 * rank and therefore constant; skip them when choosing the home bucket.  Call before inserting. */
int sgc_table_set_home_shift(sgc_table* t, int shift);

/*
 * K3 -- replaces BBHashTable.__getitem__ (kmer_idx.get_cdbg_id, cdbg/hashing.py:36-37)
 * over an array: d_id_out[i] = id or SGC_MISS.
 */
int sgc_lookup(sgc_table* t, const uint64_t* d_hash, int64_t n, uint32_t* d_id_out);

/*
 * K5 -- replaces BBHashTable.get_unique_values(hashes) = MPHF_KmerIndex.count_cdbg_matches,
 * cdbg/hashing.py:67-69: d_count_by_id[id] += 1 for every input hash present (duplicates
 * counted, exactly like iterating a list).  distinct != 0 first de-duplicates the hashes
 * (the set() that Query builds at search/query_by_sequence.py:170-176).
 * d_count_by_id must hold n_ids uint32 and is zeroed here.  h_n_miss nullable.
 */
int sgc_count_matches(sgc_table* t, const uint64_t* d_hash, int64_t n, int distinct,
                      uint32_t* d_count_by_id, int64_t n_ids, int64_t* h_n_miss,
                      int64_t* h_n_distinct);

/*
 * Batched queries -- replaces, for MANY queries in one call, Query.__init__'s `self.kmers.update(hash_sequence(...))`
 * over all records of a query (a SET: search/query_by_sequence.py:170-176) followed by
 * kmer_idx.count_cdbg_matches(self.kmers) (:189).  Sequence s belongs to query d_seq_query[s] (int32, grouped:
 * all sequences of a query are consecutive).  Results: d_q_distinct[n_queries] = len(query.kmers);
 * (d_keys_out, d_counts_out)[0 .. *h_n_pairs) = ((query << 32) | unitig id, number of the query's DISTINCT
 * k-mers on that unitig), ascending by key -- query q's dict is its run of keys.  cap = capacity of the two
 * output arrays (SGC_ERR_CAPACITY with *h_n_pairs = needed).  *h_n_hits = distinct matching k-mers over all queries.
 * Synchronises.  Two forms, same results: a table that carries the unitig sequence store runs the label kernel in
 * query mode -- per settled run the interval of store positions, per missing k-mer its hash; distinct matches = size of
 * the union of a query's intervals per unitig (no k-mer hash is written to memory) --, any other table (or
 * SGC_QUERY_SORT=1) hashes every k-mer, sorts (hash, query) and looks the first occurrences up.
 */
int sgc_query_count_batch(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                          int64_t nseq, const int32_t* d_seq_query, int n_queries, int k,
                          uint64_t* d_keys_out, uint32_t* d_counts_out, int64_t cap, int64_t* d_q_distinct,
                          int32_t* d_status, int64_t* h_n_pairs, int64_t* h_n_kmers, int64_t* h_n_hits);

/*
 * Fused K1+K3+K5 -- per-k-mer get_cdbg_id with no de-duplication, the loop of
 * search/count_dominator_abundance.py:80-85.  d_kmer_ids (nullable) receives the
 * unitig id or SGC_MISS of every k-mer (laid out by d_kmer_off like sgc_hash_batch);
 * d_count_by_id (nullable, n_ids, zeroed here) the per-unitig totals.
 */
int sgc_lookup_sequences(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes,
                         const int64_t* d_off, int64_t nseq, int k, uint32_t* d_kmer_ids,
                         int64_t ids_cap, int64_t* d_kmer_off, uint32_t* d_count_by_id,
                         int64_t n_ids, int32_t* d_status, int64_t* h_n_kmers, int64_t* h_n_hits);

/*
 * Fused K1+K3+K4 -- the per-read step of index_reads, cdbg/index_reads.py:143-152:
 * for every sequence the DISTINCT unitig ids its k-mers hit, ascending, as CSR:
 * d_ids_off[nseq+1], d_ids_out[ids_cap].  Returns SGC_ERR_CAPACITY (with
 * *h_n_labels = needed size) when ids_cap is short.  Synchronises.
 */
int sgc_label_reads(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                    int64_t nseq, int k, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
                    int32_t* d_status, int64_t* h_n_kmers, int64_t* h_n_hits, int64_t* h_n_labels);
/* asynchronous two-step form of the same (for pipelined callers): _launch enqueues the
 * work, _finish synchronises and reports; the outputs are valid after _finish.  No other call on
 * the same context may be made between the two (they share the context's scratch memory). */
int sgc_label_reads_launch(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes,
                           const int64_t* d_off, int64_t nseq, int k, int64_t* d_ids_off,
                           uint32_t* d_ids_out, int64_t ids_cap, int32_t* d_status);
int sgc_label_reads_finish(sgc_table* t, int64_t* h_n_kmers, int64_t* h_n_hits,
                           int64_t* h_n_labels);
/*
 * The same step on 2-BIT PACKED reads (a quarter of the bytes over PCIe and HBM): A=0 C=1 G=2 T=3, base b
 * of record s in bits 2(b%4).. of byte 4*U(s) + b/4 where U(s) = sum over earlier records of
 * ceil(len/16) -- every record starts on a 4-byte boundary (40 bytes per 150-bp read); d_packed itself must be
 * 16-byte aligned.  Lengths in bases: d_len[nseq] (int32), or
 * NULL when every record has uniform_len bases (nothing but the packed bytes crosses the bus then).
 * The kernel expands the codes back to the ASCII words the reference hash is defined on, so the labels are
 * those of sgc_label_reads on the unpacked reads.  Packed records cannot carry non-ACGT bytes: the packer
 * (sgc_pack_reads on the host, sgc_pack_reads_device) reports such records and the caller applies the
 * reference policy (khmer raises; cdbg/index_reads.py:147-150).  _launch/_finish pair like above.
 */
int64_t sgc_packed_bytes(const int32_t* h_len, int uniform_len, int64_t nseq); /* -1: bad argument */
int sgc_label_reads_packed_launch(sgc_table* t, const uint8_t* d_packed, int64_t packed_bytes,
                                  const int32_t* d_len, int uniform_len, int64_t nseq, int k,
                                  int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap);
int sgc_label_reads_packed(sgc_table* t, const uint8_t* d_packed, int64_t packed_bytes,
                           const int32_t* d_len, int uniform_len, int64_t nseq, int k,
                           int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
                           int64_t* h_n_kmers, int64_t* h_n_hits, int64_t* h_n_labels);
/* The CSR offsets on the wire: d_counts[s] = d_off[s+1] - d_off[s] as one byte, d_counts[n] = 1 when some count does
 * not fit (the caller then copies the offsets themselves); asynchronous.  sgc_counts8_to_offsets (host, threads)
 * rebuilds h_off[n+1] from the bytes.  16 MB instead of 134 MB back over PCIe for 2^24 reads. */
int sgc_offsets_to_counts8(sgc_ctx* ctx, const int64_t* d_off, int64_t n, uint8_t* d_counts);
int sgc_counts8_to_offsets(const uint8_t* h_counts, int64_t n, int n_threads, int64_t* h_off);
/* ASCII (d_seq, d_off) already on the device -> packed records + d_len (nullable) + d_status (nullable,
 * bit 0 = non-ACGT byte, packed as A).  *h_packed_bytes = bytes written.  Synchronises. */
int sgc_pack_reads_device(sgc_ctx* ctx, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                          int64_t nseq, uint8_t* d_packed, int64_t packed_cap, int32_t* d_len,
                          int32_t* d_status, int64_t* h_packed_bytes);

/*
 * K6 -- replaces MPHF_KmerIndex.count_catlas_matches / build_catlas_node_sizes,
 * cdbg/hashing.py:71-107 and :42-65: sums over the catlas.  The catlas is given as CSR
 * in child-before-parent level order (search/catlas.py:142-154):
 *   level-1 node v : d_node_count[v] = sum of d_count_by_id[c] for c in d_member[d_member_off[v] ..)
 *                    (members are unitig ids: catlas.layer1_to_cdbg)
 *   level>1 node v : sum of d_node_count[c] over its children (members are node ids)
 * Nodes are processed level by level: d_level_nodes[d_level_off[l] .. d_level_off[l+1]) are the
 * node ids of level l+1 (h_level_off is a HOST array of n_levels+1 offsets).
 */
int sgc_count_doms(sgc_ctx* ctx, const uint32_t* d_count_by_id, const int64_t* d_member_off,
                   const uint32_t* d_member, const uint32_t* d_level_nodes,
                   const int64_t* h_level_off, int n_levels, uint32_t* d_node_count,
                   int64_t n_nodes);

/*
 * K7 -- hash-range partitioned table across R GPUs (owner(hash) = (hash * R) >> 64).
 * sgc_hash_partition hashes every k-mer and writes it into the bucket of its owner rank:
 * d_bucketed[d_bucket_off[r] ..) holds rank r's hashes; d_perm[j] = original k-mer index of
 * bucketed element j (to un-permute the replies).  d_bucket_off has R+1 int64 (device);
 * h_bucket_off (nullable, host, R+1) receives a copy and synchronises.  The exchange itself
 * is an NCCL all-to-all issued by the host layer on the same stream.
 */
int sgc_hash_partition(sgc_ctx* ctx, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                       int64_t nseq, int k, int n_ranks, uint64_t* d_bucketed, int64_t* d_perm,
                       int64_t cap, int64_t* d_bucket_off, int64_t* d_kmer_off, int32_t* d_status,
                       int64_t* h_bucket_off);
/* bucket precomputed (hash[, id]) pairs by owner rank (partitioned table build) */
int sgc_partition_pairs(sgc_ctx* ctx, const uint64_t* d_hash, const uint32_t* d_id, int64_t n,
                        int n_ranks, uint64_t* d_hash_out, uint32_t* d_id_out, int64_t* d_perm,
                        int64_t* d_bucket_off, int64_t* h_bucket_off);
/* K4 on gathered per-k-mer ids: d_kmer_ids[n_kmers] laid out by d_kmer_off[nseq+1] ->
 * CSR of distinct ids per sequence (same outputs as sgc_label_reads).  Synchronises. */
int sgc_labels_from_ids(sgc_ctx* ctx, const uint32_t* d_kmer_ids, const int64_t* d_kmer_off,
                        int64_t nseq, int64_t n_kmers, int64_t* d_ids_off, uint32_t* d_ids_out,
                        int64_t ids_cap, int64_t* h_n_hits, int64_t* h_n_labels);
/* Fused K7 (no sort, no permutation array).  Request side: hash every k-mer and append it to the
 * send region of its owner rank: d_send = n_ranks regions of send_cap hashes; d_ridx[k-mer] =
 * owner << (32 - log2 R) | position in that region; h_counts[n_ranks] (host) = hashes per owner.
 * Synchronises.  Reply side: d_replies has the layout of d_send (one uint32 id per request, as
 * returned by the owners); outputs like sgc_label_reads (+ optional per-k-mer ids). */
int sgc_partition_hashes(sgc_ctx* ctx, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                         int64_t nseq, int k, int n_ranks, uint64_t* d_send, int64_t send_cap,
                         uint32_t* d_ridx, int64_t ridx_cap, int32_t* d_status, int64_t* h_counts);
int sgc_labels_from_replies(sgc_ctx* ctx, const int64_t* d_off, int64_t nseq, int64_t seq_bytes, int k,
                            int n_ranks, const uint32_t* d_ridx, const uint32_t* d_replies,
                            int64_t send_cap, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
                            uint32_t* d_kmer_ids, int64_t* h_n_kmers, int64_t* h_n_hits,
                            int64_t* h_n_labels);
/*
 * Fused compute + exchange over NVLink peer memory (one process per GPU, buffers shared with CUDA
 * IPC): sgc_partition_hashes_p2p stores every hash straight into its OWNER GPU's inbox
 * (h_peer_inboxes[r] = base of GPU r's inbox, mapped in this process; this rank writes region
 * my_rank of region_cap hashes); sgc_lookup_regions_p2p probes all source regions of the local
 * inbox and stores every reply straight into the SOURCE GPU's reply buffer (region my_rank, same
 * position), which is exactly the d_replies layout sgc_labels_from_replies consumes.  Between the
 * two calls the ranks only exchange the per-region counts and a barrier.  sgc_lookup_regions_p2p
 * is asynchronous (it only enqueues), so the next sub-batch's sgc_partition_hashes_p2p, issued
 * through a second context, overlaps it: integer-bound hashing under HBM-bound probing.
 */
int sgc_partition_hashes_p2p(sgc_ctx* ctx, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                             int64_t nseq, int k, int n_ranks, int my_rank,
                             uint64_t* const* h_peer_inboxes, int64_t region_cap, uint32_t* d_ridx,
                             int64_t ridx_cap, int32_t* d_status, int64_t* h_counts);
int sgc_lookup_regions_p2p(sgc_table* t, const uint64_t* d_inbox, const int64_t* h_counts_from,
                           int64_t region_cap, int n_ranks, int my_rank, uint32_t* const* h_peer_replies,
                           int ctas_per_sm /* 0 = fill the GPU; 1..7 = leave room for a concurrent kernel */);
/*
 * The whole labelling step against the partitioned table as ONE call with no host round trip between its
 * kernels: the ranks synchronise through flags in peer memory (NVLink) instead of count exchanges and barriers.
 *   request kernel : hash + store every hash into its owner's inbox; its last warp writes this rank's
 *                    per-owner counts and the step's epoch into the owners' mailboxes
 *   serve kernel   : per source region, waits for that rank's stamp, probes, stores every id straight into the
 *                    requester's reply buffer; its last CTA stamps the requesters' reply buffers
 *   gather kernel  : waits for every owner's stamp, replies -> per-read labels (outputs like sgc_label_reads)
 * h_peer_*[r] = rank r's buffer mapped into this process (CUDA IPC; [my_rank] = the local one): inbox
 * (n_ranks x region_cap hashes), replies (n_ranks x region_cap ids), mail_count (n_ranks uint64), mail_stamp
 * and reply_stamp (n_ranks uint32, zero-initialised).  epoch = 1, 2, ... identical on every rank.  Waits are
 * bounded (~4 s): a missing peer ends in SGC_ERR_CUDA, never in a hung GPU.  *h_flags (nullable): bit 2 = a
 * region overflowed region_cap (SGC_ERR_CAPACITY), bit 3 = time-out; the caller makes the retry collective.
 */
int sgc_label_reads_partitioned(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                                int64_t nseq, int k, int n_ranks, int my_rank, uint32_t epoch,
                                uint64_t* const* h_peer_inboxes, uint32_t* const* h_peer_replies,
                                uint64_t* const* h_peer_mail_count, uint32_t* const* h_peer_mail_stamp,
                                uint32_t* const* h_peer_reply_stamp, int64_t region_cap, uint32_t* d_ridx,
                                int64_t ridx_cap, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
                                int32_t* d_status, int64_t* h_n_kmers, int64_t* h_n_hits, int64_t* h_n_labels,
                                int32_t* h_flags);
/*
 * Hash-range partitioned table PROBED WHERE IT LIES (the product path of BASELINE.json configs[4]):
 * owner(hash) = its top log2(R) bits; every rank owns one table (same size everywhere, home shift log2 R) and
 * maps the other ranks' bucket arrays with CUDA IPC, so the label kernel loads the 32-byte home bucket of a
 * k-mer straight from its owner's HBM over NVLink -- no request / reply buffers, no collective in the label path.
 * The unitig sequence store is replicated on every GPU, one shard per rank (the unitigs that rank hashed);
 * unitig ids are dealt to the ranks in ascending ranges (h_id_bounds[r] = first id of rank r), so the shard a
 * hit lies in follows from its id.
 *   sgc_hash_positions      per k-mer of a shard of unitigs: hash + (position in the shard's buffer | strand bit)
 *                           (stands for pass 1 of build_mphf, cdbg/index_cdbg_by_kmer.py:40-57, on one rank)
 *   sgc_table_insert_triples  owner side: insert what arrived from every rank; repeated hashes / multicounts
 *                           are returned as break marks (shard << 32 | position), *h_n_marks of them
 *   sgc_store_build         the shard's bases, 2 bits each, into d_useq (word of position 0) and its break bits
 *                           into d_brk (bit 32 = position 0; the caller pre-sets every bit)
 *   sgc_store_apply_marks   set the marks of all owners in the replicated break bits (shards brk_stride_words apart)
 *   sgc_table_create_shared like sgc_table_create, the bucket array allocated with the CUDA virtual memory
 *                           management API so that it can be exported as a POSIX file descriptor and mapped by the
 *                           other ranks with 2 MB pages (memory opened through legacy CUDA IPC handles serves random
 *                           32-byte peer loads at 0.09 G/s, this at 6.8 G/s: scripts/peer_probe.cu)
 *   sgc_table_export_fd     local pointer, mapped size and (h_fd non-NULL) a new fd of the bucket array; the host
 *                           layer passes the fd to the other ranks (SCM_RIGHTS) and closes it
 *   sgc_peer_import_fd      map such an fd on this context's GPU;  sgc_peer_unmap undoes it
 *   sgc_table_set_peers     from now on sgc_label_reads* on this table probe h_bucket_ptrs[owner] and match against
 *                           the replicated store (d_useq = word of position 0 of shard 0); n_ranks = 0 undoes it
 *   sgc_filter_insert       set the bits of n hashes in a MISS FILTER of 2^filt_bits 64-bit words (word = the hash's top
 *                           filt_bits bits -- so an owner's hashes fill one contiguous slice of it and the full filter is
 *                           the all-gather of the owners' slices --, four bit positions from its low 24 bits)
 *   sgc_table_set_filter    give a peer-probed table the replicated filter: a k-mer that must be probed on its own (one
 *                           that overlaps a sequencing error: almost always in nobody's table) and whose owner is another
 *                           GPU is first looked up in the filter, in local HBM; a missing bit proves the miss (no false
 *                           negatives), so the labels are unchanged and the NVLink load is saved.  d_words = NULL: off
 */
int sgc_filter_insert(sgc_ctx* ctx, const uint64_t* d_hash, int64_t n, uint64_t* d_words, int filt_bits);
int sgc_table_set_filter(sgc_table* t, const uint64_t* d_words, int filt_bits, int my_rank);
#define SGC_STORE_PAD_WORDS 8 /* 32-bit words of zero padding on either side of a store shard */
#define SGC_BRK_PAD_BITS 32   /* bit index of position 0 in a shard's break-bit array */
int sgc_hash_positions(sgc_ctx* ctx, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                       uint64_t* d_hash_out, uint32_t* d_aux_out, int64_t hash_cap, int64_t* d_kmer_off,
                       int32_t* d_status, int64_t* h_total_kmers);
int sgc_table_insert_triples(sgc_table* t, const uint64_t* d_hash, const uint32_t* d_id, const uint32_t* d_aux, int64_t n,
                             const uint32_t* h_id_bounds, int n_shards, uint64_t* d_marks, int64_t marks_cap,
                             int64_t* h_n_marks);
int sgc_store_build(sgc_ctx* ctx, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                    uint32_t* d_useq, uint32_t* d_brk);
int sgc_store_apply_marks(sgc_ctx* ctx, const uint64_t* d_marks, int64_t n, uint32_t* d_brk, int64_t brk_stride_words);
int sgc_table_create_shared(sgc_ctx* ctx, int64_t max_entries, float load, sgc_table** out);
int sgc_table_export_fd(sgc_table* t, void** d_ptr, int64_t* h_bytes, int* h_fd);
int sgc_peer_import_fd(sgc_ctx* ctx, int fd, int64_t bytes, void** d_ptr);
int sgc_peer_unmap(sgc_ctx* ctx, void* d_ptr, int64_t bytes);
int sgc_table_set_peers(sgc_table* t, int n_ranks, const void* const* h_bucket_ptrs, const uint32_t* h_id_bounds,
                        const uint32_t* d_useq, int64_t useq_stride_words, const uint32_t* d_brk,
                        int64_t brk_stride_words);
int sgc_peer_alloc(sgc_ctx* ctx, int64_t bytes, void** d_ptr, uint8_t* h_handle64);
int sgc_peer_open(sgc_ctx* ctx, const uint8_t* h_handle64, void** d_ptr);
int sgc_peer_close(sgc_ctx* ctx, void* d_ptr);
int sgc_peer_free(sgc_ctx* ctx, void* d_ptr);
/* d_out[d_perm[j]] = d_in[j] */
int sgc_unpermute_u32(sgc_ctx* ctx, const uint32_t* d_in, const int64_t* d_perm, int64_t n,
                      uint32_t* d_out);

/*
 * Bench / test utilities (no reference counterpart): CUDA-event timing of the tile kernel on the
 * context's stream, and on-device generators for the synthetic workloads of SURVEY.md 8(d).
 */
int sgc_ctx_set_timing(sgc_ctx* ctx, int on);
/* synchronises; total milliseconds and number of tile-kernel launches since the last call */
int sgc_ctx_tile_ms(sgc_ctx* ctx, double* h_ms_total, int64_t* h_n);
int sgc_synth_genome(sgc_ctx* ctx, uint8_t* d_out, int64_t n, uint64_t seed);
int sgc_synth_reads(sgc_ctx* ctx, const uint8_t* d_genome, int64_t glen, int64_t nreads, int read_len,
                    uint64_t seed, uint32_t err_ppm, int rc_half, uint8_t* d_out, int64_t* d_off);
int sgc_synth_unitigs(sgc_ctx* ctx, const uint8_t* d_genome, const int64_t* d_cuts, int64_t n, int k,
                      int64_t out_bytes, uint8_t* d_out, int64_t* d_out_off);

/* ---------------------------------------------------------------------------------------------
 * BBHash (boomphf) minimal perfect hash -- the on-disk form of the reference's index.
 *
 * Replaces bbhash.PyMPHF(keys, n, num_threads=4, gamma=1.0) + mphf.save()/load_mphf() as used by
 * bbhash_table.BBHashTable (reference cdbg/hashing.py:8,30-37; index_cdbg_by_kmer.py:63-64,105):
 * the dump produced by sgc_mphf_dump is byte-compatible with boomphf's `contigs.mphf`, and
 * sgc_mphf_order produces `mphf_to_hash` / `mphf_to_value` of `contigs.indices`, so an index
 * built here loads in an unmodified CPU install.  Lookups of this library never need it.
 * --------------------------------------------------------------------------------------------- */
typedef struct sgc_mphf sgc_mphf;
#define SGC_MPHF_NOT_FOUND 0xFFFFFFFFFFFFFFFFull

/* d_keys: n DISTINCT 64-bit keys on the device (duplicates never settle and end in the final map).
 * gamma >= 1.0 (reference: 1.0), 2 <= nb_levels <= 64 (boomphf: 25).  Synchronises. */
int sgc_mphf_build(sgc_ctx* ctx, const uint64_t* d_keys, int64_t n, double gamma, int nb_levels, sgc_mphf** out);
/* parse a boomphf dump held in host memory (the bytes of a reference `contigs.mphf`) */
int sgc_mphf_load(sgc_ctx* ctx, const void* h_buf, int64_t bytes, sgc_mphf** out);
int sgc_mphf_free(sgc_mphf* m);
/* h_info[0..3] = nelem, nb_levels, lastbitsetrank, keys held by the final map; *h_gamma nullable */
int sgc_mphf_info(sgc_mphf* m, int64_t* h_info, double* h_gamma);
/* d_index[i] = MPHF index of d_keys[i]; keys outside the build set give an arbitrary index < nelem or
 * SGC_MPHF_NOT_FOUND (an MPHF has no membership test; callers compare mphf_to_hash[index]).  Asynchronous. */
int sgc_mphf_lookup(sgc_mphf* m, const uint64_t* d_keys, int64_t n, uint64_t* d_index);
/* d_hash_out[index(key)] = key, d_val_out[index(key)] = value for the n == nelem build keys.
 * SGC_ERR_ARG if a key does not map below nelem.  Synchronises. */
int sgc_mphf_order(sgc_mphf* m, const uint64_t* d_keys, const uint32_t* d_vals, int64_t n, uint64_t* d_hash_out,
                   uint32_t* d_val_out);
int sgc_mphf_dump_bytes(sgc_mphf* m, int64_t* h_bytes);
int sgc_mphf_dump(sgc_mphf* m, void* h_buf, int64_t bytes);

/* ---------------------------------------------------------------------------------------------
 * Host feeder for read labelling (no device work; all pointers are host pointers).
 *
 * Stands in for the reference's pure-Python BgzfReader (utils/bgzf/bgzf.py:458-741), its record
 * iterators (search/search_utils.py:67-154: records and the position of each record's header line)
 * and the mate-pairing test of cdbg/index_reads.py:124-141.  n_threads <= 0 = all hardware threads.
 * --------------------------------------------------------------------------------------------- */
/* Walk the BGZF block headers: compressed start and inflated size (ISIZE) of every block.
 * cap = 0 with NULL arrays only counts.  SGC_ERR_ARG if the buffer is not a sequence of BGZF blocks. */
int sgc_bgzf_index(const uint8_t* h_raw, int64_t raw_bytes, int64_t* h_cstart, int64_t* h_usize, int64_t cap,
                   int64_t* h_nblocks, int64_t* h_total_bytes);
/* Inflate all blocks in parallel into h_out (out_bytes = sum of h_usize); CRC32 of every block is checked. */
int sgc_bgzf_inflate(const uint8_t* h_raw, int64_t raw_bytes, const int64_t* h_cstart, const int64_t* h_usize,
                     int64_t nblocks, int n_threads, uint8_t* h_out, int64_t out_bytes);
/* FASTA ('>') or 4-line FASTQ ('@') text: h_info[0..2] = records, total sequence bytes, format (0 FASTA, 1 FASTQ).
 * SGC_ERR_ARG: "unknown start chr", or a FASTQ record that is not '@' ... '+' ... */
int sgc_reads_count(const uint8_t* h_data, int64_t n, int n_threads, int64_t* h_info);
/* Streaming in bounded memory: *h_cut = end of the last complete record of a buffer that may stop inside one
 * (FASTQ: a multiple of 4 lines; FASTA: the start of the last header line).  0 = no complete record. */
int sgc_reads_cut(const uint8_t* h_data, int64_t n, int n_threads, int64_t* h_cut);
/* h_rec_pos[r] = position of record r's header line; [h_name_lo[r], h_name_hi[r]) = its name inside
 * h_data (header line stripped, marker removed); h_seq_off[nrec+1] / h_seq = the packed sequences
 * (every sequence line stripped of surrounding whitespace), i.e. the (seq, off) pair sgc_label_reads takes. */
int sgc_reads_scan(const uint8_t* h_data, int64_t n, int n_threads, int64_t nrec, int64_t seq_bytes,
                   int64_t* h_rec_pos, int64_t* h_name_lo, int64_t* h_name_hi, int64_t* h_seq_off, uint8_t* h_seq);
/* Host packer for sgc_label_reads_packed: (h_seq, h_off) as produced by sgc_reads_scan -> packed records,
 * h_len[nseq], h_bad[nseq] (nullable; 1 = the record holds a byte other than A/C/G/T, packed as A). */
int sgc_pack_reads(const uint8_t* h_seq, const int64_t* h_off, int64_t nseq, int n_threads, uint8_t* h_packed,
                   int64_t packed_cap, int32_t* h_len, uint8_t* h_bad, int64_t* h_packed_bytes);
/* h_last[i] = latest earlier record with h_elig set (-1: none), h_paired[i] = record i is that record's
 * mate ("x/1" then "x/2", or two equal non-empty names).  ignore_paired: all 0 / -1. */
int sgc_reads_pair_flags(const uint8_t* h_data, const int64_t* h_name_lo, const int64_t* h_name_hi,
                         const uint8_t* h_elig, int64_t nrec, int ignore_paired, uint8_t* h_paired, int64_t* h_last);
/* The rows index_reads INSERTs (cdbg/index_reads.py:124-167), in the reference's order, from the flags
 * above, every record's BGZF offset and the CSR of distinct ascending ids sgc_label_reads returned.
 * cap = 0 counts only (*h_n_rows); SGC_ERR_CAPACITY when the arrays are short. */
int sgc_reads_rows(const uint8_t* h_elig, const uint8_t* h_paired, const int64_t* h_last, const int64_t* h_offsets,
                   const int64_t* h_ids_off, const uint32_t* h_ids, int64_t nrec, int64_t* h_rows_offset,
                   int64_t* h_rows_id, int64_t cap, int64_t* h_n_rows);

#ifdef __cplusplus
}
#endif
#endif /* SGC_B200_H */
