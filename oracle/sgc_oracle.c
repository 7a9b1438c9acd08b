/*
 * sgc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic behind spacegraphcats' k-mer -> unitig
 * indexing hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object; the
 * product path (spacegraphcats_b200/) never does and has no CPU fallback.
 *
 * The arithmetic is not in /root/reference: it lives in two third-party native
 * packages the reference imports at spacegraphcats/cdbg/hashing.py:7-8:
 *   - khmer 3.0.0a3 (environment.yml:10): Nodetable(k,1,1).get_kmer_hashes(seq),
 *     called at hashing.py:14-20.  Published algorithm restated here:
 *       h(kmer) = MurmurHash3_x64_128(kmer, k, seed 0).out[0]
 *               ^ MurmurHash3_x64_128(revcomp(kmer), k, seed 0).out[0]
 *     (smhasher MurmurHash3.cpp; khmer kmer_hash.cc _hash_murmur / _revcomp).
 *   - bbhash >= 0.5.4 (setup.py:53): bbhash_table.BBHashTable, used at
 *     hashing.py:34,37,69 and index_cdbg_by_kmer.py:63-64,89,106.  It stores the
 *     original 64-bit hash beside the u32 value and compares on lookup, so its
 *     observable behaviour is an exact map {u64 -> u32}, miss = None; the MPHF
 *     is an implementation detail.  Restated here as open addressing.
 *
 * PARITY PINS (checked in tests/test_oracle.py, CPU-only): all 212,475
 * (hash -> unitig id) pairs of the reference's own fixture
 * tests/test-data/catlas.dory_k21/contigs.indices, the four KATs at
 * tests/test_dory_workflow.py:195-199, the public Murmur3 vector
 * mmh3.hash64("foo") and the dory query/response fixtures.
 * UNPINNED: behaviour on non-ACGT bytes (see sgc_oracle_revcomp below).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ROTL64(x, r) (((x) << (r)) | ((x) >> (64 - (r))))

static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

static inline uint64_t load_le64(const uint8_t *p) {
    uint64_t v;
    memcpy(&v, p, 8); /* little-endian hosts only (x86-64 / aarch64) */
    return v;
}

/* smhasher MurmurHash3_x64_128, both output words. */
void sgc_oracle_murmur3_x64_128(const void *key, int len, uint32_t seed, uint64_t out[2]) {
    const uint8_t *data = (const uint8_t *)key;
    const int nblocks = len / 16;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    for (int i = 0; i < nblocks; i++) {
        uint64_t k1 = load_le64(data + 16 * i), k2 = load_le64(data + 16 * i + 8);
        k1 *= c1; k1 = ROTL64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = ROTL64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = ROTL64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = ROTL64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    switch (len & 15) {
    case 15: k2 ^= ((uint64_t)tail[14]) << 48; /* fallthrough */
    case 14: k2 ^= ((uint64_t)tail[13]) << 40; /* fallthrough */
    case 13: k2 ^= ((uint64_t)tail[12]) << 32; /* fallthrough */
    case 12: k2 ^= ((uint64_t)tail[11]) << 24; /* fallthrough */
    case 11: k2 ^= ((uint64_t)tail[10]) << 16; /* fallthrough */
    case 10: k2 ^= ((uint64_t)tail[9]) << 8;   /* fallthrough */
    case 9:  k2 ^= ((uint64_t)tail[8]) << 0;
             k2 *= c2; k2 = ROTL64(k2, 33); k2 *= c1; h2 ^= k2; /* fallthrough */
    case 8:  k1 ^= ((uint64_t)tail[7]) << 56; /* fallthrough */
    case 7:  k1 ^= ((uint64_t)tail[6]) << 48; /* fallthrough */
    case 6:  k1 ^= ((uint64_t)tail[5]) << 40; /* fallthrough */
    case 5:  k1 ^= ((uint64_t)tail[4]) << 32; /* fallthrough */
    case 4:  k1 ^= ((uint64_t)tail[3]) << 24; /* fallthrough */
    case 3:  k1 ^= ((uint64_t)tail[2]) << 16; /* fallthrough */
    case 2:  k1 ^= ((uint64_t)tail[1]) << 8;  /* fallthrough */
    case 1:  k1 ^= ((uint64_t)tail[0]) << 0;
             k1 *= c1; k1 = ROTL64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/*
 * khmer _revcomp: A<->T, C<->G, string reversed.  UNPINNED for any other byte
 * (no reference test feeds one): this restatement passes such a byte through
 * unchanged (and reports it via sgc_oracle_count_invalid so the host policy
 * can raise instead); the CUDA path implements exactly the same rule.
 */
static inline uint8_t comp_base(uint8_t b) {
    switch (b) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default:  return b;
    }
}

void sgc_oracle_revcomp(const uint8_t *kmer, int k, uint8_t *out) {
    for (int i = 0; i < k; i++) out[k - 1 - i] = comp_base(kmer[i]);
}

int64_t sgc_oracle_count_invalid(const uint8_t *seq, int64_t len) {
    int64_t n = 0;
    for (int64_t i = 0; i < len; i++) {
        uint8_t b = seq[i];
        n += !(b == 'A' || b == 'C' || b == 'G' || b == 'T');
    }
    return n;
}

/* khmer _hash_murmur(kmer, k): forward ^ reverse-complement, seed 0, out[0]. */
uint64_t sgc_oracle_hash_kmer(const uint8_t *kmer, int k) {
    uint8_t rc[256];
    uint64_t f[2], r[2];
    sgc_oracle_murmur3_x64_128(kmer, k, 0, f);
    sgc_oracle_revcomp(kmer, k, rc);
    sgc_oracle_murmur3_x64_128(rc, k, 0, r);
    return f[0] ^ r[0];
}

/*
 * hash_sequence(seq, ksize) -- reference hashing.py:14-20 ->
 * khmer Hashtable::get_kmer_hashes: one hash per k-mer position, in order.
 * Returns the number of k-mers written, or -1 when len < k (khmer raises).
 */
int64_t sgc_oracle_hash_sequence(const uint8_t *seq, int64_t len, int k, uint64_t *out) {
    if (k < 1 || k > 255 || len < k) return -1;
    int64_t n = len - k + 1;
    for (int64_t i = 0; i < n; i++) out[i] = sgc_oracle_hash_kmer(seq + i, k);
    return n;
}

/*
 * Batched hash_sequence over a concatenated buffer: sequence s is
 * seq[off[s] .. off[s+1]); its hashes land at out[out_off[s] ..].  Sequences
 * shorter than k contribute nothing.  threads<=0 -> all OpenMP threads.
 */
int sgc_oracle_hash_batch(const uint8_t *seq, const int64_t *off, int64_t nseq, int k,
                          uint64_t *out, const int64_t *out_off, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    (void)threads;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t s = 0; s < nseq; s++) {
        int64_t len = off[s + 1] - off[s];
        if (len >= k) sgc_oracle_hash_sequence(seq + off[s], len, k, out + out_off[s]);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Exact map u64 -> u32 (BBHashTable's observable behaviour).          */
/* ------------------------------------------------------------------ */
#define SGC_ORACLE_MISS 0xFFFFFFFFu /* = index_cdbg_by_kmer.py:24 UNSET_VALUE */

typedef struct {
    uint64_t *keys;
    uint32_t *vals; /* SGC_ORACLE_MISS = empty slot */
    uint64_t mask;
    int64_t n;
} sgc_oracle_table;

static inline uint64_t slot_of(uint64_t h, uint64_t mask) {
    return (h ^ (h >> 29)) & mask;
}

/* Build from distinct (hash, id) pairs -- e.g. contigs.indices arrays. */
sgc_oracle_table *sgc_oracle_table_new(const uint64_t *hash, const uint32_t *id, int64_t n) {
    sgc_oracle_table *t = (sgc_oracle_table *)calloc(1, sizeof(*t));
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2) cap <<= 1;
    t->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    t->vals = (uint32_t *)malloc(cap * sizeof(uint32_t));
    memset(t->vals, 0xFF, cap * sizeof(uint32_t));
    t->mask = cap - 1;
    t->n = 0;
    for (int64_t i = 0; i < n; i++) {
        uint64_t s = slot_of(hash[i], t->mask);
        while (t->vals[s] != SGC_ORACLE_MISS && t->keys[s] != hash[i]) s = (s + 1) & t->mask;
        if (t->vals[s] == SGC_ORACLE_MISS) t->n++;
        t->keys[s] = hash[i];
        t->vals[s] = id[i];
    }
    return t;
}

void sgc_oracle_table_free(sgc_oracle_table *t) {
    if (!t) return;
    free(t->keys);
    free(t->vals);
    free(t);
}

int64_t sgc_oracle_table_len(const sgc_oracle_table *t) { return t->n; }

static inline uint32_t table_get(const sgc_oracle_table *t, uint64_t h) {
    uint64_t s = slot_of(h, t->mask);
    while (t->vals[s] != SGC_ORACLE_MISS) {
        if (t->keys[s] == h) return t->vals[s];
        s = (s + 1) & t->mask;
    }
    return SGC_ORACLE_MISS;
}

/* BBHashTable.__getitem__ over an array: id or SGC_ORACLE_MISS (None). */
void sgc_oracle_lookup(const sgc_oracle_table *t, const uint64_t *hash, int64_t n, uint32_t *id_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) id_out[i] = table_get(t, hash[i]);
}

/*
 * BBHashTable.get_unique_values (hashing.py:67-69 count_cdbg_matches): dense
 * count_by_id[id] += 1 for every input hash that is present (duplicates in the
 * input are counted).  Returns the number of misses.
 */
int64_t sgc_oracle_count_matches(const sgc_oracle_table *t, const uint64_t *hash, int64_t n,
                                 uint32_t *count_by_id) {
    int64_t miss = 0;
    for (int64_t i = 0; i < n; i++) {
        uint32_t v = table_get(t, hash[i]);
        if (v == SGC_ORACLE_MISS) miss++;
        else count_by_id[v]++;
    }
    return miss;
}

/*
 * index_reads.py:143-152 inner step, batched: for every read with len >= k,
 * the set of distinct unitig ids its k-mers hit, ascending.  CSR output:
 * ids_off[nseq+1] (written here), ids_out (capacity ids_cap).  Two passes when
 * ids_out == NULL (count only).  Returns total ids, or -1 if ids_cap is short.
 * Also returns k-mer and hit totals through n_kmers / n_hits when non-NULL.
 */
static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

int64_t sgc_oracle_label_reads(const sgc_oracle_table *t, const uint8_t *seq, const int64_t *off,
                               int64_t nseq, int k, int64_t *ids_off, uint32_t *ids_out,
                               int64_t ids_cap, int64_t *n_kmers, int64_t *n_hits, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    (void)threads;
    int64_t tot_k = 0, tot_h = 0;
    /* pass 1: per-read distinct counts into ids_off[s+1] */
    ids_off[0] = 0;
#pragma omp parallel
    {
        uint32_t *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 256) reduction(+ : tot_k, tot_h)
        for (int64_t s = 0; s < nseq; s++) {
            int64_t len = off[s + 1] - off[s];
            int64_t nd = 0;
            if (len >= k) {
                int64_t nk = len - k + 1;
                if (nk > cap) { cap = nk * 2; buf = (uint32_t *)realloc(buf, cap * sizeof(uint32_t)); }
                int64_t m = 0;
                for (int64_t i = 0; i < nk; i++) {
                    uint32_t v = table_get(t, sgc_oracle_hash_kmer(seq + off[s] + i, k));
                    if (v != SGC_ORACLE_MISS) buf[m++] = v;
                }
                tot_k += nk;
                tot_h += m;
                qsort(buf, m, sizeof(uint32_t), cmp_u32);
                for (int64_t i = 0; i < m; i++)
                    if (i == 0 || buf[i] != buf[i - 1]) nd++;
            }
            ids_off[s + 1] = nd;
        }
        free(buf);
    }
    for (int64_t s = 0; s < nseq; s++) ids_off[s + 1] += ids_off[s];
    if (n_kmers) *n_kmers = tot_k;
    if (n_hits) *n_hits = tot_h;
    if (!ids_out) return ids_off[nseq];
    if (ids_off[nseq] > ids_cap) return -1;
    /* pass 2: fill */
#pragma omp parallel
    {
        uint32_t *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 256)
        for (int64_t s = 0; s < nseq; s++) {
            int64_t len = off[s + 1] - off[s];
            if (len < k || ids_off[s + 1] == ids_off[s]) continue;
            int64_t nk = len - k + 1;
            if (nk > cap) { cap = nk * 2; buf = (uint32_t *)realloc(buf, cap * sizeof(uint32_t)); }
            int64_t m = 0;
            for (int64_t i = 0; i < nk; i++) {
                uint32_t v = table_get(t, sgc_oracle_hash_kmer(seq + off[s] + i, k));
                if (v != SGC_ORACLE_MISS) buf[m++] = v;
            }
            qsort(buf, m, sizeof(uint32_t), cmp_u32);
            int64_t w = ids_off[s];
            for (int64_t i = 0; i < m; i++)
                if (i == 0 || buf[i] != buf[i - 1]) ids_out[w++] = buf[i];
        }
        free(buf);
    }
    return ids_off[nseq];
}

/*
 * Single-pass variant used as the timed CPU baseline (bench.py cpu_baseline /
 * --impl reference): hash + lookup + per-read distinct, results reduced to a
 * checksum so no output buffer sizing is needed.  Returns k-mers processed.
 */
int64_t sgc_oracle_label_reads_checksum(const sgc_oracle_table *t, const uint8_t *seq,
                                        const int64_t *off, int64_t nseq, int k,
                                        uint64_t *checksum, int64_t *n_hits, int64_t *n_labels,
                                        int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    (void)threads;
    int64_t tot_k = 0, tot_h = 0, tot_l = 0;
    uint64_t sum = 0;
#pragma omp parallel
    {
        uint32_t *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 256) reduction(+ : tot_k, tot_h, tot_l, sum)
        for (int64_t s = 0; s < nseq; s++) {
            int64_t len = off[s + 1] - off[s];
            if (len < k) continue;
            int64_t nk = len - k + 1;
            if (nk > cap) { cap = nk * 2; buf = (uint32_t *)realloc(buf, cap * sizeof(uint32_t)); }
            int64_t m = 0;
            for (int64_t i = 0; i < nk; i++) {
                uint32_t v = table_get(t, sgc_oracle_hash_kmer(seq + off[s] + i, k));
                if (v != SGC_ORACLE_MISS) buf[m++] = v;
            }
            tot_k += nk;
            tot_h += m;
            qsort(buf, m, sizeof(uint32_t), cmp_u32);
            for (int64_t i = 0; i < m; i++)
                if (i == 0 || buf[i] != buf[i - 1]) {
                    tot_l++;
                    sum += ((uint64_t)(s + 1)) * 0x9E3779B97F4A7C15ULL ^ (uint64_t)buf[i] * 0xC2B2AE3D27D4EB4FULL;
                }
        }
        free(buf);
    }
    if (checksum) *checksum = sum;
    if (n_hits) *n_hits = tot_h;
    if (n_labels) *n_labels = tot_l;
    return tot_k;
}

/* Batched queries: query q owns the sequences [qseq[q], qseq[q+1]) of (seq, off).  Per query the SET of the
 * k-mer hashes of its records (search/query_by_sequence.py:170-176) is looked up and counted per unitig
 * (kmer_idx.count_cdbg_matches, :189).  Output per query q, stored from index kq[q] (kq = the query's first
 * k-mer index, a worst-case-sized region): ids ascending in out_ids, their counts in out_cnt, n_pairs[q] of
 * them; n_distinct[q] = size of the set.  Queries run in parallel (OpenMP); returns the total k-mers. */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

int64_t sgc_oracle_query_counts(const sgc_oracle_table *t, const uint8_t *seq, const int64_t *off,
                                const int64_t *qseq, int64_t nq, int k, const int64_t *kq, uint32_t *out_ids,
                                uint32_t *out_cnt, int64_t *n_pairs, int64_t *n_distinct, int threads) {
    int64_t total = 0;
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int64_t q = 0; q < nq; q++) {
        int64_t nk = 0;
        for (int64_t s = qseq[q]; s < qseq[q + 1]; s++) {
            int64_t L = off[s + 1] - off[s];
            if (L >= k) nk += L - k + 1;
        }
        n_pairs[q] = 0;
        n_distinct[q] = 0;
        total += nk;
        if (nk == 0) continue;
        uint64_t *h = (uint64_t *)malloc((size_t)nk * 8);
        int64_t w = 0;
        for (int64_t s = qseq[q]; s < qseq[q + 1]; s++) {
            int64_t L = off[s + 1] - off[s];
            if (L >= k) w += sgc_oracle_hash_sequence(seq + off[s], L, k, h + w);
        }
        qsort(h, (size_t)nk, 8, cmp_u64);
        uint32_t *ids = out_ids + kq[q];
        int64_t nd = 0, nh = 0;
        for (int64_t i = 0; i < nk; i++) {
            if (i && h[i] == h[i - 1]) continue;
            nd++;
            uint32_t v = table_get(t, h[i]);
            if (v != SGC_ORACLE_MISS) ids[nh++] = v;
        }
        free(h);
        qsort(ids, (size_t)nh, 4, cmp_u32);
        uint32_t *cnt = out_cnt + kq[q];
        int64_t np = 0;
        for (int64_t i = 0; i < nh; i++) {
            if (np && ids[np - 1] == ids[i]) cnt[np - 1]++;
            else {
                ids[np] = ids[i];
                cnt[np] = 1;
                np++;
            }
        }
        n_pairs[q] = np;
        n_distinct[q] = nd;
    }
    return total;
}

int sgc_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
