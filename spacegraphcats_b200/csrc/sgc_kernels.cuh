// sgc_kernels.cuh -- plan kernels, stand-alone table kernels, label tail, per-dominator sums, K7 helpers and the synthetic workload generators.
// Part of libsgc_b200.so (device code; included by sgc_b200.cu only).
#pragma once
#include "sgc_kmer.cuh"
#include "sgc_table.cuh"
#include "sgc_tile.cuh"

namespace sgc {
namespace dev {

// ----------------------------------------------------------------------------------
// plan kernels: per-sequence k-mer / segment counts and the segment map
// ----------------------------------------------------------------------------------
__global__ void plan_count_kernel(const int64_t* __restrict__ off, int64_t nseq, int k, int64_t seq_bytes, int seg_k,
                                  int64_t* __restrict__ nk_out, int64_t* __restrict__ nseg_out, Scalars* scal) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nseq) return;
    int64_t nk = 0;
    if (s < nseq) {
        const int64_t o0 = off[s], o1 = off[s + 1];
        if (o0 < 0 || o1 < o0 || o1 > seq_bytes) {
            atomicOr(&scal->bad_offsets, 1);  // never touch bytes outside the caller's buffer
        } else {
            nk = o1 - o0 - k + 1;
            if (nk < 0) nk = 0;
        }
    }
    nk_out[s] = nk;
    nseg_out[s] = (nk + seg_k - 1) / seg_k;
}

__global__ void plan_expand_kernel(const int64_t* __restrict__ seg_first, const int64_t* __restrict__ off,
                                   const int64_t* __restrict__ kmer_off, int64_t nseq, int k, int seg_k,
                                   SegDesc* __restrict__ segdesc) {
    // one thread per sequence; a sequence of more than 32 segments (contigs, genome-sized queries) is expanded by
    // its whole warp
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int64_t a = 0, b = 0, o0 = 0, nkseq = 0, ko = 0;
    if (s < nseq) {
        a = seg_first[s];
        b = seg_first[s + 1];
        o0 = off[s];
        nkseq = off[s + 1] - o0 - k + 1;
        ko = kmer_off[s];
    }
    auto emit = [&](int64_t j, int64_t a_, int64_t o0_, int64_t nk_, int64_t ko_, unsigned seq_) {
        const int64_t q0 = (j - a_) * seg_k;
        SegDesc d;
        d.start = o0_ + q0;
        d.kout = ko_ + q0;
        d.seq = seq_;
        d.nk = (int)(nk_ - q0 < seg_k ? nk_ - q0 : seg_k);
        d.pad[0] = d.pad[1] = 0;
        segdesc[j] = d;
    };
    const bool big = b - a > 32;
    if (!big)
        for (int64_t j = a; j < b; j++) emit(j, a, o0, nkseq, ko, (unsigned)s);
    unsigned m = __ballot_sync(0xffffffffu, big);
    while (m) {
        const int src = __ffs((int)m) - 1;
        m &= m - 1u;
        const int64_t a_ = __shfl_sync(0xffffffffu, a, src), b_ = __shfl_sync(0xffffffffu, b, src);
        const int64_t o0_ = __shfl_sync(0xffffffffu, o0, src), nk_ = __shfl_sync(0xffffffffu, nkseq, src);
        const int64_t ko_ = __shfl_sync(0xffffffffu, ko, src);
        const unsigned seq_ = (unsigned)(s - lane + src);
        for (int64_t j = a_ + lane; j < b_; j += 32) emit(j, a_, o0_, nk_, ko_, seq_);
    }
}


// the same map, one thread per SEGMENT (batches of long sequences: queries, contigs): the segment's sequence is the
// last one whose first segment is not after it
__global__ void plan_expand_segs_kernel(const int64_t* __restrict__ seg_first, const int64_t* __restrict__ off,
                                        const int64_t* __restrict__ kmer_off, int64_t nseq, int k, int seg_k,
                                        SegDesc* __restrict__ segdesc) {
    const int64_t nsegs = seg_first[nseq];
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nsegs; g += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nseq;  // first index in (lo, hi] whose seg_first is > g; seg_first[0] = 0 <= g
        while (hi - lo > 1) {
            const int64_t mid = lo + (hi - lo) / 2;
            if (seg_first[mid] <= g) lo = mid; else hi = mid;
        }
        const int64_t s = lo, o0 = off[s], nkseq = off[s + 1] - o0 - k + 1, q0 = (g - seg_first[s]) * seg_k;
        SegDesc d;
        d.start = o0 + q0;
        d.kout = kmer_off[s] + q0;
        d.seq = (unsigned int)s;
        d.nk = (int)(nkseq - q0 < seg_k ? nkseq - q0 : seg_k);
        d.pad[0] = d.pad[1] = 0;
        segdesc[g] = d;
    }
}

// ----------------------------------------------------------------------------------
// stand-alone table kernels
// ----------------------------------------------------------------------------------
__global__ void table_clear_kernel(Bucket* tbl, uint32_t nb) {
    const Entry empty{0ULL, VAL_EMPTY, 0u};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (uint64_t)nb * 2;
         i += (uint64_t)gridDim.x * blockDim.x)
        reinterpret_cast<Entry*>(tbl)[i] = empty;
}

__global__ void insert_pairs_kernel(Bucket* tbl, uint32_t nb, uint32_t nlines, int hs,
                                    const uint64_t* __restrict__ hash, const uint32_t* __restrict__ id, int64_t n,
                                    Scalars* scal, uint32_t* brk) {
    BrkSink sink;
    sink.brk = brk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = id[i];
        if (v > SGC_MAX_ID) atomicOr(&scal->err, 2);
        else table_insert(tbl, nb, nlines, hs, hash[i], v, AUX_NONE, scal, sink);
    }
}

// one owner's part of a partitioned table: (hash, id, position | orientation) triples that arrived from every rank
__global__ void insert_triples_kernel(Bucket* tbl, uint32_t nb, uint32_t nlines, int hs,
                                      const uint64_t* __restrict__ hash, const uint32_t* __restrict__ id,
                                      const uint32_t* __restrict__ aux, int64_t n, Scalars* scal, const BrkSink sink) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = id[i];
        if (v > SGC_MAX_ID) atomicOr(&scal->err, 2);
        else table_insert(tbl, nb, nlines, hs, hash[i], v, aux[i], scal, sink);
    }
}

// marks of a partitioned build -> break bits of the replicated store (shard s starts stride_words after shard s-1)
__global__ void apply_marks_kernel(const unsigned long long* __restrict__ marks, int64_t n, uint32_t* brk, int64_t stride_words) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        brk_mark(brk + (int64_t)(marks[i] >> 32) * stride_words, (uint32_t)marks[i]);
}

__global__ void lookup_kernel(const Bucket* tbl, uint32_t nb, uint32_t nlines, int hs,
                              const uint64_t* __restrict__ hash, int64_t n, uint32_t* __restrict__ out) {
    // two probes in flight per thread
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
        const int64_t j = i + stride;
        const uint64_t h0 = hash[i], h1 = j < n ? hash[j] : 0;
        const uint32_t b0 = home_bucket(h0, nb, hs), b1 = j < n ? home_bucket(h1, nb, hs) : 0u;
        Entry a0, a1, c0, c1;
        load_bucket_nc(tbl + b0, a0, a1);
        load_bucket_nc(tbl + b1, c0, c1);
        uint32_t v;
        uint32_t aux;
        if (!bucket_match(a0, a1, h0, v)) v = bucket_overflowed(a0) ? lookup_slow(tbl, nlines, b0, h0, aux) : SGC_MISS;
        out[i] = v;
        if (j < n) {
            if (!bucket_match(c0, c1, h1, v)) v = bucket_overflowed(c0) ? lookup_slow(tbl, nlines, b1, h1, aux) : SGC_MISS;
            out[j] = v;
        }
    }
}

// get_unique_values: count_by_id[id]++ per present hash.  With distinct != 0 the input is
// sorted and a hash equal to its predecessor is skipped.
__global__ void count_kernel(const Bucket* tbl, uint32_t nb, uint32_t nlines, int hs,
                             const uint64_t* __restrict__ hash, int64_t n, int distinct, uint32_t* count_by_id,
                             Scalars* scal) {
    unsigned long long miss = 0, nd = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t h = hash[i];
        if (distinct && i > 0 && hash[i - 1] == h) continue;
        nd++;
        uint32_t v = table_lookup(tbl, nb, nlines, hs, h);
        if (v == SGC_MISS) miss++;
        else atomicAdd(&count_by_id[v], 1u);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        miss += __shfl_xor_sync(0xffffffffu, miss, d);
        nd += __shfl_xor_sync(0xffffffffu, nd, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (miss) atomicAdd(&scal->n_miss, miss);
        if (nd) atomicAdd(&scal->n_distinct, nd);
    }
}

__global__ void table_stats_kernel(const Bucket* tbl, uint32_t nb, Scalars* scal) {
    unsigned long long live = 0, multi = 0;
    unsigned mx = 0;
    const Entry* e = reinterpret_cast<const Entry*>(tbl);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (uint64_t)nb * 2;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v = e[i].val;
        if (v == VAL_MULTI) multi++;
        else if (v != VAL_EMPTY) {
            live++;
            mx = v > mx ? v : mx;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        live += __shfl_xor_sync(0xffffffffu, live, d);
        multi += __shfl_xor_sync(0xffffffffu, multi, d);
        unsigned o = __shfl_xor_sync(0xffffffffu, mx, d);
        mx = o > mx ? o : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (live) atomicAdd(&scal->n_live, live);
        if (multi) atomicAdd(&scal->n_multi, multi);
        if (mx) atomicMax(&scal->max_id, mx);
    }
}

__global__ void misses_from_total_kernel(const int64_t* __restrict__ total, Scalars* scal) {
    scal->n_miss = (unsigned long long)*total - scal->n_hits;
}

__global__ void table_export_kernel(const Bucket* tbl, uint32_t nb, uint64_t* hash_out, uint32_t* id_out,
                                    unsigned long long cap, Scalars* scal) {
    const Entry* e = reinterpret_cast<const Entry*>(tbl);
    const uint64_t total = (uint64_t)nb * 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (total + stride - 1) / stride;
    const int lane = threadIdx.x & 31;
    for (uint64_t r = 0; r < rounds; r++) {
        uint64_t i = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        Entry x{0ULL, VAL_EMPTY, 0u};
        if (i < total) x = e[i];
        bool live = x.val < VAL_MULTI;
        unsigned m = __ballot_sync(0xffffffffu, live);
        if (m) {
            unsigned long long base = 0;
            int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(&scal->n_export, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (live) {
                unsigned long long p = base + __popc(m & ((1u << lane) - 1u));
                if (p < cap) {
                    hash_out[p] = x.key;
                    id_out[p] = x.val;
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------------
// label tail: unique (seq, id) keys -> CSR
// ----------------------------------------------------------------------------------
__global__ void csr_count_kernel(const unsigned long long* __restrict__ uniq, const long long* __restrict__ n_p,
                                 unsigned long long* ids_off, uint32_t* ids_out, int64_t ids_cap) {
    const int64_t n = *n_p;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long key = uniq[j];
        atomicAdd(&ids_off[(key >> 32) + 1], 1ULL);
        if (j < ids_cap) ids_out[j] = (uint32_t)key;
    }
}

// ----------------------------------------------------------------------------------
// label tail without a global sort (the common case: a handful of candidate ids per sequence):
// count candidates per sequence -> exclusive scan -> scatter the ids into per-sequence runs -> one
// thread per sequence sorts and de-duplicates its run -> scan of the distinct counts -> compact copy.
// Sequences with more than CSR_SMALL candidates raise csr_big and the caller falls back to the
// radix-sort path (pairs_to_csr_sorted), which handles any size.
// ----------------------------------------------------------------------------------
constexpr int CSR_SMALL = 160;  // any 150-bp read fits (<= 120 k-mers, a few duplicates of a candidate)

__global__ void __launch_bounds__(256) csr2_count_kernel(const unsigned long long* __restrict__ pairs, long long n,
                                                         unsigned long long* __restrict__ cnt) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x)
        atomicAdd(&cnt[pairs[j] >> 32], 1ULL);
}

// cnt[] counts down to zero again: position = start + (remaining - 1)
__global__ void __launch_bounds__(256) csr2_scatter_kernel(const unsigned long long* __restrict__ pairs, long long n,
                                                           const long long* __restrict__ start,
                                                           unsigned long long* __restrict__ cnt,
                                                           uint32_t* __restrict__ run_ids) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = pairs[j];
        const unsigned long long s = key >> 32;
        const unsigned long long left = atomicAdd(&cnt[s], ~0ULL);  // -1
        run_ids[start[s] + (long long)left - 1] = (uint32_t)key;
    }
}

// sorts + de-duplicates run s in place, writes its distinct count to ids_off[s + 1]
__global__ void __launch_bounds__(256) csr2_finalize_kernel(const long long* __restrict__ start, long long nseq,
                                                            uint32_t* __restrict__ run_ids,
                                                            unsigned long long* __restrict__ ids_off, Scalars* scal) {
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < nseq; s += (long long)gridDim.x * blockDim.x) {
        const long long b = start[s];
        const int n = (int)(start[s + 1] - b);
        unsigned long long nu = 0;
        if (n > CSR_SMALL) {
            scal->csr_big = 1u;
        } else if (n > 0) {
            uint32_t v[CSR_SMALL];
            for (int j = 0; j < n; j++) {  // insertion sort: runs are a few ids long
                const uint32_t x = run_ids[b + j];
                int p = j;
                while (p > 0 && v[p - 1] > x) {
                    v[p] = v[p - 1];
                    p--;
                }
                v[p] = x;
            }
            int w = 0;
            for (int j = 0; j < n; j++)
                if (j == 0 || v[j] != v[j - 1]) run_ids[b + w++] = v[j];
            nu = (unsigned long long)w;
        }
        ids_off[s + 1] = nu;
    }
}

__global__ void __launch_bounds__(256) csr2_copy_kernel(const long long* __restrict__ start,
                                                        const unsigned long long* __restrict__ ids_off, long long nseq,
                                                        const uint32_t* __restrict__ run_ids, uint32_t* __restrict__ ids_out,
                                                        long long ids_cap) {
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < nseq; s += (long long)gridDim.x * blockDim.x) {
        const long long o = (long long)ids_off[s], n = (long long)ids_off[s + 1] - o, b = start[s];
        for (long long j = 0; j < n; j++)
            if (o + j < ids_cap) ids_out[o + j] = run_ids[b + j];
    }
}

// per-k-mer ids (already looked up, possibly on other ranks) -> candidate (seq, id) keys
__global__ void ids_to_pairs_kernel(const uint32_t* __restrict__ kmer_ids, const int64_t* __restrict__ kmer_off,
                                    int64_t nseq, unsigned long long* pairs, unsigned long long cap, Scalars* scal) {
    unsigned long long hits = 0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nseq; s += (int64_t)gridDim.x * blockDim.x) {
        uint32_t prev = SGC_MISS;
        for (int64_t j = kmer_off[s]; j < kmer_off[s + 1]; j++) {
            uint32_t v = kmer_ids[j];
            if (v != SGC_MISS) {
                hits++;
                if (v != prev) {
                    unsigned long long p = atomicAdd(&scal->pair_count, 1ULL);
                    if (p < cap) pairs[p] = ((unsigned long long)s << 32) | v;
                }
            }
            prev = v;
        }
    }
    if (hits) atomicAdd(&scal->n_hits, hits);
}

// ----------------------------------------------------------------------------------
// K6: per-dominator sums, one launch per catlas level
// ----------------------------------------------------------------------------------
__global__ void dom_level_kernel(const uint32_t* __restrict__ src, const int64_t* __restrict__ member_off,
                                 const uint32_t* __restrict__ member, const uint32_t* __restrict__ nodes,
                                 int64_t n, uint32_t* node_count) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t v = nodes[j];
    uint32_t s = 0;
    for (int64_t m = member_off[v]; m < member_off[v + 1]; m++) s += src[member[m]];
    node_count[v] = s;
}

// ----------------------------------------------------------------------------------
// K7 helpers
// ----------------------------------------------------------------------------------
__global__ void iota_kernel(int64_t* p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = i;
}

// sorted by the top `bits` bits: bucket_off[r] = first index whose owner >= r
__global__ void bucket_bounds_kernel(const uint64_t* __restrict__ sorted, int64_t n, int bits, int n_ranks,
                                     int64_t* bucket_off) {
    int r = threadIdx.x;
    if (r > n_ranks) return;
    if (r == n_ranks) { bucket_off[r] = n; return; }
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        unsigned owner = bits ? (unsigned)(sorted[mid] >> (64 - bits)) : 0u;
        if (owner < (unsigned)r) lo = mid + 1; else hi = mid;
    }
    bucket_off[r] = lo;
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ in, const int64_t* __restrict__ perm, int64_t n,
                                  uint32_t* out, int scatter) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (scatter) out[perm[i]] = in[i];
        else out[i] = in[perm[i]];
    }
}

// ----------------------------------------------------------------------------------
// Batched queries (search/query_by_sequence.py:153-189 over many queries at once): per query the
// DISTINCT k-mer hashes of all its records are looked up and counted per unitig.
// ----------------------------------------------------------------------------------
// query index of every k-mer: sequence s (k-mers [koff[s], koff[s+1])) belongs to query seq_query[s]
__global__ void kmer_query_kernel(const int64_t* __restrict__ koff, const int32_t* __restrict__ seq_query, int64_t nseq,
                                  uint32_t* __restrict__ qid) {
    // one warp per sequence
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t s = warp; s < nseq; s += nwarps) {
        const uint32_t q = (uint32_t)seq_query[s];
        for (int64_t i = koff[s] + lane; i < koff[s + 1]; i += 32) qid[i] = q;
    }
}

constexpr unsigned long long QMAP_EMPTY = ~0ULL;
constexpr int QHIST_MAX = 8192;  // queries whose distinct counters fit a CTA's shared memory (32 KB)

__global__ void qmap_clear_kernel(unsigned long long* keys, uint32_t* cnt, uint64_t slots) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (uint64_t)gridDim.x * blockDim.x) {
        keys[i] = QMAP_EMPTY;
        cnt[i] = 0u;
    }
}

// hs / qs: the k-mer hashes and their query indices after a STABLE radix sort on the hashes' top 32 bits.
// The input was in (query, position) order, so inside a run of equal top bits the elements are still
// ordered by query: an element is a duplicate of its query's set iff the nearest EARLIER element of the
// run with the same full hash belongs to the same query.  Every first occurrence is looked up (sorted
// hashes walk the table's buckets in ascending order) and counted under (query, unitig) in a scratch map.
__global__ void __launch_bounds__(256) query_count_kernel(const Bucket* tbl, uint32_t nb, uint32_t nlines, int hs_shift,
                                                          const uint64_t* __restrict__ hs, const uint32_t* __restrict__ qs,
                                                          int64_t n, unsigned long long* q_distinct, int n_queries,
                                                          unsigned long long* map_keys, uint32_t* map_cnt, uint64_t map_mask,
                                                          Scalars* scal) {
    // per-query distinct counters: a few hundred hot addresses hit by every element -- counted in shared memory
    // per CTA and flushed once (2.6e8 global atomics on 512 addresses cost 40 ms; measured)
    extern __shared__ unsigned int s_hist[];
    const bool use_smem = n_queries <= QHIST_MAX;
    if (use_smem) {
        for (int j = threadIdx.x; j < n_queries; j += blockDim.x) s_hist[j] = 0u;
        __syncthreads();
    }
    unsigned long long hits = 0, miss = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = hs[i];
        const uint32_t q = qs[i];
        bool dup = false;
        for (int64_t j = i - 1; j >= 0 && (hs[j] >> 32) == (h >> 32); j--) {
            if (hs[j] == h) {
                dup = qs[j] == q;
                break;
            }
        }
        if (dup) continue;
        if (use_smem) atomicAdd(&s_hist[q], 1u);
        else atomicAdd(&q_distinct[q], 1ULL);
        const uint32_t v = table_lookup(tbl, nb, nlines, hs_shift, h);
        if (v == SGC_MISS) {
            miss++;
            continue;
        }
        hits++;
        const unsigned long long key = ((unsigned long long)q << 32) | v;
        uint64_t slot = (key * 0x9E3779B97F4A7C15ULL) >> 20;
        bool placed = false;
        for (uint64_t step = 0; step <= map_mask; step++, slot++) {
            slot &= map_mask;
            unsigned long long old = map_keys[slot];
            if (old == QMAP_EMPTY) {
                old = atomicCAS(&map_keys[slot], QMAP_EMPTY, key);
                // a new (query, unitig) pair: the map is kept below 5/8 full so that probe runs stay short and
                // the map small enough to live in L2; beyond that the host redoes the pass with a larger one
                if (old == QMAP_EMPTY && atomicAdd(&scal->n_distinct, 1ULL) > (map_mask >> 3) * 5) atomicOr(&scal->err, 1);
            }
            if (old == QMAP_EMPTY || old == key) {
                atomicAdd(&map_cnt[slot], 1u);
                placed = true;
                break;
            }
            if (step > 256) break;
        }
        if (!placed) atomicOr(&scal->err, 1);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        hits += __shfl_xor_sync(0xffffffffu, hits, d);
        miss += __shfl_xor_sync(0xffffffffu, miss, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (hits) atomicAdd(&scal->n_hits, hits);
        if (miss) atomicAdd(&scal->n_miss, miss);
    }
    if (use_smem) {
        __syncthreads();
        for (int j = threadIdx.x; j < n_queries; j += blockDim.x)
            if (s_hist[j]) atomicAdd(&q_distinct[j], (unsigned long long)s_hist[j]);
    }
}

// ---- the store-position form of the batched query (labelseq_kernel<.., QUERY>) ----
// intervals sorted by key = query << pbits | first position (pbits = bits of the store's size); val = length << 32 |
// unitig id.  e[i] = query << pbits | end (exclusive) -- a running maximum of e over the sorted intervals is, above
// pbits, the query of the latest interval and, below, how far that query's earlier intervals reach.
__global__ void iv_end_kernel(const unsigned long long* __restrict__ key, const unsigned long long* __restrict__ val, int64_t n,
                              unsigned long long* __restrict__ e) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        e[i] = key[i] + (val[i] >> 32);  // (the store is padded: an end never carries into the query bits)
}
// the part of interval i that no earlier interval of its query covers -> (query << 32 | unitig, length)
__global__ void iv_union_kernel(const unsigned long long* __restrict__ key, const unsigned long long* __restrict__ val,
                                const unsigned long long* __restrict__ pmax, int64_t n, int pbits,
                                unsigned long long* __restrict__ key2, uint32_t* __restrict__ cnt) {
    const unsigned long long pmask = (1ULL << pbits) - 1ULL;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = key[i], v = val[i], pm = pmax[i];
        const uint32_t start = (uint32_t)(k & pmask), end = start + (uint32_t)(v >> 32);
        uint32_t from = start;
        if ((pm >> pbits) == (k >> pbits) && (uint32_t)(pm & pmask) > from) from = (uint32_t)(pm & pmask);
        key2[i] = ((k >> pbits) << 32) | (v & 0xFFFFFFFFULL);
        cnt[i] = end > from ? end - from : 0u;
    }
}
// final (query << 32 | unitig, count) pairs -> per-query distinct matching k-mers and the total
__global__ void pairs_tally_kernel(const unsigned long long* __restrict__ key, const uint32_t* __restrict__ cnt, int64_t n,
                                   unsigned long long* q_distinct, int n_queries, Scalars* scal) {
    extern __shared__ unsigned int s_hist[];   // per-CTA counters when the queries fit (see query_count_kernel)
    const bool use_smem = n_queries <= QHIST_MAX;
    if (use_smem) {
        for (int j = threadIdx.x; j < n_queries; j += blockDim.x) s_hist[j] = 0u;
        __syncthreads();
    }
    unsigned long long hits = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (use_smem) atomicAdd(&s_hist[key[i] >> 32], cnt[i]);
        else atomicAdd(&q_distinct[key[i] >> 32], (unsigned long long)cnt[i]);
        hits += cnt[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, d);
    if ((threadIdx.x & 31) == 0 && hits) atomicAdd(&scal->n_hits, hits);
    if (use_smem) {
        __syncthreads();
        for (int j = threadIdx.x; j < n_queries; j += blockDim.x)
            if (s_hist[j]) atomicAdd(&q_distinct[j], (unsigned long long)s_hist[j]);
    }
}
// ks / hs: the missing k-mers sorted by their 32-bit key = query << hbits | top hbits bits of the hash, and their
// full hashes.  Two elements are the same (hash, query) iff they lie in one run of equal keys and their hashes are
// equal; every first occurrence counts.  Runs are short (a key space of 2^32 for the k-mers of one batch).
__global__ void __launch_bounds__(256) miss_distinct_kernel(const uint32_t* __restrict__ ks, const uint64_t* __restrict__ hs,
                                                            int64_t n, int hbits, unsigned long long* q_distinct, int n_queries) {
    extern __shared__ unsigned int s_hist[];
    const bool use_smem = n_queries <= QHIST_MAX;
    if (use_smem) {
        for (int j = threadIdx.x; j < n_queries; j += blockDim.x) s_hist[j] = 0u;
        __syncthreads();
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t key = ks[i];
        const uint64_t h = hs[i];
        bool dup = false;
        for (int64_t j = i - 1; j >= 0 && ks[j] == key; j--)
            if (hs[j] == h) {
                dup = true;
                break;
            }
        if (dup) continue;
        const uint32_t q = key >> hbits;
        if (use_smem) atomicAdd(&s_hist[q], 1u);
        else atomicAdd(&q_distinct[q], 1ULL);
    }
    if (use_smem) {
        __syncthreads();
        for (int j = threadIdx.x; j < n_queries; j += blockDim.x)
            if (s_hist[j]) atomicAdd(&q_distinct[j], (unsigned long long)s_hist[j]);
    }
}

__global__ void qmap_compact_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ cnt, uint64_t slots,
                                    unsigned long long* out_keys, uint32_t* out_cnt, unsigned long long cap, Scalars* scal) {
    const int lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (slots + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t i = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const unsigned long long k = i < slots ? keys[i] : QMAP_EMPTY;
        const bool live = k != QMAP_EMPTY;
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (m) {
            unsigned long long base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(&scal->n_export, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (live) {
                const unsigned long long p = base + __popc(m & ((1u << lane) - 1u));
                if (p < cap) {
                    out_keys[p] = k;
                    out_cnt[p] = cnt[i];
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------------
// K7 owner side, flag protocol: probe every source rank's region of my inbox as soon as that rank has
// stamped it (count + epoch in my mailbox), store each id straight into the requester's reply buffer
// (peer memory), and stamp the requesters' reply buffers once everything has been stored.
// ----------------------------------------------------------------------------------
struct ServeArgs {
    const Bucket* tbl;
    uint32_t nbuckets, nlines;
    int home_shift;
    const uint64_t* inbox;             // n_ranks regions of region_cap hashes (local)
    unsigned long long region_cap;
    const unsigned long long* mail_count;  // local mailbox: count from rank s
    const unsigned int* mail_stamp;        // ... stamped with the epoch when rank s's requests are all here
    uint32_t* peer_replies[32];        // requester s's reply buffer (my region: [my_rank * region_cap ..))
    unsigned int* peer_reply_stamp[32];
    Scalars* scal;
    int n_ranks, my_rank;
    unsigned int epoch;
};

__global__ void __launch_bounds__(256) serve_regions_kernel(const ServeArgs a) {
    __shared__ unsigned long long s_n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int j = 0; j < a.n_ranks; j++) {
        const int s = (a.my_rank + j) % a.n_ranks;  // own region first: it needs no link
        if (threadIdx.x == 0) {
            spin_until_equal(a.mail_stamp + s, a.epoch, &a.scal->err);
            unsigned long long n = *reinterpret_cast<const volatile unsigned long long*>(a.mail_count + s);
            s_n = n > a.region_cap ? a.region_cap : n;
        }
        __syncthreads();
        const int64_t n = (int64_t)s_n;
        const uint64_t* hash = a.inbox + (size_t)s * a.region_cap;
        uint32_t* out = a.peer_replies[s] + (size_t)a.my_rank * a.region_cap;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
            const int64_t i2 = i + stride;
            const uint64_t h0 = hash[i], h1 = i2 < n ? hash[i2] : 0;
            const uint32_t b0 = home_bucket(h0, a.nbuckets, a.home_shift), b1 = i2 < n ? home_bucket(h1, a.nbuckets, a.home_shift) : 0u;
            Entry x0, x1, y0, y1;
            load_bucket_nc(a.tbl + b0, x0, x1);
            load_bucket_nc(a.tbl + b1, y0, y1);
            uint32_t v, aux;
            if (!bucket_match(x0, x1, h0, v)) v = bucket_overflowed(x0) ? lookup_slow(a.tbl, a.nlines, b0, h0, aux) : SGC_MISS;
            out[i] = v;
            if (i2 < n) {
                if (!bucket_match(y0, y1, h1, v)) v = bucket_overflowed(y0) ? lookup_slow(a.tbl, a.nlines, b1, h1, aux) : SGC_MISS;
                out[i2] = v;
            }
        }
        __syncthreads();  // s_n is rewritten for the next region
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&a.scal->done_ctr2, 1u);
        if (prev == gridDim.x - 1u) {  // the last CTA: every reply of this rank has been stored
            __threadfence_system();
            for (int s = 0; s < a.n_ranks; s++)
                *reinterpret_cast<volatile unsigned int*>(a.peer_reply_stamp[s] + a.my_rank) = a.epoch;
        }
    }
}

// ----------------------------------------------------------------------------------
// synthetic workload generators (bench / tests only)
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__global__ void synth_genome_kernel(uint8_t* out, int64_t n, uint64_t seed) {
    // 32 bases per 64-bit draw
    const int64_t nw = (n + 31) / 32;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += (int64_t)gridDim.x * blockDim.x) {
        uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ULL + (uint64_t)w);
        for (int j = 0; j < 32; j++) {
            int64_t i = w * 32 + j;
            if (i < n) out[i] = "ACGT"[(r >> (2 * j)) & 3];
        }
    }
}

// read r = genome[p .. p+len) at a uniform position, reverse-complemented when bit 0 of its
// draw is set (rc_half), each base substituted by one of the 3 others with prob err_ppm/1e6.
__global__ void synth_reads_kernel(const uint8_t* __restrict__ genome, int64_t glen, int64_t nreads, int read_len,
                                   uint64_t seed, uint32_t err_ppm, int rc_half, uint8_t* out, int64_t* off) {
    const int64_t total = nreads * read_len;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / read_len;
        int j = (int)(i - r * read_len);
        uint64_t d = splitmix64(seed ^ splitmix64((uint64_t)r));
        int64_t p = (int64_t)((d >> 1) % (uint64_t)(glen - read_len + 1));
        const bool rc = rc_half && (d & 1);
        // errors are drawn in GENOME orientation, so the rc_half=0 and rc_half=1 versions of a batch
        // are exact reverse complements of each other read by read
        const int gj = rc ? read_len - 1 - j : j;
        uint8_t b = genome[p + gj];
        const uint64_t e = splitmix64(seed * 31 + 7 + (uint64_t)(r * read_len + gj));
        if ((uint32_t)(e % 1000000ULL) < err_ppm) {
            int cur = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3;
            b = "ACGT"[(cur + 1 + (int)((e >> 40) % 3)) & 3];
        }
        if (rc) b = sgc::comp_base(b);
        out[i] = b;
        if (j == 0) off[r] = r * (int64_t)read_len;
        if (i == total - 1) off[nreads] = total;
    }
}

// unitig j = genome[cuts[j] .. cuts[j+1] + k - 1), written at out_off[j] = cuts[j] + j*(k-1)
__global__ void synth_unitigs_kernel(const uint8_t* __restrict__ genome, const int64_t* __restrict__ cuts, int64_t n,
                                     int k, uint8_t* out, int64_t* out_off) {
    const int64_t total = cuts[n] - cuts[0] + n * (int64_t)(k - 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // largest j with cuts[j]-cuts[0] + j*(k-1) <= i
        int64_t lo = 0, hi = n - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi + 1) >> 1;
            if (cuts[mid] - cuts[0] + mid * (int64_t)(k - 1) <= i) lo = mid; else hi = mid - 1;
        }
        int64_t o = cuts[lo] - cuts[0] + lo * (int64_t)(k - 1);
        out[i] = genome[cuts[lo] + (i - o)];
        if (i == o) out_off[lo] = o;
        if (i == total - 1) out_off[n] = total;
    }
}


// ----------------------------------------------------------------------------------
// 2-bit packed reads (A=0 C=1 G=2 T=3; base b of record s in bits 2(b%4).. of byte poff[s] + b/4;
// records start on 16-byte boundaries)
// ----------------------------------------------------------------------------------
// per-sequence length (len[s], or uniform_len when len is NULL) -> k-mers, segments, packed 16-byte units
__global__ void packed_count_kernel(const int32_t* __restrict__ len, int uniform_len, int64_t nseq, int k, int seg_k,
                                    int64_t* __restrict__ nk_out, int64_t* __restrict__ nseg_out,
                                    int64_t* __restrict__ units_out, Scalars* scal) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nseq) return;
    int64_t nk = 0, units = 0;
    if (s < nseq) {
        const int64_t L = len ? (int64_t)len[s] : (int64_t)uniform_len;
        if (L < 0) atomicOr(&scal->bad_offsets, 1);
        else {
            nk = L - k + 1 > 0 ? L - k + 1 : 0;
            units = (L + 15) / 16;
        }
    }
    nk_out[s] = nk;
    nseg_out[s] = (nk + seg_k - 1) / seg_k;
    units_out[s] = units;
}

// segment map of packed records (units of one 32-bit word = 16 bases): start = byte offset of the segment's first
// base (seg_k % 16 == 0, so it stays word aligned)
__global__ void packed_expand_kernel(const int64_t* __restrict__ seg_first, const int64_t* __restrict__ unit_off,
                                     const int64_t* __restrict__ nk_arr, const int64_t* __restrict__ kmer_off,
                                     int64_t nseq, int seg_k, int64_t packed_bytes, SegDesc* __restrict__ segdesc,
                                     Scalars* scal) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseq) return;
    const int64_t a = seg_first[s], b = seg_first[s + 1];
    if (unit_off[s + 1] * 4 > packed_bytes) atomicOr(&scal->bad_offsets, 1);
    if (a == b) return;
    const int64_t nkseq = nk_arr[s], ko = kmer_off[s], p0 = unit_off[s] * 4;
    for (int64_t j = a; j < b; j++) {
        const int64_t q0 = (j - a) * seg_k;
        SegDesc d;
        d.start = p0 + q0 / 4;
        d.kout = ko + q0;
        d.seq = (unsigned int)s;
        d.nk = unit_off[s + 1] * 4 > packed_bytes ? 0 : (int)(nkseq - q0 < seg_k ? nkseq - q0 : seg_k);
        d.pad[0] = d.pad[1] = 0;
        segdesc[j] = d;
    }
}

// records of ONE length: the whole plan in closed form, one pass (no scans)
__global__ void packed_uniform_plan_kernel(int uniform_len, int64_t nseq, int k, int seg_k, int64_t packed_bytes,
                                           int64_t* __restrict__ kmer_off, int64_t* __restrict__ seg_first,
                                           int64_t* __restrict__ unit_off, SegDesc* __restrict__ segdesc, Scalars* scal) {
    const int64_t nk = uniform_len - k + 1 > 0 ? uniform_len - k + 1 : 0, nseg = (nk + seg_k - 1) / seg_k;
    const int64_t units = ((int64_t)uniform_len + 15) / 16;
    const bool bad = nseq * units * 4 > packed_bytes;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s <= nseq; s += (int64_t)gridDim.x * blockDim.x) {
        kmer_off[s] = s * nk;
        seg_first[s] = s * nseg;
        unit_off[s] = s * units;
        if (s == nseq) {
            if (bad) atomicOr(&scal->bad_offsets, 1);
            break;
        }
        for (int64_t j = 0; j < nseg; j++) {
            const int64_t q0 = j * seg_k;
            SegDesc d;
            d.start = s * units * 4 + q0 / 4;
            d.kout = s * nk + q0;
            d.seq = (unsigned int)s;
            d.nk = bad ? 0 : (int)(nk - q0 < seg_k ? nk - q0 : seg_k);
            d.pad[0] = d.pad[1] = 0;
            segdesc[s * nseg + j] = d;
        }
    }
}

// CSR offsets -> one byte per sequence (its count); counts[n] = 1 if some count does not fit (> 255)
__global__ void offsets_to_counts8_kernel(const int64_t* __restrict__ off, int64_t n, uint8_t* __restrict__ counts) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = off[s + 1] - off[s];
        counts[s] = (uint8_t)c;
        if (c > 255 || c < 0) counts[n] = 1;
    }
}

__global__ void len_to_i64_kernel(const int32_t* __restrict__ len, int uniform_len, int64_t nseq, int64_t* __restrict__ out) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nseq) return;
    out[s] = s < nseq ? (len ? (int64_t)len[s] : (int64_t)uniform_len) : 0;
}

// base offsets -> 4-byte units per record (and the int32 lengths the packed API takes)
__global__ void off_to_units_kernel(const int64_t* __restrict__ off, int64_t nseq, int64_t seq_bytes,
                                    int64_t* __restrict__ units, int32_t* len_out, Scalars* scal) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nseq) return;
    int64_t u = 0;
    if (s < nseq) {
        const int64_t o0 = off[s], o1 = off[s + 1];
        if (o0 < 0 || o1 < o0 || o1 > seq_bytes || o1 - o0 > 0x7FFFFFFF) atomicOr(&scal->bad_offsets, 1);
        else {
            u = (o1 - o0 + 15) / 16;
            if (len_out) len_out[s] = (int32_t)(o1 - o0);
        }
    }
    units[s] = u;
}

// ASCII (seq, off) on the device -> packed records; status[s] bit 0 = a byte other than A/C/G/T (packed as A).
// One thread per packed 32-bit word (16 bases).
__global__ void pack_reads_kernel(const uint8_t* __restrict__ seq, const int64_t* __restrict__ off,
                                  const int64_t* __restrict__ unit_off, int64_t nseq, uint32_t* __restrict__ packed,
                                  int64_t total_words, int32_t* status) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total_words; w += (int64_t)gridDim.x * blockDim.x) {
        // record of word w: largest s with unit_off[s] <= w
        int64_t lo = 0, hi = nseq - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (unit_off[mid] <= w) lo = mid; else hi = mid - 1;
        }
        const int64_t b0 = (w - unit_off[lo]) * 16, L = off[lo + 1] - off[lo];
        uint32_t x = 0;
        bool bad = false;
        for (int j = 0; j < 16; j++) {
            if (b0 + j < L) {
                const uint8_t c = seq[off[lo] + b0 + j];
                const uint32_t code = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
                if (code > 3u) bad = true;
                x |= (code & 3u) << (2 * j);
            }
        }
        packed[w] = x;
        if (bad && status) atomicOr(&status[lo], 1);
    }
}

// packed records -> ASCII (seq, off) (general fallback for tables without a side array)
__global__ void unpack_reads_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ off,
                                    const int64_t* __restrict__ unit_off, int64_t nseq, uint8_t* __restrict__ seq) {
    const int64_t total = off[nseq];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nseq - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (off[mid] <= i) lo = mid; else hi = mid - 1;
        }
        const int64_t b = i - off[lo];
        const uint32_t x = packed[unit_off[lo] + (b >> 4)];
        seq[i] = "ACGT"[(x >> (2 * (b & 15))) & 3u];
    }
}


}  // namespace dev
}  // namespace sgc
