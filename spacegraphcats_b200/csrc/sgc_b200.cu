// sgc_b200.cu -- host side and C-ABI (include/sgc_b200.h) of libsgc_b200.so, the sm_100a
// implementation of spacegraphcats' k-mer -> cDBG-unitig indexing hot path.  Device code lives in
//   sgc_kmer.cuh    the reference hash (2 x Murmur3_x64_128) and the k-mer window geometry
//   sgc_layout.cuh  constants, table slot layout, device scalars, tile-kernel argument block
//   sgc_table.cuh   table primitives (probe sequence, CAS insert with build_mphf semantics, lookup)
//   sgc_tile.cuh    segment staging + the warp-autonomous tile kernel (all fused modes)
//   sgc_labelseq.cuh  the read-labelling kernel over the unitig sequence store (ASCII or 2-bit packed reads)
//   sgc_kernels.cuh plan / stand-alone table / label-tail / dominator / K7 / generator kernels
//   sgc_mphf.cuh    BBHash (boomphf) level construction, rank samples, lookup
//   sgc_feed.hpp    host-only feeder (BGZF inflate, FASTA/FASTQ scan, mate pairing)
// See DESIGN.md for the data layout.
//
//   K1 kmer hash        tile_kernel<K, MODE_HASH>       (khmer get_kmer_hashes, cdbg/hashing.py:14-20)
//   K2 table build      tile_kernel<K, MODE_INSERT>, insert_pairs_kernel
//                                                       (build_mphf, cdbg/index_cdbg_by_kmer.py:27-115)
//   K3 lookup           lookup_kernel / fused in tile_kernel (BBHashTable.__getitem__, hashing.py:36-37)
//   K4 per-read ids     tile_kernel<K, MODE_LABEL> + sort/unique/CSR (cdbg/index_reads.py:143-152)
//   K5 per-unitig count count_kernel, tile_kernel<K, MODE_LOOKUP> (get_unique_values, hashing.py:67-69)
//   K6 per-dominator    dom_level_kernel                (count_catlas_matches, hashing.py:71-107)
//   K7 owner bucketing  partition kernels               (new: hash-range partition across GPUs)
//   K8 BBHash writer    mphf_* kernels (sgc_mphf.cuh)   (bbhash.PyMPHF + save, the on-disk contigs.mphf)
//
// No CPU fallback: every entry point needs a CUDA device.
#include <cuda.h>  // driver API types only: the entry points are resolved at run time (no link-time libcuda)
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <new>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/sgc_b200.h"
#include "sgc_kmer.cuh"

namespace {

// ----------------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------------
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(SGC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

#define TRY(call)              \
    do {                       \
        int r_ = (call);       \
        if (r_ != SGC_OK) return r_; \
    } while (0)

}  // namespace

#include "sgc_layout.cuh"
#include "sgc_table.cuh"
#include "sgc_tile.cuh"
#include "sgc_labelseq.cuh"
#include "sgc_kernels.cuh"
#include "sgc_mphf.cuh"

using namespace sgc::dev;


// ====================================================================================
// host side
// ====================================================================================
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct sgc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 0;
    int64_t launches = 0;
    Scalars* d_scal = nullptr;
    Scalars* h_scal = nullptr;  // pinned
    DevBuf nk, nseg, seg_first, kmer_off, seg_map, cub_tmp, pairs_a, pairs_b, tmp_hash, tmp_id, tmp_idx, cursor;
    DevBuf units, unit_off, base_off, unpacked;  // packed-read plans
    DevBuf q_hash2, q_id, q_id2, q_mapk, q_mapc;  // batched queries
    DevBuf iv_a, iv_b, iv_c, iv_d;                // batched queries through the sequence store: interval keys / values
    // optional CUDA-event timing of the tile kernel (bench.py roofline leg)
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
};

struct sgc_table {
    sgc_ctx* ctx = nullptr;
    Bucket* buckets = nullptr;
    uint32_t nbuckets = 0;
    uint32_t nlines = 0;
    int home_shift = 0;
    int64_t max_entries = 0;
    // unitig sequence store (tables built from sequences): every inserted sequence buffer, 2 bits per base,
    // + one break bit per base position (sgc_labelseq.cuh)
    bool store_on = false;
    uint32_t* useq_alloc = nullptr;  // USEQ_PAD words of padding on either end
    uint32_t* brk = nullptr;
    int64_t brk_words = 0;
    int64_t store_cap = 0, store_fill = 0;  // in bases; batches start at multiples of 64
    uint32_t* useq() const { return useq_alloc ? useq_alloc + USEQ_PAD : nullptr; }
    // bucket array allocated with the virtual memory management API (sgc_table_create_shared): it can be mapped
    // into other processes with 2 MB pages (legacy CUDA IPC maps small pages: random peer loads run 75x slower)
    bool vmm = false;
    unsigned long long vmm_handle = 0;
    size_t vmm_bytes = 0;
    // hash-range partitioned table probed over peer memory (sgc_table_set_peers): every owner's bucket array,
    // the store replicated on this GPU (one shard per rank) and the id range of every shard
    int peer_n = 0;
    const unsigned long long* peer_filt = nullptr;  // replicated miss filter (sgc_table_set_filter), optional
    int peer_filt_bits = 0;
    int peer_rank = -1;
    const Bucket* peer_tbl[32] = {nullptr};
    uint32_t id_bounds[33] = {0};
    const uint32_t* peer_useq = nullptr;
    const uint32_t* peer_brk = nullptr;
    int64_t peer_ustride = 0, peer_bstride = 0;
    // pending label call
    bool label_pending = false;
    bool label_counted = false;  // the label kernel already counted the candidates per sequence into label_ids_off
    int64_t label_nseq = 0;
    int64_t* label_ids_off = nullptr;
    uint32_t* label_ids_out = nullptr;
    int64_t label_ids_cap = 0;
    unsigned long long label_pairs_cap = 0;
};

namespace {

int ensure(sgc_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return SGC_OK;
    if (b.p) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(SGC_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return SGC_OK;
}

inline int grid_for(int64_t n, int block, int cap) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

int set_device(sgc_ctx* c) {
    CU(cudaSetDevice(c->device));
    return SGC_OK;
}

int zero_scalars(sgc_ctx* c) {  // everything except bad_offsets, which the plan owns
    CU(cudaMemsetAsync(c->d_scal, 0, offsetof(Scalars, bad_offsets), c->stream));
    return SGC_OK;
}

int fetch_scalars(sgc_ctx* c) {
    CU(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SGC_OK;
}

// per-sequence k-mer offsets, segment offsets and the segment map.  d_kmer_off may be a
// caller buffer (nseq+1) or NULL (internal).
struct Plan {
    const int64_t* kmer_off;
    const int64_t* seg_first;
    const SegDesc* segdesc;
    const int64_t* nsegs_p;
};

int make_plan(sgc_ctx* c, const int64_t* d_off, int64_t nseq, int64_t seq_bytes, int k, int64_t* d_kmer_off,
              Plan* plan, int seg_k = SEG_K) {
    const size_t n1 = (size_t)nseq + 1;
    TRY(ensure(c, c->nk, n1 * 8));
    TRY(ensure(c, c->nseg, n1 * 8));
    TRY(ensure(c, c->seg_first, n1 * 8));
    if (!d_kmer_off) {
        TRY(ensure(c, c->kmer_off, n1 * 8));
        d_kmer_off = (int64_t*)c->kmer_off.p;
    }
    const int64_t max_segs = nseq + seq_bytes / seg_k + 1;
    TRY(ensure(c, c->seg_map, (size_t)max_segs * sizeof(SegDesc)));
    CU(cudaMemsetAsync(&c->d_scal->bad_offsets, 0, sizeof(int), c->stream));
    plan_count_kernel<<<grid_for((int64_t)n1, 256, 1 << 30), 256, 0, c->stream>>>(
        d_off, nseq, k, seq_bytes, seg_k, (int64_t*)c->nk.p, (int64_t*)c->nseg.p, c->d_scal);
    c->launches++;
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t*)c->nk.p, d_kmer_off, (int64_t)n1, c->stream));
    TRY(ensure(c, c->cub_tmp, tb));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->nk.p, d_kmer_off, (int64_t)n1, c->stream));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->nseg.p, (int64_t*)c->seg_first.p, (int64_t)n1,
                                     c->stream));
    if (nseq > 0) {
        if (max_segs > 8 * nseq)  // long sequences (queries, contigs): one thread per segment
            plan_expand_segs_kernel<<<grid_for(max_segs, 256, c->num_sms * 16), 256, 0, c->stream>>>(
                (int64_t*)c->seg_first.p, d_off, d_kmer_off, nseq, k, seg_k, (SegDesc*)c->seg_map.p);
        else
            plan_expand_kernel<<<grid_for(nseq, 256, 1 << 30), 256, 0, c->stream>>>(
                (int64_t*)c->seg_first.p, d_off, d_kmer_off, nseq, k, seg_k, (SegDesc*)c->seg_map.p);
        c->launches++;
    }
    CU(cudaGetLastError());
    plan->kmer_off = d_kmer_off;
    plan->seg_first = (int64_t*)c->seg_first.p;
    plan->segdesc = (const SegDesc*)c->seg_map.p;
    plan->nsegs_p = (int64_t*)c->seg_first.p + nseq;
    return SGC_OK;
}

template <int K, int MODE>
int launch_tile_k(sgc_ctx* c, const TileArgs& a) {
    size_t smem = sizeof(WarpTile<K, MODE == MODE_PARTITION>) * WPB;
    // per device: the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
    static thread_local int blocks_per_sm_of[64] = {0};
    int& blocks_per_sm = blocks_per_sm_of[c->device & 63];
    if (!blocks_per_sm) {
        CU(cudaFuncSetAttribute(tile_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, tile_kernel<K, MODE>, NT, smem));
        if (blocks_per_sm < 1) return fail(SGC_ERR_CUDA, "tile kernel does not fit on an SM (smem %zu)", smem);
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing) {
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, c->stream));
    }
    int ctas = blocks_per_sm;
    if (const char* lim = getenv("SGC_CTAS_PER_SM")) {  // experiment knob
        int v = atoi(lim);
        if (v >= 1 && v < ctas) ctas = v;
    }
    tile_kernel<K, MODE><<<c->num_sms * ctas, NT, smem, c->stream>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    if (c->timing) {
        CU(cudaEventRecord(e1, c->stream));
        c->timed.emplace_back(e0, e1);
    }
    return SGC_OK;
}

template <int MODE>
int launch_tile(sgc_ctx* c, const TileArgs& a) {
    switch (a.k) {
        case 21: return launch_tile_k<21, MODE>(c, a);
        case 31: return launch_tile_k<31, MODE>(c, a);
        default: return launch_tile_k<0, MODE>(c, a);
    }
}

int check_seq_args(const void* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k) {
    if (k < 1 || k > SGC_MAX_K) return fail(SGC_ERR_ARG, "k must be in 1..%d (got %d)", SGC_MAX_K, k);
    if (nseq < 0 || seq_bytes < 0) return fail(SGC_ERR_ARG, "negative size");
    if (nseq >= (int64_t)1 << 32) return fail(SGC_ERR_ARG, "at most 2^32-1 sequences per call");
    if (!d_off) return fail(SGC_ERR_ARG, "d_off is NULL");
    if (!d_seq && seq_bytes > 0) return fail(SGC_ERR_ARG, "d_seq is NULL");
    return SGC_OK;
}

TileArgs base_args(sgc_ctx* c, const uint8_t* d_seq, const int64_t* d_off, int k, const Plan& p, int32_t* d_status) {
    TileArgs a;
    memset(&a, 0, sizeof(a));
    a.seq = d_seq;
    a.off = d_off;
    a.segdesc = p.segdesc;
    a.nsegs_p = p.nsegs_p;
    a.kmer_off = p.kmer_off;
    a.status = d_status;
    a.scal = c->d_scal;
    a.k = k;
    return a;
}

void table_args(TileArgs& a, sgc_table* t) {
    a.tbl = t->buckets;
    a.nbuckets = t->nbuckets;
    a.nlines = t->nlines;
    a.home_shift = t->home_shift;
}

int device_err_check(sgc_ctx* c) {
    if (c->h_scal->bad_offsets)
        return fail(SGC_ERR_ARG, "sequence offsets are negative, decreasing or run past seq_bytes (those sequences were skipped)");
    if (c->h_scal->err & 1) return fail(SGC_ERR_CAPACITY, "k-mer table is full");
    if (c->h_scal->err & 2) return fail(SGC_ERR_ARG, "unitig id above SGC_MAX_ID");
    return SGC_OK;
}

int bad_offsets_check(sgc_ctx* c) {
    if (c->h_scal->bad_offsets)
        return fail(SGC_ERR_ARG, "sequence offsets are negative, decreasing or run past seq_bytes (those sequences were skipped)");
    return SGC_OK;
}

int bits_for(int64_t n) {  // bits needed to represent values 0..n-1
    int b = 0;
    while (((int64_t)1 << b) < n) b++;
    return b;
}

int pairs_to_csr_sorted(sgc_ctx* c, int64_t n, int64_t nseq, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
                        int64_t* h_n_labels);

// Candidate keys in pairs_a[0..n) -> CSR of distinct ascending ids per sequence.  Synchronises.
// Bucketed path (csr2_* kernels) when every sequence has few candidates, else the radix-sort path.
int pairs_to_csr(sgc_ctx* c, int64_t n, int64_t nseq, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
                 int64_t* h_n_labels, bool counted = false) {  // counted: d_ids_off[s] already holds sequence s's candidates
    if (n <= 0 || nseq <= 0 || getenv("SGC_CSR_SORT"))
        return pairs_to_csr_sorted(c, n, nseq, d_ids_off, d_ids_out, ids_cap, h_n_labels);
    const int cap = c->num_sms * 16;
    const unsigned long long* A = (const unsigned long long*)c->pairs_a.p;
    TRY(ensure(c, c->pairs_b, (size_t)n * 8));       // run_ids (uint32 each) live in the sort's second buffer
    TRY(ensure(c, c->tmp_idx, (size_t)(nseq + 1) * 8));  // start[] of the candidate runs
    uint32_t* run_ids = (uint32_t*)c->pairs_b.p;
    long long* start = (long long*)c->tmp_idx.p;
    unsigned long long* cnt = (unsigned long long*)d_ids_off;  // the caller's offsets double as the counters
    CU(cudaMemsetAsync(&c->d_scal->csr_big, 0, sizeof(unsigned int), c->stream));
    if (!counted) {
        CU(cudaMemsetAsync(d_ids_off, 0, (size_t)(nseq + 1) * 8, c->stream));
        csr2_count_kernel<<<grid_for(n, 256, cap), 256, 0, c->stream>>>(A, n, cnt);
    }
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, (long long*)cnt, start, nseq + 1, c->stream));
    TRY(ensure(c, c->cub_tmp, tb));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (long long*)cnt, start, nseq + 1, c->stream));
    csr2_scatter_kernel<<<grid_for(n, 256, cap), 256, 0, c->stream>>>(A, n, start, cnt, run_ids);
    csr2_finalize_kernel<<<grid_for(nseq, 256, cap), 256, 0, c->stream>>>(start, nseq, run_ids, cnt, c->d_scal);
    CU(cub::DeviceScan::InclusiveSum(nullptr, tb, d_ids_off, d_ids_off, nseq + 1, c->stream));
    TRY(ensure(c, c->cub_tmp, tb));
    CU(cub::DeviceScan::InclusiveSum(c->cub_tmp.p, tb, d_ids_off, d_ids_off, nseq + 1, c->stream));
    csr2_copy_kernel<<<grid_for(nseq, 256, cap), 256, 0, c->stream>>>(start, cnt, nseq, run_ids, d_ids_out, ids_cap);
    c->launches += 4;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&c->d_scal->n_unique, d_ids_off + nseq, 8, cudaMemcpyDeviceToDevice, c->stream));
    TRY(fetch_scalars(c));
    if (c->h_scal->csr_big)  // some sequence has many candidate ids (long contigs, genome-sized queries)
        return pairs_to_csr_sorted(c, n, nseq, d_ids_off, d_ids_out, ids_cap, h_n_labels);
    if (h_n_labels) *h_n_labels = c->h_scal->n_unique;
    TRY(bad_offsets_check(c));
    if (c->h_scal->n_unique > ids_cap)
        return fail(SGC_ERR_CAPACITY, "ids_cap %lld < %lld distinct (sequence, id) labels", (long long)ids_cap,
                    (long long)c->h_scal->n_unique);
    return SGC_OK;
}

// the general path: radix sort + unique of the candidate keys, any number of candidates per sequence
int pairs_to_csr_sorted(sgc_ctx* c, int64_t n, int64_t nseq, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
                        int64_t* h_n_labels) {
    CU(cudaMemsetAsync(d_ids_off, 0, (size_t)(nseq + 1) * 8, c->stream));
    CU(cudaMemsetAsync(&c->d_scal->n_unique, 0, sizeof(long long), c->stream));
    if (n > 0) {
        TRY(ensure(c, c->pairs_b, (size_t)n * 8));
        unsigned long long* A = (unsigned long long*)c->pairs_a.p;
        unsigned long long* B = (unsigned long long*)c->pairs_b.p;
        const int end_bit = std::min(64, 32 + std::max(1, bits_for(nseq)));
        size_t tb1 = 0, tb2 = 0;
        CU(cub::DeviceRadixSort::SortKeys(nullptr, tb1, A, B, n, 0, end_bit, c->stream));
        CU(cub::DeviceSelect::Unique(nullptr, tb2, B, A, &c->d_scal->n_unique, n, c->stream));
        TRY(ensure(c, c->cub_tmp, std::max(tb1, tb2)));
        CU(cub::DeviceRadixSort::SortKeys(c->cub_tmp.p, tb1, A, B, n, 0, end_bit, c->stream));
        CU(cub::DeviceSelect::Unique(c->cub_tmp.p, tb2, B, A, &c->d_scal->n_unique, n, c->stream));
        csr_count_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(
            A, &c->d_scal->n_unique, (unsigned long long*)d_ids_off, d_ids_out, ids_cap);
        c->launches++;
        size_t tb3 = 0;
        CU(cub::DeviceScan::InclusiveSum(nullptr, tb3, d_ids_off, d_ids_off, nseq + 1, c->stream));
        TRY(ensure(c, c->cub_tmp, tb3));
        CU(cub::DeviceScan::InclusiveSum(c->cub_tmp.p, tb3, d_ids_off, d_ids_off, nseq + 1, c->stream));
    }
    TRY(fetch_scalars(c));
    if (h_n_labels) *h_n_labels = c->h_scal->n_unique;
    TRY(bad_offsets_check(c));
    if (c->h_scal->n_unique > ids_cap)
        return fail(SGC_ERR_CAPACITY, "ids_cap %lld < %lld distinct (sequence, id) labels", (long long)ids_cap,
                    (long long)c->h_scal->n_unique);
    return SGC_OK;
}

}  // namespace

extern "C" {

const char* sgc_last_error(void) { return g_err; }
int sgc_abi_version(void) { return 1; }

int sgc_ctx_create(int device, void* stream, sgc_ctx** out) {
    if (!out) return fail(SGC_ERR_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SGC_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SGC_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SGC_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    if (const char* gran = getenv("SGC_L2_FETCH_GRANULARITY")) {  // experiment knob: 32 / 64 / 128
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(gran));
    }
    sgc_ctx* c = new (std::nothrow) sgc_ctx();
    if (!c) return fail(SGC_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess) { delete c; return fail(SGC_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e2)); }
        c->own_stream = true;
    }
    if (cudaMalloc(&c->d_scal, sizeof(Scalars)) != cudaSuccess ||
        cudaMallocHost(&c->h_scal, sizeof(Scalars)) != cudaSuccess) {
        sgc_ctx_destroy(c);
        return fail(SGC_ERR_NOMEM, "cannot allocate context scalars");
    }
    memset(c->h_scal, 0, sizeof(Scalars));
    *out = c;
    return SGC_OK;
}

void sgc_ctx_destroy(sgc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& ev : c->timed) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    c->timed.clear();
    DevBuf* bufs[] = {&c->nk, &c->nseg, &c->seg_first, &c->kmer_off, &c->seg_map, &c->cub_tmp,
                      &c->pairs_a, &c->pairs_b, &c->tmp_hash, &c->tmp_id, &c->tmp_idx, &c->cursor,
                      &c->units, &c->unit_off, &c->base_off, &c->unpacked,
                      &c->q_hash2, &c->q_id, &c->q_id2, &c->q_mapk, &c->q_mapc, &c->iv_a, &c->iv_b, &c->iv_c, &c->iv_d};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->d_scal) cudaFree(c->d_scal);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int sgc_ctx_sync(sgc_ctx* c) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    CU(cudaStreamSynchronize(c->stream));
    return SGC_OK;
}

void* sgc_ctx_stream(sgc_ctx* c) { return c ? (void*)c->stream : nullptr; }
int64_t sgc_ctx_launches(const sgc_ctx* c) { return c ? c->launches : 0; }

// ---------------------------------------------------------------------------- K1
}  // extern "C"
namespace {
int hash_batch_impl(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                    uint64_t* d_hash_out, uint32_t* d_aux_out, int64_t hash_cap, int64_t* d_kmer_off, int32_t* d_status,
                    int64_t* h_total_kmers) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    TRY(set_device(c));
    // capacity: total k-mers <= seq_bytes - nseq*(k-1) can only be bounded from above without
    // a sync; require the safe bound or verify after the plan.
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, d_kmer_off, &p));
    int64_t total = -1;
    if (hash_cap < seq_bytes || h_total_kmers) {
        TRY(fetch_scalars(c));
        TRY(bad_offsets_check(c));
        CU(cudaMemcpyAsync(&c->h_scal->n_unique, p.kmer_off + nseq, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        total = c->h_scal->n_unique;
        if (h_total_kmers) *h_total_kmers = total;
        if (total > hash_cap)
            return fail(SGC_ERR_CAPACITY, "hash_cap %lld < %lld k-mers", (long long)hash_cap, (long long)total);
    }
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    if (nseq == 0 || total == 0) return SGC_OK;
    if (!d_hash_out) return fail(SGC_ERR_ARG, "d_hash_out is NULL");
    TRY(zero_scalars(c));
    TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
    a.hash_out = d_hash_out;
    a.aux_out = d_aux_out;
    return launch_tile<MODE_HASH>(c, a);
}
}  // namespace
extern "C" {

int sgc_hash_batch(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                   uint64_t* d_hash_out, int64_t hash_cap, int64_t* d_kmer_off, int32_t* d_status,
                   int64_t* h_total_kmers) {
    return hash_batch_impl(c, d_seq, seq_bytes, d_off, nseq, k, d_hash_out, nullptr, hash_cap, d_kmer_off, d_status,
                           h_total_kmers);
}

int sgc_hash_positions(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                       uint64_t* d_hash_out, uint32_t* d_aux_out, int64_t hash_cap, int64_t* d_kmer_off, int32_t* d_status,
                       int64_t* h_total_kmers) {
    if (!d_aux_out) return fail(SGC_ERR_ARG, "d_aux_out is NULL");
    if (seq_bytes + 64 >= (int64_t)AUX_NONE)
        return fail(SGC_ERR_CAPACITY, "a store shard is limited to 2^30 bases (got %lld)", (long long)seq_bytes);
    return hash_batch_impl(c, d_seq, seq_bytes, d_off, nseq, k, d_hash_out, d_aux_out, hash_cap, d_kmer_off, d_status,
                           h_total_kmers);
}

// sourmash-style MinHash primitives (rows f3 / f4 of SURVEY.md 8f)
int sgc_minhash_batch(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                      uint64_t seed, uint64_t* d_hash_out, int64_t hash_cap, int64_t* d_kmer_off, uint64_t* d_seq_min,
                      int32_t* d_status, int64_t* h_total_kmers) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    TRY(set_device(c));
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, d_kmer_off, &p));
    int64_t total = -1;
    if ((d_hash_out && hash_cap < seq_bytes) || h_total_kmers) {
        TRY(fetch_scalars(c));
        TRY(bad_offsets_check(c));
        CU(cudaMemcpyAsync(&c->h_scal->n_unique, p.kmer_off + nseq, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        total = c->h_scal->n_unique;
        if (h_total_kmers) *h_total_kmers = total;
        if (d_hash_out && total > hash_cap)
            return fail(SGC_ERR_CAPACITY, "hash_cap %lld < %lld k-mers", (long long)hash_cap, (long long)total);
    }
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    if (d_seq_min && nseq) CU(cudaMemsetAsync(d_seq_min, 0xFF, (size_t)nseq * 8, c->stream));
    if (nseq == 0 || total == 0) return SGC_OK;
    TRY(zero_scalars(c));
    TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
    a.hash_out = d_hash_out;
    a.seq_min = (unsigned long long*)d_seq_min;
    a.seed = seed;
    return launch_tile<MODE_MINHASH>(c, a);
}

// ---------------------------------------------------------------------------- shareable device memory
}  // extern "C"
namespace {
struct VmmApi {
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*export_fd)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*import_fd)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};
const VmmApi* vmm_api() {
    static VmmApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    auto get = [](const char* name, void** fn) {
        cudaDriverEntryPointQueryResult q;
        return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    api.ok = get("cuMemCreate", (void**)&api.create) && get("cuMemRelease", (void**)&api.release) &&
             get("cuMemAddressReserve", (void**)&api.reserve) && get("cuMemAddressFree", (void**)&api.addr_free) &&
             get("cuMemMap", (void**)&api.map) && get("cuMemUnmap", (void**)&api.unmap) &&
             get("cuMemSetAccess", (void**)&api.set_access) &&
             get("cuMemExportToShareableHandle", (void**)&api.export_fd) &&
             get("cuMemImportFromShareableHandle", (void**)&api.import_fd) &&
             get("cuMemGetAllocationGranularity", (void**)&api.granularity);
    return api.ok ? &api : nullptr;
}
#define CD(x)                                                                                  \
    do {                                                                                       \
        CUresult r_ = (x);                                                                     \
        if (r_ != CUDA_SUCCESS) return fail(SGC_ERR_CUDA, "%s failed: CUresult %d", #x, (int)r_); \
    } while (0)

CUmemAllocationProp vmm_prop(int device) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}
int vmm_map(const VmmApi* v, int device, CUmemGenericAllocationHandle h, size_t bytes, size_t gran, void** out) {
    CUdeviceptr p = 0;
    CD(v->reserve(&p, bytes, gran, 0, 0));
    CD(v->map(p, bytes, 0, h, 0));
    CUmemAccessDesc ad;
    memset(&ad, 0, sizeof(ad));
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ad.location.id = device;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CD(v->set_access(p, bytes, &ad, 1));
    *out = (void*)p;
    return SGC_OK;
}
// physical memory on the context's GPU, mapped here, exportable as a POSIX fd
int vmm_alloc(sgc_ctx* c, size_t want, void** ptr, unsigned long long* handle, size_t* bytes) {
    const VmmApi* v = vmm_api();
    if (!v) return fail(SGC_ERR_CUDA, "the CUDA driver does not offer the virtual memory management API");
    CUmemAllocationProp prop = vmm_prop(c->device);
    size_t gran = 0;
    CD(v->granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (gran < ((size_t)1 << 21)) gran = (size_t)1 << 21;
    const size_t sz = (want + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    CUresult r = v->create(&h, sz, &prop, 0);
    if (r != CUDA_SUCCESS) return fail(SGC_ERR_NOMEM, "cuMemCreate(%zu) failed: CUresult %d", sz, (int)r);
    int rc = vmm_map(v, c->device, h, sz, gran, ptr);
    if (rc != SGC_OK) {
        v->release(h);
        return rc;
    }
    *handle = (unsigned long long)h;
    *bytes = sz;
    return SGC_OK;
}
void vmm_free(void* ptr, unsigned long long handle, size_t bytes) {
    const VmmApi* v = vmm_api();
    if (!v || !ptr) return;
    v->unmap((CUdeviceptr)ptr, bytes);
    v->addr_free((CUdeviceptr)ptr, bytes);
    v->release((CUmemGenericAllocationHandle)handle);
}
}  // namespace
extern "C" {

// ---------------------------------------------------------------------------- K2
}  // extern "C"
namespace {
int table_create(sgc_ctx* c, int64_t max_entries, float load, bool shared, sgc_table** out) {
    if (!c || !out) return fail(SGC_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (max_entries < 0) return fail(SGC_ERR_ARG, "max_entries < 0");
    TRY(set_device(c));
    // 32-byte buckets of two 16-byte slots, four per 128-byte line; `load` = expected hashes per
    // bucket (default 0.6: ~2.3 % of the home buckets overflow; at 0.75 it was 4 %, and since one lane that
    // leaves its home bucket drags its whole warp through the walk, the label kernel ran 3 % slower)
    if (!(load > 0.0f)) {
        const char* env = getenv("SGC_TABLE_LOAD");  // experiment knob
        load = env ? (float)atof(env) : 0.6f;
        if (!(load > 0.0f)) load = 0.6f;
    }
    if (load > 1.6f) return fail(SGC_ERR_ARG, "table load %.2f too high (max 1.6 hashes per 2-slot bucket)", load);
    int64_t nb = (int64_t)((double)std::max<int64_t>(max_entries, 64) / load) + 4;
    nb = (nb + 3) & ~(int64_t)3;
    if (nb >= ((int64_t)1 << 32)) return fail(SGC_ERR_CAPACITY, "table of %lld buckets exceeds 2^32-1", (long long)nb);
    sgc_table* t = new (std::nothrow) sgc_table();
    if (!t) return fail(SGC_ERR_NOMEM, "out of host memory");
    t->ctx = c;
    t->nbuckets = (uint32_t)nb;
    t->nlines = (uint32_t)(nb / 4);
    t->max_entries = max_entries;
    if (shared) {
        void* p = nullptr;
        int rc = vmm_alloc(c, (size_t)nb * sizeof(Bucket), &p, &t->vmm_handle, &t->vmm_bytes);
        if (rc != SGC_OK) {
            delete t;
            return rc;
        }
        t->buckets = (Bucket*)p;
        t->vmm = true;
    } else {
        cudaError_t e = cudaMalloc(&t->buckets, (size_t)nb * sizeof(Bucket));
        if (e != cudaSuccess) {
            delete t;
            return fail(SGC_ERR_NOMEM, "cudaMalloc(%zu) for the k-mer table failed: %s", (size_t)nb * sizeof(Bucket),
                        cudaGetErrorString(e));
        }
    }
    table_clear_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(t->buckets, t->nbuckets);
    c->launches++;
    CU(cudaGetLastError());
    *out = t;
    return SGC_OK;
}
}  // namespace
extern "C" {

int sgc_table_create(sgc_ctx* c, int64_t max_entries, float load, sgc_table** out) {
    return table_create(c, max_entries, load, false, out);
}
int sgc_table_create_shared(sgc_ctx* c, int64_t max_entries, float load, sgc_table** out) {
    return table_create(c, max_entries, load, true, out);
}

void sgc_table_free(sgc_table* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    if (t->buckets && t->vmm) vmm_free(t->buckets, t->vmm_handle, t->vmm_bytes);
    else if (t->buckets) cudaFree(t->buckets);
    if (t->useq_alloc) cudaFree(t->useq_alloc);
    if (t->brk) cudaFree(t->brk);
    delete t;
}

int64_t sgc_table_bytes(const sgc_table* t) { return t ? (int64_t)t->nbuckets * (int64_t)sizeof(Bucket) : 0; }

int sgc_table_enable_side(sgc_table* t) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    t->store_on = true;  // allocated (and grown) by sgc_table_insert_sequences
    return SGC_OK;
}

int64_t sgc_table_side_bytes(const sgc_table* t) {
    return t && t->useq_alloc ? (t->store_cap / 16 + 2 * USEQ_PAD) * 4 + t->brk_words * 4 : 0;
}

int sgc_table_set_home_shift(sgc_table* t, int shift) {
    if (!t || shift < 0 || shift > 16) return fail(SGC_ERR_ARG, "bad home shift");
    t->home_shift = shift;
    return SGC_OK;
}

namespace {
// room for `bases` more positions in the table's sequence store (grows by reallocation; the break bits of
// the new part start out set)
int store_reserve(sgc_table* t, int64_t bases) {
    sgc_ctx* c = t->ctx;
    const int64_t need = t->store_fill + ((bases + 63) & ~(int64_t)63);
    if (need + 64 >= (int64_t)AUX_NONE)
        return fail(SGC_ERR_CAPACITY, "unitig sequence store limited to 2^30 bases per table (needs %lld)", (long long)need);
    if (need <= t->store_cap) return SGC_OK;
    int64_t cap = std::max<int64_t>(need, t->store_cap + t->store_cap / 2);
    cap = std::min<int64_t>((cap + 1023) & ~(int64_t)1023, ((int64_t)AUX_NONE - 64) & ~(int64_t)63);
    const int64_t uwords = cap / 16 + 2 * USEQ_PAD;
    const int64_t bwords = ((cap + BRK_PAD + 31) >> 5) + 2;
    uint32_t *nu = nullptr, *nb = nullptr;
    CU(cudaStreamSynchronize(c->stream));
    cudaError_t e = cudaMalloc(&nu, (size_t)uwords * 4);
    if (e == cudaSuccess) e = cudaMalloc(&nb, (size_t)bwords * 4);
    if (e != cudaSuccess) {
        if (nu) cudaFree(nu);
        return fail(SGC_ERR_NOMEM, "cudaMalloc(%zu) for the unitig sequence store failed: %s", (size_t)uwords * 4,
                    cudaGetErrorString(e));
    }
    CU(cudaMemsetAsync(nu, 0, (size_t)uwords * 4, c->stream));
    CU(cudaMemsetAsync(nb, 0xFF, (size_t)bwords * 4, c->stream));
    if (t->useq_alloc) {
        CU(cudaMemcpyAsync(nu, t->useq_alloc, (size_t)(t->store_fill / 16 + USEQ_PAD) * 4, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(nb, t->brk, (size_t)((t->store_fill + BRK_PAD) >> 5) * 4, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(t->useq_alloc);
        cudaFree(t->brk);
    }
    t->useq_alloc = nu;
    t->brk = nb;
    t->brk_words = bwords;
    t->store_cap = cap;
    return SGC_OK;
}
}  // namespace

int sgc_table_insert_pairs(sgc_table* t, const uint64_t* d_hash, const uint32_t* d_id, int64_t n) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (n < 0) return fail(SGC_ERR_ARG, "n < 0");
    if (n == 0) return SGC_OK;
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    TRY(zero_scalars(c));
    insert_pairs_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(t->buckets, t->nbuckets, t->nlines,
                                                                               t->home_shift, d_hash, d_id, n, c->d_scal,
                                                                               t->brk);
    c->launches++;
    CU(cudaGetLastError());
    TRY(fetch_scalars(c));
    return device_err_check(c);
}

int sgc_table_insert_sequences(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                               int64_t nseq, int k, const uint32_t* d_ids, uint32_t id_base, int32_t* d_status) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    TRY(set_device(c));
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    if (nseq == 0) return SGC_OK;
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, nullptr, &p));
    if (t->store_on) {
        // the batch's bases, 2 bits each, at the next 64-base boundary of the store; break bits cleared where a
        // k-mer follows a k-mer of the same sequence (the insert kernel sets more: multicounts, repeats)
        TRY(store_reserve(t, seq_bytes));
        const int64_t nw = (seq_bytes + 15) / 16, nbw = (seq_bytes + 31) / 32;
        useq_pack_kernel<<<grid_for(nw, 256, c->num_sms * 16), 256, 0, c->stream>>>(d_seq, seq_bytes, t->useq() + t->store_fill / 16, nw);
        useq_brk_kernel<<<grid_for(nbw, 256, c->num_sms * 16), 256, 0, c->stream>>>(d_off, nseq, k, t->brk + ((t->store_fill + BRK_PAD) >> 5), nbw);
        c->launches += 2;
        CU(cudaGetLastError());
    }
    TRY(zero_scalars(c));
    TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
    table_args(a, t);
    a.seq_ids = d_ids;
    a.id_base = id_base;
    a.brk = t->store_on ? t->brk : nullptr;
    a.upos_base = (unsigned long long)t->store_fill;
    TRY(launch_tile<MODE_INSERT>(c, a));
    TRY(fetch_scalars(c));
    if (t->store_on) t->store_fill += (seq_bytes + 63) & ~(int64_t)63;
    return device_err_check(c);
}

int sgc_table_stats(sgc_table* t, int64_t* h_live, int64_t* h_multi, uint32_t* h_max_id) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    TRY(zero_scalars(c));
    table_stats_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(t->buckets, t->nbuckets, c->d_scal);
    c->launches++;
    CU(cudaGetLastError());
    TRY(fetch_scalars(c));
    if (h_live) *h_live = (int64_t)c->h_scal->n_live;
    if (h_multi) *h_multi = (int64_t)c->h_scal->n_multi;
    if (h_max_id) *h_max_id = c->h_scal->max_id;
    return SGC_OK;
}

int sgc_table_export(sgc_table* t, uint64_t* d_hash_out, uint32_t* d_id_out, int64_t cap, int64_t* h_n_out) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    if (cap < 0) return fail(SGC_ERR_ARG, "cap < 0");
    TRY(ensure(c, c->tmp_hash, (size_t)std::max<int64_t>(cap, 1) * 8));
    TRY(ensure(c, c->tmp_id, (size_t)std::max<int64_t>(cap, 1) * 4));
    TRY(zero_scalars(c));
    table_export_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(t->buckets, t->nbuckets, (uint64_t*)c->tmp_hash.p,
                                                              (uint32_t*)c->tmp_id.p, (unsigned long long)cap, c->d_scal);
    c->launches++;
    CU(cudaGetLastError());
    TRY(fetch_scalars(c));
    int64_t n = (int64_t)c->h_scal->n_export;
    if (h_n_out) *h_n_out = n;
    if (n > cap) return fail(SGC_ERR_CAPACITY, "export cap %lld < %lld live hashes", (long long)cap, (long long)n);
    if (n == 0) return SGC_OK;
    size_t tb = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, (uint64_t*)c->tmp_hash.p, d_hash_out, (uint32_t*)c->tmp_id.p,
                                       d_id_out, n, 0, 64, c->stream));
    TRY(ensure(c, c->cub_tmp, tb));
    CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, (uint64_t*)c->tmp_hash.p, d_hash_out, (uint32_t*)c->tmp_id.p,
                                       d_id_out, n, 0, 64, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SGC_OK;
}

// ---------------------------------------------------------------------------- K3
int sgc_lookup(sgc_table* t, const uint64_t* d_hash, int64_t n, uint32_t* d_id_out) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (n < 0) return fail(SGC_ERR_ARG, "n < 0");
    if (n == 0) return SGC_OK;
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    lookup_kernel<<<grid_for((n + 1) / 2, 256, c->num_sms * 8), 256, 0, c->stream>>>(
        t->buckets, t->nbuckets, t->nlines, t->home_shift, d_hash, n, d_id_out);
    c->launches++;
    CU(cudaGetLastError());
    return SGC_OK;
}

// ---------------------------------------------------------------------------- K5
int sgc_count_matches(sgc_table* t, const uint64_t* d_hash, int64_t n, int distinct, uint32_t* d_count_by_id,
                      int64_t n_ids, int64_t* h_n_miss, int64_t* h_n_distinct) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (n < 0 || n_ids < 0) return fail(SGC_ERR_ARG, "negative size");
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    if (n_ids) CU(cudaMemsetAsync(d_count_by_id, 0, (size_t)n_ids * 4, c->stream));
    TRY(zero_scalars(c));
    if (n > 0) {
        const uint64_t* src = d_hash;
        if (distinct) {
            TRY(ensure(c, c->tmp_hash, (size_t)n * 8));
            size_t tb = 0;
            CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_hash, (uint64_t*)c->tmp_hash.p, n, 0, 64, c->stream));
            TRY(ensure(c, c->cub_tmp, tb));
            CU(cub::DeviceRadixSort::SortKeys(c->cub_tmp.p, tb, d_hash, (uint64_t*)c->tmp_hash.p, n, 0, 64, c->stream));
            src = (const uint64_t*)c->tmp_hash.p;
        }
        count_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(
            t->buckets, t->nbuckets, t->nlines, t->home_shift, src, n, distinct, d_count_by_id, c->d_scal);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (h_n_miss || h_n_distinct) {
        TRY(fetch_scalars(c));
        if (h_n_miss) *h_n_miss = (int64_t)c->h_scal->n_miss;
        if (h_n_distinct) *h_n_distinct = (int64_t)c->h_scal->n_distinct;
    }
    return SGC_OK;
}

// Batched Query.__init__ + count_cdbg_matches (search/query_by_sequence.py:153-189) over MANY queries in one
// call: hash every record of every query (one launch), one stable radix sort of all hashes on their top 32
// bits with the query index as the value, then one pass that drops each query's duplicate hashes, looks the
// rest up (sorted hashes walk the table in bucket order) and counts them per (query, unitig).
}  // extern "C"

namespace {

struct MaxU64 {
    __host__ __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; }
};
struct SumU32 {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; }
};

bool use_side(const sgc_table* t);
void labelseq_args(LabelSeqArgs& a, sgc_table* t, const uint8_t* seq, const Plan& p, int k, int32_t* d_status,
                   unsigned long long pairs_cap);

int launch_labelseq_query(sgc_ctx* c, const LabelSeqArgs& a);  // defined next to launch_labelseq

// The batched query through the unitig sequence store (sgc_labelseq.cuh, QUERY): no k-mer hash ever reaches HBM.
// The label kernel's runs give, per query, intervals of store positions (one 16-byte record per settled run instead
// of one hash per k-mer) and the hashes of the k-mers the table does not hold; the number of distinct matching
// hashes of a query on a unitig is the size of the union of its intervals there.
//   intervals: sort by (query, first position) -> running maximum of the ends -> uncovered length of every interval
//              -> reduce by (query, unitig) -> sort the pairs -> merge
//   misses:    sort by a 32-bit key (query, top bits of the hash) -> first occurrences per (hash, query) inside the runs
// *fallback = true (nothing written): some hit has no store position (a table also filled hash by hash).
int query_count_batch_store(sgc_table* t, const uint8_t* d_seq, const int64_t* d_off, int64_t nseq, const Plan& p, int64_t n,
                            const int32_t* d_seq_query, int n_queries, int k, uint64_t* d_keys_out, uint32_t* d_counts_out,
                            int64_t cap, int64_t* d_q_distinct, int32_t* d_status, int64_t* h_n_pairs, int64_t* h_n_hits,
                            bool* fallback) {
    sgc_ctx* c = t->ctx;
    *fallback = false;
    const int qbits = std::max(1, bits_for(std::max(n_queries, 1)));
    if (qbits > 31) return fail(SGC_ERR_ARG, "at most 2^31 queries per batch");
    // the missing k-mers' sort key: query | top bits of the hash, 24 bits (3 radix passes) when that leaves the hash
    // at least 12 bits, else 32; the runs of equal keys only have to stay short
    const int mkbits = qbits <= 12 ? 24 : 32;
    // the intervals' sort key: query << pbits | store position (pbits = bits of the store's size)
    int pbits = 1;
    while (pbits < 32 && ((int64_t)1 << pbits) <= t->store_fill + 64) pbits++;
    // every warp of the grid may leave one chunk partly used (it is filled with copies, sgc_labelseq.cuh)
    const unsigned long long slack = (unsigned long long)c->num_sms * SGC_LS_MIN_CTAS * WPBS * QCHUNK + 4096;
    unsigned long long iv_cap = (unsigned long long)(n / 4) + slack, ms_cap = (unsigned long long)(n / 4) + slack;
    if (getenv("SGC_QUERY_SMALL_CAPS")) iv_cap = ms_cap = 2048;  // test knob: force the overflow + retry path
    int64_t n_iv = 0, n_ms = 0;
    for (int attempt = 0;; attempt++) {
        TRY(ensure(c, c->iv_a, (size_t)iv_cap * 8));
        TRY(ensure(c, c->iv_b, (size_t)iv_cap * 8));
        TRY(ensure(c, c->iv_c, (size_t)iv_cap * 8));
        TRY(ensure(c, c->iv_d, (size_t)iv_cap * 8));
        TRY(ensure(c, c->q_id, (size_t)iv_cap * 4));
        TRY(ensure(c, c->q_id2, (size_t)iv_cap * 4));
        TRY(ensure(c, c->tmp_hash, (size_t)ms_cap * 8));
        TRY(ensure(c, c->q_hash2, (size_t)ms_cap * 8));
        TRY(ensure(c, c->q_mapc, (size_t)ms_cap * 4));
        TRY(ensure(c, c->tmp_id, (size_t)ms_cap * 4));
        TRY(zero_scalars(c));
        if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
        LabelSeqArgs a;
        labelseq_args(a, t, d_seq, p, k, d_status, 0);
        a.pairs = nullptr;
        a.seq_query = d_seq_query;
        a.iv_key = (unsigned long long*)c->iv_a.p;
        a.iv_val = (unsigned long long*)c->iv_b.p;
        a.iv_cap = iv_cap;
        a.miss_hash = (unsigned long long*)c->tmp_hash.p;
        a.miss_q = (uint32_t*)c->q_mapc.p;
        a.miss_cap = ms_cap;
        a.miss_hbits = mkbits - qbits;
        a.iv_pbits = pbits;
        TRY(launch_labelseq_query(c, a));
        TRY(fetch_scalars(c));
        if (c->h_scal->err & ERR_QUERY_NOPOS) {
            *fallback = true;
            return SGC_OK;
        }
        n_iv = (int64_t)c->h_scal->n_export;
        n_ms = (int64_t)c->h_scal->n_live;
        if (!(c->h_scal->err & ERR_QUERY_CAP)) break;
        if (attempt == 1) return fail(SGC_ERR_CAPACITY, "query interval buffers overflowed twice");
        // at most one record per k-mer, a tenth of every chunk abandoned at worst (a trip writes < 96 records)
        iv_cap = ms_cap = (unsigned long long)n + (unsigned long long)n / 8 + slack;
    }
    CU(cudaMemsetAsync(&c->d_scal->n_hits, 0, 8, c->stream));
    int64_t np = 0;
    size_t tb = 0;
    if (n_iv > 0) {
        unsigned long long *A = (unsigned long long*)c->iv_a.p, *B = (unsigned long long*)c->iv_b.p,
                           *C = (unsigned long long*)c->iv_c.p, *D = (unsigned long long*)c->iv_d.p;
        uint32_t *cnt = (uint32_t*)c->q_id.p, *cnt2 = (uint32_t*)c->q_id2.p;
        const int end_bit = std::min(64, 32 + qbits);
        // (A, B) -> sorted (C, D)
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, A, C, B, D, n_iv, 0, pbits + qbits, c->stream));
        TRY(ensure(c, c->cub_tmp, tb));
        CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, A, C, B, D, n_iv, 0, pbits + qbits, c->stream));
        iv_end_kernel<<<grid_for(n_iv, 256, c->num_sms * 8), 256, 0, c->stream>>>(C, D, n_iv, A);
        CU(cub::DeviceScan::ExclusiveScan(nullptr, tb, A, B, MaxU64(), 0ULL, n_iv, c->stream));
        TRY(ensure(c, c->cub_tmp, tb));
        CU(cub::DeviceScan::ExclusiveScan(c->cub_tmp.p, tb, A, B, MaxU64(), 0ULL, n_iv, c->stream));
        iv_union_kernel<<<grid_for(n_iv, 256, c->num_sms * 8), 256, 0, c->stream>>>(C, D, B, n_iv, pbits, A, cnt);
        c->launches += 2;
        // runs of equal (query, unitig): A / cnt -> C / cnt2
        long long* d_runs = &c->d_scal->n_unique;
        CU(cub::DeviceReduce::ReduceByKey(nullptr, tb, A, C, cnt, cnt2, d_runs, SumU32(), n_iv, c->stream));
        TRY(ensure(c, c->cub_tmp, tb));
        CU(cub::DeviceReduce::ReduceByKey(c->cub_tmp.p, tb, A, C, cnt, cnt2, d_runs, SumU32(), n_iv, c->stream));
        TRY(fetch_scalars(c));
        const int64_t nruns = c->h_scal->n_unique;
        // a unitig may occupy several ranges of the store (the same id given to several sequences): sort + merge
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, C, A, cnt2, cnt, nruns, 0, end_bit, c->stream));
        TRY(ensure(c, c->cub_tmp, tb));
        CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, C, A, cnt2, cnt, nruns, 0, end_bit, c->stream));
        CU(cub::DeviceReduce::ReduceByKey(nullptr, tb, A, C, cnt, cnt2, d_runs, SumU32(), nruns, c->stream));
        TRY(ensure(c, c->cub_tmp, tb));
        CU(cub::DeviceReduce::ReduceByKey(c->cub_tmp.p, tb, A, C, cnt, cnt2, d_runs, SumU32(), nruns, c->stream));
        TRY(fetch_scalars(c));
        np = c->h_scal->n_unique;
        *h_n_pairs = np;
        if (np > cap) return fail(SGC_ERR_CAPACITY, "cap %lld < %lld (query, unitig) pairs", (long long)cap, (long long)np);
        CU(cudaMemcpyAsync(d_keys_out, C, (size_t)np * 8, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(d_counts_out, cnt2, (size_t)np * 4, cudaMemcpyDeviceToDevice, c->stream));
        pairs_tally_kernel<<<grid_for(np, 256, c->num_sms * 4), 256, n_queries <= QHIST_MAX ? (size_t)n_queries * 4 : 0, c->stream>>>(
            (const unsigned long long*)d_keys_out, d_counts_out, np, (unsigned long long*)d_q_distinct, n_queries, c->d_scal);
        c->launches++;
    }
    if (n_ms > 0) {
        uint64_t *H = (uint64_t*)c->tmp_hash.p, *H2 = (uint64_t*)c->q_hash2.p;
        uint32_t *Q = (uint32_t*)c->q_mapc.p, *Q2 = (uint32_t*)c->tmp_id.p;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, Q, Q2, H, H2, n_ms, 0, mkbits, c->stream));
        TRY(ensure(c, c->cub_tmp, tb));
        CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, Q, Q2, H, H2, n_ms, 0, mkbits, c->stream));
        const size_t hist_bytes = n_queries <= QHIST_MAX ? (size_t)n_queries * 4 : 0;
        miss_distinct_kernel<<<grid_for(n_ms, 256, c->num_sms * 8), 256, hist_bytes, c->stream>>>(
            Q2, H2, n_ms, mkbits - qbits, (unsigned long long*)d_q_distinct, n_queries);
        c->launches++;
    }
    CU(cudaGetLastError());
    TRY(fetch_scalars(c));
    if (h_n_hits) *h_n_hits = (int64_t)c->h_scal->n_hits;
    return SGC_OK;
}

}  // namespace

extern "C" {

int sgc_query_count_batch(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq,
                          const int32_t* d_seq_query, int n_queries, int k, uint64_t* d_keys_out, uint32_t* d_counts_out,
                          int64_t cap, int64_t* d_q_distinct, int32_t* d_status, int64_t* h_n_pairs, int64_t* h_n_kmers,
                          int64_t* h_n_hits) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    if (n_queries < 0 || cap < 0 || (nseq > 0 && !d_seq_query) || !d_q_distinct || !h_n_pairs)
        return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    if (n_queries) CU(cudaMemsetAsync(d_q_distinct, 0, (size_t)n_queries * 8, c->stream));
    *h_n_pairs = 0;
    if (h_n_kmers) *h_n_kmers = 0;
    if (h_n_hits) *h_n_hits = 0;
    if (nseq == 0) return SGC_OK;
    const bool store = use_side(t) && t->peer_n == 0 && !getenv("SGC_QUERY_SORT");
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, nullptr, &p, store ? SEG_KS : SEG_K));
    CU(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&c->h_scal->n_unique, p.kmer_off + nseq, 8, cudaMemcpyDeviceToHost, c->stream));  // (after the scalars)
    CU(cudaStreamSynchronize(c->stream));
    TRY(bad_offsets_check(c));
    const int64_t n = c->h_scal->n_unique;
    if (h_n_kmers) *h_n_kmers = n;
    if (n == 0) return SGC_OK;
    if (n >= ((int64_t)1 << 32)) return fail(SGC_ERR_ARG, "at most 2^32-1 k-mers per query batch");
    if (store) {
        // the table carries the unitig sequence store: intervals of store positions instead of sorted hashes
        bool fallback = false;
        TRY(query_count_batch_store(t, d_seq, d_off, nseq, p, n, d_seq_query, n_queries, k, d_keys_out, d_counts_out, cap,
                                    d_q_distinct, d_status, h_n_pairs, h_n_hits, &fallback));
        if (!fallback) return SGC_OK;
        if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
        CU(cudaMemsetAsync(d_q_distinct, 0, (size_t)n_queries * 8, c->stream));
        TRY(make_plan(c, d_off, nseq, seq_bytes, k, nullptr, &p));
    }
    TRY(ensure(c, c->tmp_hash, (size_t)n * 8));
    TRY(ensure(c, c->q_hash2, (size_t)n * 8));
    TRY(ensure(c, c->q_id, (size_t)n * 4));
    TRY(ensure(c, c->q_id2, (size_t)n * 4));
    TRY(zero_scalars(c));
    {
        TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
        a.hash_out = (uint64_t*)c->tmp_hash.p;
        TRY(launch_tile<MODE_HASH>(c, a));
    }
    kmer_query_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(p.kmer_off, d_seq_query, nseq, (uint32_t*)c->q_id.p);
    c->launches++;
    size_t tb = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, (uint64_t*)c->tmp_hash.p, (uint64_t*)c->q_hash2.p, (uint32_t*)c->q_id.p,
                                       (uint32_t*)c->q_id2.p, n, 32, 64, c->stream));
    TRY(ensure(c, c->cub_tmp, tb));
    CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, (uint64_t*)c->tmp_hash.p, (uint64_t*)c->q_hash2.p, (uint32_t*)c->q_id.p,
                                       (uint32_t*)c->q_id2.p, n, 32, 64, c->stream));
    // scratch map (query, unitig) -> count.  A query touches far fewer unitigs than it has k-mers (a unitig holds
    // tens of k-mers), so the map starts at n/32 slots -- small enough to stay in the 126 MB L2, where its atomics
    // are cheap -- and the pass is redone with 4x the slots if it fills beyond 5/8 (worst case: n slots)
    uint64_t slots = 1u << 16;
    while ((int64_t)slots < n / 32 + 1024) slots <<= 1;
    for (int attempt = 0; attempt < 5; attempt++, slots <<= 2) {
        TRY(ensure(c, c->q_mapk, (size_t)slots * 8));
        TRY(ensure(c, c->q_mapc, (size_t)slots * 4));
        qmap_clear_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>((unsigned long long*)c->q_mapk.p, (uint32_t*)c->q_mapc.p, slots);
        CU(cudaMemsetAsync(d_q_distinct, 0, (size_t)n_queries * 8, c->stream));
        TRY(zero_scalars(c));
        const size_t hist_bytes = n_queries <= QHIST_MAX ? (size_t)n_queries * 4 : 0;
        query_count_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, hist_bytes, c->stream>>>(
            t->buckets, t->nbuckets, t->nlines, t->home_shift, (const uint64_t*)c->q_hash2.p, (const uint32_t*)c->q_id2.p, n,
            (unsigned long long*)d_q_distinct, n_queries, (unsigned long long*)c->q_mapk.p, (uint32_t*)c->q_mapc.p, slots - 1,
            c->d_scal);
        c->launches += 2;
        CU(cudaGetLastError());
        TRY(fetch_scalars(c));
        if (!(c->h_scal->err & 1)) break;
        if (attempt == 4) return fail(SGC_ERR_CAPACITY, "query count map overflow");
    }
    if (h_n_hits) *h_n_hits = (int64_t)c->h_scal->n_hits;
    // live map entries -> (key, count) sorted by (query, unitig)
    TRY(ensure(c, c->tmp_hash, (size_t)std::max<int64_t>(cap, 1) * 8));  // the hashes are no longer needed
    TRY(ensure(c, c->tmp_id, (size_t)std::max<int64_t>(cap, 1) * 4));
    qmap_compact_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>((const unsigned long long*)c->q_mapk.p, (const uint32_t*)c->q_mapc.p,
                                                              slots, (unsigned long long*)c->tmp_hash.p, (uint32_t*)c->tmp_id.p,
                                                              (unsigned long long)cap, c->d_scal);
    c->launches++;
    CU(cudaGetLastError());
    TRY(fetch_scalars(c));
    const int64_t np = (int64_t)c->h_scal->n_export;
    *h_n_pairs = np;
    if (np > cap) return fail(SGC_ERR_CAPACITY, "cap %lld < %lld (query, unitig) pairs", (long long)cap, (long long)np);
    if (np > 0) {
        const int end_bit = std::min(64, 32 + std::max(1, bits_for(std::max(n_queries, 1))));
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, (uint64_t*)c->tmp_hash.p, d_keys_out, (uint32_t*)c->tmp_id.p, d_counts_out,
                                           np, 0, end_bit, c->stream));
        TRY(ensure(c, c->cub_tmp, tb));
        CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, (uint64_t*)c->tmp_hash.p, d_keys_out, (uint32_t*)c->tmp_id.p,
                                           d_counts_out, np, 0, end_bit, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return SGC_OK;
}

int sgc_lookup_sequences(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq,
                         int k, uint32_t* d_kmer_ids, int64_t ids_cap, int64_t* d_kmer_off, uint32_t* d_count_by_id,
                         int64_t n_ids, int32_t* d_status, int64_t* h_n_kmers, int64_t* h_n_hits) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    TRY(set_device(c));
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    if (d_count_by_id && n_ids) CU(cudaMemsetAsync(d_count_by_id, 0, (size_t)n_ids * 4, c->stream));
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, d_kmer_off, &p));
    if (d_kmer_ids && ids_cap < seq_bytes) {
        CU(cudaMemcpyAsync(&c->h_scal->n_unique, p.kmer_off + nseq, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (c->h_scal->n_unique > ids_cap)
            return fail(SGC_ERR_CAPACITY, "ids_cap %lld < %lld k-mers", (long long)ids_cap, (long long)c->h_scal->n_unique);
    }
    TRY(zero_scalars(c));
    if (nseq > 0) {
        TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
        table_args(a, t);
        a.kmer_ids = d_kmer_ids;
        a.count_by_id = d_count_by_id;
        TRY(launch_tile<MODE_LOOKUP>(c, a));
    }
    if (h_n_kmers || h_n_hits) {
        TRY(fetch_scalars(c));
        if (h_n_hits) *h_n_hits = (int64_t)c->h_scal->n_hits;
        if (h_n_kmers) *h_n_kmers = (int64_t)(c->h_scal->n_hits + c->h_scal->n_miss);
        TRY(bad_offsets_check(c));
    }
    return SGC_OK;
}

// ---------------------------------------------------------------------------- K4
}  // extern "C"

namespace {

template <int K, bool PACKED, bool PEER, bool QUERY = false>
int launch_labelseq_k(sgc_ctx* c, const LabelSeqArgs& a) {
    const size_t smem = sizeof(WarpTileS<K, PACKED>) * WPBS;
    static thread_local int blocks_per_sm_of[64] = {0};
    int& blocks_per_sm = blocks_per_sm_of[c->device & 63];
    if (!blocks_per_sm) {
        CU(cudaFuncSetAttribute(labelseq_kernel<K, PACKED, PEER, QUERY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, labelseq_kernel<K, PACKED, PEER, QUERY>, NTS, smem));
        if (blocks_per_sm < 1) return fail(SGC_ERR_CUDA, "label kernel does not fit on an SM (smem %zu)", smem);
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing) {
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, c->stream));
    }
    int ctas = blocks_per_sm;
    if (const char* lim = getenv("SGC_CTAS_PER_SM")) {  // experiment knob
        int v = atoi(lim);
        if (v >= 1 && v < ctas) ctas = v;
    }
    labelseq_kernel<K, PACKED, PEER, QUERY><<<c->num_sms * ctas, NTS, smem, c->stream>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    if (c->timing) {
        CU(cudaEventRecord(e1, c->stream));
        c->timed.emplace_back(e0, e1);
    }
    return SGC_OK;
}

int launch_labelseq_query(sgc_ctx* c, const LabelSeqArgs& a) {
    switch (a.k) {
        case 21: return launch_labelseq_k<21, false, false, true>(c, a);
        case 31: return launch_labelseq_k<31, false, false, true>(c, a);
        default: return launch_labelseq_k<0, false, false, true>(c, a);
    }
}

template <bool PACKED>
int launch_labelseq(sgc_ctx* c, const LabelSeqArgs& a) {
    if (a.n_shards > 1) {  // the table is partitioned: probes go to the owner GPU's buckets over NVLink
        switch (a.k) {
            case 21: return launch_labelseq_k<21, PACKED, true>(c, a);
            case 31: return launch_labelseq_k<31, PACKED, true>(c, a);
            default: return launch_labelseq_k<0, PACKED, true>(c, a);
        }
    }
    switch (a.k) {
        case 21: return launch_labelseq_k<21, PACKED, false>(c, a);
        case 31: return launch_labelseq_k<31, PACKED, false>(c, a);
        default: return launch_labelseq_k<0, PACKED, false>(c, a);
    }
}

bool use_side(const sgc_table* t) {
    return ((t->useq_alloc && t->store_fill > 0) || t->peer_n > 0) && !getenv("SGC_NO_SIDE");
}

int label_common_prologue(sgc_table* t, int64_t nseq, int64_t ids_cap, unsigned long long* pairs_cap_out) {
    sgc_ctx* c = t->ctx;
    // candidate (sequence, id) keys: a few per read in practice; start from ids_cap and a floor
    unsigned long long pairs_cap = (unsigned long long)std::max<int64_t>(std::max<int64_t>(ids_cap * 2, nseq * 4), 4096);
    TRY(ensure(c, c->pairs_a, (size_t)pairs_cap * 8));
    *pairs_cap_out = c->pairs_a.cap / 8;
    return zero_scalars(c);
}

void labelseq_args(LabelSeqArgs& a, sgc_table* t, const uint8_t* seq, const Plan& p, int k, int32_t* d_status,
                 unsigned long long pairs_cap) {
    sgc_ctx* c = t->ctx;
    memset(&a, 0, sizeof(a));
    a.seq = seq;
    a.segdesc = p.segdesc;
    a.nsegs_p = p.nsegs_p;
    a.status = d_status;
    a.scal = c->d_scal;
    a.k = k;
    a.tbl = t->buckets;
    a.nbuckets = t->nbuckets;
    a.nlines = t->nlines;
    a.home_shift = t->home_shift;
    a.useq = t->useq();
    a.brk = t->brk;
    if (t->peer_n > 0) {
        a.useq = t->peer_useq;
        a.brk = t->peer_brk;
        a.ustride = t->peer_ustride;
        a.bstride = t->peer_bstride;
        a.n_shards = t->peer_n;
        for (int r = 0; r < t->peer_n; r++) a.peer_tbl[r] = t->peer_tbl[r];
        for (int r = 0; r <= t->peer_n; r++) a.id_bounds[r] = t->id_bounds[r];
        int b = 0;
        while ((1 << b) < t->peer_n) b++;
        a.owner_bits = b;
        if (t->peer_filt && !getenv("SGC_NO_FILTER")) {
            a.filt = t->peer_filt;
            a.filt_shift = 64 - t->peer_filt_bits;
            a.my_rank = t->peer_rank;
        }
    }
    a.pairs = (unsigned long long*)c->pairs_a.p;
    a.pairs_cap = pairs_cap;
}

// the labelseq kernel counts hits only: n_miss = (k-mers of the plan) - n_hits, computed on the device
int labelseq_misses(sgc_ctx* c, const Plan& p, int64_t nseq) {
    misses_from_total_kernel<<<1, 1, 0, c->stream>>>(p.kmer_off + nseq, c->d_scal);
    c->launches++;
    CU(cudaGetLastError());
    return SGC_OK;
}

void label_pending(sgc_table* t, int64_t nseq, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
                   unsigned long long pairs_cap) {
    t->label_pending = true;
    t->label_counted = false;
    t->label_nseq = nseq;
    t->label_ids_off = d_ids_off;
    t->label_ids_out = d_ids_out;
    t->label_ids_cap = ids_cap;
    t->label_pairs_cap = pairs_cap;
}

// geometry of a batch of packed records: k-mer offsets, segment offsets, 16-byte unit offsets, (optionally)
// base offsets; the segment map points into the packed buffer
int make_packed_plan(sgc_ctx* c, const int32_t* d_len, int uniform_len, int64_t nseq, int64_t packed_bytes, int k,
                     int seg_k, bool want_base_off, Plan* plan) {
    const size_t n1 = (size_t)nseq + 1;
    TRY(ensure(c, c->nk, n1 * 8));
    TRY(ensure(c, c->nseg, n1 * 8));
    TRY(ensure(c, c->seg_first, n1 * 8));
    TRY(ensure(c, c->kmer_off, n1 * 8));
    TRY(ensure(c, c->units, n1 * 8));
    TRY(ensure(c, c->unit_off, n1 * 8));
    // every 4-byte unit holds 16 bases: at most one segment per seg_k bases plus one per record
    const int64_t max_segs = nseq + packed_bytes * 4 / seg_k + 1;
    TRY(ensure(c, c->seg_map, (size_t)max_segs * sizeof(SegDesc)));
    CU(cudaMemsetAsync(&c->d_scal->bad_offsets, 0, sizeof(int), c->stream));
    if (!d_len && !want_base_off) {  // records of one length: closed form, one kernel
        packed_uniform_plan_kernel<<<grid_for((int64_t)n1, 256, c->num_sms * 32), 256, 0, c->stream>>>(
            uniform_len, nseq, k, seg_k, packed_bytes, (int64_t*)c->kmer_off.p, (int64_t*)c->seg_first.p,
            (int64_t*)c->unit_off.p, (SegDesc*)c->seg_map.p, c->d_scal);
        c->launches++;
        CU(cudaGetLastError());
        plan->kmer_off = (int64_t*)c->kmer_off.p;
        plan->seg_first = (int64_t*)c->seg_first.p;
        plan->segdesc = (const SegDesc*)c->seg_map.p;
        plan->nsegs_p = (int64_t*)c->seg_first.p + nseq;
        return SGC_OK;
    }
    packed_count_kernel<<<grid_for((int64_t)n1, 256, 1 << 30), 256, 0, c->stream>>>(
        d_len, uniform_len, nseq, k, seg_k, (int64_t*)c->nk.p, (int64_t*)c->nseg.p, (int64_t*)c->units.p, c->d_scal);
    c->launches++;
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t*)c->nk.p, (int64_t*)c->kmer_off.p, (int64_t)n1, c->stream));
    TRY(ensure(c, c->cub_tmp, tb));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->nk.p, (int64_t*)c->kmer_off.p, (int64_t)n1, c->stream));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->nseg.p, (int64_t*)c->seg_first.p, (int64_t)n1, c->stream));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->units.p, (int64_t*)c->unit_off.p, (int64_t)n1, c->stream));
    if (nseq > 0) {
        packed_expand_kernel<<<grid_for(nseq, 256, 1 << 30), 256, 0, c->stream>>>(
            (int64_t*)c->seg_first.p, (int64_t*)c->unit_off.p, (int64_t*)c->nk.p, (int64_t*)c->kmer_off.p, nseq, seg_k,
            packed_bytes, (SegDesc*)c->seg_map.p, c->d_scal);
        c->launches++;
    }
    if (want_base_off) {  // offsets in bases (the ASCII layout): scan of len
        TRY(ensure(c, c->base_off, n1 * 8));
        len_to_i64_kernel<<<grid_for((int64_t)n1, 256, 1 << 30), 256, 0, c->stream>>>(d_len, uniform_len, nseq,
                                                                                    (int64_t*)c->nk.p);
        c->launches++;
        CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->nk.p, (int64_t*)c->base_off.p, (int64_t)n1, c->stream));
    }
    CU(cudaGetLastError());
    plan->kmer_off = (int64_t*)c->kmer_off.p;
    plan->seg_first = (int64_t*)c->seg_first.p;
    plan->segdesc = (const SegDesc*)c->seg_map.p;
    plan->nsegs_p = (int64_t*)c->seg_first.p + nseq;
    return SGC_OK;
}

}  // namespace

extern "C" {

int sgc_label_reads_launch(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq,
                           int k, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap, int32_t* d_status) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    if (!d_ids_off) return fail(SGC_ERR_ARG, "d_ids_off is NULL");
    if (ids_cap < 0) return fail(SGC_ERR_ARG, "ids_cap < 0");
    TRY(set_device(c));
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    const bool side = use_side(t);
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, nullptr, &p, side ? SEG_KS : SEG_K));
    unsigned long long pairs_cap = 0;
    TRY(label_common_prologue(t, nseq, ids_cap, &pairs_cap));
    bool counted = false;
    if (nseq > 0) {
        if (side) {
            LabelSeqArgs a;
            labelseq_args(a, t, d_seq, p, k, d_status, pairs_cap);
            a.cnt = (unsigned long long*)d_ids_off;  // the caller's offsets double as the counters (pairs_to_csr)
            CU(cudaMemsetAsync(d_ids_off, 0, (size_t)(nseq + 1) * 8, c->stream));
            counted = true;
            TRY(launch_labelseq<false>(c, a));
            TRY(labelseq_misses(c, p, nseq));
        } else {
            TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
            table_args(a, t);
            a.pairs = (unsigned long long*)c->pairs_a.p;
            a.pairs_cap = pairs_cap;
            TRY(launch_tile<MODE_LABEL>(c, a));
        }
    }
    CU(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
    label_pending(t, nseq, d_ids_off, d_ids_out, ids_cap, pairs_cap);
    t->label_counted = counted;
    return SGC_OK;
}

int64_t sgc_packed_bytes(const int32_t* h_len, int uniform_len, int64_t nseq) {
    if (nseq < 0) return -1;
    if (!h_len) return uniform_len < 0 ? -1 : nseq * (int64_t)((uniform_len + 15) / 16) * 4;
    int64_t units = 0;
    for (int64_t s = 0; s < nseq; s++) {
        if (h_len[s] < 0) return -1;
        units += ((int64_t)h_len[s] + 15) / 16;
    }
    return units * 4;
}

int sgc_label_reads_packed_launch(sgc_table* t, const uint8_t* d_packed, int64_t packed_bytes, const int32_t* d_len,
                                  int uniform_len, int64_t nseq, int k, int64_t* d_ids_off, uint32_t* d_ids_out,
                                  int64_t ids_cap) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    if (k < 1 || k > SGC_MAX_K) return fail(SGC_ERR_ARG, "k must be in 1..%d (got %d)", SGC_MAX_K, k);
    if (nseq < 0 || packed_bytes < 0 || nseq >= (int64_t)1 << 32) return fail(SGC_ERR_ARG, "bad size");
    if (!d_packed && packed_bytes > 0) return fail(SGC_ERR_ARG, "d_packed is NULL");
    if (((uintptr_t)d_packed & 15) != 0) return fail(SGC_ERR_ARG, "d_packed must be 16-byte aligned");
    if (!d_len && uniform_len < 0) return fail(SGC_ERR_ARG, "need d_len or uniform_len >= 0");
    if (!d_ids_off) return fail(SGC_ERR_ARG, "d_ids_off is NULL");
    if (ids_cap < 0) return fail(SGC_ERR_ARG, "ids_cap < 0");
    TRY(set_device(c));
    const bool side = use_side(t);
    Plan p;
    TRY(make_packed_plan(c, d_len, uniform_len, nseq, packed_bytes, k, SEG_KS, !side, &p));
    unsigned long long pairs_cap = 0;
    TRY(label_common_prologue(t, nseq, ids_cap, &pairs_cap));
    bool counted = false;
    if (nseq > 0) {
        if (side) {
            LabelSeqArgs a;
            labelseq_args(a, t, d_packed, p, k, nullptr, pairs_cap);
            a.cnt = (unsigned long long*)d_ids_off;
            CU(cudaMemsetAsync(d_ids_off, 0, (size_t)(nseq + 1) * 8, c->stream));
            counted = true;
            TRY(launch_labelseq<true>(c, a));
            TRY(labelseq_misses(c, p, nseq));
        } else {
            // table without a side array (loaded from (hash, id) pairs): expand to ASCII, per-k-mer probing
            TRY(ensure(c, c->unpacked, (size_t)std::max<int64_t>(packed_bytes * 4, 16)));
            unpack_reads_kernel<<<c->num_sms * 16, 256, 0, c->stream>>>((const uint32_t*)d_packed, (const int64_t*)c->base_off.p,
                                                                       (const int64_t*)c->unit_off.p, nseq,
                                                                       (uint8_t*)c->unpacked.p);
            c->launches++;
            Plan pa;
            TRY(make_plan(c, (const int64_t*)c->base_off.p, nseq, packed_bytes * 4, k, nullptr, &pa));
            TileArgs a = base_args(c, (const uint8_t*)c->unpacked.p, (const int64_t*)c->base_off.p, k, pa, nullptr);
            table_args(a, t);
            a.pairs = (unsigned long long*)c->pairs_a.p;
            a.pairs_cap = pairs_cap;
            TRY(launch_tile<MODE_LABEL>(c, a));
        }
    }
    CU(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
    label_pending(t, nseq, d_ids_off, d_ids_out, ids_cap, pairs_cap);
    t->label_counted = counted;
    return SGC_OK;
}

int sgc_label_reads_packed(sgc_table* t, const uint8_t* d_packed, int64_t packed_bytes, const int32_t* d_len,
                           int uniform_len, int64_t nseq, int k, int64_t* d_ids_off, uint32_t* d_ids_out,
                           int64_t ids_cap, int64_t* h_n_kmers, int64_t* h_n_hits, int64_t* h_n_labels) {
    TRY(sgc_label_reads_packed_launch(t, d_packed, packed_bytes, d_len, uniform_len, nseq, k, d_ids_off, d_ids_out, ids_cap));
    return sgc_label_reads_finish(t, h_n_kmers, h_n_hits, h_n_labels);
}

// ASCII (seq, off) on the device -> packed records (tests / bench; the host feeder packs with sgc_pack_reads)
int sgc_pack_reads_device(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq,
                          uint8_t* d_packed, int64_t packed_cap, int32_t* d_len, int32_t* d_status, int64_t* h_packed_bytes) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, 1));
    if (!h_packed_bytes) return fail(SGC_ERR_ARG, "h_packed_bytes is NULL");
    TRY(set_device(c));
    const size_t n1 = (size_t)nseq + 1;
    TRY(ensure(c, c->units, n1 * 8));
    TRY(ensure(c, c->unit_off, n1 * 8));
    CU(cudaMemsetAsync(&c->d_scal->bad_offsets, 0, sizeof(int), c->stream));
    off_to_units_kernel<<<grid_for((int64_t)n1, 256, 1 << 30), 256, 0, c->stream>>>(d_off, nseq, seq_bytes, (int64_t*)c->units.p,
                                                                                  d_len, c->d_scal);
    c->launches++;
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t*)c->units.p, (int64_t*)c->unit_off.p, (int64_t)n1, c->stream));
    TRY(ensure(c, c->cub_tmp, tb));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->units.p, (int64_t*)c->unit_off.p, (int64_t)n1, c->stream));
    int64_t units = 0;
    CU(cudaMemcpyAsync(&units, (int64_t*)c->unit_off.p + nseq, 8, cudaMemcpyDeviceToHost, c->stream));
    TRY(fetch_scalars(c));
    TRY(bad_offsets_check(c));
    *h_packed_bytes = units * 4;
    if (units * 4 > packed_cap) return fail(SGC_ERR_CAPACITY, "packed_cap %lld < %lld bytes", (long long)packed_cap, (long long)(units * 4));
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    if (units > 0) {
        pack_reads_kernel<<<grid_for(units, 256, c->num_sms * 16), 256, 0, c->stream>>>(
            d_seq, d_off, (const int64_t*)c->unit_off.p, nseq, (uint32_t*)d_packed, units, d_status);
        c->launches++;
    }
    CU(cudaGetLastError());
    return SGC_OK;
}

int sgc_label_reads_finish(sgc_table* t, int64_t* h_n_kmers, int64_t* h_n_hits, int64_t* h_n_labels) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (!t->label_pending) return fail(SGC_ERR_ARG, "no label call pending on this table");
    sgc_ctx* c = t->ctx;
    t->label_pending = false;
    TRY(set_device(c));
    CU(cudaStreamSynchronize(c->stream));
    const unsigned long long n_pairs = c->h_scal->pair_count;
    const unsigned long long hits = c->h_scal->n_hits, miss = c->h_scal->n_miss;
    if (h_n_hits) *h_n_hits = (int64_t)hits;
    if (h_n_kmers) *h_n_kmers = (int64_t)(hits + miss);
    TRY(bad_offsets_check(c));
    if (n_pairs > t->label_pairs_cap) {
        if (h_n_labels) *h_n_labels = (int64_t)n_pairs;
        return fail(SGC_ERR_CAPACITY, "candidate buffer %llu < %llu (retry with ids_cap >= %llu)", t->label_pairs_cap,
                    n_pairs, n_pairs);
    }
    return pairs_to_csr(c, (int64_t)n_pairs, t->label_nseq, t->label_ids_off, t->label_ids_out, t->label_ids_cap,
                        h_n_labels, t->label_counted);
}

int sgc_label_reads(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                    int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap, int32_t* d_status, int64_t* h_n_kmers,
                    int64_t* h_n_hits, int64_t* h_n_labels) {
    TRY(sgc_label_reads_launch(t, d_seq, seq_bytes, d_off, nseq, k, d_ids_off, d_ids_out, ids_cap, d_status));
    return sgc_label_reads_finish(t, h_n_kmers, h_n_hits, h_n_labels);
}

int sgc_offsets_to_counts8(sgc_ctx* c, const int64_t* d_off, int64_t n, uint8_t* d_counts) {
    if (!c || n < 0 || !d_counts || (n > 0 && !d_off)) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    CU(cudaMemsetAsync(d_counts + n, 0, 1, c->stream));
    if (n > 0) {
        offsets_to_counts8_kernel<<<grid_for(n, 256, c->num_sms * 16), 256, 0, c->stream>>>(d_off, n, d_counts);
        c->launches++;
        CU(cudaGetLastError());
    }
    return SGC_OK;
}

int sgc_labels_from_ids(sgc_ctx* c, const uint32_t* d_kmer_ids, const int64_t* d_kmer_off, int64_t nseq,
                        int64_t n_kmers, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap, int64_t* h_n_hits,
                        int64_t* h_n_labels) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    if (nseq < 0 || n_kmers < 0 || nseq >= (int64_t)1 << 32) return fail(SGC_ERR_ARG, "bad size");
    TRY(set_device(c));
    unsigned long long pairs_cap = (unsigned long long)std::max<int64_t>(std::max<int64_t>(ids_cap * 2, nseq * 4), 4096);
    TRY(ensure(c, c->pairs_a, (size_t)pairs_cap * 8));
    pairs_cap = c->pairs_a.cap / 8;
    TRY(zero_scalars(c));
    if (nseq > 0) {
        ids_to_pairs_kernel<<<grid_for(nseq, 128, c->num_sms * 16), 128, 0, c->stream>>>(
            d_kmer_ids, d_kmer_off, nseq, (unsigned long long*)c->pairs_a.p, pairs_cap, c->d_scal);
        c->launches++;
        CU(cudaGetLastError());
    }
    TRY(fetch_scalars(c));
    const unsigned long long n_pairs = c->h_scal->pair_count;
    if (h_n_hits) *h_n_hits = (int64_t)c->h_scal->n_hits;
    if (n_pairs > pairs_cap) {
        if (h_n_labels) *h_n_labels = (int64_t)n_pairs;
        return fail(SGC_ERR_CAPACITY, "candidate buffer %llu < %llu", pairs_cap, n_pairs);
    }
    return pairs_to_csr(c, (int64_t)n_pairs, nseq, d_ids_off, d_ids_out, ids_cap, h_n_labels);
}

// ---------------------------------------------------------------------------- K6
int sgc_count_doms(sgc_ctx* c, const uint32_t* d_count_by_id, const int64_t* d_member_off, const uint32_t* d_member,
                   const uint32_t* d_level_nodes, const int64_t* h_level_off, int n_levels, uint32_t* d_node_count,
                   int64_t n_nodes) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    if (n_levels < 0 || n_nodes < 0 || (n_levels && !h_level_off)) return fail(SGC_ERR_ARG, "bad catlas arguments");
    TRY(set_device(c));
    if (n_nodes) CU(cudaMemsetAsync(d_node_count, 0, (size_t)n_nodes * 4, c->stream));
    for (int l = 0; l < n_levels; l++) {
        int64_t n = h_level_off[l + 1] - h_level_off[l];
        if (n <= 0) continue;
        dom_level_kernel<<<grid_for(n, 128, 1 << 30), 128, 0, c->stream>>>(
            l == 0 ? d_count_by_id : d_node_count, d_member_off, d_member, d_level_nodes + h_level_off[l], n,
            d_node_count);
        c->launches++;
    }
    CU(cudaGetLastError());
    return SGC_OK;
}

// ---------------------------------------------------------------------------- K7
static int log2_ranks(int n_ranks) {
    int b = 0;
    while ((1 << b) < n_ranks) b++;
    return ((1 << b) == n_ranks && n_ranks <= 256) ? b : -1;
}

int sgc_partition_pairs(sgc_ctx* c, const uint64_t* d_hash, const uint32_t* d_id, int64_t n, int n_ranks,
                        uint64_t* d_hash_out, uint32_t* d_id_out, int64_t* d_perm, int64_t* d_bucket_off,
                        int64_t* h_bucket_off) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    const int bits = log2_ranks(n_ranks);
    if (bits < 0) return fail(SGC_ERR_ARG, "n_ranks must be a power of two <= 256 (got %d)", n_ranks);
    if (n < 0) return fail(SGC_ERR_ARG, "n < 0");
    TRY(set_device(c));
    if (n > 0) {
        size_t tb = 0;
        if (d_perm) {
            TRY(ensure(c, c->tmp_idx, (size_t)n * 8));
            iota_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>((int64_t*)c->tmp_idx.p, n);
            c->launches++;
            CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_hash, d_hash_out, (int64_t*)c->tmp_idx.p, d_perm, n,
                                               64 - bits, 64, c->stream));
            TRY(ensure(c, c->cub_tmp, tb));
            if (bits)
                CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, d_hash, d_hash_out, (int64_t*)c->tmp_idx.p, d_perm, n,
                                                   64 - bits, 64, c->stream));
            else {
                CU(cudaMemcpyAsync(d_hash_out, d_hash, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpyAsync(d_perm, c->tmp_idx.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
            }
            if (d_id && d_id_out) {
                gather_u32_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(d_id, d_perm, n, d_id_out, 0);
                c->launches++;
            }
        } else if (d_id && d_id_out) {
            if (bits) {
                CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_hash, d_hash_out, d_id, d_id_out, n, 64 - bits, 64,
                                                   c->stream));
                TRY(ensure(c, c->cub_tmp, tb));
                CU(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tb, d_hash, d_hash_out, d_id, d_id_out, n, 64 - bits, 64,
                                                   c->stream));
            } else {
                CU(cudaMemcpyAsync(d_hash_out, d_hash, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpyAsync(d_id_out, d_id, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream));
            }
        } else {
            return fail(SGC_ERR_ARG, "need d_perm or (d_id, d_id_out)");
        }
    }
    bucket_bounds_kernel<<<1, 512, 0, c->stream>>>(d_hash_out, n, bits, n_ranks, d_bucket_off);
    c->launches++;
    CU(cudaGetLastError());
    if (h_bucket_off) {
        CU(cudaMemcpyAsync(h_bucket_off, d_bucket_off, (size_t)(n_ranks + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return SGC_OK;
}

int sgc_hash_partition(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                       int n_ranks, uint64_t* d_bucketed, int64_t* d_perm, int64_t cap, int64_t* d_bucket_off,
                       int64_t* d_kmer_off, int32_t* d_status, int64_t* h_bucket_off) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    if (log2_ranks(n_ranks) < 0) return fail(SGC_ERR_ARG, "n_ranks must be a power of two <= 256 (got %d)", n_ranks);
    TRY(set_device(c));
    TRY(ensure(c, c->tmp_hash, (size_t)std::max<int64_t>(cap, 1) * 8));
    int64_t total = 0;
    TRY(sgc_hash_batch(c, d_seq, seq_bytes, d_off, nseq, k, (uint64_t*)c->tmp_hash.p, cap, d_kmer_off, d_status, &total));
    return sgc_partition_pairs(c, (const uint64_t*)c->tmp_hash.p, nullptr, total, n_ranks, d_bucketed, nullptr, d_perm,
                               d_bucket_off, h_bucket_off);
}

// fused K7: hash + owner bucketing in ONE pass (no sort, no permutation array)
static int partition_hashes_impl(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off,
                                 int64_t nseq, int k, int n_ranks, uint64_t* d_send, int64_t send_cap,
                                 uint32_t* d_ridx, int64_t ridx_cap, int32_t* d_status, int64_t* h_counts,
                                 uint64_t* const* h_peer_inboxes, int my_rank) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    const int bits = log2_ranks(n_ranks);
    if (bits < 0 || n_ranks > 32) return fail(SGC_ERR_ARG, "n_ranks must be a power of two <= 32 (got %d)", n_ranks);
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    if (send_cap < 0 || ridx_cap < 0 || !h_counts) return fail(SGC_ERR_ARG, "bad argument");
    if (send_cap >= ((int64_t)1 << (32 - bits))) return fail(SGC_ERR_ARG, "send_cap too large for 32-bit reply indices");
    TRY(set_device(c));
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, nullptr, &p));
    CU(cudaMemcpyAsync(&c->h_scal->n_unique, p.kmer_off + nseq, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const int64_t total = c->h_scal->n_unique;
    if (total > ridx_cap) return fail(SGC_ERR_CAPACITY, "ridx_cap %lld < %lld k-mers", (long long)ridx_cap, (long long)total);
    TRY(ensure(c, c->cursor, 32 * 8));
    CU(cudaMemsetAsync(c->cursor.p, 0, 32 * 8, c->stream));
    TRY(zero_scalars(c));
    if (nseq > 0 && total > 0) {
        TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
        a.send = d_send;
        a.send_cap = (unsigned long long)send_cap;
        for (int r = 0; r < n_ranks; r++)
            a.peer_send[r] = h_peer_inboxes ? h_peer_inboxes[r] : d_send + (size_t)r * (size_t)send_cap;
        a.send_region_off = h_peer_inboxes ? (unsigned long long)my_rank * (unsigned long long)send_cap : 0ULL;
        a.cursor = (unsigned long long*)c->cursor.p;
        a.ridx = d_ridx;
        a.owner_bits = bits;
        a.n_ranks = n_ranks;
        TRY(launch_tile<MODE_PARTITION>(c, a));
    }
    unsigned long long counts[32];
    CU(cudaMemcpyAsync(counts, c->cursor.p, 32 * 8, cudaMemcpyDeviceToHost, c->stream));
    TRY(fetch_scalars(c));
    TRY(bad_offsets_check(c));
    for (int r = 0; r < n_ranks; r++) h_counts[r] = (int64_t)counts[r];  // every owner, so that a retry can size for the maximum
    for (int r = 0; r < n_ranks; r++)
        if ((int64_t)counts[r] > send_cap)
            return fail(SGC_ERR_CAPACITY, "send_cap %lld < %llu hashes for owner %d", (long long)send_cap, counts[r], r);
    return SGC_OK;
}

int sgc_partition_hashes(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                         int n_ranks, uint64_t* d_send, int64_t send_cap, uint32_t* d_ridx, int64_t ridx_cap,
                         int32_t* d_status, int64_t* h_counts) {
    if (!d_send) return fail(SGC_ERR_ARG, "d_send is NULL");
    return partition_hashes_impl(c, d_seq, seq_bytes, d_off, nseq, k, n_ranks, d_send, send_cap, d_ridx, ridx_cap,
                                 d_status, h_counts, nullptr, 0);
}

// the same with the exchange fused in: every hash is stored straight into its owner GPU's inbox
int sgc_partition_hashes_p2p(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq,
                             int k, int n_ranks, int my_rank, uint64_t* const* h_peer_inboxes, int64_t region_cap,
                             uint32_t* d_ridx, int64_t ridx_cap, int32_t* d_status, int64_t* h_counts) {
    if (!h_peer_inboxes || my_rank < 0 || my_rank >= n_ranks) return fail(SGC_ERR_ARG, "bad peer arguments");
    return partition_hashes_impl(c, d_seq, seq_bytes, d_off, nseq, k, n_ranks, nullptr, region_cap, d_ridx, ridx_cap,
                                 d_status, h_counts, h_peer_inboxes, my_rank);
}

// owner side of the fused exchange: probe the requests of every source region of my inbox and store
// each reply straight into the source GPU's reply buffer (region my_rank, same position)
int sgc_lookup_regions_p2p(sgc_table* t, const uint64_t* d_inbox, const int64_t* h_counts_from, int64_t region_cap,
                           int n_ranks, int my_rank, uint32_t* const* h_peer_replies, int ctas_per_sm) {
    if (!t || !d_inbox || !h_counts_from || !h_peer_replies) return fail(SGC_ERR_ARG, "NULL argument");
    if (n_ranks < 1 || n_ranks > 32 || my_rank < 0 || my_rank >= n_ranks) return fail(SGC_ERR_ARG, "bad ranks");
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    for (int s = 0; s < n_ranks; s++) {
        const int64_t n = h_counts_from[s];
        if (n < 0 || n > region_cap) return fail(SGC_ERR_ARG, "bad region count");
        if (n == 0) continue;
        // ctas_per_sm < 8 leaves SM room for a concurrently running hash kernel (the probe is bound by
        // HBM accesses and saturates with few resident threads)
        const int per_sm = ctas_per_sm > 0 && ctas_per_sm < 8 ? ctas_per_sm : 8;
        lookup_kernel<<<grid_for((n + 1) / 2, 256, c->num_sms * per_sm), 256, 0, c->stream>>>(
            t->buckets, t->nbuckets, t->nlines, t->home_shift, d_inbox + (size_t)s * (size_t)region_cap, n,
            h_peer_replies[s] + (size_t)my_rank * (size_t)region_cap);
        c->launches++;
    }
    CU(cudaGetLastError());
    return SGC_OK;
}

// peer-visible device memory (CUDA IPC): allocate here, open in the other ranks' processes
int sgc_peer_alloc(sgc_ctx* c, int64_t bytes, void** d_ptr, uint8_t* h_handle64) {
    if (!c || !d_ptr || !h_handle64 || bytes <= 0) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(d_ptr, (size_t)bytes);
    if (e != cudaSuccess) return fail(SGC_ERR_NOMEM, "cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    CU(cudaMemset(*d_ptr, 0, (size_t)bytes));  // mailboxes must start at zero (epochs count from 1)
    cudaIpcMemHandle_t hnd;
    CU(cudaIpcGetMemHandle(&hnd, *d_ptr));
    memcpy(h_handle64, &hnd, 64);
    return SGC_OK;
}

int sgc_peer_open(sgc_ctx* c, const uint8_t* h_handle64, void** d_ptr) {
    if (!c || !d_ptr || !h_handle64) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, h_handle64, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    return SGC_OK;
}

int sgc_peer_close(sgc_ctx* c, void* d_ptr) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    TRY(set_device(c));
    if (d_ptr) CU(cudaIpcCloseMemHandle(d_ptr));
    return SGC_OK;
}

int sgc_peer_free(sgc_ctx* c, void* d_ptr) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    TRY(set_device(c));
    if (d_ptr) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(d_ptr));
    }
    return SGC_OK;
}

// ---- partitioned table probed over peer memory ----
int sgc_table_export_fd(sgc_table* t, void** d_ptr, int64_t* h_bytes, int* h_fd) {
    if (!t || !d_ptr) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(t->ctx));
    *d_ptr = t->buckets;
    if (h_bytes) *h_bytes = t->vmm ? (int64_t)t->vmm_bytes : (int64_t)t->nbuckets * (int64_t)sizeof(Bucket);
    if (h_fd) {
        if (!t->vmm) return fail(SGC_ERR_ARG, "only tables made by sgc_table_create_shared can be exported");
        const VmmApi* v = vmm_api();
        int fd = -1;
        CD(v->export_fd(&fd, (CUmemGenericAllocationHandle)t->vmm_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        *h_fd = fd;
    }
    return SGC_OK;
}

int sgc_peer_import_fd(sgc_ctx* c, int fd, int64_t bytes, void** d_ptr) {
    if (!c || !d_ptr || bytes <= 0 || fd < 0) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    const VmmApi* v = vmm_api();
    if (!v) return fail(SGC_ERR_CUDA, "the CUDA driver does not offer the virtual memory management API");
    CUmemGenericAllocationHandle h;
    CD(v->import_fd(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    int rc = vmm_map(v, c->device, h, (size_t)bytes, (size_t)1 << 21, d_ptr);
    v->release(h);  // the mapping keeps the memory alive
    return rc;
}

int sgc_peer_unmap(sgc_ctx* c, void* d_ptr, int64_t bytes) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    TRY(set_device(c));
    const VmmApi* v = vmm_api();
    if (v && d_ptr) {
        CU(cudaStreamSynchronize(c->stream));
        v->unmap((CUdeviceptr)d_ptr, (size_t)bytes);
        v->addr_free((CUdeviceptr)d_ptr, (size_t)bytes);
    }
    return SGC_OK;
}

int sgc_table_insert_triples(sgc_table* t, const uint64_t* d_hash, const uint32_t* d_id, const uint32_t* d_aux, int64_t n,
                             const uint32_t* h_id_bounds, int n_shards, uint64_t* d_marks, int64_t marks_cap,
                             int64_t* h_n_marks) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (n < 0 || n_shards < 1 || n_shards > 32 || !h_id_bounds || marks_cap < 0) return fail(SGC_ERR_ARG, "bad argument");
    if (h_n_marks) *h_n_marks = 0;
    if (n == 0) return SGC_OK;
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    TRY(zero_scalars(c));
    BrkSink sink;
    sink.list = (unsigned long long*)d_marks;
    sink.list_n = &c->d_scal->n_export;
    sink.list_cap = (unsigned long long)marks_cap;
    sink.n_shards = n_shards;
    for (int r = 0; r <= n_shards; r++) sink.bounds[r] = h_id_bounds[r];
    insert_triples_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(t->buckets, t->nbuckets, t->nlines,
                                                                                 t->home_shift, d_hash, d_id, d_aux, n,
                                                                                 c->d_scal, sink);
    c->launches++;
    CU(cudaGetLastError());
    TRY(fetch_scalars(c));
    if (h_n_marks) *h_n_marks = (int64_t)c->h_scal->n_export;
    TRY(device_err_check(c));
    if ((int64_t)c->h_scal->n_export > marks_cap)
        return fail(SGC_ERR_CAPACITY, "%lld break marks do not fit the list of %lld", (long long)c->h_scal->n_export,
                    (long long)marks_cap);
    return SGC_OK;
}

int sgc_store_build(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                    uint32_t* d_useq, uint32_t* d_brk) {
    if (!c || !d_useq || !d_brk) return fail(SGC_ERR_ARG, "bad argument");
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    TRY(set_device(c));
    if (seq_bytes == 0) return SGC_OK;
    const int64_t nw = (seq_bytes + 15) / 16, nbw = (seq_bytes + 31) / 32;
    useq_pack_kernel<<<grid_for(nw, 256, c->num_sms * 16), 256, 0, c->stream>>>(d_seq, seq_bytes, d_useq, nw);
    useq_brk_kernel<<<grid_for(nbw, 256, c->num_sms * 16), 256, 0, c->stream>>>(d_off, nseq, k, d_brk + (BRK_PAD >> 5), nbw);
    c->launches += 2;
    CU(cudaGetLastError());
    return SGC_OK;
}

int sgc_store_apply_marks(sgc_ctx* c, const uint64_t* d_marks, int64_t n, uint32_t* d_brk, int64_t brk_stride_words) {
    if (!c || n < 0 || (n > 0 && (!d_marks || !d_brk))) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    if (n == 0) return SGC_OK;
    apply_marks_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>((const unsigned long long*)d_marks, n, d_brk,
                                                                              brk_stride_words);
    c->launches++;
    CU(cudaGetLastError());
    return SGC_OK;
}

int sgc_table_set_peers(sgc_table* t, int n_ranks, const void* const* h_bucket_ptrs, const uint32_t* h_id_bounds,
                        const uint32_t* d_useq, int64_t useq_stride_words, const uint32_t* d_brk,
                        int64_t brk_stride_words) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (n_ranks == 0) {
        t->peer_n = 0;
        return SGC_OK;
    }
    if (n_ranks < 2 || n_ranks > 32 || (n_ranks & (n_ranks - 1)) || !h_bucket_ptrs || !h_id_bounds || !d_useq || !d_brk)
        return fail(SGC_ERR_ARG, "bad argument (n_ranks must be a power of two in 2..32)");
    if ((1 << t->home_shift) != n_ranks)
        return fail(SGC_ERR_ARG, "the table's home shift (%d) does not match %d owners", t->home_shift, n_ranks);
    t->peer_n = n_ranks;
    for (int r = 0; r < n_ranks; r++) t->peer_tbl[r] = (const Bucket*)h_bucket_ptrs[r];
    for (int r = 0; r <= n_ranks; r++) t->id_bounds[r] = h_id_bounds[r];
    t->peer_useq = d_useq;
    t->peer_brk = d_brk;
    t->peer_ustride = useq_stride_words;
    t->peer_bstride = brk_stride_words;
    return SGC_OK;
}

int sgc_filter_insert(sgc_ctx* c, const uint64_t* d_hash, int64_t n, uint64_t* d_words, int filt_bits) {
    if (!c || n < 0 || filt_bits < 1 || filt_bits > 40 || (n > 0 && (!d_hash || !d_words))) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    if (n > 0) {
        filter_insert_kernel<<<grid_for(n, 256, c->num_sms * 16), 256, 0, c->stream>>>(d_hash, n, (unsigned long long*)d_words,
                                                                                   64 - filt_bits);
        c->launches++;
        CU(cudaGetLastError());
    }
    return SGC_OK;
}

int sgc_table_set_filter(sgc_table* t, const uint64_t* d_words, int filt_bits, int my_rank) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (d_words && (filt_bits < 1 || filt_bits > 40 || my_rank < 0)) return fail(SGC_ERR_ARG, "bad argument");
    t->peer_filt = (const unsigned long long*)d_words;
    t->peer_filt_bits = filt_bits;
    t->peer_rank = my_rank;
    return SGC_OK;
}

int sgc_labels_from_replies(sgc_ctx* c, const int64_t* d_off, int64_t nseq, int64_t seq_bytes, int k, int n_ranks,
                            const uint32_t* d_ridx, const uint32_t* d_replies, int64_t send_cap, int64_t* d_ids_off,
                            uint32_t* d_ids_out, int64_t ids_cap, uint32_t* d_kmer_ids, int64_t* h_n_kmers,
                            int64_t* h_n_hits, int64_t* h_n_labels) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    const int bits = log2_ranks(n_ranks);
    if (bits < 0 || n_ranks > 32) return fail(SGC_ERR_ARG, "n_ranks must be a power of two <= 32 (got %d)", n_ranks);
    TRY(check_seq_args((const void*)1, seq_bytes, d_off, nseq, k));
    if (!d_ids_off || ids_cap < 0) return fail(SGC_ERR_ARG, "bad label buffers");
    TRY(set_device(c));
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, nullptr, &p));
    unsigned long long pairs_cap = (unsigned long long)std::max<int64_t>(std::max<int64_t>(ids_cap * 2, nseq * 4), 4096);
    TRY(ensure(c, c->pairs_a, (size_t)pairs_cap * 8));
    pairs_cap = c->pairs_a.cap / 8;
    TRY(zero_scalars(c));
    if (nseq > 0) {
        TileArgs a = base_args(c, nullptr, d_off, k, p, nullptr);
        a.pairs = (unsigned long long*)c->pairs_a.p;
        a.pairs_cap = pairs_cap;
        a.ridx = const_cast<uint32_t*>(d_ridx);
        a.replies = d_replies;
        a.send_cap = (unsigned long long)send_cap;
        a.owner_bits = bits;
        a.n_ranks = n_ranks;
        a.kmer_ids = d_kmer_ids;
        TRY(launch_tile<MODE_GATHER>(c, a));
    }
    TRY(fetch_scalars(c));
    const unsigned long long n_pairs = c->h_scal->pair_count;
    if (h_n_hits) *h_n_hits = (int64_t)c->h_scal->n_hits;
    if (h_n_kmers) *h_n_kmers = (int64_t)(c->h_scal->n_hits + c->h_scal->n_miss);
    if (n_pairs > pairs_cap) {
        if (h_n_labels) *h_n_labels = (int64_t)n_pairs;
        return fail(SGC_ERR_CAPACITY, "candidate buffer %llu < %llu", pairs_cap, n_pairs);
    }
    return pairs_to_csr(c, (int64_t)n_pairs, nseq, d_ids_off, d_ids_out, ids_cap, h_n_labels);
}

// One whole labelling step against the hash-range partitioned table with NO host round trip between its
// kernels (flag protocol over NVLink peer memory): request kernel (hash + stores into the owners' inboxes,
// then counts + epoch stamp into their mailboxes), serve kernel (waits per source region for the stamp, probes,
// stores the ids into the requesters' reply buffers, stamps them), gather kernel (waits for every owner's
// stamp, replies -> labels), label tail.  Every rank calls it with the same epoch (1, 2, ...).
int sgc_label_reads_partitioned(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq,
                                int k, int n_ranks, int my_rank, uint32_t epoch, uint64_t* const* h_peer_inboxes,
                                uint32_t* const* h_peer_replies, uint64_t* const* h_peer_mail_count,
                                uint32_t* const* h_peer_mail_stamp, uint32_t* const* h_peer_reply_stamp,
                                int64_t region_cap, uint32_t* d_ridx, int64_t ridx_cap, int64_t* d_ids_off,
                                uint32_t* d_ids_out, int64_t ids_cap, int32_t* d_status, int64_t* h_n_kmers,
                                int64_t* h_n_hits, int64_t* h_n_labels, int32_t* h_flags) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    const int bits = log2_ranks(n_ranks);
    if (bits < 0 || n_ranks > 32) return fail(SGC_ERR_ARG, "n_ranks must be a power of two <= 32 (got %d)", n_ranks);
    if (my_rank < 0 || my_rank >= n_ranks || epoch == 0) return fail(SGC_ERR_ARG, "bad rank / epoch");
    if (!h_peer_inboxes || !h_peer_replies || !h_peer_mail_count || !h_peer_mail_stamp || !h_peer_reply_stamp)
        return fail(SGC_ERR_ARG, "NULL peer table");
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    if (region_cap <= 0 || region_cap >= ((int64_t)1 << (32 - bits))) return fail(SGC_ERR_ARG, "bad region_cap");
    if (ridx_cap < seq_bytes) return fail(SGC_ERR_ARG, "ridx_cap must be >= seq_bytes (an upper bound of the k-mers)");
    if (!d_ids_off || ids_cap < 0 || !d_ridx) return fail(SGC_ERR_ARG, "bad label buffers");
    if (h_flags) *h_flags = 0;
    TRY(set_device(c));
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, nullptr, &p));
    TRY(ensure(c, c->cursor, 32 * 8));
    CU(cudaMemsetAsync(c->cursor.p, 0, 32 * 8, c->stream));
    unsigned long long pairs_cap = (unsigned long long)std::max<int64_t>(std::max<int64_t>(ids_cap * 2, nseq * 4), 4096);
    TRY(ensure(c, c->pairs_a, (size_t)pairs_cap * 8));
    pairs_cap = c->pairs_a.cap / 8;
    TRY(zero_scalars(c));
    {   // requests
        TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
        a.send_cap = (unsigned long long)region_cap;
        for (int r = 0; r < n_ranks; r++) {
            a.peer_send[r] = h_peer_inboxes[r];
            a.peer_mail_count[r] = (unsigned long long*)h_peer_mail_count[r];
            a.peer_mail_stamp[r] = h_peer_mail_stamp[r];
        }
        a.send_region_off = (unsigned long long)my_rank * (unsigned long long)region_cap;
        a.cursor = (unsigned long long*)c->cursor.p;
        a.ridx = d_ridx;
        a.owner_bits = bits;
        a.n_ranks = n_ranks;
        a.my_rank = my_rank;
        a.epoch = epoch;
        TRY(launch_tile<MODE_PARTITION>(c, a));
    }
    {   // owner side
        ServeArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.tbl = t->buckets;
        sa.nbuckets = t->nbuckets;
        sa.nlines = t->nlines;
        sa.home_shift = t->home_shift;
        sa.inbox = h_peer_inboxes[my_rank];
        sa.region_cap = (unsigned long long)region_cap;
        sa.mail_count = (const unsigned long long*)h_peer_mail_count[my_rank];
        sa.mail_stamp = h_peer_mail_stamp[my_rank];
        for (int r = 0; r < n_ranks; r++) {
            sa.peer_replies[r] = h_peer_replies[r];
            sa.peer_reply_stamp[r] = h_peer_reply_stamp[r];
        }
        sa.scal = c->d_scal;
        sa.n_ranks = n_ranks;
        sa.my_rank = my_rank;
        sa.epoch = epoch;
        serve_regions_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(sa);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (nseq > 0) {   // replies -> labels
        CU(cudaMemsetAsync(&c->d_scal->round_ctr, 0, sizeof(unsigned int), c->stream));
        TileArgs a = base_args(c, nullptr, d_off, k, p, nullptr);
        a.pairs = (unsigned long long*)c->pairs_a.p;
        a.pairs_cap = pairs_cap;
        a.ridx = d_ridx;
        a.replies = h_peer_replies[my_rank];
        a.send_cap = (unsigned long long)region_cap;
        a.owner_bits = bits;
        a.n_ranks = n_ranks;
        a.reply_stamp = h_peer_reply_stamp[my_rank];
        a.epoch = epoch;
        TRY(launch_tile<MODE_GATHER>(c, a));
    }
    TRY(fetch_scalars(c));
    const unsigned long long n_pairs = c->h_scal->pair_count;
    if (h_n_hits) *h_n_hits = (int64_t)c->h_scal->n_hits;
    if (h_n_kmers) *h_n_kmers = (int64_t)(c->h_scal->n_hits + c->h_scal->n_miss);
    if (h_flags) *h_flags = c->h_scal->err;
    TRY(bad_offsets_check(c));
    if (c->h_scal->err & 8) return fail(SGC_ERR_CUDA, "partitioned step: a peer GPU's stamp did not arrive (protocol time-out)");
    if (c->h_scal->err & 4) return fail(SGC_ERR_CAPACITY, "partitioned step: a request region overflowed region_cap %lld", (long long)region_cap);
    if (n_pairs > pairs_cap) {
        if (h_n_labels) *h_n_labels = (int64_t)n_pairs;
        return fail(SGC_ERR_CAPACITY, "candidate buffer %llu < %llu", pairs_cap, n_pairs);
    }
    return pairs_to_csr(c, (int64_t)n_pairs, nseq, d_ids_off, d_ids_out, ids_cap, h_n_labels);
}

int sgc_unpermute_u32(sgc_ctx* c, const uint32_t* d_in, const int64_t* d_perm, int64_t n, uint32_t* d_out) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    if (n < 0) return fail(SGC_ERR_ARG, "n < 0");
    if (n == 0) return SGC_OK;
    TRY(set_device(c));
    gather_u32_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(d_in, d_perm, n, d_out, 1);
    c->launches++;
    CU(cudaGetLastError());
    return SGC_OK;
}

// ---------------------------------------------------------------------------- bench utilities
int sgc_ctx_set_timing(sgc_ctx* c, int on) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    c->timing = on != 0;
    return SGC_OK;
}

int sgc_ctx_tile_ms(sgc_ctx* c, double* h_ms_total, int64_t* h_n) {
    if (!c) return fail(SGC_ERR_ARG, "ctx is NULL");
    TRY(set_device(c));
    CU(cudaStreamSynchronize(c->stream));
    double tot = 0;
    for (auto& ev : c->timed) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev.first, ev.second));
        tot += ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (h_ms_total) *h_ms_total = tot;
    if (h_n) *h_n = (int64_t)c->timed.size();
    c->timed.clear();
    return SGC_OK;
}

int sgc_synth_genome(sgc_ctx* c, uint8_t* d_out, int64_t n, uint64_t seed) {
    if (!c || n < 0) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    if (n == 0) return SGC_OK;
    synth_genome_kernel<<<grid_for((n + 31) / 32, 256, c->num_sms * 8), 256, 0, c->stream>>>(d_out, n, seed);
    c->launches++;
    CU(cudaGetLastError());
    return SGC_OK;
}

int sgc_synth_reads(sgc_ctx* c, const uint8_t* d_genome, int64_t glen, int64_t nreads, int read_len, uint64_t seed,
                    uint32_t err_ppm, int rc_half, uint8_t* d_out, int64_t* d_off) {
    if (!c || nreads < 0 || read_len < 1 || glen < read_len) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    if (nreads == 0) {
        CU(cudaMemsetAsync(d_off, 0, 8, c->stream));
        return SGC_OK;
    }
    synth_reads_kernel<<<grid_for(nreads * read_len, 256, c->num_sms * 16), 256, 0, c->stream>>>(
        d_genome, glen, nreads, read_len, seed, err_ppm, rc_half, d_out, d_off);
    c->launches++;
    CU(cudaGetLastError());
    return SGC_OK;
}

int sgc_synth_unitigs(sgc_ctx* c, const uint8_t* d_genome, const int64_t* d_cuts, int64_t n, int k, int64_t out_bytes,
                      uint8_t* d_out, int64_t* d_out_off) {
    if (!c || n < 1 || k < 1 || out_bytes < 1) return fail(SGC_ERR_ARG, "bad argument");
    TRY(set_device(c));
    synth_unitigs_kernel<<<grid_for(out_bytes, 256, c->num_sms * 16), 256, 0, c->stream>>>(d_genome, d_cuts, n, k, d_out,
                                                                                       d_out_off);
    c->launches++;
    CU(cudaGetLastError());
    return SGC_OK;
}

}  // extern "C"

#include "sgc_feed.hpp"  // host feeder: BGZF inflate, record scan, pairing (no device work)

// ---------------------------------------------------------------------------- K8: BBHash (boomphf) dump

struct sgc_mphf {
    sgc_ctx* ctx = nullptr;
    double gamma = 1.0;
    uint64_t nelem = 0, lastbitsetrank = 0;
    MphfLevels lv{};
    uint64_t total_words = 0;  // concatenated level words, every level padded to a multiple of 8
    uint64_t* bits = nullptr;
    uint64_t* ranks = nullptr;  // total_words / 8
    uint64_t* final_keys = nullptr;
    uint64_t* final_vals = nullptr;
    int64_t n_final = 0;
};

namespace {

inline uint64_t mphf_nchar(uint64_t size) { return 1 + size / 64; }  // boomphf bitVector: one spare word

// level sizes exactly as boomphf computes them (constructor and load both call this)
void mphf_level_sizes(uint64_t nelem, double gamma, int nb_levels, MphfLevels& lv) {
    lv.nb_levels = nb_levels;
    double gn = gamma * (double)nelem;
    double proba = nelem ? 1.0 - std::pow((gn - 1.0) / gn, (double)(nelem - 1)) : 0.0;
    uint64_t domain = (uint64_t)std::ceil((double)nelem * gamma);
    for (int l = 0; l < nb_levels; ++l) {
        uint64_t s = nelem ? (((uint64_t)((double)domain * std::pow(proba, (double)l)) + 63) / 64) * 64 : 0;
        lv.size[l] = s ? s : 64;
    }
}

void mphf_layout(sgc_mphf* m) {
    uint64_t w = 0;
    for (int l = 0; l < m->lv.nb_levels; ++l) {
        m->lv.woff[l] = w;
        w += (mphf_nchar(m->lv.size[l]) + 7) / 8 * 8;
    }
    m->total_words = w;
}

MphfView mphf_view(const sgc_mphf* m) {
    MphfView v;
    v.lv = m->lv;
    v.bits = m->bits;
    v.ranks = m->ranks;
    v.final_keys = m->final_keys;
    v.final_vals = m->final_vals;
    v.n_final = m->n_final;
    v.lastbitsetrank = m->lastbitsetrank;
    return v;
}

void mphf_release(sgc_mphf* m) {
    if (!m) return;
    cudaFree(m->bits);
    cudaFree(m->ranks);
    cudaFree(m->final_keys);
    cudaFree(m->final_vals);
    delete m;
}

struct ScopedFree {  // build-time temporaries
    void* p = nullptr;
    ~ScopedFree() { if (p) cudaFree(p); }
};

int mphf_build_impl(sgc_ctx* c, sgc_mphf* m, const uint64_t* d_keys, int64_t n) {
    const int cap = c->num_sms * 16;
    cudaStream_t st = c->stream;
    CU(cudaMalloc(&m->bits, m->total_words * 8));
    CU(cudaMalloc(&m->ranks, m->total_words / 8 * 8));
    CU(cudaMemsetAsync(m->bits, 0, m->total_words * 8, st));
    ScopedFree coll, alive_a, alive_b, counter;
    uint64_t coll_words = mphf_nchar(m->lv.size[0]);
    for (int l = 1; l < m->lv.nb_levels; ++l) coll_words = std::max(coll_words, mphf_nchar(m->lv.size[l]));
    CU(cudaMalloc(&coll.p, coll_words * 8));
    CU(cudaMemsetAsync(coll.p, 0, coll_words * 8, st));
    CU(cudaMalloc(&counter.p, 8));

    const uint64_t* cur = d_keys;
    int64_t n_cur = n;
    for (int l = 0; l < m->lv.nb_levels - 1 && n_cur > 0; ++l) {
        uint64_t* lbits = m->bits + m->lv.woff[l];
        uint64_t size = m->lv.size[l];
        mphf_mark_kernel<<<grid_for(n_cur, 256, cap), 256, 0, st>>>(cur, n_cur, l, size, (uint32_t*)lbits,
                                                                   (uint32_t*)coll.p);
        mphf_settle_kernel<<<grid_for((int64_t)(size / 64), 256, cap), 256, 0, st>>>(lbits, (uint64_t*)coll.p,
                                                                                  (long long)(size / 64));
        // survivors alternate between two buffers; the first is sized n, the second after level 0
        ScopedFree& dst = (l & 1) ? alive_b : alive_a;
        if (!dst.p) CU(cudaMalloc(&dst.p, (size_t)std::max<int64_t>(n_cur, 1) * 8));
        CU(cudaMemsetAsync(counter.p, 0, 8, st));
        mphf_filter_kernel<<<grid_for(n_cur, 256, cap), 256, 0, st>>>(cur, n_cur, l, size, lbits, (uint64_t*)dst.p,
                                                                     (unsigned long long*)counter.p);
        c->launches += 3;
        CU(cudaGetLastError());
        unsigned long long left = 0;
        CU(cudaMemcpyAsync(&left, counter.p, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        cur = (const uint64_t*)dst.p;
        n_cur = (int64_t)left;
    }

    // rank samples: exclusive scan of the 512-bit block popcounts over the concatenated levels
    int64_t nblocks = (int64_t)(m->total_words / 8);
    mphf_block_popc_kernel<<<grid_for(nblocks, 256, cap), 256, 0, st>>>(m->bits, nblocks, m->ranks);
    c->launches++;
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, m->ranks, m->ranks, nblocks, st));
    TRY(ensure(c, c->cub_tmp, tb));
    uint64_t last_rank = 0, last_pop = 0;
    // total set bits = scanned value of the last block + its own popcount (the spare word keeps it zero)
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, m->ranks, m->ranks, nblocks, st));
    CU(cudaMemcpyAsync(&last_rank, m->ranks + nblocks - 1, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    {
        uint64_t h_last[8];
        CU(cudaMemcpy(h_last, m->bits + (nblocks - 1) * 8, 64, cudaMemcpyDeviceToHost));
        for (int j = 0; j < 8; ++j) last_pop += (uint64_t)__builtin_popcountll(h_last[j]);
    }
    m->lastbitsetrank = last_rank + last_pop;

    // leftovers of the last bitset level: explicit final map, keys ascending, value = position
    m->n_final = n_cur;
    if (n_cur > 0) {
        CU(cudaMalloc(&m->final_keys, (size_t)n_cur * 8));
        CU(cudaMalloc(&m->final_vals, (size_t)n_cur * 8));
        CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, cur, m->final_keys, n_cur, 0, 64, st));
        TRY(ensure(c, c->cub_tmp, tb));
        CU(cub::DeviceRadixSort::SortKeys(c->cub_tmp.p, tb, cur, m->final_keys, n_cur, 0, 64, st));
        iota_u64_kernel<<<grid_for(n_cur, 256, cap), 256, 0, st>>>(m->final_vals, n_cur);
        c->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(st));
    return SGC_OK;
}

}  // namespace

extern "C" {

int sgc_mphf_build(sgc_ctx* c, const uint64_t* d_keys, int64_t n, double gamma, int nb_levels, sgc_mphf** out) {
    if (!c || !out) return fail(SGC_ERR_ARG, "ctx/out is NULL");
    *out = nullptr;
    if (n < 0 || (n > 0 && !d_keys)) return fail(SGC_ERR_ARG, "bad key array");
    if (!(gamma >= 1.0)) return fail(SGC_ERR_ARG, "gamma must be >= 1.0 (got %g)", gamma);
    if (nb_levels < 2 || nb_levels > MPHF_MAX_LEVELS)
        return fail(SGC_ERR_ARG, "nb_levels must be in [2, %d] (got %d)", MPHF_MAX_LEVELS, nb_levels);
    TRY(set_device(c));
    sgc_mphf* m = new (std::nothrow) sgc_mphf();
    if (!m) return fail(SGC_ERR_NOMEM, "out of host memory");
    m->ctx = c;
    m->gamma = gamma;
    m->nelem = (uint64_t)n;
    mphf_level_sizes(m->nelem, gamma, nb_levels, m->lv);
    mphf_layout(m);
    int rc = mphf_build_impl(c, m, d_keys, n);
    if (rc != SGC_OK) {
        mphf_release(m);
        return rc;
    }
    *out = m;
    return SGC_OK;
}

int sgc_mphf_load(sgc_ctx* c, const void* h_buf, int64_t bytes, sgc_mphf** out) {
    if (!c || !out || !h_buf) return fail(SGC_ERR_ARG, "ctx/buffer/out is NULL");
    *out = nullptr;
    const uint8_t* p = (const uint8_t*)h_buf;
    int64_t off = 0;
    auto need = [&](int64_t nb) { return nb >= 0 && off + nb <= bytes; };
    if (!need(28)) return fail(SGC_ERR_ARG, "mphf dump truncated (header)");
    double gamma;
    int32_t nb_levels;
    uint64_t last, nelem;
    memcpy(&gamma, p, 8);
    memcpy(&nb_levels, p + 8, 4);
    memcpy(&last, p + 12, 8);
    memcpy(&nelem, p + 20, 8);
    off = 28;
    if (nb_levels < 2 || nb_levels > MPHF_MAX_LEVELS) return fail(SGC_ERR_ARG, "mphf dump: nb_levels %d", nb_levels);
    TRY(set_device(c));
    sgc_mphf* m = new (std::nothrow) sgc_mphf();
    if (!m) return fail(SGC_ERR_NOMEM, "out of host memory");
    m->ctx = c;
    m->gamma = gamma;
    m->nelem = nelem;
    m->lastbitsetrank = last;
    m->lv.nb_levels = nb_levels;
    // first pass: level geometry
    struct Span { int64_t words_at, nchar, ranks_at, nranks; };
    std::vector<Span> spans((size_t)nb_levels);
    for (int l = 0; l < nb_levels; ++l) {
        uint64_t size, nchar, nranks;
        if (!need(16)) { mphf_release(m); return fail(SGC_ERR_ARG, "mphf dump truncated (level %d)", l); }
        memcpy(&size, p + off, 8);
        memcpy(&nchar, p + off + 8, 8);
        off += 16;
        if (size % 64 || nchar != mphf_nchar(size) || !need((int64_t)nchar * 8 + 8)) {
            mphf_release(m);
            return fail(SGC_ERR_ARG, "mphf dump: level %d has size %llu, %llu words", l, (unsigned long long)size,
                        (unsigned long long)nchar);
        }
        spans[l].words_at = off;
        spans[l].nchar = (int64_t)nchar;
        off += (int64_t)nchar * 8;
        memcpy(&nranks, p + off, 8);
        off += 8;
        if (nranks != (nchar + 7) / 8 || !need((int64_t)nranks * 8)) {
            mphf_release(m);
            return fail(SGC_ERR_ARG, "mphf dump: level %d has %llu rank samples", l, (unsigned long long)nranks);
        }
        spans[l].ranks_at = off;
        spans[l].nranks = (int64_t)nranks;
        off += (int64_t)nranks * 8;
        m->lv.size[l] = size;
    }
    mphf_layout(m);
    uint64_t nfinal = 0;
    if (!need(8)) { mphf_release(m); return fail(SGC_ERR_ARG, "mphf dump truncated (final map)"); }
    memcpy(&nfinal, p + off, 8);
    off += 8;
    if (nfinal > (uint64_t)(bytes - off) / 16 || off + (int64_t)nfinal * 16 != bytes) {
        mphf_release(m);
        return fail(SGC_ERR_ARG, "mphf dump: final map of %llu entries does not match the file size",
                    (unsigned long long)nfinal);
    }
    std::vector<uint64_t> hb(m->total_words, 0), hr(m->total_words / 8, 0);
    for (int l = 0; l < nb_levels; ++l) {
        memcpy(hb.data() + m->lv.woff[l], p + spans[l].words_at, (size_t)spans[l].nchar * 8);
        memcpy(hr.data() + m->lv.woff[l] / 8, p + spans[l].ranks_at, (size_t)spans[l].nranks * 8);
    }
    std::vector<std::pair<uint64_t, uint64_t>> fin((size_t)nfinal);
    for (uint64_t i = 0; i < nfinal; ++i) {
        memcpy(&fin[i].first, p + off + i * 16, 8);
        memcpy(&fin[i].second, p + off + i * 16 + 8, 8);
    }
    std::sort(fin.begin(), fin.end());
    auto bail = [&](cudaError_t e, const char* what) {
        mphf_release(m);
        return fail(SGC_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    cudaError_t e;
    if ((e = cudaMalloc(&m->bits, m->total_words * 8)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc(&m->ranks, m->total_words)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemcpy(m->bits, hb.data(), m->total_words * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail(e, "cudaMemcpy");
    if ((e = cudaMemcpy(m->ranks, hr.data(), m->total_words, cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail(e, "cudaMemcpy");
    m->n_final = (int64_t)nfinal;
    if (nfinal) {
        std::vector<uint64_t> fk(nfinal), fv(nfinal);
        for (uint64_t i = 0; i < nfinal; ++i) { fk[i] = fin[i].first; fv[i] = fin[i].second; }
        if ((e = cudaMalloc(&m->final_keys, nfinal * 8)) != cudaSuccess) return bail(e, "cudaMalloc");
        if ((e = cudaMalloc(&m->final_vals, nfinal * 8)) != cudaSuccess) return bail(e, "cudaMalloc");
        if ((e = cudaMemcpy(m->final_keys, fk.data(), nfinal * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
            return bail(e, "cudaMemcpy");
        if ((e = cudaMemcpy(m->final_vals, fv.data(), nfinal * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
            return bail(e, "cudaMemcpy");
    }
    *out = m;
    return SGC_OK;
}

int sgc_mphf_free(sgc_mphf* m) {
    if (!m) return SGC_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    mphf_release(m);
    return SGC_OK;
}

int sgc_mphf_info(sgc_mphf* m, int64_t* h_info, double* h_gamma) {
    if (!m || !h_info) return fail(SGC_ERR_ARG, "mphf/info is NULL");
    h_info[0] = (int64_t)m->nelem;
    h_info[1] = m->lv.nb_levels;
    h_info[2] = (int64_t)m->lastbitsetrank;
    h_info[3] = m->n_final;
    if (h_gamma) *h_gamma = m->gamma;
    return SGC_OK;
}

int sgc_mphf_lookup(sgc_mphf* m, const uint64_t* d_keys, int64_t n, uint64_t* d_index) {
    if (!m) return fail(SGC_ERR_ARG, "mphf is NULL");
    if (n < 0) return fail(SGC_ERR_ARG, "n < 0");
    if (n == 0) return SGC_OK;
    sgc_ctx* c = m->ctx;
    TRY(set_device(c));
    mphf_lookup_kernel<<<grid_for(n, 256, c->num_sms * 16), 256, 0, c->stream>>>(mphf_view(m), d_keys, n, d_index);
    c->launches++;
    CU(cudaGetLastError());
    return SGC_OK;
}

int sgc_mphf_order(sgc_mphf* m, const uint64_t* d_keys, const uint32_t* d_vals, int64_t n, uint64_t* d_hash_out,
                   uint32_t* d_val_out) {
    if (!m) return fail(SGC_ERR_ARG, "mphf is NULL");
    if (n != (int64_t)m->nelem) return fail(SGC_ERR_ARG, "n (%lld) != nelem (%llu)", (long long)n,
                                            (unsigned long long)m->nelem);
    if (n == 0) return SGC_OK;
    sgc_ctx* c = m->ctx;
    TRY(set_device(c));
    TRY(zero_scalars(c));
    mphf_order_kernel<<<grid_for(n, 256, c->num_sms * 16), 256, 0, c->stream>>>(
        mphf_view(m), d_keys, d_vals, n, m->nelem, d_hash_out, d_val_out, &c->d_scal->n_miss);
    c->launches++;
    CU(cudaGetLastError());
    TRY(fetch_scalars(c));
    if (c->h_scal->n_miss)
        return fail(SGC_ERR_ARG, "%llu keys are not in the build set of this mphf", c->h_scal->n_miss);
    return SGC_OK;
}

int sgc_mphf_dump_bytes(sgc_mphf* m, int64_t* h_bytes) {
    if (!m || !h_bytes) return fail(SGC_ERR_ARG, "mphf/out is NULL");
    int64_t b = 28;
    for (int l = 0; l < m->lv.nb_levels; ++l) {
        uint64_t nchar = mphf_nchar(m->lv.size[l]);
        b += 16 + (int64_t)nchar * 8 + 8 + (int64_t)((nchar + 7) / 8) * 8;
    }
    b += 8 + m->n_final * 16;
    *h_bytes = b;
    return SGC_OK;
}

int sgc_mphf_dump(sgc_mphf* m, void* h_buf, int64_t bytes) {
    if (!m || !h_buf) return fail(SGC_ERR_ARG, "mphf/buffer is NULL");
    int64_t want = 0;
    TRY(sgc_mphf_dump_bytes(m, &want));
    if (bytes != want) return fail(SGC_ERR_ARG, "dump buffer is %lld bytes, need %lld", (long long)bytes, (long long)want);
    sgc_ctx* c = m->ctx;
    TRY(set_device(c));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> hb(m->total_words), hr(m->total_words / 8);
    CU(cudaMemcpy(hb.data(), m->bits, m->total_words * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hr.data(), m->ranks, m->total_words, cudaMemcpyDeviceToHost));
    uint8_t* p = (uint8_t*)h_buf;
    int64_t off = 0;
    auto put = [&](const void* src, size_t nb) { memcpy(p + off, src, nb); off += (int64_t)nb; };
    int32_t nb_levels = m->lv.nb_levels;
    put(&m->gamma, 8);
    put(&nb_levels, 4);
    put(&m->lastbitsetrank, 8);
    put(&m->nelem, 8);
    for (int l = 0; l < nb_levels; ++l) {
        uint64_t size = m->lv.size[l], nchar = mphf_nchar(size), nranks = (nchar + 7) / 8;
        put(&size, 8);
        put(&nchar, 8);
        put(hb.data() + m->lv.woff[l], nchar * 8);
        put(&nranks, 8);
        put(hr.data() + m->lv.woff[l] / 8, nranks * 8);
    }
    uint64_t nf = (uint64_t)m->n_final;
    put(&nf, 8);
    if (nf) {
        std::vector<uint64_t> fk(nf), fv(nf);
        CU(cudaMemcpy(fk.data(), m->final_keys, nf * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(fv.data(), m->final_vals, nf * 8, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < nf; ++i) {
            put(&fk[i], 8);
            put(&fv[i], 8);
        }
    }
    return SGC_OK;
}

}  // extern "C"
