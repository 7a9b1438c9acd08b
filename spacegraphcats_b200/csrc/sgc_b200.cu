// sgc_b200.cu -- sm_100a kernels + C-ABI (include/sgc_b200.h) for spacegraphcats'
// k-mer -> cDBG-unitig indexing hot path.  See DESIGN.md for the data layout.
//
//   K1 kmer hash        tile_kernel<K, MODE_HASH>       (khmer get_kmer_hashes, cdbg/hashing.py:14-20)
//   K2 table build      tile_kernel<K, MODE_INSERT>, insert_pairs_kernel
//                                                       (build_mphf, cdbg/index_cdbg_by_kmer.py:27-115)
//   K3 lookup           lookup_kernel / fused in tile_kernel (BBHashTable.__getitem__, hashing.py:36-37)
//   K4 per-read ids     tile_kernel<K, MODE_LABEL> + sort/unique/CSR (cdbg/index_reads.py:143-152)
//   K5 per-unitig count count_kernel, tile_kernel<K, MODE_LOOKUP> (get_unique_values, hashing.py:67-69)
//   K6 per-dominator    dom_level_kernel                (count_catlas_matches, hashing.py:71-107)
//   K7 owner bucketing  partition kernels               (new: hash-range partition across GPUs)
//
// No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
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

// ----------------------------------------------------------------------------------
// tile geometry
// ----------------------------------------------------------------------------------
#ifndef SGC_PROBE_MIN_CTAS
#define SGC_PROBE_MIN_CTAS 2
#endif
constexpr int NT = 256;         // threads per CTA
constexpr int WPB = NT / 32;    // warps per CTA; every warp runs its own rounds (no block barriers)
constexpr int SPW = 8;          // segments staged per warp round (4 lanes stage one segment)
constexpr int SEG_K = 256;      // k-mers per segment
constexpr int PAIR_BUF = 128;   // label keys staged per warp before one global reservation
constexpr int QCAP = 256;       // MODE_LABEL_SIDE: k-mers queued per warp for a dense table probe
constexpr unsigned long long SEQ_MASK = (1ULL << 40) - 1ULL;

constexpr uint32_t VAL_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t VAL_MULTI = 0xFFFFFFFEu;

enum Mode { MODE_HASH = 0, MODE_INSERT = 1, MODE_LOOKUP = 2, MODE_LABEL = 3, MODE_PARTITION = 4, MODE_GATHER = 5, MODE_MINHASH = 6, MODE_LABEL_SIDE = 7 };

template <int K>
struct Geo {
    static constexpr int KMAX = K ? K : SGC_MAX_K;
    static constexpr int SEGB = SEG_K + KMAX - 1;                 // bases in a full segment
    static constexpr int STRIDE = ((SEGB + 32 + 15) / 16) * 16;   // bytes per staged area
    static constexpr int WPS = STRIDE / 4;
};

// The table: 32-byte buckets of two 16-byte slots; four buckets share a 128-byte line (the unit
// HBM actually moves per random access on B200: measured).  A key lives at probe
// step s >= 0 from its home bucket hb: steps 1..3 are the other buckets of the home line, steps
// 4.. walk the following lines.  It sits in the FIRST bucket of that sequence that had a free slot
// when it arrived and slots are never freed, so a walk stops at the first bucket that is not
// full -- and it only starts if the home bucket's overflow bit says that some key homed there
// lives elsewhere (a miss in an un-flagged home bucket costs ONE sector even when it is full).
// aux[30:0] = index of the k-mer in the table's unitig-ordered side array (AUX_NONE if none);
// aux[31] of slot 0 = the bucket's overflow bit.
struct __align__(16) Entry {
    unsigned long long key;
    unsigned int val;  // VAL_EMPTY = free slot, VAL_MULTI = dropped hash, else unitig id
    unsigned int aux;  // side-array position (31 bits) | overflow bit (slot 0)
};
constexpr uint32_t AUX_NONE = 0x7FFFFFFFu;
constexpr uint32_t AUX_MASK = 0x7FFFFFFFu;
constexpr uint32_t OVF_BIT = 0x80000000u;
struct __align__(32) Bucket {
    Entry e[2];
};

struct Scalars {
    unsigned long long pair_count;
    unsigned long long n_hits;
    unsigned long long n_miss;
    unsigned long long n_distinct;
    unsigned long long n_live;
    unsigned long long n_multi;
    unsigned long long n_export;
    long long n_unique;
    unsigned int round_ctr;
    unsigned int max_id;
    int err;  // bit 0: table full, bit 1: id out of range
    int bad_offsets;  // set by the plan when an offset pair is negative, decreasing or past seq_bytes
};

struct TileArgs {
    const uint8_t* seq;
    const int64_t* off;
    const int64_t* seg_map;
    const int64_t* nsegs_p;
    const int64_t* kmer_off;
    int32_t* status;
    Scalars* scal;
    int k;
    // MODE_HASH / MODE_MINHASH
    uint64_t* hash_out;
    unsigned long long* seq_min;    // MODE_MINHASH: per-sequence minimum (atomicMin), nullable
    unsigned long long seed;        // MODE_MINHASH: Murmur3 seed (sourmash: 42)
    // table (MODE_INSERT / MODE_LOOKUP / MODE_LABEL)
    Bucket* tbl;
    uint32_t nbuckets;
    uint32_t nlines;
    int home_shift;
    const uint32_t* seq_ids;
    uint32_t id_base;
    // side array (unitig-ordered {hash, id} records): written by MODE_INSERT, read by MODE_LABEL_SIDE
    uint4* side;
    unsigned long long side_base;   // MODE_INSERT: first record of this batch
    unsigned long long side_n;      // MODE_LABEL_SIDE: number of records
    // MODE_LOOKUP
    uint32_t* kmer_ids;
    uint32_t* count_by_id;
    // MODE_LABEL
    unsigned long long* pairs;
    unsigned long long pairs_cap;
    unsigned long long zero;  // always 0 (scheduling dependency, see the pipelined probe loop)
    // MODE_PARTITION (hash -> owner rank's send region) / MODE_GATHER (replies -> labels)
    uint64_t* send;                 // n_ranks regions of send_cap hashes
    uint64_t* peer_send[32];        // where owner r's requests go: send + r*send_cap locally, or the
                                    // peer-mapped inbox of GPU r (P2P stores over NVLink)
    unsigned long long send_region_off;  // my region inside a peer inbox (my_rank * send_cap), else 0
    unsigned long long send_cap;
    unsigned long long* cursor;     // n_ranks fill counters
    uint32_t* ridx;                 // per k-mer: owner << (32 - owner_bits) | position in the owner's region
    const uint32_t* replies;        // same layout as send, 4 bytes per request
    int owner_bits;
    int n_ranks;
};

// ----------------------------------------------------------------------------------
// table primitives
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t home_bucket(uint64_t h, uint32_t nb, int home_shift) {
    return __umulhi((uint32_t)((h << home_shift) >> 32), nb);
}

__device__ __forceinline__ uint32_t probe_bucket(uint32_t hb, uint32_t step, uint32_t nlines) {
    uint32_t line = (hb >> 2) + (step >> 2);
    if (line >= nlines) line -= nlines;
    return (line << 2) | ((hb + step) & 3u);
}

// one 32-byte sector, read-only path, no L1 allocation (random access, no reuse)
__device__ __forceinline__ void load_bucket_nc(const Bucket* p, Entry& e0, Entry& e1) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
    e0.key = a; e0.val = (uint32_t)b; e0.aux = (uint32_t)(b >> 32);
    e1.key = c; e1.val = (uint32_t)d; e1.aux = (uint32_t)(d >> 32);
}

// coherent variant for the build kernels (the table is being written concurrently)
__device__ __forceinline__ void load_entry_volatile(const Entry* p, Entry& e) {
    unsigned long long a, b;
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    e.key = a; e.val = (uint32_t)b; e.aux = (uint32_t)(b >> 32);
}

__device__ __forceinline__ bool cas_entry(Entry* p, const Entry& expect, const Entry& desired, Entry& old) {
    unsigned long long e0 = expect.key, e1 = ((unsigned long long)expect.aux << 32) | expect.val;
    unsigned long long d0 = desired.key, d1 = ((unsigned long long)desired.aux << 32) | desired.val;
    unsigned long long o0, o1;
    asm volatile(
        "{\n .reg .b128 e, d, o;\n mov.b128 e, {%2, %3};\n mov.b128 d, {%4, %5};\n"
        " atom.global.relaxed.gpu.cas.b128 o, [%6], e, d;\n mov.b128 {%0, %1}, o;\n}\n"
        : "=l"(o0), "=l"(o1)
        : "l"(e0), "l"(e1), "l"(d0), "l"(d1), "l"(p)
        : "memory");
    old.key = o0; old.val = (uint32_t)o1; old.aux = (uint32_t)(o1 >> 32);
    return o0 == e0 && o1 == e1;
}

__device__ __forceinline__ uint32_t resolve_val(uint32_t v) { return v >= VAL_MULTI ? SGC_MISS : v; }

__device__ __forceinline__ bool bucket_match(const Entry& e0, const Entry& e1, uint64_t h, uint32_t& out, uint32_t& aux) {
    if (e0.key == h && e0.val != VAL_EMPTY) { out = resolve_val(e0.val); aux = e0.aux & AUX_MASK; return true; }
    if (e1.key == h && e1.val != VAL_EMPTY) { out = resolve_val(e1.val); aux = e1.aux & AUX_MASK; return true; }
    return false;
}
// true when some key whose home is this bucket was placed in a later bucket
__device__ __forceinline__ bool bucket_overflowed(const Entry& e0) { return (e0.aux & OVF_BIT) != 0; }
__device__ __forceinline__ bool bucket_match(const Entry& e0, const Entry& e1, uint64_t h, uint32_t& out) {
    uint32_t aux;
    return bucket_match(e0, e1, h, out, aux);
}
__device__ __forceinline__ bool bucket_full(const Entry& e0, const Entry& e1) {
    return e0.val != VAL_EMPTY && e1.val != VAL_EMPTY;
}

// Slow path of a lookup: the home bucket is full and does not hold the key.  The other three
// buckets of the home line are read together (the line is already in L2: HBM moved all 128 bytes),
// then, rarely, the following lines; the walk ends at the first bucket that is not full.
__device__ __forceinline__ uint32_t lookup_slow(const Bucket* tbl, uint32_t nlines, uint32_t hb, uint64_t h,
                                                uint32_t& aux) {
    Entry a0, a1, b0, b1, c0, c1;
    load_bucket_nc(tbl + probe_bucket(hb, 1, nlines), a0, a1);
    load_bucket_nc(tbl + probe_bucket(hb, 2, nlines), b0, b1);
    load_bucket_nc(tbl + probe_bucket(hb, 3, nlines), c0, c1);
    uint32_t out;
    aux = AUX_NONE;
    if (bucket_match(a0, a1, h, out, aux)) return out;
    if (!bucket_full(a0, a1)) return SGC_MISS;
    if (bucket_match(b0, b1, h, out, aux)) return out;
    if (!bucket_full(b0, b1)) return SGC_MISS;
    if (bucket_match(c0, c1, h, out, aux)) return out;
    if (!bucket_full(c0, c1)) return SGC_MISS;
    const uint32_t max_steps = nlines > (1u << 20) ? (4u << 20) : nlines * 4u;
    for (uint32_t s = 4; s < max_steps; s++) {
        load_bucket_nc(tbl + probe_bucket(hb, s, nlines), a0, a1);
        if (bucket_match(a0, a1, h, out, aux)) return out;
        if (!bucket_full(a0, a1)) return SGC_MISS;
    }
    return SGC_MISS;
}

__device__ __forceinline__ uint32_t table_lookup(const Bucket* tbl, uint32_t nb, uint32_t nlines, int hs, uint64_t h,
                                                 uint32_t& aux) {
    uint32_t hb = home_bucket(h, nb, hs);
    Entry e0, e1;
    load_bucket_nc(tbl + hb, e0, e1);
    uint32_t out;
    if (bucket_match(e0, e1, h, out, aux)) return out;
    aux = AUX_NONE;
    if (!bucket_overflowed(e0)) return SGC_MISS;
    return lookup_slow(tbl, nlines, hb, h, aux);
}
__device__ __forceinline__ uint32_t table_lookup(const Bucket* tbl, uint32_t nb, uint32_t nlines, int hs, uint64_t h) {
    uint32_t aux;
    return table_lookup(tbl, nb, nlines, hs, h, aux);
}

// build_mphf insertion semantics (cdbg/index_cdbg_by_kmer.py:40-57, 88-89):
//   new hash -> stored with id;  same hash, same id -> no-op;
//   same hash, different id -> the hash is dropped for good (VAL_MULTI).
// The outcome is independent of the order in which threads arrive: slots only ever go from
// empty to occupied, every thread walks the same probe sequence, and a lost CAS re-examines
// what the winner wrote.
__device__ void table_insert(Bucket* tbl, uint32_t nb, uint32_t nlines, int hs, uint64_t h, uint32_t id, uint32_t aux,
                             Scalars* scal) {
    const uint32_t hb = home_bucket(h, nb, hs);
    const Entry empty{0ULL, VAL_EMPTY, 0u};
    const Entry mine{h, id, aux & AUX_MASK};
    const uint32_t max_steps = nlines > (1u << 20) ? (4u << 20) : nlines * 4u;
    for (uint32_t step = 0; step < max_steps; step++) {
        Bucket* bp = tbl + probe_bucket(hb, step, nlines);
#pragma unroll
        for (int s = 0; s < 2; s++) {
            Entry e;
            load_entry_volatile(&bp->e[s], e);
            if (e.val == VAL_EMPTY) {
                Entry old;
                if (cas_entry(&bp->e[s], empty, mine, old)) {
                    // the home bucket is full (its slot 0 is occupied for good): flag it as overflowed
                    if (step) atomicOr(&tbl[hb].e[0].aux, OVF_BIT);
                    return;
                }
                e = old;  // somebody else claimed the slot first: examine what they wrote
            }
            if (e.key == h) {
                if (e.val != id && e.val != VAL_MULTI) atomicExch(&bp->e[s].val, VAL_MULTI);
                return;
            }
        }
    }
    atomicOr(&scal->err, 1);
}

// ----------------------------------------------------------------------------------
// plan kernels: per-sequence k-mer / segment counts and the segment map
// ----------------------------------------------------------------------------------
__global__ void plan_count_kernel(const int64_t* __restrict__ off, int64_t nseq, int k, int64_t seq_bytes,
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
    nseg_out[s] = (nk + SEG_K - 1) / SEG_K;
}

__global__ void plan_expand_kernel(const int64_t* __restrict__ seg_first, int64_t nseq,
                                   int64_t* __restrict__ seg_map) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseq) return;
    int64_t a = seg_first[s], b = seg_first[s + 1];
    for (int64_t j = a; j < b; j++) seg_map[j] = s | ((j - a) << 40);
}

// ----------------------------------------------------------------------------------
// staging helpers
// ----------------------------------------------------------------------------------
// bytes F[f .. f+4) of a segment whose Ls bases start at global byte address `base`; bytes outside
// [0, Ls) read as 0 and no aligned word that lies wholly outside the segment is touched.
__device__ __forceinline__ uint32_t load_window4(uintptr_t base, int f, int Ls, uint32_t& mask) {
    const int lo_clip = f < 0 ? -f : 0;
    const int hi_clip = f + 4 > Ls ? f + 4 - Ls : 0;
    if (lo_clip >= 4 || hi_clip >= 4) { mask = 0; return 0; }
    mask = (0xFFFFFFFFu << (8 * lo_clip)) & (0xFFFFFFFFu >> (8 * hi_clip));
    const uintptr_t addr = base + (intptr_t)f;
    const int s = (int)(addr & 3);
    const uintptr_t a0 = addr - s;
    const uintptr_t first = base, last = base + (uintptr_t)Ls;  // valid byte range [first, last)
    uint32_t lo = 0, hi = 0;
    if (a0 + 4 > first && a0 < last) lo = __ldg((const uint32_t*)a0);
    if (s && a0 + 8 > first && a0 + 4 < last) hi = __ldg((const uint32_t*)(a0 + 4));
    return sgc::funnel_bytes(lo, hi, s) & mask;
}

// reverse complement of 4 packed bases (byte-reversed, A<->T, C<->G); ok = every byte selected by
// mask is one of A, C, G, T.  Bytes that are not pass through unchanged (oracle rule).
__device__ __forceinline__ uint32_t revcomp_word(uint32_t u, uint32_t mask, bool& ok) {
    // 2-bit code per base from ASCII bits 1..2: A=0 C=1 T=2 G=3
    const uint32_t code = (u >> 1) & 0x03030303u;
    const uint32_t t = code | (code >> 4);
    const uint32_t sel = (t & 0xFFu) | ((t >> 8) & 0xFF00u);
    const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);   // "ACTG"[code]
    uint32_t comp = __byte_perm(0x43414754u, 0u, sel);           // "TGAC"[code]
    ok = ((expect ^ u) & mask) == 0;
    if (!ok) {
        comp = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) comp |= (uint32_t)sgc::comp_base((uint8_t)(u >> (8 * b))) << (8 * b);
    }
    return __byte_perm(comp & mask, 0u, 0x0123u);
}

// ----------------------------------------------------------------------------------
// the tile kernel.  Every WARP runs its own rounds: fetch SPW segments, stage their forward and
// reverse-complement bytes in its slice of shared memory (4 lanes per segment), then each lane
// hashes groups of 4 consecutive k-mers and feeds them to the mode's sink.  No block barriers:
// warps drift apart, so one warp's exposed latencies (round fetch, staging loads, table probes)
// are covered by the other warps of the SM.
// ----------------------------------------------------------------------------------
template <int K>
struct WarpTile {
    uint32_t F[SPW * Geo<K>::WPS];
    uint32_t R[SPW * Geo<K>::WPS];
    unsigned long long pairs[PAIR_BUF];
    long long start[SPW];
    long long seq[SPW];
    long long kout[SPW];
    int nk[SPW];
    int gpre[SPW + 1];
    int pad[3];
    unsigned wcnt[32];              // MODE_PARTITION: per-owner counts of one warp iteration
    unsigned wpre[32];              // ... their exclusive prefix (slot of the owner's run in the staging area)
    uint64_t* dst[PAIR_BUF];        // ... destination address of every staged hash
    unsigned long long wbase[32];   // ... and the reserved base positions in the owners' regions
};

template <int K, int MODE>
__global__ void __launch_bounds__(NT, (MODE == MODE_LABEL || MODE == MODE_LOOKUP || MODE == MODE_LABEL_SIDE) ? SGC_PROBE_MIN_CTAS : 1) tile_kernel(const TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WPS = Geo<K>::WPS;
    const int lane = threadIdx.x & 31;
    WarpTile<K>& wt = reinterpret_cast<WarpTile<K>*>(smem_raw)[threadIdx.x >> 5];
    const int k = K ? K : a.k;
    const int64_t nsegs = *a.nsegs_p;
    const int64_t nrounds = (nsegs + SPW - 1) / SPW;
    unsigned long long my_hits = 0, my_miss = 0;
    unsigned pair_fill = 0;  // warp-uniform: keys waiting in wt.pairs

    auto flush_pairs = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&a.scal->pair_count, (unsigned long long)pair_fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (unsigned j = lane; j < pair_fill; j += 32)
            if (base + j < a.pairs_cap) a.pairs[base + j] = wt.pairs[j];
        __syncwarp();
        pair_fill = 0;
    };

    // K4 candidate rule shared by MODE_LABEL and MODE_GATHER: a hit is a candidate (sequence, id) key
    // unless the previous k-mer of the same segment has the same id (known through a warp shuffle;
    // the first k-mer of a lane-0 group is conservatively a candidate).  All 32 lanes must call this.
    auto emit_candidates = [&](const uint32_t (&ids)[4], bool active, int sg, int i, int nv) {
        const uint32_t prev_lane = __shfl_up_sync(0xffffffffu, ids[3], 1);
        unsigned long long keys[4] = {0, 0, 0, 0};
        unsigned nc = 0;
        if (active) {
            const unsigned long long sq = (unsigned long long)wt.seq[sg] << 32;
            uint32_t prev = (i > 0 && lane > 0) ? prev_lane : SGC_MISS;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (t < nv && ids[t] != SGC_MISS && ids[t] != prev) {
#pragma unroll
                    for (int u = 0; u < 4; u++)  // constant indices keep keys[] in registers
                        if (u == (int)nc) keys[u] = sq | ids[t];
                    nc++;
                }
                prev = ids[t];
            }
        }
        unsigned incl = nc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            if (pair_fill + total > PAIR_BUF) flush_pairs();
            unsigned pos = pair_fill + incl - nc;
#pragma unroll
            for (unsigned j = 0; j < 4; j++)
                if (j < nc) wt.pairs[pos++] = keys[j];
            pair_fill += total;
        }
    };

    // MODE_LABEL_SIDE: per-warp queue of k-mers that still need their own table probe
    unsigned long long* qh = nullptr;
    uint32_t* qs = nullptr;
    unsigned q_fill = 0;  // warp-uniform
    if constexpr (MODE == MODE_LABEL_SIDE) {
        unsigned char* qbase = smem_raw + sizeof(WarpTile<K>) * WPB;
        qh = reinterpret_cast<unsigned long long*>(qbase) + (threadIdx.x >> 5) * QCAP;
        qs = reinterpret_cast<uint32_t*>(qbase + (size_t)WPB * QCAP * 8) + (threadIdx.x >> 5) * QCAP;
    }
    // probe the queued k-mers densely: one item per lane, probe of item j+32 in flight while item j is
    // resolved; hits become (sequence, id) keys right away (no neighbour to compare with here, except
    // the previous lane's item)
    auto drain_queue = [&]() {
        if constexpr (MODE == MODE_LABEL_SIDE) {
            bool d_act = false;
            uint64_t d_h = 0;
            uint32_t d_sq = 0, d_b = 0;
            Entry d_e0, d_e1;
            auto settle = [&]() {
                uint32_t id = SGC_MISS;
                if (d_act) {
                    uint32_t v, aux;
                    if (bucket_match(d_e0, d_e1, d_h, v)) id = v;
                    else if (bucket_overflowed(d_e0)) id = lookup_slow(a.tbl, a.nlines, d_b, d_h, aux);
                    if (id != SGC_MISS) my_hits++; else my_miss++;
                }
                const unsigned long long key = ((unsigned long long)d_sq << 32) | id;
                const unsigned long long prevk = __shfl_up_sync(0xffffffffu, key, 1);
                const bool has = d_act && id != SGC_MISS && !(lane > 0 && prevk == key);
                const unsigned m = __ballot_sync(0xffffffffu, has);
                if (m) {
                    const unsigned total = __popc(m);
                    if (pair_fill + total > PAIR_BUF) flush_pairs();
                    if (has) wt.pairs[pair_fill + __popc(m & ((1u << lane) - 1u))] = key;
                    pair_fill += total;
                }
            };
            for (unsigned j0 = 0; j0 < q_fill; j0 += 32) {
                const unsigned j = j0 + lane;
                const bool act = j < q_fill;
                const uint64_t h = act ? qh[j] : 0ULL;
                const uint32_t sq = act ? qs[j] : 0u;
                if (j0 > 0) settle();
                d_act = act; d_h = h; d_sq = sq;
                d_b = act ? home_bucket(h, a.nbuckets, a.home_shift) : 0u;
                load_bucket_nc(a.tbl + d_b, d_e0, d_e1);
            }
            if (q_fill) settle();
            __syncwarp();
            q_fill = 0;
        }
    };

    // dynamic round scheduling; the NEXT round index is always requested one round ahead
    unsigned next_round = 0;
    if (lane == 0) next_round = atomicAdd(&a.scal->round_ctr, 1u);

    for (;;) {
        const unsigned round = __shfl_sync(0xffffffffu, next_round, 0);
        if ((int64_t)round >= nrounds) break;
        if (lane == 0) next_round = atomicAdd(&a.scal->round_ctr, 1u);
        const int64_t seg0 = (int64_t)round * SPW;

        // ---- descriptors (lanes 0..SPW-1) ----
        {
            int nk = 0;
            long long start = 0, sq = 0, kout = 0;
            const int64_t g = seg0 + lane;
            if (lane < SPW && g < nsegs) {
                const unsigned long long e = (unsigned long long)a.seg_map[g];
                sq = (long long)(e & SEQ_MASK);
                const long long sn = (long long)(e >> 40);
                const long long o0 = a.off[sq], o1 = a.off[sq + 1];
                const long long q0 = sn * SEG_K;
                const long long rem = o1 - o0 - k + 1 - q0;
                nk = (int)(rem < SEG_K ? rem : SEG_K);
                start = o0 + q0;
                if (a.kmer_off) kout = a.kmer_off[sq] + q0;
            }
            int incl = (nk + 3) >> 2;
#pragma unroll
            for (int d = 1; d < SPW; d <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane < SPW) {
                wt.start[lane] = start;
                wt.seq[lane] = sq;
                wt.kout[lane] = kout;
                wt.nk[lane] = nk;
                wt.gpre[lane + 1] = incl;
            }
            if (lane == 0) wt.gpre[0] = 0;
        }
        __syncwarp();

        // ---- stage forward + reverse-complement words straight from global memory ----
        if constexpr (MODE != MODE_GATHER) {
            const int sg = lane >> 2, sub = lane & 3;
            const int nk = wt.nk[sg];
            if (nk) {
                const int Ls = nk + k - 1;
                const uintptr_t base = (uintptr_t)a.seq + (uintptr_t)wt.start[sg];
                uint32_t* Fs = wt.F + sg * WPS;
                uint32_t* Rs = wt.R + sg * WPS;
                const int p = sgc::rc_pad(nk);
                bool all_ok = true;
                for (int w = sub; 4 * w < Ls; w += 4) {
                    uint32_t mask;
                    Fs[w] = load_window4(base, 4 * w, Ls, mask);
                }
                // R word m holds R[4m-4-p .. +4) = revcomp of F[Ls-4m+p .. +4)
                const int nrw = (Ls + 4 + p + 3) >> 2;
                for (int m = sub; m < nrw; m += 4) {
                    uint32_t mask;
                    const uint32_t u = load_window4(base, Ls - 4 * m + p, Ls, mask);
                    bool ok;
                    Rs[m] = revcomp_word(u, mask, ok);
                    all_ok &= ok;
                }
                if (!all_ok && a.status) atomicOr(&a.status[wt.seq[sg]], 1);
            }
        }
        __syncwarp();

        // ---- hash + sink ----
        const int G = wt.gpre[SPW];

        // locate group g (segment sg, first k-mer i, nv valid k-mers) and hash its 4 k-mers
        auto locate_and_hash = [&](int g, int& sg, int& i, int& nv, uint64_t (&h)[4]) {
            int lo = 0;  // largest sg with gpre[sg] <= g
#pragma unroll
            for (int step = SPW / 2; step > 0; step >>= 1)
                if (wt.gpre[lo + step] <= g) lo += step;
            sg = lo;
            i = (g - wt.gpre[sg]) << 2;
            const int nk = wt.nk[sg];
            nv = nk - i < 4 ? nk - i : 4;
            const uint32_t* Fs = wt.F + sg * WPS;
            const uint32_t* Rs = wt.R + sg * WPS;
            if constexpr (K != 0) {
                constexpr int NWIN = sgc::KW<K ? K : 1>::NWIN;
                uint32_t fw[NWIN], rw[NWIN];
                const uint32_t* fp = Fs + (i >> 2);
                const uint32_t* rp = Rs + sgc::rc_group_word(nk, i);
#pragma unroll
                for (int j = 0; j < NWIN; j++) {
                    fw[j] = fp[j];
                    rw[j] = rp[j];
                }
#pragma unroll
                for (int t = 0; t < 4; t++) h[t] = sgc::hash_group_kmer<K ? K : 1>(fw, rw, t);
            } else {
                for (int t = 0; t < nv; t++) {
                    int q = i + t;
                    h[t] = sgc::murmur3_h1_stream(Fs, q, k) ^
                           sgc::murmur3_h1_stream(Rs, sgc::rc_byte_of(nk, nk - 1 - q), k);
                }
            }
        };

        if constexpr (MODE == MODE_MINHASH) {
            // sourmash-style hash: Murmur3(min(kmer, revcomp), seed); all hashes and/or the per-sequence
            // minimum (MinHash n=1 of cdbg/sort_bcalm_unitigs.py:56-128)
            for (int g0 = 0; g0 < G; g0 += 32) {
                const int g = g0 + lane;
                if (g >= G) continue;
                int lo = 0;
#pragma unroll
                for (int step = SPW / 2; step > 0; step >>= 1)
                    if (wt.gpre[lo + step] <= g) lo += step;
                const int sg = lo, i = (g - wt.gpre[sg]) << 2;
                const int nk = wt.nk[sg];
                const int nv = nk - i < 4 ? nk - i : 4;
                const uint32_t* Fs = wt.F + sg * WPS;
                const uint32_t* Rs = wt.R + sg * WPS;
                uint64_t h[4] = {~0ULL, ~0ULL, ~0ULL, ~0ULL};
                if constexpr (K != 0) {
                    constexpr int NWIN = sgc::KW<K ? K : 1>::NWIN;
                    uint32_t fw[NWIN], rw[NWIN];
                    const uint32_t* fp = Fs + (i >> 2);
                    const uint32_t* rp = Rs + sgc::rc_group_word(nk, i);
#pragma unroll
                    for (int j = 0; j < NWIN; j++) {
                        fw[j] = fp[j];
                        rw[j] = rp[j];
                    }
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < nv) h[t] = sgc::minhash_group_kmer<K ? K : 1>(fw, rw, t, a.seed);
                } else {
                    for (int t = 0; t < nv; t++) {
                        const int q = i + t, rq = sgc::rc_byte_of(nk, nk - 1 - q);
                        h[t] = sgc::stream_less(Rs, rq, Fs, q, k) ? sgc::murmur3_h1_stream(Rs, rq, k, a.seed)
                                                                  : sgc::murmur3_h1_stream(Fs, q, k, a.seed);
                    }
                }
                if (a.hash_out) {
                    uint64_t* o = a.hash_out + wt.kout[sg] + i;
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < nv) o[t] = h[t];
                }
                if (a.seq_min) {
                    uint64_t m = h[0];
#pragma unroll
                    for (int t = 1; t < 4; t++) m = h[t] < m ? h[t] : m;
                    atomicMin(&a.seq_min[wt.seq[sg]], (unsigned long long)m);
                }
            }
        } else if constexpr (MODE == MODE_HASH || MODE == MODE_INSERT) {
            for (int g0 = 0; g0 < G; g0 += 32) {
                const int g = g0 + lane;
                if (g >= G) continue;
                int sg, i, nv;
                uint64_t h[4] = {0, 0, 0, 0};
                locate_and_hash(g, sg, i, nv, h);
                if constexpr (MODE == MODE_HASH) {
                    uint64_t* o = a.hash_out + wt.kout[sg] + i;
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < nv) o[t] = h[t];
                } else {
                    long long sq = wt.seq[sg];
                    uint32_t id = a.seq_ids ? a.seq_ids[sq] : a.id_base + (uint32_t)sq;
                    if (id > SGC_MAX_ID) atomicOr(&a.scal->err, 2);
                    else {
                        const unsigned long long g0 = a.side_base + (unsigned long long)(wt.kout[sg] + i);
                        for (int t = 0; t < nv; t++) {
                            uint32_t aux = AUX_NONE;
                            if (a.side) {  // unitig-ordered record; its id is fixed up after the build
                                aux = (uint32_t)(g0 + t);
                                a.side[g0 + t] = make_uint4((uint32_t)h[t], (uint32_t)(h[t] >> 32), id, 0u);
                            }
                            table_insert(a.tbl, a.nbuckets, a.nlines, a.home_shift, h[t], id, aux, a.scal);
                        }
                    }
                }
            }
        } else if constexpr (MODE == MODE_PARTITION) {
            // K7 request side: every hash goes to the send region of its owner rank (top owner_bits
            // bits).  Positions are reserved per warp iteration (<=128 k-mers): lanes count owners in
            // shared memory, one lane per owner reserves a run in that owner's region, so each owner
            // receives runs of consecutive hashes.  ridx remembers where every k-mer's reply will be.
            const int sh = 32 - a.owner_bits;
            for (int g0 = 0; g0 < G; g0 += 32) {
                const int g = g0 + lane;
                const bool active = g < G;
                int sg = 0, i = 0, nv = 0;
                uint64_t h[4] = {0, 0, 0, 0};
                if (active) locate_and_hash(g, sg, i, nv, h);
                if (lane < a.n_ranks) wt.wcnt[lane] = 0;
                __syncwarp();
                uint32_t own[4], rank[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    own[t] = a.owner_bits ? (uint32_t)(h[t] >> (64 - a.owner_bits)) : 0u;
                    rank[t] = (active && t < nv) ? atomicAdd(&wt.wcnt[own[t]], 1u) : 0u;
                }
                __syncwarp();
                {   // one lane per owner: reserve the run in the owner's region, prefix for the staging slots
                    const unsigned c = lane < a.n_ranks ? wt.wcnt[lane] : 0u;
                    unsigned incl = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    if (lane < a.n_ranks) {
                        wt.wpre[lane] = incl - c;
                        wt.wbase[lane] = c ? atomicAdd(&a.cursor[lane], (unsigned long long)c) : 0ULL;
                    }
                }
                __syncwarp();
                // Stage the <=128 hashes of this iteration grouped by owner (reusing the idle label
                // staging area: 128 hashes + 128 destination addresses), then copy them out with
                // consecutive lanes on consecutive addresses: each owner gets its run as full-width
                // coalesced stores -- what NVLink wants when the destination is a peer GPU's inbox.
                unsigned long long* stage_h = wt.pairs;                       // PAIR_BUF = 128 entries
                if (active) {
                    const long long ko = wt.kout[sg] + i;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        if (t < nv) {
                            const unsigned long long pos = wt.wbase[own[t]] + rank[t];
                            a.ridx[ko + t] = (a.owner_bits ? (own[t] << sh) : 0u) | (uint32_t)pos;
                        }
                    }
                }
                __syncwarp();
                if (a.n_ranks <= 2) {
                    // few owners: runs are long enough for direct stores (measured faster at R = 2)
                    if (active) {
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            if (t < nv) {
                                const unsigned long long pos = wt.wbase[own[t]] + rank[t];
                                if (pos < a.send_cap) a.peer_send[own[t]][a.send_region_off + pos] = h[t];
                            }
                        }
                    }
                } else {
                    if (active) {
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            if (t < nv) {
                                const unsigned slot = wt.wpre[own[t]] + rank[t];
                                const unsigned long long pos = wt.wbase[own[t]] + rank[t];
                                stage_h[slot] = h[t];
                                wt.dst[slot] = pos < a.send_cap ? a.peer_send[own[t]] + a.send_region_off + pos : nullptr;
                            }
                        }
                    }
                    __syncwarp();
                    const unsigned total = wt.wpre[a.n_ranks - 1] + wt.wcnt[a.n_ranks - 1];
                    for (unsigned j = lane; j < total; j += 32) {
                        uint64_t* d = wt.dst[j];
                        if (d) *d = stage_h[j];
                    }
                }
                __syncwarp();
            }
        } else if constexpr (MODE == MODE_GATHER) {
            // K7 reply side: ids come from the owners' replies instead of local probes; same K4 rule
            const int sh = 32 - a.owner_bits;
            const uint32_t pmask = a.owner_bits ? ((1u << sh) - 1u) : 0xFFFFFFFFu;
            for (int g0 = 0; g0 < G; g0 += 32) {
                const int g = g0 + lane;
                const bool active = g < G;
                int sg = 0, i = 0, nv = 0;
                uint32_t ids[4] = {SGC_MISS, SGC_MISS, SGC_MISS, SGC_MISS};
                if (active) {
                    int lo = 0;
#pragma unroll
                    for (int step = SPW / 2; step > 0; step >>= 1)
                        if (wt.gpre[lo + step] <= g) lo += step;
                    sg = lo;
                    i = (g - wt.gpre[sg]) << 2;
                    const int nk = wt.nk[sg];
                    nv = nk - i < 4 ? nk - i : 4;
                    const long long ko = wt.kout[sg] + i;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        if (t < nv) {
                            const uint32_t r = a.ridx[ko + t];
                            const unsigned long long own = a.owner_bits ? (r >> sh) : 0u;
                            ids[t] = a.replies[own * a.send_cap + (r & pmask)];
                            if (ids[t] != SGC_MISS) my_hits++; else my_miss++;
                        }
                    }
                }
                if (a.kmer_ids && active) {
                    uint32_t* o = a.kmer_ids + wt.kout[sg] + i;
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < nv) o[t] = ids[t];
                }
                emit_candidates(ids, active, sg, i, nv);
            }
        } else if constexpr (MODE == MODE_LABEL_SIDE) {
            // K4 with fewer HBM lines per k-mer.  Only the FIRST k-mer of a group (the anchor) probes the
            // hash table (software-pipelined like MODE_LABEL).  Its slot carries g, the k-mer's position in
            // the table's unitig-ordered side array of {hash, id} records (id = what a table lookup of that
            // hash returns, fixed up after the build).  Consecutive k-mers of a read are consecutive
            // k-mers of the unitig, on one strand or the other, so k-mer t is looked for at side[g+t] and
            // side[g-t]: a full 64-bit hash match there IS the table's answer for that hash.  Anything
            // not matched (sequencing errors, unitig ends, anchor missing) is queued per warp and probed
            // later DENSELY (every lane busy, probes pipelined), so the result is exactly what per-k-mer
            // probing gives.  Neighbouring threads share side lines.
            bool p_active = false;
            int p_sg = 0, p_i = 0, p_nv = 0;
            uint64_t p_h[4] = {0, 0, 0, 0};
            uint32_t p_b0 = 0;
            Entry p_e0, p_e1;

            auto resolve = [&]() {
                uint32_t ids[4] = {SGC_MISS, SGC_MISS, SGC_MISS, SGC_MISS};
                unsigned pending = 0;
                if (p_active) {
                    uint32_t v = SGC_MISS, g = AUX_NONE;
                    if (!bucket_match(p_e0, p_e1, p_h[0], v, g)) {
                        v = SGC_MISS;
                        g = AUX_NONE;
                        if (bucket_overflowed(p_e0)) v = lookup_slow(a.tbl, a.nlines, p_b0, p_h[0], g);
                    }
                    ids[0] = v;
                    pending = ((1u << p_nv) - 1u) & ~1u;  // k-mers 1..nv-1 until proven otherwise
                    if (g != AUX_NONE) {
                        uint4 up[3], um[3];
#pragma unroll
                        for (int d = 1; d < 4; d++) {
                            const bool hi_ok = d < p_nv && (unsigned long long)g + d < a.side_n;
                            const bool lo_ok = d < p_nv && g >= (uint32_t)d;
                            up[d - 1] = hi_ok ? __ldg(a.side + g + d) : make_uint4(0, 0, SGC_MISS, 1);
                            um[d - 1] = lo_ok ? __ldg(a.side + g - d) : make_uint4(0, 0, SGC_MISS, 1);
                        }
#pragma unroll
                        for (int d = 1; d < 4; d++) {
                            if (d < p_nv) {
                                const uint32_t lo = (uint32_t)p_h[d], hi = (uint32_t)(p_h[d] >> 32);
                                if (up[d - 1].x == lo && up[d - 1].y == hi && up[d - 1].w == 0) {
                                    ids[d] = up[d - 1].z;
                                    pending &= ~(1u << d);
                                } else if (um[d - 1].x == lo && um[d - 1].y == hi && um[d - 1].w == 0) {
                                    ids[d] = um[d - 1].z;
                                    pending &= ~(1u << d);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < p_nv && !((pending >> t) & 1u)) { if (ids[t] != SGC_MISS) my_hits++; else my_miss++; }
                }
                // queue the unresolved k-mers of this warp iteration (at most 96; room was made before)
                {
                    const unsigned np = __popc(pending);
                    unsigned incl = np;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total) {
                        unsigned pos = q_fill + incl - np;
                        const uint32_t sq = pending ? (uint32_t)wt.seq[p_sg] : 0u;
#pragma unroll
                        for (int d = 1; d < 4; d++) {
                            if ((pending >> d) & 1u) {
                                qh[pos] = p_h[d];
                                qs[pos] = sq;
                                pos++;
                            }
                        }
                        q_fill += total;
                    }
                }
                emit_candidates(ids, p_active, p_sg, p_i, p_nv);  // queued k-mers count as "no id" here
            };

            for (int g0 = 0; g0 < G; g0 += 32) {
                const int g = g0 + lane;
                const bool active = g < G;
                int sg = 0, i = 0, nv = 0;
                uint64_t h[4] = {0, 0, 0, 0};
                if (active) locate_and_hash(g, sg, i, nv, h);   // the previous group's anchor probe is in flight here
                {
                    const uint64_t dep = (h[0] ^ h[1] ^ h[2] ^ h[3]) & a.zero;  // see MODE_LABEL
#pragma unroll
                    for (int t = 0; t < 4; t++) p_h[t] ^= dep;
                }
                if (g0 > 0) resolve();
                if (q_fill > QCAP - 96) drain_queue();
                p_active = active; p_sg = sg; p_i = i; p_nv = nv;
#pragma unroll
                for (int t = 0; t < 4; t++) p_h[t] = h[t];
                p_b0 = active ? home_bucket(h[0], a.nbuckets, a.home_shift) : 0u;
                load_bucket_nc(a.tbl + p_b0, p_e0, p_e1);
            }
            if (G > 0) resolve();
            if (q_fill > QCAP - 96) drain_queue();
            __syncwarp();
        } else {
            // MODE_LOOKUP / MODE_LABEL, software-pipelined: the 4 table probes of a group are issued
            // and left in flight while the NEXT group's 8 Murmur3 evaluations run; they are resolved
            // one iteration later.  Integer pipe and HBM latency overlap inside a thread.
            bool p_active = false;
            int p_sg = 0, p_i = 0, p_nv = 0;
            uint64_t p_h[4] = {0, 0, 0, 0};
            uint32_t p_b[4] = {0, 0, 0, 0};
            Entry p_e0[4], p_e1[4];

            auto resolve = [&]() {
                uint32_t ids[4] = {SGC_MISS, SGC_MISS, SGC_MISS, SGC_MISS};
                unsigned slow = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (p_active && t < p_nv) {
                        uint32_t v;
                        if (bucket_match(p_e0[t], p_e1[t], p_h[t], v)) ids[t] = v;
                        else if (bucket_overflowed(p_e0[t])) slow |= 1u << t;
                    }
                }
                while (slow) {  // ~6 % of lookups: the key is not in its home bucket; one trip per key
                    const int t = __ffs(slow) - 1;
                    slow &= slow - 1;
                    const uint64_t hk = t == 0 ? p_h[0] : t == 1 ? p_h[1] : t == 2 ? p_h[2] : p_h[3];
                    const uint32_t hb = t == 0 ? p_b[0] : t == 1 ? p_b[1] : t == 2 ? p_b[2] : p_b[3];
                    uint32_t aux_unused;
                    const uint32_t v = lookup_slow(a.tbl, a.nlines, hb, hk, aux_unused);
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (u == t) ids[u] = v;
                }
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (p_active && t < p_nv) { if (ids[t] != SGC_MISS) my_hits++; else my_miss++; }

                if constexpr (MODE == MODE_LOOKUP) {
                    if (p_active) {
                        if (a.kmer_ids) {
                            uint32_t* o = a.kmer_ids + wt.kout[p_sg] + p_i;
#pragma unroll
                            for (int t = 0; t < 4; t++)
                                if (t < p_nv) o[t] = ids[t];
                        }
                        if (a.count_by_id) {
                            // run-length aggregate within the group before touching L2
#pragma unroll
                            for (int t = 0; t < 4; t++) {
                                if (t < p_nv && ids[t] != SGC_MISS) {
                                    unsigned run = 1;
                                    bool counted = false;
#pragma unroll
                                    for (int u = 0; u < 4; u++) {
                                        if (u < t && ids[u] == ids[t]) counted = true;
                                        if (u > t && u < p_nv && ids[u] == ids[t]) run++;
                                    }
                                    if (!counted) atomicAdd(&a.count_by_id[ids[t]], run);
                                }
                            }
                        }
                    }
                } else {
                    emit_candidates(ids, p_active, p_sg, p_i, p_nv);
                }
            };

            for (int g0 = 0; g0 < G; g0 += 32) {
                const int g = g0 + lane;
                const bool active = g < G;
                int sg = 0, i = 0, nv = 0;
                uint64_t h[4] = {0, 0, 0, 0};
                if (active) locate_and_hash(g, sg, i, nv, h);   // previous group's probes are in flight here
                // Order the resolve AFTER the hash computation in the instruction stream: a.zero is
                // always 0 but only known at run time, so the compares below truly depend on the new
                // hashes and neither nvcc nor ptxas can sink the Murmur3 code under the resolve.
                {
                    const uint64_t dep = (h[0] ^ h[1] ^ h[2] ^ h[3]) & a.zero;
#pragma unroll
                    for (int t = 0; t < 4; t++) p_h[t] ^= dep;
                }
                if (g0 > 0) resolve();
                p_active = active; p_sg = sg; p_i = i; p_nv = nv;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    p_h[t] = h[t];
                    // lanes / k-mers with nothing to look up probe bucket 0 (an L2 hit) so that the
                    // loads are unconditional and land directly in the pipeline registers
                    p_b[t] = (active && t < nv) ? home_bucket(h[t], a.nbuckets, a.home_shift) : 0u;
                    load_bucket_nc(a.tbl + p_b[t], p_e0[t], p_e1[t]);
                }
            }
            if (G > 0) resolve();
            __syncwarp();
        }
    }

    if constexpr (MODE == MODE_LABEL_SIDE) drain_queue();
    if constexpr (MODE == MODE_LABEL || MODE == MODE_GATHER || MODE == MODE_LABEL_SIDE) {
        if (pair_fill) flush_pairs();
    }
    if constexpr (MODE == MODE_LOOKUP || MODE == MODE_LABEL || MODE == MODE_GATHER || MODE == MODE_LABEL_SIDE) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            my_hits += __shfl_xor_sync(0xffffffffu, my_hits, d);
            my_miss += __shfl_xor_sync(0xffffffffu, my_miss, d);
        }
        if (lane == 0) {
            if (my_hits) atomicAdd(&a.scal->n_hits, my_hits);
            if (my_miss) atomicAdd(&a.scal->n_miss, my_miss);
        }
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
                                    Scalars* scal) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = id[i];
        if (v > SGC_MAX_ID) atomicOr(&scal->err, 2);
        else table_insert(tbl, nb, nlines, hs, hash[i], v, AUX_NONE, scal);
    }
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

// side[g].id := what a table lookup of side[g].hash returns (MISS for dropped hashes), so that a
// hash match against a side record is by construction the table's own answer
__global__ void side_fixup_kernel(const Bucket* tbl, uint32_t nb, uint32_t nlines, int hs, uint4* side, uint64_t n) {
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (uint64_t)gridDim.x * blockDim.x) {
        uint4 r = side[g];
        const uint64_t h = ((uint64_t)r.y << 32) | r.x;
        r.z = table_lookup(tbl, nb, nlines, hs, h);
        r.w = 0;
        side[g] = r;
    }
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

}  // namespace

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
    // unitig-ordered side array of {hash, id} records (tables built from sequences)
    uint4* side = nullptr;
    int64_t side_cap = 0, side_fill = 0;
    bool side_dirty = false;
    // pending label call
    bool label_pending = false;
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
    const int64_t* seg_map;
    const int64_t* nsegs_p;
};

int make_plan(sgc_ctx* c, const int64_t* d_off, int64_t nseq, int64_t seq_bytes, int k, int64_t* d_kmer_off,
              Plan* plan) {
    const size_t n1 = (size_t)nseq + 1;
    TRY(ensure(c, c->nk, n1 * 8));
    TRY(ensure(c, c->nseg, n1 * 8));
    TRY(ensure(c, c->seg_first, n1 * 8));
    if (!d_kmer_off) {
        TRY(ensure(c, c->kmer_off, n1 * 8));
        d_kmer_off = (int64_t*)c->kmer_off.p;
    }
    const int64_t max_segs = nseq + seq_bytes / SEG_K + 1;
    TRY(ensure(c, c->seg_map, (size_t)max_segs * 8));
    CU(cudaMemsetAsync(&c->d_scal->bad_offsets, 0, sizeof(int), c->stream));
    plan_count_kernel<<<grid_for((int64_t)n1, 256, 1 << 30), 256, 0, c->stream>>>(
        d_off, nseq, k, seq_bytes, (int64_t*)c->nk.p, (int64_t*)c->nseg.p, c->d_scal);
    c->launches++;
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t*)c->nk.p, d_kmer_off, (int64_t)n1, c->stream));
    TRY(ensure(c, c->cub_tmp, tb));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->nk.p, d_kmer_off, (int64_t)n1, c->stream));
    CU(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tb, (int64_t*)c->nseg.p, (int64_t*)c->seg_first.p, (int64_t)n1,
                                     c->stream));
    if (nseq > 0) {
        plan_expand_kernel<<<grid_for(nseq, 256, 1 << 30), 256, 0, c->stream>>>((int64_t*)c->seg_first.p, nseq,
                                                                              (int64_t*)c->seg_map.p);
        c->launches++;
    }
    CU(cudaGetLastError());
    plan->kmer_off = d_kmer_off;
    plan->seg_first = (int64_t*)c->seg_first.p;
    plan->seg_map = (int64_t*)c->seg_map.p;
    plan->nsegs_p = (int64_t*)c->seg_first.p + nseq;
    return SGC_OK;
}

template <int K, int MODE>
int launch_tile_k(sgc_ctx* c, const TileArgs& a) {
    size_t smem = sizeof(WarpTile<K>) * WPB + (MODE == MODE_LABEL_SIDE ? (size_t)WPB * QCAP * 12 : 0);
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
    a.seg_map = p.seg_map;
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

// sorted-unique the candidate keys in pairs_a[0..n) and build the CSR.  Synchronises.
int pairs_to_csr(sgc_ctx* c, int64_t n, int64_t nseq, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap,
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
                      &c->pairs_a, &c->pairs_b, &c->tmp_hash, &c->tmp_id, &c->tmp_idx, &c->cursor};
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
int sgc_hash_batch(sgc_ctx* c, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                   uint64_t* d_hash_out, int64_t hash_cap, int64_t* d_kmer_off, int32_t* d_status,
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
    return launch_tile<MODE_HASH>(c, a);
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

// ---------------------------------------------------------------------------- K2
int sgc_table_create(sgc_ctx* c, int64_t max_entries, float load, sgc_table** out) {
    if (!c || !out) return fail(SGC_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (max_entries < 0) return fail(SGC_ERR_ARG, "max_entries < 0");
    TRY(set_device(c));
    // 32-byte buckets of two 16-byte slots, four per 128-byte line; `load` = expected hashes per
    // bucket (default 0.75: ~4 % of lookups leave their home bucket, still inside the home line)
    if (!(load > 0.0f)) load = 0.75f;
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
    cudaError_t e = cudaMalloc(&t->buckets, (size_t)nb * sizeof(Bucket));
    if (e != cudaSuccess) {
        delete t;
        return fail(SGC_ERR_NOMEM, "cudaMalloc(%zu) for the k-mer table failed: %s", (size_t)nb * sizeof(Bucket),
                    cudaGetErrorString(e));
    }
    table_clear_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(t->buckets, t->nbuckets);
    c->launches++;
    CU(cudaGetLastError());
    *out = t;
    return SGC_OK;
}

void sgc_table_free(sgc_table* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    if (t->buckets) cudaFree(t->buckets);
    if (t->side) cudaFree(t->side);
    delete t;
}

int64_t sgc_table_bytes(const sgc_table* t) { return t ? (int64_t)t->nbuckets * (int64_t)sizeof(Bucket) : 0; }

int sgc_table_enable_side(sgc_table* t) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (t->side) return SGC_OK;
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    const int64_t cap = std::max<int64_t>(t->max_entries, 64);
    if (cap >= (int64_t)AUX_NONE) return fail(SGC_ERR_CAPACITY, "side array limited to 2^31-2 records");
    cudaError_t e = cudaMalloc(&t->side, (size_t)cap * sizeof(uint4));
    if (e != cudaSuccess) {
        t->side = nullptr;
        return fail(SGC_ERR_NOMEM, "cudaMalloc(%zu) for the side array failed: %s", (size_t)cap * sizeof(uint4),
                    cudaGetErrorString(e));
    }
    t->side_cap = cap;
    t->side_fill = 0;
    return SGC_OK;
}

int64_t sgc_table_side_bytes(const sgc_table* t) { return t && t->side ? t->side_cap * (int64_t)sizeof(uint4) : 0; }

int sgc_table_set_home_shift(sgc_table* t, int shift) {
    if (!t || shift < 0 || shift > 16) return fail(SGC_ERR_ARG, "bad home shift");
    t->home_shift = shift;
    return SGC_OK;
}

int sgc_table_insert_pairs(sgc_table* t, const uint64_t* d_hash, const uint32_t* d_id, int64_t n) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    if (n < 0) return fail(SGC_ERR_ARG, "n < 0");
    if (n == 0) return SGC_OK;
    sgc_ctx* c = t->ctx;
    TRY(set_device(c));
    TRY(zero_scalars(c));
    insert_pairs_kernel<<<grid_for(n, 256, c->num_sms * 8), 256, 0, c->stream>>>(t->buckets, t->nbuckets, t->nlines,
                                                                               t->home_shift, d_hash, d_id, n, c->d_scal);
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
    int64_t total = 0;
    if (t->side) {
        CU(cudaMemcpyAsync(&c->h_scal->n_unique, p.kmer_off + nseq, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        total = c->h_scal->n_unique;
        if (t->side_fill + total > t->side_cap || t->side_fill + total >= (int64_t)AUX_NONE)
            return fail(SGC_ERR_CAPACITY, "side array of %lld records cannot take %lld more", (long long)t->side_cap,
                        (long long)total);
    }
    TRY(zero_scalars(c));
    TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
    table_args(a, t);
    a.seq_ids = d_ids;
    a.id_base = id_base;
    a.side = t->side;
    a.side_base = (unsigned long long)t->side_fill;
    TRY(launch_tile<MODE_INSERT>(c, a));
    TRY(fetch_scalars(c));
    if (t->side) {
        t->side_fill += total;
        t->side_dirty = true;
    }
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
int sgc_label_reads_launch(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq,
                           int k, int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap, int32_t* d_status) {
    if (!t) return fail(SGC_ERR_ARG, "table is NULL");
    sgc_ctx* c = t->ctx;
    TRY(check_seq_args(d_seq, seq_bytes, d_off, nseq, k));
    if (!d_ids_off) return fail(SGC_ERR_ARG, "d_ids_off is NULL");
    if (ids_cap < 0) return fail(SGC_ERR_ARG, "ids_cap < 0");
    TRY(set_device(c));
    if (d_status) CU(cudaMemsetAsync(d_status, 0, (size_t)nseq * 4, c->stream));
    Plan p;
    TRY(make_plan(c, d_off, nseq, seq_bytes, k, nullptr, &p));
    // candidate (sequence, id) keys: a few per read in practice; start from ids_cap and a floor
    unsigned long long pairs_cap = (unsigned long long)std::max<int64_t>(std::max<int64_t>(ids_cap * 2, nseq * 4), 4096);
    TRY(ensure(c, c->pairs_a, (size_t)pairs_cap * 8));
    pairs_cap = c->pairs_a.cap / 8;
    TRY(zero_scalars(c));
    if (nseq > 0) {
        TileArgs a = base_args(c, d_seq, d_off, k, p, d_status);
        table_args(a, t);
        a.pairs = (unsigned long long*)c->pairs_a.p;
        a.pairs_cap = pairs_cap;
        if (t->side && t->side_fill > 0 && !getenv("SGC_NO_SIDE")) {
            if (t->side_dirty) {
                side_fixup_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(t->buckets, t->nbuckets, t->nlines, t->home_shift,
                                                                        t->side, (uint64_t)t->side_fill);
                c->launches++;
                t->side_dirty = false;
            }
            a.side = t->side;
            a.side_n = (unsigned long long)t->side_fill;
            TRY(launch_tile<MODE_LABEL_SIDE>(c, a));
        } else {
            TRY(launch_tile<MODE_LABEL>(c, a));
        }
    }
    CU(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
    t->label_pending = true;
    t->label_nseq = nseq;
    t->label_ids_off = d_ids_off;
    t->label_ids_out = d_ids_out;
    t->label_ids_cap = ids_cap;
    t->label_pairs_cap = pairs_cap;
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
                        h_n_labels);
}

int sgc_label_reads(sgc_table* t, const uint8_t* d_seq, int64_t seq_bytes, const int64_t* d_off, int64_t nseq, int k,
                    int64_t* d_ids_off, uint32_t* d_ids_out, int64_t ids_cap, int32_t* d_status, int64_t* h_n_kmers,
                    int64_t* h_n_hits, int64_t* h_n_labels) {
    TRY(sgc_label_reads_launch(t, d_seq, seq_bytes, d_off, nseq, k, d_ids_off, d_ids_out, ids_cap, d_status));
    return sgc_label_reads_finish(t, h_n_kmers, h_n_hits, h_n_labels);
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
    for (int r = 0; r < n_ranks; r++) {
        h_counts[r] = (int64_t)counts[r];
        if ((int64_t)counts[r] > send_cap)
            return fail(SGC_ERR_CAPACITY, "send_cap %lld < %llu hashes for owner %d", (long long)send_cap, counts[r], r);
    }
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
