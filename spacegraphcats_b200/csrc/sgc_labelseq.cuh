// sgc_labelseq.cuh -- the read-labelling kernel (K1+K3+K4 fused: hash, table probe, per-read unitig ids)
// for tables that carry the unitig SEQUENCE STORE, and the kernels that build the store.  Part of
// libsgc_b200.so (device code; included by sgc_b200.cu only).  Reference semantics:
// cdbg/index_reads.py:143-152 (hash_sequence + count_cdbg_matches per read; only the KEYS of the returned
// dict are used).
//
// The idea: consecutive k-mers of a read are consecutive k-mers of a unitig, on either strand.  So only ONE
// k-mer in sixteen (the anchor of a group) is hashed and looked up in the table.  Its slot says where that
// k-mer lies in the table's store of unitig sequences (2 bits per base, every inserted sequence buffer
// concatenated) and on which strand; the fifteen followers are settled by comparing the read's own 2-bit
// bases with the store -- 46 bases, three 32-bit words, one XOR each.  A follower whose k bases equal the
// store's has, by construction, the hash of that store position, and the BREAK bits say whether a table
// lookup of that position's hash returns the anchor's id (brk[p] is set at the first k-mer of every
// sequence, at every position that holds no k-mer, at every k-mer whose hash the table keeps under another
// position, and at every dropped multicount hash -- set by the build, sgc_table.cuh:table_insert).  Whatever
// is not settled that way is queued as follow-up work of the same round: a k-mer that overlaps a mismatching
// base (a sequencing error) is hashed and probed on its own; the k-mers past a break (the read runs on into the
// next unitig) become a RUN of their own, anchored at the first k-mer past the break; and when the anchor itself
// is missing, the run is turned around once and anchored at its far end (an error early in the anchor leaves the
// last followers intact).  Every k-mer is thus either probed or proven equal to a store position whose lookup
// result is known: the labels are exactly those of per-k-mer probing (tile_kernel<K, MODE_LABEL>), with about a
// fifth of the hashes and of the table lines.
//
// Layout of the work
//   * a warp round is SPW segments of <= 128 k-mers (a 150-bp read is one segment); its work items are the
//     anchors of the round's groups followed by whatever they queue; 32 items per trip, every lane: locate,
//     hash (2 x Murmur3 over the staged ASCII words, any byte alignment), probe the home bucket, then
//     anchors: compare with the store and queue the unsettled followers; all: emit (sequence, id) keys.
//   * PACKED = true: the reads arrive 2 bits per base (A=0 C=1 G=2 T=3, base b of a record in bits
//     2(b%4).. of byte b/4, every record word aligned) and are expanded to the ASCII words the reference
//     hash is defined on while they are staged (PRMT on "ACGT"); ASCII input is packed the other way.
#pragma once
#include "sgc_kmer.cuh"
#include "sgc_table.cuh"
#include "sgc_tile.cuh"

namespace sgc {
namespace dev {

constexpr int SEG_KS = 128;   // k-mers per segment of the labelseq plan
#ifndef SGC_LS_GK
#define SGC_LS_GK 16
#endif
constexpr int GK = SGC_LS_GK; // k-mers per group: one anchor, GK - 1 followers
static_assert(GK == 16 || GK == 24 || GK == 32, "group size");
constexpr unsigned GKMASK = GK >= 32 ? 0xFFFFFFFFu : ((1u << (GK & 31)) - 1u);
constexpr int PAIRS_S = 64;   // candidate keys staged per warp before one global reservation
constexpr int FOFF = 4;       // staged forward words start 4 words into their slot
constexpr int USEQ_PAD = 8;   // words (128 bases) of zero padding on either end of the sequence store
#ifndef SGC_LS_RUN_MIN
#define SGC_LS_RUN_MIN 1
#endif
constexpr int RUN_MIN = SGC_LS_RUN_MIN;  // queued runs wait for company (or for the ranges to run out) before they take a trip
static_assert(USEQ_PAD == SGC_STORE_PAD_WORDS && BRK_PAD == SGC_BRK_PAD_BITS, "include/sgc_b200.h states the store layout");
#ifndef SGC_LS_NT
#define SGC_LS_NT 256
#endif
#ifndef SGC_LS_MIN_CTAS
#define SGC_LS_MIN_CTAS 4
#endif
#ifndef SGC_LS_SPW
#define SGC_LS_SPW 4
#endif
#ifndef SGC_LS_REANCHOR
#define SGC_LS_REANCHOR 1   // 0: every unsettled k-mer is probed on its own (no follow-up runs)
#endif
constexpr int NTS = SGC_LS_NT;    // threads per CTA of the labelseq kernel
constexpr int WPBS = NTS / 32;

template <int K, bool PACKED>
struct GeoS {
    static constexpr int KMAX = K ? K : SGC_MAX_K;
    static constexpr int SPW = K ? SGC_LS_SPW : 4;                 // segments per warp round
    static constexpr int LPS = 32 / SPW;                            // lanes staging one segment
    static constexpr int SEGB = SEG_KS + KMAX - 1;                  // bases of a full segment
    static constexpr int STRIDE = ((SEGB + 32 + 15) / 16) * 16;     // bytes per staged ASCII area
    static constexpr int WPS = STRIDE / 4;
    static constexpr int PWS = (SEGB + 15) / 16 + 4;                // 2-bit words per segment (+ slack for window reads)
    // words of the raw copy of a segment (packed: 16-byte chunks around a word-aligned start)
    static constexpr int RAWW = PACKED ? (((SEGB + 15) / 16 + 3 + 3) / 4) * 4 : WPS;
    static constexpr int QCAP = SPW * SEG_KS;                       // a round queues at most its own k-mers
    static_assert(32 % SPW == 0, "lanes per segment");
    static_assert(SPW <= 4 && SEG_KS <= 128 && GK <= 32, "a queued run is 16 bits: k-mer 7, segment 2, followers 5, back, flagged");
};

struct LabelSeqArgs {
    const uint8_t* seq;       // ASCII bases, or packed 2-bit records
    const SegDesc* segdesc;
    const int64_t* nsegs_p;
    int32_t* status;          // ASCII only (packed input is validated by the packer)
    Scalars* scal;
    int k;
    const Bucket* tbl;
    uint32_t nbuckets;
    uint32_t nlines;
    int home_shift;
    const uint32_t* useq;     // word of store position 0 (USEQ_PAD words of padding before it)
    const uint32_t* brk;      // bit BRK_PAD + p = break before position p
    unsigned long long* pairs;
    unsigned long long pairs_cap;
    unsigned long long* cnt;  // nullable: per-sequence candidate counters (zeroed by the caller), bumped as keys are flushed
    // PEER = true: the table is hash-range partitioned over n_shards GPUs (owner = top owner_bits bits of the hash)
    // and probed where it lies, over NVLink peer memory; the store is replicated, one shard per rank (ustride /
    // bstride words apart), and the shard of a hit follows from its unitig id (ids are dealt out in ranges)
    const Bucket* peer_tbl[32];
    uint32_t id_bounds[33];
    long long ustride, bstride;
    int owner_bits, n_shards;
    // PEER, optional: a replicated MISS FILTER over the hashes of all owners (one 64-bit word per hash: word = its top
    // filt_bits bits, four bit positions from its low 24 bits; sgc_filter_insert).  A k-mer that has to be probed on
    // its own -- in practice one that overlaps a sequencing error and is in nobody's table -- and whose owner is
    // another GPU is first looked up here, in local HBM: a missing bit proves the miss and saves the NVLink load.
    const unsigned long long* filt;
    int filt_shift;   // 64 - filt_bits
    int my_rank;
    // QUERY = true (sgc_query_count_batch): instead of per-read labels the kernel writes, per settled run, the
    // INTERVAL of store positions its k-mers lie at -- iv_key = query << iv_pbits | first position, iv_val = length << 32 |
    // unitig id; scal->n_export counts them -- and every k-mer that is not in the table as (hash, sort key) --
    // scal->n_live counts those.  Two store positions never hold the same hash under a clear break bit and the
    // position a probe returns is the one the table keeps for that hash, so the number of DISTINCT matching hashes
    // of a query on a unitig is the size of the union of its intervals there (search/query_by_sequence.py:170-189).
    const int32_t* seq_query;
    unsigned long long* iv_key;
    unsigned long long* iv_val;
    unsigned long long iv_cap;
    unsigned long long* miss_hash;
    uint32_t* miss_q;          // the miss's sort key: query << miss_hbits | top miss_hbits bits of the hash
    unsigned long long miss_cap;
    int miss_hbits;
    int iv_pbits;              // iv_key = query << iv_pbits | first position
};
constexpr unsigned QCHUNK = 1024;     // records a warp reserves at a time (a trip writes at most 96)
constexpr int ERR_QUERY_CAP = 4;      // Scalars::err bit: iv_cap / miss_cap too small
constexpr int ERR_QUERY_NOPOS = 8;    // Scalars::err bit: a hit without a store position (table filled hash by hash)

template <int K, bool PACKED>
struct alignas(16) WarpTileS {
    using G = GeoS<K, PACKED>;
    uint32_t F[G::SPW * G::WPS];
    uint32_t R[G::SPW * G::WPS];
    uint32_t RAW[G::SPW * G::RAWW];
    uint32_t P[G::SPW * G::PWS];       // the round's segments, 2 bits per base
    SegDesc D[3][G::SPW];
    unsigned long long pairs[PAIRS_S];
    unsigned short qr[G::QCAP / 2];    // queued runs of this round: k-mer | segment << 7 | followers << 9 | back << 14 | flagged << 15
    unsigned short qs[G::QCAP];        // queued ranges of k-mers to probe one by one: first k-mer | segment << 7 | (count - 1) << 9
    int gpre[G::SPW + 1];
    unsigned segbad;                   // bit sg: segment sg holds a byte that is not A/C/G/T (ASCII input)
    unsigned pad_[2];
};

// the walk past a full home bucket, out of line
__device__ __noinline__ uint32_t lookup_slow_call(const Bucket* tbl, uint32_t nlines, uint32_t hb, uint64_t h, uint32_t* aux) {
    uint32_t ax = AUX_NONE;
    const uint32_t v = lookup_slow(tbl, nlines, hb, h, ax);
    *aux = ax;
    return v;
}

// the four bits of a hash in its word of the miss filter
__host__ __device__ __forceinline__ unsigned long long filter_bits(uint64_t h) {
    return (1ULL << (h & 63)) | (1ULL << ((h >> 6) & 63)) | (1ULL << ((h >> 12) & 63)) | (1ULL << ((h >> 18) & 63));
}
__global__ void filter_insert_kernel(const uint64_t* __restrict__ hash, int64_t n, unsigned long long* words, int shift) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = hash[i];
        atomicOr(words + (h >> shift), filter_bits(h));
    }
}

// 16 packed bases (one 32-bit word, 2 bits each) -> four words of ASCII
__device__ __forceinline__ uint4 expand_packed16(uint32_t x) {
    auto spread = [](uint32_t h) {  // 16 bits -> 8 nibbles holding one 2-bit code each
        uint32_t t = (h | (h << 8)) & 0x00FF00FFu;
        t = (t | (t << 4)) & 0x0F0F0F0Fu;
        return (t | (t << 2)) & 0x33333333u;
    };
    const uint32_t lo = spread(x & 0xFFFFu), hi = spread(x >> 16);
    const uint32_t acgt = 0x54474341u;  // "ACGT"
    return make_uint4(__byte_perm(acgt, 0u, lo), __byte_perm(acgt, 0u, lo >> 16), __byte_perm(acgt, 0u, hi),
                      __byte_perm(acgt, 0u, hi >> 16));
}

// four ASCII bases -> 8 bits (A=0 C=1 G=2 T=3; any other byte gives some code: the caller knows it is invalid)
__host__ __device__ __forceinline__ uint32_t pack_ascii4(uint32_t u) {
    const uint32_t c = ((u >> 1) & 0x03030303u) ^ ((u >> 2) & 0x01010101u);
    return (c | (c >> 6) | (c >> 12) | (c >> 18)) & 0xFFu;
}

// ----------------------------------------------------------------------------------
// build side of the sequence store
// ----------------------------------------------------------------------------------
// word w of the batch = bases [16w, 16w+16) of its ASCII buffer (bytes past the end read as 'A')
__global__ void useq_pack_kernel(const uint8_t* __restrict__ seq, int64_t seq_bytes, uint32_t* __restrict__ out, int64_t nwords) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t u = 0x41414141u;
            const int64_t b = 16 * w + 4 * j;
            if (b + 4 <= seq_bytes && ((uintptr_t)(seq + b) & 3) == 0) u = __ldg(reinterpret_cast<const uint32_t*>(seq + b));
            else {
                u = 0;
                for (int t = 0; t < 4; t++) u |= (uint32_t)(b + t < seq_bytes ? seq[b + t] : (uint8_t)'A') << (8 * t);
            }
            x |= pack_ascii4(u) << (8 * j);
        }
        out[w] = x;
    }
}

// break bits of one batch: cleared exactly where position p and p-1 are k-mers of the same sequence (the build
// kernel sets more of them afterwards: table_insert).  word0 = first break word of the batch (its bit 0 = byte 0
// of the batch's sequence buffer), nwords = words covering the buffer.
__global__ void useq_brk_kernel(const int64_t* __restrict__ off, int64_t nseq, int k, uint32_t* __restrict__ brk_words, int64_t nwords) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o0 = 32 * w;
        int64_t lo = -1, hi = nseq;  // largest s with off[s] <= o0 (-1 if none)
        while (hi - lo > 1) {
            const int64_t mid = lo + (hi - lo) / 2;
            if (off[mid] <= o0) lo = mid; else hi = mid;
        }
        int64_t s = lo;
        uint32_t bits = 0xFFFFFFFFu;
        for (int b = 0; b < 32; b++) {
            const int64_t o = o0 + b;
            while (s + 1 < nseq && off[s + 1] <= o) s++;
            if (s >= 0) {
                const int64_t r = o - off[s], nk = off[s + 1] - off[s] - k + 1;
                if (r >= 1 && r < nk) bits &= ~(1u << b);
            }
        }
        brk_words[w] = bits;
    }
}

// ----------------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------------
template <int K, bool PACKED, bool PEER, bool QUERY = false>
__global__ void __launch_bounds__(NTS, SGC_LS_MIN_CTAS) labelseq_kernel(const LabelSeqArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using G = GeoS<K, PACKED>;
    constexpr int SPW = G::SPW, LPS = G::LPS, WPS = G::WPS, PWS = G::PWS, RAWW = G::RAWW;
    const int lane = threadIdx.x & 31;
    using Tile = WarpTileS<K, PACKED>;
    Tile& wt = reinterpret_cast<Tile*>(smem_raw)[threadIdx.x >> 5];
    const int k = K ? K : a.k;
    const int64_t nsegs = *a.nsegs_p;
    const int64_t nrounds = (nsegs + SPW - 1) / SPW;
    unsigned my_hits = 0;                // per lane: a launch is far below 2^32 k-mers per lane
    unsigned pair_fill = 0;              // warp-uniform: keys waiting in wt.pairs
    // QUERY: the warp's current chunks of the interval / miss arrays (warp-uniform) and, per lane, the last record
    // it wrote (what the unused rest of the last chunks is filled with)
    unsigned long long iv_pos = 0, iv_end = 0, ms_pos = 0, ms_end = 0;
    unsigned long long last_k = 0, last_v = 0, last_h = 0;
    uint32_t last_q = 0;
    unsigned have_last = 0;

    auto flush_pairs = [&]() {  // cold: out of line
        if (a.cnt) flush_pairs_count_out(a.scal, a.pairs, a.pairs_cap, wt.pairs, pair_fill, lane, a.cnt);
        else flush_pairs_out(a.scal, a.pairs, a.pairs_cap, wt.pairs, pair_fill, lane);
        pair_fill = 0;
    };
    // one candidate (sequence, id) key per lane at most; the order of the keys is irrelevant (they are
    // grouped per sequence by the label tail)
    auto push_key = [&](bool has, unsigned long long key) {
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (m) {
            const unsigned total = __popc(m);
            if (pair_fill + total > PAIRS_S) flush_pairs();
            if (has) wt.pairs[pair_fill + __popc(m & ((1u << lane) - 1u))] = key;
            pair_fill += total;
        }
    };

    unsigned round_spare = 0xFFFFFFFFu;
    auto claim_round = [&]() -> unsigned {  // rounds are claimed two at a time
        if (round_spare != 0xFFFFFFFFu) {
            const unsigned r = round_spare;
            round_spare = 0xFFFFFFFFu;
            return r;
        }
        unsigned r = 0;
        if (lane == 0) r = atomicAdd(&a.scal->round_ctr, 2u);
        r = __shfl_sync(0xffffffffu, r, 0);
        round_spare = r + 1u;
        return r;
    };
    auto prefetch_desc = [&](int slot, unsigned round) {
        if (lane < 2 * SPW) {
            const int sgi = lane >> 1, half = lane & 1;
            const int64_t g = (int64_t)round * SPW + sgi;
            unsigned char* dst = reinterpret_cast<unsigned char*>(&wt.D[slot][sgi]) + 16 * half;
            if ((int64_t)round < nrounds && g < nsegs) {
                cp_async16(dst, reinterpret_cast<const unsigned char*>(a.segdesc + g) + 16 * half);
            } else {
                *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);  // nk = 0: nothing to do
            }
        }
    };
    auto prefetch_raw = [&](int slot) {
        const int sgi = lane / LPS, sub = lane % LPS;
        const SegDesc& d = wt.D[slot][sgi];
        if (d.nk > 0) {
            const int Ls = d.nk + k - 1;
            if constexpr (PACKED) {  // records are word aligned and a segment starts 32 bytes after the last:
                // copy the 16-byte chunks that hold its bytes (each holds valid bytes of the buffer, so it is mapped)
                const uintptr_t g0 = (uintptr_t)a.seq + (uintptr_t)d.start;
                const uintptr_t a0 = g0 & ~(uintptr_t)15;
                const int nchunks = (int)(((g0 + (uintptr_t)((Ls + 3) >> 2) + 15) & ~(uintptr_t)15) - a0) >> 4;
                unsigned char* dst = reinterpret_cast<unsigned char*>(wt.RAW + sgi * RAWW);
                for (int c = sub; c < nchunks; c += LPS) cp_async16(dst + 16 * c, reinterpret_cast<const void*>(a0 + 16 * (uintptr_t)c));
            } else {
                const uintptr_t g0 = (uintptr_t)a.seq + (uintptr_t)d.start;
                const uintptr_t a0 = g0 & ~(uintptr_t)15;
                const int nchunks = (int)(((g0 + (uintptr_t)Ls + 15) & ~(uintptr_t)15) - a0) >> 4;
                unsigned char* dst = reinterpret_cast<unsigned char*>(wt.RAW + sgi * RAWW) + 16;
                for (int c = sub; c < nchunks; c += LPS)
                    cp_async16(dst + 16 * c, reinterpret_cast<const void*>(a0 + 16 * (uintptr_t)c));
            }
        }
    };

    unsigned rq0 = claim_round(), rq1 = claim_round(), rq2 = claim_round();
    prefetch_desc(0, rq0);
    prefetch_desc(1, rq1);
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();
    prefetch_raw(0);
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();

    for (unsigned it = 0;; it++) {
        const unsigned round = rq0;
        if ((int64_t)round >= nrounds) break;
        const int cur = it % 3;
        const int nxt = (it + 1) % 3, nx2 = (it + 2) % 3;
        prefetch_desc(nx2, rq2);  // descriptors of the round after next
        cp_async_commit();
        rq0 = rq1;
        rq1 = rq2;
        rq2 = claim_round();

        // ---- group prefix of this round's segments ----
        {
            const int nk = lane < SPW ? wt.D[cur][lane].nk : 0;
            int incl = (nk + GK - 1) / GK;
#pragma unroll
            for (int d = 1; d < SPW; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane < SPW) wt.gpre[lane + 1] = incl;
            if (lane == 0) {
                wt.gpre[0] = 0;
                wt.segbad = 0;
            }
        }
        __syncwarp();

        // ---- stage forward + reverse-complement ASCII words and the 2-bit words (LPS lanes per segment) ----
        {
            const int sg = lane / LPS, sub = lane % LPS;
            const SegDesc& d = wt.D[cur][sg];
            const int nk = d.nk;
            const int Ls = nk + k - 1;
            uint32_t* Fs = wt.F + sg * WPS + FOFF;
            uint32_t* Ps = wt.P + sg * PWS;
            const int npw = (Ls + 15) >> 4;
            if (nk) {
                if constexpr (PACKED) {
                    const uint32_t* pk = wt.RAW + sg * RAWW + ((((uintptr_t)a.seq + (uintptr_t)d.start) & 15) >> 2);
                    for (int j = sub; j < npw; j += LPS) {
                        const uint32_t x = pk[j];
                        Ps[j] = x;
                        *reinterpret_cast<uint4*>(Fs + 4 * j) = expand_packed16(x);
                    }
                } else {
                    const int r0 = (int)(((uintptr_t)a.seq + (uintptr_t)d.start) & 15);
                    const int s = r0 & 3;
                    const uint32_t* rw0 = wt.RAW + sg * RAWW + 4 + (r0 >> 2);  // raw copies start 16 bytes into their slot
                    uint32_t bad = 0;
                    for (int w = sub; 4 * w < Ls; w += LPS) {
                        const uint32_t u = sgc::funnel_bytes(rw0[w], rw0[w + 1], s);
                        Fs[w] = u;
                        // every byte must be one of A C G T: "ACTG"[bits 1..2 of the byte] == the byte
                        const uint32_t code = (u >> 1) & 0x03030303u;
                        const uint32_t sel = __byte_perm(code | (code >> 4), 0u, 0x4420u);
                        const int nval = Ls - 4 * w;
                        const uint32_t mask = nval >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nval)) - 1u);
                        bad |= (__byte_perm(0x47544341u, 0u, sel) ^ u) & mask;
                    }
                    if (bad) {
                        if (a.status) atomicOr(&a.status[d.seq], 1);
                        atomicOr(&wt.segbad, 1u << sg);  // its followers are hashed like everybody else
                    }
                }
            }
            __syncwarp();
            if (nk) {
                // R word m holds R[4m-4-p .. +4) = reverse complement of F[Ls-4m+p .. +4)
                const int p = sgc::rc_pad(nk);
                const int nrw = (Ls + 4 + p + 3) >> 2;
                uint32_t* Rs = wt.R + sg * WPS;
                const uint32_t* Fw = wt.F + sg * WPS;  // F byte f lives at byte 16 + f of this slot (f >= -7 here)
                for (int m = sub; m < nrw; m += LPS) {
                    const int b = 16 + Ls - 4 * m + p;
                    const uint32_t u = sgc::funnel_bytes(Fw[b >> 2], Fw[(b >> 2) + 1], b & 3);
                    // complement and reverse in one PRMT: nibbles of the selector = the 2-bit codes, last base first
                    const uint32_t code = (u >> 1) & 0x03030303u;
                    const uint32_t sel = __byte_perm(code | (code << 12), 0u, 0x4413u);
                    Rs[m] = __byte_perm(0x43414754u, 0u, sel);  // "TGAC"[code]
                }
                if constexpr (!PACKED) {
                    for (int j = sub; j < npw; j += LPS)
                        Ps[j] = pack_ascii4(Fs[4 * j]) | (pack_ascii4(Fs[4 * j + 1]) << 8) | (pack_ascii4(Fs[4 * j + 2]) << 16) |
                                (pack_ascii4(Fs[4 * j + 3]) << 24);
                }
            }
        }
        __syncwarp();
        prefetch_raw(nxt);  // lands while this round is worked on
        cp_async_commit();

        // ---- the round's work items: the anchors of its groups, then what they queue ----
        // Two queues (their cursors are warp-uniform registers): RUNS (an anchor k-mer, hashed and probed, and the
        // followers it may settle through the store) and RANGES of k-mers that each have to be probed on their own.
        // A trip gives its first lanes to runs -- anchors first; queued runs only when a few have gathered or nothing
        // else is left, so that the compare code below runs in few trips -- and fills the others from the ranges.
        const volatile int* gpre = wt.gpre;
        const int ng = gpre[SPW];
        int a_done = 0, qr_fill = 0, qr_done = 0, qs_fill = 0, qs_done = 0;
        for (;;) {
            const int na = ng - a_done, nr = qr_fill - qr_done, nsi = qs_fill - qs_done;
            if (na + nr + nsi == 0) break;
            const int n_a = na < 32 ? na : 32;
            int n_r = 0;
            if (n_a < 32 && nr > 0 && (na > 0 || nr >= RUN_MIN || nsi == 0)) n_r = nr < 32 - n_a ? nr : 32 - n_a;
            const int n_run = n_a + n_r;        // lanes [0, n_run) hold runs
            // an item is a RUN: its anchor k-mer qa (hashed and probed here) and n-1 followers, the k-mers after
            // it (back = false) or before it (back = true); flagged = a miss may not turn the run around again
            int sg = 0, qa = 0, n = 1;
            bool back = false, flagged = false, act = lane < n_run;
            const bool is_anchor = lane < n_a;
            if (is_anchor) {
                const int item = a_done + lane;
#pragma unroll
                for (int s = 1; s < SPW; s++) sg += item >= gpre[s];
                qa = (item - gpre[sg]) * GK;
                n = GK;
            } else if (act) {
                const unsigned u = wt.qr[qr_done + lane - n_a];
                qa = (int)(u & 127u);
                sg = (int)((u >> 7) & 3u);
                n = (int)((u >> 9) & 31u) + 1;
                back = (u >> 14) & 1u;
                flagged = (u >> 15) & 1u;
            }
            a_done += n_a;
            qr_done += n_r;
            if (n_run < 32 && nsi > 0) {  // warp-uniform: lanes [n_run, 32) take single k-mers from the next ranges
                const bool has = lane < nsi;
                const unsigned u = has ? wt.qs[qs_done + lane] : 0u;
                const int cnt = has ? (int)((u >> 9) & 31u) + 1 : 0;
                int incl = cnt;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const int x = __shfl_up_sync(0xffffffffu, incl, dd);
                    if (lane >= dd) incl += x;
                }
                const int first = n_run + incl - cnt;   // the lane slot this range starts at
                // bit j = a range starts at slot j: slot j belongs to the popc(heads up to j)-th range
                const unsigned heads = __reduce_or_sync(0xffffffffu, has && first < 32 ? 1u << first : 0u);
                const int it = __popc(heads & (0xFFFFFFFFu >> (31 - lane))) - 1;
                const unsigned ui = __shfl_sync(0xffffffffu, u, it < 0 ? 0 : it);
                const int fi = __shfl_sync(0xffffffffu, first, it < 0 ? 0 : it);
                const int ci = __shfl_sync(0xffffffffu, cnt, it < 0 ? 0 : it);
                if (lane >= n_run && it >= 0 && lane - fi < ci) {
                    act = true;
                    qa = (int)(ui & 127u) + (lane - fi);
                    sg = (int)((ui >> 7) & 3u);
                }
                // ranges that fit are done; the one cut off by the end of the trip keeps what is left of it
                const bool whole = has && first + cnt <= 32;
                if (has && first < 32 && !whole) {
                    const int taken = 32 - first;
                    wt.qs[qs_done + lane] = (unsigned short)((u & 0x180u) | ((u & 127u) + (unsigned)taken) | ((unsigned)(cnt - taken - 1) << 9));
                }
                qs_done += __popc(__ballot_sync(0xffffffffu, whole));
            }
            const SegDesc& d = wt.D[cur][sg];
            const int nk = d.nk;
            if (is_anchor && nk - qa < GK) n = nk - qa;
            const uint32_t* Fs = wt.F + sg * WPS + FOFF;
            const uint32_t* Rs = wt.R + sg * WPS;

            // hash of k-mer qa of segment sg (idle lanes hash k-mer 0 of segment 0: staged or stale words, unused)
            uint64_t hf, hr;
            const int rb = sgc::rc_byte_of(nk, nk - 1 - qa);
            if constexpr (K != 0) {
                hf = sgc::murmur3_h1_window<K ? K : 1>(Fs + (qa >> 2), qa & 3);
                hr = sgc::murmur3_h1_window<K ? K : 1>(Rs + (rb >> 2), rb & 3);
            } else {
                hf = act ? sgc::murmur3_h1_stream(Fs, qa, k) : 0ULL;
                hr = act ? sgc::murmur3_h1_stream(Rs, rb, k) : 0ULL;
            }
            const uint64_t h = hf ^ hr;
            const uint32_t hb = act ? home_bucket(h, a.nbuckets, a.home_shift) : 0u;
            const Bucket* tb = a.tbl;
            uint32_t hbp = hb;
            bool known_miss = false;
            if constexpr (PEER) {
                const unsigned owner = (unsigned)(h >> (64 - a.owner_bits));
                if (act) tb = a.peer_tbl[owner];
                if (a.filt != nullptr && n_run < 32) {  // warp-uniform: this trip holds k-mers probed on their own
                    if (act && lane >= n_run && owner != (unsigned)a.my_rank) {
                        const unsigned long long w = __ldg(a.filt + (h >> a.filt_shift));
                        known_miss = (w & filter_bits(h)) != filter_bits(h);
                        if (known_miss) {  // nothing to fetch: like an idle lane
                            tb = a.tbl;
                            hbp = 0u;
                        }
                    }
                }
            }
            Entry e0, e1;
            load_bucket_nc(tb + hbp, e0, e1);  // idle lanes probe bucket 0 of the local table (an L2 hit)
            uint32_t v = SGC_MISS, aux = AUX_NONE;
            if (act && !known_miss) {
                if (!bucket_match(e0, e1, h, v, aux)) {
                    v = SGC_MISS;
                    aux = AUX_NONE;
                    if (bucket_overflowed(e0)) v = lookup_slow_call(tb, a.nlines, hb, h, &aux);
                }
            }
            unsigned nhit = v != SGC_MISS;
            // QUERY: the settled k-mers of this item, as bits of the window lo .. lo+n-1 whose anchor (bit ao) lies at
            // store position aux & AUX_MASK; k-mer lo+i lies at that position + (i - ao) (along) or - (i - ao)
            unsigned q_settled = v != SGC_MISS ? 1u : 0u;
            int q_ao = 0;
            bool q_along = true;

            const bool is_run = act && n > 1;
            if (n_run > 0) {  // warp-uniform: this trip holds runs
                unsigned singles = 0;          // followers to probe one by one, bit i = k-mer lo + i (always a contiguous range)
                unsigned run_item = 0u;        // a follow-up run (encoded, bit 15 set) or 0
                int lo = qa;
                if (is_run) {
                    lo = back ? qa - (n - 1) : qa;
                    const int ao = back ? n - 1 : 0;            // the anchor's place in the window of k-mers lo .. lo+n-1
                    const unsigned valid = n >= 32 ? 0xFFFFFFFFu : (1u << n) - 1u;
                    const uint32_t gp = aux & AUX_MASK;
                    const bool has_g = v != SGC_MISS && gp != AUX_NONE && !((wt.segbad >> sg) & 1u);
                    if (has_g) {
                        const bool along = ((aux & ORI_BIT) != 0) == (hf < hr);
                        const uint32_t* useq = a.useq;
                        const uint32_t* brkp = a.brk;
                        if constexpr (PEER) {  // the shard of the store this unitig lies in
                            int shard = 0;
                            for (int s = 1; s < a.n_shards; s++) shard += v >= a.id_bounds[s];
                            useq += shard * a.ustride;
                            brkp += shard * a.bstride;
                        }
                        // break bits of positions gp-31 .. gp (low word, bit 31 = gp) and gp+1 .. gp+32 (high word)
                        const uint32_t bi = gp + (uint32_t)(BRK_PAD - 31);
                        const uint32_t w0 = __ldg(brkp + (bi >> 5)), w1 = __ldg(brkp + (bi >> 5) + 1);
                        const uint32_t w2 = __ldg(brkp + (bi >> 5) + 2);
                        unsigned clean;  // bit i: the k bases of k-mer lo+i equal the store's
                        if constexpr (K != 0) {
                            constexpr int WL = GK - 1 + (K ? K : 1);   // bases of a full window
                            constexpr int NWD = (WL + 15) / 16;
                            static_assert(NWD <= 4 || K == 0, "window words");
                            // read k-mer lo+i lies at store position gp + (i - ao) (along) or gp - (i - ao): the window's
                            // base j is store base wsi + j, resp. the complement of store base wsi + WL - 1 - j
                            const int wsi = along ? (int)gp - ao : (int)gp + ao - (GK - 1);
                            const uint32_t* up = useq + (wsi >> 4);
                            const int sh = 2 * (wsi & 15);
                            const uint32_t* Pw = wt.P + sg * PWS + (lo >> 4);
                            const int rsh = 2 * (lo & 15);
                            uint32_t u[NWD + 1], pw[NWD + 1], un[NWD], rd[NWD];
#pragma unroll
                            for (int j = 0; j <= NWD; j++) u[j] = __ldg(up + j);
#pragma unroll
                            for (int j = 0; j <= NWD; j++) pw[j] = Pw[j];
#pragma unroll
                            for (int j = 0; j < NWD; j++) {
                                rd[j] = __funnelshift_r(pw[j], pw[j + 1], rsh);
                                un[j] = __funnelshift_r(u[j], u[j + 1], sh);
                            }
                            if (!along) {  // reverse complement of the WL-base window
                                uint32_t r[NWD + 1];
#pragma unroll
                                for (int j = 0; j < NWD; j++) r[j] = __brev(un[NWD - 1 - j]);
                                r[NWD] = 0u;
                                constexpr int RS = 32 * NWD - 2 * WL;
#pragma unroll
                                for (int j = 0; j < NWD; j++) {
                                    const uint32_t x = __funnelshift_r(r[j], r[j + 1], RS);
                                    un[j] = ~(((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1));
                                }
                            }
                            // mismatching bases among the n-1+K the run really has: first (f) and last (l)
                            const int wn = n - 1 + K;
                            int f = WL + GK, l = -1;
#pragma unroll
                            for (int j = NWD - 1; j >= 0; j--) {
                                uint32_t x = rd[j] ^ un[j];
                                x = (x | (x >> 1)) & 0x55555555u;
                                const int nb = wn - 16 * j;
                                if (nb < 16) x &= nb > 0 ? (1u << (2 * nb)) - 1u : 0u;
                                if (x) {
                                    f = 16 * j + ((__ffs((int)x) - 1) >> 1);
                                    if (l < 0) l = 16 * j + ((31 - __clz((int)x)) >> 1);
                                }
                            }
                            // k-mer i is clean when its k bases avoid every mismatch: i + K - 1 < f, or i > l
                            const unsigned clean_a = f >= K ? (f - K + 1 >= GK ? GKMASK : (1u << (f - K + 1)) - 1u) : 0u;
                            const unsigned clean_b = l < 0 ? GKMASK : (l >= GK - 1 ? 0u : (GKMASK << (l + 1)) & GKMASK);
                            clean = clean_a | clean_b;
                        } else {
                            // any k: walk the bases away from the anchor until the first mismatch
                            const uint32_t* Pw = wt.P + sg * PWS;
                            const int wn = n - 1 + k;
                            int f = wn;
                            for (int j = 0; j < wn; j++) {
                                const int t = back ? k - 1 - j : j;        // base offset from the anchor's first base
                                const int rbp = qa + t;
                                const uint32_t rbase = (Pw[rbp >> 4] >> (2 * (rbp & 15))) & 3u;
                                const int pos = along ? (int)gp + t : (int)gp + k - 1 - t;
                                uint32_t ub = (__ldg(useq + (pos >> 4)) >> (2 * (pos & 15))) & 3u;
                                if (!along) ub ^= 3u;
                                if (rbase != ub) {
                                    f = j;
                                    break;
                                }
                            }
                            // the anchor and its first c-1 followers are clean; as bits of the window lo .. lo+n-1
                            const int c = f >= k ? (f - k + 1 >= n ? n : f - k + 1) : 0;
                            const unsigned pre = c >= 32 ? 0xFFFFFFFFu : (1u << c) - 1u;
                            clean = back ? pre << (n - c) : pre;
                        }
                        const uint32_t below = __funnelshift_r(w0, w1, bi & 31u);   // bit b = position gp-31+b
                        const uint32_t above = __funnelshift_r(w1, w2, bi & 31u);   // bit b = position gp+1+b
                        // the lim followers next to the anchor may inherit its id: no break in (gp, gp+d] resp. (gp-d, gp]
                        const unsigned lim = along != back ? (unsigned)__ffs((int)(above | 0x80000000u)) - 1u
                                                           : (unsigned)__clz((int)(below | 1u));
                        const unsigned inh = back ? ((int)lim >= n - 1 ? valid : valid & ~((1u << (n - 1 - (int)lim)) - 1u))
                                                  : ((2u << lim) - 1u) & valid;
                        const unsigned settled = ((clean & inh) | (1u << ao)) & valid;
                        nhit = (unsigned)__popc(settled);
                        q_settled = settled;
                        q_ao = ao;
                        q_along = along;
                        singles = inh & ~settled;
#if SGC_LS_REANCHOR
                        if (valid & ~inh) {  // the k-mers past the break: a run of their own, anchored next to the break
                            const int n2 = n - 1 - (int)lim;
                            const int qa2 = back ? qa - ((int)lim + 1) : qa + (int)lim + 1;
                            run_item = 0x10000u | (flagged ? 1u << 15 : 0u) | (back ? 1u << 14 : 0u) | ((unsigned)(n2 - 1) << 9) |
                                       ((unsigned)sg << 7) | (unsigned)qa2;
                        }
#else
                        singles |= valid & ~inh;
#endif
                    } else {
                        // no anchor: the far end of the run may be one (a sequencing error early in the anchor leaves
                        // the last followers intact); a second miss probes what is left one by one
#if SGC_LS_REANCHOR
                        if (!flagged && n > 2 && !((wt.segbad >> sg) & 1u) && (v == SGC_MISS)) {
                            const int qa2 = back ? qa - (n - 1) : qa + (n - 1);
                            run_item = 0x10000u | (1u << 15) | (back ? 0u : 1u << 14) | ((unsigned)(n - 2) << 9) | ((unsigned)sg << 7) |
                                       (unsigned)qa2;
                        } else
#endif
                            singles = valid & ~(1u << ao);
                    }
                }
                // queue what the store did not settle: one range and one run per lane at most
                const unsigned ms = __ballot_sync(0xffffffffu, singles != 0u);
                const unsigned mr = __ballot_sync(0xffffffffu, run_item != 0u);
                const unsigned lower = (1u << lane) - 1u;
                if (singles) {
                    const int first = lo + __ffs((int)singles) - 1;
                    wt.qs[qs_fill + __popc(ms & lower)] =
                        (unsigned short)((unsigned)first | ((unsigned)sg << 7) | ((unsigned)(__popc(singles) - 1) << 9));
                }
                if (run_item) wt.qr[qr_fill + __popc(mr & lower)] = (unsigned short)(run_item & 0xFFFFu);
                qs_fill += __popc(ms);
                qr_fill += __popc(mr);
            }
            my_hits += nhit;
            if constexpr (QUERY) {
                const uint32_t qid = act ? (uint32_t)__ldg(a.seq_query + d.seq) : 0u;
                const uint32_t mkey = (qid << a.miss_hbits) | (uint32_t)(h >> (64 - a.miss_hbits));
                // Records go into chunks of QCHUNK slots the warp reserves from one global counter (one atomic per
                // chunk, not per trip); what is left of a chunk is filled with copies of a valid record -- a repeated
                // interval does not change a union, a repeated (hash, query) is a duplicate like any other.
                // k-mers the table does not hold: (hash, query)
                const bool miss = act && v == SGC_MISS;
                const unsigned mm = __ballot_sync(0xffffffffu, miss);
                if (mm) {
                    const unsigned total = (unsigned)__popc(mm);
                    if (ms_pos + total > ms_end) {
                        const int src = __ffs((int)mm) - 1;
                        const unsigned long long fh = __shfl_sync(0xffffffffu, (unsigned long long)h, src);
                        const uint32_t fq = __shfl_sync(0xffffffffu, mkey, src);
                        for (unsigned long long pp = ms_pos + lane; pp < ms_end; pp += 32)
                            if (pp < a.miss_cap) {
                                a.miss_hash[pp] = fh;
                                a.miss_q[pp] = fq;
                            }
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(&a.scal->n_live, (unsigned long long)QCHUNK);
                        ms_pos = __shfl_sync(0xffffffffu, base, 0);
                        ms_end = ms_pos + QCHUNK;
                    }
                    if (miss) {
                        const unsigned long long pos = ms_pos + (unsigned)__popc(mm & ((1u << lane) - 1u));
                        if (pos < a.miss_cap) {
                            a.miss_hash[pos] = h;
                            a.miss_q[pos] = mkey;
                        } else {
                            atomicOr(&a.scal->err, ERR_QUERY_CAP);
                        }
                        last_h = h;
                        last_q = mkey;
                        have_last |= 1u;
                    }
                    ms_pos += total;
                }
                // settled k-mers: one interval of store positions per run of set bits
                const uint32_t gp = aux & AUX_MASK;
                if (q_settled && gp == AUX_NONE) {
                    atomicOr(&a.scal->err, ERR_QUERY_NOPOS);
                    q_settled = 0u;
                }
                const int cnt = __popc(q_settled & ~(q_settled << 1));  // runs of set bits
                int incl = cnt;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const int x = __shfl_up_sync(0xffffffffu, incl, dd);
                    if (lane >= dd) incl += x;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total) {
                    auto record = [&](unsigned m, unsigned long long& rk, unsigned long long& rv) -> unsigned {
                        const int i0 = __ffs((int)m) - 1;
                        const unsigned rest = ~(m >> i0);
                        const int len = rest ? __ffs((int)rest) - 1 : 32 - i0;   // set bits from i0 upwards
                        const int i1 = i0 + len - 1;
                        const int first = q_along ? (int)gp + (i0 - q_ao) : (int)gp - (i1 - q_ao);
                        rk = ((unsigned long long)qid << a.iv_pbits) | (unsigned)first;
                        rv = ((unsigned long long)len << 32) | v;
                        return len >= 32 ? 0u : m & ~(((1u << len) - 1u) << i0);
                    };
                    if (iv_pos + (unsigned)total > iv_end) {
                        unsigned long long fk = 0, fv = 0;
                        if (cnt) record(q_settled, fk, fv);
                        const int src = __ffs((int)__ballot_sync(0xffffffffu, cnt > 0)) - 1;
                        fk = __shfl_sync(0xffffffffu, fk, src);
                        fv = __shfl_sync(0xffffffffu, fv, src);
                        for (unsigned long long pp = iv_pos + lane; pp < iv_end; pp += 32)
                            if (pp < a.iv_cap) {
                                a.iv_key[pp] = fk;
                                a.iv_val[pp] = fv;
                            }
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(&a.scal->n_export, (unsigned long long)QCHUNK);
                        iv_pos = __shfl_sync(0xffffffffu, base, 0);
                        iv_end = iv_pos + QCHUNK;
                    }
                    unsigned long long pos = iv_pos + (unsigned)(incl - cnt);
                    unsigned m = q_settled;
                    while (m) {
                        unsigned long long rk, rv;
                        m = record(m, rk, rv);
                        if (pos < a.iv_cap) {
                            a.iv_key[pos] = rk;
                            a.iv_val[pos] = rv;
                        } else {
                            atomicOr(&a.scal->err, ERR_QUERY_CAP);
                        }
                        last_k = rk;
                        last_v = rv;
                        have_last |= 2u;
                        pos++;
                    }
                    iv_pos += (unsigned)total;
                }
            } else {
                // a hit is a candidate label unless the previous item gives the same (sequence, id)
                const unsigned long long key = ((unsigned long long)d.seq << 32) | v;
                const unsigned long long prevk = __shfl_up_sync(0xffffffffu, key, 1);
                push_key(v != SGC_MISS && !(lane > 0 && prevk == key), key);
            }
            __syncwarp();
        }

        cp_async_wait_all();  // next round's bytes, next-next descriptors
        __syncwarp();
    }
    cp_async_wait_all();

    if (pair_fill) flush_pairs();
    if constexpr (QUERY) {
        if (iv_pos < iv_end) {
            const int src = __ffs((int)__ballot_sync(0xffffffffu, (have_last & 2u) != 0)) - 1;
            const unsigned long long fk = __shfl_sync(0xffffffffu, last_k, src), fv = __shfl_sync(0xffffffffu, last_v, src);
            for (unsigned long long pp = iv_pos + lane; pp < iv_end; pp += 32)
                if (pp < a.iv_cap) {
                    a.iv_key[pp] = fk;
                    a.iv_val[pp] = fv;
                }
        }
        if (ms_pos < ms_end) {
            const int src = __ffs((int)__ballot_sync(0xffffffffu, (have_last & 1u) != 0)) - 1;
            const unsigned long long fh = __shfl_sync(0xffffffffu, last_h, src);
            const uint32_t fq = __shfl_sync(0xffffffffu, last_q, src);
            for (unsigned long long pp = ms_pos + lane; pp < ms_end; pp += 32)
                if (pp < a.miss_cap) {
                    a.miss_hash[pp] = fh;
                    a.miss_q[pp] = fq;
                }
        }
    }
    unsigned long long hits = my_hits;  // the host derives the misses from the plan's k-mer total
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, dd);
    if (lane == 0 && hits) atomicAdd(&a.scal->n_hits, hits);
}

}  // namespace dev
}  // namespace sgc
