// sgc_tile.cuh -- segment staging and the warp-autonomous tile kernel (hash / insert / lookup / label / partition / gather / minhash modes).
// Part of libsgc_b200.so (device code; included by sgc_b200.cu only).
#pragma once
#include "sgc_kmer.cuh"
#include "sgc_table.cuh"

namespace sgc {
namespace dev {

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

// bytes F[f .. f+4) of a segment of Ls bases that sits at byte offset r0 of a raw shared-memory copy;
// bytes outside [0, Ls) read as 0
__device__ __forceinline__ uint32_t raw_window4(const uint32_t* raw, int r0, int f, int Ls, uint32_t& mask) {
    const int lo_clip = f < 0 ? -f : 0;
    const int hi_clip = f + 4 > Ls ? f + 4 - Ls : 0;
    if (lo_clip >= 4 || hi_clip >= 4) { mask = 0; return 0; }
    mask = (0xFFFFFFFFu << (8 * lo_clip)) & (0xFFFFFFFFu >> (8 * hi_clip));
    const int b = r0 + f + 16;  // raw copies start 16 bytes into their slot, so b - (b & 3) >= 0 for f >= -3
    const int s = b & 3, w = b >> 2;
    return sgc::funnel_bytes(raw[w], raw[w + 1], s) & mask;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Wait until a flag written by another GPU (peer store over NVLink) reaches `want`; bounded, so that a
// protocol error ends in an error bit (8) instead of a hung GPU.
__device__ __forceinline__ void spin_until_equal(const unsigned int* flag, unsigned int want, int* err) {
    const long long t0 = clock64();
    while (*reinterpret_cast<const volatile unsigned int*>(flag) != want) {
        if (clock64() - t0 > (1LL << 33)) {  // ~4 s at 2 GHz
            atomicOr(err, 8);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();  // acquire: what was stored before the flag is visible now
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
// the tile kernel.  Every WARP runs its own rounds: fetch SPW segments (asynchronous copies, two
// rounds deep), stage their forward and reverse-complement bytes in its slice of shared memory
// (8 lanes per segment), then each lane
// hashes groups of 4 consecutive k-mers and feeds them to the mode's sink.  No block barriers:
// warps drift apart, so one warp's exposed latencies (round fetch, staging loads, table probes)
// are covered by the other warps of the SM.
// ----------------------------------------------------------------------------------
// MODE_PARTITION's per-owner bookkeeping; the other modes carry an empty stand-in so that it costs
// them no shared memory (more resident warps for the probing kernels)
struct PartitionState {
    unsigned wcnt[32];              // per-owner counts of one warp iteration
    unsigned wpre[32];              // ... their exclusive prefix (slot of the owner's run in the staging area)
    uint64_t* dst[PAIR_BUF];        // ... destination address of every staged hash
    unsigned long long wbase[32];   // ... and the reserved base positions in the owners' regions
};
struct NoPartitionState {
    unsigned wcnt[1];
    unsigned wpre[1];
    uint64_t* dst[1];
    unsigned long long wbase[1];
};

template <int K, bool PART = false>
struct alignas(16) WarpTile {
    uint32_t F[SPW * Geo<K>::WPS];
    uint32_t R[SPW * Geo<K>::WPS];
    uint32_t RAW[SPW * Geo<K>::WPS];     // 16-byte-aligned raw copies of the next round's segments (cp.async)
    SegDesc D[3][SPW];                   // prefetched descriptors of this round and the next two
    unsigned long long pairs[PAIR_BUF];
    long long start[SPW];
    long long seq[SPW];
    long long kout[SPW];
    int nk[SPW];
    int gpre[SPW + 1];
    int pad[3];
    typename std::conditional<PART, PartitionState, NoPartitionState>::type part;  // last: keeps the copies above aligned
};

// one global reservation for the keys a warp has staged, then a coalesced copy
__device__ __noinline__ void flush_pairs_out(Scalars* scal, unsigned long long* pairs, unsigned long long cap,
                                             const unsigned long long* staged, unsigned n, int lane) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&scal->pair_count, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (unsigned j = lane; j < n; j += 32)
        if (base + j < cap) pairs[base + j] = staged[j];
    __syncwarp();
}

// the same, and every key counted under its sequence (the first step of the label tail, done while the keys are at hand)
__device__ __noinline__ void flush_pairs_count_out(Scalars* scal, unsigned long long* pairs, unsigned long long cap,
                                                   const unsigned long long* staged, unsigned n, int lane,
                                                   unsigned long long* cnt) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&scal->pair_count, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (unsigned j = lane; j < n; j += 32)
        if (base + j < cap) {
            const unsigned long long key = staged[j];
            pairs[base + j] = key;
            atomicAdd(&cnt[key >> 32], 1ULL);
        }
    __syncwarp();
}

template <int K, int MODE>
__global__ void __launch_bounds__(NT, (MODE == MODE_LABEL || MODE == MODE_LOOKUP) ? SGC_PROBE_MIN_CTAS : 1) tile_kernel(const TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WPS = Geo<K>::WPS;
    const int lane = threadIdx.x & 31;
    using Tile = WarpTile<K, MODE == MODE_PARTITION>;
    Tile& wt = reinterpret_cast<Tile*>(smem_raw)[threadIdx.x >> 5];
    const int k = K ? K : a.k;
    const int64_t nsegs = *a.nsegs_p;
    const int64_t nrounds = (nsegs + SPW - 1) / SPW;
    unsigned long long my_hits = 0, my_miss = 0;
    unsigned pair_fill = 0;  // warp-uniform: keys waiting in wt.pairs

    auto flush_pairs = [&]() {  // cold (once per ~40 trips): kept out of line so the hot loops stay small
        flush_pairs_out(a.scal, a.pairs, a.pairs_cap, wt.pairs, pair_fill, lane);
        pair_fill = 0;
    };

    // K4 candidate rule shared by MODE_LABEL and MODE_GATHER: a hit is a candidate (sequence, id) key
    // unless the previous k-mer of the same segment has the same id (known through a warp shuffle;
    // the first k-mer of a lane-0 group is conservatively a candidate).  All 32 lanes must call this.
    auto emit_candidates = [&](const uint32_t (&ids)[4], bool active, int sg, int i, int nv) {
        // ids[] of k-mers that do not exist (t >= nv, idle lanes) or have no id yet are SGC_MISS, so the
        // rule needs no validity tests; slots are handed out with one ballot per k-mer position (the
        // order of the keys inside the staging buffer is irrelevant, they are grouped per sequence later)
        const uint32_t prev_lane = __shfl_up_sync(0xffffffffu, ids[3], 1);
        const uint32_t prev0 = (i > 0 && lane > 0) ? prev_lane : SGC_MISS;
        bool cand[4];
        unsigned m[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            cand[t] = ids[t] != SGC_MISS && ids[t] != (t == 0 ? prev0 : ids[t - 1]);
            m[t] = __ballot_sync(0xffffffffu, cand[t]);
        }
        const unsigned c0 = __popc(m[0]), c1 = __popc(m[1]), c2 = __popc(m[2]), c3 = __popc(m[3]);
        const unsigned total = c0 + c1 + c2 + c3;
        if (total) {
            if (pair_fill + total > PAIR_BUF) flush_pairs();
            const unsigned long long sq = (unsigned long long)wt.seq[sg] << 32;
            const unsigned lt = (1u << lane) - 1u;
            const unsigned base[4] = {pair_fill, pair_fill + c0, pair_fill + c0 + c1, pair_fill + c0 + c1 + c2};
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (cand[t]) wt.pairs[base[t] + __popc(m[t] & lt)] = sq | ids[t];
            pair_fill += total;
        }
        (void)active;
        (void)nv;
    };

    // ------------------------------------------------------------------------------------------
    // Round pipeline.  A warp claims rounds (SPW segments each) from one atomic counter, always three
    // ahead, and keeps two asynchronous copies in flight while it works on round r:
    //   descriptors of round r+2  (cp.async of SPW SegDesc, 32 bytes each)
    //   raw bytes of round r+1    (cp.async of the 16-byte chunks that cover each segment)
    // the first issued at the top of round r, the second as soon as the staging of round r has released
    // the single raw area; both are waited for at the end of the round, so they land while it is hashed.
    // The staging of round r itself (re-alignment + reverse complement) reads shared memory only.
    // ------------------------------------------------------------------------------------------
    // rounds are claimed two at a time (adjacent): half as many atomics to wait for
    unsigned round_spare = 0xFFFFFFFFu;  // warp-uniform: the second round of the last claim, if still unused
    auto claim_round = [&]() -> unsigned {
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
        if constexpr (MODE != MODE_GATHER) {
            const int sgi = lane >> 3, sub = lane & 7;
            const SegDesc& d = wt.D[slot][sgi];
            if (d.nk > 0) {
                const int Ls = d.nk + k - 1;
                const uintptr_t g0 = (uintptr_t)a.seq + (uintptr_t)d.start;
                const uintptr_t a0 = g0 & ~(uintptr_t)15;
                const int nchunks = (int)(((g0 + (uintptr_t)Ls + 15) & ~(uintptr_t)15) - a0) >> 4;
                unsigned char* dst = reinterpret_cast<unsigned char*>(wt.RAW + sgi * WPS) + 16;
                for (int c = sub; c < nchunks; c += 8)
                    cp_async16(dst + 16 * c, reinterpret_cast<const void*>(a0 + 16 * (uintptr_t)c));
            }
        }
    };

    if constexpr (MODE == MODE_GATHER) {
        // flag protocol: every owner stamps my reply buffer when all its replies to me have been stored
        if (a.epoch && lane < a.n_ranks) spin_until_equal(a.reply_stamp + lane, a.epoch, &a.scal->err);
        __syncwarp();
    }
    unsigned rq0 = claim_round(), rq1 = claim_round(), rq2 = claim_round();
    // prologue: descriptors of the first two rounds, raw bytes of the first
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
        const int cur = it % 3, nxt = (it + 1) % 3, nx2 = (it + 2) % 3;
        // top of the round: launch the descriptor copy for the round after next (the raw bytes of the
        // next round follow as soon as this round's staging has released the raw area)
        prefetch_desc(nx2, rq2);
        cp_async_commit();
        rq0 = rq1;
        rq1 = rq2;
        rq2 = claim_round();

        // ---- descriptors of THIS round (arrived a round ago) ----
        {
            int nk = 0;
            if (lane < SPW) {
                const SegDesc d = wt.D[cur][lane];
                nk = d.nk;
                wt.start[lane] = d.start;
                wt.seq[lane] = (long long)d.seq;
                wt.kout[lane] = d.kout;
                wt.nk[lane] = nk;
            }
            int incl = (nk + 3) >> 2;
#pragma unroll
            for (int d = 1; d < SPW; d <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane < SPW) wt.gpre[lane + 1] = incl;
            if (lane == 0) wt.gpre[0] = 0;
        }
        __syncwarp();

        // ---- stage forward + reverse-complement words from the raw copy (shared memory only) ----
        if constexpr (MODE != MODE_GATHER) {
            const int sg = lane >> 3, sub = lane & 7;
            const int nk = wt.nk[sg];
            if (nk) {
                const int Ls = nk + k - 1;
                const int r0 = (int)(((uintptr_t)a.seq + (uintptr_t)wt.start[sg]) & 15);
                const uint32_t* raw = wt.RAW + sg * WPS;
                uint32_t* Fs = wt.F + sg * WPS;
                uint32_t* Rs = wt.R + sg * WPS;
                const int p = sgc::rc_pad(nk);
                bool all_ok = true;
                for (int w = sub; 4 * w < Ls; w += 8) {
                    uint32_t mask;
                    Fs[w] = raw_window4(raw, r0, 4 * w, Ls, mask);
                }
                // R word m holds R[4m-4-p .. +4) = revcomp of F[Ls-4m+p .. +4)
                const int nrw = (Ls + 4 + p + 3) >> 2;
                for (int m = sub; m < nrw; m += 8) {
                    uint32_t mask;
                    const uint32_t u = raw_window4(raw, r0, Ls - 4 * m + p, Ls, mask);
                    bool ok;
                    Rs[m] = revcomp_word(u, mask, ok);
                    all_ok &= ok;
                }
                if (!all_ok && a.status) atomicOr(&a.status[wt.seq[sg]], 1);
                if constexpr (MODE == MODE_INSERT) {
                    // a unitig with a byte that is not A/C/G/T has no 2-bit form: nothing may be matched against
                    // this stretch of the sequence store (its k-mers are still found by their hashes)
                    if (!all_ok && a.brk)
                        for (int q = 0; q <= nk; q++) brk_mark(a.brk, (uint32_t)(a.upos_base + (unsigned long long)(wt.start[sg] + q)));
                }
            }
        }
        __syncwarp();
        prefetch_raw(nxt);  // lands while this round is hashed
        cp_async_commit();

        // ---- hash + sink ----
        const int G = wt.gpre[SPW];

        // locate group g (segment sg, first k-mer i, nv valid k-mers) and hash its 4 k-mers
        // ori (MODE_INSERT only): bit t = forward k-mer t hashes below its reverse complement
        constexpr bool WANT_ORI = MODE == MODE_INSERT || MODE == MODE_HASH;
        auto locate_and_hash = [&](int g, int& sg, int& i, int& nv, uint64_t (&h)[4], unsigned& ori) {
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
                for (int t = 0; t < 4; t++) {
                    if constexpr (WANT_ORI) {
                        bool fl;
                        h[t] = sgc::hash_group_kmer_ori<K ? K : 1>(fw, rw, t, fl);
                        ori |= (unsigned)fl << t;
                    } else {
                        h[t] = sgc::hash_group_kmer<K ? K : 1>(fw, rw, t);
                    }
                }
            } else {
                for (int t = 0; t < nv; t++) {
                    int q = i + t;
                    const uint64_t hf = sgc::murmur3_h1_stream(Fs, q, k);
                    const uint64_t hr = sgc::murmur3_h1_stream(Rs, sgc::rc_byte_of(nk, nk - 1 - q), k);
                    h[t] = hf ^ hr;
                    if constexpr (WANT_ORI) ori |= (unsigned)(hf < hr) << t;
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
                unsigned ori = 0;
                locate_and_hash(g, sg, i, nv, h, ori);
                if constexpr (MODE == MODE_HASH) {
                    uint64_t* o = a.hash_out + wt.kout[sg] + i;
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (t < nv) o[t] = h[t];
                    if (a.aux_out) {  // position of the k-mer in its sequence buffer | orientation (partitioned build)
                        const uint32_t p0 = (uint32_t)(a.upos_base + (unsigned long long)(wt.start[sg] + i));
#pragma unroll
                        for (int t = 0; t < 4; t++)
                            if (t < nv) a.aux_out[wt.kout[sg] + i + t] = (p0 + t) | (((ori >> t) & 1u) ? ORI_BIT : 0u);
                    }
                } else {
                    long long sq = wt.seq[sg];
                    uint32_t id = a.seq_ids ? a.seq_ids[sq] : a.id_base + (uint32_t)sq;
                    if (id > SGC_MAX_ID) atomicOr(&a.scal->err, 2);
                    else {
                        // position of the k-mer in the table's unitig sequence store = store offset of this batch
                        // + byte offset of its first base in the batch (sgc_labelseq.cuh matches reads against it)
                        const unsigned long long g0 = a.upos_base + (unsigned long long)(wt.start[sg] + i);
                        for (int t = 0; t < nv; t++) {
                            uint32_t aux = AUX_NONE;
                            if (a.brk) aux = (uint32_t)(g0 + t) | (((ori >> t) & 1u) ? ORI_BIT : 0u);
                            BrkSink sink;
                            sink.brk = a.brk;
                            table_insert(a.tbl, a.nbuckets, a.nlines, a.home_shift, h[t], id, aux, a.scal, sink);
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
                unsigned ori_unused = 0;
                if (active) locate_and_hash(g, sg, i, nv, h, ori_unused);
                if (lane < a.n_ranks) wt.part.wcnt[lane] = 0;
                __syncwarp();
                uint32_t own[4], rank[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    own[t] = a.owner_bits ? (uint32_t)(h[t] >> (64 - a.owner_bits)) : 0u;
                    rank[t] = (active && t < nv) ? atomicAdd(&wt.part.wcnt[own[t]], 1u) : 0u;
                }
                __syncwarp();
                {   // one lane per owner: reserve the run in the owner's region, prefix for the staging slots
                    const unsigned c = lane < a.n_ranks ? wt.part.wcnt[lane] : 0u;
                    unsigned incl = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    if (lane < a.n_ranks) {
                        wt.part.wpre[lane] = incl - c;
                        wt.part.wbase[lane] = c ? atomicAdd(&a.cursor[lane], (unsigned long long)c) : 0ULL;
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
                            const unsigned long long pos = wt.part.wbase[own[t]] + rank[t];
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
                                const unsigned long long pos = wt.part.wbase[own[t]] + rank[t];
                                if (pos < a.send_cap) a.peer_send[own[t]][a.send_region_off + pos] = h[t];
                            }
                        }
                    }
                } else {
                    if (active) {
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            if (t < nv) {
                                const unsigned slot = wt.part.wpre[own[t]] + rank[t];
                                const unsigned long long pos = wt.part.wbase[own[t]] + rank[t];
                                stage_h[slot] = h[t];
                                wt.part.dst[slot] = pos < a.send_cap ? a.peer_send[own[t]] + a.send_region_off + pos : nullptr;
                            }
                        }
                    }
                    __syncwarp();
                    const unsigned total = wt.part.wpre[a.n_ranks - 1] + wt.part.wcnt[a.n_ranks - 1];
                    for (unsigned j = lane; j < total; j += 32) {
                        uint64_t* d = wt.part.dst[j];
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
        } else {
            // MODE_LOOKUP / MODE_LABEL: every k-mer of a group probes its home bucket (4 loads in flight per
            // thread), resolved right away; the latency is covered by the other resident warps.
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

            // no cross-trip pipeline (ptxas serialises it anyway): hash, probe, resolve inside one trip
            for (int g0 = 0; g0 < G; g0 += 32) {
                const int g = g0 + lane;
                p_active = g < G;
                p_sg = 0; p_i = 0; p_nv = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) p_h[t] = 0;
                unsigned ori_unused = 0;
                if (p_active) locate_and_hash(g, p_sg, p_i, p_nv, p_h, ori_unused);
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    // lanes / k-mers with nothing to look up probe bucket 0 (an L2 hit): no branch around the loads
                    p_b[t] = (p_active && t < p_nv) ? home_bucket(p_h[t], a.nbuckets, a.home_shift) : 0u;
                    load_bucket_nc(a.tbl + p_b[t], p_e0[t], p_e1[t]);
                }
                resolve();
            }
            __syncwarp();
        }

        // end of the round: the copies launched at its top (next round's bytes, next-next descriptors)
        cp_async_wait_all();
        __syncwarp();
    }
    cp_async_wait_all();  // copies launched for rounds past the end (none carry data) must not outlive the CTA

    if constexpr (MODE == MODE_LABEL || MODE == MODE_GATHER) {
        if (pair_fill) flush_pairs();
    }
    if constexpr (MODE == MODE_PARTITION) {
        if (a.epoch) {
            // flag protocol: the LAST warp of the grid to get here knows that every hash of this rank has been
            // stored (each warp fences its peer stores before it counts itself); it publishes the per-owner
            // counts and then the epoch stamp into the owners' mailboxes
            __threadfence_system();
            unsigned prev = 0;
            if (lane == 0) prev = atomicAdd(&a.scal->done_ctr, 1u);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev == gridDim.x * (unsigned)WPB - 1u) {
                __threadfence();
                if (lane < a.n_ranks) {
                    const unsigned long long cnt = *reinterpret_cast<volatile unsigned long long*>(a.cursor + lane);
                    if (cnt > a.send_cap) atomicOr(&a.scal->err, 4);  // a region overflowed: the step is void
                    *reinterpret_cast<volatile unsigned long long*>(a.peer_mail_count[lane] + a.my_rank) = cnt;
                    __threadfence_system();
                    *reinterpret_cast<volatile unsigned int*>(a.peer_mail_stamp[lane] + a.my_rank) = a.epoch;
                }
            }
        }
    }
    if constexpr (MODE == MODE_LOOKUP || MODE == MODE_LABEL || MODE == MODE_GATHER) {
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


}  // namespace dev
}  // namespace sgc
