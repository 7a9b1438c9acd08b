// sgc_label8.cuh -- the read-labelling kernel (K1+K3+K4 fused: hash, table probe, per-read unitig ids)
// for tables that carry the unitig-ordered side array.  Part of libsgc_b200.so (device code; included
// by sgc_b200.cu only).  Reference semantics: cdbg/index_reads.py:143-152 (hash_sequence +
// count_cdbg_matches per read; only the KEYS of the returned dict are used).
//
// Layout of the work
//   * a lane owns a GROUP of 8 consecutive k-mers of one segment (<= 128 k-mers of one sequence); a warp
//     round is 4 segments = up to 64 groups = two full trips of the warp.
//   * only the FIRST k-mer of a group (the anchor) probes the hash table (one 128-byte HBM line).  Its
//     slot carries g, the k-mer's position in the table's unitig-ordered side array of 64-bit hashes,
//     and an orientation bit; the seven followers are compared with side[g+1..g+7] (read runs along
//     the unitig) or side[g-1..g-7] (against it): 56 contiguous bytes that neighbouring lanes share.
//   * a follower whose hash equals its side record inherits the anchor's id when no BREAK bit lies
//     between the two positions (brk[p] = "a table lookup of side[p] and of side[p-1] give different
//     answers": unitig starts, dropped multicount hashes, 64-bit collisions), so the inherited id is by
//     construction what a table lookup of that hash returns.  Everything else -- sequencing errors,
//     unitig ends, a missing anchor -- is queued per warp and probed densely (one item per lane).
//     The labels are therefore exactly those of per-k-mer probing (tile_kernel<K, MODE_LABEL>).
//   * PACKED = true: the reads arrive 2 bits per base (A=0 C=1 G=2 T=3, base b of a record in bits
//     2(b%4).. of byte b/4, every record 16-byte aligned) and are expanded to the ASCII words the
//     reference hash is defined on while they are staged (PRMT on "ACGT"); 0.3 instead of 1.25 bytes
//     per k-mer over PCIe and HBM.
#pragma once
#include "sgc_kmer.cuh"
#include "sgc_table.cuh"
#include "sgc_tile.cuh"

namespace sgc {
namespace dev {

constexpr int SEG_K8 = 128;   // k-mers per segment of the label8 plan
constexpr int SPW8 = 4;       // segments per warp round
constexpr int PAIR8 = 64;     // candidate keys staged per warp before one global reservation
#ifndef SGC_QCAP8
#define SGC_QCAP8 288
#endif
constexpr int QCAP8 = SGC_QCAP8;          // queued k-mers per warp; a trip adds at most 7 x 32 and takes 32 off
constexpr int QTHR8 = QCAP8 - 7 * 32 + 32;  // fill above which a trip might overflow the queue
constexpr int FOFF = 4;                   // staged forward words start 4 words into their slot
constexpr int SIDE_PAD = 8;               // records of padding on either end of the side array
constexpr int BRK_PAD = 32;               // bit index of position 0 in the break-bit array
static_assert(QCAP8 >= 7 * 32 + 32, "the queue must take a whole trip");
#ifndef SGC_L8_NT
#define SGC_L8_NT 256
#endif
#ifndef SGC_L8_MIN_CTAS
#define SGC_L8_MIN_CTAS 4
#endif
constexpr int NT8 = SGC_L8_NT;    // threads per CTA of the label8 kernel
constexpr int WPB8 = NT8 / 32;

template <int K>
struct Geo8 {
    static constexpr int KMAX = K ? K : SGC_MAX_K;
    static constexpr int SEGB = SEG_K8 + KMAX - 1;
    static constexpr int STRIDE = ((SEGB + 32 + 15) / 16) * 16;
    static constexpr int WPS = STRIDE / 4;
};

struct Label8Args {
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
    const unsigned long long* side8;  // position 0 (SIDE_PAD records of padding before it)
    const uint32_t* brk;              // bit BRK_PAD + p = break before position p
    unsigned long long* pairs;
    unsigned long long pairs_cap;
};

template <int K>
struct alignas(16) WarpTile8 {
    uint32_t F[SPW8 * Geo8<K>::WPS];
    uint32_t R[SPW8 * Geo8<K>::WPS];
    uint32_t RAW[SPW8 * Geo8<K>::WPS];
    SegDesc D[3][SPW8];
    unsigned long long pairs[PAIR8];
    unsigned long long qh[QCAP8];
    uint32_t qs[QCAP8];
    int gpre[SPW8 + 1];
    unsigned qfill;
    unsigned pad_[2];
};

// the walk past a full home bucket, out of line: one copy of its code for the anchor and the queue probes
__device__ __noinline__ uint32_t lookup_slow_call(const Bucket* tbl, uint32_t nlines, uint32_t hb, uint64_t h, uint32_t* aux) {
    uint32_t ax = AUX_NONE;
    const uint32_t v = lookup_slow(tbl, nlines, hb, h, ax);
    *aux = ax;
    return v;
}

__device__ __forceinline__ unsigned long long ldg_nc_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
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

template <int K, bool PACKED>
__global__ void __launch_bounds__(NT8, SGC_L8_MIN_CTAS) label8_kernel(const Label8Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WPS = Geo8<K>::WPS;
    const int lane = threadIdx.x & 31;
    using Tile = WarpTile8<K>;
    Tile& wt = reinterpret_cast<Tile*>(smem_raw)[threadIdx.x >> 5];
    const int k = K ? K : a.k;
    const int64_t nsegs = *a.nsegs_p;
    const int64_t nrounds = (nsegs + SPW8 - 1) / SPW8;
    unsigned my_hits = 0;                // per lane: a launch is far below 2^32 k-mers per lane
    unsigned pair_fill = 0;              // warp-uniform: keys waiting in wt.pairs
    int cur = 0;                         // descriptor slot of the round being worked on

    auto flush_pairs = [&]() {  // cold: out of line
        flush_pairs_out(a.scal, a.pairs, a.pairs_cap, wt.pairs, pair_fill, lane);
        pair_fill = 0;
    };
    // one candidate (sequence, id) key per lane at most; the order of the keys is irrelevant (they are
    // grouped per sequence by the label tail)
    auto push_key = [&](bool has, unsigned long long key) {
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (m) {
            const unsigned total = __popc(m);
            if (pair_fill + total > PAIR8) flush_pairs();
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
        if (lane < 2 * SPW8) {
            const int sgi = lane >> 1, half = lane & 1;
            const int64_t g = (int64_t)round * SPW8 + sgi;
            unsigned char* dst = reinterpret_cast<unsigned char*>(&wt.D[slot][sgi]) + 16 * half;
            if ((int64_t)round < nrounds && g < nsegs) {
                cp_async16(dst, reinterpret_cast<const unsigned char*>(a.segdesc + g) + 16 * half);
            } else {
                *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);  // nk = 0: nothing to do
            }
        }
    };
    auto prefetch_raw = [&](int slot) {
        const int sgi = lane >> 3, sub = lane & 7;
        const SegDesc& d = wt.D[slot][sgi];
        if (d.nk > 0) {
            const int Ls = d.nk + k - 1;
            if constexpr (PACKED) {  // records are 16-byte aligned and a segment starts 32 bytes after the last
                const int nchunks = (((Ls + 3) >> 2) + 15) >> 4;
                unsigned char* dst = reinterpret_cast<unsigned char*>(wt.RAW + sgi * WPS);
                const unsigned char* src = a.seq + d.start;
                for (int c = sub; c < nchunks; c += 8) cp_async16(dst + 16 * c, src + 16 * c);
            } else {
                const uintptr_t g0 = (uintptr_t)a.seq + (uintptr_t)d.start;
                const uintptr_t a0 = g0 & ~(uintptr_t)15;
                const int nchunks = (int)(((g0 + (uintptr_t)Ls + 15) & ~(uintptr_t)15) - a0) >> 4;
                unsigned char* dst = reinterpret_cast<unsigned char*>(wt.RAW + sgi * WPS) + 16;
                for (int c = sub; c < nchunks; c += 8)
                    cp_async16(dst + 16 * c, reinterpret_cast<const void*>(a0 + 16 * (uintptr_t)c));
            }
        }
    };

    // One queued k-mer resolved: its home bucket has arrived.  A hit becomes a (sequence, id) key unless the
    // previous lane's item gives the same key (neighbouring items are neighbouring k-mers of one read).
    auto settle = [&](bool act, uint64_t h, uint32_t sq, uint32_t b, const Entry& e0, const Entry& e1) {
        uint32_t id = SGC_MISS;
        if (act) {
            uint32_t v, aux;
            if (bucket_match(e0, e1, h, v)) id = v;
            else if (bucket_overflowed(e0)) id = lookup_slow_call(a.tbl, a.nlines, b, h, &aux);
            if (id != SGC_MISS) my_hits++;
        }
        const unsigned long long key = ((unsigned long long)sq << 32) | id;
        const unsigned long long prevk = __shfl_up_sync(0xffffffffu, key, 1);
        push_key(act && id != SGC_MISS && !(lane > 0 && prevk == key), key);
    };
    // Stand-alone drain (rare: the trips themselves take 32 items off the queue each, overlapped with their
    // side-array loads): probe the top of the queue, two items per lane in flight, until <= keep are left.
    auto drain_down = [&](unsigned keep) {
        for (;;) {
            const unsigned n = wt.qfill;
            if (n <= keep) break;
            const unsigned npop = n < 64u ? n : 64u, base = n - npop;
            const unsigned j0 = base + lane, j1 = base + 32u + lane;
            const bool act0 = j0 < n, act1 = j1 < n;
            const uint64_t h0 = act0 ? wt.qh[j0] : 0ULL, h1 = act1 ? wt.qh[j1] : 0ULL;
            const uint32_t s0 = act0 ? wt.qs[j0] : 0u, s1 = act1 ? wt.qs[j1] : 0u;
            const uint32_t b0 = act0 ? home_bucket(h0, a.nbuckets, a.home_shift) : 0u;
            const uint32_t b1 = act1 ? home_bucket(h1, a.nbuckets, a.home_shift) : 0u;
            Entry x0, x1, y0, y1;
            load_bucket_nc(a.tbl + b0, x0, x1);
            load_bucket_nc(a.tbl + b1, y0, y1);
            __syncwarp();
            if (lane == 0) wt.qfill = base;
            settle(act0, h0, s0, b0, x0, x1);
            settle(act1, h1, s1, b1, y0, y1);
            __syncwarp();
        }
    };

    if (lane == 0) wt.qfill = 0;
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
        cur = it % 3;
        const int nxt = (it + 1) % 3, nx2 = (it + 2) % 3;
        prefetch_desc(nx2, rq2);  // descriptors of the round after next
        cp_async_commit();
        rq0 = rq1;
        rq1 = rq2;
        rq2 = claim_round();

        // ---- group prefix of this round's segments ----
        {
            const int nk = lane < SPW8 ? wt.D[cur][lane].nk : 0;
            int incl = (nk + 7) >> 3;
#pragma unroll
            for (int d = 1; d < SPW8; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane < SPW8) wt.gpre[lane + 1] = incl;
            if (lane == 0) wt.gpre[0] = 0;
        }

        // ---- stage forward + reverse-complement words (8 lanes per segment, shared memory only) ----
        // Forward words first (re-aligned raw bytes, or expanded 2-bit codes), then the reverse complement
        // from them.  Bytes outside the segment are whatever lies next to it: no valid k-mer ever reads them
        // (a k-mer's window is exactly its k bytes), so only the validity test needs a mask.
        {
            const int sg = lane >> 3, sub = lane & 7;
            const SegDesc& d = wt.D[cur][sg];
            const int nk = d.nk;
            const int Ls = nk + k - 1;
            uint32_t* Fs = wt.F + sg * WPS + FOFF;
            if (nk) {
                if constexpr (PACKED) {
                    const uint32_t* pk = wt.RAW + sg * WPS;
                    const int npw = (Ls + 15) >> 4;
                    for (int j = sub; j < npw; j += 8) *reinterpret_cast<uint4*>(Fs + 4 * j) = expand_packed16(pk[j]);
                } else {
                    const int r0 = (int)(((uintptr_t)a.seq + (uintptr_t)d.start) & 15);
                    const int s = r0 & 3;
                    const uint32_t* rw0 = wt.RAW + sg * WPS + 4 + (r0 >> 2);  // raw copies start 16 bytes into their slot
                    uint32_t bad = 0;
                    for (int w = sub; 4 * w < Ls; w += 8) {
                        const uint32_t u = sgc::funnel_bytes(rw0[w], rw0[w + 1], s);
                        Fs[w] = u;
                        // every byte must be one of A C G T: "ACTG"[bits 1..2 of the byte] == the byte
                        const uint32_t code = (u >> 1) & 0x03030303u;
                        const uint32_t sel = __byte_perm(code | (code >> 4), 0u, 0x4420u);
                        const int nval = Ls - 4 * w;
                        const uint32_t mask = nval >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nval)) - 1u);
                        bad |= (__byte_perm(0x47544341u, 0u, sel) ^ u) & mask;
                    }
                    if (bad && a.status) atomicOr(&a.status[d.seq], 1);
                }
            }
            __syncwarp();
            if (nk) {
                // R word m holds R[4m-4-p .. +4) = reverse complement of F[Ls-4m+p .. +4)
                const int p = sgc::rc_pad(nk);
                const int nrw = (Ls + 4 + p + 3) >> 2;
                uint32_t* Rs = wt.R + sg * WPS;
                const uint32_t* Fw = wt.F + sg * WPS;  // F byte f lives at byte 16 + f of this slot (f >= -7 here)
                for (int m = sub; m < nrw; m += 8) {
                    const int b = 16 + Ls - 4 * m + p;
                    const uint32_t u = sgc::funnel_bytes(Fw[b >> 2], Fw[(b >> 2) + 1], b & 3);
                    // complement and reverse in one PRMT: nibbles of the selector = the 2-bit codes, last base first
                    const uint32_t code = (u >> 1) & 0x03030303u;
                    const uint32_t sel = __byte_perm(code | (code << 12), 0u, 0x4413u);
                    Rs[m] = __byte_perm(0x43414754u, 0u, sel);  // "TGAC"[code]
                }
            }
        }
        __syncwarp();
        prefetch_raw(nxt);  // lands while this round is hashed
        cp_async_commit();

        // ---- hash + probe + match ----
        // warp-uniform state is re-read from shared memory where it is used (volatile: not hoisted), so that it
        // does not sit in registers across the hash block (ptxas spilled exactly these to local memory)
        const volatile int* gpre = wt.gpre;
        for (int g0 = 0; g0 < gpre[SPW8]; g0 += 32) {
            if (wt.qfill > (unsigned)QTHR8) drain_down(32u);  // make room for the worst case of this trip
            const int g = g0 + lane;
            const bool active = g < gpre[SPW8];
            const int sg = (g >= gpre[1]) + (g >= gpre[2]) + (g >= gpre[3]);
            const SegDesc& d = wt.D[cur][active ? sg : 0];
            const int nk = d.nk;
            const int i = active ? (g - gpre[sg]) << 3 : 0;
            const int nv = active ? (nk - i < 8 ? nk - i : 8) : 0;
            const uint32_t* Fs = wt.F + sg * WPS + FOFF;
            const uint32_t* Rs = wt.R + sg * WPS;
            uint64_t h[8];
            bool fl = false;  // the anchor's forward k-mer hashes below its reverse complement
            uint32_t hb = 0;
            Entry e0, e1;
            if constexpr (K != 0) {
                constexpr int NWIN = sgc::KW<K ? K : 1>::NWIN;
                // k-mers i..i+3 (sub-group 0) use forward words [0, NWIN) of fp and RC words [0, NWIN) of rp;
                // k-mers i+4..i+7 (sub-group 1) start one forward word later and one RC word earlier.  The two
                // sub-groups run as a real loop: eight unrolled hashes (24 KB of code) do not fit the
                // instruction cache next to the rest of the trip (measured: icc hit rate 78 %, 3 stall cycles
                // per issue in no_instruction).  The anchor's probe leaves after the first sub-group and is
                // in flight while the second is hashed.
                const uint32_t* fp = Fs + (i >> 2);
                const uint32_t* rp = Rs + sgc::rc_group_word(nk, i);
#pragma unroll 1
                for (int sgrp = 0; sgrp < 2; sgrp++) {
                    uint32_t fw[NWIN], rw[NWIN];
#pragma unroll
                    for (int j = 0; j < NWIN; j++) {
                        fw[j] = fp[j + sgrp];
                        rw[j] = rp[j - sgrp];
                    }
                    uint64_t hh[4];
                    bool f0;
                    hh[0] = sgc::hash_group_kmer_ori<K ? K : 1>(fw, rw, 0, f0);
#pragma unroll
                    for (int t = 1; t < 4; t++) hh[t] = sgc::hash_group_kmer<K ? K : 1>(fw, rw, t);
                    if (sgrp == 0) {
                        fl = f0;
#pragma unroll
                        for (int t = 0; t < 4; t++) h[t] = hh[t];
                        hb = active ? home_bucket(h[0], a.nbuckets, a.home_shift) : 0u;
                        load_bucket_nc(a.tbl + hb, e0, e1);  // idle lanes probe bucket 0 (an L2 hit)
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; t++) h[4 + t] = hh[t];
                    }
                }
            } else {
#pragma unroll 1
                for (int t = 0; t < 8; t++) {
                    h[t] = 0;
                    if (t < nv) {
                        const int q = i + t;
                        const uint64_t hf = sgc::murmur3_h1_stream(Fs, q, k);
                        const uint64_t hr = sgc::murmur3_h1_stream(Rs, sgc::rc_byte_of(nk, nk - 1 - q), k);
                        h[t] = hf ^ hr;
                        if (t == 0) fl = hf < hr;
                    }
                }
                hb = active ? home_bucket(h[0], a.nbuckets, a.home_shift) : 0u;
                load_bucket_nc(a.tbl + hb, e0, e1);
            }

            // anchor
            uint32_t v = SGC_MISS, aux = AUX_NONE;
            if (active) {
                if (!bucket_match(e0, e1, h[0], v, aux)) {
                    v = SGC_MISS;
                    aux = AUX_NONE;
                    if (bucket_overflowed(e0)) v = lookup_slow_call(a.tbl, a.nlines, hb, h[0], &aux);
                }
            }
            // followers: their side records and break bits leave now ...
            const uint32_t gp = aux & AUX_MASK;
            const bool has_g = gp != AUX_NONE;
            const uint32_t gq = has_g ? gp : 0u;
            const bool along = ((aux & ORI_BIT) != 0) == fl;
            const char* sp = reinterpret_cast<const char*>(a.side8 + gq);
            const int stepb = along ? 8 : -8;  // bytes per follower
            unsigned long long rec[7];
#pragma unroll
            for (int dd = 1; dd < 8; dd++)
                rec[dd - 1] = ldg_nc_u64(reinterpret_cast<const unsigned long long*>(sp + (long long)(dd * stepb)));
            const uint32_t bi = gq + (uint32_t)(BRK_PAD - 6);  // break bits of positions gq-6 .. gq+7
            const uint32_t w0 = __ldg(a.brk + (bi >> 5)), w1 = __ldg(a.brk + (bi >> 5) + 1);
            // ... together with the table probe of one queued k-mer per lane (the top 32 of the queue), so that a
            // trip waits for memory once
            const unsigned qn = wt.qfill;
            const unsigned qpop = qn < 32u ? qn : 32u, qbase = qn - qpop;
            const bool qact = lane < qpop;
            const uint64_t qhash = qact ? wt.qh[qbase + lane] : 0ULL;
            const uint32_t qseq = qact ? wt.qs[qbase + lane] : 0u;
            const uint32_t qb = qact ? home_bucket(qhash, a.nbuckets, a.home_shift) : 0u;
            Entry q0, q1;
            load_bucket_nc(a.tbl + qb, q0, q1);
            __syncwarp();
            if (lane == 0) wt.qfill = qbase;
            __syncwarp();

            const uint32_t bits = __funnelshift_r(w0, w1, bi & 31u);
            // followers 1..lim may inherit: no break in (g, g+d] resp. (g-d, g]
            unsigned lim = along ? (unsigned)__ffs((bits >> 7) | 0x80u) - 1u
                                 : (unsigned)__clz((int)(((bits & 0x7Fu) << 25) | (1u << 24)));
            if (!has_g) lim = 0;
            unsigned m = 0;
#pragma unroll
            for (int dd = 1; dd < 8; dd++) m |= (h[dd] == rec[dd - 1]) ? (1u << dd) : 0u;
            const unsigned valid = (1u << nv) - 1u;
            m &= ((2u << lim) - 1u) & valid;
            const unsigned pending = valid & ~1u & ~m;
            if (v != SGC_MISS) my_hits += 1u + (unsigned)__popc(m);

            if (pending) {  // queue what the side array did not settle
                unsigned pos = atomicAdd(&wt.qfill, (unsigned)__popc(pending));
#pragma unroll
                for (int dd = 1; dd < 8; dd++) {
                    if ((pending >> dd) & 1u) {
                        wt.qh[pos] = h[dd];
                        wt.qs[pos] = d.seq;
                        pos++;
                    }
                }
            }
            settle(qact, qhash, qseq, qb, q0, q1);
            // the anchor's id is a candidate label unless the previous group of the same segment has it too
            const uint32_t prev_v = __shfl_up_sync(0xffffffffu, v, 1);
            push_key(v != SGC_MISS && !(lane > 0 && i > 0 && prev_v == v), ((unsigned long long)d.seq << 32) | v);
            __syncwarp();
        }

        cp_async_wait_all();  // next round's bytes, next-next descriptors
        __syncwarp();
    }
    cp_async_wait_all();
    drain_down(0u);

    if (pair_fill) flush_pairs();
    unsigned long long hits = my_hits;  // the host derives the misses from the plan's k-mer total
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, dd);
    if (lane == 0 && hits) atomicAdd(&a.scal->n_hits, hits);
}

}  // namespace dev
}  // namespace sgc
