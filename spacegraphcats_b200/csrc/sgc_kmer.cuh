// sgc_kmer.cuh -- k-mer window extraction and the reference hash, shared by the
// sm_100a kernels (device) and the host-side self-test (tests/ compile this header
// with g++ to emulate one thread of the tile kernel on the CPU; no GPU needed).
//
// The hash reproduced bit-for-bit is the one the reference obtains from
//   khmer.Nodetable(k,1,1).get_kmer_hashes(seq)      (spacegraphcats/cdbg/hashing.py:14-20)
// i.e.  h = MurmurHash3_x64_128(kmer, k, seed 0).lo64 ^ MurmurHash3_x64_128(revcomp(kmer), k, 0).lo64
//
// Layout contract used by the tile kernel (see sgc_b200.cu):
//   * a "segment" is a run of nk <= SEG_K consecutive k-mers of one sequence, i.e.
//     Ls = nk + k - 1 bases.  Its ASCII bytes are staged at a 16-byte aligned
//     shared-memory address F[0..Ls).
//   * its reverse complement R[j] = comp(F[Ls-1-j]) is staged at byte offset
//     4 + p of a second 16-byte aligned area, p = (-nk) mod 4, so that the four RC
//     windows of a 4-aligned k-mer group start inside one aligned 32-bit word.
//   * a thread owns one "group": the 4 consecutive k-mers i..i+3 (i % 4 == 0) of a
//     segment; k-mer i+t reads its forward window at byte shift t of the words at
//     F32[i/4 ..] and its RC window at byte shift 3-t of R32[(p + nk - i)/4 ..].
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SGC_HD __host__ __device__ __forceinline__
#else
#define SGC_HD inline
#endif

namespace sgc {

constexpr uint64_t MM_C1 = 0x87c37b91114253d5ULL;
constexpr uint64_t MM_C2 = 0x4cf5ad432745937fULL;

// On the device the 64-bit rotate and the 64-bit multiply by a constant are spelled as
// opaque PTX (two funnel shifts; one mul.wide + two mad.lo).  Left to itself nvcc distributes
// rotl(a*c1, r)*c2 into sums of products and spends ~25 % more integer instructions per k-mer
// (measured with cuobjdump: 248 -> 186 SASS instructions per canonical k=31 hash).
SGC_HD uint64_t rotl64(uint64_t x, int r) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32), olo, ohi;
    if (r >= 32) { uint32_t t = lo; lo = hi; hi = t; r -= 32; }
    asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(ohi) : "r"(lo), "r"(hi), "r"(r));
    asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(olo) : "r"(hi), "r"(lo), "r"(r));
    return ((uint64_t)ohi << 32) | olo;
#else
    return (x << r) | (x >> (64 - r));
#endif
}

SGC_HD uint64_t mul64(uint64_t a, uint64_t c) {
#if defined(__CUDA_ARCH__)
    const uint32_t alo = (uint32_t)a, ahi = (uint32_t)(a >> 32);
    const uint32_t clo = (uint32_t)c, chi = (uint32_t)(c >> 32);
    uint64_t r;
    uint32_t t;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(alo), "r"(clo));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(alo), "r"(chi), "r"((uint32_t)(r >> 32)));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(ahi), "r"(clo), "r"(t));
    return ((uint64_t)t << 32) | (uint32_t)r;
#else
    return a * c;
#endif
}

SGC_HD uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k = mul64(k, 0xff51afd7ed558ccdULL);
    k ^= k >> 33;
    k = mul64(k, 0xc4ceb9fe1a85ec53ULL);
    k ^= k >> 33;
    return k;
}

SGC_HD uint64_t mix_k1(uint64_t k1) { k1 = mul64(k1, MM_C1); k1 = rotl64(k1, 31); return mul64(k1, MM_C2); }
SGC_HD uint64_t mix_k2(uint64_t k2) { k2 = mul64(k2, MM_C2); k2 = rotl64(k2, 33); return mul64(k2, MM_C1); }

// bytes [8*shift .. 8*shift+4) of the 8-byte little-endian pair (lo, hi); shift in 0..3
SGC_HD uint32_t funnel_bytes(uint32_t lo, uint32_t hi, int shift) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, 8 * shift);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (8 * shift));
#endif
}

SGC_HD uint64_t u64_of(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// ---------------------------------------------------------------------------
// MurmurHash3_x64_128(key, len = K, seed = 0).h1 for a key given as ceil(K/4)
// little-endian 32-bit words whose bytes beyond K are ZERO.  K is a compile-time
// constant so that the block loop and the tail are fully unrolled.
// ---------------------------------------------------------------------------
template <int K>
SGC_HD uint64_t murmur3_h1_words(const uint32_t* w, uint64_t seed = 0) {
    constexpr int NB = K / 16;
    constexpr int TAIL = K % 16;
    uint64_t h1 = seed, h2 = seed;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        uint64_t k1 = u64_of(w[4 * b], w[4 * b + 1]);
        uint64_t k2 = u64_of(w[4 * b + 2], w[4 * b + 3]);
        h1 ^= mix_k1(k1);
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        h2 ^= mix_k2(k2);
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    if constexpr (TAIL > 8) {
        uint32_t hi = 0u;
        if constexpr (TAIL > 12) hi = w[4 * NB + 3];
        h2 ^= mix_k2(u64_of(w[4 * NB + 2], hi));
    }
    if constexpr (TAIL > 0) {
        uint32_t hi = 0u;
        if constexpr (TAIL > 4) hi = w[4 * NB + 1];
        h1 ^= mix_k1(u64_of(w[4 * NB], hi));
    }
    h1 ^= (uint64_t)K; h2 ^= (uint64_t)K;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

// number of 32-bit words covering K bytes, and covering K+3 bytes (window + shift)
template <int K> struct KW {
    static constexpr int NWK = (K + 3) / 4;
    static constexpr int NWIN = (K + 3 + 3) / 4;
    static constexpr uint32_t LASTMASK = (K % 4) ? ((1u << (8 * (K % 4))) - 1u) : 0xFFFFFFFFu;
};

// Extract the K-byte window that starts `shift` bytes (0..3) into win[0..NWIN) and hash it.
template <int K>
SGC_HD uint64_t murmur3_h1_window(const uint32_t* win, int shift) {
    constexpr int NWK = KW<K>::NWK;
    constexpr int NWIN = KW<K>::NWIN;
    uint32_t w[NWK];
#pragma unroll
    for (int j = 0; j < NWK; j++) {
        uint32_t hi = (j + 1 < NWIN) ? win[j + 1] : 0u;
        w[j] = funnel_bytes(win[j], hi, shift);
    }
    w[NWK - 1] &= KW<K>::LASTMASK;
    return murmur3_h1_words<K>(w);
}

SGC_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0u, 0x0123u);
#else
    return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24);
#endif
}

// sourmash-style DNA k-mer hash: MurmurHash3_x64_128(min(kmer, revcomp(kmer)), K, seed).h1 with the
// minimum taken lexicographically on the ASCII strings (sourmash add_sequence / MinHash; used by
// cdbg/sort_bcalm_unitigs.py:56-128 with seed 42, n=1).  Same window convention as hash_group_kmer.
template <int K>
SGC_HD uint64_t minhash_group_kmer(const uint32_t* fwin, const uint32_t* rwin, int t, uint64_t seed) {
    constexpr int NWK = KW<K>::NWK;
    constexpr int NWIN = KW<K>::NWIN;
    uint32_t f[NWK], r[NWK];
#pragma unroll
    for (int j = 0; j < NWK; j++) {
        f[j] = funnel_bytes(fwin[j], (j + 1 < NWIN) ? fwin[j + 1] : 0u, t);
        r[j] = funnel_bytes(rwin[j], (j + 1 < NWIN) ? rwin[j + 1] : 0u, 3 - t);
    }
    f[NWK - 1] &= KW<K>::LASTMASK;
    r[NWK - 1] &= KW<K>::LASTMASK;
    bool r_less = false;  // scan from the last word to the first: the earliest difference decides
#pragma unroll
    for (int j = NWK - 1; j >= 0; j--) {
        const uint32_t a = bswap32(f[j]), b = bswap32(r[j]);
        r_less = (b < a) || (b == a && r_less);
    }
    uint32_t w[NWK];
#pragma unroll
    for (int j = 0; j < NWK; j++) w[j] = r_less ? r[j] : f[j];
    return murmur3_h1_words<K>(w, seed);
}

// Hash of k-mer t (0..3) of a group: fwin = forward words at the group's aligned
// start, rwin = RC words at the group's aligned RC start (see header comment).
template <int K>
SGC_HD uint64_t hash_group_kmer(const uint32_t* fwin, const uint32_t* rwin, int t) {
    return murmur3_h1_window<K>(fwin, t) ^ murmur3_h1_window<K>(rwin, 3 - t);
}
// ... and whether the forward k-mer hashes below its reverse complement (the side array's orientation bit)
template <int K>
SGC_HD uint64_t hash_group_kmer_ori(const uint32_t* fwin, const uint32_t* rwin, int t, bool& fwd_less) {
    const uint64_t hf = murmur3_h1_window<K>(fwin, t), hr = murmur3_h1_window<K>(rwin, 3 - t);
    fwd_less = hf < hr;
    return hf ^ hr;
}

// ---------------------------------------------------------------------------
// Runtime-k path (any 1 <= k <= 255): stream the key out of a 4-byte aligned
// word array starting at an arbitrary byte offset.
// ---------------------------------------------------------------------------
SGC_HD uint64_t load_le64_at(const uint32_t* base, int byteoff) {
    int idx = byteoff >> 2, s = byteoff & 3;
    uint32_t a = base[idx], b = base[idx + 1], c = base[idx + 2];
    return u64_of(funnel_bytes(a, b, s), funnel_bytes(b, c, s));
}

SGC_HD uint64_t keep_low_bytes(uint64_t v, int n /*1..8*/) {
    return n >= 8 ? v : (v & ((1ULL << (8 * n)) - 1ULL));
}

SGC_HD uint64_t murmur3_h1_stream(const uint32_t* base, int byteoff, int k, uint64_t seed = 0) {
    const int nb = k >> 4, tail = k & 15;
    uint64_t h1 = seed, h2 = seed;
    for (int b = 0; b < nb; b++) {
        uint64_t k1 = load_le64_at(base, byteoff + 16 * b);
        uint64_t k2 = load_le64_at(base, byteoff + 16 * b + 8);
        h1 ^= mix_k1(k1);
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        h2 ^= mix_k2(k2);
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    if (tail > 8) h2 ^= mix_k2(keep_low_bytes(load_le64_at(base, byteoff + 16 * nb + 8), tail - 8));
    if (tail > 0) h1 ^= mix_k1(keep_low_bytes(load_le64_at(base, byteoff + 16 * nb), tail > 8 ? 8 : tail));
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

// true when the k bytes at (rbase, roff) are lexicographically smaller than those at (fbase, foff)
SGC_HD bool stream_less(const uint32_t* rbase, int roff, const uint32_t* fbase, int foff, int k) {
    for (int j = 0; j < k; j += 8) {
        const int n = k - j < 8 ? k - j : 8;
        const uint64_t a = keep_low_bytes(load_le64_at(rbase, roff + j), n);
        const uint64_t b = keep_low_bytes(load_le64_at(fbase, foff + j), n);
        if (a != b) {
            const uint64_t d = a ^ b;  // lowest differing byte = earliest string position
            const uint64_t low = d & (~d + 1);
            int sh = 0;
            while (!((low >> sh) & 0xFFULL)) sh += 8;
            return ((a >> sh) & 0xFFULL) < ((b >> sh) & 0xFFULL);
        }
    }
    return false;
}

// khmer _revcomp base map; every byte other than A,C,G,T maps to itself (that is how
// an invalid base is recognised: comp(x) == x).  Policy for such bytes is host-side.
SGC_HD uint8_t comp_base(uint8_t b) {
    return b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : b;
}

// Geometry of one staged segment (all byte counts).
SGC_HD int rc_pad(int nk) { return (4 - (nk & 3)) & 3; }               // p
SGC_HD int rc_group_word(int nk, int i) { return (rc_pad(nk) + nk - i) >> 2; }  // word index of the RC window block
SGC_HD int rc_byte_of(int nk, int j) { return 4 + rc_pad(nk) + j; }     // where R[j] lives

// One 32-bit word (index m) of a segment's staged RC area, built from its forward
// bytes F8[0..Ls).  Bytes outside R[0..Ls) are zero.  *invalid is OR-ed with
// "some base here is not A/C/G/T".
SGC_HD uint32_t rc_word(const uint8_t* F8, int Ls, int nk, int m, bool* invalid) {
    const int j0 = 4 * m - 4 - rc_pad(nk);
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        int j = j0 + b;
        if (j >= 0 && j < Ls) {
            uint8_t x = F8[Ls - 1 - j];
            uint8_t c = comp_base(x);
            if (c == x) *invalid = true;
            r |= (uint32_t)c << (8 * b);
        }
    }
    return r;
}

}  // namespace sgc
