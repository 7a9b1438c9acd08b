// CPU emulation of ONE thread of the sm_100a tile kernel's staging + hashing
// geometry (spacegraphcats_b200/csrc/sgc_kmer.cuh), so the window/RC alignment
// arithmetic can be checked against the oracle without a GPU.  Test-only.
#include <cstring>
#include <vector>
#include "../../spacegraphcats_b200/csrc/sgc_kmer.cuh"

using namespace sgc;

template <int K>
static void hash_segment_fixed(const uint32_t* F32, const uint32_t* R32, int nk, uint64_t* out) {
    for (int i = 0; i < nk; i += 4) {
        const uint32_t* fwin = F32 + (i >> 2);
        const uint32_t* rwin = R32 + rc_group_word(nk, i);
        for (int t = 0; t < 4 && i + t < nk; t++) out[i + t] = hash_group_kmer<K>(fwin, rwin, t);
    }
}

extern "C" int emul_hash_sequence(const uint8_t* seq, long L, int k, int seg_k, uint64_t* out, int* invalid_out) {
    if (L < k) return -1;
    long nkseq = L - k + 1;
    bool invalid = false;
    for (long q0 = 0; q0 < nkseq; q0 += seg_k) {
        int nk = (int)((nkseq - q0) < seg_k ? (nkseq - q0) : seg_k);
        int Ls = nk + k - 1;
        int stride = ((seg_k + k - 1 + 16 + 15) / 16) * 16 + 16;
        // poison the staging areas: anything the kernel must not depend on is 0xEE
        std::vector<uint32_t> F32(stride / 4 + 4, 0xEEEEEEEEu), R32(stride / 4 + 4, 0xEEEEEEEEu);
        uint8_t* F8 = (uint8_t*)F32.data();
        std::memcpy(F8, seq + q0, Ls);
        int nrw = (4 + 3 + Ls + 3) / 4 + 1;
        for (int m = 0; m < nrw; m++) R32[m] = rc_word(F8, Ls, nk, m, &invalid);
        if (k == 21) hash_segment_fixed<21>(F32.data(), R32.data(), nk, out + q0);
        else if (k == 31) hash_segment_fixed<31>(F32.data(), R32.data(), nk, out + q0);
        else if (k == 32) hash_segment_fixed<32>(F32.data(), R32.data(), nk, out + q0);
        else if (k == 5) hash_segment_fixed<5>(F32.data(), R32.data(), nk, out + q0);
        else {
            for (int q = 0; q < nk; q++)
                out[q0 + q] = murmur3_h1_stream(F32.data(), q, k) ^
                              murmur3_h1_stream(R32.data(), rc_byte_of(nk, nk - 1 - q), k);
        }
    }
    if (invalid_out) *invalid_out = invalid ? 1 : 0;
    return 0;
}

// sourmash-style hash (min(kmer, rc), seeded Murmur3) through the same staged geometry
template <int K>
static void minhash_segment_fixed(const uint32_t* F32, const uint32_t* R32, int nk, uint64_t seed, uint64_t* out) {
    for (int i = 0; i < nk; i += 4) {
        const uint32_t* fwin = F32 + (i >> 2);
        const uint32_t* rwin = R32 + rc_group_word(nk, i);
        for (int t = 0; t < 4 && i + t < nk; t++) out[i + t] = minhash_group_kmer<K>(fwin, rwin, t, seed);
    }
}

extern "C" int emul_minhash_sequence(const uint8_t* seq, long L, int k, int seg_k, uint64_t seed, uint64_t* out) {
    if (L < k) return -1;
    long nkseq = L - k + 1;
    bool invalid = false;
    for (long q0 = 0; q0 < nkseq; q0 += seg_k) {
        int nk = (int)((nkseq - q0) < seg_k ? (nkseq - q0) : seg_k);
        int Ls = nk + k - 1;
        int stride = ((seg_k + k - 1 + 16 + 15) / 16) * 16 + 16;
        std::vector<uint32_t> F32(stride / 4 + 4, 0xEEEEEEEEu), R32(stride / 4 + 4, 0xEEEEEEEEu);
        uint8_t* F8 = (uint8_t*)F32.data();
        std::memcpy(F8, seq + q0, Ls);
        int nrw = (4 + 3 + Ls + 3) / 4 + 1;
        for (int m = 0; m < nrw; m++) R32[m] = rc_word(F8, Ls, nk, m, &invalid);
        if (k == 21) minhash_segment_fixed<21>(F32.data(), R32.data(), nk, seed, out + q0);
        else if (k == 31) minhash_segment_fixed<31>(F32.data(), R32.data(), nk, seed, out + q0);
        else {
            for (int q = 0; q < nk; q++) {
                int rq = rc_byte_of(nk, nk - 1 - q);
                out[q0 + q] = stream_less(R32.data(), rq, F32.data(), q, k) ? murmur3_h1_stream(R32.data(), rq, k, seed)
                                                                           : murmur3_h1_stream(F32.data(), q, k, seed);
            }
        }
    }
    return 0;
}
