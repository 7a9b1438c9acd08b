// sgc_mphf.cuh -- BBHash (boomphf) minimal perfect hash built and queried on the device.
//
// The reference persists its k-mer index through bbhash_table.BBHashTable (reference
// spacegraphcats/cdbg/hashing.py:8,30-37; index_cdbg_by_kmer.py:63-64,89,105): `contigs.mphf` is a
// boomphf::mphf<uint64_t> dump (gamma 1.0, 25 levels) and `contigs.indices` holds the hash and
// cDBG-id arrays in MPHF order.  These kernels build the same level bitsets from the table's keys
// so that the index written by this library loads in an unmodified CPU install.  The algorithm is
// restated in oracle/bbhash.py, which is pinned byte-for-byte on the reference's dory fixture.
//
// All kernels are HBM/L2-bound bit work: one 32-bit atomicOr (mark) or one 8-byte load (filter)
// per key and level, grid-stride over the surviving keys.
#pragma once
#include <cstdint>

namespace sgc {
namespace dev {

constexpr int MPHF_MAX_LEVELS = 64;
constexpr uint64_t MPHF_NOT_FOUND = ~0ull;

struct MphfLevels {  // by-value kernel argument
    int nb_levels;
    uint64_t size[MPHF_MAX_LEVELS];  // bits in the level (multiple of 64)
    uint64_t woff[MPHF_MAX_LEVELS];  // first 64-bit word of the level in the concatenated array (multiple of 8)
};

__host__ __device__ __forceinline__ uint64_t bb_hash64(uint64_t key, uint64_t seed) {
    uint64_t h = seed;
    h ^= (h << 7) ^ (key * (h >> 3)) ^ (~((h << 11) + (key ^ (h >> 5))));
    h = (~h) + (h << 21);
    h ^= h >> 24;
    h = (h + (h << 3)) + (h << 8);
    h ^= h >> 14;
    h = (h + (h << 2)) + (h << 4);
    h ^= h >> 28;
    h += h << 31;
    return h;
}

// h_0, h_1, then the xorshift128+ stream seeded with them; call next() once per level, ascending.
struct BbLevelHash {
    uint64_t s0, s1;
    int level;
    __device__ __forceinline__ explicit BbLevelHash(uint64_t key)
        : s0(bb_hash64(key, 0xAAAAAAAA55555555ull)), s1(bb_hash64(key, 0x33333333CCCCCCCCull)), level(0) {}
    __device__ __forceinline__ uint64_t next() {
        int lv = level++;
        if (lv == 0) return s0;
        if (lv == 1) return s1;
        uint64_t a = s0;
        const uint64_t b = s1;
        s0 = b;
        a ^= a << 23;
        s1 = a ^ b ^ (a >> 17) ^ (b >> 26);
        return s1 + b;
    }
    __device__ __forceinline__ uint64_t at(int lv) {  // hash of level lv (from a fresh state)
        uint64_t h = 0;
        for (int i = 0; i <= lv; ++i) h = next();
        return h;
    }
};

__device__ __forceinline__ uint64_t bb_pos(uint64_t h, uint64_t size) { return __umul64hi(h, size); }

// first hit sets the level bit, a later hit on the same bit records the collision
__global__ void __launch_bounds__(256) mphf_mark_kernel(const uint64_t* __restrict__ keys, long long n, int level,
                                                        uint64_t size, uint32_t* __restrict__ bits32,
                                                        uint32_t* __restrict__ coll32) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        BbLevelHash lh(keys[i]);
        uint64_t pos = bb_pos(lh.at(level), size);
        uint32_t m = 1u << (pos & 31);
        uint32_t old = atomicOr(&bits32[pos >> 5], m);
        if (old & m) atomicOr(&coll32[pos >> 5], m);
    }
}

__global__ void __launch_bounds__(256) mphf_settle_kernel(uint64_t* __restrict__ bits, uint64_t* __restrict__ coll,
                                                          long long nwords) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords;
         i += (long long)gridDim.x * blockDim.x) {
        uint64_t c = coll[i];
        if (c) {
            bits[i] &= ~c;
            coll[i] = 0;
        }
    }
}

// keys whose bit did not survive go on to the next level (order is irrelevant to the result)
__global__ void __launch_bounds__(256) mphf_filter_kernel(const uint64_t* __restrict__ keys, long long n, int level,
                                                          uint64_t size, const uint64_t* __restrict__ bits,
                                                          uint64_t* __restrict__ out, unsigned long long* counter) {
    long long stride = (long long)gridDim.x * blockDim.x;
    long long n_up = (n + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_up; i += stride) {
        bool keep = false;
        uint64_t key = 0;
        if (i < n) {
            key = keys[i];
            BbLevelHash lh(key);
            uint64_t pos = bb_pos(lh.at(level), size);
            keep = !((bits[pos >> 6] >> (pos & 63)) & 1);
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            int lane = threadIdx.x & 31;
            int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (keep) out[base + __popc(m & ((1u << lane) - 1))] = key;
        }
    }
}

// popcount of every 512-bit block; an exclusive scan over the output gives boomphf's rank samples
__global__ void __launch_bounds__(256) mphf_block_popc_kernel(const uint64_t* __restrict__ bits, long long nblocks,
                                                              uint64_t* __restrict__ out) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks;
         b += (long long)gridDim.x * blockDim.x) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(bits + b * 8);
        int s = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ulonglong2 v = p[j];
            s += __popcll(v.x) + __popcll(v.y);
        }
        out[b] = (uint64_t)s;
    }
}

struct MphfView {
    MphfLevels lv;
    const uint64_t* bits;
    const uint64_t* ranks;       // one per 512-bit block of the concatenated array
    const uint64_t* final_keys;  // ascending
    const uint64_t* final_vals;  // index - lastbitsetrank
    long long n_final;
    uint64_t lastbitsetrank;
};

__device__ __forceinline__ uint64_t mphf_index(const MphfView& v, uint64_t key) {
    BbLevelHash lh(key);
    for (int l = 0; l < v.lv.nb_levels - 1; ++l) {
        uint64_t pos = bb_pos(lh.next(), v.lv.size[l]);
        const uint64_t* w = v.bits + v.lv.woff[l];
        uint64_t wi = pos >> 6;
        uint64_t word = w[wi];
        if ((word >> (pos & 63)) & 1) {
            uint64_t blk = wi >> 3;
            uint64_t r = v.ranks[(v.lv.woff[l] >> 3) + blk];
            for (uint64_t j = blk << 3; j < wi; ++j) r += __popcll(w[j]);
            r += __popcll(word & ((1ull << (pos & 63)) - 1));
            return r;
        }
    }
    long long lo = 0, hi = v.n_final;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        uint64_t k = v.final_keys[mid];
        if (k == key) return v.final_vals[mid] + v.lastbitsetrank;
        if (k < key) lo = mid + 1; else hi = mid;
    }
    return MPHF_NOT_FOUND;
}

__global__ void __launch_bounds__(256) mphf_lookup_kernel(MphfView v, const uint64_t* __restrict__ keys, long long n,
                                                          uint64_t* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = mphf_index(v, keys[i]);
}

// scatter {key, value} to MPHF order: the two arrays of the reference's contigs.indices
__global__ void __launch_bounds__(256) mphf_order_kernel(MphfView v, const uint64_t* __restrict__ keys,
                                                         const uint32_t* __restrict__ vals, long long n, uint64_t nelem,
                                                         uint64_t* __restrict__ out_hash, uint32_t* __restrict__ out_val,
                                                         unsigned long long* n_bad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint64_t key = keys[i];
        uint64_t idx = mphf_index(v, key);
        if (idx >= nelem) {
            atomicAdd(n_bad, 1ull);
            continue;
        }
        out_hash[idx] = key;
        out_val[idx] = vals[i];
    }
}

__global__ void __launch_bounds__(256) iota_u64_kernel(uint64_t* out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (uint64_t)i;
}

}  // namespace dev
}  // namespace sgc
