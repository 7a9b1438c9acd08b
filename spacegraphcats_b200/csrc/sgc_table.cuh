// sgc_table.cuh -- device primitives of the GPU-resident k-mer -> unitig table (probe sequence, 128-bit CAS insert with build_mphf semantics, lookup).
// Part of libsgc_b200.so (device code; included by sgc_b200.cu only).
#pragma once
#include "sgc_layout.cuh"

namespace sgc {
namespace dev {

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
    if (e0.key == h && e0.val != VAL_EMPTY) { out = resolve_val(e0.val); aux = e0.aux & ~OVF_BIT; return true; }
    if (e1.key == h && e1.val != VAL_EMPTY) { out = resolve_val(e1.val); aux = e1.aux & ~OVF_BIT; return true; }
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

// break bit of store position p (sgc_labelseq.cuh): set = "a k-mer that sequence-matches position p must not
// inherit the id found for position p-1 (resp. p+1 when walking down)"
constexpr int BRK_PAD = 32;  // bit index of position 0 in the break-bit array
__device__ __forceinline__ void brk_mark(uint32_t* brk, uint32_t pos) {
    const uint32_t b0 = pos + (uint32_t)BRK_PAD, b1 = b0 + 1u;
    atomicOr(brk + (b0 >> 5), 1u << (b0 & 31u));
    atomicOr(brk + (b1 >> 5), 1u << (b1 & 31u));
}

// build_mphf insertion semantics (cdbg/index_cdbg_by_kmer.py:40-57, 88-89):
//   new hash -> stored with id;  same hash, same id -> no-op;
//   same hash, different id -> the hash is dropped for good (VAL_MULTI).
// The outcome is independent of the order in which threads arrive: slots only ever go from
// empty to occupied, every thread walks the same probe sequence, and a lost CAS re-examines
// what the winner wrote.
// brk (nullable) = break bits of the table's unitig sequence store: a k-mer whose hash is already in the
// table under another position fences its own position off (it is not the one the slot points to), and a
// hash that becomes a multicount fences the slot's position off too -- so whatever lies between two break
// bits is a run of k-mers whose slots point at themselves and hold their unitig's id.
// Where the build records its break bits: straight into the table's own store (brk), or -- building one owner's
// part of a hash-range partitioned table, whose store is replicated on every GPU -- as a list of
// (shard << 32 | position) marks that every rank applies afterwards.  The shard of a position is the rank whose
// unitigs were stored there; unitig ids are dealt to the ranks in ascending ranges, so it follows from the id.
struct BrkSink {
    uint32_t* brk = nullptr;
    unsigned long long* list = nullptr;
    unsigned long long* list_n = nullptr;
    unsigned long long list_cap = 0;
    uint32_t bounds[33] = {0};  // shard s holds the ids in [bounds[s], bounds[s+1])
    int n_shards = 1;
    __device__ __forceinline__ bool on() const { return brk != nullptr || list != nullptr; }
    __device__ void mark(uint32_t pos, uint32_t id) const {
        if (brk) brk_mark(brk, pos);
        else if (list && id < VAL_MULTI) {
            unsigned shard = 0;
            for (int s = 1; s < n_shards; s++) shard += id >= bounds[s];
            const unsigned long long at = atomicAdd(list_n, 1ULL);
            if (at < list_cap) list[at] = ((unsigned long long)shard << 32) | pos;
        }
    }
};

__device__ void table_insert(Bucket* tbl, uint32_t nb, uint32_t nlines, int hs, uint64_t h, uint32_t id, uint32_t aux,
                             Scalars* scal, const BrkSink& sink) {
    const uint32_t hb = home_bucket(h, nb, hs);
    const Entry empty{0ULL, VAL_EMPTY, 0u};
    const Entry mine{h, id, aux & (AUX_MASK | ORI_BIT)};
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
                const uint32_t mypos = aux & AUX_MASK, slotpos = e.aux & AUX_MASK;
                const bool marks = sink.on();
                if (marks && mypos != AUX_NONE && mypos != slotpos) sink.mark(mypos, id);
                if (e.val != id) {
                    if (e.val != VAL_MULTI) atomicExch(&bp->e[s].val, VAL_MULTI);
                    // (a slot that is a multicount already had its position fenced off by whoever made it one)
                    if (marks && slotpos != AUX_NONE) sink.mark(slotpos, e.val);
                }
                return;
            }
        }
    }
    atomicOr(&scal->err, 1);
}


}  // namespace dev
}  // namespace sgc
