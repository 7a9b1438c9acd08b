// sgc_layout.cuh -- tile geometry constants, the table slot layout, device scalars and the tile kernel argument block.
// Part of libsgc_b200.so (device code; included by sgc_b200.cu only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sgc_b200.h"

namespace sgc {
namespace dev {

// ----------------------------------------------------------------------------------
// tile geometry
// ----------------------------------------------------------------------------------
// Resident CTAs per SM the probing kernels are compiled for.  These kernels wait on dependent HBM/L2
// loads, and ptxas will not overlap a load with arithmetic across a loop trip or a forward branch (it
// drains the load's scoreboard there), so the latency is hidden by resident warps instead.  The
// label / lookup kernels fit in 64 registers without spilling once they carry no cross-trip state:
// 4 CTAs x 256 threads = 32 warps per SM (side-array kernel: 16 warps 53 G k-mers/s, 24: 62-67, 27: 70, 32: 71.5).
#ifndef SGC_PROBE_MIN_CTAS
#define SGC_PROBE_MIN_CTAS 4
#endif
#ifndef SGC_NT
#define SGC_NT 256
#endif
constexpr int NT = SGC_NT;      // threads per CTA
constexpr int WPB = NT / 32;    // warps per CTA; every warp runs its own rounds (no block barriers)
constexpr int SPW = 4;          // segments staged per warp round (8 lanes stage one segment)
constexpr int SEG_K = 256;      // k-mers per segment
constexpr int PAIR_BUF = 128;   // label keys staged per warp before one global reservation
constexpr unsigned long long SEQ_MASK = (1ULL << 40) - 1ULL;

constexpr uint32_t VAL_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t VAL_MULTI = 0xFFFFFFFEu;

enum Mode { MODE_HASH = 0, MODE_INSERT = 1, MODE_LOOKUP = 2, MODE_LABEL = 3, MODE_PARTITION = 4, MODE_GATHER = 5, MODE_MINHASH = 6 };

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
// aux[29:0] = position of the k-mer's first base in the table's unitig sequence store (AUX_NONE if none);
// aux[30] = orientation of the k-mer as stored: set when Murmur3(kmer) < Murmur3(revcomp(kmer)) for
// the k-mer as it is written in its unitig.  A read k-mer with the same hash and the same bit lies on
// the unitig's strand (its successors are at g+1, g+2, ...), with the opposite bit on the other strand
// (g-1, g-2, ...), so the read is compared with the store in one direction only;
// aux[31] of slot 0 = the bucket's overflow bit.
struct __align__(16) Entry {
    unsigned long long key;
    unsigned int val;  // VAL_EMPTY = free slot, VAL_MULTI = dropped hash, else unitig id
    unsigned int aux;  // sequence-store position (30 bits) | orientation bit | overflow bit (slot 0)
};
constexpr uint32_t AUX_NONE = 0x3FFFFFFFu;   // as a position: "no side record"
constexpr uint32_t AUX_MASK = 0x3FFFFFFFu;
constexpr uint32_t ORI_BIT = 0x40000000u;
constexpr uint32_t OVF_BIT = 0x80000000u;
struct __align__(32) Bucket {
    Entry e[2];
};

// One segment (<= SEG_K consecutive k-mers of one sequence), precomputed by the plan so that the
// tile kernel can fetch it with one independent asynchronous copy (no dependent loads).
struct __align__(16) SegDesc {
    long long start;   // byte offset of the segment's first base in the sequence buffer
    long long kout;    // global index of its first k-mer
    unsigned int seq;  // sequence index
    int nk;            // k-mers in this segment (0 = nothing to do)
    unsigned int pad[2];
};
static_assert(sizeof(SegDesc) == 32, "SegDesc is copied as two 16-byte cp.async");

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
    unsigned int csr_big;  // label tail: a sequence has more candidate ids than the per-thread path sorts
    unsigned int done_ctr;   // flag protocol of the partitioned path: warps of the request kernel that have finished
    unsigned int done_ctr2;  // ... CTAs of the owner-side probe kernel that have finished
    unsigned int pad_;
    int err;  // bit 0: table full, bit 1: id out of range
    int bad_offsets;  // set by the plan when an offset pair is negative, decreasing or past seq_bytes
};

struct TileArgs {
    const uint8_t* seq;
    const int64_t* off;
    const SegDesc* segdesc;
    const int64_t* nsegs_p;
    const int64_t* kmer_off;
    int32_t* status;
    Scalars* scal;
    int k;
    // MODE_HASH / MODE_MINHASH
    uint64_t* hash_out;
    uint32_t* aux_out;              // MODE_HASH, nullable: per k-mer (upos_base + byte offset of its first base) | ORI_BIT
    unsigned long long* seq_min;    // MODE_MINHASH: per-sequence minimum (atomicMin), nullable
    unsigned long long seed;        // MODE_MINHASH: Murmur3 seed (sourmash: 42)
    // table (MODE_INSERT / MODE_LOOKUP / MODE_LABEL)
    Bucket* tbl;
    uint32_t nbuckets;
    uint32_t nlines;
    int home_shift;
    const uint32_t* seq_ids;
    uint32_t id_base;
    // unitig sequence store (sgc_labelseq.cuh): MODE_INSERT records every k-mer's store position in its slot
    uint32_t* brk;                  // break bits of the store (NULL: table without a store)
    unsigned long long upos_base;   // MODE_INSERT: store position of byte 0 of this batch's sequence buffer
    // MODE_LOOKUP
    uint32_t* kmer_ids;
    uint32_t* count_by_id;
    // MODE_LABEL
    unsigned long long* pairs;
    unsigned long long pairs_cap;
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
    // flag protocol (sgc_label_reads_partitioned): no host round trip between request, probe and gather kernels.
    // The last warp of the request kernel publishes this rank's per-owner counts and the step's epoch into every
    // owner's mailbox (peer memory); the gather kernel waits until every owner has stamped its replies.
    unsigned long long* peer_mail_count[32];  // owner r's count mailbox (my slot: [my_rank])
    unsigned int* peer_mail_stamp[32];        // owner r's request stamps
    const unsigned int* reply_stamp;          // LOCAL: reply_stamp[r] == epoch once owner r has stored all my replies
    int my_rank;
    unsigned int epoch;                       // 0 = flag protocol off
};


}  // namespace dev
}  // namespace sgc
