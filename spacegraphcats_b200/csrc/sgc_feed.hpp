// sgc_feed.hpp -- host feeder for read labelling: multi-threaded BGZF block inflate, FASTA/FASTQ
// record scan into the packed (seq, off) layout the kernels take, and the mate-pairing test.
// No device work.  Included by sgc_b200.cu (shares its error plumbing).
//
// What it stands in for (reference, all pure Python):
//   utils/bgzf/bgzf.py:458-741      BgzfReader (block inflate, tell() = virtual offset)
//   search/search_utils.py:67-154   iterate_bgzf / my_fasta_iter / my_fastq_iter (records + their positions)
//   cdbg/index_reads.py:124-141     the is_paired test against the previous long-enough record
// The GPU labels 10^7 reads in tens of milliseconds; this is what keeps the whole index_reads run from
// being bound by per-record Python.
#pragma once
#include <zlib.h>

#include <atomic>
#include <thread>

namespace feed {

inline int pick_threads(int n_threads, int64_t work_items) {
    int t = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if (t > 256) t = 256;
    if ((int64_t)t > work_items) t = (int)std::max<int64_t>(work_items, 1);
    return t;
}

template <class F>
void parallel_for(int nthreads, F&& body) {  // body(thread index)
    if (nthreads <= 1) {
        body(0);
        return;
    }
    std::vector<std::thread> th;
    th.reserve((size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&body, t] { body(t); });
    for (auto& x : th) x.join();
}

inline bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

// [lo, hi) with ASCII whitespace stripped from both ends (Python's line.strip())
inline void strip(const uint8_t* d, int64_t& lo, int64_t& hi) {
    while (lo < hi && is_space(d[lo])) ++lo;
    while (hi > lo && is_space(d[hi - 1])) --hi;
}

inline int64_t line_end(const uint8_t* d, int64_t pos, int64_t n) {  // index of '\n' or n
    const void* p = memchr(d + pos, '\n', (size_t)(n - pos));
    return p ? (int64_t)((const uint8_t*)p - d) : n;
}

inline int64_t count_newlines(const uint8_t* d, int64_t lo, int64_t hi) {
    int64_t c = 0;
    while (lo < hi) {
        const void* p = memchr(d + lo, '\n', (size_t)(hi - lo));
        if (!p) break;
        ++c;
        lo = (int64_t)((const uint8_t*)p - d) + 1;
    }
    return c;
}

// Chunked view of a text buffer: chunk t owns the line starts in [b[t], b[t+1]).
struct Chunks {
    int nt;
    std::vector<int64_t> b;          // nt + 1 byte bounds
    std::vector<int64_t> first;      // first line start owned by chunk t (or b[t+1] if none)
    std::vector<int64_t> lines_before;  // number of line starts before chunk t
    int64_t total_lines;
};

inline Chunks make_chunks(const uint8_t* d, int64_t n, int nthreads) {
    Chunks c;
    c.nt = pick_threads(nthreads, std::max<int64_t>(n >> 16, 1));
    c.b.resize((size_t)c.nt + 1);
    for (int t = 0; t <= c.nt; ++t) c.b[(size_t)t] = n * t / c.nt;
    c.first.assign((size_t)c.nt, 0);
    std::vector<int64_t> cnt((size_t)c.nt, 0);
    parallel_for(c.nt, [&](int t) {
        int64_t lo = c.b[(size_t)t], hi = c.b[(size_t)t + 1];
        // a line start is position 0 or any position p < n with d[p-1] == '\n'
        int64_t k = count_newlines(d, lo > 0 ? lo - 1 : 0, hi > 0 ? hi - 1 : 0);
        if (lo == 0 && hi > 0) ++k;
        cnt[(size_t)t] = k;
        int64_t f = hi;
        if (lo == 0) f = 0;
        else {
            int64_t e = line_end(d, lo - 1, n);  // first '\n' at or after lo-1
            if (e + 1 < hi) f = e + 1;
        }
        c.first[(size_t)t] = f;
    });
    c.lines_before.assign((size_t)c.nt + 1, 0);
    for (int t = 0; t < c.nt; ++t) c.lines_before[(size_t)t + 1] = c.lines_before[(size_t)t] + cnt[(size_t)t];
    c.total_lines = c.lines_before[(size_t)c.nt];
    return c;
}

struct RecordSink {  // all nullable: counting passes leave them out
    int64_t* rec_pos;
    int64_t* name_lo;
    int64_t* name_hi;
    int64_t* seq_len;  // per record
};

// ---- FASTQ: record r is lines 4r .. 4r+3 (search_utils.py my_fastq_iter) ----------------------
// returns 0, or 1 + the index of the first malformed record
inline int64_t fastq_pass(const uint8_t* d, int64_t n, const Chunks& c, int64_t nrec, const RecordSink& s,
                          const int64_t* seq_off, uint8_t* seq_out, int64_t* total_seq) {
    std::atomic<int64_t> bad{0}, tot{0};
    parallel_for(c.nt, [&](int t) {
        int64_t pos = c.first[(size_t)t], hi = c.b[(size_t)t + 1], line = c.lines_before[(size_t)t];
        int64_t local = 0;
        // advance to a record boundary
        while (pos < hi && (line & 3)) {
            pos = line_end(d, pos, n) + 1;
            ++line;
        }
        while (pos < hi && pos < n) {
            int64_t r = line >> 2;
            if (r >= nrec) break;
            int64_t h_lo = pos, h_hi = line_end(d, pos, n);
            int64_t s_lo = std::min(h_hi + 1, n), s_hi = line_end(d, s_lo, n);
            int64_t p_lo = std::min(s_hi + 1, n), p_hi = line_end(d, p_lo, n);
            int64_t q_lo = std::min(p_hi + 1, n), q_hi = line_end(d, q_lo, n);
            int64_t a = h_lo, b = h_hi;
            strip(d, a, b);
            int64_t pa = p_lo, pb = p_hi;
            strip(d, pa, pb);
            if (d[h_lo] != '@' || pb - pa != 1 || d[pa] != '+') {
                int64_t want = r + 1, cur = bad.load();
                while ((cur == 0 || want < cur) && !bad.compare_exchange_weak(cur, want)) {}
                return;
            }
            strip(d, s_lo, s_hi);
            if (s.rec_pos) {
                s.rec_pos[r] = h_lo;
                s.name_lo[r] = a + 1;  // name = line.strip()[1:]
                s.name_hi[r] = b;
                s.seq_len[r] = s_hi - s_lo;
            }
            if (seq_out) memcpy(seq_out + seq_off[r], d + s_lo, (size_t)(s_hi - s_lo));
            local += s_hi - s_lo;
            pos = q_hi + 1;
            line += 4;
        }
        tot += local;
    });
    if (total_seq) *total_seq = tot.load();
    return bad.load();
}

// ---- FASTA: a record starts at every line beginning with '>' (my_fasta_iter) ---------------------
inline void fasta_headers_per_chunk(const uint8_t* d, int64_t n, const Chunks& c, std::vector<int64_t>& before) {
    std::vector<int64_t> cnt((size_t)c.nt, 0);
    parallel_for(c.nt, [&](int t) {
        int64_t pos = c.first[(size_t)t], hi = c.b[(size_t)t + 1], k = 0;
        while (pos < hi && pos < n) {
            if (d[pos] == '>') ++k;
            pos = line_end(d, pos, n) + 1;
        }
        cnt[(size_t)t] = k;
    });
    before.assign((size_t)c.nt + 1, 0);
    for (int t = 0; t < c.nt; ++t) before[(size_t)t + 1] = before[(size_t)t] + cnt[(size_t)t];
}

inline void fasta_pass(const uint8_t* d, int64_t n, const Chunks& c, const std::vector<int64_t>& before,
                       const RecordSink& s, const int64_t* seq_off, uint8_t* seq_out, int64_t* total_seq) {
    std::atomic<int64_t> tot{0};
    parallel_for(c.nt, [&](int t) {
        int64_t pos = c.first[(size_t)t], hi = c.b[(size_t)t + 1], r = before[(size_t)t], local = 0;
        // skip the tail of a record owned by an earlier chunk
        while (pos < hi && pos < n && d[pos] != '>') pos = line_end(d, pos, n) + 1;
        while (pos < hi && pos < n) {  // d[pos] == '>'
            int64_t h_hi = line_end(d, pos, n);
            int64_t a = pos, b = h_hi;
            strip(d, a, b);  // line.strip()
            ++a;             // [1:]
            strip(d, a, b);  // name.strip()
            int64_t len = 0;
            uint8_t* out = seq_out ? seq_out + seq_off[r] : nullptr;
            int64_t q = std::min(h_hi + 1, n);
            while (q < n && d[q] != '>') {  // sequence lines run past the chunk end if need be
                int64_t e = line_end(d, q, n);
                int64_t lo = q, up = e;
                strip(d, lo, up);
                if (out) memcpy(out + len, d + lo, (size_t)(up - lo));
                len += up - lo;
                q = std::min(e + 1, n);
                if (e >= n) break;
            }
            if (s.rec_pos) {
                s.rec_pos[r] = pos;
                s.name_lo[r] = a;
                s.name_hi[r] = b;
                s.seq_len[r] = len;
            }
            local += len;
            ++r;
            pos = q;
            if (q >= n) break;
        }
        tot += local;
    });
    if (total_seq) *total_seq = tot.load();
}

}  // namespace feed

extern "C" {

// ---------------------------------------------------------------------------- BGZF
int sgc_bgzf_index(const uint8_t* h_raw, int64_t raw_bytes, int64_t* h_cstart, int64_t* h_usize, int64_t cap,
                   int64_t* h_nblocks, int64_t* h_total_bytes) {
    if (!h_raw || raw_bytes < 0 || cap < 0) return fail(SGC_ERR_ARG, "bad argument");
    int64_t pos = 0, nb = 0, total = 0;
    while (pos < raw_bytes) {
        if (pos + 18 > raw_bytes || h_raw[pos] != 0x1f || h_raw[pos + 1] != 0x8b || h_raw[pos + 2] != 8 ||
            !(h_raw[pos + 3] & 4))
            return fail(SGC_ERR_ARG, "not a BGZF block at byte %lld", (long long)pos);
        int64_t xlen = h_raw[pos + 10] | (h_raw[pos + 11] << 8);
        if (pos + 12 + xlen > raw_bytes) return fail(SGC_ERR_ARG, "BGZF block truncated at byte %lld", (long long)pos);
        int64_t bsize = -1;
        for (int64_t i = pos + 12; i + 4 <= pos + 12 + xlen;) {
            int64_t slen = h_raw[i + 2] | (h_raw[i + 3] << 8);
            if (h_raw[i] == 66 && h_raw[i + 1] == 67 && slen == 2 && i + 6 <= pos + 12 + xlen)
                bsize = (int64_t)(h_raw[i + 4] | (h_raw[i + 5] << 8)) + 1;
            i += 4 + slen;
        }
        if (bsize < 0) return fail(SGC_ERR_ARG, "BGZF block without BC subfield at byte %lld", (long long)pos);
        if (bsize < 12 + xlen + 8 || pos + bsize > raw_bytes)
            return fail(SGC_ERR_ARG, "BGZF block truncated at byte %lld", (long long)pos);
        const uint8_t* tail = h_raw + pos + bsize - 4;
        int64_t isize = (int64_t)tail[0] | ((int64_t)tail[1] << 8) | ((int64_t)tail[2] << 16) | ((int64_t)tail[3] << 24);
        if (nb < cap) {
            if (h_cstart) h_cstart[nb] = pos;
            if (h_usize) h_usize[nb] = isize;
        }
        ++nb;
        total += isize;
        pos += bsize;
    }
    if (h_nblocks) *h_nblocks = nb;
    if (h_total_bytes) *h_total_bytes = total;
    if (cap && nb > cap) return fail(SGC_ERR_CAPACITY, "%lld BGZF blocks, room for %lld", (long long)nb, (long long)cap);
    return SGC_OK;
}

int sgc_bgzf_inflate(const uint8_t* h_raw, int64_t raw_bytes, const int64_t* h_cstart, const int64_t* h_usize,
                     int64_t nblocks, int n_threads, uint8_t* h_out, int64_t out_bytes) {
    if (!h_raw || !h_cstart || !h_usize || nblocks < 0 || (out_bytes > 0 && !h_out))
        return fail(SGC_ERR_ARG, "bad argument");
    std::vector<int64_t> ustart((size_t)nblocks + 1, 0);
    for (int64_t b = 0; b < nblocks; ++b) {
        if (h_usize[b] < 0 || h_cstart[b] < 0 || h_cstart[b] + 26 > raw_bytes)
            return fail(SGC_ERR_ARG, "bad block table entry %lld", (long long)b);
        ustart[(size_t)b + 1] = ustart[(size_t)b] + h_usize[b];
    }
    if (ustart[(size_t)nblocks] != out_bytes)
        return fail(SGC_ERR_ARG, "output buffer is %lld bytes, blocks inflate to %lld", (long long)out_bytes,
                    (long long)ustart[(size_t)nblocks]);
    int nt = feed::pick_threads(n_threads, nblocks);
    std::atomic<int64_t> next{0}, bad{-1};
    feed::parallel_for(nt, [&](int) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) {
            bad.store(0);
            return;
        }
        for (;;) {
            int64_t b0 = next.fetch_add(16);
            if (b0 >= nblocks || bad.load() >= 0) break;
            for (int64_t b = b0; b < std::min(b0 + 16, nblocks); ++b) {
                int64_t pos = h_cstart[b];
                int64_t xlen = h_raw[pos + 10] | (h_raw[pos + 11] << 8);
                int64_t end = (b + 1 < nblocks) ? h_cstart[b + 1] : raw_bytes;
                // BSIZE again (the table carries only the start): the payload ends 8 bytes before the block end
                int64_t bsize = -1;
                for (int64_t i = pos + 12; i + 6 <= pos + 12 + xlen && i + 6 <= raw_bytes;) {
                    int64_t slen = h_raw[i + 2] | (h_raw[i + 3] << 8);
                    if (h_raw[i] == 66 && h_raw[i + 1] == 67 && slen == 2) bsize = (int64_t)(h_raw[i + 4] | (h_raw[i + 5] << 8)) + 1;
                    i += 4 + slen;
                }
                if (bsize < 12 + xlen + 8 || pos + bsize > end) {
                    bad.store(b);
                    break;
                }
                const uint8_t* cdata = h_raw + pos + 12 + xlen;
                int64_t clen = bsize - 12 - xlen - 8;
                uint8_t* dst = h_out + ustart[(size_t)b];
                int64_t want = h_usize[b];
                if (want == 0) continue;  // empty block (the EOF marker)
                inflateReset(&zs);
                zs.next_in = const_cast<Bytef*>(cdata);
                zs.avail_in = (uInt)clen;
                zs.next_out = dst;
                zs.avail_out = (uInt)want;
                int rc = inflate(&zs, Z_FINISH);
                const uint8_t* tail = h_raw + pos + bsize - 8;
                uint32_t crc = (uint32_t)tail[0] | ((uint32_t)tail[1] << 8) | ((uint32_t)tail[2] << 16) | ((uint32_t)tail[3] << 24);
                if (rc != Z_STREAM_END || (int64_t)zs.total_out != want ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, (uInt)want) != crc) {
                    bad.store(b);
                    break;
                }
            }
        }
        inflateEnd(&zs);
    });
    if (bad.load() >= 0) return fail(SGC_ERR_ARG, "BGZF block %lld does not inflate (corrupt data or CRC)", (long long)bad.load());
    return SGC_OK;
}

// ---------------------------------------------------------------------------- 2-bit packing
// ASCII (seq, off) -> packed records for sgc_label_reads_packed: A=0 C=1 G=2 T=3, base b of record s in bits
// 2(b%4).. of byte 4*unit_off[s] + b/4, every record padded to a multiple of 4 bytes (16 bases).
// h_len[s] = length in bases; h_bad[s] (nullable) = 1 when the record holds a byte other than A/C/G/T (packed
// as A: the caller applies the reference's policy -- khmer raises on such reads).  A quarter of the bytes
// cross PCIe and HBM; the kernel expands them back to the ASCII words the reference hash is defined on.
int sgc_pack_reads(const uint8_t* h_seq, const int64_t* h_off, int64_t nseq, int n_threads, uint8_t* h_packed,
                   int64_t packed_cap, int32_t* h_len, uint8_t* h_bad, int64_t* h_packed_bytes) {
    if (nseq < 0 || !h_off || !h_packed_bytes || (nseq > 0 && !h_len)) return fail(SGC_ERR_ARG, "bad argument");
    std::vector<int64_t> uoff((size_t)nseq + 1, 0);
    for (int64_t s = 0; s < nseq; ++s) {
        const int64_t L = h_off[s + 1] - h_off[s];
        if (L < 0 || L > 0x7FFFFFFF) return fail(SGC_ERR_ARG, "bad offsets at record %lld", (long long)s);
        uoff[(size_t)s + 1] = uoff[(size_t)s] + (L + 15) / 16;
    }
    const int64_t total = uoff[(size_t)nseq] * 4;
    *h_packed_bytes = total;
    if (total > packed_cap) return fail(SGC_ERR_CAPACITY, "packed_cap %lld < %lld bytes", (long long)packed_cap, (long long)total);
    if (total > 0 && (!h_packed || !h_seq)) return fail(SGC_ERR_ARG, "NULL buffer");
    static const struct Lut { uint8_t v[256]; Lut() { memset(v, 4, 256); v['A'] = 0; v['C'] = 1; v['G'] = 2; v['T'] = 3; } } lut;
    const int nt = feed::pick_threads(n_threads, std::max<int64_t>(nseq >> 10, 1));
    feed::parallel_for(nt, [&](int t) {
        const int64_t s0 = nseq * t / nt, s1 = nseq * (t + 1) / nt;
        for (int64_t s = s0; s < s1; ++s) {
            const uint8_t* src = h_seq + h_off[s];
            const int64_t L = h_off[s + 1] - h_off[s];
            uint8_t* dst = h_packed + uoff[(size_t)s] * 4;
            const int64_t nbytes = (uoff[(size_t)s + 1] - uoff[(size_t)s]) * 4;
            uint8_t bad = 0;
            int64_t b = 0, o = 0;
            for (; b + 4 <= L; b += 4, ++o) {
                const uint8_t c0 = lut.v[src[b]], c1 = lut.v[src[b + 1]], c2 = lut.v[src[b + 2]], c3 = lut.v[src[b + 3]];
                bad |= (uint8_t)((c0 | c1 | c2 | c3) & 4);
                dst[o] = (uint8_t)((c0 & 3) | ((c1 & 3) << 2) | ((c2 & 3) << 4) | ((c3 & 3) << 6));
            }
            if (b < L) {
                uint8_t x = 0;
                for (int j = 0; b + j < L; ++j) {
                    const uint8_t c = lut.v[src[b + j]];
                    bad |= (uint8_t)(c & 4);
                    x |= (uint8_t)((c & 3) << (2 * j));
                }
                dst[o++] = x;
            }
            if (o < nbytes) memset(dst + o, 0, (size_t)(nbytes - o));
            h_len[s] = (int32_t)L;
            if (h_bad) h_bad[s] = bad ? 1 : 0;
        }
    });
    return SGC_OK;
}

// ---------------------------------------------------------------------------- label offsets on the wire
// The CSR offsets of a labelled batch travel device -> host as one byte per read (its number of labels; see
// sgc_offsets_to_counts8) and are rebuilt here: h_off[0] = 0, h_off[s+1] = h_off[s] + h_counts[s].  Two passes over
// thread-sized ranges (sums, then the offsets).
int sgc_counts8_to_offsets(const uint8_t* h_counts, int64_t n, int n_threads, int64_t* h_off) {
    if (n < 0 || !h_off || (n > 0 && !h_counts)) return fail(SGC_ERR_ARG, "bad argument");
    const int nt = feed::pick_threads(n_threads, std::max<int64_t>(n >> 18, 1));
    std::vector<int64_t> base((size_t)nt + 1, 0);
    feed::parallel_for(nt, [&](int t) {
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        int64_t s = 0;
        for (int64_t i = a; i < b; ++i) s += h_counts[i];
        base[(size_t)t + 1] = s;
    });
    for (int t = 0; t < nt; ++t) base[(size_t)t + 1] += base[(size_t)t];
    feed::parallel_for(nt, [&](int t) {
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        int64_t s = base[(size_t)t];
        for (int64_t i = a; i < b; ++i) {
            h_off[i] = s;
            s += h_counts[i];
        }
    });
    h_off[n] = base[(size_t)nt];
    return SGC_OK;
}

// ---------------------------------------------------------------------------- records
// h_info[0..2] = number of records, total sequence bytes, format (0 FASTA, 1 FASTQ)
int sgc_reads_count(const uint8_t* h_data, int64_t n, int n_threads, int64_t* h_info) {
    if (n < 0 || (n > 0 && !h_data) || !h_info) return fail(SGC_ERR_ARG, "bad argument");
    h_info[0] = h_info[1] = 0;
    h_info[2] = 0;
    if (n == 0) return SGC_OK;
    feed::RecordSink none{nullptr, nullptr, nullptr, nullptr};
    feed::Chunks c = feed::make_chunks(h_data, n, n_threads);
    if (h_data[0] == '@') {
        int64_t nrec = c.total_lines / 4, tot = 0;
        int64_t bad = feed::fastq_pass(h_data, n, c, nrec, none, nullptr, nullptr, &tot);
        if (bad) return fail(SGC_ERR_ARG, "FASTQ records must be exactly 4 lines (record %lld)", (long long)(bad - 1));
        h_info[0] = nrec;
        h_info[1] = tot;
        h_info[2] = 1;
        return SGC_OK;
    }
    if (h_data[0] == '>') {
        std::vector<int64_t> before;
        feed::fasta_headers_per_chunk(h_data, n, c, before);
        int64_t tot = 0;
        feed::fasta_pass(h_data, n, c, before, none, nullptr, nullptr, &tot);
        h_info[0] = before[(size_t)c.nt];
        h_info[1] = tot;
        return SGC_OK;
    }
    return fail(SGC_ERR_ARG, "unknown start chr %c", (char)h_data[0]);
}

// Fills rec_pos / name_lo / name_hi (nrec each), seq_off (nrec + 1) and the packed sequence bytes.
// End of the last COMPLETE record of a buffer that may stop in the middle of one (streaming reads in bounded
// memory: the tail from *h_cut on is carried over to the next chunk).  FASTQ: records are exactly 4 lines, so
// the cut is the start of line 4*floor(L/4) (L = complete lines); FASTA: the start of the last header line
// (the record it opens ends only at the next header or at the end of the file).  *h_cut = 0: no complete record.
int sgc_reads_cut(const uint8_t* h_data, int64_t n, int n_threads, int64_t* h_cut) {
    if (n < 0 || (n > 0 && !h_data) || !h_cut) return fail(SGC_ERR_ARG, "bad argument");
    *h_cut = 0;
    if (n == 0) return SGC_OK;
    if (h_data[0] == '@') {
        const int nt = feed::pick_threads(n_threads, std::max<int64_t>(n >> 20, 1));
        std::vector<int64_t> cnt((size_t)nt, 0);
        feed::parallel_for(nt, [&](int t) { cnt[(size_t)t] = feed::count_newlines(h_data, n * t / nt, n * (t + 1) / nt); });
        int64_t lines = 0;
        for (int64_t c : cnt) lines += c;
        int64_t drop = lines % 4;  // complete lines of the unfinished record
        int64_t pos = n;           // walk back over the partial line and `drop` complete lines
        if (h_data[n - 1] != '\n') {
            while (pos > 0 && h_data[pos - 1] != '\n') --pos;
        }
        for (int64_t d = 0; d < drop; ++d) {
            --pos;  // the newline that ends this line
            while (pos > 0 && h_data[pos - 1] != '\n') --pos;
        }
        *h_cut = pos;
        return SGC_OK;
    }
    if (h_data[0] == '>') {
        int64_t pos = n;
        for (;;) {  // start of the line that holds pos-1
            while (pos > 0 && h_data[pos - 1] != '\n') --pos;
            if (pos == 0 || h_data[pos] == '>') break;   // pos < n here: a line start
            --pos;
        }
        *h_cut = (pos < n && h_data[pos] == '>') ? pos : 0;
        return SGC_OK;
    }
    return fail(SGC_ERR_ARG, "unknown start chr %c", (char)h_data[0]);
}

int sgc_reads_scan(const uint8_t* h_data, int64_t n, int n_threads, int64_t nrec, int64_t seq_bytes,
                   int64_t* h_rec_pos, int64_t* h_name_lo, int64_t* h_name_hi, int64_t* h_seq_off, uint8_t* h_seq) {
    if (n < 0 || (n > 0 && !h_data) || nrec < 0 || seq_bytes < 0 || !h_seq_off) return fail(SGC_ERR_ARG, "bad argument");
    h_seq_off[0] = 0;
    if (nrec == 0) return SGC_OK;
    if (!h_rec_pos || !h_name_lo || !h_name_hi || (seq_bytes > 0 && !h_seq)) return fail(SGC_ERR_ARG, "NULL output array");
    feed::Chunks c = feed::make_chunks(h_data, n, n_threads);
    feed::RecordSink sink{h_rec_pos, h_name_lo, h_name_hi, h_seq_off + 1};
    bool fastq = h_data[0] == '@';
    std::vector<int64_t> before;
    int64_t tot = 0;
    if (fastq) {
        if (c.total_lines / 4 != nrec) return fail(SGC_ERR_ARG, "nrec does not match the buffer");
        int64_t bad = feed::fastq_pass(h_data, n, c, nrec, sink, nullptr, nullptr, &tot);
        if (bad) return fail(SGC_ERR_ARG, "FASTQ records must be exactly 4 lines (record %lld)", (long long)(bad - 1));
    } else if (h_data[0] == '>') {
        feed::fasta_headers_per_chunk(h_data, n, c, before);
        if (before[(size_t)c.nt] != nrec) return fail(SGC_ERR_ARG, "nrec does not match the buffer");
        feed::fasta_pass(h_data, n, c, before, sink, nullptr, nullptr, &tot);
    } else {
        return fail(SGC_ERR_ARG, "unknown start chr %c", (char)h_data[0]);
    }
    if (tot != seq_bytes) return fail(SGC_ERR_ARG, "seq_bytes does not match the buffer");
    for (int64_t r = 0; r < nrec; ++r) h_seq_off[r + 1] += h_seq_off[r];  // lengths -> offsets
    feed::RecordSink none{nullptr, nullptr, nullptr, nullptr};
    if (fastq) feed::fastq_pass(h_data, n, c, nrec, none, h_seq_off, h_seq, nullptr);
    else feed::fasta_pass(h_data, n, c, before, none, h_seq_off, h_seq, nullptr);
    return SGC_OK;
}

// index_reads.py:124-141 for every record: h_last[i] = the latest earlier record with h_elig set
// (-1: none; always -1 with ignore_paired), h_paired[i] = 1 when record i is that record's mate.
int sgc_reads_pair_flags(const uint8_t* h_data, const int64_t* h_name_lo, const int64_t* h_name_hi,
                         const uint8_t* h_elig, int64_t nrec, int ignore_paired, uint8_t* h_paired, int64_t* h_last) {
    if (nrec < 0 || (nrec > 0 && (!h_data || !h_name_lo || !h_name_hi || !h_elig || !h_paired || !h_last)))
        return fail(SGC_ERR_ARG, "bad argument");
    int64_t last = -1;
    for (int64_t i = 0; i < nrec; ++i) {
        h_last[i] = ignore_paired ? -1 : last;
        uint8_t paired = 0;
        if (!ignore_paired && last >= 0) {
            const uint8_t* a = h_data + h_name_lo[last];
            const uint8_t* b = h_data + h_name_lo[i];
            int64_t la = h_name_hi[last] - h_name_lo[last], lb = h_name_hi[i] - h_name_lo[i];
            bool a1 = la >= 2 && a[la - 2] == '/' && a[la - 1] == '1';
            bool b2 = lb >= 2 && b[lb - 2] == '/' && b[lb - 1] == '2';
            if (a1 && b2) paired = (la == lb && la > 2 && memcmp(a, b, (size_t)(la - 2)) == 0);
            else paired = (la == lb && la > 0 && memcmp(a, b, (size_t)la) == 0);
        }
        h_paired[i] = paired;
        if (h_elig[i]) last = i;
    }
    return SGC_OK;
}

// The row-producing half of the index_reads loop (index_reads.py:124-167) given every record's
// distinct ascending unitig ids (CSR from sgc_label_reads): the id set carries over from a record
// to its mate, is reset by any record that is not paired (long enough or not), and a long-enough
// record emits the set under its own offset and, when paired, first under its mate's offset.
// h_rows_* may be NULL (cap 0) to count only; *h_n_rows is always the number of rows.
int sgc_reads_rows(const uint8_t* h_elig, const uint8_t* h_paired, const int64_t* h_last, const int64_t* h_offsets,
                   const int64_t* h_ids_off, const uint32_t* h_ids, int64_t nrec, int64_t* h_rows_offset,
                   int64_t* h_rows_id, int64_t cap, int64_t* h_n_rows) {
    if (nrec < 0 || !h_n_rows || (nrec > 0 && (!h_elig || !h_paired || !h_last || !h_offsets || !h_ids_off)))
        return fail(SGC_ERR_ARG, "bad argument");
    if (cap > 0 && (!h_rows_offset || !h_rows_id)) return fail(SGC_ERR_ARG, "NULL output array");
    std::vector<uint32_t> cur, merged;
    int64_t n_rows = 0;
    for (int64_t i = 0; i < nrec; ++i) {
        if (!h_paired[i]) cur.clear();
        if (!h_elig[i]) continue;
        const uint32_t* a = h_ids + h_ids_off[i];
        int64_t na = h_ids_off[i + 1] - h_ids_off[i];
        if (na < 0) return fail(SGC_ERR_ARG, "ids_off decreases at record %lld", (long long)i);
        if (cur.empty()) cur.assign(a, a + na);
        else if (na) {
            merged.resize(cur.size() + (size_t)na);
            merged.resize((size_t)(std::set_union(cur.begin(), cur.end(), a, a + na, merged.begin()) - merged.begin()));
            cur.swap(merged);
        }
        bool mate = h_paired[i] && h_last[i] >= 0;
        int64_t emit = (int64_t)cur.size() * (mate ? 2 : 1);
        if (n_rows + emit <= cap) {
            int64_t w = n_rows;
            if (mate)
                for (uint32_t id : cur) { h_rows_offset[w] = h_offsets[h_last[i]]; h_rows_id[w++] = id; }
            for (uint32_t id : cur) { h_rows_offset[w] = h_offsets[i]; h_rows_id[w++] = id; }
        }
        n_rows += emit;
    }
    *h_n_rows = n_rows;
    if (cap > 0 && n_rows > cap) return fail(SGC_ERR_CAPACITY, "%lld rows, room for %lld", (long long)n_rows, (long long)cap);
    return SGC_OK;
}

}  // extern "C"
