// Host code: <trait>.score.gz / <trait>.full_score.gz writer (bin/scdrs:276-288 of the reference,
// format docs/file_format.rst:81-108): a gzip-compressed TSV with the cell index in the first column
// and float32 columns printed exactly as pandas' DataFrame.to_csv prints them (numpy's shortest
// round-trip float32 text: positional for 1e-4 <= |x| < 1e6 with a trailing ".0" for integers,
// scientific with a two-digit exponent otherwise, empty field for NaN).
//
// With the default --flag-return-ctrl-norm-score True a trait is n_cell x 1006 numbers; pandas
// formats and deflates that on one thread (minutes per trait at 1e5 cells), which would dominate
// the run once scoring takes milliseconds.  Here row blocks are formatted and deflated in parallel
// as independent gzip members and written in order; any gzip reader concatenates the members.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "common.cuh"
#include "score_format.cuh"

namespace {

inline char* fmt_f32(char* p, float v) {
    if (std::isnan(v)) return p;                       // na_rep = ''
    if (std::isinf(v)) {
        if (v < 0) *p++ = '-';
        memcpy(p, "inf", 3);
        return p + 3;
    }
    const double a = std::fabs((double)v);             // no float lies between the double and the exact 1e-4 / 1e6
    if (a == 0.0 || (a >= 1.e-4 && a < 1.e6)) {
        auto r = std::to_chars(p, p + 64, v, std::chars_format::fixed);
        const bool has_dot = memchr(p, '.', (size_t)(r.ptr - p)) != nullptr;
        p = r.ptr;
        if (!has_dot) { *p++ = '.'; *p++ = '0'; }
        return p;
    }
    auto r = std::to_chars(p, p + 64, v, std::chars_format::scientific);
    return r.ptr;
}

struct Block {
    std::string gz;
    std::atomic<int> ready{0};
};

bool zlib_member(const char* text, size_t n_text, int level, std::string& out) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, n_text) + 64);
    zs.next_in = (Bytef*)text;
    zs.avail_in = (uInt)n_text;
    zs.next_out = (Bytef*)out.data();
    zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return rc == Z_STREAM_END;
}

// ---- level 1: one literal-only dynamic-Huffman deflate block per gzip member --------------------------------
// Score files are digits, '.', '-', 'e', tabs and cell names: LZ77 matches buy nothing (zlib level 1 reaches 0.48 of
// the text, the order-0 entropy is 0.46) and cost almost all of zlib's time.  This writes RFC 1951 by hand: byte
// histogram -> length-limited Huffman code -> block header -> one table look-up and a shift per byte.  Any gzip
// reader accepts the members (BTYPE = 2, a single unused distance code).
struct BitSink {
    unsigned char* p;
    uint64_t buf = 0;
    int n = 0;
    inline void put(uint32_t code, int len) {          // len <= 32 at a time; the buffer is drained below 32 bits
        buf |= (uint64_t)code << n;
        n += len;
        if (n >= 32) {
            memcpy(p, &buf, 4);                         // little endian
            p += 4;
            buf >>= 32;
            n -= 32;
        }
    }
    inline unsigned char* finish() {
        while (n > 0) {
            *p++ = (unsigned char)buf;
            buf >>= 8;
            n -= 8;
        }
        return p;
    }
};

// code lengths (<= max_len) of a Huffman code for freq[0 .. n): plain Huffman; when the tree is too deep the
// frequencies are flattened (halved, floor 1) and the tree rebuilt -- a few iterations at most, rarely any
void huff_lengths(const uint64_t* freq, int n, int max_len, unsigned char* len) {
    std::vector<uint64_t> f(freq, freq + n);
    for (;;) {
        struct Node { uint64_t w; int l, r; };
        std::vector<Node> nodes;
        std::vector<int> live;
        for (int i = 0; i < n; ++i) {
            len[i] = 0;
            if (f[i] > 0) { nodes.push_back({f[i], -1 - i, -1}); live.push_back((int)nodes.size() - 1); }
        }
        if (live.empty()) return;
        if (live.size() == 1) { len[-1 - nodes[(size_t)live[0]].l] = 1; return; }
        while (live.size() > 1) {                    // the alphabet is tiny (<= 257): selection by scanning is fine
            int a = 0, b = 1;
            if (nodes[(size_t)live[(size_t)b]].w < nodes[(size_t)live[(size_t)a]].w) std::swap(a, b);
            for (int k = 2; k < (int)live.size(); ++k) {
                const uint64_t w = nodes[(size_t)live[(size_t)k]].w;
                if (w < nodes[(size_t)live[(size_t)a]].w) { b = a; a = k; }
                else if (w < nodes[(size_t)live[(size_t)b]].w) b = k;
            }
            const int ia = live[(size_t)a], ib = live[(size_t)b];
            nodes.push_back({nodes[(size_t)ia].w + nodes[(size_t)ib].w, ia, ib});
            if (a > b) std::swap(a, b);
            live.erase(live.begin() + b);
            live.erase(live.begin() + a);
            live.push_back((int)nodes.size() - 1);
        }
        int deepest = 0;
        std::vector<std::pair<int, int>> stack{{live[0], 0}};
        while (!stack.empty()) {
            const auto [id, d] = stack.back();
            stack.pop_back();
            const Node& nd = nodes[(size_t)id];
            if (nd.r < 0) {                           // leaf: l = -1 - symbol
                len[-1 - nd.l] = (unsigned char)(d > 255 ? 255 : d);
                deepest = d > deepest ? d : deepest;
            } else {
                stack.push_back({nd.l, d + 1});
                stack.push_back({nd.r, d + 1});
            }
        }
        if (deepest <= max_len) return;
        for (auto& x : f)
            if (x > 0) x = (x >> 1) > 0 ? (x >> 1) : 1;
    }
}

// canonical codes of RFC 1951 section 3.2.2, bit-reversed (deflate packs Huffman codes most significant bit first)
void huff_codes(const unsigned char* len, int n, uint32_t* code) {
    int bl_count[16] = {0};
    for (int i = 0; i < n; ++i) bl_count[len[i]]++;
    bl_count[0] = 0;
    uint32_t next[16] = {0}, c = 0;
    for (int b = 1; b < 16; ++b) {
        c = (c + (uint32_t)bl_count[b - 1]) << 1;
        next[b] = c;
    }
    for (int i = 0; i < n; ++i) {
        code[i] = 0;
        if (len[i] == 0) continue;
        uint32_t v = next[len[i]]++, r = 0;
        for (int b = 0; b < len[i]; ++b) { r = (r << 1) | (v & 1u); v >>= 1; }
        code[i] = r;
    }
}

bool huffman_member(const char* text, size_t n_text, std::string& out) {
    const unsigned char* in = reinterpret_cast<const unsigned char*>(text);
    // byte histogram in four lanes (consecutive equal bytes would otherwise serialise on one counter)
    uint64_t freq[257] = {0};
    {
        uint32_t h[4][256];
        memset(h, 0, sizeof(h));
        size_t i = 0;
        for (; i + 4 <= n_text; i += 4) { h[0][in[i]]++; h[1][in[i + 1]]++; h[2][in[i + 2]]++; h[3][in[i + 3]]++; }
        for (; i < n_text; ++i) h[0][in[i]]++;
        for (int c = 0; c < 256; ++c) freq[c] = (uint64_t)h[0][c] + h[1][c] + h[2][c] + h[3][c];
    }
    freq[256] = 1;                                     // end of block
    unsigned char llen[258];
    uint32_t lcode[257];
    huff_lengths(freq, 257, 15, llen);
    huff_codes(llen, 257, lcode);
    llen[257] = 0;                                     // the one distance code: zero bits = no distances used
    // the 258 code lengths are sent one by one with the code-length code (no run-length symbols: < 150 bytes per member)
    uint64_t cfreq[19] = {0};
    for (int i = 0; i < 258; ++i) cfreq[llen[i]]++;
    unsigned char clen[19];
    uint32_t ccode[19];
    huff_lengths(cfreq, 19, 7, clen);
    huff_codes(clen, 19, ccode);
    size_t bits = 0;
    for (int c = 0; c < 257; ++c) bits += freq[c] * llen[c];
    out.resize(10 + 64 + 19 * 3 / 8 + 258 + (bits + 7) / 8 + 8 + 16);
    unsigned char* o = reinterpret_cast<unsigned char*>(&out[0]);
    static const unsigned char gz_head[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3};
    memcpy(o, gz_head, 10);
    BitSink bs{o + 10};
    bs.put(1u, 1);                                     // BFINAL
    bs.put(2u, 2);                                     // BTYPE = dynamic Huffman
    bs.put(0u, 5);                                     // HLIT: 257 literal / length codes
    bs.put(0u, 5);                                     // HDIST: 1 distance code
    bs.put(15u, 4);                                    // HCLEN: all 19 code-length codes
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (int k = 0; k < 19; ++k) bs.put(clen[order[k]], 3);
    for (int i = 0; i < 258; ++i) bs.put(ccode[llen[i]], clen[llen[i]]);
    // the text: code and length of a byte in one table entry
    uint32_t tab[256];
    for (int c = 0; c < 256; ++c) tab[c] = (lcode[c] << 4) | llen[c];
    {
        // two bytes per drain check: the buffer holds fewer than 32 bits before and at most 30 more after
        unsigned char* p = bs.p;
        uint64_t buf = bs.buf;
        int n = bs.n;
        size_t i = 0;
        for (; i + 2 <= n_text; i += 2) {
            const uint32_t e0 = tab[in[i]], e1 = tab[in[i + 1]];
            buf |= (uint64_t)(e0 >> 4) << n;
            n += (int)(e0 & 15u);
            buf |= (uint64_t)(e1 >> 4) << n;
            n += (int)(e1 & 15u);
            if (n >= 32) {
                memcpy(p, &buf, 4);
                p += 4;
                buf >>= 32;
                n -= 32;
            }
        }
        bs.p = p;
        bs.buf = buf;
        bs.n = n;
        for (; i < n_text; ++i) {
            const uint32_t e = tab[in[i]];
            bs.put(e >> 4, (int)(e & 15u));
        }
    }
    bs.put(lcode[256], llen[256]);
    unsigned char* end = bs.finish();
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), in, (uInt)n_text), isize = (uint32_t)n_text;
    memcpy(end, &crc, 4);
    memcpy(end + 4, &isize, 4);
    out.resize((size_t)(end + 8 - o));
    return true;
}

bool deflate_member(const char* text, size_t n_text, int level, std::string& out) {
    if (level == 1 && n_text < ((size_t)1 << 31)) return huffman_member(text, n_text, out);
    return zlib_member(text, n_text, level, out);
}

}  // namespace

extern "C" int scdrs_gzip_member(const char* text, int64_t n_text, int32_t level, uint8_t* out, int64_t cap,
                                 int64_t* out_len) {
    SCDRS_CHECK(n_text >= 0 && (n_text == 0 || text) && out && out_len, "null pointer");
    std::string gz;
    SCDRS_CHECK(deflate_member(text ? text : "", (size_t)n_text, level, gz), "deflate failed");
    SCDRS_CHECK((int64_t)gz.size() <= cap, "output buffer too small (%zu bytes needed)", gz.size());
    memcpy(out, gz.data(), gz.size());
    *out_len = (int64_t)gz.size();
    return 0;
}

extern "C" int scdrs_format_f32(float v, char* out64) {
    char* e = fmt_f32(out64, v);
    *e = 0;
    return (int)(e - out64);
}

// fmt_f32_shortest (the function the GPU formats with) against std::to_chars for the float32 bit patterns
// [first, first + count): returns the number of mismatches, the first offending pattern in *bad_bits.
extern "C" int64_t scdrs_selftest_format(uint32_t first, uint64_t count, int32_t n_thread, uint32_t* bad_bits) {
    if (n_thread < 1) n_thread = 1;
    std::atomic<int64_t> bad{0};
    std::atomic<uint32_t> first_bad{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thread; ++t)
        pool.emplace_back([&, t]() {
            char a[80], b[80];
            for (uint64_t i = (uint64_t)t; i < count; i += (uint64_t)n_thread) {
                const uint32_t bits = first + (uint32_t)i;
                float v;
                memcpy(&v, &bits, 4);
                const int la = (int)(fmt_f32(a, v) - a);
                const int lb = scdrs::fmt_f32_shortest(v, b);
                if (la != lb || memcmp(a, b, (size_t)la) != 0) {
                    if (bad.fetch_add(1) == 0) first_bad.store(bits);
                }
            }
        });
    for (auto& th : pool) th.join();
    if (bad_bits) *bad_bits = first_bad.load();
    return bad.load();
}

// host restatement of format_slots_kernel (score_format.cu): what the GPU writes, for tests and timing probes
extern "C" int scdrs_format_slots_host(int64_t n_val, const float* vals, uint8_t* slots, int32_t n_thread) {
    SCDRS_CHECK(n_val >= 0 && (n_val == 0 || (vals && slots)), "null pointer");
    if (n_thread < 1) n_thread = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thread; ++t)
        pool.emplace_back([=]() {
            for (int64_t i = t; i < n_val; i += n_thread) {
                char c[16] = {0};
                c[15] = (char)scdrs::fmt_f32_shortest(vals[i], c);
                memcpy(slots + i * 16, c, 16);
            }
        });
    for (auto& th : pool) th.join();
    return 0;
}

extern "C" int scdrs_format_f32_shortest(float v, char* out64) {
    const int n = scdrs::fmt_f32_shortest(v, out64);
    out64[n] = 0;
    return n;
}

namespace {

// Row blocks of ~4 MB of text are produced by `row_fn(row, p) -> end of the row's fields` (after the row name),
// deflated as independent gzip members by n_thread threads and written in order.
template <typename RowFn>
int write_blocks(const char* path, const char* header, int64_t n_row, const char* names, const int64_t* name_off,
                 size_t field_bytes_per_row, RowFn row_fn, int32_t n_thread, int32_t level) {
    if (n_thread < 1) n_thread = 1;
    if (level < 0 || level > 9) level = 6;
    FILE* fh = fopen(path, "wb");
    SCDRS_CHECK(fh != nullptr, "cannot open %s for writing", path);
    int64_t rows_per_block = (int64_t)(4 << 20) / ((int64_t)(field_bytes_per_row * 11 / 16) + 32);
    if (rows_per_block < 16) rows_per_block = 16;
    const int64_t n_block = (n_row + rows_per_block - 1) / rows_per_block;
    std::vector<Block> blocks((size_t)n_block + 1);
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    const char* tr_env = getenv("SCDRS_B200_TRACE");
    const bool trace = tr_env != nullptr && tr_env[0] == '1';
    std::atomic<long long> ns_fmt{0}, ns_gz{0};
    const auto t_begin = std::chrono::steady_clock::now();
    auto work = [&]() {
        std::vector<char> text;          // one buffer per thread, written through a raw pointer
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b > n_block) return;
            const auto t0 = std::chrono::steady_clock::now();
            size_t n_text = 0;
            if (b == 0) {
                n_text = strlen(header);
                text.assign(header, header + n_text);
            } else {
                const int64_t r0 = (b - 1) * rows_per_block;
                const int64_t r1 = r0 + rows_per_block < n_row ? r0 + rows_per_block : n_row;
                const size_t cap = (size_t)(name_off[r1] - name_off[r0]) + (size_t)(r1 - r0) * (field_bytes_per_row + 2) + 64;
                if (text.size() < cap) text.resize(cap);
                char* p = text.data();
                for (int64_t r = r0; r < r1; ++r) {
                    const size_t ln = (size_t)(name_off[r + 1] - name_off[r]);
                    memcpy(p, names + name_off[r], ln);
                    p = row_fn(r, p + ln);
                    *p++ = '\n';
                }
                n_text = (size_t)(p - text.data());
            }
            const auto t1 = std::chrono::steady_clock::now();
            if (!deflate_member(text.data(), n_text, level, blocks[(size_t)b].gz)) failed.store(1);
            blocks[(size_t)b].ready.store(1, std::memory_order_release);
            if (trace) {
                const auto t2 = std::chrono::steady_clock::now();
                ns_fmt += std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
                ns_gz += std::chrono::duration_cast<std::chrono::nanoseconds>(t2 - t1).count();
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thread; ++t) pool.emplace_back(work);
    bool ok = true;
    for (int64_t b = 0; b <= n_block && ok; ++b) {
        while (!blocks[(size_t)b].ready.load(std::memory_order_acquire)) std::this_thread::yield();
        const std::string& gz = blocks[(size_t)b].gz;
        ok = fwrite(gz.data(), 1, gz.size(), fh) == gz.size();
        std::string().swap(blocks[(size_t)b].gz);
    }
    for (auto& t : pool) t.join();
    ok = (fclose(fh) == 0) && ok && !failed.load();
    if (trace)
        fprintf(stderr, "score writer %s: %.3f s wall, %d threads, thread-seconds: text %.2f, deflate %.2f\n", path,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(), n_thread,
                ns_fmt.load() * 1e-9, ns_gz.load() * 1e-9);
    SCDRS_CHECK(ok, "writing %s failed", path);
    return 0;
}

}  // namespace

extern "C" int scdrs_write_score_gz(const char* path, const char* header, int64_t n_row, int32_t n_col,
                                    const char* names, const int64_t* name_off, const float* values,
                                    int32_t n_thread, int32_t level) {
    SCDRS_CHECK(path && header && (n_row == 0 || (names && name_off && values)), "null pointer");
    SCDRS_CHECK(n_row >= 0 && n_col >= 0, "negative shape");
    auto row_fn = [=](int64_t r, char* p) {
        const float* row = values + r * (int64_t)n_col;
        for (int j = 0; j < n_col; ++j) {
            *p++ = '\t';
            p = fmt_f32(p, row[j]);
        }
        return p;
    };
    return write_blocks(path, header, n_row, names, name_off, (size_t)n_col * 16, row_fn, n_thread, level);
}

// The same file from two sources: n_lead float32 columns formatted here, then n_slot_col columns that the GPU
// already turned into text slots (score_format.cu: 16 bytes per value, text in bytes 0 .. len-1, len in byte 15).
extern "C" int scdrs_write_score_gz_slots(const char* path, const char* header, int64_t n_row, const char* names,
                                          const int64_t* name_off, int32_t n_lead, const float* lead,
                                          int32_t n_slot_col, const uint8_t* slots, int32_t n_thread, int32_t level) {
    SCDRS_CHECK(path && header && (n_row == 0 || (names && name_off)), "null pointer");
    SCDRS_CHECK(n_row >= 0 && n_lead >= 0 && n_slot_col >= 0, "negative shape");
    SCDRS_CHECK((n_lead == 0 || lead) && (n_slot_col == 0 || slots), "null pointer");
    auto row_fn = [=](int64_t r, char* p) {
        const float* row = lead + r * (int64_t)n_lead;
        for (int j = 0; j < n_lead; ++j) {
            *p++ = '\t';
            p = fmt_f32(p, row[j]);
        }
        const uint8_t* s = slots + (size_t)r * n_slot_col * 16;
        for (int j = 0; j < n_slot_col; ++j, s += 16) {
            *p++ = '\t';
            memcpy(p, s, 16);             // the tail beyond the length is overwritten by the next field
            p += s[15];
        }
        return p;
    };
    return write_blocks(path, header, n_row, names, name_off, (size_t)(n_lead + n_slot_col) * 17 + 32, row_fn, n_thread, level);
}
