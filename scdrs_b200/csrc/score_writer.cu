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
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <charconv>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

inline char* fmt_f32(char* p, float v) {
    if (std::isnan(v)) return p;                       // na_rep = ''
    if (std::isinf(v)) {
        if (v < 0) *p++ = '-';
        memcpy(p, "inf", 3);
        return p + 3;
    }
    const long double a = std::fabs((long double)v);
    if (a == 0.0L || (a >= 1.e-4L && a < 1.e6L)) {
        auto r = std::to_chars(p, p + 64, v, std::chars_format::fixed);
        bool has_dot = false;
        for (char* q = p; q < r.ptr; ++q) has_dot |= (*q == '.');
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

bool deflate_member(const std::string& text, int level, std::string& out) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, text.size()) + 64);
    zs.next_in = (Bytef*)text.data();
    zs.avail_in = (uInt)text.size();
    zs.next_out = (Bytef*)out.data();
    zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return rc == Z_STREAM_END;
}

}  // namespace

extern "C" int scdrs_format_f32(float v, char* out64) {
    char* e = fmt_f32(out64, v);
    *e = 0;
    return (int)(e - out64);
}

extern "C" int scdrs_write_score_gz(const char* path, const char* header, int64_t n_row, int32_t n_col,
                                    const char* names, const int64_t* name_off, const float* values,
                                    int32_t n_thread, int32_t level) {
    SCDRS_CHECK(path && header && (n_row == 0 || (names && name_off && values)), "null pointer");
    SCDRS_CHECK(n_row >= 0 && n_col >= 0, "negative shape");
    if (n_thread < 1) n_thread = 1;
    if (level < 0 || level > 9) level = 6;
    FILE* fh = fopen(path, "wb");
    SCDRS_CHECK(fh != nullptr, "cannot open %s for writing", path);
    // block size: about 4 MB of text per gzip member
    int64_t rows_per_block = (int64_t)(4 << 20) / ((int64_t)n_col * 11 + 32);
    if (rows_per_block < 16) rows_per_block = 16;
    const int64_t n_block = (n_row + rows_per_block - 1) / rows_per_block;
    std::vector<Block> blocks((size_t)n_block + 1);
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        std::string text;
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b > n_block) return;
            text.clear();
            if (b == 0) {
                text = header;
            } else {
                const int64_t r0 = (b - 1) * rows_per_block;
                const int64_t r1 = r0 + rows_per_block < n_row ? r0 + rows_per_block : n_row;
                text.reserve((size_t)(r1 - r0) * ((size_t)n_col * 12 + 48));
                char buf[80];
                for (int64_t r = r0; r < r1; ++r) {
                    text.append(names + name_off[r], (size_t)(name_off[r + 1] - name_off[r]));
                    const float* row = values + r * (int64_t)n_col;
                    for (int j = 0; j < n_col; ++j) {
                        buf[0] = '\t';
                        char* e = fmt_f32(buf + 1, row[j]);
                        text.append(buf, (size_t)(e - buf));
                    }
                    text.push_back('\n');
                }
            }
            if (!deflate_member(text, level, blocks[(size_t)b].gz)) failed.store(1);
            blocks[(size_t)b].ready.store(1, std::memory_order_release);
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
    SCDRS_CHECK(ok, "writing %s failed", path);
    return 0;
}
