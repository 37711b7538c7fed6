// float32 -> the text pandas' DataFrame.to_csv writes for it (numpy's shortest round-trip digits; positional for
// 1e-4 <= |x| < 1e6 with a trailing ".0" for integers, scientific with a two-digit exponent otherwise, empty for
// NaN), as one __host__ __device__ function: the GPU formats the n_cell x n_ctrl control-score block of a
// `.full_score.gz` file (docs/file_format.rst:81-108 of the reference; bin/scdrs:276-288), the host formats the
// rest, and `scdrs_selftest_format` checks this function against std::to_chars for every float32 bit pattern.
//
// Shortest digits: the Ryu algorithm (Ulf Adams, "Ryu: fast float-to-string conversion", PLDI 2018), restated for
// 32-bit floats from the paper's description; tables in fmt_tables.cuh.
#pragma once
#include <stdint.h>
#include <string.h>

#include "fmt_tables.cuh"

namespace scdrs {

#ifdef __CUDA_ARCH__
#define SCDRS_POW5_INV(i) kPow5InvSplitDev[i]
#define SCDRS_POW5(i) kPow5SplitDev[i]
#else
#define SCDRS_POW5_INV(i) kPow5InvSplitHost[i]
#define SCDRS_POW5(i) kPow5SplitHost[i]
#endif

__host__ __device__ inline uint32_t ryu_pow5bits(int32_t e) { return (uint32_t)(((uint32_t)e * 1217359u) >> 19) + 1u; }
__host__ __device__ inline uint32_t ryu_log10pow2(int32_t e) { return ((uint32_t)e * 78913u) >> 18; }
__host__ __device__ inline uint32_t ryu_log10pow5(int32_t e) { return ((uint32_t)e * 732923u) >> 20; }

__host__ __device__ inline uint32_t ryu_mulshift(uint32_t m, uint64_t factor, int32_t shift) {
    const uint64_t lo = (uint64_t)m * (uint32_t)factor;
    const uint64_t hi = (uint64_t)m * (uint32_t)(factor >> 32);
    return (uint32_t)(((lo >> 32) + hi) >> (shift - 32));
}

__host__ __device__ inline bool ryu_multiple_of_pow5(uint32_t v, uint32_t p) {
    uint32_t c = 0;
    for (;;) {
        const uint32_t q = v / 5u, r = v - 5u * q;
        if (r != 0u) break;
        v = q;
        ++c;
    }
    return c >= p;
}

// value = digits * 10^exp10 (digits has no trailing zeros beyond what the shortest representation keeps)
__host__ __device__ inline void f32_shortest_decimal(uint32_t mant, uint32_t expo, uint32_t* digits_out, int32_t* exp10_out) {
    int32_t e2;
    uint32_t m2;
    if (expo == 0u) {
        e2 = 1 - 127 - 23 - 2;
        m2 = mant;
    } else {
        e2 = (int32_t)expo - 127 - 23 - 2;
        m2 = (1u << 23) | mant;
    }
    const bool accept = (m2 & 1u) == 0u;
    const uint32_t mv = 4u * m2, mp = 4u * m2 + 2u;
    const uint32_t mm_shift = (mant != 0u || expo <= 1u) ? 1u : 0u;
    const uint32_t mm = 4u * m2 - 1u - mm_shift;
    uint32_t vr, vp, vm;
    int32_t e10;
    bool vm_tz = false, vr_tz = false;
    uint32_t last = 0u;
    if (e2 >= 0) {
        const uint32_t q = ryu_log10pow2(e2);
        e10 = (int32_t)q;
        const int32_t k = 59 + (int32_t)ryu_pow5bits((int32_t)q) - 1;
        const int32_t i = -e2 + (int32_t)q + k;
        vr = ryu_mulshift(mv, SCDRS_POW5_INV(q), i);
        vp = ryu_mulshift(mp, SCDRS_POW5_INV(q), i);
        vm = ryu_mulshift(mm, SCDRS_POW5_INV(q), i);
        if (q != 0u && (vp - 1u) / 10u <= vm / 10u) {
            const int32_t l = 59 + (int32_t)ryu_pow5bits((int32_t)q - 1) - 1;
            last = ryu_mulshift(mv, SCDRS_POW5_INV(q - 1u), -e2 + (int32_t)q - 1 + l) % 10u;
        }
        if (q <= 9u) {
            if (mv % 5u == 0u) vr_tz = ryu_multiple_of_pow5(mv, q);
            else if (accept) vm_tz = ryu_multiple_of_pow5(mm, q);
            else vp -= ryu_multiple_of_pow5(mp, q) ? 1u : 0u;
        }
    } else {
        const uint32_t q = ryu_log10pow5(-e2);
        e10 = (int32_t)q + e2;
        const int32_t i = -e2 - (int32_t)q;
        const int32_t k = (int32_t)ryu_pow5bits(i) - 61;
        int32_t j = (int32_t)q - k;
        vr = ryu_mulshift(mv, SCDRS_POW5(i), j);
        vp = ryu_mulshift(mp, SCDRS_POW5(i), j);
        vm = ryu_mulshift(mm, SCDRS_POW5(i), j);
        if (q != 0u && (vp - 1u) / 10u <= vm / 10u) {
            j = (int32_t)q - 1 - ((int32_t)ryu_pow5bits(i + 1) - 61);
            last = ryu_mulshift(mv, SCDRS_POW5(i + 1), j) % 10u;
        }
        if (q <= 1u) {
            vr_tz = true;
            if (accept) vm_tz = mm_shift == 1u;
            else --vp;
        } else if (q < 31u) {
            vr_tz = (mv & ((1u << (q - 1u)) - 1u)) == 0u;
        }
    }
    int32_t removed = 0;
    uint32_t out;
    if (vm_tz || vr_tz) {
        while (vp / 10u > vm / 10u) {
            vm_tz &= vm % 10u == 0u;
            vr_tz &= last == 0u;
            last = vr % 10u;
            vr /= 10u; vp /= 10u; vm /= 10u;
            ++removed;
        }
        if (vm_tz) {
            while (vm % 10u == 0u) {
                vr_tz &= last == 0u;
                last = vr % 10u;
                vr /= 10u; vp /= 10u; vm /= 10u;
                ++removed;
            }
        }
        if (vr_tz && last == 5u && vr % 2u == 0u) last = 4u;      // round half to even
        out = vr + (((vr == vm && (!accept || !vm_tz)) || last >= 5u) ? 1u : 0u);
    } else {
        while (vp / 10u > vm / 10u) {
            last = vr % 10u;
            vr /= 10u; vp /= 10u; vm /= 10u;
            ++removed;
        }
        out = vr + ((vr == vm || last >= 5u) ? 1u : 0u);
    }
    *digits_out = out;
    *exp10_out = e10 + removed;
}

// Writes the text of v at p (at most 16 characters, no terminator); returns the length.  NaN -> 0 characters.
__host__ __device__ inline int fmt_f32_shortest(float v, char* p) {
    uint32_t bits;
#ifdef __CUDA_ARCH__
    bits = __float_as_uint(v);
#else
    memcpy(&bits, &v, 4);
#endif
    const uint32_t mant = bits & 0x7fffffu, expo = (bits >> 23) & 0xffu;
    char* const p0 = p;
    if (expo == 0xffu) {
        if (mant != 0u) return 0;                                   // NaN: na_rep = ''
        if (bits >> 31) *p++ = '-';
        *p++ = 'i'; *p++ = 'n'; *p++ = 'f';
        return (int)(p - p0);
    }
    if (bits >> 31) *p++ = '-';
    if (expo == 0u && mant == 0u) {
        *p++ = '0'; *p++ = '.'; *p++ = '0';
        return (int)(p - p0);
    }
    uint32_t dig;
    int32_t e10;
    f32_shortest_decimal(mant, expo, &dig, &e10);
    char d[10];
    int n = 0;
    {
        char rev[10];
        uint32_t t = dig;
        while (t != 0u) { rev[n++] = (char)('0' + t % 10u); t /= 10u; }
        for (int i = 0; i < n; ++i) d[i] = rev[n - 1 - i];
    }
    const int32_t sci = e10 + n - 1;                                // decimal exponent of the first digit
    // positional for 1e-4 <= |v| < 1e6, decided on the VALUE like the host writer (the float just below 1e-4 has
    // the shortest digits "1e-04" and stays scientific)
    const double a = v < 0.f ? -(double)v : (double)v;
    if (a >= 1.e-4 && a < 1.e6) {
        if (sci >= 0) {
            const int n_int = sci + 1;
            for (int i = 0; i < n_int; ++i) *p++ = i < n ? d[i] : '0';
            *p++ = '.';
            if (n > n_int) for (int i = n_int; i < n; ++i) *p++ = d[i];
            else *p++ = '0';
        } else {
            *p++ = '0'; *p++ = '.';
            for (int i = 0; i < -sci - 1; ++i) *p++ = '0';
            for (int i = 0; i < n; ++i) *p++ = d[i];
        }
        return (int)(p - p0);
    }
    *p++ = d[0];
    if (n > 1) {
        *p++ = '.';
        for (int i = 1; i < n; ++i) *p++ = d[i];
    }
    *p++ = 'e';
    int32_t e = sci;
    if (e < 0) { *p++ = '-'; e = -e; } else { *p++ = '+'; }
    *p++ = (char)('0' + e / 10);
    *p++ = (char)('0' + e % 10);
    return (int)(p - p0);
}

#undef SCDRS_POW5_INV
#undef SCDRS_POW5

}  // namespace scdrs
