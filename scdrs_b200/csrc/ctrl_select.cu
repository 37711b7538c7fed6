// Host code: the random draws of _select_ctrl_geneset (scdrs/method.py:276, 295-307 of the
// reference) without the Python interpreter in the loop.
//
// The reference calls np.random.choice(bin_genes, size=k, replace=False) on numpy's legacy global
// generator once per (bin, control set): about n_bins * n_ctrl = 4e5 calls per trait.  That call
// is permutation(len(bin_genes))[:k]: MT19937 seeded with init_genrand(seed), a Fisher-Yates
// shuffle from the top (i = n-1 .. 1) whose index draw is "next 32-bit output AND smallest
// all-ones mask >= i, rejected while > i" (numpy/random/mtrand.pyx: RandomState.shuffle /
// legacy random_interval).  Restated here so control sets stay bit-identical to numpy's; the
// tests compare it with numpy.random.RandomState draw for draw.
#include <stdint.h>

#include <vector>

#include "common.cuh"

namespace {

struct MT19937 {
    uint32_t mt[624];
    int pos;
    explicit MT19937(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        pos = 624;
    }
    void refill() {
        const uint32_t hi = 0x80000000u, lo = 0x7fffffffu, a = 0x9908b0dfu;
        int i = 0;
        for (; i < 624 - 397; ++i) {
            const uint32_t y = (mt[i] & hi) | (mt[i + 1] & lo);
            mt[i] = mt[i + 397] ^ (y >> 1) ^ ((y & 1u) ? a : 0u);
        }
        for (; i < 623; ++i) {
            const uint32_t y = (mt[i] & hi) | (mt[i + 1] & lo);
            mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? a : 0u);
        }
        const uint32_t y = (mt[623] & hi) | (mt[0] & lo);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? a : 0u);
        pos = 0;
    }
    inline uint32_t next() {
        if (pos == 624) refill();
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    inline uint32_t interval(uint32_t max) {  // uniform on [0, max]
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        uint32_t v;
        while ((v = next() & mask) > max) {}
        return v;
    }
};

}  // namespace

extern "C" int scdrs_select_ctrl(uint32_t seed, int32_t n_bin, const int32_t* bin_ptr,
                                 const int32_t* bin_genes, const int32_t* bin_ntrait, int32_t n_ctrl,
                                 int32_t n_s, int32_t* ctrl_genes) {
    SCDRS_CHECK(n_bin >= 0 && n_ctrl >= 0 && n_s >= 0, "negative size");
    SCDRS_CHECK(bin_ptr && bin_genes && bin_ntrait && (ctrl_genes || n_ctrl == 0), "null pointer");
    MT19937 rng(seed);
    std::vector<int32_t> perm;
    int64_t off = 0;
    for (int b = 0; b < n_bin; ++b) {
        const int32_t k = bin_ntrait[b];
        if (k <= 0) continue;
        const int32_t n = bin_ptr[b + 1] - bin_ptr[b];
        SCDRS_CHECK(k <= n, "bin %d holds %d genes but %d trait genes", b, n, k);
        SCDRS_CHECK(off + k <= n_s, "trait genes per bin exceed n_s");
        const int32_t* genes = bin_genes + bin_ptr[b];
        perm.resize(n);
        for (int i = 0; i < n_ctrl; ++i) {
            for (int32_t t = 0; t < n; ++t) perm[t] = t;
            for (int32_t t = n - 1; t >= 1; --t) {
                const uint32_t j = rng.interval((uint32_t)t);
                const int32_t tmp = perm[t];
                perm[t] = perm[j];
                perm[j] = tmp;
            }
            int32_t* out = ctrl_genes + (int64_t)i * n_s + off;
            for (int32_t t = 0; t < k; ++t) out[t] = genes[perm[t]];
        }
        off += k;
    }
    SCDRS_CHECK(off == n_s, "trait genes per bin sum to %lld, expected n_s = %d", (long long)off, n_s);
    return 0;
}
