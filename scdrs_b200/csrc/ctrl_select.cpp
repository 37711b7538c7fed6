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
#include <string.h>

#include <vector>

#include "common.cuh"

// Plain C++ (compiled by the host compiler through nvcc): the refill is cloned for AVX2 and picked at load time.
//
// Speed.  A trait at config-2 size is ~1.8e7 index draws (~2.4e7 generator outputs) on one core and sits on
// the critical path of `score_traits` (the GPU scores a trait in 11 ms).  Two things make the loop 2.1x
// faster than the textbook form (164 -> 78 ms on the build host): the generator refills and tempers 624
// outputs at a time (vectorised), and the rejection loop has no data-dependent branch -- all bounds t that
// share one bit mask are handled by one inner loop whose only loop-carried dependency is `t -= accepted`;
// a rejected draw swaps perm[t] with itself.

namespace {

struct MT19937 {
    uint32_t mt[624];
    uint32_t out[624];      // the 624 tempered outputs of the current block
    int pos;
    explicit MT19937(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        pos = 624;
    }
};

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
void mt_refill(uint32_t* __restrict__ mt, uint32_t* __restrict__ out) {
    const uint32_t hi = 0x80000000u, lo = 0x7fffffffu, a = 0x9908b0dfu;
    int i = 0;
    for (; i < 624 - 397; ++i) {
        const uint32_t y = (mt[i] & hi) | (mt[i + 1] & lo);
        mt[i] = mt[i + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & a);
    }
    for (; i < 623; ++i) {
        const uint32_t y = (mt[i] & hi) | (mt[i + 1] & lo);
        mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((0u - (y & 1u)) & a);
    }
    const uint32_t y = (mt[623] & hi) | (mt[0] & lo);
    mt[623] = mt[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & a);
    for (int k = 0; k < 624; ++k) {
        uint32_t v = mt[k];
        v ^= v >> 11;
        v ^= (v << 7) & 0x9d2c5680u;
        v ^= (v << 15) & 0xefc60000u;
        v ^= v >> 18;
        out[k] = v;
    }
}

// permutation(n) of numpy's legacy generator into perm[0 .. n): for t = n-1 .. 1, swap perm[t] with perm[j],
// j = the next output AND (smallest all-ones mask >= t), redrawn while j > t.
inline void shuffle_from_top(MT19937& rng, int32_t* __restrict__ perm, int32_t n) {
    int32_t t = n - 1;
    int pos = rng.pos;
    while (t >= 1) {
        const uint32_t mask = 0xffffffffu >> __builtin_clz((uint32_t)t);
        const int32_t lo = (int32_t)((mask >> 1) + 1);        // bounds lo .. t share this mask
        while (t >= lo) {
            if (pos == 624) {
                mt_refill(rng.mt, rng.out);
                pos = 0;
            }
            const uint32_t* w = rng.out;
            while (pos < 624 && t >= lo) {
                const uint32_t v = w[pos++] & mask;
                const bool accepted = v <= (uint32_t)t;
                const uint32_t j = accepted ? v : (uint32_t)t;
                const int32_t x = perm[t], y = perm[j];
                perm[t] = y;
                perm[j] = x;
                t -= accepted;
            }
        }
    }
    rng.pos = pos;
}

}  // namespace

extern "C" int scdrs_select_ctrl_state(uint32_t seed, int32_t n_bin, const int32_t* bin_ptr,
                                       const int32_t* bin_genes, const int32_t* bin_ntrait, int32_t n_ctrl,
                                       int32_t n_s, int32_t* ctrl_genes, uint32_t* mt_state_out) {
    SCDRS_CHECK(n_bin >= 0 && n_ctrl >= 0 && n_s >= 0, "negative size");
    SCDRS_CHECK(bin_ptr && bin_genes && bin_ntrait && (ctrl_genes || n_ctrl == 0), "null pointer");
    MT19937 rng(seed);
    std::vector<int32_t> perm, ident;
    int64_t off = 0;
    for (int b = 0; b < n_bin; ++b) {
        const int32_t k = bin_ntrait[b];
        if (k <= 0) continue;
        const int32_t n = bin_ptr[b + 1] - bin_ptr[b];
        SCDRS_CHECK(k <= n, "bin %d holds %d genes but %d trait genes", b, n, k);
        SCDRS_CHECK(off + k <= n_s, "trait genes per bin exceed n_s");
        const int32_t* genes = bin_genes + bin_ptr[b];
        perm.resize(n);
        ident.resize(n);
        for (int32_t t = 0; t < n; ++t) ident[t] = t;
        for (int i = 0; i < n_ctrl; ++i) {
            memcpy(perm.data(), ident.data(), (size_t)n * sizeof(int32_t));
            shuffle_from_top(rng, perm.data(), n);
            int32_t* out = ctrl_genes + (int64_t)i * n_s + off;
            for (int32_t t = 0; t < k; ++t) out[t] = genes[perm[t]];
        }
        off += k;
    }
    SCDRS_CHECK(off == n_s, "trait genes per bin sum to %lld, expected n_s = %d", (long long)off, n_s);
    if (mt_state_out != nullptr) {      // numpy's legacy state after the same draws: key[624], then pos
        memcpy(mt_state_out, rng.mt, 624 * sizeof(uint32_t));
        mt_state_out[624] = (uint32_t)rng.pos;
    }
    return 0;
}

extern "C" int scdrs_select_ctrl(uint32_t seed, int32_t n_bin, const int32_t* bin_ptr,
                                 const int32_t* bin_genes, const int32_t* bin_ntrait, int32_t n_ctrl,
                                 int32_t n_s, int32_t* ctrl_genes) {
    return scdrs_select_ctrl_state(seed, n_bin, bin_ptr, bin_genes, bin_ntrait, n_ctrl, n_s, ctrl_genes, nullptr);
}
