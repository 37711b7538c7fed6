// K3: Monte-Carlo counts and the pooled empirical null of the reference
// (scdrs/method.py:205-206 and _get_p_from_empi_null, scdrs/method.py:601-632).
//
// The reference sorts the N = n_cell * n_ctrl pooled null values and binary-searches the
// M = n_cell trait scores in them; the only data-dependent output is the integer
// #{null >= t_i}.  Here the SMALL side is sorted instead:
//   1. radix-sort the M queries (LSD, 8-bit digits, stable warp-level ranking; keys + cell ids)
//   2. index them with a uniform grid: cell(v) = clamp(floor((v - lo) * scale)) is monotone in v,
//      so "queries <= v" = all queries in lower cells + a short search inside v's own cell.
//      The SAME function is applied to queries and null values, which makes the integer result
//      exact whatever the rounding of (v - lo) * scale is.
//   3. stream the null values ONCE (they are produced on the fly from the resident raw-score
//      matrix, so the normalised control matrix is never materialised unless asked for),
//      bump hist[#queries <= v] with one 64-bit reduction each, and count per cell how many
//      controls beat the trait (mc_pval).
//   4. #{null >= q_(s)} = N - sum_{p <= s} hist[p]  (exclusive scan), un-permuted to cell order.
// Across GPUs the queries are all-gathered, every rank counts its own null values against all of
// them and the histograms are summed (NCCL all-reduce): the pooled null never crosses NVLink.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace scdrs {

// ---- exclusive scan (3 kernels: tile sums, scan of sums, apply) --------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__device__ T block_exclusive_scan(T v, T* total, T* sh /* [33] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < (int)(blockDim.x >> 5) ? sh[lane] : (T)0;
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        sh[lane] = winc - w;
        if (lane == 31) sh[32] = winc;
    }
    __syncthreads();
    *total = sh[32];
    return sh[warp] + inc - v;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums_kernel(const T* __restrict__ in, int64_t n, T* __restrict__ sums) {
    __shared__ T sh[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) s += in[base + i];
    T total;
    block_exclusive_scan<T>(s, &total, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
scan_sums_kernel(T* __restrict__ sums, int64_t nb) {
    __shared__ T sh[33];
    T carry = 0;
    for (int64_t b0 = 0; b0 < nb; b0 += kScanThreads) {
        const int64_t i = b0 + threadIdx.x;
        const T v = i < nb ? sums[i] : (T)0;
        T total;
        const T ex = block_exclusive_scan<T>(v, &total, sh);
        if (i < nb) sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n,
                  const T* __restrict__ sums) {
    __shared__ T sh[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T v[kScanItems];
    T s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = base + i < n ? in[base + i] : (T)0;
        s += v[i];
    }
    T total;
    T ex = block_exclusive_scan<T>(s, &total, sh) + (sums != nullptr ? sums[blockIdx.x] : (T)0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

template <typename T>
static int exclusive_scan(scdrs_ctx* c, const T* in, T* out, int64_t n) {
    if (n <= 0) return 0;
    const int64_t nb = (n + kScanTile - 1) / kScanTile;
    if (nb == 1) {
        scan_apply_kernel<T><<<1, kScanThreads, 0, c->stream>>>(in, out, n, nullptr);
        SCDRS_LAUNCH_CHECK(c);
        return 0;
    }
    SCDRS_TRY(c->scan_tmp.reserve((size_t)nb * sizeof(T)));
    T* sums = c->scan_tmp.as<T>();
    scan_tile_sums_kernel<T><<<(unsigned)nb, kScanThreads, 0, c->stream>>>(in, n, sums);
    SCDRS_LAUNCH_CHECK(c);
    scan_sums_kernel<T><<<1, kScanThreads, 0, c->stream>>>(sums, nb);
    SCDRS_LAUNCH_CHECK(c);
    scan_apply_kernel<T><<<(unsigned)nb, kScanThreads, 0, c->stream>>>(in, out, n, sums);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

int exclusive_scan_u32(scdrs_ctx* c, const uint32_t* in, uint32_t* out, int64_t n) {
    return exclusive_scan<uint32_t>(c, in, out, n);
}
int exclusive_scan_u64(scdrs_ctx* c, const unsigned long long* in, unsigned long long* out,
                       int64_t n) {
    return exclusive_scan<unsigned long long>(c, in, out, n);
}

// ---- order-preserving keys ------------------------------------------------------------------
template <typename F> struct KeyOf;
template <> struct KeyOf<float> {
    typedef uint32_t type;
    __host__ __device__ static uint32_t make(float f) {
        f += 0.0f;  // -0 -> +0 so that key order == numeric order
#ifdef __CUDA_ARCH__
        uint32_t u = __float_as_uint(f);
#else
        uint32_t u; memcpy(&u, &f, 4);
#endif
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
};
template <> struct KeyOf<double> {
    typedef unsigned long long type;
    __host__ __device__ static unsigned long long make(double f) {
        f += 0.0;
#ifdef __CUDA_ARCH__
        unsigned long long u = (unsigned long long)__double_as_longlong(f);
#else
        unsigned long long u; memcpy(&u, &f, 8);
#endif
        return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
    }
};

template <typename F>
__global__ void make_keys_kernel(int64_t m, const F* __restrict__ q,
                                 typename KeyOf<F>::type* __restrict__ keys,
                                 uint32_t* __restrict__ idx) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m) {
        keys[i] = KeyOf<F>::make(q[i]);
        idx[i] = (uint32_t)i;
    }
}

// ---- LSD radix sort: one warp per chunk of kChunk keys ------------------------------------------
constexpr int kChunk = 512;
constexpr int kSortWarps = 4;

template <typename K>
__global__ void __launch_bounds__(kSortWarps * 32)
radix_hist_kernel(int64_t m, int64_t n_chunk, int shift, const K* __restrict__ keys,
                  uint32_t* __restrict__ table) {
    __shared__ uint32_t cnt[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t chunk = (int64_t)blockIdx.x * kSortWarps + warp;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    if (chunk < n_chunk) {
        const int64_t b = chunk * kChunk;
        const int64_t e = b + kChunk < m ? b + kChunk : m;
        for (int64_t i = b + lane; i < e; i += 32)
            atomicAdd(&cnt[warp][(int)((keys[i] >> shift) & 0xff)], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) table[(size_t)d * n_chunk + chunk] = cnt[warp][d];
    }
}

template <typename K>
__global__ void __launch_bounds__(kSortWarps * 32)
radix_scatter_kernel(int64_t m, int64_t n_chunk, int shift, const K* __restrict__ keys,
                     const uint32_t* __restrict__ idx, const uint32_t* __restrict__ table,
                     K* __restrict__ keys_out, uint32_t* __restrict__ idx_out) {
    __shared__ uint32_t base[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t chunk = (int64_t)blockIdx.x * kSortWarps + warp;
    if (chunk >= n_chunk) return;
    for (int d = lane; d < 256; d += 32) base[warp][d] = table[(size_t)d * n_chunk + chunk];
    __syncwarp();
    const int64_t b = chunk * kChunk;
    const int64_t e = b + kChunk < m ? b + kChunk : m;
    const uint32_t lt = (1u << lane) - 1u;
    for (int64_t i0 = b; i0 < e; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool ok = i < e;
        K key = 0;
        uint32_t id = 0;
        int digit = 256 + lane;  // inactive lanes never match anybody
        if (ok) {
            key = keys[i];
            id = idx[i];
            digit = (int)((key >> shift) & 0xff);
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const int rank = __popc(peers & lt);
        uint32_t pos = 0;
        if (ok) pos = base[warp][digit] + rank;
        __syncwarp();
        if (ok && rank == 0) base[warp][digit] += __popc(peers);
        __syncwarp();
        if (ok) {
            keys_out[pos] = key;
            idx_out[pos] = id;
        }
    }
}

// sorts c->sort_keys[0]/sort_idx[0] (m elements); returns which of the two buffers holds the result
template <typename K>
static int radix_sort(scdrs_ctx* c, int64_t m, int* result_buf) {
    const int64_t n_chunk = (m + kChunk - 1) / kChunk;
    SCDRS_TRY(c->sort_tab.reserve((size_t)256 * n_chunk * sizeof(uint32_t)));
    uint32_t* table = c->sort_tab.as<uint32_t>();
    const unsigned nblk = (unsigned)((n_chunk + kSortWarps - 1) / kSortWarps);
    int cur = 0;
    for (int pass = 0; pass < (int)sizeof(K); ++pass) {
        K* kin = c->sort_keys[cur].as<K>();
        K* kout = c->sort_keys[cur ^ 1].as<K>();
        uint32_t* iin = c->sort_idx[cur].as<uint32_t>();
        uint32_t* iout = c->sort_idx[cur ^ 1].as<uint32_t>();
        radix_hist_kernel<K><<<nblk, kSortWarps * 32, 0, c->stream>>>(m, n_chunk, pass * 8, kin, table);
        SCDRS_LAUNCH_CHECK(c);
        SCDRS_TRY(exclusive_scan_u32(c, table, table, 256 * n_chunk));
        radix_scatter_kernel<K><<<nblk, kSortWarps * 32, 0, c->stream>>>(m, n_chunk, pass * 8, kin, iin,
                                                                        table, kout, iout);
        SCDRS_LAUNCH_CHECK(c);
        cur ^= 1;
    }
    *result_buf = cur;
    return 0;
}

// ---- uniform grid over the sorted queries -----------------------------------------------------
template <typename F>
struct GridParams {
    F lo, scale;
    int n_cells;
};

template <typename F>
__device__ __forceinline__ int grid_cell(F v, F lo, F scale, int n_cells) {
    F t = (v - lo) * scale;
    t = t > (F)0 ? t : (F)0;  // also maps NaN to cell 0
    const F top = (F)(n_cells - 1);
    t = t < top ? t : top;
    return (int)t;
}

template <typename F> __device__ __forceinline__ F key_to_float(typename KeyOf<F>::type k);
template <> __device__ __forceinline__ float key_to_float<float>(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
template <> __device__ __forceinline__ double key_to_float<double>(unsigned long long k) {
    return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}

// params[0] = lo, params[1] = scale, taken from the extreme sorted queries
template <typename F>
__global__ void grid_params_kernel(int64_t m, int n_cells, const typename KeyOf<F>::type* __restrict__ keys,
                                   F* __restrict__ params) {
    const F lo = key_to_float<F>(keys[0]);
    const F hi = key_to_float<F>(keys[m - 1]);
    F scale = (F)0;
    if (hi > lo && isfinite(hi - lo)) scale = (F)n_cells / (hi - lo);
    if (!isfinite(scale)) scale = (F)0;
    params[0] = isfinite(lo) ? lo : (F)0;
    params[1] = scale;
}

// grid[c] = number of sorted queries whose cell is < c, for c in [0, n_cells]
// one thread per grid cell: binary search over the sorted queries (their cell index is monotone)
template <typename F>
__global__ void grid_build_kernel(int64_t m, int n_cells, const typename KeyOf<F>::type* __restrict__ keys,
                                  const F* __restrict__ params, uint32_t* __restrict__ grid) {
    const int64_t cc = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cc > n_cells) return;
    const F lo = params[0], scale = params[1];
    int64_t a = 0, b = m;
    while (a < b) {
        const int64_t mid = (a + b) >> 1;
        const F q = key_to_float<F>(keys[mid]);     // NaN queries sort last: treat them as beyond the grid
        const int64_t qc = (q != q) ? (int64_t)n_cells + 1 : grid_cell<F>(q, lo, scale, n_cells);
        if (qc < cc) a = mid + 1; else b = mid;
    }
    grid[cc] = (uint32_t)a;
}

// number of sorted queries <= v
template <typename F>
__device__ __forceinline__ uint32_t queries_le(F v, F lo, F scale, int n_cells,
                                               const uint32_t* __restrict__ grid,
                                               const typename KeyOf<F>::type* __restrict__ keys) {
    const int cell = grid_cell<F>(v, lo, scale, n_cells);
    uint32_t a = grid[cell], b = grid[cell + 1];
    const typename KeyOf<F>::type kv = KeyOf<F>::make(v);
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (keys[mid] <= kv) a = mid + 1; else b = mid;
    }
    return a;
}

// Cell table of the pooled-null pass: 8 bytes per grid cell, ONE load answers almost every null value.
//   .x = first sorted query of the cell | min(queries in the cell, 31) << 27      (m < 2^27)
//   .y = key of that first query
// no query in the cell -> position = first; one query -> first + (key <= value); more -> first when the value is
// below the first query, else a search among the cell's queries (31: the cell ends where the next one starts).
constexpr int kCellCntShift = 27;
constexpr uint32_t kCellCntMax = 31u, kCellStartMask = (1u << kCellCntShift) - 1u;

__global__ void grid_pack8_kernel(int n_cells, const uint32_t* __restrict__ grid, const uint32_t* __restrict__ keys,
                                  uint2* __restrict__ cells) {
    const int64_t cc = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cc > n_cells) return;
    const uint32_t a = grid[cc];
    uint32_t cnt = 0u, key0 = 0u;
    if (cc < n_cells) {
        cnt = grid[cc + 1] - a;
        if (cnt > 0u) key0 = keys[a];
        if (cnt > kCellCntMax) cnt = kCellCntMax;
    }
    cells[cc] = make_uint2(a | (cnt << kCellCntShift), key0);
}

// number of sorted queries <= v (float keys); same result as queries_le
__device__ __forceinline__ uint32_t queries_le_cells(float v, float lo, float scale, int n_cells,
                                                     const uint2* __restrict__ cells,
                                                     const uint32_t* __restrict__ keys) {
    const int cell = grid_cell<float>(v, lo, scale, n_cells);
    const uint2 e = __ldg(cells + cell);
    const uint32_t cnt = e.x >> kCellCntShift;
    uint32_t a = e.x & kCellStartMask;
    if (cnt == 0u) return a;
    const uint32_t kv = KeyOf<float>::make(v);
    if (kv < e.y) return a;
    ++a;                                              // the first query of the cell is <= v
    if (cnt == 1u) return a;
    uint32_t b = cnt < kCellCntMax ? a + cnt - 1u : (__ldg(&cells[cell + 1].x) & kCellStartMask);
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (keys[mid] <= kv) a = mid + 1; else b = mid;
    }
    return a;
}

static int choose_grid_cells(int64_t m) {
    // 8 - 16 cells per query (most cells then hold no query or one), at most 2^23 cells (64 MB of cell table)
    int64_t g = 1024;
    while (g < 8 * m && g < (1 << 23)) g <<= 1;
    return (int)g;
}

template <typename F>
static int prepare_queries(scdrs_ctx* c, const F* dev_q, int64_t m, int* sorted_buf) {
    typedef typename KeyOf<F>::type K;
    SCDRS_CHECK(m >= 1 && m < (int64_t)0xffffffff, "query count %lld out of range", (long long)m);
    for (int i = 0; i < 2; ++i) {
        SCDRS_TRY(c->sort_keys[i].reserve((size_t)m * sizeof(K)));
        SCDRS_TRY(c->sort_idx[i].reserve((size_t)m * sizeof(uint32_t)));
    }
    make_keys_kernel<F><<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(
        m, dev_q, c->sort_keys[0].as<K>(), c->sort_idx[0].as<uint32_t>());
    SCDRS_LAUNCH_CHECK(c);
    SCDRS_TRY(radix_sort<K>(c, m, sorted_buf));
    const int n_cells = choose_grid_cells(m);
    c->grid_bits = n_cells;
    SCDRS_TRY(c->grid.reserve(((size_t)n_cells + 1) * sizeof(uint32_t)));
    SCDRS_TRY(c->scalars.reserve(8 * sizeof(double)));
    F* params = reinterpret_cast<F*>(c->scalars.as<double>() + 2);
    grid_params_kernel<F><<<1, 1, 0, c->stream>>>(m, n_cells, c->sort_keys[*sorted_buf].as<K>(), params);
    SCDRS_LAUNCH_CHECK(c);
    grid_build_kernel<F><<<(unsigned)((n_cells + 256) / 256), 256, 0, c->stream>>>(
        m, n_cells, c->sort_keys[*sorted_buf].as<K>(), params, c->grid.as<uint32_t>());
    SCDRS_LAUNCH_CHECK(c);
    c->cells_ready = false;
    if (sizeof(F) == 4 && m < ((int64_t)1 << kCellCntShift)) {
        SCDRS_TRY(c->grid_cells.reserve(((size_t)n_cells + 1) * sizeof(uint2)));
        grid_pack8_kernel<<<(unsigned)((n_cells + 256) / 256), 256, 0, c->stream>>>(
            n_cells, c->grid.as<uint32_t>(), reinterpret_cast<const uint32_t*>(c->sort_keys[*sorted_buf].p),
            c->grid_cells.as<uint2>());
        SCDRS_LAUNCH_CHECK(c);
        c->cells_ready = true;
    }
    return 0;
}

int null_prepare_queries(scdrs_ctx* c, const float* dev_query_all, int64_t m) {
    return prepare_queries<float>(c, dev_query_all, m, &c->sorted_buf);
}

// ---- the pooled-null pass over the resident raw-score matrix ------------------------------------
// One warp per cell.  Recomputes the final normalised control scores (second alignment and zero
// rule, scdrs/method.py:565, :583), optionally stores them, counts controls >= trait (mc_pval) and
// bumps the pooled-null histogram.
template <bool CELLS>
__global__ void __launch_bounds__(256)
null_count_kernel(int64_t n_cell, int ncol, int ld, const float* __restrict__ dev_mat,
                  const double* __restrict__ row_ref, const double* __restrict__ colmean,
                  const double* __restrict__ ratio_a, const double* __restrict__ row_mean,
                  const double* __restrict__ row_istd, const double* __restrict__ col2,
                  const double* __restrict__ scalars, const float* __restrict__ query_local,
                  const uint32_t* __restrict__ grid, const uint2* __restrict__ cells,
                  const uint32_t* __restrict__ qkeys, int n_cells,
                  unsigned long long* __restrict__ hist, int32_t* __restrict__ mc_count,
                  float* __restrict__ norm_mat) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float gmin = (float)scalars[0];
    const float* gp = reinterpret_cast<const float*>(scalars + 2);
    const float lo = gp[0], scale = gp[1];
    const int n_ctrl = ncol - 1;
    for (int64_t row = warp0; row < n_cell; row += nwarp) {
        const double ref = row_ref[row], mu = row_mean[row], is = row_istd[row];
        const float t = query_local[row];
        const float* p = dev_mat + row * ld;
        int mc = 0;
        for (int j = 1 + lane; j < ncol; j += 32) {
            const float d = __ldcs(p + j);             // streamed once: the look-up tables should stay in L2
            float v;
            if (isnan(d)) {
                v = gmin;
            } else {
                const double z = __dmul_rn(ref + (double)d - colmean[j], ratio_a[j]);
                v = (float)(__dmul_rn(z - mu, is) - col2[j]);
            }
            if (norm_mat != nullptr) norm_mat[row * n_ctrl + (j - 1)] = v;
            mc += (v >= t) ? 1 : 0;
            const uint32_t pos = CELLS ? queries_le_cells(v, lo, scale, n_cells, cells, qkeys)
                                       : queries_le<float>(v, lo, scale, n_cells, grid, qkeys);
            atomicAdd(&hist[pos], 1ull);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mc += __shfl_xor_sync(0xffffffffu, mc, o);
        if (lane == 0) mc_count[row] = mc;
    }
}

int null_count_pass(scdrs_ctx* c, unsigned long long* dev_hist, int64_t m, bool write_norm) {
    SCDRS_TRY(c->mc_count.reserve((size_t)c->n_cell * sizeof(int32_t)));
    if (write_norm)
        SCDRS_TRY(c->norm_mat.reserve((size_t)c->n_cell * (c->n_ctrl > 0 ? c->n_ctrl : 1) * sizeof(float)));
    SCDRS_CUDA(cudaMemsetAsync(dev_hist, 0, (size_t)(m + 1) * sizeof(unsigned long long), c->stream));
    int64_t grid = (c->n_cell + 7) / 8;
    if (grid > kNumSM * 8) grid = kNumSM * 8;
    if (grid < 1) grid = 1;
    // the 8-byte cell table when the queries fit its 27-bit positions, else the plain grid + search
    auto kernel = c->cells_ready ? null_count_kernel<true> : null_count_kernel<false>;
    kernel<<<(unsigned)grid, 256, 0, c->stream>>>(
        c->n_cell, c->ncol, c->ld, c->dev_mat.as<float>(), c->row_ref.as<double>(),
        c->colmean.as<double>(), c->ratio_a.as<double>(), c->row_mean.as<double>(),
        c->row_istd.as<double>(), c->col2.as<double>(), c->scalars.as<double>(),
        c->query.as<float>(), c->grid.as<uint32_t>(), c->grid_cells.as<uint2>(),
        c->sort_keys[c->sorted_buf].as<uint32_t>(), c->grid_bits, dev_hist, c->mc_count.as<int32_t>(),
        write_norm ? c->norm_mat.as<float>() : nullptr);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

// ge[perm[s] - offset] = n_null_total - (exclusive prefix of hist)[s + 1]
__global__ void ge_scatter_kernel(int64_t m, int64_t offset, int64_t n_local, long long n_null_total,
                                  const unsigned long long* __restrict__ prefix,
                                  const uint32_t* __restrict__ perm, long long* __restrict__ ge_out) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= m) return;
    const int64_t i = (int64_t)perm[s] - offset;
    if (i >= 0 && i < n_local) ge_out[i] = n_null_total - (long long)prefix[s + 1];
}

static int finish_common(scdrs_ctx* c, const unsigned long long* dev_hist_all, int64_t m,
                         int64_t offset, int64_t n_null_total, int64_t n_local, int sorted_buf,
                         long long* dev_ge_out) {
    SCDRS_TRY(c->ge_sorted.reserve((size_t)(m + 2) * sizeof(unsigned long long)));
    unsigned long long* prefix = c->ge_sorted.as<unsigned long long>();
    SCDRS_TRY(exclusive_scan_u64(c, dev_hist_all, prefix, m + 1));
    ge_scatter_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(
        m, offset, n_local, (long long)n_null_total, prefix, c->sort_idx[sorted_buf].as<uint32_t>(),
        dev_ge_out);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

int null_finish(scdrs_ctx* c, const unsigned long long* dev_hist_all, int64_t m, int64_t offset,
                int64_t n_null_total, int64_t n_local, long long* dev_ge_out) {
    return finish_common(c, dev_hist_all, m, offset, n_null_total, n_local, c->sorted_buf, dev_ge_out);
}

// ---- stand-alone _get_p_from_empi_null over explicit arrays --------------------------------------
template <typename F>
__global__ void __launch_bounds__(256)
null_count_array_kernel(int64_t n_null, const F* __restrict__ vals, const F* __restrict__ params,
                        const uint32_t* __restrict__ grid,
                        const typename KeyOf<F>::type* __restrict__ qkeys, int n_cells,
                        unsigned long long* __restrict__ hist) {
    const F lo = params[0], scale = params[1];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_null;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t pos = queries_le<F>(vals[i], lo, scale, n_cells, grid, qkeys);
        atomicAdd(&hist[pos], 1ull);
    }
}

template <typename F>
static int null_count_generic(scdrs_ctx* c, int64_t n_t, const F* dev_t, int64_t n_null,
                              const F* dev_null, long long* dev_ge) {
    typedef typename KeyOf<F>::type K;
    int sorted_buf = 0;
    SCDRS_TRY(prepare_queries<F>(c, dev_t, n_t, &sorted_buf));
    SCDRS_TRY(c->hist.reserve((size_t)(n_t + 1) * sizeof(unsigned long long)));
    unsigned long long* hist = c->hist.as<unsigned long long>();
    SCDRS_CUDA(cudaMemsetAsync(hist, 0, (size_t)(n_t + 1) * sizeof(unsigned long long), c->stream));
    if (n_null > 0) {
        int64_t grid = (n_null + 255) / 256;
        if (grid > kNumSM * 16) grid = kNumSM * 16;
        const F* params = reinterpret_cast<const F*>(c->scalars.as<double>() + 2);
        null_count_array_kernel<F><<<(unsigned)grid, 256, 0, c->stream>>>(
            n_null, dev_null, params, c->grid.as<uint32_t>(), c->sort_keys[sorted_buf].as<K>(),
            c->grid_bits, hist);
        SCDRS_LAUNCH_CHECK(c);
    }
    return finish_common(c, hist, n_t, 0, n_null, n_t, sorted_buf, dev_ge);
}

int null_count_generic_f32(scdrs_ctx* c, int64_t n_t, const float* dev_t, int64_t n_null,
                           const float* dev_null, long long* dev_ge) {
    return null_count_generic<float>(c, n_t, dev_t, n_null, dev_null, dev_ge);
}
int null_count_generic_f64(scdrs_ctx* c, int64_t n_t, const double* dev_t, int64_t n_null,
                           const double* dev_null, long long* dev_ge) {
    return null_count_generic<double>(c, n_t, dev_t, n_null, dev_null, dev_ge);
}

}  // namespace scdrs
