// K2: _correct_background of the reference (scdrs/method.py:482-598) as fused row/column
// reductions over the resident raw-score matrix.
//
// The matrix is held as  raw[c, j] = row_ref[c] + dev_mat[c, j]  (fp64 reference + fp32
// deviation; NaN deviation = raw score exactly 0, scdrs/method.py:522-523).  All arithmetic is
// fp64 in registers, as in the reference; only storage is fp32.
//
//   colsum     : column sums of raw                                   (:526-527, the means)
//   rowstats   : z = (raw - colmean) / sqrt(ratio)                    (:527-528)
//                row mean / std(ddof=0) over the control columns      (:544-545)
//                n = (z - mean) / std ; column sums and minima of n   (:547-548, :564-565, :581)
//   queries    : trait column: n - colmean2, zero rule                (:564, :582)
//   (the control columns are finalised inside the pooled-null pass, empi_null.cu)
//
// Every column reduction is "per-CTA partial in a fixed order, then a fixed-order reduce", so
// results are bitwise reproducible.
#include "common.cuh"

namespace scdrs {

constexpr int kRowBlock = 32;    // rows per CTA work item
constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double raw_value(double ref, float d) {
    return isnan(d) ? 0.0 : ref + (double)d;
}

// minimum that propagates NaN like numpy's min (scdrs/method.py:581)
__device__ __forceinline__ double min_nan(double a, double b) {
    return (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000ll) : fmin(a, b);
}

// first gene-set alignment of one entry; rounded product (no fma contraction) so that every kernel
// that recomputes it gets the same bits
__device__ __forceinline__ double aligned_value(double ref, float d, double colmean, double a) {
    return __dmul_rn(raw_value(ref, d) - colmean, a);
}

// ---- column sums of the raw matrix -----------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
colsum_kernel(int64_t n_cell, int ncol, int ld, const float* __restrict__ dev_mat,
              const double* __restrict__ row_ref, double* __restrict__ partial) {
    extern __shared__ double sh_part[];  // [ncol]
    __shared__ double sh_ref[kRowBlock];
    for (int j = threadIdx.x; j < ncol; j += kThreads) sh_part[j] = 0.0;
    const int64_t nblk = (n_cell + kRowBlock - 1) / kRowBlock;
    for (int64_t b = blockIdx.x; b < nblk; b += gridDim.x) {
        const int64_t r0 = b * kRowBlock;
        const int nr = (int)((n_cell - r0) < kRowBlock ? (n_cell - r0) : kRowBlock);
        __syncthreads();
        if (threadIdx.x < nr) sh_ref[threadIdx.x] = row_ref[r0 + threadIdx.x];
        __syncthreads();
        for (int j = threadIdx.x; j < ncol; j += kThreads) {
            double s = 0.0;
            const float* p = dev_mat + r0 * ld + j;
#pragma unroll 8
            for (int r = 0; r < nr; ++r) s += raw_value(sh_ref[r], p[(int64_t)r * ld]);
            sh_part[j] += s;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ncol; j += kThreads)
        partial[(size_t)blockIdx.x * ncol + j] = sh_part[j];
}

// out[j] = sum_b partial[b][j]  (op 0) or min_b (op 1), in a fixed order: 32 columns per block,
// 16 warps stride over the partials (coalesced 256-byte rows), then one ordered pass over the 16.
constexpr int kRedWarps = 16;
__global__ void __launch_bounds__(kRedWarps * 32)
reduce_partials_kernel(int nblk, int ncol, int op, const double* __restrict__ partial,
                       double* __restrict__ out) {
    __shared__ double sh[kRedWarps][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    double v = op == 0 ? 0.0 : INFINITY;
    if (j < ncol) {
        for (int b = ty; b < nblk; b += kRedWarps) {
            const double x = partial[(size_t)b * ncol + j];
            v = op == 0 ? v + x : min_nan(v, x);
        }
    }
    sh[ty][tx] = v;
    __syncthreads();
    if (ty == 0 && j < ncol) {
        double r = sh[0][tx];
        for (int w = 1; w < kRedWarps; ++w) r = op == 0 ? r + sh[w][tx] : min_nan(r, sh[w][tx]);
        out[j] = r;
    }
}

static int grid_for(int64_t n_cell) {
    const int64_t nblk = (n_cell + kRowBlock - 1) / kRowBlock;
    const int64_t g = (int64_t)kNumSM * 4;
    return (int)(nblk < g ? (nblk < 1 ? 1 : nblk) : g);
}

int bg_colsum(scdrs_ctx* c, double* dev_colsum_out) {
    SCDRS_CHECK(c->trait_ready, "no raw scores resident (call scdrs_trait_raw first)");
    if (c->colsum_ready) {
        // already known by algebra from the per-gene sums of X (prep_sets_kernel): no pass over the matrix
        SCDRS_CUDA(cudaMemcpyAsync(dev_colsum_out, c->colsum_alg.p, (size_t)c->ncol * sizeof(double),
                                   cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    const int g = grid_for(c->n_cell);
    SCDRS_TRY(c->partial.reserve((size_t)g * c->ncol * sizeof(double) * 2));
    SCDRS_CHECK((size_t)c->ncol * 8 <= 160 * 1024, "1 + n_ctrl too large");
    static bool attr = false;
    if (!attr) {
        SCDRS_CUDA(cudaFuncSetAttribute(colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    colsum_kernel<<<g, kThreads, (size_t)c->ncol * sizeof(double), c->stream>>>(
        c->n_cell, c->ncol, c->ld, c->dev_mat.as<float>(), c->row_ref.as<double>(),
        c->partial.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    reduce_partials_kernel<<<(c->ncol + 31) / 32, kRedWarps * 32, 0, c->stream>>>(
        g, c->ncol, 0, c->partial.as<double>(), dev_colsum_out);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

// ---- row statistics + second-alignment column statistics --------------------------------------
__global__ void colmean_kernel(int ncol, double inv_n, const double* __restrict__ colsum,
                               double* __restrict__ colmean) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ncol) colmean[j] = colsum[j] * inv_n;
}

__global__ void __launch_bounds__(kThreads)
rowstats_kernel(int64_t n_cell, int ncol, int ld, const float* __restrict__ dev_mat,
                const double* __restrict__ row_ref, const double* __restrict__ colmean,
                const double* __restrict__ ratio_a, double* __restrict__ row_mean,
                double* __restrict__ row_istd, double* __restrict__ part_sum,
                double* __restrict__ part_min) {
    extern __shared__ double sh[];
    double* sh_cm = sh;             // [ncol] column means of raw
    double* sh_a = sh + ncol;       // [ncol] 1/sqrt(ratio)
    double* sh_sum = sh + 2 * ncol; // [ncol] partial sums of n
    double* sh_min = sh + 3 * ncol; // [ncol] partial minima of n
    __shared__ double sh_ref[kRowBlock], sh_mu[kRowBlock], sh_is[kRowBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kThreads / 32;
    for (int j = threadIdx.x; j < ncol; j += kThreads) {
        sh_cm[j] = colmean[j];
        sh_a[j] = ratio_a[j];
        sh_sum[j] = 0.0;
        sh_min[j] = INFINITY;
    }
    const int n_ctrl = ncol - 1;
    const int64_t nblk = (n_cell + kRowBlock - 1) / kRowBlock;
    for (int64_t b = blockIdx.x; b < nblk; b += gridDim.x) {
        const int64_t r0 = b * kRowBlock;
        const int nr = (int)((n_cell - r0) < kRowBlock ? (n_cell - r0) : kRowBlock);
        __syncthreads();
        // phase 1: one warp per row, mean and std over the control columns
        for (int r = warp; r < nr; r += kWarps) {
            const int64_t row = r0 + r;
            const double ref = row_ref[row];
            const float* p = dev_mat + row * ld;
            double s = 0.0;
            for (int j = 1 + lane; j < ncol; j += 32)
                s += aligned_value(ref, p[j], sh_cm[j], sh_a[j]);
            s = warp_sum_d(s);
            const double mu = n_ctrl > 0 ? s / (double)n_ctrl : 0.0;
            double q = 0.0;
            for (int j = 1 + lane; j < ncol; j += 32) {
                const double z = aligned_value(ref, p[j], sh_cm[j], sh_a[j]) - mu;
                q = fma(z, z, q);
            }
            q = warp_sum_d(q);
            const double sd = n_ctrl > 0 ? sqrt(q / (double)n_ctrl) : 0.0;
            if (lane == 0) {
                const double is = 1.0 / sd;
                sh_ref[r] = ref;
                sh_mu[r] = mu;
                sh_is[r] = is;
                row_mean[row] = mu;
                row_istd[row] = is;
            }
        }
        __syncthreads();
        // phase 2: one thread per column, sums and minima of the standardised scores
        for (int j = threadIdx.x; j < ncol; j += kThreads) {
            const double cm = sh_cm[j], a = sh_a[j];
            const float* p = dev_mat + r0 * ld + j;
            double s = 0.0, mn = INFINITY;
#pragma unroll 8
            for (int r = 0; r < nr; ++r) {
                const double z = aligned_value(sh_ref[r], p[(int64_t)r * ld], cm, a);
                const double v = __dmul_rn(z - sh_mu[r], sh_is[r]);
                s += v;
                mn = min_nan(mn, v);
            }
            sh_sum[j] += s;
            sh_min[j] = min_nan(sh_min[j], mn);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ncol; j += kThreads) {
        part_sum[(size_t)blockIdx.x * ncol + j] = sh_sum[j];
        part_min[(size_t)blockIdx.x * ncol + j] = sh_min[j];
    }
}


// Same statistics for score matrices of at most 1 + 1024 columns (one pass of the fixed-point scatter), with ONE read of
// the matrix: a CTA of 8 warps owns blocks of 16 rows (two CTAs per SM: one loads while the other sums columns); a warp loads a whole row (32 independent 128-byte loads per
// lane in flight), keeps its aligned control scores in registers for the mean and the centred second moment, and
// leaves the raw deviations in a shared-memory tile from which the per-column pass (a thread per column, rows in
// order) reads them again -- rowstats_kernel above reads every value three times (once from DRAM, twice from L2).
constexpr int kTileThreads = 256, kTileWarps = kTileThreads / 32, kTileMaxCtrl = 1024, kTileRows = 16, kTilePerSM = 2;

__global__ void __launch_bounds__(kTileThreads, kTilePerSM)
rowstats_tile_kernel(int64_t n_cell, int ncol, int ld, const float* __restrict__ dev_mat,
                     const double* __restrict__ row_ref, const double* __restrict__ colmean,
                     const double* __restrict__ ratio_a, double* __restrict__ row_mean,
                     double* __restrict__ row_istd, double* __restrict__ part_sum,
                     double* __restrict__ part_min) {
    extern __shared__ double sh[];
    double* sh_cm = sh;             // [ncol] column means of raw
    double* sh_a = sh + ncol;       // [ncol] 1/sqrt(ratio)
    double* sh_sum = sh + 2 * ncol; // [ncol] partial sums of n
    double* sh_min = sh + 3 * ncol; // [ncol] partial minima of n
    float* tile = reinterpret_cast<float*>(sh + 4 * ncol);      // [kTileRows][ncol] raw deviations
    __shared__ double sh_ref[kTileRows], sh_mu[kTileRows], sh_is[kTileRows];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < ncol; j += kTileThreads) {
        sh_cm[j] = colmean[j];
        sh_a[j] = ratio_a[j];
        sh_sum[j] = 0.0;
        sh_min[j] = INFINITY;
    }
    const int n_ctrl = ncol - 1;
    const int64_t nblk = (n_cell + kTileRows - 1) / kTileRows;
    for (int64_t b = blockIdx.x; b < nblk; b += gridDim.x) {
        const int64_t r0 = b * kTileRows;
        const int nr = (int)((n_cell - r0) < kTileRows ? (n_cell - r0) : kTileRows);
        __syncthreads();
        // phase 1: one warp per row; lane l owns the control columns 1 + l + 32 t
        for (int r = warp; r < nr; r += kTileWarps) {
            const int64_t row = r0 + r;
            const double ref = row_ref[row];
            const float* p = dev_mat + row * ld;
            float* trow = tile + (size_t)r * ncol;
            float d[kTileMaxCtrl / 32];
#pragma unroll
            for (int t = 0; t < kTileMaxCtrl / 32; ++t) {
                const int j = 1 + lane + 32 * t;
                d[t] = j < ncol ? __ldg(p + j) : 0.f;
            }
            if (lane == 0) trow[0] = __ldg(p);
            double z[kTileMaxCtrl / 32];
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < kTileMaxCtrl / 32; ++t) {
                const int j = 1 + lane + 32 * t;
                if (j < ncol) {
                    trow[j] = d[t];
                    z[t] = aligned_value(ref, d[t], sh_cm[j], sh_a[j]);
                    s += z[t];
                } else {
                    z[t] = 0.0;
                }
            }
            s = warp_sum_d(s);
            const double mu = n_ctrl > 0 ? s / (double)n_ctrl : 0.0;
            double q = 0.0;
#pragma unroll
            for (int t = 0; t < kTileMaxCtrl / 32; ++t) {
                const int j = 1 + lane + 32 * t;
                if (j < ncol) {
                    const double c = z[t] - mu;
                    q = fma(c, c, q);
                }
            }
            q = warp_sum_d(q);
            const double sd = n_ctrl > 0 ? sqrt(q / (double)n_ctrl) : 0.0;
            if (lane == 0) {
                const double is = 1.0 / sd;
                sh_ref[r] = ref;
                sh_mu[r] = mu;
                sh_is[r] = is;
                row_mean[row] = mu;
                row_istd[row] = is;
            }
        }
        __syncthreads();
        // phase 2: one thread per column, sums and minima of the standardised scores, rows in order
        for (int j = threadIdx.x; j < ncol; j += kTileThreads) {
            const double cm = sh_cm[j], a = sh_a[j];
            double s = 0.0, mn = INFINITY;
#pragma unroll 8
            for (int r = 0; r < nr; ++r) {
                const double zz = aligned_value(sh_ref[r], tile[(size_t)r * ncol + j], cm, a);
                const double v = __dmul_rn(zz - sh_mu[r], sh_is[r]);
                s += v;
                mn = min_nan(mn, v);
            }
            sh_sum[j] += s;
            sh_min[j] = min_nan(sh_min[j], mn);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ncol; j += kTileThreads) {
        part_sum[(size_t)blockIdx.x * ncol + j] = sh_sum[j];
        part_min[(size_t)blockIdx.x * ncol + j] = sh_min[j];
    }
}

int bg_rowstats(scdrs_ctx* c, const double* dev_colsum, int64_t n_cell_total, double* dev_col2,
                double* dev_colmin) {
    SCDRS_CHECK(c->trait_ready, "no raw scores resident (call scdrs_trait_raw first)");
    SCDRS_CHECK(n_cell_total >= c->n_cell, "n_cell_total smaller than the local cell count");
    const int ncol = c->ncol;
    c->n_cell_total = n_cell_total;
    const bool tiled = ncol <= 1 + kTileMaxCtrl;
    const int64_t nblk_rows = (c->n_cell + kTileRows - 1) / kTileRows, g_tile = (int64_t)kNumSM * kTilePerSM;
    const int g = tiled ? (int)(nblk_rows < g_tile ? (nblk_rows < 1 ? 1 : nblk_rows) : g_tile) : grid_for(c->n_cell);
    SCDRS_TRY(c->partial.reserve((size_t)g * ncol * sizeof(double) * 2));
    SCDRS_TRY(c->colmean.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->row_mean.reserve((size_t)c->n_cell * sizeof(double)));
    SCDRS_TRY(c->row_istd.reserve((size_t)c->n_cell * sizeof(double)));
    colmean_kernel<<<(ncol + 127) / 128, 128, 0, c->stream>>>(ncol, 1.0 / (double)n_cell_total,
                                                              dev_colsum, c->colmean.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    double* psum = c->partial.as<double>();
    double* pmin = psum + (size_t)g * ncol;
    static bool attr = false;
    if (!attr) {
        SCDRS_CUDA(cudaFuncSetAttribute(rowstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SCDRS_CUDA(cudaFuncSetAttribute(rowstats_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    if (tiled) {
        const size_t smem = (size_t)ncol * 4 * sizeof(double) + (size_t)kTileRows * ncol * sizeof(float);
        rowstats_tile_kernel<<<g, kTileThreads, smem, c->stream>>>(
            c->n_cell, ncol, c->ld, c->dev_mat.as<float>(), c->row_ref.as<double>(),
            c->colmean.as<double>(), c->ratio_a.as<double>(), c->row_mean.as<double>(),
            c->row_istd.as<double>(), psum, pmin);
    } else {
        const size_t smem = (size_t)ncol * 4 * sizeof(double);
        SCDRS_CHECK(smem <= 200 * 1024, "1 + n_ctrl too large for the row-statistics kernel");
        rowstats_kernel<<<g, kThreads, smem, c->stream>>>(
            c->n_cell, ncol, c->ld, c->dev_mat.as<float>(), c->row_ref.as<double>(),
            c->colmean.as<double>(), c->ratio_a.as<double>(), c->row_mean.as<double>(),
            c->row_istd.as<double>(), psum, pmin);
    }
    SCDRS_LAUNCH_CHECK(c);
    reduce_partials_kernel<<<(ncol + 31) / 32, kRedWarps * 32, 0, c->stream>>>(g, ncol, 0, psum, dev_col2);
    SCDRS_LAUNCH_CHECK(c);
    reduce_partials_kernel<<<(ncol + 31) / 32, kRedWarps * 32, 0, c->stream>>>(g, ncol, 1, pmin, dev_colmin);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

// ---- second alignment constants + trait column ------------------------------------------------
// col2[j] <- column mean of n ; scalars[0] = global minimum of the final scores (:581)
__global__ void finalize_cols_kernel(int ncol, double inv_n, const double* __restrict__ colsum2,
                                     const double* __restrict__ colmin, double* __restrict__ col2,
                                     double* __restrict__ scalars) {
    __shared__ double sh[32];
    double mn = INFINITY;
    for (int j = threadIdx.x; j < ncol; j += blockDim.x) {
        const double m2 = colsum2[j] * inv_n;
        col2[j] = m2;
        mn = min_nan(mn, colmin[j] - m2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mn = min_nan(mn, sh[i]);
        scalars[0] = mn;
    }
}

__global__ void queries_kernel(int64_t n_cell, int ld, const float* __restrict__ dev_mat,
                               const double* __restrict__ row_ref, const double* __restrict__ colmean,
                               const double* __restrict__ ratio_a, const double* __restrict__ row_mean,
                               const double* __restrict__ row_istd, const double* __restrict__ col2,
                               const double* __restrict__ scalars, float* __restrict__ query) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_cell) return;
    const float d = dev_mat[r * ld];
    double v;
    if (isnan(d)) {
        v = scalars[0] - 1e-3;
    } else {
        const double z = aligned_value(row_ref[r], d, colmean[0], ratio_a[0]);
        v = __dmul_rn(z - row_mean[r], row_istd[r]) - col2[0];
    }
    query[r] = (float)v;
}

int bg_queries(scdrs_ctx* c, const double* dev_col2, const double* dev_colmin, float* dev_query) {
    SCDRS_CHECK(c->n_cell_total > 0, "row statistics missing (call scdrs_trait_rowstats first)");
    SCDRS_TRY(c->col2.reserve(c->ncol * sizeof(double)));
    SCDRS_TRY(c->scalars.reserve(8 * sizeof(double)));
    finalize_cols_kernel<<<1, 256, 0, c->stream>>>(c->ncol, 1.0 / (double)c->n_cell_total, dev_col2,
                                                   dev_colmin, c->col2.as<double>(),
                                                   c->scalars.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    queries_kernel<<<(unsigned)((c->n_cell + 255) / 256), 256, 0, c->stream>>>(
        c->n_cell, c->ld, c->dev_mat.as<float>(), c->row_ref.as<double>(), c->colmean.as<double>(),
        c->ratio_a.as<double>(), c->row_mean.as<double>(), c->row_istd.as<double>(),
        c->col2.as<double>(), c->scalars.as<double>(), dev_query);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

// ---- stand-alone _correct_background: pack an explicit (v_raw, mat_ctrl) pair -------------------
__global__ void pack_dense_kernel(int64_t n_cell, int ncol, int ld, const double* __restrict__ vraw,
                                  const double* __restrict__ mat, float* __restrict__ dev_mat,
                                  double* __restrict__ row_ref) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_cell) return;
    const int n_ctrl = ncol - 1;
    double s = 0.0;
    for (int j = lane; j < ncol; j += 32)
        s += j == 0 ? vraw[row] : mat[row * n_ctrl + (j - 1)];
    s = warp_sum_d(s);
    const double ref = s / (double)ncol;
    for (int j = lane; j < ncol; j += 32) {
        const double v = j == 0 ? vraw[row] : mat[row * n_ctrl + (j - 1)];
        dev_mat[row * ld + j] = v == 0.0 ? __int_as_float(0x7fc00000) : (float)(v - ref);
    }
    if (lane == 0) row_ref[row] = ref;
}

__global__ void ratio_to_a_kernel(int ncol, const double* __restrict__ ratio, double* __restrict__ a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ncol) a[j] = j == 0 ? 1.0 : 1.0 / sqrt(ratio[j - 1]);
}

int bg_pack_dense(scdrs_ctx* c, int64_t n_cell, int32_t n_ctrl, const double* dev_vraw,
                  const double* dev_mat, const double* dev_ratio) {
    c->ncol = n_ctrl + 1;
    c->n_ctrl = n_ctrl;
    c->ld = (c->ncol + 3) & ~3;
    SCDRS_TRY(c->dev_mat.reserve((size_t)n_cell * c->ld * sizeof(float)));
    SCDRS_TRY(c->row_ref.reserve((size_t)n_cell * sizeof(double)));
    SCDRS_TRY(c->ratio_a.reserve(c->ncol * sizeof(double)));
    pack_dense_kernel<<<(unsigned)((n_cell * 32 + 255) / 256), 256, 0, c->stream>>>(
        n_cell, c->ncol, c->ld, dev_vraw, dev_mat, c->dev_mat.as<float>(), c->row_ref.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    ratio_to_a_kernel<<<(c->ncol + 127) / 128, 128, 0, c->stream>>>(c->ncol, dev_ratio,
                                                                    c->ratio_a.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

// raw_score[c] (f64) and optionally ctrl_raw[c, j-1] (f32) from the packed representation
__global__ void export_raw_kernel(int64_t n_cell, int ncol, int ld, const float* __restrict__ dev_mat,
                                  const double* __restrict__ row_ref, double* __restrict__ raw_score,
                                  float* __restrict__ ctrl_raw) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_cell) return;
    const double ref = row_ref[row];
    const float* p = dev_mat + row * ld;
    if (lane == 0) raw_score[row] = raw_value(ref, p[0]);
    if (ctrl_raw != nullptr) {
        const int n_ctrl = ncol - 1;
        for (int j = 1 + lane; j < ncol; j += 32)
            ctrl_raw[row * n_ctrl + (j - 1)] = (float)raw_value(ref, p[j]);
    }
}

int bg_export_raw(scdrs_ctx* c, double* dev_raw_score, float* dev_ctrl_raw) {
    export_raw_kernel<<<(unsigned)((c->n_cell * 32 + 255) / 256), 256, 0, c->stream>>>(
        c->n_cell, c->ncol, c->ld, c->dev_mat.as<float>(), c->row_ref.as<double>(), dev_raw_score,
        dev_ctrl_raw);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

}  // namespace scdrs
