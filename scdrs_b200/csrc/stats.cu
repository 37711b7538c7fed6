// K4: the reductions behind pp.compute_stats of the reference (scdrs/pp.py:353-373, 409-417,
// _get_mean_var :467-517, _weighted_sparse_average :520-547, _get_mean_var_implicit_cov_corr
// :550-641): per-gene and per-cell sums of T(v) and T(v)^2 with
//     v[c, g] = X[c, g] + COV_MAT[c, :] . COV_BETA[g, :] + COV_GENE_MEAN[g]   (implicit mode)
//     v[c, g] = X[c, g]                                                       (normal mode)
// and T = identity or expm1.  The reference densifies the matrix in n_chunk float64 blocks.  Here
// per-gene sums are accumulated cell by cell in the reference's own order (see below); per-cell
// sums split into a dense part (what every entry would be if X were all zero) and a sparse part
// (the change each stored entry makes):  sum_g T(v) = sum_g T(b) + sum_{nnz} (T(b + x) - T(b)).
// All accumulation is fp64; nothing uses atomics, so results are bitwise reproducible.
#include "common.cuh"

namespace scdrs {

template <int TRANSFORM>
__device__ __forceinline__ double apply_t(double v) {
    return TRANSFORM == 0 ? v : expm1(v);
}

__device__ __forceinline__ double warp_sum_s(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- per-gene statistics: strictly sequential over cells, like the reference ---------------------
// The reference reduces a dense (n_cell x chunk) float64 block along axis 0, which numpy does by
// adding row after row; np.var subtracts the mean first (scdrs/pp.py:613-621, 498-501).  For genes
// whose corrected expression is (numerically) constant the resulting variance is pure rounding noise
// of that exact operation order, and it feeds the LOESS trend -- so the order is reproduced:
// one thread owns one gene and walks the cells in order; a CTA owns kGeneTile genes and stages
// kRowTile x kGeneTile dense tiles of X in shared memory (rows are located by binary search in the
// sorted CSR column ids).  v = (x + sum_k C[c,k] B[g,k]) + mu with the k-sum accumulated in index
// order by fma, as a BLAS dgemm micro-kernel does.  Bitwise reproducible.
constexpr int kGeneTile = 64;    // 313 tiles at 20k genes: ~2 resident CTAs per SM, no tail wave
constexpr int kRowTile = 64;

template <int TRANSFORM>
__global__ void __launch_bounds__(kGeneTile)
stats_gene_seq_kernel(int64_t n_cell, int n_gene, int k, int center, double denom,
                      const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                      const float* __restrict__ data, const double* __restrict__ cov_mat,
                      const double* __restrict__ cov_beta, const double* __restrict__ gene_mean,
                      const double* __restrict__ cell_w, double* __restrict__ gsum,
                      double* __restrict__ gsq) {
    extern __shared__ double sdyn[];                 // [k][kGeneTile] covariate effects of this tile
    __shared__ float tile[kRowTile][kGeneTile + 1];
    const int g0 = blockIdx.x * kGeneTile;
    const int g = g0 + threadIdx.x;
    const bool ok = g < n_gene;
    for (int kk = 0; kk < k; ++kk) sdyn[kk * kGeneTile + threadIdx.x] = ok ? cov_beta[(size_t)g * k + kk] : 0.0;
    const double mu = (ok && k > 0) ? gene_mean[g] : 0.0;
    const int n_pass = center ? 2 : 1;
    double s = 0.0, q = 0.0, mean = 0.0;
    for (int pass = 0; pass < n_pass; ++pass) {
        for (int64_t r0 = 0; r0 < n_cell; r0 += kRowTile) {
            const int nr = (int)((n_cell - r0) < kRowTile ? (n_cell - r0) : kRowTile);
            __syncthreads();
            for (int i = threadIdx.x; i < kRowTile * (kGeneTile + 1); i += kGeneTile) (&tile[0][0])[i] = 0.f;
            __syncthreads();
            if ((int)threadIdx.x < nr) {
                const int64_t p0 = indptr[r0 + threadIdx.x], p1 = indptr[r0 + threadIdx.x + 1];
                int64_t lo = p0, hi = p1;
                while (lo < hi) {                     // first stored column >= g0
                    const int64_t mid = (lo + hi) >> 1;
                    if (indices[mid] < g0) lo = mid + 1; else hi = mid;
                }
                for (int64_t t = lo; t < p1; ++t) {
                    const int col = indices[t] - g0;
                    if (col >= kGeneTile) break;
                    tile[threadIdx.x][col] = data[t];
                }
            }
            __syncthreads();
            if (ok) {
                // four cells at a time: the transforms are independent (instruction-level parallelism),
                // the accumulation stays strictly in cell order
                for (int rb = 0; rb < nr; rb += 4) {
                    double t[4], w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = rb + u < nr ? rb + u : nr - 1;
                        double v = (double)tile[r][threadIdx.x];
                        if (k > 0) {
                            const double* crow = cov_mat + (r0 + r) * k;
                            double cb = crow[0] * sdyn[threadIdx.x];
                            for (int kk = 1; kk < k; ++kk) cb = fma(crow[kk], sdyn[kk * kGeneTile + threadIdx.x], cb);
                            v = (v + cb) + mu;
                        }
                        t[u] = apply_t<TRANSFORM>(v);
                        w[u] = cell_w != nullptr ? cell_w[r0 + r] : 1.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (rb + u >= nr) break;
                        // explicit _rn intrinsics: no fma contraction, numpy multiplies then adds
                        if (pass == 0) {
                            s = __dadd_rn(s, cell_w != nullptr ? __dmul_rn(t[u], w[u]) : t[u]);
                            if (!center) {
                                const double t2 = __dmul_rn(t[u], t[u]);
                                q = __dadd_rn(q, cell_w != nullptr ? __dmul_rn(t2, w[u]) : t2);
                            }
                        } else {
                            const double d = __dadd_rn(t[u], -mean);
                            const double d2 = __dmul_rn(d, d);
                            q = __dadd_rn(q, cell_w != nullptr ? __dmul_rn(d2, w[u]) : d2);
                        }
                    }
                }
            }
        }
        if (pass == 0) mean = s / denom;
    }
    if (ok) {
        gsum[g] = s;
        gsq[g] = q;
    }
}

// ---- gene means and C^T X for the covariate regression (scdrs/pp.py:206, :209-212) --------------
// Same tiling as above.  The reference's `X.mean(axis=0)` on float32 CSR data multiplies every stored
// value by float32(1/n) and accumulates in float32 in cell order (scipy csc_matvec); that sum is
// reproduced operation by operation so COV_GENE_MEAN carries the same bits.  C^T X is accumulated
// in float64 in cell order (multiply, then add), as scipy's csc_matvecs does.
constexpr int kMaxCovReg = 16;

__global__ void __launch_bounds__(kGeneTile)
gene_mean_xtc_kernel(int64_t n_cell, int n_gene, int k, float inv_n,
                     const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                     const float* __restrict__ data, const double* __restrict__ cov_mat,
                     float* __restrict__ gmean32, double* __restrict__ xtc) {
    __shared__ float tile[kRowTile][kGeneTile + 1];
    const int g0 = blockIdx.x * kGeneTile;
    const int g = g0 + threadIdx.x;
    const bool ok = g < n_gene;
    float s32 = 0.f;
    double acc[kMaxCovReg];
#pragma unroll
    for (int kk = 0; kk < kMaxCovReg; ++kk) acc[kk] = 0.0;
    for (int64_t r0 = 0; r0 < n_cell; r0 += kRowTile) {
        const int nr = (int)((n_cell - r0) < kRowTile ? (n_cell - r0) : kRowTile);
        __syncthreads();
        for (int i = threadIdx.x; i < kRowTile * (kGeneTile + 1); i += kGeneTile) (&tile[0][0])[i] = 0.f;
        __syncthreads();
        if ((int)threadIdx.x < nr) {
            const int64_t p0 = indptr[r0 + threadIdx.x], p1 = indptr[r0 + threadIdx.x + 1];
            int64_t lo = p0, hi = p1;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (indices[mid] < g0) lo = mid + 1; else hi = mid;
            }
            for (int64_t t = lo; t < p1; ++t) {
                const int col = indices[t] - g0;
                if (col >= kGeneTile) break;
                tile[threadIdx.x][col] = data[t];
            }
        }
        __syncthreads();
        if (ok) {
            for (int r = 0; r < nr; ++r) {
                const float x = tile[r][threadIdx.x];
                if (x != 0.f) {      // only stored entries take part, as in the sparse products
                    s32 = __fadd_rn(s32, __fmul_rn(x, inv_n));
                    const double xd = (double)x;
                    const double* crow = cov_mat + (r0 + r) * k;
#pragma unroll
                    for (int kk = 0; kk < kMaxCovReg; ++kk)
                        if (kk < k) acc[kk] = __dadd_rn(acc[kk], __dmul_rn(xd, crow[kk]));
                }
            }
        }
    }
    if (ok) {
        gmean32[g] = s32;
#pragma unroll
        for (int kk = 0; kk < kMaxCovReg; ++kk)
            if (kk < k) xtc[(size_t)g * k + kk] = acc[kk];
    }
}

int gene_mean_xtc(scdrs_ctx* c, int k, const double* dev_cov, float* dev_gmean32, double* dev_xtc) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    SCDRS_CHECK(k >= 0 && k <= kMaxCovReg, "this kernel handles at most %d covariates (got %d)", kMaxCovReg, k);
    const int tiles = (c->n_gene + kGeneTile - 1) / kGeneTile;
    gene_mean_xtc_kernel<<<tiles, kGeneTile, 0, c->stream>>>(
        c->n_cell, c->n_gene, k, (float)(1.0 / (double)c->n_cell), c->indptr, c->indices, c->data, dev_cov,
        dev_gmean32, dev_xtc);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

// ---- sparse part of the per-cell statistics: one warp per cell ----------------------------------
template <int TRANSFORM>
__global__ void __launch_bounds__(256)
stats_sparse_kernel(int64_t n_cell, int k, const int64_t* __restrict__ indptr,
                    const int32_t* __restrict__ indices, const float* __restrict__ data,
                    const double* __restrict__ cov_mat, const double* __restrict__ cov_beta,
                    const double* __restrict__ gene_mean, double* __restrict__ csum,
                    double* __restrict__ csq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < n_cell; row += nwarp) {
        const int64_t p0 = indptr[row], p1 = indptr[row + 1];
        double cs = 0.0, cq = 0.0;
        for (int64_t q = p0 + lane; q < p1; q += 32) {
            const int g = indices[q];
            const double x = (double)data[q];
            double b = 0.0;
            if (k > 0) {
                b = gene_mean[g];
                for (int kk = 0; kk < k; ++kk)
                    b = fma(cov_mat[row * k + kk], cov_beta[(size_t)g * k + kk], b);
            }
            const double t1 = apply_t<TRANSFORM>(b + x);
            const double t0 = k > 0 ? apply_t<TRANSFORM>(b) : 0.0;
            cs += t1 - t0;
            cq += t1 * t1 - t0 * t0;
        }
        cs = warp_sum_s(cs);
        cq = warp_sum_s(cq);
        if (lane == 0) {
            csum[row] += cs;
            csq[row] += cq;
        }
    }
}

// ---- dense part, per cell: one warp per cell, lanes over genes ------------------------------------
template <int TRANSFORM>
__global__ void __launch_bounds__(256)
stats_dense_cell_kernel(int64_t n_cell, int n_gene, int k, const double* __restrict__ cov_mat,
                        const double* __restrict__ cov_beta, const double* __restrict__ gene_mean,
                        double* __restrict__ csum, double* __restrict__ csq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < n_cell; row += nwarp) {
        double s = 0.0, q = 0.0;
        for (int g = lane; g < n_gene; g += 32) {
            double b = gene_mean[g];
            for (int kk = 0; kk < k; ++kk)
                b = fma(cov_mat[row * k + kk], cov_beta[(size_t)g * k + kk], b);
            const double t = apply_t<TRANSFORM>(b);
            s += t;
            q = fma(t, t, q);
        }
        s = warp_sum_s(s);
        q = warp_sum_s(q);
        if (lane == 0) {
            csum[row] = s;
            csq[row] = q;
        }
    }
}

template <int TRANSFORM>
static int run_stats(scdrs_ctx* c, const double* dev_cell_w, int center, double denom, double* gsum,
                     double* gsq, double* csum, double* csq) {
    const int k = c->k, n_gene = c->n_gene;
    const int64_t n_cell = c->n_cell;
    if (gsum != nullptr) {
        const int tiles = (n_gene + kGeneTile - 1) / kGeneTile;
        const size_t smem = (size_t)(k > 0 ? k : 1) * kGeneTile * sizeof(double);
        SCDRS_CHECK(smem <= 150 * 1024, "too many covariates (%d) for the statistics kernel", k);
        SCDRS_CUDA(cudaFuncSetAttribute(stats_gene_seq_kernel<TRANSFORM>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024));
        stats_gene_seq_kernel<TRANSFORM><<<tiles, kGeneTile, smem, c->stream>>>(
            n_cell, n_gene, k, center, denom, c->indptr, c->indices, c->data, c->cov_mat.as<double>(),
            c->cov_beta.as<double>(), c->gene_mean.as<double>(), dev_cell_w, gsum, gsq);
        SCDRS_LAUNCH_CHECK(c);
    }
    if (csum != nullptr) {
        int64_t grid = (n_cell + 7) / 8;
        if (grid > kNumSM * 8) grid = kNumSM * 8;
        if (grid < 1) grid = 1;
        if (k > 0) {
            stats_dense_cell_kernel<TRANSFORM><<<(unsigned)grid, 256, 0, c->stream>>>(
                n_cell, n_gene, k, c->cov_mat.as<double>(), c->cov_beta.as<double>(),
                c->gene_mean.as<double>(), csum, csq);
            SCDRS_LAUNCH_CHECK(c);
        } else {
            SCDRS_CUDA(cudaMemsetAsync(csum, 0, (size_t)n_cell * sizeof(double), c->stream));
            SCDRS_CUDA(cudaMemsetAsync(csq, 0, (size_t)n_cell * sizeof(double), c->stream));
        }
        stats_sparse_kernel<TRANSFORM><<<(unsigned)grid, 256, 0, c->stream>>>(
            n_cell, k, c->indptr, c->indices, c->data, c->cov_mat.as<double>(), c->cov_beta.as<double>(),
            c->gene_mean.as<double>(), csum, csq);
        SCDRS_LAUNCH_CHECK(c);
    }
    return 0;
}

int gene_cell_stats(scdrs_ctx* c, int transform, const double* dev_cell_w, int center, double denom,
                    double* dev_gsum, double* dev_gsumsq, double* dev_csum, double* dev_csumsq) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    SCDRS_CHECK(transform == 0 || transform == 1, "transform must be 0 (identity) or 1 (expm1)");
    SCDRS_CHECK((dev_gsum == nullptr) == (dev_gsumsq == nullptr), "gene outputs must come in pairs");
    SCDRS_CHECK((dev_csum == nullptr) == (dev_csumsq == nullptr), "cell outputs must come in pairs");
    SCDRS_CHECK(!center || denom > 0, "centered sums need a positive denominator");
    if (transform == 0) return run_stats<0>(c, dev_cell_w, center, denom, dev_gsum, dev_gsumsq, dev_csum, dev_csumsq);
    return run_stats<1>(c, dev_cell_w, center, denom, dev_gsum, dev_gsumsq, dev_csum, dev_csumsq);
}

}  // namespace scdrs
