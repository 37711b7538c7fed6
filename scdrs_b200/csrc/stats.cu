// K4: the reductions behind pp.compute_stats of the reference (scdrs/pp.py:353-373, 409-417,
// _get_mean_var :467-517, _weighted_sparse_average :520-547, _get_mean_var_implicit_cov_corr
// :550-641): per-gene and per-cell sums of T(v) and T(v)^2 with
//     v[c, g] = X[c, g] + COV_MAT[c, :] . COV_BETA[g, :] + COV_GENE_MEAN[g]   (implicit mode)
//     v[c, g] = X[c, g]                                                       (normal mode)
// and T = identity or expm1.  The reference densifies the matrix in n_chunk float64 blocks; here
// the dense part (what every entry would be if X were all zero) and the sparse part (the change
// each stored entry makes) are reduced separately and X is read exactly once:
//     sum_c w_c T(v) = sum_c w_c T(b[c, g])  +  sum_{nnz} w_c (T(b + x) - T(b)),  b = v - x.
// In normal mode b = 0 and T(0) = 0, so only the sparse pass runs.  All accumulation is fp64.
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

// ---- sparse part: one warp per cell ------------------------------------------------------------
template <int TRANSFORM>
__global__ void __launch_bounds__(256)
stats_sparse_kernel(int64_t n_cell, int k, const int64_t* __restrict__ indptr,
                    const int32_t* __restrict__ indices, const float* __restrict__ data,
                    const double* __restrict__ cov_mat, const double* __restrict__ cov_beta,
                    const double* __restrict__ gene_mean, const double* __restrict__ cell_w,
                    double* __restrict__ gsum, double* __restrict__ gsq,
                    double* __restrict__ csum, double* __restrict__ csq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < n_cell; row += nwarp) {
        const double w = cell_w != nullptr ? cell_w[row] : 1.0;
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
            const double dv = t1 - t0, dq = t1 * t1 - t0 * t0;
            if (gsum != nullptr) {
                atomicAdd(&gsum[g], w * dv);
                atomicAdd(&gsq[g], w * dq);
            }
            cs += dv;
            cq += dq;
        }
        if (csum != nullptr) {
            cs = warp_sum_s(cs);
            cq = warp_sum_s(cq);
            if (lane == 0) {
                csum[row] += cs;
                csq[row] += cq;
            }
        }
    }
}

// ---- dense part, per gene: thread per gene, CTA = (gene tile, cell split) ------------------------
constexpr int kGeneTile = 128;

template <int TRANSFORM>
__global__ void __launch_bounds__(kGeneTile)
stats_dense_gene_kernel(int64_t n_cell, int n_gene, int k, int n_split,
                        const double* __restrict__ cov_mat, const double* __restrict__ cov_beta,
                        const double* __restrict__ gene_mean, const double* __restrict__ cell_w,
                        double* __restrict__ part) {  // [n_split][2][n_gene]
    extern __shared__ double sB[];  // [k][kGeneTile]
    const int g = blockIdx.x * kGeneTile + threadIdx.x;
    const int split = blockIdx.y;
    const bool ok = g < n_gene;
    for (int kk = 0; kk < k; ++kk) sB[kk * kGeneTile + threadIdx.x] = ok ? cov_beta[(size_t)g * k + kk] : 0.0;
    const double mu = ok ? gene_mean[g] : 0.0;
    const int64_t per = (n_cell + n_split - 1) / n_split;
    const int64_t r0 = split * per, r1 = (r0 + per) < n_cell ? (r0 + per) : n_cell;
    double s = 0.0, q = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        double b = mu;
        for (int kk = 0; kk < k; ++kk) b = fma(cov_mat[r * k + kk], sB[kk * kGeneTile + threadIdx.x], b);
        const double t = apply_t<TRANSFORM>(b);
        const double w = cell_w != nullptr ? cell_w[r] : 1.0;
        s = fma(w, t, s);
        q = fma(w * t, t, q);
    }
    if (ok) {
        part[((size_t)split * 2 + 0) * n_gene + g] = s;
        part[((size_t)split * 2 + 1) * n_gene + g] = q;
    }
}

__global__ void stats_gene_reduce_kernel(int n_gene, int n_split, const double* __restrict__ part,
                                         double* __restrict__ gsum, double* __restrict__ gsq) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gene) return;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < n_split; ++i) {
        s += part[((size_t)i * 2 + 0) * n_gene + g];
        q += part[((size_t)i * 2 + 1) * n_gene + g];
    }
    gsum[g] = s;
    gsq[g] = q;
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
static int run_stats(scdrs_ctx* c, const double* dev_cell_w, double* gsum, double* gsq,
                     double* csum, double* csq) {
    const int k = c->k, n_gene = c->n_gene;
    const int64_t n_cell = c->n_cell;
    const bool want_gene = gsum != nullptr, want_cell = csum != nullptr;
    if (want_gene) {
        if (k > 0) {
            const int tiles = (n_gene + kGeneTile - 1) / kGeneTile;
            int n_split = (kNumSM * 8 + tiles - 1) / tiles;
            if (n_split < 1) n_split = 1;
            if ((int64_t)n_split > n_cell) n_split = (int)(n_cell > 0 ? n_cell : 1);
            SCDRS_TRY(c->partial.reserve((size_t)n_split * 2 * n_gene * sizeof(double)));
            const size_t smem = (size_t)k * kGeneTile * sizeof(double);
            SCDRS_CHECK(smem <= 160 * 1024, "too many covariates (%d) for the statistics kernel", k);
            SCDRS_CUDA(cudaFuncSetAttribute(stats_dense_gene_kernel<TRANSFORM>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            stats_dense_gene_kernel<TRANSFORM><<<dim3(tiles, n_split), kGeneTile, smem, c->stream>>>(
                n_cell, n_gene, k, n_split, c->cov_mat.as<double>(), c->cov_beta.as<double>(),
                c->gene_mean.as<double>(), dev_cell_w, c->partial.as<double>());
            SCDRS_LAUNCH_CHECK(c);
            stats_gene_reduce_kernel<<<(n_gene + 127) / 128, 128, 0, c->stream>>>(
                n_gene, n_split, c->partial.as<double>(), gsum, gsq);
            SCDRS_LAUNCH_CHECK(c);
        } else {
            SCDRS_CUDA(cudaMemsetAsync(gsum, 0, (size_t)n_gene * sizeof(double), c->stream));
            SCDRS_CUDA(cudaMemsetAsync(gsq, 0, (size_t)n_gene * sizeof(double), c->stream));
        }
    }
    int64_t grid = (n_cell + 7) / 8;
    if (grid > kNumSM * 8) grid = kNumSM * 8;
    if (grid < 1) grid = 1;
    if (want_cell) {
        if (k > 0) {
            stats_dense_cell_kernel<TRANSFORM><<<(unsigned)grid, 256, 0, c->stream>>>(
                n_cell, n_gene, k, c->cov_mat.as<double>(), c->cov_beta.as<double>(),
                c->gene_mean.as<double>(), csum, csq);
            SCDRS_LAUNCH_CHECK(c);
        } else {
            SCDRS_CUDA(cudaMemsetAsync(csum, 0, (size_t)n_cell * sizeof(double), c->stream));
            SCDRS_CUDA(cudaMemsetAsync(csq, 0, (size_t)n_cell * sizeof(double), c->stream));
        }
    }
    stats_sparse_kernel<TRANSFORM><<<(unsigned)grid, 256, 0, c->stream>>>(
        n_cell, k, c->indptr, c->indices, c->data, c->cov_mat.as<double>(), c->cov_beta.as<double>(),
        c->gene_mean.as<double>(), dev_cell_w, gsum, gsq, csum, csq);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

int gene_cell_stats(scdrs_ctx* c, int transform, const double* dev_cell_w, double* dev_gsum,
                    double* dev_gsumsq, double* dev_csum, double* dev_csumsq) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    SCDRS_CHECK(transform == 0 || transform == 1, "transform must be 0 (identity) or 1 (expm1)");
    SCDRS_CHECK((dev_gsum == nullptr) == (dev_gsumsq == nullptr), "gene outputs must come in pairs");
    SCDRS_CHECK((dev_csum == nullptr) == (dev_csumsq == nullptr), "cell outputs must come in pairs");
    if (transform == 0) return run_stats<0>(c, dev_cell_w, dev_gsum, dev_gsumsq, dev_csum, dev_csumsq);
    return run_stats<1>(c, dev_cell_w, dev_gsum, dev_gsumsq, dev_csum, dev_csumsq);
}

}  // namespace scdrs
