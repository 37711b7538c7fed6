// K4, parallel form: per-gene moments of the resident CSR matrix for pp.preprocess / pp.compute_stats of the
// reference (scdrs/pp.py:206-212 gene means and C^T X; :353-373 mean / var and ct_mean / ct_var;
// _get_mean_var :467-517, _get_mean_var_implicit_cov_corr :550-641).
//
// The reference densifies v = X + COV_MAT . COV_BETA^T + COV_GENE_MEAN in chunks and lets numpy reduce them.
// Here nothing dense is formed for the identity transform: with b[c, g] = C[c, :] . B[g, :] + mu[g],
//     sum_c w_c v          = sum_nnz w x  +  (sum_c w_c C_c) . B_g  +  (sum_c w_c) mu_g
//     sum_c w_c (v - m)^2  = B_g' (sum_c w_c C_c C_c') B_g + 2 d (sum_c w_c C_c) . B_g + (sum_c w_c) d^2
//                            + 2 (sum_nnz w x C_c) . B_g + 2 d sum_nnz w x + sum_nnz w x^2,      d = mu_g - m
// so ONE pass over the stored entries that accumulates, per gene, sum w x, sum w x^2 and sum w x C[c, j] gives
// the gene means, C^T X of the covariate regression, and mean / var of the corrected matrix (the k x k terms
// are host arithmetic on n_gene x k numbers).  The expm1 transform has no such identity: a dense pass evaluates
// sum_c w (expm1(b) - K_g) and its square for every (cell, gene) pair (n_cell x n_gene fp64 expm1 on the FP64
// pipe), and a sparse pass adds what each stored entry changes, expm1(b + x) versus expm1(b); K_g is a per-gene
// shift near the mean, so that var = E[(T-K)^2] - E[T-K]^2 does not cancel.
//
// Per-gene sums from a cell-major matrix are a scatter.  A CTA owns (a block of rows) x (a tile of genes whose
// accumulators fit its shared memory); each warp owns a sub-range of the tile's genes: for every row it finds its
// segment of the (column-sorted) row by binary search -- 32 rows searched at once, a lane each -- and its lanes
// add consecutive entries (distinct genes, no conflict between lanes; nobody else touches the sub-range), rows in
// order.  No atomics: a gene's partial sum is the sum over the block's cells in cell order, partials of the row
// blocks are added in block order (gene_reduce_kernel), so results repeat bit for bit and do not depend on the
// grid; sums over the shards of several GPUs are added by the host side (all-reduce).  Genes whose variance is
// numerically zero against their mean are recomputed by the sequential kernel of stats.cu in numpy's own order
// (scdrs_gene_stats_exact), because there the variance IS rounding noise of that order.
#include "common.cuh"

namespace scdrs {

constexpr int kGaWarps = 16;
constexpr int kGaThreads = kGaWarps * 32;
constexpr size_t kGaSmemBudget = 200 * 1024;

// first entry of row [p0, p1) with column >= g
__device__ __forceinline__ int64_t row_lower_bound(const int32_t* __restrict__ indices, int64_t p0, int64_t p1, int g) {
    while (p0 < p1) {
        const int64_t mid = (p0 + p1) >> 1;
        if (__ldg(indices + mid) < g) p0 = mid + 1; else p1 = mid;
    }
    return p0;
}

// MODE 0: sum w x, sum w x^2, sum w x C[c, j]   (n_acc = 2 + k)
// MODE 1: sum w expm1(x), sum w expm1(x)^2      (n_acc = 2)
// MODE 2: sum w (expm1(b + x) - expm1(b)), sum w ((expm1(b + x) - K)^2 - (expm1(b) - K)^2)   (n_acc = 2)
template <int MODE>
__global__ void __launch_bounds__(kGaThreads, 1)
gene_accum_kernel(int64_t n_cell, int n_gene, int k, int n_acc, int gpt, int n_tile, int64_t rows_per_block,
                  const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const float* __restrict__ data, const double* __restrict__ cov_mat,
                  const double* __restrict__ cov_beta, const double* __restrict__ gene_mean,
                  const double* __restrict__ shift, const double* __restrict__ cell_w,
                  double* __restrict__ partial) {
    extern __shared__ double ga_acc[];          // [n_acc][gpt]
    const int tile = blockIdx.x % n_tile;
    const int64_t rb = blockIdx.x / n_tile;
    const int g0 = tile * gpt;
    const int g1 = (g0 + gpt < n_gene) ? g0 + gpt : n_gene;
    const int64_t r0 = rb * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < n_cell) ? r0 + rows_per_block : n_cell;
    for (int i = threadIdx.x; i < n_acc * gpt; i += kGaThreads) ga_acc[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = (g1 - g0 + kGaWarps - 1) / kGaWarps;
    const int my_lo = g0 + warp * sub;
    const int my_hi = (my_lo + sub < g1) ? my_lo + sub : g1;
    if (my_lo < my_hi) {
        for (int64_t base = r0; base < r1; base += 32) {
            const int64_t row_l = base + lane;
            int64_t s_l = 0, e_l = 0;
            if (row_l < r1) {
                const int64_t p0 = indptr[row_l], p1 = indptr[row_l + 1];
                s_l = row_lower_bound(indices, p0, p1, my_lo);
                e_l = row_lower_bound(indices, s_l, p1, my_hi);
            }
            const int nr = (int)((r1 - base) < 32 ? (r1 - base) : 32);
            // The first 32 entries of row l + 1's segment (usually all of it) are loaded while row l is accumulated:
            // a warp is one dependent chain load -> shared-memory read-modify-write per row, and with 16 warps per SM
            // the load latency would otherwise be most of the kernel.
            int64_t s = __shfl_sync(0xffffffffu, s_l, 0), e = __shfl_sync(0xffffffffu, e_l, 0);
            int g_cur = 0;
            float x_cur = 0.f;
            if (s + lane < e) { g_cur = __ldg(indices + s + lane); x_cur = __ldg(data + s + lane); }
            for (int l = 0; l < nr; ++l) {
                int64_t s_n = 0, e_n = 0;
                int g_nxt = 0;
                float x_nxt = 0.f;
                if (l + 1 < nr) {
                    s_n = __shfl_sync(0xffffffffu, s_l, l + 1);
                    e_n = __shfl_sync(0xffffffffu, e_l, l + 1);
                    if (s_n + lane < e_n) { g_nxt = __ldg(indices + s_n + lane); x_nxt = __ldg(data + s_n + lane); }
                }
                const int64_t row = base + l;
                const double w = cell_w != nullptr ? cell_w[row] : 1.0;
                // the row's first covariates, once per row (MODE 0); the accumulators of one entry are distinct words,
                // so they are all read before any is written: six independent read-modify-writes instead of a chain
                // the compiler has to keep in order (it cannot know that a[0] and a[gpt] differ)
                double cv[4] = {0.0, 0.0, 0.0, 0.0};
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < k) cv[j] = cov_mat[row * k + j];
                }
                for (int64_t q = s + lane; q < e; q += 32) {
                    const bool first = q < s + 32;
                    const int g = first ? g_cur : __ldg(indices + q);
                    const double x = (double)(first ? x_cur : __ldg(data + q));
                    double* a = ga_acc + (g - g0);
                    if (MODE == 0) {
                        const double wx = w * x;
                        const double v0 = a[0], v1 = a[gpt];
                        double t[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) t[j] = j < k ? a[(size_t)(2 + j) * gpt] : 0.0;
                        a[0] = v0 + wx;
                        a[gpt] = v1 + wx * x;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < k) a[(size_t)(2 + j) * gpt] = t[j] + wx * cv[j];
                        for (int j = 4; j < k; ++j) a[(size_t)(2 + j) * gpt] += wx * cov_mat[row * k + j];
                    } else if (MODE == 1) {
                        const double t = expm1(x);
                        const double v0 = a[0], v1 = a[gpt];
                        a[0] = v0 + w * t;
                        a[gpt] = v1 + w * t * t;
                    } else {
                        double b = gene_mean[g];
                        for (int j = 0; j < k; ++j) b = fma(cov_mat[row * k + j], cov_beta[(size_t)g * k + j], b);
                        const double kg = shift[g];
                        const double t1 = expm1(b + x) - kg, t0 = expm1(b) - kg;
                        const double v0 = a[0], v1 = a[gpt];
                        a[0] = v0 + w * (t1 - t0);
                        a[gpt] = v1 + w * (t1 * t1 - t0 * t0);
                    }
                }
                s = s_n;
                e = e_n;
                g_cur = g_nxt;
                x_cur = x_nxt;
            }
        }
    }
    __syncthreads();
    // partial[rb][a][g]
    double* out = partial + (size_t)rb * n_acc * n_gene;
    for (int i = threadIdx.x; i < n_acc * (g1 - g0); i += kGaThreads) {
        const int a = i / (g1 - g0), gi = i % (g1 - g0);
        out[(size_t)a * n_gene + g0 + gi] = ga_acc[(size_t)a * gpt + gi];
    }
}

// out[g * n_acc + a] (+)= sum over row blocks, in block order
__global__ void __launch_bounds__(256)
gene_reduce_kernel(int n_gene, int n_acc, int64_t n_rb, int accumulate, const double* __restrict__ partial,
                   double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_gene * n_acc) return;
    const int a = (int)(i / n_gene), g = (int)(i % n_gene);
    double s = 0.0;
    for (int64_t rb = 0; rb < n_rb; ++rb) s += partial[((size_t)rb * n_acc + a) * n_gene + g];
    double* dst = out + (size_t)g * n_acc + a;
    *dst = accumulate ? *dst + s : s;
}

// Dense part of the expm1 statistics: for every gene, over ALL cells of a cell block,
//   sum_c w_c (expm1(b[c, g]) - K_g)  and its square,  b = C[c, :] . B[g, :] + mu_g.
// One thread per gene (B_g, mu_g, K_g in registers / shared memory), covariate rows staged in shared memory.
constexpr int kEdGenes = 128;       // threads per block = genes per block
constexpr int kEdCells = 128;       // covariate rows staged per step
constexpr int kEdMaxK = 16;

__global__ void __launch_bounds__(kEdGenes)
gene_expm1_dense_kernel(int64_t n_cell, int n_gene, int k, int64_t cells_per_block,
                        const double* __restrict__ cov_mat, const double* __restrict__ cov_beta,
                        const double* __restrict__ gene_mean, const double* __restrict__ shift,
                        const double* __restrict__ cell_w, double* __restrict__ partial /* [n_cb][2][n_gene] */) {
    __shared__ double sc[kEdCells][kEdMaxK + 1];
    const int g = blockIdx.x * kEdGenes + threadIdx.x;
    const int64_t cb = blockIdx.y;
    const int64_t c0 = cb * cells_per_block;
    const int64_t c1 = (c0 + cells_per_block < n_cell) ? c0 + cells_per_block : n_cell;
    double beta[kEdMaxK];
#pragma unroll
    for (int j = 0; j < kEdMaxK; ++j) beta[j] = (j < k && g < n_gene) ? cov_beta[(size_t)g * k + j] : 0.0;
    const double mu = g < n_gene ? gene_mean[g] : 0.0, kg = g < n_gene ? shift[g] : 0.0;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t base = c0; base < c1; base += kEdCells) {
        const int nc = (int)((c1 - base) < kEdCells ? (c1 - base) : kEdCells);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * (k + 1); i += kEdGenes) {
            const int r = i / (k + 1), j = i % (k + 1);
            sc[r][j] = j < k ? cov_mat[(base + r) * k + j] : (cell_w != nullptr ? cell_w[base + r] : 1.0);
        }
        __syncthreads();
        for (int r = 0; r < nc; ++r) {
            double b = mu;
#pragma unroll
            for (int j = 0; j < kEdMaxK; ++j)
                if (j < k) b = fma(sc[r][j], beta[j], b);
            const double t = expm1(b) - kg;
            const double wt = sc[r][k] * t;
            s1 += wt;
            s2 = fma(wt, t, s2);
        }
    }
    if (g < n_gene) {
        partial[((size_t)cb * 2 + 0) * n_gene + g] = s1;
        partial[((size_t)cb * 2 + 1) * n_gene + g] = s2;
    }
}

static int launch_accum(scdrs_ctx* c, int mode, int k, int n_acc, const double* d_cov, const double* d_shift,
                        const double* d_w, double* d_out, int accumulate) {
    const int n_gene = c->n_gene;
    int gpt = (int)(kGaSmemBudget / (sizeof(double) * n_acc));
    if (gpt > n_gene) gpt = n_gene;
    const int n_tile = (n_gene + gpt - 1) / gpt;
    gpt = (n_gene + n_tile - 1) / n_tile;                        // equal tiles
    int64_t n_rb = (2 * kNumSM + n_tile - 1) / n_tile;           // about two CTAs per SM in total
    if (n_rb > (c->n_cell + 31) / 32) n_rb = (c->n_cell + 31) / 32;
    if (n_rb < 1) n_rb = 1;
    const int64_t rows_per_block = (c->n_cell + n_rb - 1) / n_rb;
    n_rb = (c->n_cell + rows_per_block - 1) / rows_per_block;
    const size_t smem = (size_t)gpt * n_acc * sizeof(double);
    SCDRS_TRY(c->partial.reserve((size_t)n_rb * n_acc * n_gene * sizeof(double)));
    double* part = c->partial.as<double>();
    const unsigned grid = (unsigned)(n_rb * n_tile);
#define GA_LAUNCH(M)                                                                                           \
    do {                                                                                                       \
        SCDRS_CUDA(cudaFuncSetAttribute(gene_accum_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                        (int)kGaSmemBudget));                                                  \
        gene_accum_kernel<M><<<grid, kGaThreads, smem, c->stream>>>(                                            \
            c->n_cell, n_gene, k, n_acc, gpt, n_tile, rows_per_block, c->indptr, c->indices, c->data, d_cov,   \
            c->cov_beta.as<double>(), c->gene_mean.as<double>(), d_shift, d_w, part);                          \
    } while (0)
    if (mode == 0) GA_LAUNCH(0);
    else if (mode == 1) GA_LAUNCH(1);
    else GA_LAUNCH(2);
#undef GA_LAUNCH
    SCDRS_LAUNCH_CHECK(c);
    const int64_t n_out = (int64_t)n_gene * n_acc;
    gene_reduce_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, c->stream>>>(n_gene, n_acc, n_rb, accumulate, part, d_out);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

// mode 0 / 1 of scdrs_gene_moments; d_cov [n_cell x k] (mode 0, k > 0), d_w [n_cell] or null; d_out [n_gene x n_acc]
int gene_moments(scdrs_ctx* c, int mode, int k, const double* d_cov, const double* d_w, double* d_out) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    SCDRS_CHECK(mode == 0 || mode == 1, "mode must be 0 (moments) or 1 (expm1 moments)");
    SCDRS_CHECK(k >= 0 && k <= 64, "number of covariates %d out of range", k);
    const int n_acc = mode == 0 ? 2 + k : 2;
    return launch_accum(c, mode, mode == 0 ? k : 0, n_acc, d_cov, nullptr, d_w, d_out, 0);
}

// sum_c w (expm1(v) - K) and its square over all cells, v = X + C B' + mu with the resident covariate terms
int gene_expm1_implicit(scdrs_ctx* c, const double* d_shift, const double* d_w, double* d_out /* [n_gene x 2] */) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    SCDRS_CHECK(c->k > 0 && c->k <= kEdMaxK, "needs 1 .. %d resident covariates (have %d)", kEdMaxK, c->k);
    const int n_gene = c->n_gene, k = c->k;
    // dense part
    int64_t n_cb = (8 * kNumSM) / ((n_gene + kEdGenes - 1) / kEdGenes);
    if (n_cb < 1) n_cb = 1;
    if (n_cb > (c->n_cell + kEdCells - 1) / kEdCells) n_cb = (c->n_cell + kEdCells - 1) / kEdCells;
    const int64_t cells_per_block = (c->n_cell + n_cb - 1) / n_cb;
    n_cb = (c->n_cell + cells_per_block - 1) / cells_per_block;
    SCDRS_CHECK(n_cb <= 65535, "too many cell blocks");
    SCDRS_TRY(c->partial.reserve((size_t)n_cb * 2 * n_gene * sizeof(double)));
    dim3 grid((n_gene + kEdGenes - 1) / kEdGenes, (unsigned)n_cb);
    gene_expm1_dense_kernel<<<grid, kEdGenes, 0, c->stream>>>(c->n_cell, n_gene, k, cells_per_block, c->cov_mat.as<double>(),
                                                             c->cov_beta.as<double>(), c->gene_mean.as<double>(), d_shift,
                                                             d_w, c->partial.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    gene_reduce_kernel<<<(unsigned)(((int64_t)n_gene * 2 + 255) / 256), 256, 0, c->stream>>>(n_gene, 2, n_cb, 0,
                                                                                               c->partial.as<double>(), d_out);
    SCDRS_LAUNCH_CHECK(c);
    // what the stored entries change
    return launch_accum(c, 2, k, 2, c->cov_mat.as<double>(), d_shift, d_w, d_out, 1);
}

}  // namespace scdrs
