// K1: raw scores of the trait gene set and its n_ctrl control gene sets for every cell.
//
// Replaces the 1 + n_ctrl sequential calls of _compute_raw_score (scdrs/method.py:172-183,
// 312-402 of the reference): raw[c, j] = sum_g X[c, g] * W[g, j]  (+ COV_MAT[c,:] . M[:, j] + m[j]
// in implicit-covariate mode), where column j of W holds the normalised weights of gene set j.
//
// Layout: X stays CSR (cells-major, the way cells are sharded across GPUs).  W is stored
// gene-major ("for gene g: the (set, weight) pairs that contain g"), so that one warp owns one
// cell, walks that cell's non-zeros once and scatters x * w into 1 + n_ctrl accumulators that
// live in shared memory.  Lanes split ONE gene's (set, weight) list, so the targets of a warp
// instruction are distinct columns; no atomics anywhere.  Accumulation is fp64 (the reference
// accumulates float32 data in float64, scdrs/method.py:392-400); the result is written as an
// fp64 per-cell reference value plus fp32 deviations from it, with NaN marking exact zeros
// (the reference's `raw == 0` rule, scdrs/method.py:522-523).
#include "common.cuh"

namespace scdrs {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block-wide sum, result broadcast; sh needs 32 doubles
__device__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nwarp; ++i) t += sh[i];
    return t;
}

// One block per gene set: normalised weights (scdrs/method.py:366-375), the covariate
// projections M[:, j] = COV_BETA[S_j]^T w_j and m[j] = COV_GENE_MEAN[S_j] . w_j (:393-394) and the
// variance-ratio numerator sum_g var_g w_g^2 (:191-195).
__global__ void prep_sets_kernel(int n_s0, int n_sc, int ncol, int k, int weight_opt,
                                 const int32_t* __restrict__ set_genes,
                                 const double* __restrict__ gw,  // [n_s0] trait, then [n_sc] control
                                 const double* __restrict__ explicit_w,
                                 const double* __restrict__ var, const double* __restrict__ var_tech,
                                 const double* __restrict__ beta, const double* __restrict__ gmean,
                                 float* __restrict__ set_w, double* __restrict__ covM,
                                 double* __restrict__ covm, double* __restrict__ rnum) {
    __shared__ double sh[32];
    const int j = blockIdx.x;
    const int n_s = j == 0 ? n_s0 : n_sc;
    const size_t off = j == 0 ? 0 : (size_t)n_s0 + (size_t)(j - 1) * n_sc;
    const int32_t* genes = set_genes + off;
    float* wout = set_w + off;
    double s = 1.0;
    if (explicit_w == nullptr) {
        const double* g = gw + (j == 0 ? 0 : n_s0);
        double part = 0.0;
        for (int p = threadIdx.x; p < n_s; p += blockDim.x) {
            const int gi = genes[p];
            double base = 1.0;
            if (weight_opt == SCDRS_WEIGHT_VS) base = 1.0 / sqrt(var_tech[gi] + 1e-2);
            if (weight_opt == SCDRS_WEIGHT_INV_STD) base = 1.0 / sqrt(var[gi] + 1e-2);
            part += base * g[p];
        }
        s = block_sum(part, sh);
    }
    double pm = 0.0, pr = 0.0;
    for (int p = threadIdx.x; p < n_s; p += blockDim.x) {
        const int gi = genes[p];
        double w;
        if (explicit_w != nullptr) {
            w = explicit_w[off + p];
        } else {
            const double* g = gw + (j == 0 ? 0 : n_s0);
            double base = 1.0;
            if (weight_opt == SCDRS_WEIGHT_VS) base = 1.0 / sqrt(var_tech[gi] + 1e-2);
            if (weight_opt == SCDRS_WEIGHT_INV_STD) base = 1.0 / sqrt(var[gi] + 1e-2);
            w = base * g[p] / s;
        }
        wout[p] = (float)w;
        // use the weight the scatter kernel will actually multiply with
        const double wf = (double)(float)w;
        if (k > 0) pm += gmean[gi] * wf;
        if (var != nullptr) pr += var[gi] * w * w;
    }
    pm = block_sum(pm, sh);
    pr = block_sum(pr, sh);
    if (threadIdx.x == 0) {
        covm[j] = pm;
        rnum[j] = pr;
    }
    for (int kk = 0; kk < k; ++kk) {
        double pb = 0.0;
        for (int p = threadIdx.x; p < n_s; p += blockDim.x)
            pb += beta[(size_t)genes[p] * k + kk] * (double)wout[p];
        pb = block_sum(pb, sh);
        if (threadIdx.x == 0) covM[(size_t)kk * ncol + j] = pb;
    }
}

// 1/sqrt(var ratio) per column; column 0 (the trait) and the no-ratio case get 1.
__global__ void ratio_kernel(int ncol, int use_ratio, const double* __restrict__ rnum,
                             double* __restrict__ ratio_a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncol) return;
    double a = 1.0;
    if (use_ratio && j > 0) a = 1.0 / sqrt(rnum[j] / rnum[0]);
    ratio_a[j] = a;
}

// ---- gene-major weight lists ---------------------------------------------------------------
__global__ void w_count_kernel(int64_t n_entry, const int32_t* __restrict__ set_genes,
                               uint32_t* __restrict__ cnt) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_entry;
         e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[set_genes[e]], 1u);
}

__global__ void w_fill_kernel(int64_t n_entry, int n_s0, int n_sc, const int32_t* __restrict__ set_genes,
                              const float* __restrict__ set_w, const uint32_t* __restrict__ w_ptr,
                              uint32_t* __restrict__ cursor, uint16_t* __restrict__ w_col,
                              float* __restrict__ w_val) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_entry;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int g = set_genes[e];
        const uint32_t pos = w_ptr[g] + atomicAdd(&cursor[g], 1u);
        w_col[pos] = (uint16_t)(e < n_s0 ? 0 : 1 + (e - n_s0) / n_sc);
        w_val[pos] = set_w[e];
    }
}

// Sort each gene's list by column so the layout (and hence the shared-memory access pattern) is
// reproducible run to run.  Lists are short (about (1+n_ctrl)*n_s/n_gene entries).
__global__ void w_sort_rows_kernel(int n_gene, const uint32_t* __restrict__ w_ptr,
                                   uint16_t* __restrict__ w_col, float* __restrict__ w_val) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gene) return;
    const uint32_t s = w_ptr[g], e = w_ptr[g + 1];
    for (uint32_t i = s + 1; i < e; ++i) {
        const uint16_t c = w_col[i];
        const float v = w_val[i];
        uint32_t q = i;
        while (q > s && w_col[q - 1] > c) {
            w_col[q] = w_col[q - 1];
            w_val[q] = w_val[q - 1];
            --q;
        }
        w_col[q] = c;
        w_val[q] = v;
    }
}

static int upload_sets(scdrs_ctx* c, const int32_t* trait_genes, const int32_t* ctrl_genes,
                       const double* gw_trait, const double* gw_ctrl, const double* explicit_w) {
    const int32_t n_s = c->n_s, n_sc = c->n_sc;
    const size_t n_entry = (size_t)n_s + (size_t)c->n_ctrl * n_sc;
    SCDRS_TRY(c->set_genes.reserve(n_entry * sizeof(int32_t)));
    SCDRS_TRY(c->set_w.reserve(n_entry * sizeof(float)));
    SCDRS_TRY(c->set_gw.reserve(((size_t)n_s + n_sc) * sizeof(double)));
    if (trait_genes != nullptr) {
        SCDRS_CUDA(cudaMemcpyAsync(c->set_genes.p, trait_genes, n_s * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, c->stream));
        if (n_entry > (size_t)n_s)
            SCDRS_CUDA(cudaMemcpyAsync(c->set_genes.as<int32_t>() + n_s, ctrl_genes,
                                       (n_entry - n_s) * sizeof(int32_t), cudaMemcpyHostToDevice,
                                       c->stream));
        SCDRS_CUDA(cudaMemcpyAsync(c->set_gw.p, gw_trait, n_s * sizeof(double),
                                   cudaMemcpyHostToDevice, c->stream));
        if (n_sc > 0 && c->n_ctrl > 0)
            SCDRS_CUDA(cudaMemcpyAsync(c->set_gw.as<double>() + n_s, gw_ctrl, n_sc * sizeof(double),
                                       cudaMemcpyHostToDevice, c->stream));
    } else {
        SCDRS_CUDA(cudaMemcpyAsync(c->set_genes.p, ctrl_genes, n_entry * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, c->stream));
        SCDRS_TRY(c->stage.reserve(n_entry * sizeof(double)));
        SCDRS_CUDA(cudaMemcpyAsync(c->stage.p, explicit_w, n_entry * sizeof(double),
                                   cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
}

static int build_w(scdrs_ctx* c, int weight_opt, int use_var_ratio, bool explicit_w) {
    const int ncol = c->ncol, n_s = c->n_s, n_sc = c->n_sc, n_gene = c->n_gene;
    const int64_t n_entry = (int64_t)n_s + (int64_t)c->n_ctrl * n_sc;
    SCDRS_TRY(c->covM.reserve((size_t)(c->k > 0 ? c->k : 1) * ncol * sizeof(double)));
    SCDRS_TRY(c->covm.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->ratio_a.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->colmean.reserve(ncol * sizeof(double)));  // reused as rnum scratch here
    SCDRS_TRY(c->w_ptr.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    SCDRS_TRY(c->w_cursor.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    SCDRS_TRY(c->w_col.reserve(n_entry * sizeof(uint16_t)));
    SCDRS_TRY(c->w_val.reserve(n_entry * sizeof(float)));

    double* rnum = c->colmean.as<double>();
    prep_sets_kernel<<<ncol, 256, 0, c->stream>>>(
        n_s, n_sc, ncol, c->k, weight_opt, c->set_genes.as<int32_t>(), c->set_gw.as<double>(),
        explicit_w ? c->stage.as<double>() : nullptr,
        c->have_stats ? c->var.as<double>() : nullptr,
        c->have_stats ? c->var_tech.as<double>() : nullptr, c->cov_beta.as<double>(),
        c->gene_mean.as<double>(), c->set_w.as<float>(), c->covM.as<double>(),
        c->covm.as<double>(), rnum);
    SCDRS_LAUNCH_CHECK(c);
    ratio_kernel<<<(ncol + 255) / 256, 256, 0, c->stream>>>(ncol, use_var_ratio, rnum,
                                                            c->ratio_a.as<double>());
    SCDRS_LAUNCH_CHECK(c);

    SCDRS_CUDA(cudaMemsetAsync(c->w_cursor.p, 0, (size_t)(n_gene + 1) * sizeof(uint32_t), c->stream));
    const int nb = (int)((n_entry + 255) / 256 < 1184 ? (n_entry + 255) / 256 : 1184);
    w_count_kernel<<<nb, 256, 0, c->stream>>>(n_entry, c->set_genes.as<int32_t>(),
                                              c->w_cursor.as<uint32_t>());
    SCDRS_LAUNCH_CHECK(c);
    SCDRS_TRY(exclusive_scan_u32(c, c->w_cursor.as<uint32_t>(), c->w_ptr.as<uint32_t>(), n_gene + 1));
    SCDRS_CUDA(cudaMemsetAsync(c->w_cursor.p, 0, (size_t)(n_gene + 1) * sizeof(uint32_t), c->stream));
    w_fill_kernel<<<nb, 256, 0, c->stream>>>(n_entry, n_s, n_sc > 0 ? n_sc : 1, c->set_genes.as<int32_t>(),
                                             c->set_w.as<float>(), c->w_ptr.as<uint32_t>(),
                                             c->w_cursor.as<uint32_t>(), c->w_col.as<uint16_t>(),
                                             c->w_val.as<float>());
    SCDRS_LAUNCH_CHECK(c);
    w_sort_rows_kernel<<<(n_gene + 127) / 128, 128, 0, c->stream>>>(
        n_gene, c->w_ptr.as<uint32_t>(), c->w_col.as<uint16_t>(), c->w_val.as<float>());
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

static int set_shape(scdrs_ctx* c, int32_t ncol, int32_t n_s, int32_t n_sc) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    SCDRS_CHECK(ncol >= 1 && ncol <= 65535, "1 + n_ctrl = %d outside [1, 65535]", ncol);
    SCDRS_CHECK(n_s >= 1, "empty gene set");
    SCDRS_CHECK(n_sc >= 0 && n_sc <= n_s, "control set size %d outside [0, %d]", n_sc, n_s);
    c->n_sc = n_sc;
    c->ncol = ncol;
    c->n_ctrl = ncol - 1;
    c->n_s = n_s;
    c->ld = (ncol + 3) & ~3;
    c->trait_ready = false;
    return 0;
}

int trait_prepare(scdrs_ctx* c, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                  const int32_t* trait_genes, const double* trait_gw, const int32_t* ctrl_genes,
                  const double* ctrl_gw, int32_t weight_opt, int32_t use_var_ratio) {
    SCDRS_CHECK(n_ctrl >= 0, "n_ctrl < 0");
    SCDRS_CHECK(weight_opt >= 0 && weight_opt <= 2, "illegal weight_opt %d", weight_opt);
    SCDRS_CHECK(weight_opt == SCDRS_WEIGHT_UNIFORM || c->have_stats,
                "weight_opt needs gene statistics (call scdrs_set_gene_stats)");
    SCDRS_CHECK(!use_var_ratio || c->have_stats, "variance ratio needs gene statistics");
    SCDRS_TRY(set_shape(c, n_ctrl + 1, n_s, n_s_ctrl));
    SCDRS_TRY(upload_sets(c, trait_genes, ctrl_genes, trait_gw, ctrl_gw, nullptr));
    return build_w(c, weight_opt, use_var_ratio, false);
}

int explicit_prepare(scdrs_ctx* c, int32_t n_set, int32_t n_s, const int32_t* genes,
                     const double* w) {
    SCDRS_TRY(set_shape(c, n_set, n_s, n_s));
    SCDRS_TRY(upload_sets(c, nullptr, genes, nullptr, nullptr, w));
    return build_w(c, SCDRS_WEIGHT_UNIFORM, 0, true);
}

// ---- the scatter kernel ----------------------------------------------------------------------
// One warp per cell.  acc[] (1 + n_ctrl doubles) lives in shared memory.
__global__ void __launch_bounds__(256)
spmm_score_kernel(int64_t n_cell, int ncol, int ld, int acc_stride,
                  const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const float* __restrict__ data, const uint32_t* __restrict__ w_ptr,
                  const uint16_t* __restrict__ w_col, const float* __restrict__ w_val, int k,
                  const double* __restrict__ cov_mat, const double* __restrict__ covM,
                  const double* __restrict__ covm, float* __restrict__ dev_mat,
                  double* __restrict__ row_ref) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* acc = reinterpret_cast<double*>(smem_raw) + (size_t)warp * acc_stride;

    for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < n_cell;
         row += (int64_t)gridDim.x * wpb) {
        for (int j = lane; j < ncol; j += 32) acc[j] = 0.0;
        __syncwarp();
        const int64_t p0 = indptr[row], p1 = indptr[row + 1];
        for (int64_t base = p0; base < p1; base += 32) {
            const int64_t q = base + lane;
            uint32_t ws = 0, we = 0;
            float x = 0.f;
            if (q < p1) {
                const int g = indices[q];
                x = data[q];
                ws = w_ptr[g];
                we = w_ptr[g + 1];
            }
            const int n = (int)((p1 - base) < 32 ? (p1 - base) : 32);
            for (int i = 0; i < n; ++i) {
                const uint32_t s = __shfl_sync(0xffffffffu, ws, i);
                const uint32_t e = __shfl_sync(0xffffffffu, we, i);
                const double xi = (double)__shfl_sync(0xffffffffu, x, i);
                for (uint32_t t = s + lane; t < e; t += 32) {
                    const int col = w_col[t];
                    acc[col] = fma(xi, (double)w_val[t], acc[col]);
                }
                __syncwarp();
            }
        }
        // epilogue: covariate terms, row reference, deviations
        double rs = 0.0;
        for (int j = lane; j < ncol; j += 32) {
            double v = acc[j];
            if (k > 0) {
                double cterm = covm[j];
                for (int kk = 0; kk < k; ++kk)
                    cterm = fma(cov_mat[row * k + kk], covM[(size_t)kk * ncol + j], cterm);
                v += cterm;
                acc[j] = v;
            }
            rs += v;
        }
        rs = warp_sum(rs);
        const double ref = rs / (double)ncol;
        float* out = dev_mat + row * (int64_t)ld;
        for (int j = lane; j < ncol; j += 32) {
            const double v = acc[j];
            out[j] = (v == 0.0) ? __int_as_float(0x7fc00000) : (float)(v - ref);
        }
        if (lane == 0) row_ref[row] = ref;
        __syncwarp();
    }
}

int spmm_score(scdrs_ctx* c) {
    const int ncol = c->ncol;
    SCDRS_TRY(c->dev_mat.reserve((size_t)c->n_cell * c->ld * sizeof(float)));
    SCDRS_TRY(c->row_ref.reserve((size_t)c->n_cell * sizeof(double)));
    const int acc_stride = (ncol + 1) & ~1;
    const size_t per_warp = (size_t)acc_stride * sizeof(double);
    SCDRS_CHECK(per_warp <= 200 * 1024, "1 + n_ctrl = %d needs more shared memory than one SM has", ncol);
    int wpb = (int)((64 * 1024) / per_warp);
    if (wpb > 8) wpb = 8;
    if (wpb < 1) wpb = 1;
    const size_t smem = per_warp * wpb;
    static bool attr_set = false;
    if (!attr_set) {
        SCDRS_CUDA(cudaFuncSetAttribute(spmm_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024));
        attr_set = true;
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * wpb > 48) per_sm = 48 / wpb;
    int64_t grid = (int64_t)kNumSM * per_sm;
    const int64_t need = (c->n_cell + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (c->timed_spmm) SCDRS_CUDA(cudaEventRecord(c->ev[0], c->stream));
    spmm_score_kernel<<<(unsigned)grid, wpb * 32, smem, c->stream>>>(
        c->n_cell, ncol, c->ld, acc_stride, c->indptr, c->indices, c->data,
        c->w_ptr.as<uint32_t>(), c->w_col.as<uint16_t>(), c->w_val.as<float>(), c->k,
        c->cov_mat.as<double>(), c->covM.as<double>(), c->covm.as<double>(),
        c->dev_mat.as<float>(), c->row_ref.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    if (c->timed_spmm) SCDRS_CUDA(cudaEventRecord(c->ev[1], c->stream));
    c->trait_ready = true;
    return 0;
}

}  // namespace scdrs
