// The step right after the scoring path: what scdrs.method.downstream_corr_analysis,
// downstream_group_analysis, test_gearysc and gearys_c (scdrs/method.py:712-865, 915-1085 of the reference) do
// with the n_cell x (1 + n_ctrl) matrix of normalised scores of one trait:
//   * Pearson correlation of a cell-level variable with every score column          (:853-861)
//   * per cell group, the 0.95 quantile of every score column                       (:797-806)
//   * per cell group, Geary's C of every column on the group's neighbour graph, the control columns first
//     given the trait's value distribution by ordinal rank                          (:967-990, :1044-1085)
// The reference runs the last one as a Python loop over cells, once per control and group; here the matrix
// is uploaded once and the whole of it is a segmented sort, two streaming reductions and one pass over the
// graph's edges.  Small per-group arithmetic (p-values, z-scores) stays in numpy on the host.
//
// The segmented sort is CUB's (a library call, like the reference's np.sort / rankdata); everything else is
// written here.  All arithmetic is fp64, like the reference's.
#include <cub/device/device_segmented_sort.cuh>

#include <vector>

#include "common.cuh"

namespace scdrs {

// ---- column moments of the row-major score matrix against V variables ---------------------------
// pass 0: sums of every score column and every variable; pass 1: centred products.
constexpr int kDsRowChunks = 64;

__global__ void __launch_bounds__(256)
ds_col_sum_kernel(int64_t n, int ncol, const double* __restrict__ S, double* __restrict__ part) {
    __shared__ double sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const int64_t r0 = n * blockIdx.y / gridDim.y, r1 = n * (blockIdx.y + 1) / gridDim.y;
    double s = 0.0;
    if (col < ncol)
        for (int64_t r = r0 + ty; r < r1; r += 8) s += S[r * ncol + col];
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < ncol) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][tx];
        part[(size_t)blockIdx.y * ncol + col] = t;
    }
}

__global__ void ds_part_reduce_kernel(int n_part, int m, const double* __restrict__ part, double scale,
                                      double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double t = 0.0;
    for (int k = 0; k < n_part; ++k) t += part[(size_t)k * m + j];
    out[j] = t * scale;
}

// centred second moments: part[chunk][0][col] = sum (y - ybar)^2, part[chunk][1 + v][col] = sum (x_v - xbar_v)(y - ybar)
__global__ void __launch_bounds__(256)
ds_col_cov_kernel(int64_t n, int ncol, int n_var, const double* __restrict__ S, const double* __restrict__ ymean,
                  const double* __restrict__ vars, const double* __restrict__ vmean, double* __restrict__ part) {
    __shared__ double sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const int v = (int)blockIdx.z - 1;          // -1: the column's own variance
    const int64_t r0 = n * blockIdx.y / gridDim.y, r1 = n * (blockIdx.y + 1) / gridDim.y;
    double s = 0.0;
    if (col < ncol) {
        const double ym = ymean[col];
        const double vm = v >= 0 ? vmean[v] : 0.0;
        for (int64_t r = r0 + ty; r < r1; r += 8) {
            const double dy = S[r * ncol + col] - ym;
            const double dx = v >= 0 ? vars[(size_t)v * n + r] - vm : dy;
            s += dx * dy;
        }
    }
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < ncol) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][tx];
        part[((size_t)blockIdx.y * (n_var + 1) + blockIdx.z) * ncol + col] = t;
    }
}

__global__ void ds_var_moments_kernel(int64_t n, const double* __restrict__ vars, double* __restrict__ vmean,
                                      double* __restrict__ vss) {
    // one block per variable: mean, then centred sum of squares (two passes, fixed tree order)
    __shared__ double sh[256];
    const double* x = vars + (size_t)blockIdx.x * n;
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += 256) s += x[r];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    const double m = sh[0] / (double)n;
    __syncthreads();
    s = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += 256) {
        const double d = x[r] - m;
        s += d * d;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        vmean[blockIdx.x] = m;
        vss[blockIdx.x] = sh[0];
    }
}

// corr[v][col] = Sxy / sqrt(Sxx) / sqrt(Syy)   (np.corrcoef: c / stddev[:, None] / stddev[None, :])
__global__ void ds_corr_finish_kernel(int ncol, int n_var, const double* __restrict__ mom, const double* __restrict__ vss,
                                      double* __restrict__ corr) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const double syy = mom[col];
    for (int v = 0; v < n_var; ++v) corr[(size_t)v * ncol + col] = mom[(size_t)(1 + v) * ncol + col] / sqrt(vss[v]) / sqrt(syy);
}

int ds_corr(scdrs_ctx* c, int n_var, const double* dev_vars, double* dev_corr) {
    const int64_t n = c->ds_n;
    const int ncol = c->ds_ncol;
    const double* S = c->ds_S.as<double>();
    const size_t m = (size_t)(n_var + 1) * ncol;
    SCDRS_TRY(c->ds_tmp.reserve(((size_t)kDsRowChunks * m + m + ncol + 2 * (size_t)n_var + 8) * sizeof(double)));
    double* part = c->ds_tmp.as<double>();
    double* mom = part + (size_t)kDsRowChunks * m;
    double* ymean = mom + m;
    double* vmean = ymean + ncol;
    double* vss = vmean + n_var;
    const dim3 g1((ncol + 31) / 32, kDsRowChunks);
    ds_col_sum_kernel<<<g1, 256, 0, c->stream>>>(n, ncol, S, part);
    SCDRS_LAUNCH_CHECK(c);
    ds_part_reduce_kernel<<<(ncol + 255) / 256, 256, 0, c->stream>>>(kDsRowChunks, ncol, part, 1.0 / (double)n, ymean);
    SCDRS_LAUNCH_CHECK(c);
    ds_var_moments_kernel<<<n_var, 256, 0, c->stream>>>(n, dev_vars, vmean, vss);
    SCDRS_LAUNCH_CHECK(c);
    const dim3 g2((ncol + 31) / 32, kDsRowChunks, n_var + 1);
    ds_col_cov_kernel<<<g2, 256, 0, c->stream>>>(n, ncol, n_var, S, ymean, dev_vars, vmean, part);
    SCDRS_LAUNCH_CHECK(c);
    ds_part_reduce_kernel<<<(int)((m + 255) / 256), 256, 0, c->stream>>>(kDsRowChunks, (int)m, part, 1.0, mom);
    SCDRS_LAUNCH_CHECK(c);
    ds_corr_finish_kernel<<<(ncol + 255) / 256, 256, 0, c->stream>>>(ncol, n_var, mom, vss, dev_corr);
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

// ---- per-group order statistics -------------------------------------------------------------------
// T[col][p] = S[order[p]][col] + 0.0 (cells in group order, one column after the other; -0 becomes +0 so
// that equal values tie like they do for np.sort), I[col][p] = position of p inside its group.
__global__ void __launch_bounds__(256)
ds_gather_cols_kernel(int64_t n, int ncol, int col0, int ncb, const double* __restrict__ S,
                      const int64_t* __restrict__ order, const uint32_t* __restrict__ local_pos,
                      double* __restrict__ T, uint32_t* __restrict__ I) {
    __shared__ double tile[32][33];
    __shared__ int64_t rows[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    if (threadIdx.x < 32) rows[threadIdx.x] = p0 + threadIdx.x < n ? order[p0 + threadIdx.x] : -1;
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = rows[k];
        const int col = c0 + tx;
        tile[k][tx] = (r >= 0 && col < ncb) ? S[r * ncol + col0 + col] + 0.0 : 0.0;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int col = c0 + k;
        const int64_t p = p0 + tx;
        if (col < ncb && p < n) {
            T[(size_t)col * n + p] = tile[tx][k];
            I[(size_t)col * n + p] = local_pos[p];
        }
    }
}

__global__ void ds_local_pos_kernel(int n_group, const int64_t* __restrict__ gptr, uint32_t* __restrict__ local_pos,
                                    int32_t* __restrict__ group_of) {
    const int g = blockIdx.x;
    const int64_t a = gptr[g], b = gptr[g + 1];
    for (int64_t p = a + threadIdx.x; p < b; p += blockDim.x) {
        local_pos[p] = (uint32_t)(p - a);
        group_of[p] = g;
    }
}

__global__ void ds_seg_offsets_kernel(int64_t n, int ncb, int n_group, const int64_t* __restrict__ gptr,
                                      int64_t* __restrict__ off) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n_seg = (int64_t)ncb * n_group;
    if (s > n_seg) return;
    if (s == n_seg) { off[s] = (int64_t)ncb * n; return; }
    const int col = (int)(s / n_group), g = (int)(s % n_group);
    off[s] = (int64_t)col * n + gptr[g];
}

// np.quantile(..., method="linear") of every (group, column): lo and gamma per group come from the host
// (numpy's own index arithmetic); lerp as numpy's _lerp (a + (b-a) t, or b - (b-a)(1-t) when t >= 0.5).
// A float32 score frame is interpolated like numpy does it for float32 input: b - a rounded to float32.
__global__ void ds_quantile_kernel(int64_t n, int ncol, int col0, int ncb, int n_group, int is_f32, const int64_t* __restrict__ gptr,
                                   const int32_t* __restrict__ q_lo, const double* __restrict__ q_gamma,
                                   const double* __restrict__ T, double* __restrict__ out) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= (int64_t)ncb * n_group) return;
    const int col = (int)(s / n_group), g = (int)(s % n_group);
    const int64_t a0 = gptr[g], len = gptr[g + 1] - a0;
    double res = __longlong_as_double(0x7ff8000000000000LL);
    if (len > 0) {
        const int64_t lo = q_lo[g], hi = lo + 1 < len ? lo + 1 : len - 1;
        const double a = T[(size_t)col * n + a0 + lo], b = T[(size_t)col * n + a0 + hi], t = q_gamma[g];
        double d = __dsub_rn(b, a);
        if (is_f32) d = (double)__double2float_rn(d);
        res = t >= 0.5 ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(a, __dmul_rn(d, t));
    }
    out[(size_t)g * ncol + col0 + col] = res;
}

// control_distribution_match (scdrs/method.py:976-981): the cell holding the r-th smallest value of a column
// (ties by cell order) receives the r-th smallest trait score of its group.  sorted trait column = ref.
__global__ void __launch_bounds__(256)
ds_match_kernel(int64_t n, int ncol, int col0, int ncb, const int64_t* __restrict__ gptr, const int32_t* __restrict__ group_of,
                const uint32_t* __restrict__ I_sorted, const double* __restrict__ ref_sorted, double* __restrict__ M) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= (int64_t)ncb * n) return;
    const int col = (int)(e / n);
    const int64_t p = e % n;
    const int64_t base = gptr[group_of[p]];
    M[(size_t)(base + I_sorted[e]) * ncol + col0 + col] = ref_sorted[p];
}

__global__ void __launch_bounds__(256)
ds_gather_rows_kernel(int64_t n, int ncol, const double* __restrict__ S, const int64_t* __restrict__ order,
                      double* __restrict__ M) {
    const int64_t p = blockIdx.x;
    const int64_t r = order[p];
    for (int col = threadIdx.x; col < ncol; col += blockDim.x) M[p * ncol + col] = S[r * ncol + col];
}

// per (group, column) of the row-major matrix in group order: sum (x - mean)^2 (two passes; the denominator
// of Geary's C, scdrs/method.py:1081).  A group's rows are split over kSsSplit blocks (one group may hold half
// of all cells); partial results are combined in split order.
constexpr int kSsSplit = 16;

// pass = 0: part[g][z][col] = sum of the split's rows; pass = 1: sum (x - mean)^2 with mean from pass 0
__global__ void __launch_bounds__(256)
ds_group_ss_kernel(int ncol, int pass, const int64_t* __restrict__ gptr, const double* __restrict__ M,
                   const double* __restrict__ sums, double* __restrict__ part) {
    __shared__ double sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int g = blockIdx.y, z = blockIdx.z;
    const int col = blockIdx.x * 32 + tx;
    const int64_t a = gptr[g], b = gptr[g + 1];
    const int64_t r0 = a + (b - a) * z / kSsSplit, r1 = a + (b - a) * (z + 1) / kSsSplit;
    double m = 0.0;
    if (pass == 1 && col < ncol) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kSsSplit; ++k) t += sums[((size_t)g * kSsSplit + k) * ncol + col];
        m = t / (double)(b - a);
    }
    double s = 0.0;
    if (col < ncol)
        for (int64_t p = r0 + ty; p < r1; p += 8) {
            const double d = M[p * ncol + col] - m;
            s += pass == 1 ? d * d : d;
        }
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < ncol) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][tx];
        part[((size_t)g * kSsSplit + z) * ncol + col] = t;
    }
}

__global__ void ds_group_ss_finish_kernel(int ncol, const double* __restrict__ part, double* __restrict__ out) {
    const int g = blockIdx.y;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kSsSplit; ++k) t += part[((size_t)g * kSsSplit + k) * ncol + col];
    out[(size_t)g * ncol + col] = t;
}

// Geary's C numerators: sum over the stored within-group edges (p, q) of w (x_p - x_q)^2, for every column at
// once.  A block owns up to kGearyRows consecutive cells of one group and 1024 columns (4 per thread); every
// edge costs one row read of the neighbour (L2: the cells of a group sit next to each other).
constexpr int kGearyRows = 8;

__global__ void __launch_bounds__(256)
ds_geary_kernel(int ncol, const int64_t* __restrict__ chunk_p0, const int32_t* __restrict__ chunk_len,
                const int64_t* __restrict__ g_indptr, const int32_t* __restrict__ g_nbr, const double* __restrict__ g_w,
                const double* __restrict__ M, double* __restrict__ part) {
    const int64_t p0 = chunk_p0[blockIdx.x];
    const int len = chunk_len[blockIdx.x];
    for (int ct = 0; ct < ncol; ct += 1024) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < len; ++k) {
            const int64_t p = p0 + k;
            double xi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int col = ct + (int)threadIdx.x + 256 * u;
                xi[u] = col < ncol ? M[p * ncol + col] : 0.0;
            }
            const int64_t e0 = g_indptr[p], e1 = g_indptr[p + 1];
            for (int64_t e = e0; e < e1; ++e) {
                const int64_t q = g_nbr[e];
                const double w = g_w[e];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int col = ct + (int)threadIdx.x + 256 * u;
                    if (col < ncol) {
                        const double d = xi[u] - M[q * ncol + col];
                        acc[u] += w * (d * d);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int col = ct + (int)threadIdx.x + 256 * u;
            if (col < ncol) part[(size_t)blockIdx.x * ncol + col] = acc[u];
        }
    }
}

// totals per group: eight interleaved running sums over the group's chunks, combined in a fixed order
// (results repeat bit for bit)
__global__ void __launch_bounds__(256)
ds_geary_reduce_kernel(int ncol, const int64_t* __restrict__ grp_chunk_ptr, const double* __restrict__ part,
                       double* __restrict__ out) {
    __shared__ double sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const int col = blockIdx.x * 32 + tx;
    double t = 0.0;
    if (col < ncol)
        for (int64_t k = grp_chunk_ptr[g] + ty; k < grp_chunk_ptr[g + 1]; k += 8) t += part[(size_t)k * ncol + col];
    sh[ty][tx] = t;
    __syncthreads();
    if (ty == 0 && col < ncol) {
        double r = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) r += sh[k][tx];
        out[(size_t)g * ncol + col] = r;
    }
}

// Everything per grouping.  dev pointers: order[n], gptr[n_group + 1] (int64), q_lo / q_gamma per group (or null),
// graph in group order (or null).  Outputs dev: quantile / total / denom [n_group x ncol] (null = skipped).
int ds_group_stats(scdrs_ctx* c, int n_group, const int64_t* h_gptr, const int64_t* d_order, const int64_t* d_gptr,
                   const int32_t* d_qlo, const double* d_qgamma, double* d_quant, int geary_mode,
                   const int64_t* d_gindptr, const int32_t* d_gnbr, const double* d_gw, double* d_total,
                   double* d_denom) {
    const int64_t n = c->ds_n;
    const int ncol = c->ds_ncol;
    const double* S = c->ds_S.as<double>();
    const bool need_sort = d_quant != nullptr || geary_mode == 1;
    const bool need_M = geary_mode != 0;
    SCDRS_TRY(c->ds_pos.reserve((size_t)n * 8 + 64));
    uint32_t* local_pos = c->ds_pos.as<uint32_t>();
    int32_t* group_of = reinterpret_cast<int32_t*>(local_pos + n);
    ds_local_pos_kernel<<<n_group, 256, 0, c->stream>>>(n_group, d_gptr, local_pos, group_of);
    SCDRS_LAUNCH_CHECK(c);
    double* M = nullptr;
    if (need_M) {
        SCDRS_TRY(c->ds_M.reserve((size_t)n * ncol * sizeof(double)));
        M = c->ds_M.as<double>();
    }
    if (need_sort) {
        // column batches keep every sort below 2^31 keys and the sort buffers bounded
        int ncb_max = (int)(((int64_t)1 << 30) / (n > 0 ? n : 1));
        if (ncb_max < 1) ncb_max = 1;
        if (ncb_max > ncol) ncb_max = ncol;
        SCDRS_CHECK(n < ((int64_t)1 << 31), "too many cells for one grouping");
        const size_t items = (size_t)ncb_max * n;
        SCDRS_TRY(c->ds_keys.reserve(items * 2 * sizeof(double) + 64));
        SCDRS_TRY(c->ds_vals.reserve(items * 2 * sizeof(uint32_t) + 64));
        SCDRS_TRY(c->ds_off.reserve(((size_t)ncb_max * n_group + 2) * sizeof(int64_t)));
        SCDRS_TRY(c->ds_ref.reserve((size_t)n * sizeof(double)));
        double* T = c->ds_keys.as<double>();
        double* T2 = T + items;
        uint32_t* I = c->ds_vals.as<uint32_t>();
        uint32_t* I2 = I + items;
        int64_t* off = c->ds_off.as<int64_t>();
        for (int col0 = 0; col0 < ncol; col0 += ncb_max) {
            const int ncb = ncol - col0 < ncb_max ? ncol - col0 : ncb_max;
            const dim3 gg((unsigned)((n + 31) / 32), (ncb + 31) / 32);
            ds_gather_cols_kernel<<<gg, 256, 0, c->stream>>>(n, ncol, col0, ncb, S, d_order, local_pos, T, I);
            SCDRS_LAUNCH_CHECK(c);
            const int64_t n_seg = (int64_t)ncb * n_group;
            ds_seg_offsets_kernel<<<(unsigned)((n_seg + 256) / 256), 256, 0, c->stream>>>(n, ncb, n_group, d_gptr, off);
            SCDRS_LAUNCH_CHECK(c);
            size_t tmp_bytes = 0;
            SCDRS_CUDA(cub::DeviceSegmentedSort::StableSortPairs(nullptr, tmp_bytes, T, T2, I, I2, (int64_t)ncb * n, n_seg,
                                                                 off, off + 1, c->stream));
            SCDRS_TRY(c->ds_tmp.reserve(tmp_bytes + 64));
            SCDRS_CUDA(cub::DeviceSegmentedSort::StableSortPairs(c->ds_tmp.p, tmp_bytes, T, T2, I, I2, (int64_t)ncb * n,
                                                                 n_seg, off, off + 1, c->stream));
            c->launches += 3;
            if (d_quant) {
                ds_quantile_kernel<<<(unsigned)((n_seg + 255) / 256), 256, 0, c->stream>>>(n, ncol, col0, ncb, n_group, c->ds_f32 ? 1 : 0, d_gptr,
                                                                                          d_qlo, d_qgamma, T2, d_quant);
                SCDRS_LAUNCH_CHECK(c);
            }
            if (geary_mode == 1) {
                if (col0 == 0)      // the sorted trait scores of every group: the reference distribution
                    SCDRS_CUDA(cudaMemcpyAsync(c->ds_ref.p, T2, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
                const int64_t ne = (int64_t)ncb * n;
                ds_match_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, c->stream>>>(n, ncol, col0, ncb, d_gptr, group_of, I2,
                                                                                    c->ds_ref.as<double>(), M);
                SCDRS_LAUNCH_CHECK(c);
            }
        }
    }
    if (geary_mode == 2) {
        ds_gather_rows_kernel<<<(unsigned)n, 256, 0, c->stream>>>(n, ncol, S, d_order, M);
        SCDRS_LAUNCH_CHECK(c);
    }
    if (geary_mode != 0) {
        const dim3 gs((ncol + 31) / 32, n_group, kSsSplit);
        SCDRS_TRY(c->ds_tmp.reserve((size_t)2 * n_group * kSsSplit * ncol * sizeof(double)));
        double* ss_sum = c->ds_tmp.as<double>();
        double* ss_part = ss_sum + (size_t)n_group * kSsSplit * ncol;
        ds_group_ss_kernel<<<gs, 256, 0, c->stream>>>(ncol, 0, d_gptr, M, nullptr, ss_sum);
        SCDRS_LAUNCH_CHECK(c);
        ds_group_ss_kernel<<<gs, 256, 0, c->stream>>>(ncol, 1, d_gptr, M, ss_sum, ss_part);
        SCDRS_LAUNCH_CHECK(c);
        ds_group_ss_finish_kernel<<<dim3((ncol + 255) / 256, n_group), 256, 0, c->stream>>>(ncol, ss_part, d_denom);
        SCDRS_LAUNCH_CHECK(c);
        // chunks of up to kGearyRows cells that never straddle a group
        std::vector<int64_t> cp0, gcp(n_group + 1, 0);
        std::vector<int32_t> clen;
        for (int g = 0; g < n_group; ++g) {
            for (int64_t p = h_gptr[g]; p < h_gptr[g + 1]; p += kGearyRows) {
                cp0.push_back(p);
                clen.push_back((int32_t)(h_gptr[g + 1] - p < kGearyRows ? h_gptr[g + 1] - p : kGearyRows));
            }
            gcp[g + 1] = (int64_t)cp0.size();
        }
        const size_t nch = cp0.size();
        SCDRS_TRY(c->ds_chunk.reserve(nch * 12 + (size_t)(n_group + 1) * 8 + 64));
        int64_t* d_cp0 = c->ds_chunk.as<int64_t>();
        int64_t* d_gcp = d_cp0 + nch;
        int32_t* d_clen = reinterpret_cast<int32_t*>(d_gcp + n_group + 1);
        SCDRS_CUDA(cudaMemcpyAsync(d_cp0, cp0.data(), nch * 8, cudaMemcpyHostToDevice, c->stream));
        SCDRS_CUDA(cudaMemcpyAsync(d_gcp, gcp.data(), (size_t)(n_group + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        SCDRS_CUDA(cudaMemcpyAsync(d_clen, clen.data(), nch * 4, cudaMemcpyHostToDevice, c->stream));
        SCDRS_TRY(c->ds_part.reserve((nch > 0 ? nch : 1) * (size_t)ncol * sizeof(double)));
        if (nch > 0) {
            ds_geary_kernel<<<(unsigned)nch, 256, 0, c->stream>>>(ncol, d_cp0, d_clen, d_gindptr, d_gnbr, d_gw, M,
                                                                 c->ds_part.as<double>());
            SCDRS_LAUNCH_CHECK(c);
        }
        const dim3 gr((ncol + 31) / 32, n_group);
        ds_geary_reduce_kernel<<<gr, 256, 0, c->stream>>>(ncol, d_gcp, c->ds_part.as<double>(), d_total);
        SCDRS_LAUNCH_CHECK(c);
        SCDRS_CUDA(cudaStreamSynchronize(c->stream));     // the chunk vectors go out of scope
    }
    return 0;
}

// The caller's frame -> S[n x ncol] fp64 row-major: element (r, j) of the scores is src[r * rs + cols[j] * cs]
// (a row-major frame has cs = 1, a column-major one -- what pandas holds after most operations -- rs = 1).
template <typename T>
__global__ void __launch_bounds__(256)
ds_import_rowmajor_kernel(int64_t n, int ncol, int64_t rs, const int32_t* __restrict__ cols, const T* __restrict__ src,
                          double* __restrict__ S) {
    const int64_t r = blockIdx.x;
    for (int j = threadIdx.x; j < ncol; j += blockDim.x) S[r * ncol + j] = (double)src[r * rs + cols[j]];
}

template <typename T>
__global__ void __launch_bounds__(256)
ds_import_colmajor_kernel(int64_t n, int ncol, int64_t cs, const int32_t* __restrict__ cols, const T* __restrict__ src,
                          double* __restrict__ S) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    for (int k = ty; k < 32; k += 8) {
        const int j = j0 + k;
        const int64_t r = r0 + tx;
        tile[k][tx] = (j < ncol && r < n) ? (double)src[(int64_t)cols[j] * cs + r] : 0.0;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = r0 + k;
        const int j = j0 + tx;
        if (r < n && j < ncol) S[r * ncol + j] = tile[tx][k];
    }
}

int ds_import(scdrs_ctx* c, int64_t n, int ncol, int is_f32, int64_t rs, int64_t cs, const int32_t* dev_cols,
              const void* dev_src, double* dev_S) {
    if (cs == 1) {
        if (is_f32) ds_import_rowmajor_kernel<float><<<(unsigned)n, 256, 0, c->stream>>>(n, ncol, rs, dev_cols, (const float*)dev_src, dev_S);
        else ds_import_rowmajor_kernel<double><<<(unsigned)n, 256, 0, c->stream>>>(n, ncol, rs, dev_cols, (const double*)dev_src, dev_S);
    } else {
        const dim3 g((unsigned)((n + 31) / 32), (ncol + 31) / 32);
        if (is_f32) ds_import_colmajor_kernel<float><<<g, 256, 0, c->stream>>>(n, ncol, cs, dev_cols, (const float*)dev_src, dev_S);
        else ds_import_colmajor_kernel<double><<<g, 256, 0, c->stream>>>(n, ncol, cs, dev_cols, (const double*)dev_src, dev_S);
    }
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

}  // namespace scdrs
