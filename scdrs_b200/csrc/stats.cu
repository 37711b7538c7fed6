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
constexpr int kGeneTile = 32;    // 626 tiles at 20k genes: ~4 resident CTAs per SM (each is one latency chain)
constexpr int kRowTile = 64;
constexpr int kStatThreadsSeq = 2 * kGeneTile;   // consumers (one per gene) + producers (one per staged row)

// tptr[t][r] = number of stored entries of row r with column < t * kGeneTile (t = 0 .. n_tile): where
// a gene tile starts inside every CSR row.  Built once per matrix by one warp per row; it replaces an
// 11-step dependent binary search per (row, tile) by one coalesced load.
__global__ void __launch_bounds__(256)
tile_ptr_kernel(int64_t n_cell, int n_tile, const int64_t* __restrict__ indptr,
                const int32_t* __restrict__ indices, uint32_t* __restrict__ tptr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < n_cell; row += nwarp) {
        const int64_t p0 = indptr[row], p1 = indptr[row + 1];
        const int64_t len = p1 - p0;
        // entry i opens the tiles (tile(i-1), tile(i)]; the tail (tile(last), n_tile] gets len
        for (int64_t i = lane; i <= len; i += 32) {
            const int prev = i > 0 ? indices[p0 + i - 1] / kGeneTile : -1;
            const int cur = i < len ? indices[p0 + i] / kGeneTile : n_tile;
            for (int t = prev + 1; t <= cur; ++t) tptr[(size_t)t * n_cell + row] = (uint32_t)i;
        }
    }
}

static int ensure_tile_ptr(scdrs_ctx* c) {
    const int n_tile = (c->n_gene + kGeneTile - 1) / kGeneTile;
    const size_t bytes = (size_t)(n_tile + 1) * c->n_cell * sizeof(uint32_t);
    if (bytes > ((size_t)6 << 30)) {      // too large to be worth it: the kernels fall back to searching
        c->tile_ptr_valid = false;
        return 0;
    }
    if (c->tile_ptr_valid) return 0;
    SCDRS_TRY(c->tile_ptr.reserve(bytes));
    int64_t grid = (c->n_cell + 7) / 8;
    if (grid > kNumSM * 8) grid = kNumSM * 8;
    tile_ptr_kernel<<<(unsigned)grid, 256, 0, c->stream>>>(c->n_cell, n_tile, c->indptr, c->indices,
                                                          c->tile_ptr.as<uint32_t>());
    SCDRS_LAUNCH_CHECK(c);
    c->tile_ptr_valid = true;
    return 0;
}

// Fill rows [0, nr) of a dense tile (cells r0.., genes g0 .. g0 + kGeneTile) of X from the CSR rows; the
// calling threads (tid of nthreads) take rows round-robin.  The loads of a row segment are issued four
// at a time (their addresses are known up front), so a segment costs one memory round trip, not one per
// stored entry.
__device__ __forceinline__ void stage_rows(float (*buf)[kGeneTile + 1], int tid, int nthreads, int64_t r0, int nr,
                                           int g0, int tile_id, int64_t n_cell,
                                           const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                           const float* __restrict__ data, const uint32_t* __restrict__ tptr) {
    for (int rr = tid; rr < nr; rr += nthreads) {
        const int64_t row = r0 + rr;
        const int64_t p0 = indptr[row];
        int64_t lo, hi;
        if (tptr != nullptr) {
            lo = p0 + tptr[(size_t)tile_id * n_cell + row];
            hi = p0 + tptr[(size_t)(tile_id + 1) * n_cell + row];
        } else {
            const int64_t p1 = indptr[row + 1];
            lo = p0;
            hi = p1;
            while (lo < hi) {                     // first stored column >= g0
                const int64_t mid = (lo + hi) >> 1;
                if (indices[mid] < g0) lo = mid + 1; else hi = mid;
            }
            hi = p1;
        }
        bool more = true;
        for (int64_t t = lo; t < hi && more; t += 4) {
            int col[4];
            float val[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool in = t + u < hi;
                col[u] = in ? indices[t + u] - g0 : kGeneTile;
                val[u] = in ? data[t + u] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (col[u] < kGeneTile) buf[rr][col[u]] = val[u];
                else more = false;                // past the tile (only reachable without the table) or past hi
            }
        }
    }
}

// Producer side of the two-group kernels: the threads [kGeneTile, 2 kGeneTile) fill one tile.
__device__ __forceinline__ void stage_tile(float (*buf)[kGeneTile + 1], int ptid, int64_t r0, int nr, int g0,
                                           int tile_id, int64_t n_cell, const int64_t* __restrict__ indptr,
                                           const int32_t* __restrict__ indices, const float* __restrict__ data,
                                           const uint32_t* __restrict__ tptr) {
    for (int i = ptid; i < kRowTile * (kGeneTile + 1); i += kGeneTile) (&buf[0][0])[i] = 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(kGeneTile));      // producers only
    stage_rows(buf, ptid, kGeneTile, r0, nr, g0, tile_id, n_cell, indptr, indices, data, tptr);
}

// Work split inside a CTA (kGeneTile genes): the transform T(x + C.B + mu) of every (cell, gene) pair is
// independent and goes to the "transform" warps, which also stage the tiles of X; only the summation
// has to be sequential, and the "owner" warp (one thread per gene) does nothing else: one shared
// load and one add per cell.  Tiles of kSeqRows cells are double-buffered between the two groups.
constexpr int kSeqRows = 32;
constexpr int kSeqThreads = 128;
constexpr int kSeqXform = kSeqThreads - kGeneTile;   // transform threads

template <int TRANSFORM>
__global__ void __launch_bounds__(kSeqThreads)
stats_gene_seq_kernel(int64_t n_cell, int n_gene, int k, int center, double denom,
                      const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                      const float* __restrict__ data, const uint32_t* __restrict__ tptr,
                      const double* __restrict__ cov_mat,
                      const double* __restrict__ cov_beta, const double* __restrict__ gene_mean,
                      const double* __restrict__ cell_w, const int32_t* __restrict__ tile_ids,
                      const double* __restrict__ mean_in, double* __restrict__ gsum,
                      double* __restrict__ gsq) {
    extern __shared__ double sdyn[];                 // [k][kGeneTile] covariate effects, then [kGeneTile] mu
    __shared__ float xt[kSeqRows][kGeneTile + 1];
    __shared__ double tt[2][kSeqRows][kGeneTile];
    const bool owner = threadIdx.x < kGeneTile;
    // all gene tiles, or the listed ones (the exact-order recomputation of a few genes)
    const int tile_id = tile_ids != nullptr ? tile_ids[blockIdx.x] : (int)blockIdx.x;
    const int g0 = tile_id * kGeneTile;
    const int g = g0 + (int)threadIdx.x;
    const bool ok = owner && g < n_gene;
    double* smu = sdyn + (size_t)k * kGeneTile;
    if (owner) {
        for (int kk = 0; kk < k; ++kk) sdyn[kk * kGeneTile + threadIdx.x] = g < n_gene ? cov_beta[(size_t)g * k + kk] : 0.0;
        smu[threadIdx.x] = (g < n_gene && k > 0) ? gene_mean[g] : 0.0;
    }
    __syncthreads();
    const int xid = (int)threadIdx.x - kGeneTile;    // 0 .. kSeqXform-1 for the transform threads
    const int64_t n_rt = (n_cell + kSeqRows - 1) / kSeqRows;

    // transform threads: stage the X tile of row tile rt, then write T(v) for all its pairs into tt[buf]
    auto produce = [&](int64_t rt, int buf) {
        const int64_t r0 = rt * kSeqRows;
        const int nr = (int)((n_cell - r0) < kSeqRows ? (n_cell - r0) : kSeqRows);
        for (int i = xid; i < kSeqRows * (kGeneTile + 1); i += kSeqXform) (&xt[0][0])[i] = 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(kSeqXform));
        stage_rows(xt, xid, kSeqXform, r0, nr, g0, tile_id, n_cell, indptr, indices, data, tptr);
        asm volatile("bar.sync 1, %0;" ::"n"(kSeqXform));
        for (int p = xid; p < nr * kGeneTile; p += kSeqXform) {
            const int r = p / kGeneTile, gi = p % kGeneTile;
            double v = (double)xt[r][gi];
            if (k > 0) {
                const double* crow = cov_mat + (r0 + r) * k;
                double cb = crow[0] * sdyn[gi];
                for (int kk = 1; kk < k; ++kk) cb = fma(crow[kk], sdyn[kk * kGeneTile + gi], cb);
                v = (v + cb) + smu[gi];
            }
            tt[buf][r][gi] = apply_t<TRANSFORM>(v);
        }
    };

    // mean_in: the centred pass alone, around a mean supplied by the caller (cells sharded over several GPUs: the
    // mean is that of all shards)
    const int n_pass = center ? 2 : 1;
    double s = 0.0, q = 0.0, mean = (mean_in != nullptr && ok) ? mean_in[g] : 0.0;
    for (int pass = (mean_in != nullptr && center) ? 1 : 0; pass < n_pass; ++pass) {
        if (!owner) produce(0, 0);
        __syncthreads();
        for (int64_t rt = 0; rt < n_rt; ++rt) {
            if (!owner) {
                if (rt + 1 < n_rt) produce(rt + 1, (int)((rt + 1) & 1));
            } else if (ok) {
                const int64_t r0 = rt * kSeqRows;
                const int nr = (int)((n_cell - r0) < kSeqRows ? (n_cell - r0) : kSeqRows);
                const double (*cur)[kGeneTile] = tt[rt & 1];
                // strictly in cell order; explicit _rn intrinsics: no fma contraction, numpy multiplies then adds
                if (pass == 0) {
                    if (cell_w == nullptr) {
                        for (int r = 0; r < nr; ++r) {
                            const double t = cur[r][threadIdx.x];
                            s = __dadd_rn(s, t);
                            if (!center) q = __dadd_rn(q, __dmul_rn(t, t));
                        }
                    } else {
                        for (int r = 0; r < nr; ++r) {
                            const double t = cur[r][threadIdx.x], w = cell_w[r0 + r];
                            s = __dadd_rn(s, __dmul_rn(t, w));
                            if (!center) q = __dadd_rn(q, __dmul_rn(__dmul_rn(t, t), w));
                        }
                    }
                } else {
                    for (int r = 0; r < nr; ++r) {
                        const double d = __dadd_rn(cur[r][threadIdx.x], -mean);
                        const double d2 = __dmul_rn(d, d);
                        q = __dadd_rn(q, cell_w != nullptr ? __dmul_rn(d2, cell_w[r0 + r]) : d2);
                    }
                }
            }
            __syncthreads();
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

__global__ void __launch_bounds__(kStatThreadsSeq)
gene_mean_xtc_kernel(int64_t n_cell, int n_gene, int k, float inv_n,
                     const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                     const float* __restrict__ data, const uint32_t* __restrict__ tptr,
                     const double* __restrict__ cov_mat,
                     float* __restrict__ gmean32, double* __restrict__ xtc) {
    __shared__ float tile[2][kRowTile][kGeneTile + 1];
    const bool producer = threadIdx.x >= kGeneTile;
    const int tid = threadIdx.x & (kGeneTile - 1);
    const int g0 = blockIdx.x * kGeneTile;
    const int g = g0 + tid;
    const bool ok = !producer && g < n_gene;
    float s32 = 0.f;
    double acc[kMaxCovReg];
#pragma unroll
    for (int kk = 0; kk < kMaxCovReg; ++kk) acc[kk] = 0.0;
    const int64_t n_rt = (n_cell + kRowTile - 1) / kRowTile;
    if (producer)
        stage_tile(tile[0], tid, 0, (int)(n_cell < kRowTile ? n_cell : kRowTile), g0, blockIdx.x, n_cell, indptr,
                   indices, data, tptr);
    __syncthreads();
    for (int64_t rt = 0; rt < n_rt; ++rt) {
        const int64_t r0 = rt * kRowTile;
        const int nr = (int)((n_cell - r0) < kRowTile ? (n_cell - r0) : kRowTile);
        if (producer) {
            if (rt + 1 < n_rt) {
                const int64_t r1 = r0 + kRowTile;
                stage_tile(tile[(rt + 1) & 1], tid, r1, (int)((n_cell - r1) < kRowTile ? (n_cell - r1) : kRowTile), g0,
                           blockIdx.x, n_cell, indptr, indices, data, tptr);
            }
        } else if (ok) {
            float (*cur)[kGeneTile + 1] = tile[rt & 1];
            for (int r = 0; r < nr; ++r) {
                const float x = cur[r][tid];
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
        __syncthreads();
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
    SCDRS_TRY(ensure_tile_ptr(c));
    const int tiles = (c->n_gene + kGeneTile - 1) / kGeneTile;
    gene_mean_xtc_kernel<<<tiles, kStatThreadsSeq, 0, c->stream>>>(
        c->n_cell, c->n_gene, k, (float)(1.0 / (double)c->n_cell), c->indptr, c->indices, c->data,
        c->tile_ptr_valid ? c->tile_ptr.as<uint32_t>() : nullptr, dev_cov, dev_gmean32, dev_xtc);
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

// ---- dense part of the per-cell statistics by algebra (identity transform) ---------------------------
// sum_g b = C_c . (sum_g B_g) + sum_g mu_g;  sum_g b^2 = C_c' (sum_g B_g B_g') C_c + 2 C_c . (sum_g mu_g B_g) + sum_g mu_g^2
// side[]: [k] sum B, [k*k] sum B B', [k] sum mu B, [1] sum mu, [1] sum mu^2 -- one block, fixed order.
__global__ void __launch_bounds__(256)
gene_side_sums_kernel(int n_gene, int k, const double* __restrict__ cov_beta, const double* __restrict__ gene_mean,
                      double* __restrict__ side) {
    __shared__ double sh[256];
    const int n_q = k + k * k + k + 2;
    for (int q = 0; q < n_q; ++q) {
        double part = 0.0;
        for (int g = threadIdx.x; g < n_gene; g += 256) {
            const double* b = cov_beta + (size_t)g * k;
            const double mu = gene_mean[g];
            double v;
            if (q < k) v = b[q];
            else if (q < k + k * k) v = b[(q - k) / k] * b[(q - k) % k];
            else if (q < 2 * k + k * k) v = mu * b[q - k - k * k];
            else if (q == 2 * k + k * k) v = mu;
            else v = mu * mu;
            part += v;
        }
        sh[threadIdx.x] = part;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) side[q] = sh[0];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
stats_dense_cell_alg_kernel(int64_t n_cell, int k, const double* __restrict__ cov_mat, const double* __restrict__ side,
                            double* __restrict__ csum, double* __restrict__ csq) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n_cell) return;
    const double* c = cov_mat + row * k;
    const double* s_b = side;
    const double* s_bb = side + k;
    const double* s_mub = side + k + k * k;
    double s = side[2 * k + k * k], q = side[2 * k + k * k + 1];
    for (int i = 0; i < k; ++i) {
        s = fma(c[i], s_b[i], s);
        double row_i = 2.0 * s_mub[i];
        for (int j = 0; j < k; ++j) row_i = fma(s_bb[i * k + j], c[j], row_i);
        q = fma(c[i], row_i, q);
    }
    csum[row] = s;
    csq[row] = q;
}

template <int TRANSFORM>
static int run_stats(scdrs_ctx* c, const double* dev_cell_w, int center, double denom, double* gsum,
                     double* gsq, double* csum, double* csq, int n_tile_sel = 0, const int32_t* dev_tile_ids = nullptr,
                     const double* dev_mean_in = nullptr) {
    const int k = c->k, n_gene = c->n_gene;
    const int64_t n_cell = c->n_cell;
    if (gsum != nullptr) {
        const int tiles = (n_gene + kGeneTile - 1) / kGeneTile;
        const size_t smem = (size_t)((k > 0 ? k : 1) + 1) * kGeneTile * sizeof(double);
        SCDRS_CHECK(smem <= 150 * 1024, "too many covariates (%d) for the statistics kernel", k);
        SCDRS_CUDA(cudaFuncSetAttribute(stats_gene_seq_kernel<TRANSFORM>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024));
        // a few listed tiles: rows are searched instead of building the (n_tile + 1) x n_cell table of tile starts
        if (dev_tile_ids == nullptr) SCDRS_TRY(ensure_tile_ptr(c));
        const bool table = dev_tile_ids == nullptr && c->tile_ptr_valid;
        stats_gene_seq_kernel<TRANSFORM><<<dev_tile_ids != nullptr ? n_tile_sel : tiles, kSeqThreads, smem, c->stream>>>(
            n_cell, n_gene, k, center, denom, c->indptr, c->indices, c->data,
            table ? c->tile_ptr.as<uint32_t>() : nullptr, c->cov_mat.as<double>(),
            c->cov_beta.as<double>(), c->gene_mean.as<double>(), dev_cell_w, dev_tile_ids, dev_mean_in, gsum, gsq);
        SCDRS_LAUNCH_CHECK(c);
    }
    if (csum != nullptr) {
        int64_t grid = (n_cell + 7) / 8;
        if (grid > kNumSM * 8) grid = kNumSM * 8;
        if (grid < 1) grid = 1;
        if (k > 0 && TRANSFORM == 0) {
            SCDRS_TRY(c->scan_tmp.reserve((size_t)(2 * k + k * k + 2) * sizeof(double)));
            gene_side_sums_kernel<<<1, 256, 0, c->stream>>>(n_gene, k, c->cov_beta.as<double>(), c->gene_mean.as<double>(),
                                                            c->scan_tmp.as<double>());
            SCDRS_LAUNCH_CHECK(c);
            stats_dense_cell_alg_kernel<<<(unsigned)((n_cell + 255) / 256), 256, 0, c->stream>>>(
                n_cell, k, c->cov_mat.as<double>(), c->scan_tmp.as<double>(), csum, csq);
            SCDRS_LAUNCH_CHECK(c);
        } else if (k > 0) {
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

// the sequential (numpy-order) per-gene sums for the listed gene tiles only
int gene_stats_exact(scdrs_ctx* c, int transform, const double* dev_cell_w, int center, double denom, int n_tile_sel,
                     const int32_t* dev_tile_ids, const double* dev_mean_in, double* dev_gsum, double* dev_gsumsq) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    SCDRS_CHECK(transform == 0 || transform == 1, "transform must be 0 (identity) or 1 (expm1)");
    SCDRS_CHECK(n_tile_sel >= 1 && dev_tile_ids != nullptr, "no gene tiles listed");
    if (transform == 0)
        return run_stats<0>(c, dev_cell_w, center, denom, dev_gsum, dev_gsumsq, nullptr, nullptr, n_tile_sel, dev_tile_ids, dev_mean_in);
    return run_stats<1>(c, dev_cell_w, center, denom, dev_gsum, dev_gsumsq, nullptr, nullptr, n_tile_sel, dev_tile_ids, dev_mean_in);
}

int gene_tile_size() { return kGeneTile; }

}  // namespace scdrs
