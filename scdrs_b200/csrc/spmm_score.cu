// K1: raw scores of the trait gene set and its n_ctrl control gene sets for every cell.
//
// Replaces the 1 + n_ctrl sequential calls of _compute_raw_score (scdrs/method.py:172-183,
// 312-402 of the reference): raw[c, j] = sum_g X[c, g] * W[g, j]  (+ COV_MAT[c,:] . M[:, j] + m[j]
// in implicit-covariate mode), where column j of W holds the normalised weights of gene set j.
//
// Layout.  X stays CSR (cells-major, the way cells are sharded across GPUs).  The weight
//   w[g, j] = base_g * gw / s_j        (scdrs/method.py:366-375; base = 1, 1/sqrt(var_tech + .01) ...)
// is factored: base_g is folded into the expression value once per stored entry, 1/s_j is applied
// once per output column, and only gw (the .gs gene weight; 1 for binary traits, then nothing at
// all) is kept per (gene, set) pair.  W is stored gene-major: for gene g the list of set ids that
// contain g.  One warp owns one cell: it walks the cell's non-zeros once and, lanes splitting one
// gene's list, adds into 1 + n_ctrl fp64 accumulators in shared memory (no atomics: the set ids
// of one gene are distinct).  The kernel is bound by the shared-memory pipe (one 64-bit load and
// one 64-bit store per multiply-add), so each gene's list is laid out to make those accesses
// conflict-free: it is padded to a multiple of 16 slots and its entries are dealt into groups of
// 16 slots (one shared-memory wavefront each) such that a group holds as few entries of the same
// bank pair (set id mod 16) as possible.
//
// Accumulation is fp64 like the reference's (float32 data times float64 weights,
// scdrs/method.py:392-400).  The result is written as an fp64 per-cell reference value plus fp32
// deviations from it, NaN marking exact zeros (the `raw == 0` rule, scdrs/method.py:522-523).
#include <math.h>

#include <utility>

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

__device__ __forceinline__ double base_weight(int weight_opt, const double* var, const double* var_tech, int g) {
    if (weight_opt == SCDRS_WEIGHT_VS) return 1.0 / sqrt(var_tech[g] + 1e-2);
    if (weight_opt == SCDRS_WEIGHT_INV_STD) return 1.0 / sqrt(var[g] + 1e-2);
    return 1.0;
}

__global__ void gene_base_kernel(int n_gene, int weight_opt, const double* __restrict__ var,
                                 const double* __restrict__ var_tech, double* __restrict__ gbase) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_gene) gbase[g] = base_weight(weight_opt, var, var_tech, g);
}

// One block per gene set: s_j, the per-entry value the scatter multiplies with, the covariate
// projections M[:, j] = COV_BETA[S_j]^T w_j and m[j] = COV_GENE_MEAN[S_j] . w_j (:393-394) and the
// variance-ratio numerator sum_g var_g w_g^2 (:191-195), w being the normalised weights.
__global__ void prep_sets_kernel(int n_s0, int n_sc, int ncol, int k, int binary,
                                 const int32_t* __restrict__ set_genes,
                                 const double* __restrict__ gw,  // [n_s0] trait, then [n_sc] control
                                 const double* __restrict__ explicit_w,
                                 const double* __restrict__ gbase, const double* __restrict__ var,
                                 const double* __restrict__ beta, const double* __restrict__ gmean,
                                 float* __restrict__ set_val, double* __restrict__ colscale,
                                 double* __restrict__ covM, double* __restrict__ covm,
                                 double* __restrict__ rnum, const double* __restrict__ xsum,
                                 const double* __restrict__ cov_colsum, double n_cell_d,
                                 double* __restrict__ colsum_alg) {
    __shared__ double sh[32];
    const int j = blockIdx.x;
    const int n_s = j == 0 ? n_s0 : n_sc;
    const size_t off = j == 0 ? 0 : (size_t)n_s0 + (size_t)(j - 1) * n_sc;
    const int32_t* genes = set_genes + off;
    float* vout = set_val + off;
    const double* g = explicit_w != nullptr ? explicit_w + off : gw + (j == 0 ? 0 : n_s0);
    double s = 1.0;
    if (explicit_w == nullptr) {
        double part = 0.0;
        for (int p = threadIdx.x; p < n_s; p += blockDim.x) part += gbase[genes[p]] * g[p];
        s = block_sum(part, sh);
    }
    // binary sets scatter the constant 1 and fold the common gene weight into the column scale
    const double cs = binary ? (n_s > 0 ? g[0] : 1.0) / s : 1.0 / s;
    double pm = 0.0, pr = 0.0, px = 0.0;
    for (int p = threadIdx.x; p < n_s; p += blockDim.x) {
        const int gi = genes[p];
        const double w = gbase[gi] * g[p] / s;                          // the reference's weight
        const float val = binary ? 1.0f : (float)g[p];
        vout[p] = val;
        const double wf = (double)val * gbase[gi] * cs;                 // the weight actually applied
        if (k > 0) pm += gmean[gi] * wf;
        if (var != nullptr) pr += var[gi] * w * w;
        if (xsum != nullptr) px += xsum[gi] * wf;
    }
    pm = block_sum(pm, sh);
    pr = block_sum(pr, sh);
    px = block_sum(px, sh);
    if (threadIdx.x == 0) {
        covm[j] = pm;
        rnum[j] = pr;
        colscale[j] = cs;
    }
    // column sum of the raw scores of this set over all cells, without touching the score matrix:
    // sum_c raw[c, j] = (sum_c X[c, :]) . w_j + (sum_c C[c, :]) . M[:, j] + n_cell m[j]
    double total = px + n_cell_d * pm;
    for (int kk = 0; kk < k; ++kk) {
        double pb = 0.0;
        for (int p = threadIdx.x; p < n_s; p += blockDim.x)
            pb += beta[(size_t)genes[p] * k + kk] * ((double)vout[p] * gbase[genes[p]] * cs);
        pb = block_sum(pb, sh);
        if (threadIdx.x == 0) covM[(size_t)kk * ncol + j] = pb;
        if (cov_colsum != nullptr) total += cov_colsum[kk] * pb;
    }
    if (threadIdx.x == 0 && colsum_alg != nullptr) colsum_alg[j] = total;
}

// ---- per-matrix sums behind the algebraic column sums -------------------------------------------------
// sum_c X[c, g] in 2^-24 fixed point (integer atomics: exact and independent of the order)
__global__ void __launch_bounds__(256)
xsum_kernel(int64_t nnz, const int32_t* __restrict__ indices, const float* __restrict__ data,
            unsigned long long* __restrict__ acc) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&acc[indices[q]], (unsigned long long)__double2ll_rn((double)data[q] * 16777216.0));
}
__global__ void xsum_finish_kernel(int n_gene, const unsigned long long* __restrict__ acc, double* __restrict__ xsum) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_gene) xsum[g] = (double)(long long)acc[g] * (1.0 / 16777216.0);
}
// sum_c COV_MAT[c, kk], one block per covariate, fixed order
__global__ void __launch_bounds__(256)
cov_colsum_kernel(int64_t n_cell, int k, const double* __restrict__ cov_mat, double* __restrict__ out) {
    __shared__ double sh[32];
    const int kk = blockIdx.x;
    double part = 0.0;
    for (int64_t r = threadIdx.x; r < n_cell; r += blockDim.x) part += cov_mat[r * k + kk];
    part = block_sum(part, sh);
    if (threadIdx.x == 0) out[kk] = part;
}

static int ensure_matrix_sums(scdrs_ctx* c) {
    if (!c->xsum_valid) {
        SCDRS_TRY(c->gene_xsum.reserve((size_t)c->n_gene * sizeof(double)));
        SCDRS_TRY(c->stage.reserve((size_t)c->n_gene * sizeof(unsigned long long)));
        SCDRS_CUDA(cudaMemsetAsync(c->stage.p, 0, (size_t)c->n_gene * sizeof(unsigned long long), c->stream));
        if (c->nnz > 0) {
            xsum_kernel<<<kNumSM * 8, 256, 0, c->stream>>>(c->nnz, c->indices, c->data, c->stage.as<unsigned long long>());
            SCDRS_LAUNCH_CHECK(c);
        }
        xsum_finish_kernel<<<(c->n_gene + 255) / 256, 256, 0, c->stream>>>(c->n_gene, c->stage.as<unsigned long long>(),
                                                                          c->gene_xsum.as<double>());
        SCDRS_LAUNCH_CHECK(c);
        c->xsum_valid = true;
    }
    if (c->k > 0 && !c->covsum_valid) {
        SCDRS_TRY(c->cov_colsum.reserve((size_t)c->k * sizeof(double)));
        cov_colsum_kernel<<<c->k, 256, 0, c->stream>>>(c->n_cell, c->k, c->cov_mat.as<double>(), c->cov_colsum.as<double>());
        SCDRS_LAUNCH_CHECK(c);
        c->covsum_valid = true;
    }
    return 0;
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
__global__ void w_count_kernel(int64_t e_begin, int64_t e_end, const int32_t* __restrict__ set_genes,
                               uint32_t* __restrict__ cnt) {
    for (int64_t e = e_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < e_end;
         e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[set_genes[e]], 1u);
}

__global__ void w_padded_len_kernel(int n_gene, const uint32_t* __restrict__ cnt,
                                    uint32_t* __restrict__ plen) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_gene) plen[g] = (cnt[g] + 63u) & ~63u;
    if (g == n_gene) plen[g] = 0;
}

// Storage index of slot q of a gene's list (group q/16, position q%16).  Lists are stored in blocks of
// 64 slots; inside a block the slots that one lane of the scatter kernel consumes together (slot l of
// groups 0/1 and slot l of groups 2/3) are adjacent, so that the lane fetches both set ids with one
// 32-bit load (and both weights with one 64-bit load).
__host__ __device__ __forceinline__ uint32_t slot_index(uint32_t q) {
    const uint32_t grp = q >> 4, pos = q & 15u;
    return (((grp >> 2) * 32u + (grp & 1u) * 16u + pos) << 1) | ((grp >> 1) & 1u);
}

// GeneInfo (common.cuh): meta = (first 64-slot block << 12) | number of 16-slot groups

__global__ void w_meta_kernel(int n_gene, const uint32_t* __restrict__ w_ptr, const uint32_t* __restrict__ p_ptr,
                              const double* __restrict__ gbase, GeneInfo* __restrict__ info) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_gene) {
        GeneInfo gi;
        gi.base = gbase[g];
        gi.meta = ((p_ptr[g] >> 6) << 12) | ((w_ptr[g + 1] - w_ptr[g] + 15u) >> 4);
        gi.pad = 0;
        info[g] = gi;
    }
}

// entries [e_begin, e_end) of the set-major list; the stored set id is relative to col0
__global__ void w_fill_kernel(int64_t e_begin, int64_t e_end, int n_s0, int n_sc, int col0,
                              const int32_t* __restrict__ set_genes,
                              const float* __restrict__ set_val, const uint32_t* __restrict__ w_ptr,
                              uint32_t* __restrict__ cursor, uint16_t* __restrict__ w_col,
                              float* __restrict__ w_val) {
    for (int64_t e = e_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < e_end;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int g = set_genes[e];
        const uint32_t pos = w_ptr[g] + atomicAdd(&cursor[g], 1u);
        w_col[pos] = (uint16_t)((e < n_s0 ? 0 : 1 + (e - n_s0) / n_sc) - col0);
        w_val[pos] = set_val[e];
    }
}

// padding slots carry the set id 0xffff, which the scatter kernel predicates off
__global__ void w_pad_init_kernel(int64_t n_slot, uint16_t* __restrict__ col, float* __restrict__ val) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_slot;
         i += (int64_t)gridDim.x * blockDim.x) {
        col[i] = (uint16_t)0xffff;
        if (val != nullptr) val[i] = 0.f;
    }
}

// One warp per gene: deal the gene's entries, ordered by bank pair (set id mod 16), round-robin
// over its groups of 16 slots.  Entry number i of that order goes to group i % R, slot i / R,
// R = padded_length / 16, so same-pair entries land in different groups whenever there are at
// most R of them.
__global__ void __launch_bounds__(128)
w_layout_deal_kernel(int n_gene, const uint32_t* __restrict__ w_ptr, const uint32_t* __restrict__ p_ptr,
                const uint16_t* __restrict__ w_col, const float* __restrict__ w_val,
                uint16_t* __restrict__ p_col, float* __restrict__ p_val) {
    __shared__ uint32_t hist[4][16], run[4][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * 4 + warp;
    if (g >= n_gene) return;
    const uint32_t s = w_ptr[g], e = w_ptr[g + 1];
    const uint32_t ps = p_ptr[g], R = (e - s + 15u) >> 4;
    if (e == s || R <= 8) return;   // short lists are placed by w_layout_greedy_kernel
    if (lane < 16) { hist[warp][lane] = 0; run[warp][lane] = 0; }
    __syncwarp();
    for (uint32_t t = s + lane; t < e; t += 32) atomicAdd(&hist[warp][w_col[t] & 15], 1u);
    __syncwarp();
    // exclusive prefix over the 16 pairs
    uint32_t h = lane < 16 ? hist[warp][lane] : 0u, inc = h;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncwarp();
    if (lane < 16) hist[warp][lane] = inc - h;   // now the base of each pair
    __syncwarp();
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t t0 = s; t0 < e; t0 += 32) {
        const uint32_t t = t0 + lane;
        const bool ok = t < e;
        int col = 0, pair = 16 + lane;
        float v = 0.f;
        if (ok) {
            col = w_col[t];
            v = w_val[t];
            pair = col & 15;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, pair);
        const uint32_t rank = __popc(peers & lt);
        uint32_t i = 0;
        if (ok) i = hist[warp][pair] + run[warp][pair] + rank;
        __syncwarp();
        if (ok && rank == 0) run[warp][pair] += __popc(peers);
        __syncwarp();
        if (ok) {
            const uint32_t dst = ps + slot_index((i % R) * 16 + (i / R));
            p_col[dst] = (uint16_t)col;
            if (p_val != nullptr) p_val[dst] = v;
        }
    }
}

// One thread per gene, lists of at most 8 groups (128 entries): greedy placement that minimises the
// number of shared-memory wavefronts, sum over groups of the largest multiplicity of a bank pair (set
// id mod 16) inside the group.  Bank pairs are taken in order of decreasing multiplicity; every entry
// goes to the group where it does not raise the group's maximum, then where its pair is rarest, then
// to the emptiest group.  The list is first counting-sorted into that order in shared memory, then
// placed in one pass; per-group state lives in registers (16 four-bit counters packed in a uint64).
constexpr int kGreedyThreads = 64;
constexpr int kGreedyMaxLen = 128;

__global__ void __launch_bounds__(kGreedyThreads)
w_layout_greedy_kernel(int n_gene, const uint32_t* __restrict__ w_ptr, const uint32_t* __restrict__ p_ptr,
                       const uint16_t* __restrict__ w_col, const float* __restrict__ w_val,
                       uint16_t* __restrict__ p_col, float* __restrict__ p_val) {
    extern __shared__ uint32_t stage[];   // [kGreedyMaxLen][kGreedyThreads]: (offset in list << 16) | set id
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gene) return;
    const uint32_t s = w_ptr[g], e = w_ptr[g + 1];
    const uint32_t ps = p_ptr[g], R = (e - s + 15u) >> 4;
    if (e == s || R > 8) return;
    uint32_t c[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) c[b] = 0;
    for (uint32_t t = s; t < e; ++t) {
        const int b = w_col[t] & 15;
#pragma unroll
        for (int bb = 0; bb < 16; ++bb) c[bb] += (bb == b);
    }
    // start[b] = number of entries whose pair comes before b in (multiplicity desc, pair asc) order
    uint32_t start[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        uint32_t acc = 0;
#pragma unroll
        for (int o = 0; o < 16; ++o)
            if (c[o] > c[b] || (c[o] == c[b] && o < b)) acc += c[o];
        start[b] = acc;
    }
    for (uint32_t t = s; t < e; ++t) {
        const uint32_t col = w_col[t];
        const int b = col & 15;
        uint32_t pos = 0;
#pragma unroll
        for (int bb = 0; bb < 16; ++bb)
            if (bb == b) pos = start[bb]++;
        stage[pos * kGreedyThreads + threadIdx.x] = ((t - s) << 16) | col;
    }
    unsigned long long cnt[8];            // per group: 16 x 4-bit counters (saturating)
    uint32_t tot[8], gmax[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { cnt[r] = 0; tot[r] = 0; gmax[r] = 0; }
    const uint32_t n = e - s;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t ent = stage[i * kGreedyThreads + threadIdx.x];
        const uint32_t col = ent & 0xffffu, off = ent >> 16;
        const int sh = 4 * (int)(col & 15);
        uint32_t best = 0, best_key = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if ((uint32_t)r >= R || tot[r] >= 16) continue;
            const uint32_t cb = (uint32_t)((cnt[r] >> sh) & 15ull);
            const uint32_t key = ((uint32_t)(cb + 1 > gmax[r]) << 16) | (cb << 8) | tot[r];
            if (key < best_key) { best_key = key; best = r; }
        }
        uint32_t slot = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if ((uint32_t)r == best) {
                slot = tot[r]++;
                const uint32_t cb = (uint32_t)((cnt[r] >> sh) & 15ull) + 1;
                if (cb < 16) cnt[r] += 1ull << sh;
                if (cb > gmax[r]) gmax[r] = cb;
            }
        }
        const uint32_t dst = ps + slot_index(best * 16 + slot);
        p_col[dst] = (uint16_t)col;
        if (p_val != nullptr) p_val[dst] = w_val[s + off];
    }
}

static int upload_sets(scdrs_ctx* c, const int32_t* trait_genes, const int32_t* ctrl_genes,
                       const double* gw_trait, const double* gw_ctrl, const double* explicit_w,
                       bool ctrl_on_device = false) {
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
                                       (n_entry - n_s) * sizeof(int32_t),
                                       ctrl_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
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

// Compact gene-major lists of the sets [col0, col1): for gene g the ids (relative to col0) and .gs
// weights of the sets that contain it -> c->w_ptr (offsets), c->w_col, c->w_val; c->w_cursor holds the
// per-gene counts afterwards.
int build_gene_lists(scdrs_ctx* c, int col0, int col1) {
    const int n_s = c->n_s, n_sc = c->n_sc, n_gene = c->n_gene;
    const int64_t e_begin = col0 == 0 ? 0 : (int64_t)n_s + (int64_t)(col0 - 1) * n_sc;
    const int64_t e_end = col1 == 0 ? 0 : (int64_t)n_s + (int64_t)(col1 - 1) * n_sc;
    const int64_t n_entry = e_end - e_begin;
    SCDRS_CUDA(cudaMemsetAsync(c->w_cursor.p, 0, (size_t)(n_gene + 1) * sizeof(uint32_t), c->stream));
    const int nb = (int)((n_entry + 255) / 256 < 1184 ? (n_entry + 255) / 256 : 1184);
    if (n_entry > 0) {
        w_count_kernel<<<nb, 256, 0, c->stream>>>(e_begin, e_end, c->set_genes.as<int32_t>(), c->w_cursor.as<uint32_t>());
        SCDRS_LAUNCH_CHECK(c);
    }
    SCDRS_TRY(exclusive_scan_u32(c, c->w_cursor.as<uint32_t>(), c->w_ptr.as<uint32_t>(), n_gene + 1));
    if (n_entry > 0) {
        SCDRS_TRY(c->stage.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
        SCDRS_CUDA(cudaMemsetAsync(c->stage.p, 0, (size_t)(n_gene + 1) * sizeof(uint32_t), c->stream));
        w_fill_kernel<<<nb, 256, 0, c->stream>>>(e_begin, e_end, n_s, n_sc > 0 ? n_sc : 1, col0, c->set_genes.as<int32_t>(),
                                                 c->set_w.as<float>(), c->w_ptr.as<uint32_t>(), c->stage.as<uint32_t>(),
                                                 c->w_col.as<uint16_t>(), c->w_val.as<float>());
        SCDRS_LAUNCH_CHECK(c);
    }
    return 0;
}

static int build_w(scdrs_ctx* c, int weight_opt, int use_var_ratio, bool explicit_w, bool binary,
                   double gw_absmax) {
    const int ncol = c->ncol, n_s = c->n_s, n_sc = c->n_sc, n_gene = c->n_gene;
    const int64_t n_entry = (int64_t)n_s + (int64_t)c->n_ctrl * n_sc;
    SCDRS_TRY(c->covM.reserve((size_t)(c->k > 0 ? c->k : 1) * ncol * sizeof(double)));
    SCDRS_TRY(c->covm.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->colscale.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->gbase.reserve((size_t)n_gene * sizeof(double)));
    SCDRS_TRY(c->ratio_a.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->rnum.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->w_ptr.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    SCDRS_TRY(c->w_cursor.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    SCDRS_TRY(c->p_ptr.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    SCDRS_TRY(c->p_meta.reserve((size_t)(n_gene + 1) * sizeof(GeneInfo)));
    SCDRS_CHECK(n_entry + 64 * (int64_t)n_gene < ((int64_t)1 << 26), "gene-set lists too large for the packed index");
    SCDRS_CHECK(ncol <= 4095 * 16, "too many gene sets");
    SCDRS_TRY(c->w_col.reserve(n_entry * sizeof(uint16_t)));
    SCDRS_TRY(c->w_val.reserve(n_entry * sizeof(float)));
    c->w_binary = binary;
    c->w_fx = false;
    c->colsum_ready = false;
    if (!explicit_w) {
        SCDRS_TRY(ensure_matrix_sums(c));
        SCDRS_TRY(c->colsum_alg.reserve(ncol * sizeof(double)));
    }

    gene_base_kernel<<<(n_gene + 255) / 256, 256, 0, c->stream>>>(
        n_gene, explicit_w ? SCDRS_WEIGHT_UNIFORM : weight_opt,
        c->have_stats ? c->var.as<double>() : nullptr,
        c->have_stats ? c->var_tech.as<double>() : nullptr, c->gbase.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    double* rnum = c->rnum.as<double>();
    prep_sets_kernel<<<ncol, 256, 0, c->stream>>>(
        n_s, n_sc, ncol, c->k, binary ? 1 : 0, c->set_genes.as<int32_t>(), c->set_gw.as<double>(),
        explicit_w ? c->stage.as<double>() : nullptr, c->gbase.as<double>(),
        c->have_stats ? c->var.as<double>() : nullptr, c->cov_beta.as<double>(),
        c->gene_mean.as<double>(), c->set_w.as<float>(), c->colscale.as<double>(),
        c->covM.as<double>(), c->covm.as<double>(), rnum, explicit_w ? nullptr : c->gene_xsum.as<double>(),
        (!explicit_w && c->k > 0) ? c->cov_colsum.as<double>() : nullptr, (double)c->n_cell,
        explicit_w ? nullptr : c->colsum_alg.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    c->colsum_ready = !explicit_w;
    ratio_kernel<<<(ncol + 255) / 256, 256, 0, c->stream>>>(ncol, use_var_ratio, rnum,
                                                            c->ratio_a.as<double>());
    SCDRS_LAUNCH_CHECK(c);

    // which scatter: fixed point when it applies (spmm_score_fx.cu), else fp64 accumulators
    bool use_fx = fx_applicable(c, ncol, explicit_w);
    if (use_fx) {
        SCDRS_TRY(fx_value_stats(c, weight_opt));
        use_fx = c->fx_vmax > 0.0 && c->fx_vmax <= 128.0 * c->fx_vmean && gw_absmax > 0.0 &&
                 gw_absmax < 1e30;
    }
    SCDRS_CHECK(use_fx || c->spmm_mode != 2,
                "fixed-point scatter requested but not applicable (needs a trait prepared by scdrs_trait_raw, "
                "1 + n_ctrl <= %d and max|x*base| <= 128 mean|x*base|)", kFxMaxCol * kFxMaxTiles);
    if (use_fx) return fx_prepare(c, binary, binary ? 1.0 : gw_absmax);

    SCDRS_TRY(build_gene_lists(c, 0, ncol));
    w_padded_len_kernel<<<(n_gene + 256) / 256, 256, 0, c->stream>>>(n_gene, c->w_cursor.as<uint32_t>(),
                                                                     c->p_ptr.as<uint32_t>());
    SCDRS_LAUNCH_CHECK(c);
    // padded size: every non-empty gene rounds up to a multiple of 64 slots
    const int64_t n_slot_max = n_entry + 63 * (int64_t)(n_gene < n_entry ? n_gene : n_entry) + 64;
    SCDRS_TRY(c->p_col.reserve((size_t)n_slot_max * sizeof(uint16_t)));
    if (!binary) SCDRS_TRY(c->p_val.reserve((size_t)n_slot_max * sizeof(float)));
    SCDRS_TRY(exclusive_scan_u32(c, c->p_ptr.as<uint32_t>(), c->p_ptr.as<uint32_t>(), n_gene + 1));
    w_meta_kernel<<<(n_gene + 255) / 256, 256, 0, c->stream>>>(n_gene, c->w_ptr.as<uint32_t>(), c->p_ptr.as<uint32_t>(), c->gbase.as<double>(),
                                                               c->p_meta.as<GeneInfo>());
    SCDRS_LAUNCH_CHECK(c);
    const int nb = (int)((n_slot_max + 255) / 256 < 1184 ? (n_slot_max + 255) / 256 : 1184);
    w_pad_init_kernel<<<nb, 256, 0, c->stream>>>(n_slot_max, c->p_col.as<uint16_t>(),
                                                 binary ? nullptr : c->p_val.as<float>());
    SCDRS_LAUNCH_CHECK(c);
    w_layout_deal_kernel<<<(n_gene + 3) / 4, 128, 0, c->stream>>>(
        n_gene, c->w_ptr.as<uint32_t>(), c->p_ptr.as<uint32_t>(), c->w_col.as<uint16_t>(),
        c->w_val.as<float>(), c->p_col.as<uint16_t>(), binary ? nullptr : c->p_val.as<float>());
    SCDRS_LAUNCH_CHECK(c);
    static bool greedy_attr = false;
    if (!greedy_attr) {
        SCDRS_CUDA(cudaFuncSetAttribute(w_layout_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kGreedyMaxLen * kGreedyThreads * (int)sizeof(uint32_t)));
        greedy_attr = true;
    }
    w_layout_greedy_kernel<<<(n_gene + kGreedyThreads - 1) / kGreedyThreads, kGreedyThreads,
                             kGreedyMaxLen * kGreedyThreads * sizeof(uint32_t), c->stream>>>(
        n_gene, c->w_ptr.as<uint32_t>(), c->p_ptr.as<uint32_t>(), c->w_col.as<uint16_t>(),
        c->w_val.as<float>(), c->p_col.as<uint16_t>(), binary ? nullptr : c->p_val.as<float>());
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

static int set_shape(scdrs_ctx* c, int32_t ncol, int32_t n_s, int32_t n_sc) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    SCDRS_CHECK(ncol >= 1 && ncol <= 65000, "1 + n_ctrl = %d outside [1, 65000]", ncol);
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

// Runs a trait preparation on the context's second stream (see scdrs_ctx::prep_stream): everything launched inside
// the scope goes to that stream and uses that stream's scratch; ratio_a / colsum_alg switch to the buffers the
// already enqueued tail of the previous trait does not read.  The preparation starts when the previous scatter has
// finished -- or, when the inputs may have changed since or the control sets are produced on the main stream, behind
// everything enqueued so far -- and the main stream continues behind it.
namespace {
struct PrepScope {
    scdrs_ctx* c;
    bool on;
    int rc = 0;
    PrepScope(scdrs_ctx* ctx, bool after_whole_stream) : c(ctx), on(ctx->prep_overlap && ctx->prep_stream != nullptr) {
        std::swap(c->ratio_a, c->ratio_a_alt);
        std::swap(c->colsum_alg, c->colsum_alg_alt);
        if (!on) return;
        cudaEvent_t after = c->ev_scatter;
        if (after_whole_stream || c->inputs_dirty || !c->scatter_marked) {
            after = c->ev_mark;
            if (cudaEventRecord(after, c->stream) != cudaSuccess) rc = 1;
        }
        if (rc == 0 && cudaStreamWaitEvent(c->prep_stream, after, 0) != cudaSuccess) rc = 1;
        c->inputs_dirty = false;
        std::swap(c->stream, c->prep_stream);
        std::swap(c->stage, c->stage_prep);
        std::swap(c->scan_tmp, c->scan_tmp_prep);
        if (c->tl.on) cudaEventRecord(c->tl.prep0[c->tl.n & 15], c->stream);
    }
    ~PrepScope() {
        if (!on) return;
        if (c->tl.on) cudaEventRecord(c->tl.prep1[c->tl.n & 15], c->stream);
        std::swap(c->stream, c->prep_stream);
        std::swap(c->stage, c->stage_prep);
        std::swap(c->scan_tmp, c->scan_tmp_prep);
        // the main stream continues behind the preparation (also after an error: nothing may overtake it)
        if (cudaEventRecord(c->ev_prep, c->prep_stream) == cudaSuccess) cudaStreamWaitEvent(c->stream, c->ev_prep, 0);
    }
};
}  // namespace

int trait_prepare(scdrs_ctx* c, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                  const int32_t* trait_genes, const double* trait_gw, const int32_t* ctrl_genes,
                  const double* ctrl_gw, int32_t weight_opt, int32_t use_var_ratio, bool ctrl_on_device) {
    PrepScope scope(c, ctrl_on_device && !c->resident_sets);
    SCDRS_CHECK(scope.rc == 0, "could not order the trait preparation behind the main stream");
    SCDRS_CHECK(n_ctrl >= 0, "n_ctrl < 0");
    SCDRS_CHECK(weight_opt >= 0 && weight_opt <= 2, "illegal weight_opt %d", weight_opt);
    SCDRS_CHECK(weight_opt == SCDRS_WEIGHT_UNIFORM || c->have_stats,
                "weight_opt needs gene statistics (call scdrs_set_gene_stats)");
    SCDRS_CHECK(!use_var_ratio || c->have_stats, "variance ratio needs gene statistics");
    SCDRS_TRY(set_shape(c, n_ctrl + 1, n_s, n_s_ctrl));
    for (int i = 0; i < n_s; ++i) SCDRS_CHECK(trait_genes[i] >= 0 && trait_genes[i] < c->n_gene, "gene index out of range");
    // binary gene set: every gene weight identical -> the per-entry value is dropped
    bool binary = true;
    for (int i = 1; i < n_s && binary; ++i) binary = trait_gw[i] == trait_gw[0];
    for (int i = 0; i < n_s_ctrl && binary && n_ctrl > 0; ++i) binary = ctrl_gw[i] == trait_gw[0];
    double gw_absmax = 0.0;
    for (int i = 0; i < n_s; ++i) gw_absmax = fabs(trait_gw[i]) > gw_absmax ? fabs(trait_gw[i]) : gw_absmax;
    for (int i = 0; i < n_s_ctrl && n_ctrl > 0; ++i) gw_absmax = fabs(ctrl_gw[i]) > gw_absmax ? fabs(ctrl_gw[i]) : gw_absmax;
    SCDRS_TRY(upload_sets(c, trait_genes, ctrl_genes, trait_gw, ctrl_gw, nullptr, ctrl_on_device));
    return build_w(c, weight_opt, use_var_ratio, false, binary, gw_absmax);
}

int explicit_prepare(scdrs_ctx* c, int32_t n_set, int32_t n_s, const int32_t* genes,
                     const double* w) {
    SCDRS_TRY(set_shape(c, n_set, n_s, n_s));
    SCDRS_TRY(upload_sets(c, nullptr, genes, nullptr, nullptr, w));
    return build_w(c, SCDRS_WEIGHT_UNIFORM, 0, true, false, 1.0);
}

// ---- the scatter kernel ----------------------------------------------------------------------
// Padding slots carry the set id 0xffff and are predicated off, so they touch no bank.
__device__ __forceinline__ void acc_add(uint32_t acc_s, uint32_t col, double x) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f64 v;\n\t.reg .u32 a;\n\t"
        "setp.ne.u32 p, %1, 0xffff;\n\t"
        "mad.lo.u32 a, %1, 8, %0;\n\t"
        "mov.f64 v, 0d0000000000000000;\n\t"
        "@p ld.shared.f64 v, [a];\n\t"
        "add.f64 v, v, %2;\n\t"
        "@p st.shared.f64 [a], v;\n\t}"
        ::"r"(acc_s), "r"(col), "d"(x)
        : "memory");
}
__device__ __forceinline__ void acc_fma(uint32_t acc_s, uint32_t col, double x, float w) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f64 v, wd;\n\t.reg .u32 a;\n\t"
        "setp.ne.u32 p, %1, 0xffff;\n\t"
        "mad.lo.u32 a, %1, 8, %0;\n\t"
        "cvt.f64.f32 wd, %3;\n\t"
        "mov.f64 v, 0d0000000000000000;\n\t"
        "@p ld.shared.f64 v, [a];\n\t"
        "fma.rn.f64 v, %2, wd, v;\n\t"
        "@p st.shared.f64 [a], v;\n\t}"
        ::"r"(acc_s), "r"(col), "d"(x), "f"(w)
        : "memory");
}

// Two independent read-modify-writes (the set ids of one gene are distinct, so the addresses
// differ): loads first, then adds, then stores, to overlap the shared-memory latencies.
__device__ __forceinline__ void acc_add2(uint32_t acc_s, uint32_t both, double x) {
    const uint32_t q0 = both & 0xffffu, q1 = both >> 16;
    asm volatile(
        "{\n\t.reg .pred p0, p1;\n\t.reg .f64 v0, v1;\n\t.reg .u32 a0, a1;\n\t"
        "setp.ne.u32 p0, %1, 0xffff;\n\tsetp.ne.u32 p1, %2, 0xffff;\n\t"
        "mad.lo.u32 a0, %1, 8, %0;\n\tmad.lo.u32 a1, %2, 8, %0;\n\t"
        "mov.f64 v0, 0d0000000000000000;\n\tmov.f64 v1, v0;\n\t"
        "@p0 ld.shared.f64 v0, [a0];\n\t@p1 ld.shared.f64 v1, [a1];\n\t"
        "add.f64 v0, v0, %3;\n\tadd.f64 v1, v1, %3;\n\t"
        "@p0 st.shared.f64 [a0], v0;\n\t@p1 st.shared.f64 [a1], v1;\n\t}"
        ::"r"(acc_s), "r"(q0), "r"(q1), "d"(x)
        : "memory");
}
__device__ __forceinline__ void acc_fma2(uint32_t acc_s, uint32_t both, double x, float w0, float w1) {
    const uint32_t q0 = both & 0xffffu, q1 = both >> 16;
    asm volatile(
        "{\n\t.reg .pred p0, p1;\n\t.reg .f64 v0, v1, d0, d1;\n\t.reg .u32 a0, a1;\n\t"
        "setp.ne.u32 p0, %1, 0xffff;\n\tsetp.ne.u32 p1, %2, 0xffff;\n\t"
        "mad.lo.u32 a0, %1, 8, %0;\n\tmad.lo.u32 a1, %2, 8, %0;\n\t"
        "cvt.f64.f32 d0, %4;\n\tcvt.f64.f32 d1, %5;\n\t"
        "mov.f64 v0, 0d0000000000000000;\n\tmov.f64 v1, v0;\n\t"
        "@p0 ld.shared.f64 v0, [a0];\n\t@p1 ld.shared.f64 v1, [a1];\n\t"
        "fma.rn.f64 v0, %3, d0, v0;\n\tfma.rn.f64 v1, %3, d1, v1;\n\t"
        "@p0 st.shared.f64 [a0], v0;\n\t@p1 st.shared.f64 [a1], v1;\n\t}"
        ::"r"(acc_s), "r"(q0), "r"(q1), "d"(x), "f"(w0), "f"(w1)
        : "memory");
}

// Lane l of the warp consumes slots l and l + 32 of every 64-slot block of a gene's list: lanes 0-15
// groups 0 and 2, lanes 16-31 groups 1 and 3; lists longer than one block (rare) continue in a loop.
// One warp per cell.  p_meta[g] = (first 16-slot group of gene g's list << 12) | number of groups.
template <bool WEIGHTED>
__global__ void __launch_bounds__(256, 4)
spmm_score_kernel(int64_t n_cell, int ncol, int ld, int acc_stride,
                  const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const float* __restrict__ data, const GeneInfo* __restrict__ ginfo,
                  const uint16_t* __restrict__ p_col,
                  const float* __restrict__ p_val, const double* __restrict__ colscale, int k,
                  const double* __restrict__ cov_mat, const double* __restrict__ covM,
                  const double* __restrict__ covm, float* __restrict__ dev_mat,
                  double* __restrict__ row_ref) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* acc = reinterpret_cast<double*>(smem_raw) + (size_t)warp * acc_stride;
    const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(acc);
    const uint32_t* __restrict__ p_col32 = reinterpret_cast<const uint32_t*>(p_col);
    const float2* __restrict__ p_val2 = reinterpret_cast<const float2*>(p_val);

    for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < n_cell;
         row += (int64_t)gridDim.x * wpb) {
        for (int j = lane; j < acc_stride; j += 32) acc[j] = 0.0;
        __syncwarp();
        const int64_t p0 = indptr[row], p1 = indptr[row + 1];
        for (int64_t base = p0; base < p1; base += 32) {
            const int64_t q = base + lane;
            uint32_t meta = 0;
            double xd = 0.0;
            if (q < p1) {
                const int4 gi = __ldg(reinterpret_cast<const int4*>(ginfo + indices[q]));
                xd = (double)data[q] * __hiloint2double(gi.y, gi.x);
                meta = (uint32_t)gi.z;
            }
            const int n = (int)((p1 - base) < 32 ? (p1 - base) : 32);
            // software pipeline in blocks of four entries: the set ids of the next block are in
            // flight (L2 latency, the lists rarely hit in L1) while the current block is applied
            uint32_t a_ids[4], a_m[4], b_ids[4], b_m[4];      // packed set ids of groups (0|1),(2|3); meta
            float aw0[4], aw1[4], bw0[4], bw1[4];
            auto fetch4 = [&](uint32_t (&ids)[4], uint32_t (&mm)[4], float (&w0)[4], float (&w1)[4], int j) {
                if (j >= n) {                                    // warp-uniform: nothing left in this chunk
#pragma unroll
                    for (int u = 0; u < 4; ++u) { ids[u] = 0xffffffffu; mm[u] = 0; }
                    return;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t m = __shfl_sync(0xffffffffu, meta, (j + u) & 31);
                    if (j + u >= n) m = 0;
                    // one 32-bit load brings this lane's set ids of groups (0|1) and (2|3); slots past
                    // the end of the list hold 0xffff
                    const uint32_t t = (m >> 12) * 32u + lane;
                    ids[u] = (m & 4095u) ? __ldg(p_col32 + t) : 0xffffffffu;
                    mm[u] = m;
                    if (WEIGHTED) {
                        const float2 w = (m & 4095u) ? __ldg(p_val2 + t) : make_float2(0.f, 0.f);
                        w0[u] = w.x;
                        w1[u] = w.y;
                    }
                }
            };
            auto apply4 = [&](uint32_t (&ids)[4], uint32_t (&mm)[4], float (&w0)[4], float (&w1)[4], int j) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (j + u >= n) break;                       // warp-uniform
                    const double xi = __shfl_sync(0xffffffffu, xd, j + u);
                    if (WEIGHTED)
                        acc_fma2(acc_s, ids[u], xi, w0[u], w1[u]);
                    else
                        acc_add2(acc_s, ids[u], xi);
                    const uint32_t m = mm[u], ngc = m & 4095u;
                    if (ngc > 4) {                               // long lists (rare); warp-uniform
                        const uint32_t t0 = (m >> 12) * 32u + lane;
#pragma unroll 1
                        for (uint32_t blk = 1; blk * 4 < ngc; ++blk) {
                            const uint32_t both = __ldg(p_col32 + t0 + blk * 32u);
                            if (WEIGHTED) {
                                const float2 w = __ldg(p_val2 + t0 + blk * 32u);
                                acc_fma2(acc_s, both, xi, w.x, w.y);
                            } else {
                                acc_add2(acc_s, both, xi);
                            }
                        }
                    }
                    __syncwarp();
                }
            };
            fetch4(a_ids, a_m, aw0, aw1, 0);
#pragma unroll 1
            for (int i = 0; i < n; i += 8) {
                fetch4(b_ids, b_m, bw0, bw1, i + 4);
                apply4(a_ids, a_m, aw0, aw1, i);
                fetch4(a_ids, a_m, aw0, aw1, i + 8);
                apply4(b_ids, b_m, bw0, bw1, i + 4);
            }
        }
        // epilogue: column scale, covariate terms, row reference, deviations
        double rs = 0.0;
        for (int j = lane; j < ncol; j += 32) {
            double v = acc[j] * colscale[j];
            if (k > 0) {
                double cterm = covm[j];
                for (int kk = 0; kk < k; ++kk)
                    cterm = fma(cov_mat[row * k + kk], covM[(size_t)kk * ncol + j], cterm);
                v += cterm;
            }
            acc[j] = v;
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

static int spmm_score_f64(scdrs_ctx* c);

int spmm_score(scdrs_ctx* c) {
    if (c->tl.on) cudaEventRecord(c->tl.sc0[c->tl.n & 15], c->stream);
    SCDRS_TRY(c->w_fx ? fx_spmm(c) : spmm_score_f64(c));
    if (c->tl.on) {
        cudaEventRecord(c->tl.sc1[c->tl.n & 15], c->stream);
        ++c->tl.n;
    }
    // the preparation of the next trait may overwrite the lists from here on (PrepScope)
    if (c->ev_scatter != nullptr) {
        SCDRS_CUDA(cudaEventRecord(c->ev_scatter, c->stream));
        c->scatter_marked = true;
    }
    return 0;
}

static int spmm_score_f64(scdrs_ctx* c) {
    const int ncol = c->ncol;
    SCDRS_TRY(c->dev_mat.reserve((size_t)c->n_cell * c->ld * sizeof(float)));
    SCDRS_TRY(c->row_ref.reserve((size_t)c->n_cell * sizeof(double)));
    const int acc_stride = (ncol + 15) & ~15;
    const size_t per_warp = (size_t)acc_stride * sizeof(double);          // one cell's accumulators
    SCDRS_CHECK(per_warp <= 200 * 1024, "1 + n_ctrl = %d needs more shared memory than one SM has", ncol);
    // warps per block / blocks per SM that keep the most cells in flight within 227 KB of shared
    // memory per SM (each resident block also costs 1 KB of reserved shared memory)
    int wpb = 1, per_sm = 1, best = 0;
    for (int w = 1; w <= 8; ++w) {
        const size_t blk = per_warp * w + 1024;
        if (per_warp * w > 200 * 1024) break;
        int b = (int)((227 * 1024) / blk);
        if (b > 32) b = 32;
        if (b * w > 64) b = 64 / w;
        if (b >= 1 && b * w >= best) { best = b * w; wpb = w; per_sm = b; }
    }
    const size_t smem = per_warp * wpb;
    static bool attr_set = false;
    if (!attr_set) {
        SCDRS_CUDA(cudaFuncSetAttribute(spmm_score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SCDRS_CUDA(cudaFuncSetAttribute(spmm_score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    int64_t grid = (int64_t)kNumSM * per_sm;
    const int64_t need = (c->n_cell + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (c->timed_spmm) SCDRS_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (c->w_binary)
        spmm_score_kernel<false><<<(unsigned)grid, wpb * 32, smem, c->stream>>>(
            c->n_cell, ncol, c->ld, acc_stride, c->indptr, c->indices, c->data, c->p_meta.as<GeneInfo>(),
            c->p_col.as<uint16_t>(), nullptr, c->colscale.as<double>(), c->k,
            c->cov_mat.as<double>(), c->covM.as<double>(), c->covm.as<double>(), c->dev_mat.as<float>(),
            c->row_ref.as<double>());
    else
        spmm_score_kernel<true><<<(unsigned)grid, wpb * 32, smem, c->stream>>>(
            c->n_cell, ncol, c->ld, acc_stride, c->indptr, c->indices, c->data, c->p_meta.as<GeneInfo>(),
            c->p_col.as<uint16_t>(), c->p_val.as<float>(), c->colscale.as<double>(),
            c->k, c->cov_mat.as<double>(), c->covM.as<double>(), c->covm.as<double>(),
            c->dev_mat.as<float>(), c->row_ref.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    if (c->timed_spmm) SCDRS_CUDA(cudaEventRecord(c->ev[1], c->stream));
    c->trait_ready = true;
    return 0;
}

}  // namespace scdrs
