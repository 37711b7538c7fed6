// K1, fixed-point variant: raw scores of the trait gene set and its n_ctrl control gene sets.
//
// Same contract as spmm_score.cu (the 1 + n_ctrl calls of _compute_raw_score, scdrs/method.py:172-183,
// 312-402 of the reference), different arithmetic.  The fp64 kernel is bound by the shared-memory
// pipe: one 64-bit load + store per add and, on top, bank conflicts that no placement of a gene's
// ~50 set ids over 16 bank pairs can avoid (the busiest pair holds ~6.5 of them).  Here
//
//   * every stored entry x * base_g is scaled by 2^s and rounded to an integer ONCE; the per-set
//     sums are then exact integer sums (order independent, bit reproducible) in 32-bit shared-memory
//     words, half the traffic of fp64.  The scale is the largest power of two that keeps one entry
//     below 2^25; a running bound on the sum of the |entries| applied so far triggers a flush of the
//     words into 64-bit per-lane register totals before any word can overflow, so the result is
//     exact for any input.  Quantisation error: 2^-26 of the largest entry per item (about 1e-7 of
//     the normalised score at the benchmark shapes; the reference computes in float64).
//   * every set has TWO accumulator words, in different banks (array A: bank j mod 32, array B:
//     bank (j + j/32) mod 32), and each (gene, set) entry is assigned to one of them by a matching
//     (augmenting paths, fx_assign_kernel) such that round r of a gene's list touches every bank at
//     most once.  An entry that lives in bank b is handled by lane b, so an access is conflict-free
//     by construction, unused lanes add into a per-lane trash word (no predication), and all adds
//     to one word come from the same lane in program order (no warp barriers).  ceil(n / 32) rounds
//     per gene suffice in practice: 2 rounds = 4 wavefronts per stored entry instead of ~14.
//
// Score matrices wider than 1024 columns are produced in column tiles of 1024 sets (own lists, own pass
// over X; the first tile fixes the row reference).  Used when max|x * base| <= 128 * mean|x * base|;
// otherwise, and for explicit-weight calls, the fp64 kernel runs.
#include <math.h>

#include "common.cuh"

namespace scdrs {

constexpr int kFxNB = kFxMaxCol;              // words per accumulator array
constexpr int kFxTrashOff = 2 * kFxNB * 4;    // byte offset of the 32 per-lane trash words
constexpr int kFxItemOff = kFxTrashOff + 128;   // byte offset of the staged work items (list block, value)
constexpr int kFxItemCap = 128;                 // four passes over a chunk of 32 stored entries
constexpr int kFxMaxDepth = 8;                  // register slots of the item pipeline
constexpr int kFxWarpsPerBlock = 10, kFxBlocksPerSM = 2;   // 20 warps per SM at <= 96 registers
constexpr int kFxWarpBytes = kFxItemOff + (kFxItemCap + 2 * kFxMaxDepth) * 8;
constexpr int kFxItemBits = 25;               // |item| <= 2^25: 32 stored entries add at most 2^30
constexpr uint32_t kFxFlushAt = 0x7f000000u;  // flush before the bound on any word reaches this
constexpr uint32_t kFxInflight = ((uint32_t)kFxMaxDepth << kFxItemBits) + 64u;   // bound on the items held in the register slots

__host__ __device__ __forceinline__ uint32_t fx_bank_a(uint32_t j) { return j & 31u; }
__host__ __device__ __forceinline__ uint32_t fx_bank_b(uint32_t j) { return (j + (j >> 5)) & 31u; }
__host__ __device__ __forceinline__ uint32_t fx_off_a(uint32_t j) { return j * 4u; }
__host__ __device__ __forceinline__ uint32_t fx_off_b(uint32_t j) {
    return ((uint32_t)kFxNB + (j & ~31u) + fx_bank_b(j)) * 4u;
}

// ---- value statistics (once per expression matrix and weight_opt) ---------------------------------
constexpr int kStatBlocks = 592, kStatThreads = 256;

__global__ void __launch_bounds__(kStatThreads)
fx_vstat_kernel(int64_t nnz, const int32_t* __restrict__ indices, const float* __restrict__ data,
                const double* __restrict__ gbase, double* __restrict__ part) {
    __shared__ double smax[kStatThreads / 32], ssum[kStatThreads / 32];
    double mx = 0.0, sm = 0.0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
        const double v = fabs((double)data[q] * gbase[indices[q]]);
        mx = v > mx ? v : mx;   // NaN never wins; non-finite values end in the fp64 kernel via the ratio test
        sm += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = m2 > mx ? m2 : mx;
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { smax[warp] = mx; ssum[warp] = sm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kStatThreads / 32; ++w) {
            mx = smax[w] > mx ? smax[w] : mx;
            sm += ssum[w];
        }
        part[2 * blockIdx.x] = mx;
        part[2 * blockIdx.x + 1] = sm;
    }
}

__global__ void fx_vstat_final_kernel(int nb, double* __restrict__ part) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double mx = 0.0, sm = 0.0;
    for (int b = 0; b < nb; ++b) {
        mx = part[2 * b] > mx ? part[2 * b] : mx;
        sm += part[2 * b + 1];
    }
    part[2 * nb] = mx;
    part[2 * nb + 1] = sm;
}

bool fx_applicable(scdrs_ctx* c, int ncol, bool explicit_w) {
    // n_gene bounds the stored entries of a row, which keeps the 32-bit column totals exact (see the kernel)
    return c->spmm_mode != 1 && !explicit_w && ncol <= kFxMaxCol * kFxMaxTiles && c->nnz > 0 && c->n_gene < (1 << 18);
}

// max and mean of |x * base_g| over the stored entries; c->gbase must hold the base weights of
// `weight_opt`.  Cached per weight_opt (reset when the matrix or the gene statistics change).
int fx_value_stats(scdrs_ctx* c, int weight_opt) {
    if (c->fx_stat_key == weight_opt) return 0;
    SCDRS_TRY(c->fx_part.reserve((size_t)(2 * kStatBlocks + 2) * sizeof(double)));
    fx_vstat_kernel<<<kStatBlocks, kStatThreads, 0, c->stream>>>(c->nnz, c->indices, c->data, c->gbase.as<double>(),
                                                                 c->fx_part.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    fx_vstat_final_kernel<<<1, 32, 0, c->stream>>>(kStatBlocks, c->fx_part.as<double>());
    SCDRS_LAUNCH_CHECK(c);
    double out[2];
    SCDRS_CUDA(cudaMemcpyAsync(out, c->fx_part.as<double>() + 2 * kStatBlocks, sizeof(out), cudaMemcpyDeviceToHost, c->stream));
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    c->fx_vmax = out[0];
    c->fx_vmean = out[1] / (double)c->nnz;
    c->fx_stat_key = weight_opt;
    return 0;
}

// ---- assignment of a gene's entries to (round, accumulator array) ----------------------------------
// One thread per gene.  Lists of at most 128 entries: bank b (0..31) has R slots (one per round); an
// entry may go to bank_a(j) or bank_b(j); entries are inserted one by one into the emptier bank and,
// when both are full, along a shortest augmenting path (breadth-first over banks, moving resident
// entries to their other bank); R starts at ceil(n / 32) and grows only when no path exists.  Longer
// lists (and R > 8) use plain two-choice greedy.  code[t] = (round << 1) | array of entry t.
constexpr int kAsgThreads = 8;      // genes per block: ONE gene per warp (lane 0 works), so divergent genes do not serialise each other
constexpr int kAsgBlock = kAsgThreads * 32;
constexpr int kAsgMaxLen = 128;
constexpr int kAsgMaxR = 8;
constexpr int kAsgBytesPerThread = 2 * 32 * kAsgMaxR + 32 + 32 + 32 + 64;

__global__ void __launch_bounds__(kAsgBlock)
fx_assign_kernel(int n_gene, const uint32_t* __restrict__ w_ptr, const uint16_t* __restrict__ w_col,
                 uint16_t* __restrict__ code, uint32_t* __restrict__ rounds) {
    extern __shared__ unsigned char asg_sm[];
    constexpr int T = kAsgThreads;
    if ((threadIdx.x & 31) != 0) return;
    const int tid = threadIdx.x >> 5;
    unsigned char* sm_slot = asg_sm;                              // [32 * kAsgMaxR][T] entry in the slot
    unsigned char* sm_alt = sm_slot + 32 * kAsgMaxR * T;          // [32 * kAsgMaxR][T] its other bank
    unsigned char* sm_par = sm_alt + 32 * kAsgMaxR * T;           // [32][T]
    unsigned char* sm_via = sm_par + 32 * T;                      // [32][T]
    unsigned char* sm_que = sm_via + 32 * T;                      // [32][T]
    unsigned short* sm_load = reinterpret_cast<unsigned short*>(sm_que + 32 * T);  // [32][T]
#define FX_SLOT(b, k) sm_slot[((b) * kAsgMaxR + (k)) * T + tid]
#define FX_ALT(b, k) sm_alt[((b) * kAsgMaxR + (k)) * T + tid]
#define FX_LOAD(b) sm_load[(b) * T + tid]
    const int g = blockIdx.x * T + tid;
    if (g >= n_gene) return;
    if (g == 0) rounds[n_gene] = 0;
    const uint32_t s = w_ptr[g], e = w_ptr[g + 1], n = e - s;
    if (n == 0) {
        rounds[g] = 0;
        return;
    }
    constexpr uint32_t ROOT = 255;
    for (int b = 0; b < 32; ++b) FX_LOAD(b) = 0;
    bool greedy = n > (uint32_t)kAsgMaxLen;
    uint32_t R = (n + 31u) >> 5;
    if (!greedy) {
        uint32_t jn = w_col[s];
        for (uint32_t t = 0; t < n; ++t) {
            const uint32_t j = jn;
            if (t + 1 < n) jn = w_col[s + t + 1];
            const uint32_t a = fx_bank_a(j), b = fx_bank_b(j);
            const uint32_t la = FX_LOAD(a), lb = FX_LOAD(b);
            if (la < R && (la <= lb || lb >= R)) {
                FX_SLOT(a, la) = (unsigned char)t;
                FX_ALT(a, la) = (unsigned char)b;
                FX_LOAD(a) = (unsigned short)(la + 1);
                continue;
            }
            if (lb < R) {
                FX_SLOT(b, lb) = (unsigned char)t;
                FX_ALT(b, lb) = (unsigned char)a;
                FX_LOAD(b) = (unsigned short)(lb + 1);
                continue;
            }
            // both banks full: breadth-first search for a bank with a free slot
            uint32_t visited = (1u << a) | (1u << b);
            uint32_t qh = 0, qt = 0;
            sm_que[(qt++) * T + tid] = (unsigned char)a;
            sm_par[a * T + tid] = (unsigned char)ROOT;
            if (b != a) {
                sm_que[(qt++) * T + tid] = (unsigned char)b;
                sm_par[b * T + tid] = (unsigned char)ROOT;
            }
            int target = -1;
            while (qh < qt && target < 0) {
                const uint32_t cb = sm_que[(qh++) * T + tid];
                const uint32_t lcb = FX_LOAD(cb);
                for (uint32_t k = 0; k < lcb; ++k) {
                    const uint32_t alt = FX_ALT(cb, k);
                    if ((visited >> alt) & 1u) continue;     // covers alt == cb
                    visited |= 1u << alt;
                    sm_par[alt * T + tid] = (unsigned char)cb;
                    sm_via[alt * T + tid] = (unsigned char)k;
                    if (FX_LOAD(alt) < R) {
                        target = (int)alt;
                        break;
                    }
                    sm_que[(qt++) * T + tid] = (unsigned char)alt;
                }
            }
            if (target < 0) {
                if (R < (uint32_t)kAsgMaxR) {     // no augmenting path: one more round
                    FX_SLOT(a, la) = (unsigned char)t;
                    FX_ALT(a, la) = (unsigned char)b;
                    FX_LOAD(a) = (unsigned short)(la + 1);
                    ++R;
                    continue;
                }
                greedy = true;
                break;
            }
            // shift entries along the path; the hole moves back to one of the new entry's banks
            uint32_t hb = (uint32_t)target, hk = FX_LOAD(hb), cur = hb;
            FX_LOAD(hb) = (unsigned short)(hk + 1);
            while (sm_par[cur * T + tid] != ROOT) {
                const uint32_t p = sm_par[cur * T + tid], k = sm_via[cur * T + tid];
                FX_SLOT(hb, hk) = FX_SLOT(p, k);
                FX_ALT(hb, hk) = (unsigned char)p;
                hb = p;
                hk = k;
                cur = p;
            }
            FX_SLOT(hb, hk) = (unsigned char)t;
            FX_ALT(hb, hk) = (unsigned char)(hb == a ? b : a);
        }
    }
    uint32_t rmax = 0;
    if (!greedy) {
        for (uint32_t b = 0; b < 32; ++b) {
            const uint32_t lb = FX_LOAD(b);
            rmax = lb > rmax ? lb : rmax;
            for (uint32_t k = 0; k < lb; ++k) {
                const uint32_t i = FX_SLOT(b, k);
                const uint32_t side = fx_bank_a(w_col[s + i]) == b ? 0u : 1u;
                code[s + i] = (uint16_t)((k << 1) | side);
            }
        }
    } else {
        for (int b = 0; b < 32; ++b) FX_LOAD(b) = 0;
        for (uint32_t t = 0; t < n; ++t) {
            const uint32_t j = w_col[s + t];
            const uint32_t a = fx_bank_a(j), b = fx_bank_b(j);
            const uint32_t la = FX_LOAD(a), lb = FX_LOAD(b);
            if (la <= lb) {
                code[s + t] = (uint16_t)(la << 1);
                FX_LOAD(a) = (unsigned short)(la + 1);
                rmax = la + 1 > rmax ? la + 1 : rmax;
            } else {
                code[s + t] = (uint16_t)((lb << 1) | 1u);
                FX_LOAD(b) = (unsigned short)(lb + 1);
                rmax = lb + 1 > rmax ? lb + 1 : rmax;
            }
        }
    }
    rounds[g] = rmax;
#undef FX_SLOT
#undef FX_ALT
#undef FX_LOAD
}

// padded list length in 16-bit entries: blocks of two rounds x 32 lanes
__global__ void fx_plen_kernel(int n_gene, const uint32_t* __restrict__ rounds, uint32_t* __restrict__ plen) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= n_gene) plen[g] = g < n_gene ? ((rounds[g] + 1u) >> 1) * 64u : 0u;
}

// One warp per gene: fill the gene's blocks with the per-lane trash offsets, then drop every entry
// into (round, lane = its bank).  A block holds, for lane l, the byte offsets of rounds 2k and 2k+1
// next to each other, so the scatter kernel fetches both with one 32-bit load.
__global__ void __launch_bounds__(128)
fx_place_kernel(int n_gene, const uint32_t* __restrict__ w_ptr, const uint32_t* __restrict__ p_ptr,
                const uint32_t* __restrict__ rounds, const uint16_t* __restrict__ w_col,
                const uint16_t* __restrict__ code, const float* __restrict__ w_val,
                uint16_t* __restrict__ p_col, float* __restrict__ p_val) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (g >= n_gene) return;
    const uint32_t s = w_ptr[g], e = w_ptr[g + 1], ps = p_ptr[g] + 64u, nblk = (rounds[g] + 1u) >> 1;
    uint32_t* col32 = reinterpret_cast<uint32_t*>(p_col) + (ps >> 1);
    const uint32_t trash = (uint32_t)(kFxTrashOff + 4 * lane) * 0x10001u;
    if (g == 0) {   // block 0: every lane adds into its trash word (what pipeline bubbles fetch)
        reinterpret_cast<uint32_t*>(p_col)[lane] = trash;
        if (p_val != nullptr) { p_val[lane] = 0.f; p_val[32 + lane] = 0.f; }
    }
    for (uint32_t w = lane; w < nblk * 32u; w += 32) col32[w] = trash;
    if (p_val != nullptr)
        for (uint32_t w = lane; w < nblk * 64u; w += 32) p_val[ps + w] = 0.f;
    __syncwarp();
    for (uint32_t t = s + lane; t < e; t += 32) {
        const uint32_t j = w_col[t], cd = code[t], r = cd >> 1;
        const uint32_t bank = (cd & 1u) ? fx_bank_b(j) : fx_bank_a(j);
        const uint32_t off = (cd & 1u) ? fx_off_b(j) : fx_off_a(j);
        const uint32_t dst = ps + (r >> 1) * 64u + bank * 2u + (r & 1u);
        p_col[dst] = (uint16_t)off;
        if (p_val != nullptr) p_val[dst] = w_val[t];
    }
}

// GeneInfo for the fixed-point kernel: base = base_g * 2^s, meta = (first block << 8) | blocks
// (block 0 of the list storage is the all-trash block, a gene's blocks start at 1 + p_ptr / 64)
__global__ void fx_meta_kernel(int n_gene, const uint32_t* __restrict__ p_ptr, const uint32_t* __restrict__ rounds,
                               const double* __restrict__ gbase, double scale, GeneInfo* __restrict__ info,
                               uint32_t* __restrict__ meta32) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_gene) {
        GeneInfo gi;
        gi.base = gbase[g] * scale;
        gi.meta = (((p_ptr[g] >> 6) + 1u) << 8) | ((rounds[g] + 1u) >> 1);
        gi.pad = 0;
        info[g] = gi;
        meta32[g] = gi.meta;      // what the pre-quantised path gathers: 4 bytes per gene instead of 16
    }
}

// Pre-quantised stored entries round(x * base_g * 2^s), once per (matrix, weight_opt, s): binary traits share
// the scale, so the scatter of every trait reads these integers instead of converting x on the fly, and
// gathers a 4-byte list location per entry instead of a 16-byte record.
__global__ void __launch_bounds__(256)
fx_quantise_kernel(int64_t nnz, const int32_t* __restrict__ indices, const float* __restrict__ data,
                   const double* __restrict__ gbase, double scale, int32_t* __restrict__ xq) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x)
        xq[q] = __double2int_rn((double)data[q] * (gbase[indices[q]] * scale));
}

// Column means of the per-set constants: what the scatter kernel needs to place a cell's reference value
// (an estimate of its mean raw score) without a pass over the per-column constants.
// out[0] = mean colscale, out[1] = mean covm, out[2 + kk] = mean covM[kk, :].  One block, fixed order.
__global__ void __launch_bounds__(256)
fx_scalars_kernel(int ncol, int cov_stride, int k, const double* __restrict__ colscale, const double* __restrict__ covm,
                  const double* __restrict__ covM, double* __restrict__ out) {
    __shared__ double sh[256];
    for (int q = 0; q < 2 + k; ++q) {
        const double* src = q == 0 ? colscale : (q == 1 ? covm : covM + (size_t)(q - 2) * cov_stride);
        double part = 0.0;
        if (q != 1 || k > 0)
            for (int j = threadIdx.x; j < ncol; j += 256) part += src[j];
        sh[threadIdx.x] = part;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[q] = sh[0] / (double)ncol;
        __syncthreads();
    }
}

// Scale and tiling of the fixed-point scatter for the trait whose per-set constants build_w has just
// computed.  A score matrix wider than kFxMaxCol columns is produced in column tiles (sets [1024 T,
// 1024 (T + 1))), each with its own gene-major lists and its own pass over X; with one tile the lists
// are built here, otherwise per tile right before its pass (fx_spmm).
static int fx_build_tile(scdrs_ctx* c, int tile);

int fx_prepare(scdrs_ctx* c, bool binary, double gw_absmax) {
    SCDRS_CHECK(c->fx_vmax > 0.0 && gw_absmax > 0.0 && isfinite(c->fx_vmax) && isfinite(gw_absmax),
                "fixed-point scatter needs finite non-zero values");
    int shift = ilogb(ldexp(1.0, kFxItemBits) / (c->fx_vmax * gw_absmax));
    if (shift > 200) shift = 200;
    if (shift < -200) shift = -200;
    c->fx_shift = shift;
    c->fx_gwmax = (float)gw_absmax * 1.000001f;
    c->fx_n_tile = (c->ncol + kFxMaxCol - 1) / kFxMaxCol;
    c->w_binary = binary;
    c->w_fx = true;
    c->fx_tile0_built = false;
    c->fx_preq = false;
    if (binary && (size_t)c->nnz * sizeof(int32_t) <= ((size_t)48 << 30)) {
        if (!(c->fx_xq_key == c->fx_stat_key && c->fx_xq_shift == shift && c->fx_xq_key >= 0)) {
            SCDRS_TRY(c->fx_xq.reserve((size_t)c->nnz * sizeof(int32_t)));
            fx_quantise_kernel<<<kNumSM * 8, 256, 0, c->stream>>>(c->nnz, c->indices, c->data, c->gbase.as<double>(),
                                                                 ldexp(1.0, shift), c->fx_xq.as<int32_t>());
            SCDRS_LAUNCH_CHECK(c);
            c->fx_xq_key = c->fx_stat_key;
            c->fx_xq_shift = shift;
        }
        c->fx_preq = true;
    }
    if (c->fx_n_tile == 1) {
        SCDRS_TRY(fx_build_tile(c, 0));
        c->fx_tile0_built = true;
    }
    return 0;
}

static int fx_build_tile(scdrs_ctx* c, int tile) {
    const int n_gene = c->n_gene;
    const bool binary = c->w_binary;
    const int col0 = tile * kFxMaxCol, col1 = (col0 + kFxMaxCol < c->ncol) ? col0 + kFxMaxCol : c->ncol;
    const int ncol_t = col1 - col0;
    const int64_t e_begin = col0 == 0 ? 0 : (int64_t)c->n_s + (int64_t)(col0 - 1) * c->n_sc;
    const int64_t e_end = (int64_t)c->n_s + (int64_t)(col1 - 1) * c->n_sc;
    const int64_t n_entry = e_end - e_begin;
    // worst case: every round of a gene holds one entry
    const int64_t n_slot_max = 32 * n_entry + 64 * (int64_t)n_gene + 128;
    SCDRS_CHECK(n_slot_max < ((int64_t)1 << 30), "gene-set lists too large for the fixed-point layout");
    SCDRS_TRY(build_gene_lists(c, col0, col1));
    SCDRS_TRY(c->w_code.reserve((size_t)(n_entry > 0 ? n_entry : 1) * sizeof(uint16_t)));
    SCDRS_TRY(c->w_rounds.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    SCDRS_TRY(c->fx_meta32.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    SCDRS_TRY(c->p_col.reserve((size_t)n_slot_max * sizeof(uint16_t)));
    if (!binary) SCDRS_TRY(c->p_val.reserve((size_t)n_slot_max * sizeof(float)));
    static bool attr = false;
    const int asg_smem = kAsgBytesPerThread * kAsgThreads;
    if (!attr) {
        SCDRS_CUDA(cudaFuncSetAttribute(fx_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, asg_smem));
        attr = true;
    }
    fx_assign_kernel<<<(n_gene + kAsgThreads - 1) / kAsgThreads, kAsgBlock, asg_smem, c->stream>>>(
        n_gene, c->w_ptr.as<uint32_t>(), c->w_col.as<uint16_t>(), c->w_code.as<uint16_t>(), c->w_rounds.as<uint32_t>());
    SCDRS_LAUNCH_CHECK(c);
    fx_plen_kernel<<<(n_gene + 256) / 256, 256, 0, c->stream>>>(n_gene, c->w_rounds.as<uint32_t>(), c->p_ptr.as<uint32_t>());
    SCDRS_LAUNCH_CHECK(c);
    SCDRS_TRY(exclusive_scan_u32(c, c->p_ptr.as<uint32_t>(), c->p_ptr.as<uint32_t>(), n_gene + 1));
    fx_place_kernel<<<(n_gene + 3) / 4, 128, 0, c->stream>>>(
        n_gene, c->w_ptr.as<uint32_t>(), c->p_ptr.as<uint32_t>(), c->w_rounds.as<uint32_t>(), c->w_col.as<uint16_t>(),
        c->w_code.as<uint16_t>(), c->w_val.as<float>(), c->p_col.as<uint16_t>(), binary ? nullptr : c->p_val.as<float>());
    SCDRS_LAUNCH_CHECK(c);
    fx_meta_kernel<<<(n_gene + 255) / 256, 256, 0, c->stream>>>(n_gene, c->p_ptr.as<uint32_t>(), c->w_rounds.as<uint32_t>(),
                                                               c->gbase.as<double>(), ldexp(1.0, c->fx_shift), c->p_meta.as<GeneInfo>(),
                                                               c->fx_meta32.as<uint32_t>());
    SCDRS_LAUNCH_CHECK(c);
    if (tile == 0) {
        // column means of the per-set constants over the first tile: they place the row reference
        SCDRS_TRY(c->fx_scal.reserve((size_t)(2 + (c->k > 0 ? c->k : 0)) * sizeof(double)));
        fx_scalars_kernel<<<1, 256, 0, c->stream>>>(ncol_t, c->ncol, c->k, c->colscale.as<double>(), c->covm.as<double>(),
                                                     c->covM.as<double>(), c->fx_scal.as<double>());
        SCDRS_LAUNCH_CHECK(c);
    }
    return 0;
}

// ---- the scatter kernel ------------------------------------------------------------------------------
__device__ __forceinline__ double fx_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// two read-modify-writes on the words at byte offsets (both & 0xffff) and (both >> 16)
__device__ __forceinline__ void fx_add2(uint32_t acc_s, uint32_t both, int q) {
    asm volatile(
        "{\n\t.reg .u32 a0, a1;\n\t.reg .s32 v0, v1;\n\t"
        "and.b32 a0, %1, 0xffff;\n\tshr.u32 a1, %1, 16;\n\t"
        "add.u32 a0, a0, %0;\n\tadd.u32 a1, a1, %0;\n\t"
        "ld.shared.s32 v0, [a0];\n\tld.shared.s32 v1, [a1];\n\t"
        "add.s32 v0, v0, %2;\n\tadd.s32 v1, v1, %2;\n\t"
        "st.shared.s32 [a0], v0;\n\tst.shared.s32 [a1], v1;\n\t}"
        ::"r"(acc_s), "r"(both), "r"(q)
        : "memory");
}
__device__ __forceinline__ void fx_fma2(uint32_t acc_s, uint32_t both, float vs, float w0, float w1) {
    asm volatile(
        "{\n\t.reg .u32 a0, a1;\n\t.reg .s32 v0, v1, q0, q1;\n\t.reg .f32 f0, f1;\n\t"
        "and.b32 a0, %1, 0xffff;\n\tshr.u32 a1, %1, 16;\n\t"
        "add.u32 a0, a0, %0;\n\tadd.u32 a1, a1, %0;\n\t"
        "mul.rn.f32 f0, %2, %3;\n\tmul.rn.f32 f1, %2, %4;\n\t"
        "cvt.rni.s32.f32 q0, f0;\n\tcvt.rni.s32.f32 q1, f1;\n\t"
        "ld.shared.s32 v0, [a0];\n\tld.shared.s32 v1, [a1];\n\t"
        "add.s32 v0, v0, q0;\n\tadd.s32 v1, v1, q1;\n\t"
        "st.shared.s32 [a0], v0;\n\tst.shared.s32 [a1], v1;\n\t}"
        ::"r"(acc_s), "r"(both), "f"(vs), "f"(w0), "f"(w1)
        : "memory");
}

// volatile, so that the compiler keeps the load where it is written (between the volatile
// read-modify-writes) instead of sinking it towards its use
__device__ __forceinline__ uint32_t fx_ldg_pinned(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// One warp per cell.  ginfo[g].meta = (first two-round block of gene g's list << 8) | number of blocks.
//
// The shared-memory work per list block is ~9 instructions, so everything around it is kept as lean
// and as branch-free as possible, and three nested software pipelines hide the global latencies:
//   * column ids / values of the row are loaded two chunks of 32 entries ahead, the per-gene record
//     (base weight, list location) one chunk ahead;
//   * per chunk the work items (list block, quantised value) are compacted into shared memory, pass
//     by pass (all first blocks, then all second blocks, ...; legal because integer adds commute and
//     a word is only ever touched by one lane), padded with bubbles (the all-trash block 0, value 0)
//     to a multiple of D;
//   * the items flow through D register slots with no per-slot control flow: slot u is applied, then
//     refilled with the list words of the item D positions later (one broadcast shared load, one
//     global load), across chunk and pass-group boundaries.
// Per-lane column totals are 32-bit: a flush moves word >> 12 into the total and leaves the low 12 bits
// in the word (|sum over a row| < 2^18 entries * 2^25 / 2^12 = 2^31), which keeps the kernel at 96
// registers = 20 warps per SM.
template <bool WEIGHTED, int D, bool PREQ>
__global__ void __launch_bounds__(kFxWarpsPerBlock * 32, kFxBlocksPerSM)
spmm_score_fx_kernel(int64_t n_cell, int ncol, int cov_stride, int write_ref, int ld, const int64_t* __restrict__ indptr,
                     const int32_t* __restrict__ indices, const float* __restrict__ data,
                     const int32_t* __restrict__ xq_arr, const uint32_t* __restrict__ meta32,
                     const GeneInfo* __restrict__ ginfo, const uint32_t* __restrict__ p_col32,
                     const float2* __restrict__ p_val2, const double* __restrict__ colscale,
                     double inv_scale, float gwmax, int k, const double* __restrict__ cov_mat,
                     const double* __restrict__ covM, const double* __restrict__ covm,
                     const double* __restrict__ scal, float* __restrict__ dev_mat, double* __restrict__ row_ref,
                     unsigned long long* __restrict__ next_row) {
    static_assert(D <= kFxMaxDepth && (kFxItemCap % D) == 0, "pipeline depth");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned char* mine = smem_raw + (size_t)warp * kFxWarpBytes;
    uint32_t* accw = reinterpret_cast<uint32_t*>(mine);
    uint2* items = reinterpret_cast<uint2*>(mine + kFxItemOff);
    const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(mine);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t trash_ids = (uint32_t)(kFxTrashOff + 4 * lane) * 0x10001u;
    for (int i = lane; i < 2 * kFxNB / 4; i += 32) reinterpret_cast<uint4*>(accw)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();

    // rows are handed out dynamically (one atomic per cell): cells differ in stored entries and every
    // warp should finish at the same time; the result of a cell does not depend on who computes it
    (void)wpb;
    for (;;) {
        long long row_l = 0;
        if (lane == 0) row_l = (long long)atomicAdd(next_row, 1ull);
        const int64_t row = __shfl_sync(0xffffffffu, row_l, 0);
        if (row >= n_cell) break;
        int hi[kFxNB / 32];
#pragma unroll
        for (int t = 0; t < kFxNB / 32; ++t) hi[t] = 0;
        // item slots, initially bubbles
        uint32_t s_ids[D];
        int s_x[D];
        float2 s_w[WEIGHTED ? D : 1];
#pragma unroll
        for (int u = 0; u < D; ++u) {
            s_ids[u] = trash_ids;
            s_x[u] = 0;
            if (WEIGHTED) s_w[u] = make_float2(0.f, 0.f);
        }
        auto apply = [&](int u) {
            if (WEIGHTED)
                fx_fma2(acc_s, s_ids[u], __int_as_float(s_x[u]), s_w[u].x, s_w[u].y);
            else
                fx_add2(acc_s, s_ids[u], s_x[u]);
        };

        uint32_t lbound = 0;
        uint32_t n_left = 0;          // staged items not yet in the slots (fewer than D), carried to the next group
        const int64_t p0 = indptr[row], p1 = indptr[row + 1];
        // chunk pipeline registers: record + value of the current chunk, ids + values of the next two
        // (pre-quantised path: the value is the integer entry, the record just the 4-byte list location)
        auto load_val = [&](int64_t q) -> float {
            return PREQ ? __int_as_float(__ldg(xq_arr + q)) : __ldg(data + q);
        };
        auto load_rec = [&](int g) -> int4 {
            if (PREQ) return make_int4(0, 0, (int)__ldg(meta32 + g), 0);
            return __ldg(reinterpret_cast<const int4*>(ginfo + g));
        };
        int4 gi0 = make_int4(0, 0, 0, 0);
        float dat0 = 0.f, dat1 = 0.f, dat2 = 0.f;
        int idx1 = -1, idx2 = -1;
        {
            int idx0 = -1;
            if (p0 + lane < p1) { idx0 = __ldg(indices + p0 + lane); dat0 = load_val(p0 + lane); }
            if (p0 + 32 + lane < p1) { idx1 = __ldg(indices + p0 + 32 + lane); dat1 = load_val(p0 + 32 + lane); }
            if (idx0 >= 0) gi0 = load_rec(idx0);
        }
        for (int64_t base = p0;; base += 32) {
            const bool done = base >= p1;
            if (done) {
                // drain the slots, then the staged left-overs through the slots, before the last flush
                __syncwarp();
#pragma unroll
                for (int u = 0; u < D; ++u) apply(u);
#pragma unroll
                for (int u = 0; u < D; ++u) {
                    if ((uint32_t)u < n_left) {                         // warp-uniform
                        const uint2 it = items[u];
                        const uint32_t t = it.x * 32u + lane;
                        s_ids[u] = __ldg(p_col32 + t);
                        if (WEIGHTED) s_w[u] = __ldg(p_val2 + t);
                        s_x[u] = (int)it.y;
                    } else {
                        s_ids[u] = trash_ids;
                        s_x[u] = 0;
                    }
                }
#pragma unroll
                for (int u = 0; u < D; ++u) apply(u);
            }
            // this chunk: quantise, bound
            const uint32_t meta = (uint32_t)gi0.z, nblk = done ? 0u : (meta & 255u), blk0 = meta >> 8;
            uint32_t qb = 0;
            int xq = 0;      // binary: the quantised entry; weighted: the bits of the scaled entry (float)
            if (PREQ) {
                if (nblk) {
                    xq = __float_as_int(dat0);
                    qb = (uint32_t)(xq < 0 ? -xq : xq);
                }
            } else if (nblk) {
                const double vs = (double)dat0 * __hiloint2double(gi0.y, gi0.x);
                if (WEIGHTED) {
                    const float vf = (float)vs;
                    xq = __float_as_int(vf);
                    qb = __float2uint_ru(fabsf(vf) * gwmax) + 2u;
                } else {
                    xq = __double2int_rn(vs);
                    qb = (uint32_t)(xq < 0 ? -xq : xq);
                }
            }
            // Overflow guard without a warp reduction: every lane sums the bounds of its own entries; all
            // lanes stay below 1/32 of the budget, so the sum over lanes (the true bound) stays below it.
            // The one flush site: words -> totals, the low 12 bits stay in the word.  Lane l owns the
            // columns j = 32 t + l (A in bank l, B in bank (l + t) % 32).
            if (done || __any_sync(0xffffffffu, lbound + qb >= kFxFlushAt / 32u)) {
#pragma unroll
                for (int t = 0; t < kFxNB / 32; ++t) {
                    if (t * 32 < ncol) {
                        const int ia = t * 32 + lane, ib = kFxNB + t * 32 + ((lane + t) & 31);
                        const int a = (int)accw[ia], b = (int)accw[ib];
                        accw[ia] = (uint32_t)(a & 0xfff);
                        accw[ib] = (uint32_t)(b & 0xfff);
                        hi[t] += (a >> 12) + (b >> 12);
                    }
                }
                lbound = 2u * kFxInflight / 32u + 4096u / 32u + 1u;   // slot and staged items land after the flush; words keep 12 bits
            }
            if (done) break;
            lbound += qb;

            // stage the loads of the chunks after this one
            int4 gi1 = make_int4(0, 0, 0, 0);
            if (idx1 >= 0) gi1 = load_rec(idx1);
            idx2 = -1;
            dat2 = 0.f;
            if (base + 64 + lane < p1) { idx2 = __ldg(indices + base + 64 + lane); dat2 = load_val(base + 64 + lane); }

            for (uint32_t pg = 0;; pg += 4) {
                // compact the items of passes pg .. pg+3 into shared memory
                uint32_t m = __ballot_sync(0xffffffffu, nblk > pg);
                if (m == 0u) break;
                __syncwarp();
                uint32_t n_items = n_left;
#pragma unroll
                for (uint32_t dp = 0; dp < 4; ++dp) {
                    const uint32_t pass = pg + dp;
                    if (dp > 0) m = __ballot_sync(0xffffffffu, nblk > pass);
                    if (nblk > pass) items[n_items + __popc(m & lt_mask)] = make_uint2(blk0 + pass, (uint32_t)xq);
                    n_items += __popc(m);
                }
                // whole rounds of D now; the remainder waits (no bubbles inside a row)
                n_left = n_items % (uint32_t)D;
                n_items -= n_left;
                __syncwarp();
#pragma unroll 1
                for (uint32_t i = 0; i < n_items; i += D) {
                    uint2 itn = items[i];                                // broadcast
#pragma unroll
                    for (int u = 0; u < D; ++u) {
                        const uint2 it = itn;
                        if (u + 1 < D) itn = items[i + u + 1];           // in flight while slot u is applied
                        apply(u);
                        const uint32_t t = it.x * 32u + lane;
                        s_ids[u] = fx_ldg_pinned(p_col32 + t);           // D items before its use
                        if (WEIGHTED) s_w[u] = __ldg(p_val2 + t);
                        s_x[u] = (int)it.y;
                    }
                }
                if (n_left != 0u && n_items != 0u) {                     // warp-uniform
                    uint2 keep = make_uint2(0u, 0u);
                    if (lane < n_left) keep = items[n_items + lane];
                    __syncwarp();
                    if (lane < n_left) items[lane] = keep;
                }
            }
            gi0 = gi1;
            dat0 = dat1;
            idx1 = idx2;
            dat1 = dat2;
        }
        // Fold the totals into the words (A: total >> 12, B: the low 12 bits), so that the epilogue can
        // walk the columns with a rolled loop (dynamic shared addresses instead of 32 unrolled copies).
#pragma unroll
        for (int t = 0; t < kFxNB / 32; ++t) {
            if (t * 32 < ncol) {
                const int ia = t * 32 + lane, ib = kFxNB + t * 32 + ((lane + t) & 31);
                const int lo = (int)accw[ia] + (int)accw[ib];      // two 12-bit residues
                accw[ia] = (uint32_t)(hi[t] + (lo >> 12));
                accw[ib] = (uint32_t)(lo & 0xfff);
            }
        }
        // epilogue: row reference, then column scale, covariate terms and deviations in one pass.
        // The reference only has to lie near the cell's mean raw score (consumers use ref + deviation):
        // it is estimated from the integer totals and the column means of the per-set constants, which
        // saves a second pass over those constants.
        long long isum = 0;
#pragma unroll 1
        for (int t = 0; t * 32 < ncol; ++t) {
            const int j = t * 32 + lane;
            if (j < ncol)
                isum += ((long long)(int)accw[t * 32 + lane] << 12) + (long long)accw[kFxNB + t * 32 + ((lane + t) & 31)];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) isum += __shfl_xor_sync(0xffffffffu, isum, o);
        double ref;
        if (write_ref) {
            ref = ((double)isum * inv_scale) * scal[0] / (double)ncol;
            if (k > 0) {
                double cterm = scal[1];
                for (int kk = 0; kk < k; ++kk) cterm = fma(cov_mat[row * k + kk], scal[2 + kk], cterm);
                ref += cterm;
            }
        } else {
            ref = row_ref[row];       // later column tiles share the reference the first tile chose
        }
        float* out = dev_mat + row * (int64_t)ld;
#pragma unroll 1
        for (int t = 0; t * 32 < ncol; ++t) {
            const int j = t * 32 + lane;
            const int ia = t * 32 + lane, ib = kFxNB + t * 32 + ((lane + t) & 31);
            if (j < ncol) {
                const long long total = ((long long)(int)accw[ia] << 12) + (long long)accw[ib];
                double v = ((double)total * inv_scale) * colscale[j];
                if (k > 0) {
                    double cterm = covm[j];
                    for (int kk = 0; kk < k; ++kk)
                        cterm = fma(cov_mat[row * k + kk], covM[(size_t)kk * cov_stride + j], cterm);
                    v += cterm;
                }
                out[j] = (v == 0.0) ? __int_as_float(0x7fc00000) : (float)(v - ref);
            }
            accw[ia] = 0u;
            accw[ib] = 0u;
        }
        if (lane == 0 && write_ref) row_ref[row] = ref;
    }
}

constexpr int kFxDepthBinary = 8, kFxDepthWeighted = 4;

int fx_spmm(scdrs_ctx* c) {
    SCDRS_CHECK(c->w_fx && c->ncol <= kFxMaxCol * kFxMaxTiles, "fixed-point lists were not built");
    SCDRS_TRY(c->dev_mat.reserve((size_t)c->n_cell * c->ld * sizeof(float)));
    SCDRS_TRY(c->row_ref.reserve((size_t)c->n_cell * sizeof(double)));
    const int wpb = kFxWarpsPerBlock, per_sm = kFxBlocksPerSM;
    const size_t smem = (size_t)kFxWarpBytes * wpb;
    static bool attr_set = false;
    if (!attr_set) {
        SCDRS_CUDA(cudaFuncSetAttribute(spmm_score_fx_kernel<true, kFxDepthWeighted, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SCDRS_CUDA(cudaFuncSetAttribute(spmm_score_fx_kernel<false, kFxDepthBinary, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SCDRS_CUDA(cudaFuncSetAttribute(spmm_score_fx_kernel<false, kFxDepthBinary, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int64_t grid = (int64_t)kNumSM * per_sm;
    const int64_t need = (c->n_cell + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    const double inv_scale = ldexp(1.0, -c->fx_shift);
    SCDRS_TRY(c->fx_rowctr.reserve((size_t)kFxMaxTiles * sizeof(unsigned long long)));
    SCDRS_CUDA(cudaMemsetAsync(c->fx_rowctr.p, 0, (size_t)kFxMaxTiles * sizeof(unsigned long long), c->stream));
    if (c->timed_spmm) SCDRS_CUDA(cudaEventRecord(c->ev[0], c->stream));
    for (int tile = 0; tile < c->fx_n_tile; ++tile) {
        // with several column tiles the gene-major lists of a tile are built right before its pass over X
        if (!(tile == 0 && c->fx_tile0_built)) SCDRS_TRY(fx_build_tile(c, tile));
        const int col0 = tile * kFxMaxCol;
        const int ncol_t = (col0 + kFxMaxCol < c->ncol ? col0 + kFxMaxCol : c->ncol) - col0;
        if (c->w_binary && c->fx_preq)
            spmm_score_fx_kernel<false, kFxDepthBinary, true><<<(unsigned)grid, wpb * 32, smem, c->stream>>>(
                c->n_cell, ncol_t, c->ncol, tile == 0 ? 1 : 0, c->ld, c->indptr, c->indices, c->data,
                c->fx_xq.as<int32_t>(), c->fx_meta32.as<uint32_t>(), c->p_meta.as<GeneInfo>(), c->p_col.as<uint32_t>(),
                nullptr, c->colscale.as<double>() + col0, inv_scale, c->fx_gwmax, c->k, c->cov_mat.as<double>(),
                c->covM.as<double>() + col0, c->covm.as<double>() + col0, c->fx_scal.as<double>(),
                c->dev_mat.as<float>() + col0, c->row_ref.as<double>(), c->fx_rowctr.as<unsigned long long>() + tile);
        else if (c->w_binary)
            spmm_score_fx_kernel<false, kFxDepthBinary, false><<<(unsigned)grid, wpb * 32, smem, c->stream>>>(
                c->n_cell, ncol_t, c->ncol, tile == 0 ? 1 : 0, c->ld, c->indptr, c->indices, c->data, nullptr, nullptr,
                c->p_meta.as<GeneInfo>(), c->p_col.as<uint32_t>(), nullptr, c->colscale.as<double>() + col0, inv_scale,
                c->fx_gwmax, c->k, c->cov_mat.as<double>(), c->covM.as<double>() + col0, c->covm.as<double>() + col0,
                c->fx_scal.as<double>(), c->dev_mat.as<float>() + col0, c->row_ref.as<double>(),
                c->fx_rowctr.as<unsigned long long>() + tile);
        else
            spmm_score_fx_kernel<true, kFxDepthWeighted, false><<<(unsigned)grid, wpb * 32, smem, c->stream>>>(
                c->n_cell, ncol_t, c->ncol, tile == 0 ? 1 : 0, c->ld, c->indptr, c->indices, c->data, nullptr, nullptr,
                c->p_meta.as<GeneInfo>(), c->p_col.as<uint32_t>(), c->p_val.as<float2>(), c->colscale.as<double>() + col0,
                inv_scale, c->fx_gwmax, c->k, c->cov_mat.as<double>(), c->covM.as<double>() + col0,
                c->covm.as<double>() + col0, c->fx_scal.as<double>(), c->dev_mat.as<float>() + col0,
                c->row_ref.as<double>(), c->fx_rowctr.as<unsigned long long>() + tile);
        SCDRS_LAUNCH_CHECK(c);
    }
    if (c->fx_n_tile > 1) c->fx_tile0_built = false;     // the lists now belong to the last tile
    if (c->timed_spmm) SCDRS_CUDA(cudaEventRecord(c->ev[1], c->stream));
    c->trait_ready = true;
    return 0;
}

}  // namespace scdrs
