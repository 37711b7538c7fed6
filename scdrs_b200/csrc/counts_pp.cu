// The step right before the scoring path: what scdrs.util.load_h5ad does to a raw count matrix
// (scdrs/util.py:72-91 of the reference, which calls scanpy): NaN / negative checks, filter_cells
// (cells with fewer than min_genes stored entries), filter_genes (genes stored in fewer than min_cells of
// the remaining cells), normalize_per_cell(counts_per_cell_after = 1e4) (drops cells with a total below 1,
// scales every cell to 1e4) and log1p -- on the CSR matrix resident on the GPU, which then stays resident
// for preprocess / score_cell.  On the host these are scipy format conversions and fancy slices of a
// 2e8-entry matrix (seconds); here they are five streaming kernels and two prefix sums.
//
// Integer outputs (which cells and genes survive, entries per cell / gene) are exact; the float32
// arithmetic is the host's (x * float32(1e4 / n_counts), log1p in float32).
#include <math.h>

#include "common.cuh"

namespace scdrs {

__global__ void __launch_bounds__(256)
cp_check_kernel(int64_t nnz, const float* __restrict__ data, int* __restrict__ flags) {
    int f = 0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
        const float x = data[q];
        if (x != x) f |= 1;
        if (x < 0.f) f |= 2;
    }
    if (f) atomicOr(flags, f);
}

// stored entries per cell and the first cell filter
__global__ void cp_row_kernel(int64_t n_cell, const int64_t* __restrict__ indptr, int do_filter, int min_genes,
                              int32_t* __restrict__ n_genes, uint8_t* __restrict__ keep_cell) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_cell) return;
    const int64_t n = indptr[r + 1] - indptr[r];
    n_genes[r] = (int32_t)n;
    keep_cell[r] = (!do_filter || n >= min_genes) ? 1 : 0;
}

// cells per gene over the kept cells (integer atomics: exact, order independent); one warp per cell
__global__ void __launch_bounds__(256)
cp_gene_count_kernel(int64_t n_cell, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                     const uint8_t* __restrict__ keep_cell, uint32_t* __restrict__ n_cells) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_cell; r += nwarp) {
        if (!keep_cell[r]) continue;
        for (int64_t q = indptr[r] + lane; q < indptr[r + 1]; q += 32) atomicAdd(&n_cells[indices[q]], 1u);
    }
}

__global__ void cp_gene_keep_kernel(int n_gene, int do_filter, int min_cells, const uint32_t* __restrict__ n_cells,
                                    uint8_t* __restrict__ keep_gene, uint32_t* __restrict__ keep_gene32) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_gene) return;
    const uint32_t k = (g < n_gene && (!do_filter || n_cells[g] >= (uint32_t)min_cells)) ? 1u : 0u;
    if (g < n_gene) keep_gene[g] = (uint8_t)k;
    keep_gene32[g] = k;      // [n_gene + 1], last = 0: input of the exclusive scan that renumbers the genes
}

// per kept cell: surviving entries and their total (float32, in storage order like a row sum of the CSR
// matrix); with normalisation, cells whose total is below 1 go too (scanpy's min_counts = 1)
__global__ void __launch_bounds__(256)
cp_row_sum_kernel(int64_t n_cell, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const float* __restrict__ data, const uint8_t* __restrict__ keep_gene, int do_norm,
                  uint8_t* __restrict__ keep_cell, float* __restrict__ n_counts,
                  unsigned long long* __restrict__ new_len, uint32_t* __restrict__ keep_cell32) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_cell; r += nwarp) {
        unsigned long long len = 0;
        float total = 0.f;
        uint8_t keep = keep_cell[r];
        if (keep) {
            double s = 0.0;
            uint32_t cnt = 0;
            for (int64_t q = indptr[r] + lane; q < indptr[r + 1]; q += 32) {
                if (keep_gene[indices[q]]) {
                    s += (double)data[q];
                    ++cnt;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            }
            total = (float)s;
            if (do_norm && !(total >= 1.0f)) keep = 0;
            len = keep ? cnt : 0;
        }
        if (lane == 0) {
            keep_cell[r] = keep;
            keep_cell32[r] = keep;
            n_counts[r] = total;
            new_len[r] = len;
        }
    }
    if (warp0 == 0 && lane == 0) {
        keep_cell32[n_cell] = 0;
        new_len[n_cell] = 0;
    }
}

// compact, renumber, normalise: one warp per kept cell, entries keep their order
__global__ void __launch_bounds__(256)
cp_write_kernel(int64_t n_cell, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const float* __restrict__ data, const uint8_t* __restrict__ keep_cell,
                const uint8_t* __restrict__ keep_gene, const uint32_t* __restrict__ gene_new,
                const uint32_t* __restrict__ cell_new, const unsigned long long* __restrict__ new_ptr,
                const float* __restrict__ n_counts, int do_norm, float target, int64_t* __restrict__ out_indptr,
                int32_t* __restrict__ out_indices, float* __restrict__ out_data) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_cell; r += nwarp) {
        if (!keep_cell[r]) continue;
        unsigned long long pos = new_ptr[r];
        if (lane == 0) out_indptr[cell_new[r]] = (int64_t)pos;
        const float scale = do_norm ? __fdiv_rn(target, n_counts[r]) : 1.f;
        const int64_t p0 = indptr[r], p1 = indptr[r + 1];
        for (int64_t base = p0; base < p1; base += 32) {
            const int64_t q = base + lane;
            int g = 0;
            float x = 0.f;
            bool ok = false;
            if (q < p1) {
                g = indices[q];
                x = data[q];
                ok = keep_gene[g] != 0;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const unsigned long long dst = pos + __popc(m & lt);
                out_indices[dst] = (int32_t)gene_new[g];
                out_data[dst] = do_norm ? log1pf(__fmul_rn(x, scale)) : x;
            }
            pos += __popc(m);
        }
    }
}

__global__ void cp_tail_kernel(int64_t n_cell_new, const unsigned long long* __restrict__ new_ptr, int64_t n_cell,
                               int64_t* __restrict__ out_indptr) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out_indptr[n_cell_new] = (int64_t)new_ptr[n_cell];
}

// On success the context's resident matrix is the processed one (owned by the context).
int counts_preprocess(scdrs_ctx* c, int do_filter, int min_genes, int min_cells, int do_norm, double target_sum,
                      int* has_nan, int* has_negative, int64_t* n_cell_out, int32_t* n_gene_out, int64_t* nnz_out,
                      uint8_t* h_keep_cell, uint8_t* h_keep_gene, int32_t* h_n_genes, int32_t* h_n_cells,
                      float* h_n_counts) {
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set (call scdrs_set_csr_* first)");
    const int64_t n_cell = c->n_cell, nnz = c->nnz;
    const int n_gene = c->n_gene;
    DevBuf flags, keep_cell, keep_gene, n_genes, n_cells, keep_gene32, keep_cell32, n_counts, new_len, gene_new,
        cell_new, new_ptr, o_indptr, o_indices, o_data;
    DevBuf* tmp[] = {&flags, &keep_cell, &keep_gene, &n_genes, &n_cells, &keep_gene32, &keep_cell32, &n_counts,
                     &new_len, &gene_new, &cell_new, &new_ptr};
    auto cleanup = [&]() { for (DevBuf* b : tmp) b->release(); };
    auto fail = [&]() { cleanup(); o_indptr.release(); o_indices.release(); o_data.release(); return 1; };
#define CP_TRY(expr) do { if ((expr) != 0) return fail(); } while (0)
#define CP_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e__)); return fail(); } } while (0)
    int64_t grid = (n_cell + 7) / 8;
    if (grid > kNumSM * 16) grid = kNumSM * 16;
    if (grid < 1) grid = 1;
    // 1. checks (scdrs/util.py:72-84)
    CP_TRY(flags.reserve(sizeof(int)));
    CP_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int), c->stream));
    if (nnz > 0) {
        cp_check_kernel<<<kNumSM * 8, 256, 0, c->stream>>>(nnz, c->data, flags.as<int>());
        c->launches++;
    }
    int h_flags = 0;
    CP_CUDA(cudaMemcpyAsync(&h_flags, flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaStreamSynchronize(c->stream));
    *has_nan = h_flags & 1;
    *has_negative = (h_flags >> 1) & 1;
    *n_cell_out = n_cell;
    *n_gene_out = n_gene;
    *nnz_out = nnz;
    if (h_flags != 0 || (!do_filter && !do_norm)) {
        cleanup();
        return 0;
    }
    // 2. filters
    CP_TRY(keep_cell.reserve((size_t)n_cell));
    CP_TRY(keep_gene.reserve((size_t)n_gene));
    CP_TRY(n_genes.reserve((size_t)n_cell * sizeof(int32_t)));
    CP_TRY(n_cells.reserve((size_t)n_gene * sizeof(uint32_t)));
    CP_TRY(keep_gene32.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    CP_TRY(keep_cell32.reserve((size_t)(n_cell + 1) * sizeof(uint32_t)));
    CP_TRY(n_counts.reserve((size_t)n_cell * sizeof(float)));
    CP_TRY(new_len.reserve((size_t)(n_cell + 1) * sizeof(unsigned long long)));
    CP_TRY(gene_new.reserve((size_t)(n_gene + 1) * sizeof(uint32_t)));
    CP_TRY(cell_new.reserve((size_t)(n_cell + 1) * sizeof(uint32_t)));
    CP_TRY(new_ptr.reserve((size_t)(n_cell + 1) * sizeof(unsigned long long)));
    cp_row_kernel<<<(unsigned)((n_cell + 255) / 256), 256, 0, c->stream>>>(n_cell, c->indptr, do_filter, min_genes,
                                                                          n_genes.as<int32_t>(), keep_cell.as<uint8_t>());
    CP_CUDA(cudaMemsetAsync(n_cells.p, 0, (size_t)n_gene * sizeof(uint32_t), c->stream));
    cp_gene_count_kernel<<<(unsigned)grid, 256, 0, c->stream>>>(n_cell, c->indptr, c->indices, keep_cell.as<uint8_t>(),
                                                               n_cells.as<uint32_t>());
    cp_gene_keep_kernel<<<(n_gene + 256) / 256, 256, 0, c->stream>>>(n_gene, do_filter, min_cells, n_cells.as<uint32_t>(),
                                                                    keep_gene.as<uint8_t>(), keep_gene32.as<uint32_t>());
    cp_row_sum_kernel<<<(unsigned)grid, 256, 0, c->stream>>>(n_cell, c->indptr, c->indices, c->data, keep_gene.as<uint8_t>(),
                                                            do_norm, keep_cell.as<uint8_t>(), n_counts.as<float>(),
                                                            new_len.as<unsigned long long>(), keep_cell32.as<uint32_t>());
    c->launches += 4;
    CP_CUDA(cudaGetLastError());
    // 3. renumbering: exclusive scans of the keep flags and of the new row lengths
    CP_TRY(exclusive_scan_u32(c, keep_gene32.as<uint32_t>(), gene_new.as<uint32_t>(), n_gene + 1));
    CP_TRY(exclusive_scan_u32(c, keep_cell32.as<uint32_t>(), cell_new.as<uint32_t>(), n_cell + 1));
    CP_TRY(exclusive_scan_u64(c, new_len.as<unsigned long long>(), new_ptr.as<unsigned long long>(), n_cell + 1));
    uint32_t h_ng = 0, h_nc = 0;
    unsigned long long h_nnz = 0;
    CP_CUDA(cudaMemcpyAsync(&h_ng, gene_new.as<uint32_t>() + n_gene, 4, cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaMemcpyAsync(&h_nc, cell_new.as<uint32_t>() + n_cell, 4, cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaMemcpyAsync(&h_nnz, new_ptr.as<unsigned long long>() + n_cell, 8, cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaStreamSynchronize(c->stream));
    // 4. the new matrix
    CP_TRY(o_indptr.reserve((size_t)(h_nc + 1) * sizeof(int64_t)));
    CP_TRY(o_indices.reserve((size_t)h_nnz * sizeof(int32_t)));
    CP_TRY(o_data.reserve((size_t)h_nnz * sizeof(float)));
    cp_write_kernel<<<(unsigned)grid, 256, 0, c->stream>>>(
        n_cell, c->indptr, c->indices, c->data, keep_cell.as<uint8_t>(), keep_gene.as<uint8_t>(), gene_new.as<uint32_t>(),
        cell_new.as<uint32_t>(), new_ptr.as<unsigned long long>(), n_counts.as<float>(), do_norm, (float)target_sum,
        o_indptr.as<int64_t>(), o_indices.as<int32_t>(), o_data.as<float>());
    cp_tail_kernel<<<1, 32, 0, c->stream>>>((int64_t)h_nc, new_ptr.as<unsigned long long>(), n_cell, o_indptr.as<int64_t>());
    c->launches += 2;
    CP_CUDA(cudaGetLastError());
    CP_CUDA(cudaMemcpyAsync(h_keep_cell, keep_cell.p, (size_t)n_cell, cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaMemcpyAsync(h_keep_gene, keep_gene.p, (size_t)n_gene, cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaMemcpyAsync(h_n_genes, n_genes.p, (size_t)n_cell * 4, cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaMemcpyAsync(h_n_cells, n_cells.p, (size_t)n_gene * 4, cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaMemcpyAsync(h_n_counts, n_counts.p, (size_t)n_cell * 4, cudaMemcpyDeviceToHost, c->stream));
    CP_CUDA(cudaStreamSynchronize(c->stream));
    // 5. the processed matrix becomes the resident one
    c->own_indptr.release();
    c->own_indices.release();
    c->own_data.release();
    c->own_indptr = o_indptr;
    c->own_indices = o_indices;
    c->own_data = o_data;
    c->indptr = c->own_indptr.as<int64_t>();
    c->indices = c->own_indices.as<int32_t>();
    c->data = c->own_data.as<float>();
    c->n_cell = (int64_t)h_nc;
    c->n_gene = (int32_t)h_ng;
    c->nnz = (int64_t)h_nnz;
    c->trait_ready = false;
    c->have_stats = false;
    c->fx_stat_key = -1;
    c->fx_xq_key = -1;
    c->tile_ptr_valid = false;
    c->xsum_valid = false;
    c->covsum_valid = false;
    c->colsum_ready = false;
    c->k = 0;
    *n_cell_out = c->n_cell;
    *n_gene_out = c->n_gene;
    *nnz_out = c->nnz;
    cleanup();
    return 0;
#undef CP_TRY
#undef CP_CUDA
}

}  // namespace scdrs
