// Internal declarations shared by the kernel translation units.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/scdrs_b200.h"

namespace scdrs {

void set_error(const char* fmt, ...);

#define SCDRS_CUDA(call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            scdrs::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,           \
                             cudaGetErrorString(e__));                                    \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

#define SCDRS_CHECK(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            scdrs::set_error(__VA_ARGS__);     \
            return 1;                          \
        }                                      \
    } while (0)

#define SCDRS_TRY(expr)        \
    do {                       \
        if ((expr) != 0) return 1; \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int kNumSM = 148;  // B200

}  // namespace scdrs

struct scdrs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool timed_spmm = false, timed_trait = false;
    // The list building of trait i + 1 (prep_sets ... fx_place, ~0.3 ms of small kernels) runs on a second stream
    // while the main stream still normalises and counts trait i: it starts when the scatter of trait i has finished
    // (ev_scatter) and the scatter of trait i + 1 waits for it (ev_prep).  Scratch that both sides use exists twice
    // (swapped in by PrepScope, spmm_score.cu); what the tail of trait i reads and the preparation of trait i + 1
    // writes (ratio_a, colsum_alg) alternates between two buffers.
    cudaStream_t prep_stream = nullptr;
    cudaEvent_t ev_scatter = nullptr, ev_prep = nullptr, ev_mark = nullptr;
    bool prep_overlap = true;        // SCDRS_B200_PREP_OVERLAP=0 turns it off
    bool scatter_marked = false;     // ev_scatter has been recorded since the inputs last changed
    bool inputs_dirty = true;        // an API call other than the scoring family ran since the last preparation
    // SCDRS_B200_TIMELINE=1: events around the preparation and the scatter of the last 16 traits, printed (relative
    // to the first) by scdrs_ctx_sync -- where the device spends the time between two scatters
    struct Timeline {
        bool on = false;
        int n = 0;
        cudaEvent_t prep0[16], prep1[16], sc0[16], sc1[16];
    } tl;
    bool resident_sets = false;      // device-side control sets are complete long before they are handed over

    // resident expression matrix (CSR)
    int64_t n_cell = 0, nnz = 0;
    int32_t n_gene = 0;
    const int64_t* indptr = nullptr;
    const int32_t* indices = nullptr;
    const float* data = nullptr;
    scdrs::DevBuf own_indptr, own_indices, own_data;
    scdrs::DevBuf tile_ptr;        // where every gene tile starts inside every CSR row (stats.cu)
    bool tile_ptr_valid = false;
    // column sums of the raw scores by algebra: sum_c raw[c, j] = (sum_c X[c, :]) . w_j + covariate terms
    scdrs::DevBuf gene_xsum, cov_colsum, colsum_alg, fx_scal;
    bool xsum_valid = false, covsum_valid = false, colsum_ready = false;

    // gene-level vectors
    scdrs::DevBuf var, var_tech;
    bool have_stats = false;
    int32_t k = 0;  // number of covariates (0 = normal mode)
    scdrs::DevBuf cov_mat, cov_beta, gene_mean;

    // pipelined scoring (scdrs_score_trait_submit / _collect): two slots of pinned host staging
    struct Slot {
        void* pin_in = nullptr;
        size_t in_cap = 0;
        void* pin_out = nullptr;
        size_t out_cap = 0;
        cudaEvent_t done = nullptr;
        bool busy = false;
        int32_t n_ctrl = 0;
        bool ctrl_raw = false, ctrl_norm = false;
        void* pin_text = nullptr;       // text slots of the normalised control scores (score_format.cu)
        size_t text_cap = 0;
        bool text = false;
    } slot[2];
    int next_slot = 0;
    bool finish_async = false;      // scdrs_trait_finish leaves the final stream synchronisation to the caller
    bool text_output = false;       // submit: the control norm scores leave the device as text slots, not floats
    bool force_norm_mat = false;    // keep the normalised control matrix on the device although nobody downloads it
    scdrs::DevBuf text_slots;

    // downstream analyses (downstream.cu): the score matrix of one trait, [ds_n x ds_ncol] fp64 row-major
    scdrs::DevBuf ds_S, ds_tmp, ds_pos, ds_M, ds_keys, ds_vals, ds_off, ds_ref, ds_chunk, ds_part;
    int64_t ds_n = 0;
    int32_t ds_ncol = 0;
    bool ds_f32 = false;      // the caller's frame was float32 (numpy interpolates quantiles of float32 differently)

    // per-trait state
    int32_t n_ctrl = 0, n_s = 0, n_sc = 0, ncol = 0, ld = 0;  // n_s: trait set size, n_sc: control set size
    bool trait_ready = false;
    scdrs::DevBuf set_genes, set_gw, set_w, covM, covm, ratio_a;  // ratio_a: 1/sqrt(var ratio)
    scdrs::DevBuf ratio_a_alt, colsum_alg_alt, rnum, stage_prep, scan_tmp_prep;   // see prep_stream
    scdrs::DevBuf w_ptr, w_cursor, w_col, w_val;        // compact gene-major lists (build scratch)
    scdrs::DevBuf p_ptr, p_meta, p_col, p_val, colscale, gbase;  // padded / bank-dealt lists the scatter reads
    bool w_binary = false;
    // fixed-point scatter (spmm_score_fx.cu)
    int spmm_mode = 0;        // 0 auto, 1 fp64 accumulators, 2 fixed point (error if not applicable)
    bool w_fx = false;        // the lists currently built are laid out for the fixed-point kernel
    int fx_shift = 0;         // scale = 2^fx_shift
    int fx_n_tile = 1;        // column tiles of the fixed-point scatter
    bool fx_preq = false;     // binary trait: the scatter reads pre-quantised entries (fx_xq) and 4-byte records
    int fx_xq_key = -1, fx_xq_shift = 0;   // weight_opt / scale the pre-quantised entries were built for
    scdrs::DevBuf fx_xq, fx_meta32;
    bool fx_tile0_built = false;
    int fx_stat_key = -1;     // weight_opt the value statistics below were computed for (-1: none)
    double fx_vmax = 0.0, fx_vmean = 0.0;   // max and mean of |x * base_g| over the stored entries
    float fx_gwmax = 1.f;
    scdrs::DevBuf w_code, w_rounds, fx_part, fx_rowctr;
    scdrs::DevBuf dev_mat;   // float [n_cell x ld]: raw score minus the row reference; NaN = exact 0
    scdrs::DevBuf row_ref;   // double [n_cell]
    scdrs::DevBuf partial;   // double scratch for per-CTA column partials
    scdrs::DevBuf colmean, col2, colmin;  // double [ncol]
    scdrs::DevBuf row_mean, row_istd;     // double [n_cell]
    scdrs::DevBuf scalars;                // double [8]: gmin, ...
    scdrs::DevBuf query, mc_count, norm_mat;
    int64_t n_cell_total = 0;
    // pooled-null workspace
    scdrs::DevBuf sort_keys[2], sort_idx[2], sort_tab, grid, grid_cells, hist, ge_sorted, ge_local, scan_tmp;
    bool cells_ready = false;   // grid_cells holds the 8-byte cell table of the current sorted queries
    int32_t grid_bits = 0;  // number of grid cells
    int sorted_buf = 0;     // which sort buffer holds the sorted queries
    // single-GPU exchange buffers
    scdrs::DevBuf x_colsum, x_col2, x_colmin, x_hist;
    // staging
    scdrs::DevBuf stage;
};

namespace scdrs {

// ---- spmm_score.cu ------------------------------------------------------------------------
int trait_prepare(scdrs_ctx* c, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                  const int32_t* trait_genes, const double* trait_gw, const int32_t* ctrl_genes,
                  const double* ctrl_gw, int32_t weight_opt, int32_t use_var_ratio, bool ctrl_on_device = false);
int explicit_prepare(scdrs_ctx* c, int32_t n_set, int32_t n_s, const int32_t* genes,
                     const double* w);
int spmm_score(scdrs_ctx* c);

// per gene, fetched with ONE 16-byte load per stored entry of X: the base weight folded into the
// expression value and the packed location of the gene's set-id list
struct __align__(16) GeneInfo {
    double base;
    uint32_t meta;
    uint32_t pad;
};

// ---- spmm_score_fx.cu ---------------------------------------------------------------------
constexpr int kFxMaxCol = 1024;   // columns per pass: the fixed-point kernel keeps one total per column and lane in registers
constexpr int kFxMaxTiles = 32;   // wider score matrices are scattered in column tiles of kFxMaxCol sets, one pass over X each
bool fx_applicable(scdrs_ctx* c, int ncol, bool explicit_w);
int fx_value_stats(scdrs_ctx* c, int weight_opt);
int fx_prepare(scdrs_ctx* c, bool binary, double gw_absmax);
int fx_spmm(scdrs_ctx* c);
int build_gene_lists(scdrs_ctx* c, int col0, int col1);   // spmm_score.cu

// ---- background.cu ------------------------------------------------------------------------
int bg_colsum(scdrs_ctx* c, double* dev_colsum_out);
int bg_rowstats(scdrs_ctx* c, const double* dev_colsum, int64_t n_cell_total, double* dev_col2,
                double* dev_colmin);
int bg_queries(scdrs_ctx* c, const double* dev_col2, const double* dev_colmin, float* dev_query);
int bg_pack_dense(scdrs_ctx* c, int64_t n_cell, int32_t n_ctrl, const double* dev_vraw,
                  const double* dev_mat, const double* dev_ratio);
int bg_export_raw(scdrs_ctx* c, double* dev_raw_score, float* dev_ctrl_raw);

// ---- empi_null.cu -------------------------------------------------------------------------
int null_prepare_queries(scdrs_ctx* c, const float* dev_query_all, int64_t m);
int null_count_pass(scdrs_ctx* c, unsigned long long* dev_hist, int64_t m, bool write_norm);
int null_finish(scdrs_ctx* c, const unsigned long long* dev_hist_all, int64_t m, int64_t offset,
                int64_t n_null_total, int64_t n_local, long long* dev_ge_out);
int null_count_generic_f32(scdrs_ctx* c, int64_t n_t, const float* dev_t, int64_t n_null,
                           const float* dev_null, long long* dev_ge);
int null_count_generic_f64(scdrs_ctx* c, int64_t n_t, const double* dev_t, int64_t n_null,
                           const double* dev_null, long long* dev_ge);

// ---- stats.cu -----------------------------------------------------------------------------
int gene_cell_stats(scdrs_ctx* c, int transform, const double* dev_cell_w, int center, double denom,
                    double* dev_gsum, double* dev_gsumsq, double* dev_csum, double* dev_csumsq);

int gene_mean_xtc(scdrs_ctx* c, int k, const double* dev_cov, float* dev_gmean32, double* dev_xtc);
int gene_stats_exact(scdrs_ctx* c, int transform, const double* dev_cell_w, int center, double denom, int n_tile_sel,
                     const int32_t* dev_tile_ids, const double* dev_mean_in, double* dev_gsum, double* dev_gsumsq);
int gene_tile_size();

// ---- stats_par.cu ---------------------------------------------------------------------------
int gene_moments(scdrs_ctx* c, int mode, int k, const double* d_cov, const double* d_w, double* d_out);
int gene_expm1_implicit(scdrs_ctx* c, const double* d_shift, const double* d_w, double* d_out);

// ---- score_format.cu ------------------------------------------------------------------------
int format_text_slots(scdrs_ctx* c, int64_t n_val, const float* dev_vals, void* dev_slots);

// ---- counts_pp.cu ---------------------------------------------------------------------------
int counts_preprocess(scdrs_ctx* c, int do_filter, int min_genes, int min_cells, int do_norm, double target_sum,
                      int* has_nan, int* has_negative, int64_t* n_cell_out, int32_t* n_gene_out, int64_t* nnz_out,
                      uint8_t* h_keep_cell, uint8_t* h_keep_gene, int32_t* h_n_genes, int32_t* h_n_cells,
                      float* h_n_counts);

// ---- downstream.cu ----------------------------------------------------------------------------
int ds_import(scdrs_ctx* c, int64_t n, int ncol, int is_f32, int64_t rs, int64_t cs, const int32_t* dev_cols,
              const void* dev_src, double* dev_S);
int ds_corr(scdrs_ctx* c, int n_var, const double* dev_vars, double* dev_corr);
int ds_group_stats(scdrs_ctx* c, int n_group, const int64_t* h_gptr, const int64_t* d_order, const int64_t* d_gptr,
                   const int32_t* d_qlo, const double* d_qgamma, double* d_quant, int geary_mode,
                   const int64_t* d_gindptr, const int32_t* d_gnbr, const double* d_gw, double* d_total,
                   double* d_denom);

// ---- scan (empi_null.cu) --------------------------------------------------------------------
int exclusive_scan_u32(scdrs_ctx* c, const uint32_t* in, uint32_t* out, int64_t n);
int exclusive_scan_u64(scdrs_ctx* c, const unsigned long long* in, unsigned long long* out,
                       int64_t n);

}  // namespace scdrs

#define SCDRS_LAUNCH_CHECK(c)                      \
    do {                                           \
        (c)->launches++;                           \
        SCDRS_CUDA(cudaGetLastError());            \
    } while (0)
