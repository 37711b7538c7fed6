/*
 * scdrs_b200 -- C ABI of the B200 (sm_100a) scoring hot path of scDRS.
 *
 * The reference (martinjzhang/scDRS) is pure Python and has no FFI; the functions below are
 * the native boundary its hot path would bind (the ctypes stub a maintainer would add is shown
 * in INTEGRATION.md and is what scdrs_b200/_native.py uses).  Each entry point cites the
 * reference code it replaces, paths relative to the reference root.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on failure, and
 *     scdrs_last_error() then returns a thread-local message.  There is no CPU fallback.
 *   - "host" pointers are ordinary host memory (pinned or pageable); "dev" pointers are device
 *     memory on the context's GPU.  All work is enqueued on the context's stream; functions
 *     that fill host outputs synchronise that stream before returning.
 *   - matrices are row-major.  Column 0 of every (1 + n_ctrl)-column object is the trait gene
 *     set, columns 1..n_ctrl are the control gene sets.
 */
#ifndef SCDRS_B200_H
#define SCDRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scdrs_ctx scdrs_ctx;

/* weight_opt of scdrs/method.py:366-371 */
enum { SCDRS_WEIGHT_UNIFORM = 0, SCDRS_WEIGHT_VS = 1, SCDRS_WEIGHT_INV_STD = 2 };

const char* scdrs_last_error(void);
int scdrs_abi_version(void);

/* ---- context: one per (GPU, expression matrix) ------------------------------------------ */
int scdrs_ctx_create(int device, scdrs_ctx** out);
int scdrs_ctx_destroy(scdrs_ctx* ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the context's own */
int scdrs_ctx_set_stream(scdrs_ctx* ctx, void* cuda_stream);
int scdrs_ctx_sync(scdrs_ctx* ctx);
/* number of kernel launches issued through this context since creation (bench bookkeeping) */
int64_t scdrs_ctx_launch_count(scdrs_ctx* ctx);
/* device time (ms) of the last scdrs_trait_raw scatter kernel / whole scdrs_score_trait call,
 * measured with CUDA events on the context's stream; valid after a sync */
int scdrs_ctx_last_timing(scdrs_ctx* ctx, float* spmm_ms, float* trait_ms);

/* Arithmetic of the raw-score scatter (phase 1).  mode 0 (default): exact fixed-point integer sums
 * when the stored values have no extreme outliers (in column tiles of 1024 sets when
 * 1 + n_ctrl > 1024), fp64 accumulators otherwise; mode 1: always fp64; mode 2: fixed point or fail.  Both reproduce the reference's
 * float64 _compute_raw_score (scdrs/method.py:377-400) far inside the 1e-5 tolerance. */
int scdrs_ctx_set_spmm_mode(scdrs_ctx* ctx, int32_t mode);
/* which arithmetic the last prepared trait uses, and log2 of the fixed-point scale */
int scdrs_ctx_last_spmm_kind(scdrs_ctx* ctx, int32_t* fixed_point, int32_t* scale_log2);

/* ---- resident inputs ---------------------------------------------------------------------
 * adata.X as CSR (cells x genes): float32 values, int32 column ids, int64 row pointers.
 * Replaces the CSC copy of scdrs/method.py:98-99.  _host copies H2D, _dev borrows pointers. */
int scdrs_set_csr_host(scdrs_ctx* ctx, int64_t n_cell, int32_t n_gene, const int64_t* indptr,
                       const int32_t* indices, const float* data);
int scdrs_set_csr_dev(scdrs_ctx* ctx, int64_t n_cell, int32_t n_gene, const int64_t* dev_indptr,
                      const int32_t* dev_indices, const float* dev_data);
/* ---- the step before the path: load_h5ad's checks, filters and normalisation -----------------
 * scdrs/util.py:72-91 (NaN / negative checks; scanpy filter_cells(min_genes), filter_genes(min_cells),
 * normalize_per_cell(counts_per_cell_after = target_sum, cells with a total < 1 dropped), log1p) on the
 * resident CSR matrix, which is replaced by the processed one and stays resident.  has_nan / has_negative:
 * the reference raises ValueError for either; nothing is modified then.  With both flags 0 only the checks
 * run.  Host outputs, indexed by ORIGINAL cell / gene: keep_cell[n_cell], keep_gene[n_gene] (final
 * survivors), n_genes_per_cell[n_cell] (stored entries), n_cells_per_gene[n_gene] (over the cells that
 * passed min_genes), n_counts[n_cell] (total over the surviving genes; valid for cells that passed
 * min_genes).  n_cell_out / n_gene_out / nnz_out: shape of the matrix now resident. */
int scdrs_preprocess_counts(scdrs_ctx* ctx, int32_t flag_filter_data, int32_t min_genes, int32_t min_cells,
                            int32_t flag_raw_count, double target_sum, int32_t* has_nan, int32_t* has_negative,
                            int64_t* n_cell_out, int32_t* n_gene_out, int64_t* nnz_out, uint8_t* keep_cell,
                            uint8_t* keep_gene, int32_t* n_genes_per_cell, int32_t* n_cells_per_gene, float* n_counts);
/* download the resident CSR matrix (host arrays sized from the shape the context reports) */
int scdrs_get_csr(scdrs_ctx* ctx, int64_t* indptr, int32_t* indices, float* data);

/* GENE_STATS["var"], GENE_STATS["var_tech"] (scdrs/method.py:369-371, 192-195); host, n_gene each */
int scdrs_set_gene_stats(scdrs_ctx* ctx, const double* var, const double* var_tech);
/* implicit covariate correction terms COV_MAT (n_cell x k), COV_BETA (n_gene x k, = -beta) and
 * COV_GENE_MEAN (n_gene) of scdrs/method.py:377-395; host pointers; k = 0 switches it off */
int scdrs_set_cov(scdrs_ctx* ctx, int32_t k, const double* cov_mat, const double* cov_beta,
                  const double* gene_mean);

/* ---- one trait, phase by phase (the multi-GPU driver runs collectives between phases) ------
 * Phase 1 = _compute_raw_score x (1 + n_ctrl), scdrs/method.py:172-183, 312-402, plus the
 * variance ratio of :186-195.  trait_genes[n_s] / trait_gw[n_s]: trait gene columns and their
 * gene_weight in gene-name order; ctrl_genes[n_ctrl x n_s_ctrl]: control gene columns;
 * ctrl_gw[n_s_ctrl]: the gene_weight each control position inherits (scdrs/method.py:305-307).
 * n_s_ctrl == n_s unless some trait genes fall in no bin (NaN match key).  Host pointers.
 * dev_colsum_out[1 + n_ctrl] receives this rank's column sums of the raw-score matrix. */
int scdrs_trait_raw(scdrs_ctx* ctx, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                    const int32_t* trait_genes, const double* trait_gw, const int32_t* ctrl_genes,
                    const double* ctrl_gw, int32_t weight_opt, int32_t use_var_ratio,
                    double* dev_colsum_out);
/* Phase 2 = first gene-set alignment + cell-wise standardisation, scdrs/method.py:526-548.
 * dev_colsum: column sums over ALL ranks; n_cell_total: cells over all ranks.
 * Outputs (device, 1 + n_ctrl each): column sums and column minima of the standardised scores. */
int scdrs_trait_rowstats(scdrs_ctx* ctx, const double* dev_colsum, int64_t n_cell_total,
                         double* dev_colsum2_out, double* dev_colmin_out);
/* Phase 3 = second alignment + zero rule for the trait column, scdrs/method.py:564-582.
 * dev_colsum2 / dev_colmin: all-rank sums / minima.  dev_query_out[n_cell]: this rank's final
 * normalised trait scores (float32). */
int scdrs_trait_queries(scdrs_ctx* ctx, const double* dev_colsum2, const double* dev_colmin,
                        float* dev_query_out);
/* Phase 4 = mc_pval counts (scdrs/method.py:205) and the pooled empirical null
 * (_get_p_from_empi_null, scdrs/method.py:601-632): radix-sorts all ranks' queries, then
 * bucket-counts this rank's n_cell x n_ctrl null values between consecutive sorted queries.
 * dev_hist_out[n_query_all + 1] (uint64): dev_hist_out[p] = #{local null with exactly p queries <= it}. */
int scdrs_trait_count(scdrs_ctx* ctx, const float* dev_query_all, int64_t n_query_all,
                      unsigned long long* dev_hist_out);
/* Phase 5: suffix-sum the all-rank histogram into #{null >= t} and copy results to the host.
 * query_offset: position of this rank's first cell inside dev_query_all.  Host outputs:
 * raw_score[n_cell] f64, norm_score[n_cell] f32, mc_count[n_cell] i32 = #{ctrl_norm >= norm},
 * ge_count[n_cell] i64 = #{pooled null >= norm}; ctrl_raw / ctrl_norm [n_cell x n_ctrl] f32 or NULL. */
int scdrs_trait_finish(scdrs_ctx* ctx, const unsigned long long* dev_hist_all, int64_t n_query_all,
                       int64_t query_offset, int64_t n_null_total, double* raw_score,
                       float* norm_score, int32_t* mc_count, int64_t* ge_count, float* ctrl_raw,
                       float* ctrl_norm);

/* ---- one trait, single GPU: phases 1-5 back to back = scdrs.method.score_cell's device part */
int scdrs_score_trait(scdrs_ctx* ctx, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                      const int32_t* trait_genes, const double* trait_gw, const int32_t* ctrl_genes,
                      const double* ctrl_gw, int32_t weight_opt, int32_t use_var_ratio,
                      double* raw_score, float* norm_score, int32_t* mc_count, int64_t* ge_count,
                      float* ctrl_raw, float* ctrl_norm);

/* The same in two halves, for callers that score many traits (`scdrs compute-score` loops over a .gs
 * file, bin/scdrs:255-274): submit copies the gene sets into pinned memory, enqueues phases 1-5 and the
 * download of the results into pinned memory, and returns at once with a slot number (two slots: at most
 * two traits in flight); collect waits for that slot and copies the results out (NULL = skip; the control
 * matrices only if requested at submit).  The next trait's work is thus queued while the current one runs. */
int scdrs_score_trait_submit(scdrs_ctx* ctx, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                             const int32_t* trait_genes, const double* trait_gw,
                             const int32_t* ctrl_genes, const double* ctrl_gw, int32_t weight_opt,
                             int32_t use_var_ratio, int32_t want_ctrl_raw, int32_t want_ctrl_norm,
                             int32_t* slot_out);
int scdrs_score_trait_collect(scdrs_ctx* ctx, int32_t slot, double* raw_score, float* norm_score,
                              int32_t* mc_count, int64_t* ge_count, float* ctrl_raw, float* ctrl_norm);
/* Text output (the score-file writer's fast path): after scdrs_ctx_set_text_output(ctx, 1), a trait submitted with
 * want_ctrl_norm != 0 keeps its normalised control scores on the device and turns them there into the text pandas
 * writes for them -- one 16-byte slot per value, see scdrs_write_score_gz_slots -- instead of downloading n_cell x
 * n_ctrl floats for host threads to print.  After the trait was collected, scdrs_slot_text returns the pinned host
 * block [n_cell x n_ctrl x 16] of its slot; it stays valid until the next submit that takes this slot (the one
 * after next).  The float matrix is then not available from scdrs_score_trait_collect (pass ctrl_norm = NULL). */
int scdrs_ctx_set_text_output(scdrs_ctx* ctx, int32_t on);
/* The list building of a trait runs on a second stream of the context, concurrently with the normalisation and
 * null counting of the trait before it (it starts when that trait's raw-score scatter has finished).  Control sets
 * handed over in DEVICE memory (scdrs_trait_raw_dev) are by default assumed to be produced on the context's stream
 * right before the call (the receive buffer of a broadcast), which orders the preparation behind everything
 * enqueued so far; a caller whose device-side control sets were complete before the previous trait was submitted
 * (sets resident for the whole run) says so with on = 1 and gets the overlap there too. */
int scdrs_ctx_set_resident_sets(scdrs_ctx* ctx, int32_t on);
int scdrs_slot_text(scdrs_ctx* ctx, int32_t slot, const void** text, int64_t* n_bytes);

/* The same split for the phase-by-phase (multi-GPU) driver: phase 1 with the control sets already in
 * device memory (dev_ctrl_genes[n_ctrl x n_s_ctrl], e.g. the receive buffer of a broadcast; the small
 * host arrays are staged through pinned memory), and phase 5 into pinned memory behind an event --
 * collect with scdrs_score_trait_collect(slot). */
int scdrs_trait_raw_dev(scdrs_ctx* ctx, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                        const int32_t* trait_genes, const double* trait_gw,
                        const int32_t* dev_ctrl_genes, const double* ctrl_gw, int32_t weight_opt,
                        int32_t use_var_ratio, double* dev_colsum_out);
int scdrs_trait_finish_submit(scdrs_ctx* ctx, const unsigned long long* dev_hist_all,
                              int64_t n_query_all, int64_t query_offset, int64_t n_null_total,
                              int32_t want_ctrl_raw, int32_t want_ctrl_norm, int32_t* slot_out);

/* ---- the reference's helper functions as stand-alone calls (host buffers) ------------------ */
/* _compute_raw_score (scdrs/method.py:377-400) for n_set explicit gene sets with FINAL weights
 * w[n_set x n_s] (already normalised); raw_out[n_cell x n_set] f64. */
int scdrs_compute_raw_score(scdrs_ctx* ctx, int32_t n_set, int32_t n_s, const int32_t* genes,
                            const double* w, double* raw_out);
/* _correct_background (scdrs/method.py:482-598): v_raw[n_cell], mat_ctrl[n_cell x n_ctrl],
 * var_ratio[n_ctrl] -> v_norm[n_cell], mat_norm[n_cell x n_ctrl] (all f64 host) */
int scdrs_correct_background(scdrs_ctx* ctx, int64_t n_cell, int32_t n_ctrl, const double* v_raw,
                             const double* mat_ctrl, const double* var_ratio, double* v_norm,
                             double* mat_norm);
/* _get_p_from_empi_null (scdrs/method.py:601-632): ge_count[i] = #{null >= t[i]};
 * p = (ge_count + 1) / (n_null + 1).  f32 and f64 key variants. */
int scdrs_empi_null_count_f32(scdrs_ctx* ctx, int64_t n_t, const float* t, int64_t n_null,
                              const float* null_vals, int64_t* ge_count);
int scdrs_empi_null_count_f64(scdrs_ctx* ctx, int64_t n_t, const double* t, int64_t n_null,
                              const double* null_vals, int64_t* ge_count);
/* pp.compute_stats reductions (scdrs/pp.py:353-373, 409-417, 467-641) over the resident CSR,
 * of T(X + COV_MAT*COV_BETA + COV_GENE_MEAN) when covariates are set (implicit mode) else T(X);
 * T = identity (transform 0) or expm1 (transform 1).  cell_weight[n_cell] or NULL (weights 1).
 * Host outputs (f64; NULL pairs are skipped):
 *   gene_sum[n_gene]   = sum_c w_c T(v), accumulated cell by cell like numpy's axis-0 reduction;
 *   gene_sumsq[n_gene] = sum_c w_c T(v)^2                     when center == 0 (scdrs/pp.py:508-515)
 *                      = sum_c w_c (T(v) - gene_sum/denom)^2  when center == 1 (np.var, :501);
 *   cell_sum[n_cell], cell_sumsq[n_cell] = unweighted row sums of T(v) and T(v)^2. */
int scdrs_gene_cell_stats(scdrs_ctx* ctx, int32_t transform, const double* cell_weight,
                          int32_t center, double denom, double* gene_sum, double* gene_sumsq,
                          double* cell_sum, double* cell_sumsq);

/* Gene means and C^T X for the covariate regression of pp.preprocess (scdrs/pp.py:206, :209-212).
 * cov_mat[n_cell x k] (host f64, k <= 16).  Outputs (host): gene_mean[n_gene] f32 = X.mean(axis=0)
 * with scipy's float32 operation order; xtc[n_gene x k] f64 = (C^T X)^T accumulated in cell order. */
int scdrs_gene_mean_xtc(scdrs_ctx* ctx, int32_t k, const double* cov_mat, float* gene_mean, double* xtc);

/* The same reductions in parallel form: per-gene moments of THIS context's cells in one pass over the stored
 * entries, no dense block, bitwise reproducible (fixed-order sums, no atomics).  They are plain sums over cells,
 * so the moments of the shards of several GPUs add up (all-reduce) to those of the whole atlas -- the distributed
 * form of scdrs/pp.py:206-212 and :353-373.  cell_weight[n_cell] or NULL (weights 1); host pointers.
 *   mode 0: out[n_gene x (2 + k)] = sum w x, sum w x^2, sum w x cov_mat[c, j] (j < k; cov_mat[n_cell x k], k >= 0)
 *           -> gene means, C^T X of the regression, and -- with the k x k covariate moments, host algebra, see
 *           scdrs_b200/pp.py -- mean / var of X + COV_MAT*COV_BETA + COV_GENE_MEAN (_get_mean_var_implicit_cov_corr)
 *   mode 1: out[n_gene x 2] = sum w expm1(x), sum w expm1(x)^2   (normal mode, scdrs/pp.py:358-364) */
int scdrs_gene_moments(scdrs_ctx* ctx, int32_t mode, int32_t k, const double* cov_mat, const double* cell_weight,
                       double* out);
/* Implicit mode, expm1 scale (scdrs/pp.py:371-373): with v = X + COV_MAT*COV_BETA + COV_GENE_MEAN of the resident
 * covariate terms (scdrs_set_cov) and a per-gene shift K[n_gene] near the mean,
 *   out[n_gene x 2] = sum_c w (expm1(v) - K), sum_c w (expm1(v) - K)^2   over ALL cells of this context
 * (a dense pass over cells x genes plus what the stored entries change). */
int scdrs_gene_expm1_implicit(scdrs_ctx* ctx, const double* shift, const double* cell_weight, double* out);
/* scdrs_gene_cell_stats' per-gene sums, strictly in numpy's operation order, for the listed tiles of
 * scdrs_gene_tile_size() consecutive genes only (genes whose variance is rounding noise of that order).
 * mean_in[n_gene] != NULL with center == 1: only the centred pass, around the given means (sharded atlases).
 * gene_sum / gene_sumsq [n_gene]: entries outside the listed tiles are left untouched. */
int scdrs_gene_stats_exact(scdrs_ctx* ctx, int32_t transform, const double* cell_weight, int32_t center, double denom,
                           int32_t n_tile, const int32_t* tile_ids, const double* mean_in, double* gene_sum,
                           double* gene_sumsq);
int scdrs_gene_tile_size(void);

/* ---- downstream analyses of one trait's score matrix -------------
 * scdrs.method.downstream_corr_analysis / downstream_group_analysis / test_gearysc / gearys_c
 * (scdrs/method.py:712-865, 915-1085).  The caller uploads the trait's scores once, straight from the
 * memory of its score frame: `frame` holds n_elem values (float32 when is_f32, else float64); score
 * (r, j) is frame[r * row_stride + cols[j] * col_stride] -- col_stride 1 for a row-major frame,
 * row_stride 1 for a column-major one (what pandas usually holds).  cols[0] = the norm_score column,
 * cols[1..] = the ctrl_norm_score_i columns.  On the device the scores become fp64 [n_cell x ncol]; the
 * rows may be any subset / order of the cells, the other calls refer to them by row number. */
int scdrs_ds_set_scores(scdrs_ctx* ctx, int64_t n_cell, int32_t ncol, const void* frame, int32_t is_f32,
                        int64_t n_elem, int64_t row_stride, int64_t col_stride, const int32_t* cols);
/* np.corrcoef(var, score column)[0, 1] for every variable and column (scdrs/method.py:853-857):
 * vars[n_var x n_cell] host f64 -> corr[n_var x ncol] host f64. */
int scdrs_ds_corr(scdrs_ctx* ctx, int32_t n_var, const double* vars, double* corr);
/* Per cell group (order[n_cell]: cells sorted by group, the caller's order inside a group;
 * grp_ptr[n_group + 1]) and per score column, all outputs host f64 [n_group x ncol]:
 *   quantile (NULL = skip): np.quantile(column of the group, q), q_lo[g] / q_gamma[g] = numpy's lower
 *     index and interpolation weight for a group of that size (scdrs/method.py:797-801);
 *   geary_mode 0: nothing more.  1: every column is first given the distribution of column 0 inside
 *     the group by ordinal rank (control_distribution_match, scdrs/method.py:976-990); 2: columns as
 *     they are.  Then geary_total = sum over the stored within-group edges (p, q) of w (x_p - x_q)^2
 *     and geary_denom = sum_p (x_p - mean)^2 (scdrs/method.py:1070-1082; C = (N-1) total / (2 W denom)).
 *     Graph: CSR over cells in group order, g_nbr = neighbour positions in group order, only edges
 *     inside a group (g_indptr[n_cell + 1] i64, g_nbr i32, g_w f64). */
int scdrs_ds_group_stats(scdrs_ctx* ctx, int32_t n_group, const int64_t* order, const int64_t* grp_ptr,
                         const int32_t* q_lo, const double* q_gamma, double* quantile,
                         int32_t geary_mode, const int64_t* g_indptr, const int32_t* g_nbr,
                         const double* g_w, double* geary_total, double* geary_denom);

/* ---- host code: the random draws of _select_ctrl_geneset (scdrs/method.py:276, 295-307) --------
 * numpy-legacy MT19937 seeded with `seed`, bins visited in the given order; for every bin b with
 * bin_ntrait[b] > 0 and every control set i: permutation(len(bin b))[:bin_ntrait[b]] indexes
 * bin_genes[bin_ptr[b] .. bin_ptr[b+1]) (gene columns ordered by gene name).  Writes
 * ctrl_genes[n_ctrl x n_s]; bit-identical to np.random.seed(seed) + np.random.choice(...). */
int scdrs_select_ctrl(uint32_t seed, int32_t n_bin, const int32_t* bin_ptr, const int32_t* bin_genes,
                      const int32_t* bin_ntrait, int32_t n_ctrl, int32_t n_s, int32_t* ctrl_genes);
/* Same, and mt_state_out[625] (may be NULL) receives the generator after the draws -- numpy's key[624] and pos --
 * so that the caller can leave numpy's GLOBAL generator where the reference leaves it (the reference draws on
 * it, scdrs/method.py:96, 276, 301: callers observe that side effect). */
int scdrs_select_ctrl_state(uint32_t seed, int32_t n_bin, const int32_t* bin_ptr, const int32_t* bin_genes,
                            const int32_t* bin_ntrait, int32_t n_ctrl, int32_t n_s, int32_t* ctrl_genes,
                            uint32_t* mt_state_out);

/* ---- host code: score-file writer (bin/scdrs:276-288; format docs/file_format.rst:81-108) --------
 * Gzip TSV: `header` (one line, '\n'-terminated), then per row: name, then n_col float32 fields in
 * pandas' to_csv text (shortest round-trip; NaN -> empty).  names/name_off: concatenated UTF-8 row
 * names and their n_row + 1 offsets.  Row blocks are formatted and deflated by n_thread threads. */
int scdrs_write_score_gz(const char* path, const char* header, int64_t n_row, int32_t n_col,
                         const char* names, const int64_t* name_off, const float* values,
                         int32_t n_thread, int32_t level);
/* The same file when the GPU has already formatted the trailing n_slot_col columns (the normalised control scores):
 * lead[n_row x n_lead] float32 are printed by the host threads, slots[n_row x n_slot_col x 16] are the text slots of
 * scdrs_slot_text -- text in bytes 0 .. len-1 (at most 15), len in byte 15 -- and are only copied. */
int scdrs_write_score_gz_slots(const char* path, const char* header, int64_t n_row, const char* names,
                               const int64_t* name_off, int32_t n_lead, const float* lead, int32_t n_slot_col,
                               const uint8_t* slots, int32_t n_thread, int32_t level);
/* one float32 in that text format into out64 (NUL-terminated); returns its length */
int scdrs_format_f32(float v, char* out64);
/* The __host__ __device__ formatter the GPU uses for the control-score block of a .full_score.gz file, and its
 * self-test against scdrs_format_f32 over the float32 bit patterns [first, first + count) (returns the number of
 * mismatches; the whole 2^32 range takes about a minute on 32 threads). */
int scdrs_format_f32_shortest(float v, char* out64);
/* Host restatement of the device formatting kernel: slots[n_val x 16] from vals[n_val] (tests, timing probes). */
int scdrs_format_slots_host(int64_t n_val, const float* vals, uint8_t* slots, int32_t n_thread);
int64_t scdrs_selftest_format(uint32_t first, uint64_t count, int32_t n_thread, uint32_t* bad_bits);
/* One gzip member of `text` as the score-file writer produces it at `level` (1: the literal-only dynamic-Huffman
 * encoder of score_writer.cu, 0 / 2..9: zlib) into out[0 .. cap); *out_len = bytes written.  For tests and timing. */
int scdrs_gzip_member(const char* text, int64_t n_text, int32_t level, uint8_t* out, int64_t cap, int64_t* out_len);

#ifdef __cplusplus
}
#endif
#endif /* SCDRS_B200_H */
