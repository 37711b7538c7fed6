// C ABI (include/scdrs_b200.h): context management, host<->device staging and phase sequencing.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

namespace scdrs {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (bytes <= cap) return 0;
    if (p != nullptr) {
        cudaError_t e = cudaFree(p);  // synchronises; contents are scratch
        p = nullptr;
        cap = 0;
        if (e != cudaSuccess) {
            set_error("cudaFree failed: %s", cudaGetErrorString(e));
            return 1;
        }
    }
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        return 1;
    }
    cap = want;
    return 0;
}

void DevBuf::release() {
    if (p != nullptr) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static int h2d(scdrs_ctx* c, DevBuf& buf, const void* host, size_t bytes) {
    SCDRS_TRY(buf.reserve(bytes));
    if (bytes) SCDRS_CUDA(cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

// Large uploads from ordinary (pageable) host memory: the driver stages such copies through one pinned
// buffer at ~11 GB/s.  Here a few threads each own a pinned staging buffer and a stream: memcpy a chunk in,
// DMA it out, next chunk -- the CPU copies of some threads overlap the DMA of the others (~2x).  The pinned
// buffers are process-wide and allocated once (pinning memory is slow).  Synchronous: returns when the
// data is on the device.
namespace {
constexpr size_t kUpChunk = (size_t)8 << 20;
constexpr int kUpThreads = 8;
struct Uploader {
    void* pin[kUpThreads] = {};
    cudaStream_t st[kUpThreads] = {};
    int device = -1;
    std::mutex mu;
};
Uploader g_up;
}  // namespace

static int h2d_big(scdrs_ctx* c, DevBuf& buf, const void* host, size_t bytes) {
    SCDRS_TRY(buf.reserve(bytes));
    // a source that is already page-locked (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch tensor behind
    // a numpy view) is copied by the DMA engine directly: one asynchronous copy at link speed, no staging
    cudaPointerAttributes attr;
    const bool pinned = bytes > 0 && cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();        // older drivers report unregistered memory as an error: clear it
    if (pinned || bytes < 4 * kUpChunk) {
        if (bytes) SCDRS_CUDA(cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
        if (pinned && bytes >= 4 * kUpChunk) SCDRS_CUDA(cudaStreamSynchronize(c->stream));   // like the staged path
        return 0;
    }
    std::lock_guard<std::mutex> lock(g_up.mu);
    if (g_up.device != c->device) {
        for (int i = 0; i < kUpThreads; ++i) {
            if (g_up.pin[i]) cudaFreeHost(g_up.pin[i]);
            if (g_up.st[i]) cudaStreamDestroy(g_up.st[i]);
            g_up.pin[i] = nullptr;
            g_up.st[i] = nullptr;
        }
        for (int i = 0; i < kUpThreads; ++i) {
            SCDRS_CUDA(cudaHostAlloc(&g_up.pin[i], kUpChunk, cudaHostAllocDefault));
            SCDRS_CUDA(cudaStreamCreateWithFlags(&g_up.st[i], cudaStreamNonBlocking));
        }
        g_up.device = c->device;
    }
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));      // the destination may still be in use
    const size_t n_chunk = (bytes + kUpChunk - 1) / kUpChunk;
    std::atomic<int> failed(0);
    std::vector<std::thread> workers;
    const int dev = c->device;
    for (int i = 0; i < kUpThreads; ++i) {
        workers.emplace_back([&, i]() {
            if (cudaSetDevice(dev) != cudaSuccess) { failed = 1; return; }
            for (size_t ch = (size_t)i; ch < n_chunk && !failed; ch += kUpThreads) {
                const size_t off = ch * kUpChunk, len = bytes - off < kUpChunk ? bytes - off : kUpChunk;
                memcpy(g_up.pin[i], static_cast<const char*>(host) + off, len);
                if (cudaMemcpyAsync(static_cast<char*>(buf.p) + off, g_up.pin[i], len, cudaMemcpyHostToDevice,
                                    g_up.st[i]) != cudaSuccess ||
                    cudaStreamSynchronize(g_up.st[i]) != cudaSuccess)
                    failed = 1;
            }
        });
    }
    for (auto& w : workers) w.join();
    SCDRS_CHECK(!failed, "chunked upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

static int d2h(scdrs_ctx* c, void* host, const void* dev, size_t bytes) {
    if (bytes) SCDRS_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

static void release_all(scdrs_ctx* c) {
    DevBuf* bufs[] = {&c->own_indptr, &c->own_indices, &c->own_data, &c->var, &c->var_tech,
                      &c->cov_mat, &c->cov_beta, &c->gene_mean, &c->set_genes, &c->set_gw, &c->set_w,
                      &c->covM, &c->covm, &c->ratio_a, &c->w_ptr, &c->w_cursor, &c->w_col, &c->w_val,
                      &c->p_ptr, &c->p_meta, &c->p_col, &c->p_val, &c->colscale, &c->gbase,
                      &c->dev_mat, &c->row_ref, &c->partial, &c->colmean, &c->col2, &c->colmin,
                      &c->row_mean, &c->row_istd, &c->scalars, &c->query, &c->mc_count, &c->norm_mat,
                      &c->sort_keys[0], &c->sort_keys[1], &c->sort_idx[0], &c->sort_idx[1],
                      &c->sort_tab, &c->grid, &c->grid_cells, &c->hist, &c->ge_sorted, &c->ge_local, &c->scan_tmp,
                      &c->x_colsum, &c->x_col2, &c->x_colmin, &c->x_hist, &c->stage,
                      &c->w_code, &c->w_rounds, &c->fx_part, &c->tile_ptr,
                      &c->gene_xsum, &c->cov_colsum, &c->colsum_alg, &c->fx_scal, &c->fx_rowctr, &c->ratio_a_alt, &c->colsum_alg_alt,
                      &c->rnum, &c->stage_prep, &c->scan_tmp_prep,
                      &c->fx_xq, &c->fx_meta32, &c->text_slots, &c->ds_S, &c->ds_tmp, &c->ds_pos, &c->ds_M, &c->ds_keys,
                      &c->ds_vals, &c->ds_off, &c->ds_ref, &c->ds_chunk, &c->ds_part};
    for (DevBuf* b : bufs) b->release();
}

}  // namespace scdrs

using namespace scdrs;

// SCDRS_B200_TRACE=1: host time of the stages of a call, to stderr (where does a submit spend its time?)
#include <chrono>
namespace {
struct StageTrace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::string line;
    explicit StageTrace(const char* what) : on(false) {
        const char* e = getenv("SCDRS_B200_TRACE");
        on = e != nullptr && e[0] == '1';
        if (on) { t0 = std::chrono::steady_clock::now(); line = what; }
    }
    void mark(const char* stage) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof(buf), " | %s %.2f ms", stage, std::chrono::duration<double, std::milli>(t1 - t0).count());
        line += buf;
        t0 = t1;
    }
    ~StageTrace() { if (on) fprintf(stderr, "%s\n", line.c_str()); }
};
}  // namespace

// every API call but the scoring family may change what a trait preparation reads (matrix, gene statistics,
// covariates, lists): the next preparation then orders itself behind the whole main stream
#define CTX_GUARD(c)                                                       \
    SCDRS_CHECK((c) != nullptr, "null context");                           \
    SCDRS_CUDA(cudaSetDevice((c)->device));                                \
    (c)->inputs_dirty = true
#define CTX_GUARD_SCORING(c)                                               \
    SCDRS_CHECK((c) != nullptr, "null context");                           \
    SCDRS_CUDA(cudaSetDevice((c)->device))

template <typename F>
static int empi_null_count(scdrs_ctx* c, int64_t n_t, const F* t, int64_t n_null, const F* null_vals,
                           int64_t* ge_count) {
    SCDRS_CHECK(n_t >= 1 && n_null >= 0, "empty input");
    SCDRS_CHECK(t && ge_count && (null_vals || n_null == 0), "null pointer");
    scdrs_ctx s;
    s.device = c->device;
    s.stream = c->stream;
    DevBuf d_t, d_n, d_ge;
    int rc = 1;
    do {
        if (h2d(&s, d_t, t, (size_t)n_t * sizeof(F)) || h2d(&s, d_n, null_vals, (size_t)n_null * sizeof(F)) ||
            d_ge.reserve((size_t)n_t * sizeof(long long))) break;
        int r2;
        if (sizeof(F) == 4)
            r2 = null_count_generic_f32(&s, n_t, (const float*)d_t.p, n_null, (const float*)d_n.p, d_ge.as<long long>());
        else
            r2 = null_count_generic_f64(&s, n_t, (const double*)d_t.p, n_null, (const double*)d_n.p, d_ge.as<long long>());
        if (r2) break;
        if (d2h(&s, ge_count, d_ge.p, (size_t)n_t * sizeof(int64_t))) break;
        if (cudaStreamSynchronize(s.stream) != cudaSuccess) { set_error("stream sync failed"); break; }
        rc = 0;
    } while (0);
    d_t.release(); d_n.release(); d_ge.release();
    c->launches += s.launches;
    release_all(&s);
    return rc;
}

extern "C" {

const char* scdrs_last_error(void) { return g_err; }
int scdrs_abi_version(void) { return 1; }

int scdrs_ctx_create(int device, scdrs_ctx** out) {
    SCDRS_CHECK(out != nullptr, "null output pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    SCDRS_CHECK(e == cudaSuccess && n > 0,
                "no CUDA device available (%s); scdrs_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    SCDRS_CHECK(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    SCDRS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SCDRS_CUDA(cudaGetDeviceProperties(&prop, device));
    SCDRS_CHECK(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                device, prop.major, prop.minor);
    scdrs_ctx* c = new (std::nothrow) scdrs_ctx();
    SCDRS_CHECK(c != nullptr, "out of host memory");
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        set_error("cudaStreamCreate failed");
        return 1;
    }
    c->own_stream = true;
    for (int i = 0; i < 4; ++i) cudaEventCreate(&c->ev[i]);
    c->timed_spmm = c->timed_trait = true;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        bool ok = cudaStreamCreateWithPriority(&c->prep_stream, cudaStreamNonBlocking, hi) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&c->ev_scatter, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&c->ev_prep, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&c->ev_mark, cudaEventDisableTiming) == cudaSuccess;
        const char* e = getenv("SCDRS_B200_PREP_OVERLAP");
        c->prep_overlap = ok && !(e != nullptr && e[0] == '0');
        e = getenv("SCDRS_B200_TIMELINE");
        if (e != nullptr && e[0] == '1') {
            c->tl.on = true;
            for (int i = 0; i < 16; ++i) {
                cudaEventCreate(&c->tl.prep0[i]); cudaEventCreate(&c->tl.prep1[i]);
                cudaEventCreate(&c->tl.sc0[i]); cudaEventCreate(&c->tl.sc1[i]);
            }
        }
    }
    *out = c;
    return 0;
}

int scdrs_ctx_destroy(scdrs_ctx* c) {
    if (c == nullptr) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    release_all(c);
    for (auto& sl : c->slot) {
        if (sl.pin_in) cudaFreeHost(sl.pin_in);
        if (sl.pin_out) cudaFreeHost(sl.pin_out);
        if (sl.pin_text) cudaFreeHost(sl.pin_text);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    for (int i = 0; i < 4; ++i)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->prep_stream) { cudaStreamSynchronize(c->prep_stream); cudaStreamDestroy(c->prep_stream); }
    for (cudaEvent_t e : {c->ev_scatter, c->ev_prep, c->ev_mark})
        if (e) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int scdrs_ctx_set_stream(scdrs_ctx* c, void* cuda_stream) {
    CTX_GUARD(c);
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    c->own_stream = false;
    return 0;
}

int scdrs_ctx_sync(scdrs_ctx* c) {
    CTX_GUARD(c);
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    if (c->tl.on && c->tl.n >= 16 && c->prep_overlap) {
        SCDRS_CUDA(cudaStreamSynchronize(c->prep_stream));
        const int first = c->tl.n - 15;          // the oldest complete entry of the ring
        auto at = [&](cudaEvent_t e) { float ms = 0.f; cudaEventElapsedTime(&ms, c->tl.sc0[first & 15], e); return ms; };
        for (int t = first; t < c->tl.n; ++t)
            fprintf(stderr, "timeline trait %d: prep %.3f .. %.3f  scatter %.3f .. %.3f\n", t, at(c->tl.prep0[t & 15]),
                    at(c->tl.prep1[t & 15]), at(c->tl.sc0[t & 15]), at(c->tl.sc1[t & 15]));
        c->tl.n = 0;
    }
    return 0;
}

int64_t scdrs_ctx_launch_count(scdrs_ctx* c) { return c ? c->launches : -1; }

int scdrs_ctx_set_spmm_mode(scdrs_ctx* c, int32_t mode) {
    CTX_GUARD(c);
    SCDRS_CHECK(mode >= 0 && mode <= 2, "spmm mode %d outside [0, 2]", mode);
    c->spmm_mode = mode;
    c->trait_ready = false;
    return 0;
}

int scdrs_ctx_last_spmm_kind(scdrs_ctx* c, int32_t* fixed_point, int32_t* scale_log2) {
    CTX_GUARD_SCORING(c);
    if (fixed_point) *fixed_point = c->w_fx ? 1 : 0;
    if (scale_log2) *scale_log2 = c->w_fx ? c->fx_shift : 0;
    return 0;
}

int scdrs_ctx_last_timing(scdrs_ctx* c, float* spmm_ms, float* trait_ms) {
    CTX_GUARD_SCORING(c);
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    if (spmm_ms) {
        *spmm_ms = -1.f;
        if (cudaEventElapsedTime(spmm_ms, c->ev[0], c->ev[1]) != cudaSuccess) { *spmm_ms = -1.f; cudaGetLastError(); }
    }
    if (trait_ms) {
        *trait_ms = -1.f;
        if (cudaEventElapsedTime(trait_ms, c->ev[2], c->ev[3]) != cudaSuccess) { *trait_ms = -1.f; cudaGetLastError(); }
    }
    return 0;
}

// ---- resident inputs ---------------------------------------------------------------------------
static int check_shape(int64_t n_cell, int32_t n_gene) {
    SCDRS_CHECK(n_cell >= 1, "n_cell must be >= 1");
    SCDRS_CHECK(n_gene >= 1, "n_gene must be >= 1");
    return 0;
}

int scdrs_set_csr_host(scdrs_ctx* c, int64_t n_cell, int32_t n_gene, const int64_t* indptr,
                       const int32_t* indices, const float* data) {
    CTX_GUARD(c);
    SCDRS_TRY(check_shape(n_cell, n_gene));
    SCDRS_CHECK(indptr && (indices || indptr[n_cell] == 0), "null CSR arrays");
    const int64_t nnz = indptr[n_cell];
    SCDRS_CHECK(indptr[0] == 0 && nnz >= 0, "malformed indptr");
    SCDRS_TRY(h2d(c, c->own_indptr, indptr, (size_t)(n_cell + 1) * sizeof(int64_t)));
    SCDRS_TRY(h2d_big(c, c->own_indices, indices, (size_t)nnz * sizeof(int32_t)));
    SCDRS_TRY(h2d_big(c, c->own_data, data, (size_t)nnz * sizeof(float)));
    c->indptr = c->own_indptr.as<int64_t>();
    c->indices = c->own_indices.as<int32_t>();
    c->data = c->own_data.as<float>();
    c->n_cell = n_cell;
    c->n_gene = n_gene;
    c->nnz = nnz;
    c->trait_ready = false;
    c->have_stats = false;
    c->fx_stat_key = -1;
    c->fx_xq_key = -1;
    c->tile_ptr_valid = false;
    c->xsum_valid = false;
    c->covsum_valid = false;
    c->colsum_ready = false;
    c->k = 0;
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int scdrs_set_csr_dev(scdrs_ctx* c, int64_t n_cell, int32_t n_gene, const int64_t* dev_indptr,
                      const int32_t* dev_indices, const float* dev_data) {
    CTX_GUARD(c);
    SCDRS_TRY(check_shape(n_cell, n_gene));
    SCDRS_CHECK(dev_indptr != nullptr, "null CSR arrays");
    int64_t nnz = 0;
    SCDRS_CUDA(cudaMemcpyAsync(&nnz, dev_indptr + n_cell, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    c->indptr = dev_indptr;
    c->indices = dev_indices;
    c->data = dev_data;
    c->n_cell = n_cell;
    c->n_gene = n_gene;
    c->nnz = nnz;
    c->trait_ready = false;
    c->have_stats = false;
    c->fx_stat_key = -1;
    c->fx_xq_key = -1;
    c->tile_ptr_valid = false;
    c->xsum_valid = false;
    c->covsum_valid = false;
    c->colsum_ready = false;
    c->k = 0;
    return 0;
}

int scdrs_preprocess_counts(scdrs_ctx* c, int32_t flag_filter_data, int32_t min_genes, int32_t min_cells,
                            int32_t flag_raw_count, double target_sum, int32_t* has_nan, int32_t* has_negative,
                            int64_t* n_cell_out, int32_t* n_gene_out, int64_t* nnz_out, uint8_t* keep_cell,
                            uint8_t* keep_gene, int32_t* n_genes_per_cell, int32_t* n_cells_per_gene, float* n_counts) {
    CTX_GUARD(c);
    SCDRS_CHECK(has_nan && has_negative && n_cell_out && n_gene_out && nnz_out, "null output pointer");
    SCDRS_CHECK((!flag_filter_data && !flag_raw_count) ||
                    (keep_cell && keep_gene && n_genes_per_cell && n_cells_per_gene && n_counts),
                "null output arrays");
    SCDRS_CHECK(!flag_raw_count || target_sum > 0, "target_sum must be positive");
    int hn = 0, hneg = 0;
    const int rc = counts_preprocess(c, flag_filter_data ? 1 : 0, min_genes, min_cells, flag_raw_count ? 1 : 0,
                                     target_sum, &hn, &hneg, n_cell_out, n_gene_out, nnz_out, keep_cell, keep_gene,
                                     n_genes_per_cell, n_cells_per_gene, n_counts);
    *has_nan = hn;
    *has_negative = hneg;
    return rc;
}

int scdrs_get_csr(scdrs_ctx* c, int64_t* indptr, int32_t* indices, float* data) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set");
    SCDRS_CHECK(indptr && (c->nnz == 0 || (indices && data)), "null output pointer");
    SCDRS_TRY(d2h(c, indptr, c->indptr, (size_t)(c->n_cell + 1) * sizeof(int64_t)));
    SCDRS_TRY(d2h(c, indices, c->indices, (size_t)c->nnz * sizeof(int32_t)));
    SCDRS_TRY(d2h(c, data, c->data, (size_t)c->nnz * sizeof(float)));
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int scdrs_set_gene_stats(scdrs_ctx* c, const double* var, const double* var_tech) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->indptr != nullptr, "set the expression matrix first");
    SCDRS_CHECK(var && var_tech, "null gene statistics");
    SCDRS_TRY(h2d(c, c->var, var, (size_t)c->n_gene * sizeof(double)));
    SCDRS_TRY(h2d(c, c->var_tech, var_tech, (size_t)c->n_gene * sizeof(double)));
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    c->have_stats = true;
    c->fx_stat_key = -1;
    c->fx_xq_key = -1;
    return 0;
}

int scdrs_set_cov(scdrs_ctx* c, int32_t k, const double* cov_mat, const double* cov_beta,
                  const double* gene_mean) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->indptr != nullptr, "set the expression matrix first");
    SCDRS_CHECK(k >= 0 && k <= 1024, "number of covariates %d out of range", k);
    if (k > 0) {
        SCDRS_CHECK(cov_mat && cov_beta && gene_mean, "null covariate arrays");
        SCDRS_TRY(h2d(c, c->cov_mat, cov_mat, (size_t)c->n_cell * k * sizeof(double)));
        SCDRS_TRY(h2d(c, c->cov_beta, cov_beta, (size_t)c->n_gene * k * sizeof(double)));
        SCDRS_TRY(h2d(c, c->gene_mean, gene_mean, (size_t)c->n_gene * sizeof(double)));
        SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->k = k;
    c->trait_ready = false;
    c->covsum_valid = false;
    c->colsum_ready = false;
    return 0;
}

// ---- phases -------------------------------------------------------------------------------------
int scdrs_trait_raw(scdrs_ctx* c, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                    const int32_t* trait_genes, const double* trait_gw, const int32_t* ctrl_genes,
                    const double* ctrl_gw, int32_t weight_opt, int32_t use_var_ratio,
                    double* dev_colsum_out) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(trait_genes && trait_gw && (n_ctrl == 0 || n_s_ctrl == 0 || (ctrl_genes && ctrl_gw)), "null gene-set arrays");
    SCDRS_CHECK(dev_colsum_out != nullptr, "null output");
    SCDRS_TRY(trait_prepare(c, n_ctrl, n_s, n_s_ctrl, trait_genes, trait_gw, ctrl_genes, ctrl_gw,
                            weight_opt, use_var_ratio));
    SCDRS_TRY(spmm_score(c));
    return bg_colsum(c, dev_colsum_out);
}

int scdrs_trait_rowstats(scdrs_ctx* c, const double* dev_colsum, int64_t n_cell_total,
                         double* dev_colsum2_out, double* dev_colmin_out) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(dev_colsum && dev_colsum2_out && dev_colmin_out, "null pointer");
    return bg_rowstats(c, dev_colsum, n_cell_total, dev_colsum2_out, dev_colmin_out);
}

int scdrs_trait_queries(scdrs_ctx* c, const double* dev_colsum2, const double* dev_colmin,
                        float* dev_query_out) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(dev_colsum2 && dev_colmin, "null pointer");
    SCDRS_TRY(c->query.reserve((size_t)c->n_cell * sizeof(float)));
    SCDRS_TRY(bg_queries(c, dev_colsum2, dev_colmin, c->query.as<float>()));
    if (dev_query_out != nullptr && dev_query_out != c->query.as<float>())
        SCDRS_CUDA(cudaMemcpyAsync(dev_query_out, c->query.p, (size_t)c->n_cell * sizeof(float),
                                   cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

static int trait_count_impl(scdrs_ctx* c, const float* dev_query_all, int64_t n_query_all,
                            unsigned long long* dev_hist_out, bool write_norm) {
    SCDRS_CHECK(n_query_all >= c->n_cell, "fewer queries than local cells");
    SCDRS_TRY(null_prepare_queries(c, dev_query_all, n_query_all));
    return null_count_pass(c, dev_hist_out, n_query_all, write_norm);
}

int scdrs_trait_count(scdrs_ctx* c, const float* dev_query_all, int64_t n_query_all,
                      unsigned long long* dev_hist_out) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(dev_query_all && dev_hist_out, "null pointer");
    // the normalised control matrix is always kept in the phase API (finish may ask for it)
    return trait_count_impl(c, dev_query_all, n_query_all, dev_hist_out, true);
}

int scdrs_trait_finish(scdrs_ctx* c, const unsigned long long* dev_hist_all, int64_t n_query_all,
                       int64_t query_offset, int64_t n_null_total, double* raw_score,
                       float* norm_score, int32_t* mc_count, int64_t* ge_count, float* ctrl_raw,
                       float* ctrl_norm) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(dev_hist_all != nullptr, "null pointer");
    const int64_t n = c->n_cell;
    SCDRS_CHECK(query_offset >= 0 && query_offset + n <= n_query_all, "query_offset out of range");
    SCDRS_TRY(c->ge_local.reserve((size_t)n * sizeof(long long)));
    SCDRS_TRY(null_finish(c, dev_hist_all, n_query_all, query_offset, n_null_total, n,
                          c->ge_local.as<long long>()));
    if (ge_count) SCDRS_TRY(d2h(c, ge_count, c->ge_local.p, (size_t)n * sizeof(int64_t)));
    if (norm_score) SCDRS_TRY(d2h(c, norm_score, c->query.p, (size_t)n * sizeof(float)));
    if (mc_count) SCDRS_TRY(d2h(c, mc_count, c->mc_count.p, (size_t)n * sizeof(int32_t)));
    if (raw_score || ctrl_raw) {
        const size_t need = (size_t)n * sizeof(double) + (ctrl_raw ? (size_t)n * c->n_ctrl * sizeof(float) : 0);
        SCDRS_TRY(c->stage.reserve(need));
        double* d_raw = c->stage.as<double>();
        float* d_ctrl = ctrl_raw ? reinterpret_cast<float*>(d_raw + n) : nullptr;
        SCDRS_TRY(bg_export_raw(c, d_raw, d_ctrl));
        if (raw_score) SCDRS_TRY(d2h(c, raw_score, d_raw, (size_t)n * sizeof(double)));
        if (ctrl_raw) SCDRS_TRY(d2h(c, ctrl_raw, d_ctrl, (size_t)n * c->n_ctrl * sizeof(float)));
    }
    if (ctrl_norm) {
        SCDRS_CHECK(c->norm_mat.p != nullptr, "normalised control scores were not kept");
        SCDRS_TRY(d2h(c, ctrl_norm, c->norm_mat.p, (size_t)n * c->n_ctrl * sizeof(float)));
    }
    if (!c->finish_async) SCDRS_CUDA(cudaStreamSynchronize(c->stream));     // async: the caller waits on an event
    return 0;
}

int scdrs_score_trait(scdrs_ctx* c, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                      const int32_t* trait_genes, const double* trait_gw, const int32_t* ctrl_genes,
                      const double* ctrl_gw, int32_t weight_opt, int32_t use_var_ratio,
                      double* raw_score, float* norm_score, int32_t* mc_count, int64_t* ge_count,
                      float* ctrl_raw, float* ctrl_norm) {
    CTX_GUARD_SCORING(c);
    const int ncol = n_ctrl + 1;
    SCDRS_CHECK(ncol >= 1, "n_ctrl < 0");
    SCDRS_TRY(c->x_colsum.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->x_col2.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->x_colmin.reserve(ncol * sizeof(double)));
    SCDRS_TRY(c->x_hist.reserve((size_t)(c->n_cell + 1) * sizeof(unsigned long long)));
    StageTrace tr("scdrs_score_trait");
    if (c->timed_trait) SCDRS_CUDA(cudaEventRecord(c->ev[2], c->stream));
    SCDRS_TRY(trait_prepare(c, n_ctrl, n_s, n_s_ctrl, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt,
                            use_var_ratio));
    tr.mark("prepare");
    SCDRS_TRY(spmm_score(c));
    tr.mark("scatter");
    SCDRS_TRY(bg_colsum(c, c->x_colsum.as<double>()));
    SCDRS_TRY(scdrs_trait_rowstats(c, c->x_colsum.as<double>(), c->n_cell, c->x_col2.as<double>(),
                                   c->x_colmin.as<double>()));
    tr.mark("rowstats");
    SCDRS_TRY(scdrs_trait_queries(c, c->x_col2.as<double>(), c->x_colmin.as<double>(), nullptr));
    SCDRS_TRY(trait_count_impl(c, c->query.as<float>(), c->n_cell,
                               c->x_hist.as<unsigned long long>(), ctrl_norm != nullptr || c->force_norm_mat));
    tr.mark("queries+count");
    const int rc = scdrs_trait_finish(c, c->x_hist.as<unsigned long long>(), c->n_cell, 0,
                                      (int64_t)c->n_cell * n_ctrl, raw_score, norm_score, mc_count,
                                      ge_count, ctrl_raw, ctrl_norm);
    tr.mark("finish");
    if (c->timed_trait) cudaEventRecord(c->ev[3], c->stream);
    return rc;
}

// ---- pipelined scoring: submit trait i + 1 while trait i finishes ------------------------------------
// The inputs are copied into pinned memory (so their upload does not wait for the stream) and the results
// are downloaded into pinned memory behind an event; nothing here blocks on the GPU.  Two slots.
static int pin_reserve(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    SCDRS_CUDA(cudaHostAlloc(p, bytes + bytes / 8, cudaHostAllocDefault));
    *cap = bytes + bytes / 8;
    return 0;
}

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// pinned result block of a slot: [raw f64 | ge i64 | norm f32 | mc i32 | ctrl_raw f32 | ctrl_norm f32]
struct SlotOut {
    double* raw;
    int64_t* ge;
    float* norm;
    int32_t* mc;
    float *craw, *cnorm;
    size_t bytes;
};
static SlotOut slot_layout(char* out, int64_t n, int32_t n_ctrl, bool want_raw, bool want_norm) {
    const size_t mat = align16((size_t)n * n_ctrl * 4), a8 = align16((size_t)n * 8), a4 = align16((size_t)n * 4);
    SlotOut o;
    o.raw = reinterpret_cast<double*>(out);
    o.ge = reinterpret_cast<int64_t*>(out + a8);
    o.norm = reinterpret_cast<float*>(out + 2 * a8);
    o.mc = reinterpret_cast<int32_t*>(out + 2 * a8 + a4);
    char* tail = out + 2 * a8 + 2 * a4;
    o.craw = want_raw ? reinterpret_cast<float*>(tail) : nullptr;
    o.cnorm = want_norm ? reinterpret_cast<float*>(tail + (want_raw ? mat : 0)) : nullptr;
    o.bytes = 2 * a8 + 2 * a4 + (want_raw ? mat : 0) + (want_norm ? mat : 0);
    return o;
}

int scdrs_score_trait_submit(scdrs_ctx* c, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl,
                             const int32_t* trait_genes, const double* trait_gw, const int32_t* ctrl_genes,
                             const double* ctrl_gw, int32_t weight_opt, int32_t use_var_ratio,
                             int32_t want_ctrl_raw, int32_t want_ctrl_norm, int32_t* slot_out) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(slot_out != nullptr && trait_genes && trait_gw && n_ctrl >= 0 && n_s >= 0 && n_s_ctrl >= 0, "null pointer");
    SCDRS_CHECK(n_ctrl == 0 || n_s_ctrl == 0 || (ctrl_genes && ctrl_gw), "null gene-set arrays");
    const int id = c->next_slot;
    scdrs_ctx::Slot& sl = c->slot[id];
    SCDRS_CHECK(!sl.busy, "both scoring slots are in flight: collect one first");
    const int64_t n = c->n_cell;
    // pinned input: [trait_genes | ctrl_genes | trait_gw | ctrl_gw]
    const size_t o_tg = 0, o_cg = align16((size_t)n_s * 4), o_tw = o_cg + align16((size_t)n_ctrl * n_s_ctrl * 4),
                 o_cw = o_tw + align16((size_t)n_s * 8), in_bytes = o_cw + align16((size_t)n_s_ctrl * 8);
    SCDRS_TRY(pin_reserve(&sl.pin_in, &sl.in_cap, in_bytes));
    char* in = static_cast<char*>(sl.pin_in);
    memcpy(in + o_tg, trait_genes, (size_t)n_s * 4);
    if (n_ctrl > 0 && n_s_ctrl > 0) {
        memcpy(in + o_cg, ctrl_genes, (size_t)n_ctrl * n_s_ctrl * 4);
        memcpy(in + o_cw, ctrl_gw, (size_t)n_s_ctrl * 8);
    }
    memcpy(in + o_tw, trait_gw, (size_t)n_s * 8);
    // text output: the normalised control scores stay on the device as floats and leave it as 16-byte text slots
    const bool as_text = c->text_output && want_ctrl_norm != 0 && n_ctrl > 0;
    const bool norm_f32 = want_ctrl_norm != 0 && !as_text;
    SlotOut o = slot_layout(nullptr, n, n_ctrl, want_ctrl_raw != 0, norm_f32);
    SCDRS_TRY(pin_reserve(&sl.pin_out, &sl.out_cap, o.bytes));
    const size_t text_bytes = as_text ? (size_t)n * n_ctrl * 16 : 0;
    if (as_text) {
        SCDRS_TRY(pin_reserve(&sl.pin_text, &sl.text_cap, text_bytes + 64));
        SCDRS_TRY(c->text_slots.reserve(text_bytes));
    }
    if (!sl.done) SCDRS_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    o = slot_layout(static_cast<char*>(sl.pin_out), n, n_ctrl, want_ctrl_raw != 0, norm_f32);
    c->finish_async = true;
    c->force_norm_mat = as_text;
    int rc = scdrs_score_trait(c, n_ctrl, n_s, n_s_ctrl, reinterpret_cast<const int32_t*>(in + o_tg),
                               reinterpret_cast<const double*>(in + o_tw), reinterpret_cast<const int32_t*>(in + o_cg),
                               reinterpret_cast<const double*>(in + o_cw), weight_opt, use_var_ratio, o.raw, o.norm,
                               o.mc, o.ge, o.craw, o.cnorm);
    c->finish_async = false;
    c->force_norm_mat = false;
    if (rc) return rc;
    if (as_text) {
        SCDRS_TRY(format_text_slots(c, (int64_t)n * n_ctrl, c->norm_mat.as<float>(), c->text_slots.p));
        SCDRS_TRY(d2h(c, sl.pin_text, c->text_slots.p, text_bytes));
    }
    SCDRS_CUDA(cudaEventRecord(sl.done, c->stream));
    sl.busy = true;
    sl.n_ctrl = n_ctrl;
    sl.ctrl_raw = want_ctrl_raw != 0;
    sl.ctrl_norm = norm_f32;
    sl.text = as_text;
    c->next_slot = id ^ 1;
    *slot_out = id;
    return 0;
}

int scdrs_score_trait_collect(scdrs_ctx* c, int32_t slot, double* raw_score, float* norm_score, int32_t* mc_count,
                              int64_t* ge_count, float* ctrl_raw, float* ctrl_norm) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(slot == 0 || slot == 1, "slot must be 0 or 1");
    scdrs_ctx::Slot& sl = c->slot[slot];
    SCDRS_CHECK(sl.busy, "nothing was submitted to this slot");
    SCDRS_CHECK((ctrl_raw == nullptr || sl.ctrl_raw) && (ctrl_norm == nullptr || sl.ctrl_norm),
                "control scores were not requested at submit");
    SCDRS_CUDA(cudaEventSynchronize(sl.done));
    sl.busy = false;
    const int64_t n = c->n_cell;
    const size_t mat = (size_t)n * sl.n_ctrl * 4;
    const SlotOut o = slot_layout(static_cast<char*>(sl.pin_out), n, sl.n_ctrl, sl.ctrl_raw, sl.ctrl_norm);
    if (raw_score) memcpy(raw_score, o.raw, (size_t)n * 8);
    if (ge_count) memcpy(ge_count, o.ge, (size_t)n * 8);
    if (norm_score) memcpy(norm_score, o.norm, (size_t)n * 4);
    if (mc_count) memcpy(mc_count, o.mc, (size_t)n * 4);
    if (ctrl_raw) memcpy(ctrl_raw, o.craw, mat);
    if (ctrl_norm) memcpy(ctrl_norm, o.cnorm, mat);
    return 0;
}

int scdrs_ctx_set_text_output(scdrs_ctx* c, int32_t on) {
    CTX_GUARD(c);
    c->text_output = on != 0;
    return 0;
}

int scdrs_ctx_set_resident_sets(scdrs_ctx* c, int32_t on) {
    CTX_GUARD(c);
    c->resident_sets = on != 0;
    return 0;
}

int scdrs_slot_text(scdrs_ctx* c, int32_t slot, const void** text, int64_t* n_bytes) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(slot == 0 || slot == 1, "slot must be 0 or 1");
    SCDRS_CHECK(text != nullptr && n_bytes != nullptr, "null pointer");
    scdrs_ctx::Slot& sl = c->slot[slot];
    SCDRS_CHECK(!sl.busy, "collect the slot first");
    SCDRS_CHECK(sl.text && sl.pin_text != nullptr, "the trait in this slot was not submitted with text output");
    *text = sl.pin_text;
    *n_bytes = (int64_t)c->n_cell * sl.n_ctrl * 16;
    return 0;
}

// Phase-level halves for the multi-GPU driver: phase 1 with the control sets already on the device (they
// arrive there by broadcast) and the small host arrays staged through the slot's pinned block, and phase 5
// into the slot's pinned result block behind an event (collected with scdrs_score_trait_collect).
int scdrs_trait_raw_dev(scdrs_ctx* c, int32_t n_ctrl, int32_t n_s, int32_t n_s_ctrl, const int32_t* trait_genes,
                        const double* trait_gw, const int32_t* dev_ctrl_genes, const double* ctrl_gw,
                        int32_t weight_opt, int32_t use_var_ratio, double* dev_colsum_out) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(trait_genes && trait_gw && n_ctrl >= 0 && n_s >= 0 && n_s_ctrl >= 0, "null pointer");
    SCDRS_CHECK(n_ctrl == 0 || n_s_ctrl == 0 || (dev_ctrl_genes && ctrl_gw), "null gene-set arrays");
    SCDRS_CHECK(dev_colsum_out != nullptr, "null output");
    scdrs_ctx::Slot& sl = c->slot[c->next_slot];
    SCDRS_CHECK(!sl.busy, "both scoring slots are in flight: collect one first");
    const size_t o_tw = align16((size_t)n_s * 4), o_cw = o_tw + align16((size_t)n_s * 8),
                 in_bytes = o_cw + align16((size_t)n_s_ctrl * 8);
    SCDRS_TRY(pin_reserve(&sl.pin_in, &sl.in_cap, in_bytes));
    char* in = static_cast<char*>(sl.pin_in);
    memcpy(in, trait_genes, (size_t)n_s * 4);
    memcpy(in + o_tw, trait_gw, (size_t)n_s * 8);
    if (n_s_ctrl > 0 && ctrl_gw) memcpy(in + o_cw, ctrl_gw, (size_t)n_s_ctrl * 8);
    SCDRS_TRY(trait_prepare(c, n_ctrl, n_s, n_s_ctrl, reinterpret_cast<const int32_t*>(in),
                            reinterpret_cast<const double*>(in + o_tw), dev_ctrl_genes,
                            reinterpret_cast<const double*>(in + o_cw), weight_opt, use_var_ratio, true));
    SCDRS_TRY(spmm_score(c));
    return bg_colsum(c, dev_colsum_out);
}

int scdrs_trait_finish_submit(scdrs_ctx* c, const unsigned long long* dev_hist_all, int64_t n_query_all,
                              int64_t query_offset, int64_t n_null_total, int32_t want_ctrl_raw,
                              int32_t want_ctrl_norm, int32_t* slot_out) {
    CTX_GUARD_SCORING(c);
    SCDRS_CHECK(slot_out != nullptr, "null pointer");
    const int id = c->next_slot;
    scdrs_ctx::Slot& sl = c->slot[id];
    SCDRS_CHECK(!sl.busy, "both scoring slots are in flight: collect one first");
    const int64_t n = c->n_cell;
    SlotOut o = slot_layout(nullptr, n, c->n_ctrl, want_ctrl_raw != 0, want_ctrl_norm != 0);
    SCDRS_TRY(pin_reserve(&sl.pin_out, &sl.out_cap, o.bytes));
    if (!sl.done) SCDRS_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    o = slot_layout(static_cast<char*>(sl.pin_out), n, c->n_ctrl, want_ctrl_raw != 0, want_ctrl_norm != 0);
    c->finish_async = true;
    const int rc = scdrs_trait_finish(c, dev_hist_all, n_query_all, query_offset, n_null_total, o.raw, o.norm, o.mc,
                                      o.ge, o.craw, o.cnorm);
    c->finish_async = false;
    if (rc) return rc;
    SCDRS_CUDA(cudaEventRecord(sl.done, c->stream));
    sl.busy = true;
    sl.n_ctrl = c->n_ctrl;
    sl.ctrl_raw = want_ctrl_raw != 0;
    sl.ctrl_norm = want_ctrl_norm != 0;
    c->next_slot = id ^ 1;
    *slot_out = id;
    return 0;
}

// ---- stand-alone helpers --------------------------------------------------------------------------
int scdrs_compute_raw_score(scdrs_ctx* c, int32_t n_set, int32_t n_s, const int32_t* genes,
                            const double* w, double* raw_out) {
    CTX_GUARD(c);
    SCDRS_CHECK(genes && w && raw_out, "null pointer");
    SCDRS_TRY(explicit_prepare(c, n_set, n_s, genes, w));
    SCDRS_TRY(spmm_score(c));
    // export as f64 [n_cell x n_set]: column 0 through raw_score, the rest through ctrl_raw (f32)
    // would lose precision, so unpack on the host from the packed representation instead.
    const int64_t n = c->n_cell;
    std::vector<float> dev((size_t)n * c->ld);
    std::vector<double> ref((size_t)n);
    SCDRS_TRY(d2h(c, dev.data(), c->dev_mat.p, dev.size() * sizeof(float)));
    SCDRS_TRY(d2h(c, ref.data(), c->row_ref.p, ref.size() * sizeof(double)));
    SCDRS_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t r = 0; r < n; ++r)
        for (int j = 0; j < n_set; ++j) {
            const float d = dev[(size_t)r * c->ld + j];
            raw_out[(size_t)r * n_set + j] = (d != d) ? 0.0 : ref[r] + (double)d;
        }
    return 0;
}

int scdrs_correct_background(scdrs_ctx* c, int64_t n_cell, int32_t n_ctrl, const double* v_raw,
                             const double* mat_ctrl, const double* var_ratio, double* v_norm,
                             double* mat_norm) {
    CTX_GUARD(c);
    SCDRS_CHECK(n_cell >= 1 && n_ctrl >= 1, "empty input");
    SCDRS_CHECK(v_raw && mat_ctrl && var_ratio && v_norm && mat_norm, "null pointer");
    // scratch context sharing the device and stream, so the caller's resident state is untouched
    scdrs_ctx t;
    t.device = c->device;
    t.stream = c->stream;
    t.n_cell = n_cell;
    int rc = 1;
    do {
        DevBuf d_v, d_m, d_r;
        auto cleanup = [&]() { d_v.release(); d_m.release(); d_r.release(); };
        const int ncol = n_ctrl + 1;
        if (h2d(&t, d_v, v_raw, (size_t)n_cell * sizeof(double)) ||
            h2d(&t, d_m, mat_ctrl, (size_t)n_cell * n_ctrl * sizeof(double)) ||
            h2d(&t, d_r, var_ratio, (size_t)n_ctrl * sizeof(double))) { cleanup(); break; }
        if (bg_pack_dense(&t, n_cell, n_ctrl, d_v.as<double>(), d_m.as<double>(), d_r.as<double>())) { cleanup(); break; }
        t.trait_ready = true;
        if (t.x_colsum.reserve(ncol * 8) || t.x_col2.reserve(ncol * 8) || t.x_colmin.reserve(ncol * 8) ||
            t.x_hist.reserve((size_t)(n_cell + 1) * 8) || t.query.reserve((size_t)n_cell * 4)) { cleanup(); break; }
        if (bg_colsum(&t, t.x_colsum.as<double>()) ||
            bg_rowstats(&t, t.x_colsum.as<double>(), n_cell, t.x_col2.as<double>(), t.x_colmin.as<double>()) ||
            bg_queries(&t, t.x_col2.as<double>(), t.x_colmin.as<double>(), t.query.as<float>()) ||
            null_prepare_queries(&t, t.query.as<float>(), n_cell) ||
            null_count_pass(&t, t.x_hist.as<unsigned long long>(), n_cell, true)) { cleanup(); break; }
        std::vector<float> q((size_t)n_cell), m((size_t)n_cell * n_ctrl);
        if (d2h(&t, q.data(), t.query.p, q.size() * 4) || d2h(&t, m.data(), t.norm_mat.p, m.size() * 4)) { cleanup(); break; }
        if (cudaStreamSynchronize(t.stream) != cudaSuccess) { set_error("stream sync failed"); cleanup(); break; }
        for (size_t i = 0; i < q.size(); ++i) v_norm[i] = (double)q[i];
        for (size_t i = 0; i < m.size(); ++i) mat_norm[i] = (double)m[i];
        cleanup();
        rc = 0;
    } while (0);
    c->launches += t.launches;
    release_all(&t);
    return rc;
}

int scdrs_empi_null_count_f32(scdrs_ctx* c, int64_t n_t, const float* t, int64_t n_null,
                              const float* null_vals, int64_t* ge_count) {
    CTX_GUARD(c);
    return empi_null_count<float>(c, n_t, t, n_null, null_vals, ge_count);
}

int scdrs_empi_null_count_f64(scdrs_ctx* c, int64_t n_t, const double* t, int64_t n_null,
                              const double* null_vals, int64_t* ge_count) {
    CTX_GUARD(c);
    return empi_null_count<double>(c, n_t, t, n_null, null_vals, ge_count);
}

int scdrs_gene_cell_stats(scdrs_ctx* c, int32_t transform, const double* cell_weight, int32_t center,
                          double denom, double* gene_sum, double* gene_sumsq, double* cell_sum,
                          double* cell_sumsq) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set");
    const int64_t n = c->n_cell;
    const int g = c->n_gene;
    DevBuf d_w;
    if (cell_weight) SCDRS_TRY(h2d(c, d_w, cell_weight, (size_t)n * sizeof(double)));
    const bool wg = gene_sum != nullptr, wc = cell_sum != nullptr;
    SCDRS_TRY(c->stage.reserve(((size_t)2 * g + (size_t)2 * n) * sizeof(double)));
    double* d_gs = c->stage.as<double>();
    double* d_gq = d_gs + g;
    double* d_cs = d_gq + g;
    double* d_cq = d_cs + n;
    int rc = gene_cell_stats(c, transform, cell_weight ? d_w.as<double>() : nullptr, center, denom, wg ? d_gs : nullptr,
                             wg ? d_gq : nullptr, wc ? d_cs : nullptr, wc ? d_cq : nullptr);
    if (rc == 0 && wg) rc = d2h(c, gene_sum, d_gs, (size_t)g * 8) || d2h(c, gene_sumsq, d_gq, (size_t)g * 8);
    if (rc == 0 && wc) rc = d2h(c, cell_sum, d_cs, (size_t)n * 8) || d2h(c, cell_sumsq, d_cq, (size_t)n * 8);
    if (rc == 0 && cudaStreamSynchronize(c->stream) != cudaSuccess) { set_error("stream sync failed"); rc = 1; }
    d_w.release();
    return rc;
}

int scdrs_gene_mean_xtc(scdrs_ctx* c, int32_t k, const double* cov_mat, float* gene_mean, double* xtc) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set");
    SCDRS_CHECK(gene_mean != nullptr && (k == 0 || (cov_mat != nullptr && xtc != nullptr)), "null pointer");
    const int64_t n = c->n_cell;
    const int g = c->n_gene;
    DevBuf d_cov;
    SCDRS_TRY(h2d(c, d_cov, cov_mat, (size_t)n * k * sizeof(double)));
    SCDRS_TRY(c->stage.reserve((size_t)g * sizeof(float) + (size_t)g * (k > 0 ? k : 1) * sizeof(double) + 64));
    double* d_xtc = c->stage.as<double>();
    float* d_gm = reinterpret_cast<float*>(d_xtc + (size_t)g * (k > 0 ? k : 1));
    int rc = gene_mean_xtc(c, k, d_cov.as<double>(), d_gm, d_xtc);
    if (rc == 0) rc = d2h(c, gene_mean, d_gm, (size_t)g * sizeof(float));
    if (rc == 0 && k > 0) rc = d2h(c, xtc, d_xtc, (size_t)g * k * sizeof(double));
    if (rc == 0 && cudaStreamSynchronize(c->stream) != cudaSuccess) { set_error("stream sync failed"); rc = 1; }
    d_cov.release();
    return rc;
}

int scdrs_gene_moments(scdrs_ctx* c, int32_t mode, int32_t k, const double* cov_mat, const double* cell_weight,
                       double* out) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set");
    SCDRS_CHECK(out != nullptr && (mode != 0 || k == 0 || cov_mat != nullptr), "null pointer");
    const int64_t n = c->n_cell;
    const int g = c->n_gene;
    const int n_acc = mode == 0 ? 2 + k : 2;
    DevBuf d_cov, d_w;
    int rc = 0;
    if (mode == 0 && k > 0) rc = h2d(c, d_cov, cov_mat, (size_t)n * k * sizeof(double));
    if (rc == 0 && cell_weight) rc = h2d(c, d_w, cell_weight, (size_t)n * sizeof(double));
    if (rc == 0) rc = c->stage.reserve((size_t)g * n_acc * sizeof(double));
    if (rc == 0) rc = gene_moments(c, mode, k, d_cov.as<double>(), cell_weight ? d_w.as<double>() : nullptr, c->stage.as<double>());
    if (rc == 0) rc = d2h(c, out, c->stage.p, (size_t)g * n_acc * sizeof(double));
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == 0) { set_error("stream sync failed"); rc = 1; }
    d_cov.release();
    d_w.release();
    return rc;
}

int scdrs_gene_expm1_implicit(scdrs_ctx* c, const double* shift, const double* cell_weight, double* out) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set");
    SCDRS_CHECK(shift != nullptr && out != nullptr, "null pointer");
    const int64_t n = c->n_cell;
    const int g = c->n_gene;
    DevBuf d_shift, d_w;
    int rc = h2d(c, d_shift, shift, (size_t)g * sizeof(double));
    if (rc == 0 && cell_weight) rc = h2d(c, d_w, cell_weight, (size_t)n * sizeof(double));
    if (rc == 0) rc = c->stage.reserve((size_t)g * 2 * sizeof(double));
    if (rc == 0) rc = gene_expm1_implicit(c, d_shift.as<double>(), cell_weight ? d_w.as<double>() : nullptr, c->stage.as<double>());
    if (rc == 0) rc = d2h(c, out, c->stage.p, (size_t)g * 2 * sizeof(double));
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == 0) { set_error("stream sync failed"); rc = 1; }
    d_shift.release();
    d_w.release();
    return rc;
}

int scdrs_gene_tile_size(void) { return gene_tile_size(); }

int scdrs_gene_stats_exact(scdrs_ctx* c, int32_t transform, const double* cell_weight, int32_t center, double denom,
                           int32_t n_tile, const int32_t* tile_ids, const double* mean_in, double* gene_sum,
                           double* gene_sumsq) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->indptr != nullptr, "no expression matrix set");
    SCDRS_CHECK(n_tile >= 1 && tile_ids && gene_sum && gene_sumsq, "null pointer");
    const int64_t n = c->n_cell;
    const int g = c->n_gene, ts = gene_tile_size();
    for (int i = 0; i < n_tile; ++i)
        SCDRS_CHECK(tile_ids[i] >= 0 && (int64_t)tile_ids[i] * ts < g, "gene tile %d out of range", tile_ids[i]);
    DevBuf d_w, d_tiles, d_mean;
    int rc = h2d(c, d_tiles, tile_ids, (size_t)n_tile * sizeof(int32_t));
    if (rc == 0 && cell_weight) rc = h2d(c, d_w, cell_weight, (size_t)n * sizeof(double));
    if (rc == 0 && mean_in) rc = h2d(c, d_mean, mean_in, (size_t)g * sizeof(double));
    if (rc == 0) rc = c->stage.reserve((size_t)2 * g * sizeof(double));
    double* d_gs = c->stage.as<double>();
    double* d_gq = d_gs + g;
    if (rc == 0)
        rc = gene_stats_exact(c, transform, cell_weight ? d_w.as<double>() : nullptr, center, denom, n_tile,
                              d_tiles.as<int32_t>(), mean_in ? d_mean.as<double>() : nullptr, d_gs, d_gq);
    std::vector<double> hs((size_t)g), hq((size_t)g);
    if (rc == 0) rc = d2h(c, hs.data(), d_gs, (size_t)g * 8) || d2h(c, hq.data(), d_gq, (size_t)g * 8);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == 0) { set_error("stream sync failed"); rc = 1; }
    if (rc == 0)
        for (int i = 0; i < n_tile; ++i)
            for (int j = tile_ids[i] * ts; j < (tile_ids[i] + 1) * ts && j < g; ++j) {
                gene_sum[j] = hs[j];
                gene_sumsq[j] = hq[j];
            }
    d_w.release();
    d_tiles.release();
    d_mean.release();
    return rc;
}

// ---- downstream analyses of one trait's score matrix (downstream.cu) ------------------------------
int scdrs_ds_set_scores(scdrs_ctx* c, int64_t n_cell, int32_t ncol, const void* frame, int32_t is_f32,
                        int64_t n_elem, int64_t row_stride, int64_t col_stride, const int32_t* cols) {
    CTX_GUARD(c);
    SCDRS_CHECK(frame != nullptr && cols != nullptr && n_cell > 0 && ncol > 0, "empty score matrix");
    SCDRS_CHECK(row_stride == 1 || col_stride == 1,
                "the frame must be row-major (col_stride 1) or column-major (row_stride 1)");
    for (int j = 0; j < ncol; ++j) {
        const int64_t last = (n_cell - 1) * row_stride + (int64_t)cols[j] * col_stride;
        SCDRS_CHECK(cols[j] >= 0 && last < n_elem, "score column %d outside the frame", j);
    }
    c->ds_n = 0;
    DevBuf d_cols;
    SCDRS_TRY(c->ds_S.reserve((size_t)n_cell * ncol * sizeof(double)));
    SCDRS_TRY(h2d_big(c, c->ds_M, frame, (size_t)n_elem * (is_f32 ? sizeof(float) : sizeof(double))));
    int rc = h2d(c, d_cols, cols, (size_t)ncol * sizeof(int32_t));
    if (rc == 0) rc = ds_import(c, n_cell, ncol, is_f32, row_stride, col_stride, d_cols.as<int32_t>(), c->ds_M.p, c->ds_S.as<double>());
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == 0) { set_error("stream sync failed"); rc = 1; }
    d_cols.release();
    if (rc) return rc;
    c->ds_n = n_cell;
    c->ds_ncol = ncol;
    c->ds_f32 = is_f32 != 0;
    return 0;
}

int scdrs_ds_corr(scdrs_ctx* c, int32_t n_var, const double* vars, double* corr) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->ds_n > 0, "no score matrix set (scdrs_ds_set_scores)");
    SCDRS_CHECK(n_var > 0 && vars != nullptr && corr != nullptr, "null pointer");
    DevBuf d_vars, d_corr;
    int rc = h2d(c, d_vars, vars, (size_t)n_var * c->ds_n * sizeof(double)) ||
             d_corr.reserve((size_t)n_var * c->ds_ncol * sizeof(double));
    if (rc == 0) rc = ds_corr(c, n_var, d_vars.as<double>(), d_corr.as<double>());
    if (rc == 0) rc = d2h(c, corr, d_corr.p, (size_t)n_var * c->ds_ncol * sizeof(double));
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == 0) { set_error("stream sync failed"); rc = 1; }
    d_vars.release();
    d_corr.release();
    return rc;
}

int scdrs_ds_group_stats(scdrs_ctx* c, int32_t n_group, const int64_t* order, const int64_t* grp_ptr,
                         const int32_t* q_lo, const double* q_gamma, double* quantile, int32_t geary_mode,
                         const int64_t* g_indptr, const int32_t* g_nbr, const double* g_w, double* geary_total,
                         double* geary_denom) {
    CTX_GUARD(c);
    SCDRS_CHECK(c->ds_n > 0, "no score matrix set (scdrs_ds_set_scores)");
    SCDRS_CHECK(n_group > 0 && order != nullptr && grp_ptr != nullptr, "null pointer");
    SCDRS_CHECK(n_group <= 65535, "at most 65535 cell groups per call (got %d)", n_group);
    SCDRS_CHECK(geary_mode >= 0 && geary_mode <= 2, "geary_mode must be 0, 1 or 2");
    SCDRS_CHECK(quantile == nullptr || (q_lo != nullptr && q_gamma != nullptr), "quantile positions missing");
    SCDRS_CHECK(geary_mode == 0 || (g_indptr && geary_total && geary_denom), "graph or outputs missing");
    const int64_t n = c->ds_n;
    SCDRS_CHECK(grp_ptr[0] == 0 && grp_ptr[n_group] == n, "grp_ptr must cover all cells");
    for (int g = 0; g < n_group; ++g) SCDRS_CHECK(grp_ptr[g] <= grp_ptr[g + 1], "grp_ptr must not decrease");
    const size_t gm = (size_t)n_group * c->ds_ncol * sizeof(double);
    DevBuf d_order, d_gptr, d_qlo, d_qg, d_gi, d_gn, d_gw, d_out;
    int rc = h2d(c, d_order, order, (size_t)n * 8) || h2d(c, d_gptr, grp_ptr, (size_t)(n_group + 1) * 8) ||
             d_out.reserve(3 * gm);
    if (rc == 0 && quantile) rc = h2d(c, d_qlo, q_lo, (size_t)n_group * 4) || h2d(c, d_qg, q_gamma, (size_t)n_group * 8);
    if (rc == 0 && geary_mode) {
        const int64_t ne = g_indptr[n];
        SCDRS_CHECK(ne == 0 || (g_nbr != nullptr && g_w != nullptr), "graph arrays missing");
        rc = h2d(c, d_gi, g_indptr, (size_t)(n + 1) * 8) || d_gn.reserve(8) || d_gw.reserve(8) ||
             h2d(c, d_gn, g_nbr, (size_t)ne * 4) || h2d(c, d_gw, g_w, (size_t)ne * 8);
    }
    double* d_quant = d_out.as<double>();
    double* d_total = d_quant + (size_t)n_group * c->ds_ncol;
    double* d_denom = d_total + (size_t)n_group * c->ds_ncol;
    if (rc == 0)
        rc = ds_group_stats(c, n_group, grp_ptr, d_order.as<int64_t>(), d_gptr.as<int64_t>(), d_qlo.as<int32_t>(),
                            d_qg.as<double>(), quantile ? d_quant : nullptr, geary_mode, d_gi.as<int64_t>(),
                            d_gn.as<int32_t>(), d_gw.as<double>(), d_total, d_denom);
    if (rc == 0 && quantile) rc = d2h(c, quantile, d_quant, gm);
    if (rc == 0 && geary_mode) rc = d2h(c, geary_total, d_total, gm) || d2h(c, geary_denom, d_denom, gm);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == 0) { set_error("stream sync failed"); rc = 1; }
    DevBuf* tmp[] = {&d_order, &d_gptr, &d_qlo, &d_qg, &d_gi, &d_gn, &d_gw, &d_out};
    for (DevBuf* b : tmp) b->release();
    return rc;
}

}  // extern "C"
