"""ctypes binding of ``libscdrs_b200.so`` (C ABI in ``include/scdrs_b200.h``).

The product path has NO CPU fallback: if the library is missing or no B200 is visible, the
first call raises ``RuntimeError``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscdrs_b200.so")

_lib = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_u64p = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); must list every symbol include/scdrs_b200.h declares
SIGNATURES = {
    "scdrs_last_error": (C.c_char_p, []),
    "scdrs_abi_version": (C.c_int, []),
    "scdrs_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "scdrs_ctx_destroy": (C.c_int, [C.c_void_p]),
    "scdrs_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "scdrs_ctx_sync": (C.c_int, [C.c_void_p]),
    "scdrs_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "scdrs_ctx_last_timing": (C.c_int, [C.c_void_p, c_f32p, c_f32p]),
    "scdrs_ctx_set_spmm_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "scdrs_ctx_last_spmm_kind": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "scdrs_set_csr_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_set_csr_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_preprocess_counts": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_get_csr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_set_gene_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_set_cov": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_trait_raw": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "scdrs_trait_rowstats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "scdrs_trait_queries": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_trait_count": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "scdrs_trait_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_score_trait": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_score_trait_submit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.POINTER(C.c_int32)]),
    "scdrs_score_trait_collect": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "scdrs_trait_raw_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "scdrs_trait_finish_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                            C.c_int32, C.POINTER(C.c_int32)]),
    "scdrs_compute_raw_score": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_correct_background": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_empi_null_count_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "scdrs_empi_null_count_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "scdrs_gene_cell_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_gene_mean_xtc": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_gene_moments": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_gene_expm1_implicit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_gene_stats_exact": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scdrs_gene_tile_size": (C.c_int, []),
    "scdrs_ds_set_scores": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int64,
                                      C.c_int64, C.c_void_p]),
    "scdrs_ds_corr": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "scdrs_ds_group_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "scdrs_select_ctrl": (C.c_int, [C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    "scdrs_select_ctrl_state": (C.c_int, [C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_void_p]),
    "scdrs_write_score_gz": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_int32]),
    "scdrs_write_score_gz_slots": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                             C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]),
    "scdrs_ctx_set_text_output": (C.c_int, [C.c_void_p, C.c_int32]),
    "scdrs_ctx_set_resident_sets": (C.c_int, [C.c_void_p, C.c_int32]),
    "scdrs_slot_text": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "scdrs_format_f32": (C.c_int, [C.c_float, C.c_char_p]),
    "scdrs_format_f32_shortest": (C.c_int, [C.c_float, C.c_char_p]),
    "scdrs_format_slots_host": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]),
    "scdrs_selftest_format": (C.c_int64, [C.c_uint32, C.c_uint64, C.c_int32, C.POINTER(C.c_uint32)]),
    "scdrs_gzip_member": (C.c_int, [C.c_char_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
}

WEIGHT_OPT = {"uniform": 0, "vs": 1, "inv_std": 2}


def lib():
    """Load the shared library (once); raise loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "scdrs_b200: %s not found -- build it with `python -m scdrs_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("scdrs_b200: " + lib().scdrs_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Raw pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One GPU context holding one resident expression matrix."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._L = lib()
        check(self._L.scdrs_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self.n_cell = 0
        self.n_gene = 0
        self._keep = []  # arrays / tensors the device side borrows

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            self._L.scdrs_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- resident inputs --
    def set_csr_host(self, indptr, indices, data, n_gene):
        indptr = as_c(indptr, np.int64)
        indices = as_c(indices, np.int32)
        data = as_c(data, np.float32)
        n_cell = indptr.shape[0] - 1
        check(self._L.scdrs_set_csr_host(self._h, n_cell, int(n_gene), ptr(indptr), ptr(indices), ptr(data)))
        self.n_cell, self.n_gene = n_cell, int(n_gene)
        self.h2d_bytes = indptr.nbytes + indices.nbytes + data.nbytes

    def set_csr_dev(self, n_cell, n_gene, d_indptr, d_indices, d_data, keep=()):
        check(self._L.scdrs_set_csr_dev(self._h, int(n_cell), int(n_gene), int(d_indptr), int(d_indices), int(d_data)))
        self.n_cell, self.n_gene = int(n_cell), int(n_gene)
        self._keep = list(keep)

    def preprocess_counts(self, flag_filter_data, min_genes, min_cells, flag_raw_count, target_sum=1e4):
        """load_h5ad's checks / filters / normalisation on the resident matrix (scdrs_preprocess_counts).
        Returns a dict; the processed matrix stays resident (``get_csr`` downloads it)."""
        n_cell, n_gene = self.n_cell, self.n_gene
        out = dict(keep_cell=np.zeros(n_cell, np.uint8), keep_gene=np.zeros(n_gene, np.uint8),
                   n_genes=np.zeros(n_cell, np.int32), n_cells=np.zeros(n_gene, np.int32),
                   n_counts=np.zeros(n_cell, np.float32))
        has_nan, has_neg = C.c_int32(), C.c_int32()
        nc, ng, nz = C.c_int64(), C.c_int32(), C.c_int64()
        check(self._L.scdrs_preprocess_counts(
            self._h, int(bool(flag_filter_data)), int(min_genes), int(min_cells), int(bool(flag_raw_count)),
            float(target_sum), C.byref(has_nan), C.byref(has_neg), C.byref(nc), C.byref(ng), C.byref(nz),
            ptr(out["keep_cell"]), ptr(out["keep_gene"]), ptr(out["n_genes"]), ptr(out["n_cells"]), ptr(out["n_counts"])))
        out.update(has_nan=bool(has_nan.value), has_negative=bool(has_neg.value), n_cell=int(nc.value),
                   n_gene=int(ng.value), nnz=int(nz.value))
        if not (out["has_nan"] or out["has_negative"]):
            self.n_cell, self.n_gene = out["n_cell"], out["n_gene"]
        return out

    def get_csr(self, nnz):
        indptr = np.empty(self.n_cell + 1, np.int64)
        indices, data = np.empty(nnz, np.int32), np.empty(nnz, np.float32)
        check(self._L.scdrs_get_csr(self._h, ptr(indptr), ptr(indices), ptr(data)))
        return indptr, indices, data

    def set_gene_stats(self, var, var_tech):
        var, var_tech = as_c(var, np.float64), as_c(var_tech, np.float64)
        assert var.shape[0] == self.n_gene and var_tech.shape[0] == self.n_gene
        check(self._L.scdrs_set_gene_stats(self._h, ptr(var), ptr(var_tech)))

    def set_cov(self, cov_mat=None, cov_beta=None, gene_mean=None):
        if cov_mat is None:
            check(self._L.scdrs_set_cov(self._h, 0, None, None, None))
            return
        cov_mat, cov_beta = as_c(cov_mat, np.float64), as_c(cov_beta, np.float64)
        gene_mean = as_c(gene_mean, np.float64)
        k = cov_mat.shape[1]
        assert cov_mat.shape == (self.n_cell, k) and cov_beta.shape == (self.n_gene, k)
        check(self._L.scdrs_set_cov(self._h, k, ptr(cov_mat), ptr(cov_beta), ptr(gene_mean)))

    def set_stream(self, cuda_stream):
        check(self._L.scdrs_ctx_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def sync(self):
        check(self._L.scdrs_ctx_sync(self._h))

    def launch_count(self):
        return int(self._L.scdrs_ctx_launch_count(self._h))

    def set_spmm_mode(self, mode):
        """0 auto, 1 fp64 accumulators, 2 fixed point (raises when it does not apply)."""
        check(self._L.scdrs_ctx_set_spmm_mode(self._h, {"auto": 0, "f64": 1, "fixed": 2}.get(mode, mode)))

    def last_spmm_kind(self):
        a, b = C.c_int32(), C.c_int32()
        check(self._L.scdrs_ctx_last_spmm_kind(self._h, C.byref(a), C.byref(b)))
        return ("fixed" if a.value else "f64"), b.value

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        check(self._L.scdrs_ctx_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- one trait, single GPU --
    def score_trait(self, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_var_ratio,
                    want_ctrl_raw=False, want_ctrl_norm=False):
        trait_genes = as_c(trait_genes, np.int32)
        ctrl_genes = as_c(ctrl_genes, np.int32)
        trait_gw, ctrl_gw = as_c(trait_gw, np.float64), as_c(ctrl_gw, np.float64)
        n_s = trait_genes.shape[0]
        n_ctrl, n_sc = ctrl_genes.shape
        assert ctrl_gw.shape[0] == n_sc and trait_gw.shape[0] == n_s
        n = self.n_cell
        out = dict(
            raw_score=np.empty(n, np.float64), norm_score=np.empty(n, np.float32),
            mc_count=np.empty(n, np.int32), ge_count=np.empty(n, np.int64),
            ctrl_raw=np.empty((n, n_ctrl), np.float32) if want_ctrl_raw else None,
            ctrl_norm=np.empty((n, n_ctrl), np.float32) if want_ctrl_norm else None,
        )
        check(self._L.scdrs_score_trait(
            self._h, n_ctrl, n_s, n_sc, ptr(trait_genes), ptr(trait_gw), ptr(ctrl_genes), ptr(ctrl_gw),
            WEIGHT_OPT[weight_opt], int(bool(use_var_ratio)), ptr(out["raw_score"]),
            ptr(out["norm_score"]), ptr(out["mc_count"]), ptr(out["ge_count"]), ptr(out["ctrl_raw"]),
            ptr(out["ctrl_norm"])))
        return out

    def score_trait_submit(self, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_var_ratio,
                           want_ctrl_raw=False, want_ctrl_norm=False):
        """Enqueue one trait without waiting for it; returns a ticket for ``score_trait_collect``.  At most two
        tickets may be outstanding."""
        trait_genes = as_c(trait_genes, np.int32)
        ctrl_genes = as_c(ctrl_genes, np.int32)
        trait_gw, ctrl_gw = as_c(trait_gw, np.float64), as_c(ctrl_gw, np.float64)
        n_s = trait_genes.shape[0]
        n_ctrl, n_sc = ctrl_genes.shape
        assert ctrl_gw.shape[0] == n_sc and trait_gw.shape[0] == n_s
        slot = C.c_int32(-1)
        check(self._L.scdrs_score_trait_submit(
            self._h, n_ctrl, n_s, n_sc, ptr(trait_genes), ptr(trait_gw), ptr(ctrl_genes), ptr(ctrl_gw),
            WEIGHT_OPT[weight_opt], int(bool(use_var_ratio)), int(bool(want_ctrl_raw)), int(bool(want_ctrl_norm)),
            C.byref(slot)))
        as_text = bool(want_ctrl_norm) and getattr(self, "text_output", False) and n_ctrl > 0
        return (slot.value, n_ctrl, bool(want_ctrl_raw), bool(want_ctrl_norm) and not as_text, as_text)

    def set_text_output(self, on):
        """Traits submitted with want_ctrl_norm leave the device as text slots (scdrs_ctx_set_text_output)."""
        check(self._L.scdrs_ctx_set_text_output(self._h, int(bool(on))))
        self.text_output = bool(on)

    def set_resident_sets(self, on):
        """Device-side control sets given to trait_raw_dev were complete before the previous trait was submitted
        (scdrs_ctx_set_resident_sets): their list building may overlap with that trait's tail."""
        check(self._L.scdrs_ctx_set_resident_sets(self._h, int(bool(on))))

    def slot_text(self, slot, n_ctrl):
        """uint8 view [n_cell, n_ctrl, 16] of the pinned text block of a collected slot (valid until the submit after next)."""
        p, nb = C.c_void_p(), C.c_int64()
        check(self._L.scdrs_slot_text(self._h, int(slot), C.byref(p), C.byref(nb)))
        assert nb.value == self.n_cell * n_ctrl * 16
        buf = (C.c_uint8 * nb.value).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint8).reshape(self.n_cell, n_ctrl, 16)

    def score_trait_collect(self, ticket):
        slot, n_ctrl, want_ctrl_raw, want_ctrl_norm = ticket[:4]
        as_text = len(ticket) > 4 and ticket[4]
        n = self.n_cell
        out = dict(
            raw_score=np.empty(n, np.float64), norm_score=np.empty(n, np.float32),
            mc_count=np.empty(n, np.int32), ge_count=np.empty(n, np.int64),
            ctrl_raw=np.empty((n, n_ctrl), np.float32) if want_ctrl_raw else None,
            ctrl_norm=np.empty((n, n_ctrl), np.float32) if want_ctrl_norm else None,
        )
        check(self._L.scdrs_score_trait_collect(
            self._h, slot, ptr(out["raw_score"]), ptr(out["norm_score"]), ptr(out["mc_count"]), ptr(out["ge_count"]),
            ptr(out["ctrl_raw"]), ptr(out["ctrl_norm"])))
        if as_text:
            out["ctrl_norm_text"] = self.slot_text(slot, n_ctrl)
            out["slot"] = slot
        return out

    # -- phases (device pointers are plain ints) --
    def trait_raw(self, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_var_ratio, d_colsum):
        trait_genes, ctrl_genes = as_c(trait_genes, np.int32), as_c(ctrl_genes, np.int32)
        trait_gw, ctrl_gw = as_c(trait_gw, np.float64), as_c(ctrl_gw, np.float64)
        check(self._L.scdrs_trait_raw(self._h, ctrl_genes.shape[0], trait_genes.shape[0], ctrl_genes.shape[1], ptr(trait_genes),
                                      ptr(trait_gw), ptr(ctrl_genes), ptr(ctrl_gw), WEIGHT_OPT[weight_opt],
                                      int(bool(use_var_ratio)), int(d_colsum)))

    def trait_raw_dev(self, trait_genes, trait_gw, d_ctrl_genes, n_ctrl, n_sc, ctrl_gw, weight_opt, use_var_ratio,
                      d_colsum):
        """Phase 1 with the control sets [n_ctrl x n_sc] int32 already in device memory."""
        trait_genes = as_c(trait_genes, np.int32)
        trait_gw, ctrl_gw = as_c(trait_gw, np.float64), as_c(ctrl_gw, np.float64)
        assert ctrl_gw.shape[0] == n_sc
        check(self._L.scdrs_trait_raw_dev(self._h, int(n_ctrl), trait_genes.shape[0], int(n_sc), ptr(trait_genes),
                                          ptr(trait_gw), int(d_ctrl_genes), ptr(ctrl_gw), WEIGHT_OPT[weight_opt],
                                          int(bool(use_var_ratio)), int(d_colsum)))

    def trait_finish_submit(self, d_hist_all, n_query_all, query_offset, n_null_total, n_ctrl,
                            want_ctrl_raw=False, want_ctrl_norm=False):
        """Phase 5 without waiting; returns a ticket for ``score_trait_collect``."""
        slot = C.c_int32(-1)
        check(self._L.scdrs_trait_finish_submit(self._h, int(d_hist_all), int(n_query_all), int(query_offset),
                                                int(n_null_total), int(bool(want_ctrl_raw)),
                                                int(bool(want_ctrl_norm)), C.byref(slot)))
        return (slot.value, n_ctrl, bool(want_ctrl_raw), bool(want_ctrl_norm))

    def trait_rowstats(self, d_colsum, n_cell_total, d_colsum2, d_colmin):
        check(self._L.scdrs_trait_rowstats(self._h, int(d_colsum), int(n_cell_total), int(d_colsum2), int(d_colmin)))

    def trait_queries(self, d_colsum2, d_colmin, d_query):
        check(self._L.scdrs_trait_queries(self._h, int(d_colsum2), int(d_colmin), int(d_query)))

    def trait_count(self, d_query_all, n_query_all, d_hist):
        check(self._L.scdrs_trait_count(self._h, int(d_query_all), int(n_query_all), int(d_hist)))

    def trait_finish(self, d_hist_all, n_query_all, query_offset, n_null_total, n_ctrl,
                     want_ctrl_raw=False, want_ctrl_norm=False):
        n = self.n_cell
        out = dict(
            raw_score=np.empty(n, np.float64), norm_score=np.empty(n, np.float32),
            mc_count=np.empty(n, np.int32), ge_count=np.empty(n, np.int64),
            ctrl_raw=np.empty((n, n_ctrl), np.float32) if want_ctrl_raw else None,
            ctrl_norm=np.empty((n, n_ctrl), np.float32) if want_ctrl_norm else None,
        )
        check(self._L.scdrs_trait_finish(
            self._h, int(d_hist_all), int(n_query_all), int(query_offset), int(n_null_total),
            ptr(out["raw_score"]), ptr(out["norm_score"]), ptr(out["mc_count"]), ptr(out["ge_count"]),
            ptr(out["ctrl_raw"]), ptr(out["ctrl_norm"])))
        return out

    # -- stand-alone helpers --
    def compute_raw_score(self, genes, w):
        genes, w = as_c(genes, np.int32), as_c(w, np.float64)
        if genes.ndim == 1:
            genes, w = genes[None, :], w[None, :]
        n_set, n_s = genes.shape
        out = np.empty((self.n_cell, n_set), np.float64)
        check(self._L.scdrs_compute_raw_score(self._h, n_set, n_s, ptr(genes), ptr(w), ptr(out)))
        return out

    def correct_background(self, v_raw, mat_ctrl, var_ratio):
        v_raw, mat_ctrl = as_c(v_raw, np.float64), as_c(mat_ctrl, np.float64)
        var_ratio = as_c(var_ratio, np.float64)
        n_cell, n_ctrl = mat_ctrl.shape
        v_norm, mat_norm = np.empty(n_cell, np.float64), np.empty((n_cell, n_ctrl), np.float64)
        check(self._L.scdrs_correct_background(self._h, n_cell, n_ctrl, ptr(v_raw), ptr(mat_ctrl),
                                               ptr(var_ratio), ptr(v_norm), ptr(mat_norm)))
        return v_norm, mat_norm

    def empi_null_count(self, v_t, v_null):
        v_t, v_null = np.asarray(v_t), np.asarray(v_null)
        f32 = v_t.dtype == np.float32 and v_null.dtype == np.float32
        dt = np.float32 if f32 else np.float64
        v_t, v_null = as_c(v_t.reshape(-1), dt), as_c(v_null.reshape(-1), dt)
        out = np.empty(v_t.shape[0], np.int64)
        fn = self._L.scdrs_empi_null_count_f32 if f32 else self._L.scdrs_empi_null_count_f64
        check(fn(self._h, v_t.shape[0], ptr(v_t), v_null.shape[0], ptr(v_null), ptr(out)))
        return out

    def gene_mean_xtc(self, cov_mat):
        """(X.mean(axis=0) as float32, (C^T X)^T as float64 [n_gene x k])."""
        cov_mat = as_c(cov_mat, np.float64)
        k = cov_mat.shape[1]
        gm = np.empty(self.n_gene, np.float32)
        xtc = np.empty((self.n_gene, k), np.float64)
        check(self._L.scdrs_gene_mean_xtc(self._h, k, ptr(cov_mat), ptr(gm), ptr(xtc)))
        return gm, xtc

    def gene_cell_stats(self, transform, cell_weight=None, want_gene=True, want_cell=True, center=False, denom=0.0):
        w = None if cell_weight is None else as_c(cell_weight, np.float64)
        gs = np.empty(self.n_gene) if want_gene else None
        gq = np.empty(self.n_gene) if want_gene else None
        cs = np.empty(self.n_cell) if want_cell else None
        cq = np.empty(self.n_cell) if want_cell else None
        check(self._L.scdrs_gene_cell_stats(self._h, int(transform), ptr(w), int(bool(center)), float(denom),
                                            ptr(gs), ptr(gq), ptr(cs), ptr(cq)))
        return gs, gq, cs, cq

    def gene_moments(self, mode, cov_mat=None, cell_weight=None):
        """Per-gene moments of this context's cells (scdrs_gene_moments): mode 0 -> [n_gene x (2 + k)] =
        sum w x, sum w x^2, sum w x C[:, j]; mode 1 -> [n_gene x 2] = sum w expm1(x), sum w expm1(x)^2."""
        k = 0
        if mode == 0 and cov_mat is not None:
            cov_mat = as_c(cov_mat, np.float64)
            assert cov_mat.shape[0] == self.n_cell
            k = cov_mat.shape[1]
        w = None if cell_weight is None else as_c(cell_weight, np.float64)
        out = np.empty((self.n_gene, 2 + k if mode == 0 else 2), np.float64)
        check(self._L.scdrs_gene_moments(self._h, int(mode), int(k), ptr(cov_mat) if k else None, ptr(w), ptr(out)))
        return out

    def gene_expm1_implicit(self, shift, cell_weight=None):
        """[n_gene x 2] = sum_c w (expm1(v) - shift), sum_c w (expm1(v) - shift)^2 over all cells (implicit mode)."""
        shift = as_c(shift, np.float64)
        assert shift.shape[0] == self.n_gene
        w = None if cell_weight is None else as_c(cell_weight, np.float64)
        out = np.empty((self.n_gene, 2), np.float64)
        check(self._L.scdrs_gene_expm1_implicit(self._h, ptr(shift), ptr(w), ptr(out)))
        return out

    def gene_stats_exact(self, transform, genes, gene_sum, gene_sumsq, cell_weight=None, center=False, denom=0.0,
                         mean_in=None):
        """Overwrite gene_sum / gene_sumsq (float64 [n_gene], in place) for the tiles that hold `genes` with the
        sequential, numpy-ordered sums of scdrs_gene_cell_stats.  Returns the gene indices that were recomputed."""
        ts = int(self._L.scdrs_gene_tile_size())
        tiles = np.unique(np.asarray(genes, dtype=np.int64) // ts).astype(np.int32)
        w = None if cell_weight is None else as_c(cell_weight, np.float64)
        m = None if mean_in is None else as_c(mean_in, np.float64)
        assert gene_sum.dtype == np.float64 and gene_sumsq.dtype == np.float64
        check(self._L.scdrs_gene_stats_exact(self._h, int(transform), ptr(w), int(bool(center)), float(denom),
                                             tiles.shape[0], ptr(tiles), ptr(m), ptr(gene_sum), ptr(gene_sumsq)))
        idx = (tiles[:, None].astype(np.int64) * ts + np.arange(ts)[None, :]).reshape(-1)
        return idx[idx < self.n_gene]

    # ---- downstream analyses of one trait's score matrix (include/scdrs_b200.h) ----
    def ds_set_scores(self, frame, cols=None):
        """frame[n_cell x m] float32 or float64, C- or F-contiguous (uploaded as it lies in memory); cols: the
        columns that make up the score matrix (norm_score first, then the control columns; default all)."""
        frame = np.asarray(frame)
        assert frame.ndim == 2
        if frame.dtype not in (np.float32, np.float64):
            frame = frame.astype(np.float64)
        if not (frame.flags["C_CONTIGUOUS"] or frame.flags["F_CONTIGUOUS"]):
            frame = np.ascontiguousarray(frame)
        n, m = frame.shape
        rs, cs = (m, 1) if frame.flags["C_CONTIGUOUS"] else (1, n)
        cols = as_c(np.arange(m) if cols is None else cols, np.int32)
        check(self._L.scdrs_ds_set_scores(self._h, n, cols.shape[0], frame.ctypes.data, int(frame.dtype == np.float32),
                                          n * m, rs, cs, ptr(cols)))
        self.ds_shape = (n, cols.shape[0])

    def ds_corr(self, variables):
        """variables[n_var x n_cell] -> Pearson correlation with every score column [n_var x ncol]."""
        v = as_c(np.atleast_2d(variables), np.float64)
        assert v.shape[1] == self.ds_shape[0]
        out = np.empty((v.shape[0], self.ds_shape[1]), np.float64)
        check(self._L.scdrs_ds_corr(self._h, v.shape[0], ptr(v), ptr(out)))
        return out

    def ds_group_stats(self, order, grp_ptr, q_lo=None, q_gamma=None, geary_mode=0, graph=None):
        """(quantile, geary_total, geary_denom), each [n_group x ncol] or None.  graph = (indptr, nbr, w) of
        the within-group edges with cells in group order."""
        order, grp_ptr = as_c(order, np.int64), as_c(grp_ptr, np.int64)
        n_group = grp_ptr.shape[0] - 1
        shape = (n_group, self.ds_shape[1])
        quant = None
        if q_lo is not None:
            q_lo, q_gamma = as_c(q_lo, np.int32), as_c(q_gamma, np.float64)
            quant = np.empty(shape, np.float64)
        total = denom = gi = gn = gw = None
        if geary_mode:
            gi, gn, gw = as_c(graph[0], np.int64), as_c(graph[1], np.int32), as_c(graph[2], np.float64)
            total, denom = np.empty(shape, np.float64), np.empty(shape, np.float64)
        check(self._L.scdrs_ds_group_stats(self._h, n_group, ptr(order), ptr(grp_ptr), ptr(q_lo), ptr(q_gamma),
                                           ptr(quant), int(geary_mode), ptr(gi), ptr(gn), ptr(gw), ptr(total),
                                           ptr(denom)))
        return quant, total, denom


def select_ctrl(seed, bin_ptr, bin_genes, bin_ntrait, n_ctrl, n_s, want_state=False):
    """Host-only: numpy-legacy-RNG control gene draws (see include/scdrs_b200.h).  With ``want_state`` also the
    generator after the draws, as the tuple ``np.random.set_state`` takes."""
    bin_ptr, bin_genes = as_c(bin_ptr, np.int32), as_c(bin_genes, np.int32)
    bin_ntrait = as_c(bin_ntrait, np.int32)
    out = np.empty((n_ctrl, n_s), np.int32)
    state = np.empty(625, np.uint32) if want_state else None
    check(lib().scdrs_select_ctrl_state(int(seed) & 0xFFFFFFFF, bin_ntrait.shape[0], ptr(bin_ptr), ptr(bin_genes),
                                        ptr(bin_ntrait), int(n_ctrl), int(n_s), ptr(out), ptr(state)))
    if want_state:
        return out, ("MT19937", state[:624].copy(), int(state[624]), 0, 0.0)
    return out


_NAME_BLOBS = {}


def _index_blob(index):
    """(concatenated UTF-8 row names, offsets) of a pandas index; cached for the index object last seen (a run
    writes the same 1e5 .. 1e6 cell names once or twice per trait)."""
    key = id(index)
    hit = _NAME_BLOBS.get(key)
    if hit is not None and hit[0] is index:
        return hit[1], hit[2]
    names = [str(x).encode("utf-8") for x in index]
    off = np.zeros(len(names) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in names], out=off[1:])
    blob = b"".join(names)
    buf = C.create_string_buffer(blob, len(blob) + 1)
    _NAME_BLOBS.clear()
    _NAME_BLOBS[key] = (index, buf, off)
    return buf, off


def gzip_level():
    """Deflate level of the score files: SCDRS_B200_GZIP_LEVEL (0 - 9), default 1.  The text inside is the same at
    every level; level 1 deflates 5x faster than zlib's default 6 for files 10 % larger (pandas writes level 9)."""
    import os

    try:
        return min(9, max(0, int(os.environ.get("SCDRS_B200_GZIP_LEVEL", "1"))))
    except ValueError:
        return 1


def write_score_gz(path, df, n_thread=None, level=None):
    """Write a float32 DataFrame as the gzip TSV ``DataFrame.to_csv(path, sep="\\t", compression="gzip")`` gives
    (the same text after gunzip)."""
    import os

    values = df.values
    if values.dtype != np.float32 or not values.flags["C_CONTIGUOUS"]:
        values = as_c(values, np.float32)
    buf, off = _index_blob(df.index)
    header = "\t".join([df.index.name or ""] + [str(c) for c in df.columns]) + "\n"
    n_thread = n_thread or max(1, min(32, (os.cpu_count() or 2)))
    level = gzip_level() if level is None else level
    check(lib().scdrs_write_score_gz(os.fsencode(path), header.encode("utf-8"), values.shape[0], values.shape[1],
                                     C.cast(buf, C.c_void_p), ptr(off), ptr(values), int(n_thread), int(level)))


def write_score_gz_slots(path, df_lead, extra_columns, slots, n_thread=None, level=None):
    """The score file whose trailing columns the GPU already formatted: ``df_lead`` (float32 frame) is printed by the
    host threads, ``slots`` uint8 [n_row, n_extra, 16] are copied (scdrs_write_score_gz_slots)."""
    import os

    values = df_lead.values
    if values.dtype != np.float32 or not values.flags["C_CONTIGUOUS"]:
        values = as_c(values, np.float32)
    assert slots.dtype == np.uint8 and slots.flags["C_CONTIGUOUS"] and slots.shape[0] == values.shape[0]
    buf, off = _index_blob(df_lead.index)
    header = "\t".join([df_lead.index.name or ""] + [str(c) for c in df_lead.columns] + list(extra_columns)) + "\n"
    n_thread = n_thread or max(1, min(32, (os.cpu_count() or 2)))
    level = gzip_level() if level is None else level
    check(lib().scdrs_write_score_gz_slots(os.fsencode(path), header.encode("utf-8"), values.shape[0],
                                           C.cast(buf, C.c_void_p), ptr(off), values.shape[1], ptr(values),
                                           slots.shape[1], slots.ctypes.data, int(n_thread), int(level)))


def gzip_member(text, level=1):
    """One gzip member as the score-file writer emits it (level 1: the literal-only Huffman encoder; else zlib)."""
    text = bytes(text)
    cap = len(text) + len(text) // 8 + 1024
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_int64(0)
    check(lib().scdrs_gzip_member(text, len(text), int(level), out.ctypes.data, cap, C.byref(n)))
    return out[: n.value].tobytes()


def format_f32(v):
    buf = C.create_string_buffer(64)
    n = lib().scdrs_format_f32(float(np.float32(v)), buf)
    return buf.raw[:n].decode()
