"""Cells sharded across the GPUs of one box (one process per GPU, ``torch.distributed``).

The reference has no distributed path (single process; SLURM array jobs over ``.gs`` batches,
experiments/job.compute_score/*.sh).  Cells are independent units of the scoring path, so every
rank keeps a contiguous block of cells (its own CSR shard) and runs the same five device phases
(``include/scdrs_b200.h``); only small per-trait statistics cross NVLink:

    phase 1  raw scores              -> all-reduce(sum)  of 1 + n_ctrl column sums          (f64)
    phase 2  row statistics          -> all-reduce(sum)  of 1 + n_ctrl column sums,
                                        all-reduce(min)  of 1 + n_ctrl column minima        (f64)
    phase 3  trait scores (queries)  -> all-gather       of n_cell float32 scores
    phase 4  pooled-null histogram   -> all-reduce(sum)  of n_cell + 1 int64 counts
    phase 5  suffix sums, local slice of the ranks

The n_cell x n_ctrl null values never leave their GPU: every rank counts its own null values
against ALL ranks' (sorted) trait scores, and the global rank of a trait score is the sum of the
per-rank ranks -- the splitters of the pooled "sample sort" are the gathered trait scores and the
exchanged counts are the histogram.

The collectives are plain ``torch.distributed`` calls on torch tensors whose storage the C ABI
reads and writes through raw device pointers; the phase object is pluggable so that the exchange
logic is testable on CPU with the ``gloo`` backend (tests/test_distributed.py).
"""
import numpy as np
import torch
import torch.distributed as dist


class CudaPhases:
    """The five device phases of ``libscdrs_b200.so`` on torch CUDA tensors."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = device
        self.n_cell = ctx.n_cell
        # kernels and NCCL collectives are ordered on torch's current stream
        ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)

    def raw(self, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio, colsum):
        if isinstance(ctrl_genes, torch.Tensor):      # already on the device (the receive buffer of a broadcast)
            assert ctrl_genes.is_cuda and ctrl_genes.dtype == torch.int32 and ctrl_genes.is_contiguous()
            self.ctx.trait_raw_dev(trait_genes, trait_gw, ctrl_genes.data_ptr(), ctrl_genes.shape[0], ctrl_genes.shape[1],
                                   ctrl_gw, weight_opt, use_ratio, colsum.data_ptr())
        else:
            self.ctx.trait_raw(trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio, colsum.data_ptr())

    def rowstats(self, colsum, n_cell_total, col2, colmin):
        self.ctx.trait_rowstats(colsum.data_ptr(), n_cell_total, col2.data_ptr(), colmin.data_ptr())

    def queries(self, col2, colmin, q_local):
        self.ctx.trait_queries(col2.data_ptr(), colmin.data_ptr(), q_local.data_ptr())

    def count(self, q_all, hist):
        self.ctx.trait_count(q_all.data_ptr(), q_all.shape[0], hist.data_ptr())

    def finish(self, hist, n_query_all, offset, n_null_total, n_ctrl, want_ctrl_raw, want_ctrl_norm):
        return self.ctx.trait_finish(hist.data_ptr(), n_query_all, offset, n_null_total, n_ctrl,
                                     want_ctrl_raw, want_ctrl_norm)


    def finish_submit(self, hist, n_query_all, offset, n_null_total, n_ctrl, want_ctrl_raw, want_ctrl_norm):
        return self.ctx.trait_finish_submit(hist.data_ptr(), n_query_all, offset, n_null_total, n_ctrl,
                                            want_ctrl_raw, want_ctrl_norm)

    def collect(self, ticket):
        return self.ctx.score_trait_collect(ticket)


class ShardedScorer:
    """Scores one trait over cells sharded across the ranks of the default process group."""

    def __init__(self, phases, device, group=None):
        """group: the ranks that share one atlas (None: the default group; False: this process alone, whatever
        torch.distributed says -- trait-sharded runs, where every rank holds all cells)."""
        self.p = phases
        self.device = device
        self.group = None if group is False else group
        alone = group is False or not dist.is_initialized()
        self.world = 1 if alone else dist.get_world_size(self.group)
        self.rank = 0 if alone else dist.get_rank(self.group)
        n_local = int(phases.n_cell)
        if self.world > 1:
            sizes = torch.zeros(self.world, dtype=torch.int64, device=device)
            sizes[self.rank] = n_local
            dist.all_reduce(sizes, group=self.group)
            self.sizes = [int(s) for s in sizes.tolist()]
        else:
            self.sizes = [n_local]
        self.n_total = sum(self.sizes)
        self.offset = sum(self.sizes[: self.rank])
        self._buf = {}
        self.timing = None        # dict -> per-phase CUDA-event times are collected (bench.py)
        self._marks = []

    # ---- optional per-phase timing: CUDA events on the compute stream around phases and collectives ----
    def _mark(self, name):
        if self.timing is None or self.device.type != "cuda":
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._marks.append((name, ev))

    def _close_marks(self):
        if self.timing is None or len(self._marks) < 2:
            self._marks = []
            return
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(self._marks[:-1], self._marks[1:]):
            self.timing.setdefault(name, []).append(e0.elapsed_time(e1))
        self._marks = []

    def timing_summary(self):
        """{phase_ms: mean over the timed traits} (None when nothing was timed)."""
        if not self.timing:
            return None
        return {k + "_ms": float(np.mean(v)) for k, v in self.timing.items()}

    def _tensor(self, name, n, dtype):
        t = self._buf.get(name)
        if t is None or t.shape[0] != n or t.dtype != dtype:
            t = torch.empty(n, dtype=dtype, device=self.device)
            self._buf[name] = t
        return t

    def _gather_queries(self, q_local):
        if self.world == 1:
            return q_local
        q_all = self._tensor("q_all", self.n_total, torch.float32)
        if len(set(self.sizes)) == 1:
            dist.all_gather_into_tensor(q_all, q_local, group=self.group)
        else:
            # ragged shards: gather blocks padded to the largest shard, then compact
            m = max(self.sizes)
            pad = self._tensor("q_pad", m, torch.float32)
            pad[: q_local.shape[0]].copy_(q_local)
            blocks = self._tensor("q_blocks", m * self.world, torch.float32)
            dist.all_gather_into_tensor(blocks, pad, group=self.group)
            off = 0
            for r, n in enumerate(self.sizes):
                q_all[off: off + n].copy_(blocks[r * m: r * m + n])
                off += n
        return q_all

    def score(self, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio,
              want_ctrl_raw=False, want_ctrl_norm=False):
        """Returns the dict of ``Context.score_trait`` for this rank's cells, with pooled-null
        counts over ALL ranks' cells (``n_null_total`` = total cells x n_ctrl)."""
        hist, n_ctrl = self._phases_1_to_4(trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio)
        n_null_total = self.n_total * n_ctrl
        out = self.p.finish(hist, self.n_total, self.offset, n_null_total, n_ctrl, want_ctrl_raw, want_ctrl_norm)
        out["n_null_total"] = n_null_total
        return out

    def submit(self, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio,
               want_ctrl_raw=False, want_ctrl_norm=False):
        """``score`` without waiting for the GPU: everything (collectives included) is queued on the stream and
        the results land in pinned memory; ``collect(ticket)`` returns them.  At most two tickets outstanding."""
        hist, n_ctrl = self._phases_1_to_4(trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio)
        ticket = self.p.finish_submit(hist, self.n_total, self.offset, self.n_total * n_ctrl, n_ctrl,
                                      want_ctrl_raw, want_ctrl_norm)
        self._mark("finish")
        self._close_marks()
        return ticket

    def collect(self, ticket):
        out = self.p.collect(ticket)
        out["n_null_total"] = self.n_total * ticket[1]
        return out

    def _phases_1_to_4(self, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio):
        if not isinstance(ctrl_genes, torch.Tensor):
            ctrl_genes = np.asarray(ctrl_genes)
        n_ctrl = ctrl_genes.shape[0]
        ncol = n_ctrl + 1
        colsum = self._tensor("colsum", ncol, torch.float64)
        col2 = self._tensor("col2", ncol, torch.float64)
        colmin = self._tensor("colmin", ncol, torch.float64)
        q_local = self._tensor("q_local", self.sizes[self.rank], torch.float32)
        hist = self._tensor("hist", self.n_total + 1, torch.int64)

        self._mark("start")
        self.p.raw(trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio, colsum)
        self._mark("raw")
        if self.world > 1:
            dist.all_reduce(colsum, group=self.group)
            self._mark("colsum_allreduce")
        self.p.rowstats(colsum, self.n_total, col2, colmin)
        self._mark("rowstats")
        if self.world > 1:
            dist.all_reduce(col2, group=self.group)
            dist.all_reduce(colmin, op=dist.ReduceOp.MIN, group=self.group)
            self._mark("col2_min_allreduce")
        self.p.queries(col2, colmin, q_local)
        self._mark("queries")
        q_all = self._gather_queries(q_local)
        if self.world > 1:
            self._mark("query_allgather")
        self.p.count(q_all, hist)
        self._mark("count")
        if self.world > 1:
            dist.all_reduce(hist, group=self.group)
            self._mark("hist_allreduce")
        return hist, n_ctrl


class DrawBroadcaster:
    """Ships the control sets of one trait (the only random part of ``method._prepare_trait``) from the rank that
    drew them into a device buffer on every rank, without a host round trip on the receivers: pinned staging on
    the owner, NCCL broadcast on the stream the kernels run on.  Two buffers alternate (two traits in flight)."""

    def __init__(self, device, cap, group=None):
        self.device = device
        self.group = group
        self.dev = [torch.empty(cap, dtype=torch.int32, device=device) for _ in range(2)]
        self.pin = [torch.empty(cap, dtype=torch.int32, pin_memory=(device.type == "cuda")) for _ in range(2)]
        self.k = 0

    def ship(self, ctrl_cols, src, n_ctrl, n_sc):
        """ctrl_cols: the drawn sets on rank ``src`` (None elsewhere).  Returns an int32 [n_ctrl, n_sc] device view
        that stays valid until the call after next."""
        k, self.k = self.k, self.k ^ 1
        m = n_ctrl * n_sc
        buf = self.dev[k][:m]
        if ctrl_cols is not None:
            self.pin[k][:m].copy_(torch.from_numpy(np.ascontiguousarray(ctrl_cols, dtype=np.int32).reshape(-1)))
            buf.copy_(self.pin[k][:m], non_blocking=True)
        dist.broadcast(buf, src=src if self.group is None else dist.get_global_rank(self.group, src), group=self.group)
        return buf.view(n_ctrl, n_sc)
