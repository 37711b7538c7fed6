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

    def __init__(self, phases, device):
        self.p = phases
        self.device = device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        n_local = int(phases.n_cell)
        if self.world > 1:
            sizes = torch.zeros(self.world, dtype=torch.int64, device=device)
            sizes[self.rank] = n_local
            dist.all_reduce(sizes)
            self.sizes = [int(s) for s in sizes.tolist()]
        else:
            self.sizes = [n_local]
        self.n_total = sum(self.sizes)
        self.offset = sum(self.sizes[: self.rank])
        self._buf = {}

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
            dist.all_gather_into_tensor(q_all, q_local)
        else:
            # ragged shards: gather blocks padded to the largest shard, then compact
            m = max(self.sizes)
            pad = self._tensor("q_pad", m, torch.float32)
            pad[: q_local.shape[0]].copy_(q_local)
            blocks = self._tensor("q_blocks", m * self.world, torch.float32)
            dist.all_gather_into_tensor(blocks, pad)
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
        return self.p.finish_submit(hist, self.n_total, self.offset, self.n_total * n_ctrl, n_ctrl,
                                    want_ctrl_raw, want_ctrl_norm)

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

        self.p.raw(trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio, colsum)
        if self.world > 1:
            dist.all_reduce(colsum)
        self.p.rowstats(colsum, self.n_total, col2, colmin)
        if self.world > 1:
            dist.all_reduce(col2)
            dist.all_reduce(colmin, op=dist.ReduceOp.MIN)
        self.p.queries(col2, colmin, q_local)
        q_all = self._gather_queries(q_local)
        self.p.count(q_all, hist)
        if self.world > 1:
            dist.all_reduce(hist)
        return hist, n_ctrl


class DrawBroadcaster:
    """Ships the control sets of one trait (the only random part of ``method._prepare_trait``) from the rank that
    drew them into a device buffer on every rank, without a host round trip on the receivers: pinned staging on
    the owner, NCCL broadcast on the stream the kernels run on.  Two buffers alternate (two traits in flight)."""

    def __init__(self, device, cap):
        self.device = device
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
        dist.broadcast(buf, src=src)
        return buf.view(n_ctrl, n_sc)


def broadcast_prep(prep, src, n_ctrl, n_s_max, device):
    """Ship one trait's host-drawn gene sets (``method._prepare_trait``) from rank ``src`` to all ranks.

    One int32 buffer over NCCL: [n_s, n_sc, trait_cols, ctrl_cols, trait_gw (f64 bits), ctrl_gw (f64 bits)],
    sized for the largest possible trait (n_s_max genes) so that receivers need no size exchange."""
    cap = 2 + n_s_max + n_ctrl * n_s_max + 4 * n_s_max
    if prep is not None:
        n_s, n_sc = int(prep["trait_cols"].shape[0]), int(prep["ctrl_cols"].shape[1])
        host = np.empty(cap, dtype=np.int32)
        host[0], host[1] = n_s, n_sc
        o = 2
        host[o:o + n_s] = prep["trait_cols"]; o += n_s
        host[o:o + n_ctrl * n_sc] = np.ascontiguousarray(prep["ctrl_cols"], dtype=np.int32).reshape(-1); o += n_ctrl * n_sc
        host[o:o + 2 * n_s] = np.ascontiguousarray(prep["trait_gw"], dtype=np.float64).view(np.int32); o += 2 * n_s
        host[o:o + 2 * n_sc] = np.ascontiguousarray(prep["ctrl_gw"], dtype=np.float64).view(np.int32)
        buf = torch.from_numpy(host).to(device)
    else:
        buf = torch.empty(cap, dtype=torch.int32, device=device)
    dist.broadcast(buf, src=src)
    if prep is not None:
        return prep
    host = buf.cpu().numpy()
    n_s, n_sc = int(host[0]), int(host[1])
    o = 2
    trait_cols = host[o:o + n_s].copy(); o += n_s
    ctrl_cols = host[o:o + n_ctrl * n_sc].reshape(n_ctrl, n_sc).copy(); o += n_ctrl * n_sc
    trait_gw = host[o:o + 2 * n_s].copy().view(np.float64); o += 2 * n_s
    ctrl_gw = host[o:o + 2 * n_sc].copy().view(np.float64)
    return dict(trait_cols=trait_cols, trait_gw=trait_gw, ctrl_cols=ctrl_cols, ctrl_gw=ctrl_gw)
