"""Batched device-side surface of the hot path (SURVEY.md §8b "new batched surface").

Every function takes CUDA tensors, hands raw pointers to the C ABI on the current torch
stream and returns CUDA tensors.  torch is plumbing here (allocation, streams, autograd
bookkeeping); all arithmetic is in libsbir_b200.so.  There is no CPU path: CPU tensors or
a missing library raise.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _binding as B

_METRICS = {"euclidean": B.SBIR_EUCLIDEAN, "cosine": B.SBIR_COSINE}


def metric_id(loss_type: str) -> int:
    """Maps the reference's `loss_type` strings (inference.py:43-48, train.py:164-175)."""
    try:
        return _METRICS[loss_type]
    except KeyError:
        # same message as the reference (inference.py:48)
        raise Exception(f"loss type not correct {loss_type}") from None


def _dtype_id(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return B.SBIR_F32
    if t.dtype == torch.bfloat16:
        return B.SBIR_BF16
    raise TypeError(f"sbir_b200 kernels take float32 or bfloat16 embeddings, got {t.dtype}")


def _dev(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the sbir_b200 path has no CPU fallback")
    return t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _workspace(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator returns 512-byte aligned blocks; the ABI asks for 256.
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------- H9 ----
def l2_normalize(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """x / max(‖x‖₂, eps) per row (the normalisation inside nn.CosineSimilarity, utils.py:34)."""
    x = _dev(x, "x")
    if x.dim() != 2:
        raise ValueError("l2_normalize expects [rows, dim]")
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        B.check(B.load().sbir_l2_normalize(x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], _dtype_id(x),
                                           float(eps), _stream()), "sbir_l2_normalize")
    return y


def row_sqnorm(x: torch.Tensor) -> torch.Tensor:
    x = _dev(x, "x")
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        B.check(B.load().sbir_row_sqnorm(x.data_ptr(), x.shape[0], x.shape[1], _dtype_id(x), out.data_ptr(),
                                         _stream()), "sbir_row_sqnorm")
    return out


# --------------------------------------------------------------------------- N1 ----
class GalleryBuffer:
    """Preallocated device gallery [capacity, dim] (fp32 or bf16) + ‖row‖² that encoder output is appended
    to block by block (sbir_gallery_append) — what replaces the reference's torch.cat growth, .cpu() and
    CSV dump in compute_image_features (inference.py:72-92).  `rows` / `sqnorm` are views of what has been
    written; the buffer never leaves the GPU.  With one encoder replica per GPU every rank fills the buffer
    of its own shard (sharded.shard_bounds) and the row-sharded layout of the multi-GPU path exists at once."""

    def __init__(self, capacity: int, dim: int, dtype: torch.dtype = torch.float32, device=None, normalize: bool = False):
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("gallery storage is float32 or bfloat16")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("GalleryBuffer lives on a CUDA device: the sbir_b200 path has no CPU fallback")
        self.storage = torch.empty((capacity, dim), dtype=dtype, device=device)
        self.sqnorm_storage = torch.empty(capacity, dtype=torch.float32, device=device)
        self.normalize = bool(normalize)
        self.filled = 0

    def append(self, block: torch.Tensor) -> None:
        block = _dev(block.reshape(-1, block.shape[-1]), "block")
        if block.dtype not in (torch.float32, torch.bfloat16):
            block = block.float()
        n, d = block.shape
        if d != self.storage.shape[1] or block.device != self.storage.device:
            raise ValueError(f"block [{n},{d}] on {block.device} does not match the gallery {tuple(self.storage.shape)} on {self.storage.device}")
        if self.filled + n > self.storage.shape[0]:
            raise ValueError(f"gallery buffer of {self.storage.shape[0]} rows cannot take {n} more after {self.filled}")
        with torch.cuda.device(block.device):
            B.check(B.load().sbir_gallery_append(block.data_ptr(), _dtype_id(block), n, d, self.storage.data_ptr(),
                                                 _dtype_id(self.storage), self.storage.shape[0], self.filled,
                                                 self.sqnorm_storage.data_ptr(), int(self.normalize), _stream()),
                    "sbir_gallery_append")
        self.filled += n

    @property
    def rows(self) -> torch.Tensor:
        return self.storage[:self.filled]

    @property
    def sqnorm(self) -> torch.Tensor:
        return self.sqnorm_storage[:self.filled]


# ------------------------------------------------------------------------ H1 / H2 ----
class _PairwiseDistanceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, metric):
        x1c, x2c = _dev(x1, "x1"), _dev(x2, "x2")
        r1, r2, d = x1c.shape[0], x2c.shape[0], x1c.shape[1]
        out = torch.empty(max(r1, r2), dtype=torch.float32, device=x1c.device)
        with torch.cuda.device(x1c.device):
            B.check(B.load().sbir_pairwise_distance(x1c.data_ptr(), r1, x2c.data_ptr(), r2, d, _dtype_id(x1c),
                                                    metric, out.data_ptr(), _stream()), "sbir_pairwise_distance")
        ctx.save_for_backward(x1c, x2c)
        ctx.metric = metric
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x1, x2 = ctx.saved_tensors
        if x1.dtype != torch.float32:
            raise RuntimeError("pairwise distance backward is implemented for float32 embeddings")
        g1 = torch.empty_like(x1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(x2) if ctx.needs_input_grad[1] else None
        go = grad_out.contiguous().float()
        with torch.cuda.device(x1.device):
            B.check(B.load().sbir_pairwise_distance_bwd(x1.data_ptr(), x1.shape[0], x2.data_ptr(), x2.shape[0],
                                                        x1.shape[1], ctx.metric, go.data_ptr(), _ptr(g1),
                                                        _ptr(g2), _stream()), "sbir_pairwise_distance_bwd")
        return g1, g2, None


def pairwise_distance(x1: torch.Tensor, x2: torch.Tensor, loss_type: str = "euclidean") -> torch.Tensor:
    """Row-wise distance with torch broadcasting of a single row, differentiable.

    euclidean: ‖x1 − x2 + 1e-6‖₂ (nn.PairwiseDistance, utils.py:42);
    cosine: 1 − cos(x1, x2) (utils.CosineLoss, utils.py:31-40)."""
    if x1.dim() == 1:
        x1 = x1.unsqueeze(0)
    if x2.dim() == 1:
        x2 = x2.unsqueeze(0)
    if x1.dim() != 2 or x2.dim() != 2 or x1.shape[1] != x2.shape[1]:
        raise ValueError(f"pairwise_distance expects [rows, dim] operands, got {tuple(x1.shape)} and {tuple(x2.shape)}")
    if x1.shape[0] != x2.shape[0] and 1 not in (x1.shape[0], x2.shape[0]):
        raise ValueError("row counts must match or one operand must have a single row")
    if x1.dtype != x2.dtype:
        # the reference promotes fp32 queries against fp64 CSV-loaded galleries (F8); we compute in fp32
        x1, x2 = x1.float(), x2.float()
    return _PairwiseDistanceFn.apply(x1, x2, metric_id(loss_type))


# ------------------------------------------------------------- H1+H3+H4 batched ----
def _check_retrieval_args(queries: torch.Tensor, gallery: torch.Tensor, k: Optional[int] = None,
                          per_query: Sequence[Tuple[str, Optional[torch.Tensor]]] = ()):
    """Validation shared by every entry point that hands raw pointers of a (queries, gallery) pair to the
    C ABI: CUDA + contiguous, 2-D, equal dim, equal dtype (mixed dtypes — e.g. fp32 queries against a
    float64 CSV-loaded or bf16 gallery, F8 — are scored in fp32), same device, k in range, and per-query
    vectors of length Q on the same device.  Returns the (possibly converted) operands."""
    q, g = _dev(queries, "queries"), _dev(gallery, "gallery")
    if q.dim() != 2 or g.dim() != 2 or q.shape[1] != g.shape[1]:
        raise ValueError(f"expected [Q,D] and [N,D], got {tuple(q.shape)} and {tuple(g.shape)}")
    if q.device != g.device:
        raise ValueError(f"queries ({q.device}) and gallery ({g.device}) must live on the same device")
    if q.dtype != g.dtype or q.dtype not in (torch.float32, torch.bfloat16):
        q, g = q.float(), g.float()
    if k is not None and not 1 <= k <= B.MAX_K:
        raise ValueError(f"k must be in [1, {B.MAX_K}]")
    for name, t in per_query:
        if t is None:
            continue
        if t.dim() != 1 or t.shape[0] != q.shape[0]:
            raise ValueError(f"{name} must have one entry per query ({q.shape[0]}), got {tuple(t.shape)}")
        if t.device != q.device:
            raise ValueError(f"{name} ({t.device}) must live on the queries' device ({q.device})")
    return q, g


def _gallery_sqnorm(gallery_sqnorm: Optional[torch.Tensor], g: torch.Tensor, gallery: torch.Tensor):
    """Stored ‖g‖² handed to the kernels — only valid for the rows exactly as they are scored."""
    if gallery_sqnorm is None:
        return None
    if g.dtype != gallery.dtype:
        return None  # the gallery was converted for scoring (e.g. float64 CSV features): norms of other values
    sq = _dev(gallery_sqnorm, "gallery_sqnorm")
    if sq.dtype != torch.float32 or sq.dim() != 1 or sq.shape[0] != g.shape[0] or sq.device != g.device:
        raise ValueError("gallery_sqnorm must be float32 [num_gallery] on the gallery's device")
    return sq


def pairwise_topk(queries: torch.Tensor, gallery: torch.Tensor, k: int, loss_type: str = "euclidean",
                  pos_index: Optional[torch.Tensor] = None, index_offset: int = 0,
                  return_uncertified: bool = False, gallery_sqnorm: Optional[torch.Tensor] = None):
    """k nearest gallery rows per query (ascending distance, ties by index) and, when
    `pos_index` (int64 [num_q], <0 = no positive) is given, the 0-based rank of that
    gallery row: the batched form of inference.py:30-69.

    `gallery_sqnorm` (fp32 [N], optional): ‖g‖² of the rows as stored (GalleryBuffer / sidecar) — spares the
    pass its own read of the gallery for the norms.
    Returns (values fp32 [Q,k], indices int64 [Q,k]) or (values, indices, rank int64 [Q])."""
    want_rank = pos_index is not None
    pos = _dev(pos_index.to(torch.int64), "pos_index") if want_rank else None
    q, g = _check_retrieval_args(queries, gallery, k, (("pos_index", pos),))
    gsq = _gallery_sqnorm(gallery_sqnorm, g, gallery)
    metric = metric_id(loss_type)
    nq, ng, d = q.shape[0], g.shape[0], q.shape[1]
    dev = q.device
    vals = torch.empty((nq, k), dtype=torch.float32, device=dev)
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    rank = torch.empty(nq, dtype=torch.int64, device=dev) if want_rank else None
    unc = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = B.load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.sbir_pairwise_topk_workspace_bytes(nq, ng, d, k, _dtype_id(q), metric, int(want_rank)), dev)
        B.check(lib.sbir_pairwise_topk(q.data_ptr(), nq, g.data_ptr(), _ptr(gsq), ng, d, _dtype_id(q), metric, k,
                                       int(index_offset), _ptr(pos), vals.data_ptr(), idx.data_ptr(), _ptr(rank),
                                       unc.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "sbir_pairwise_topk")
    out = (vals, idx) + ((rank,) if want_rank else ())
    return out + ((unc,) if return_uncertified else ())


def rank_of_positive(queries: torch.Tensor, gallery: torch.Tensor, pos_index: torch.Tensor,
                     loss_type: str = "euclidean") -> torch.Tensor:
    """0-based rank of gallery row pos_index[q] in the ascending distance order of query q
    (count of strictly closer rows) — what get_ranking_position finds with a full sort
    (inference.py:49-52); len(gallery) where pos_index[q] < 0 (inference.py:39-41)."""
    return pairwise_topk(queries, gallery, 1, loss_type, pos_index=pos_index)[2]


def positive_distance(queries, gallery_shard, pos_index_local, loss_type="euclidean") -> torch.Tensor:
    """Exact d(q_i, shard[pos_index_local[i]]) as fp64 [Q]; NaN where the index is outside the shard."""
    pos = _dev(pos_index_local.to(torch.int64), "pos_index")
    q, g = _check_retrieval_args(queries, gallery_shard, None, (("pos_index_local", pos),))
    out = torch.empty(q.shape[0], dtype=torch.float64, device=q.device)
    with torch.cuda.device(q.device):
        B.check(B.load().sbir_positive_distance(q.data_ptr(), q.shape[0], g.data_ptr(), g.shape[0], q.shape[1],
                                                _dtype_id(q), metric_id(loss_type), pos.data_ptr(), out.data_ptr(),
                                                _stream()), "sbir_positive_distance")
    return out


def pairwise_topk_shard(queries, gallery_shard, k, loss_type, index_offset, pos_dist=None, pos_index_global=None,
                        gallery_sqnorm=None):
    """One gallery shard's contribution: local top-k with global indices and, if pos_dist
    (fp64 [Q], NaN = no positive) is given, the local count of rows closer than it."""
    want = pos_dist is not None
    pd = _dev(pos_dist.to(torch.float64), "pos_dist") if want else None
    pg = _dev(pos_index_global.to(torch.int64), "pos_index_global") if (want and pos_index_global is not None) else None
    q, g = _check_retrieval_args(queries, gallery_shard, k, (("pos_dist", pd), ("pos_index_global", pg)))
    gsq = _gallery_sqnorm(gallery_sqnorm, g, gallery_shard)
    metric = metric_id(loss_type)
    nq, ng, d = q.shape[0], g.shape[0], q.shape[1]
    dev = q.device
    vals = torch.empty((nq, k), dtype=torch.float32, device=dev)
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    cnt = torch.zeros(nq, dtype=torch.int64, device=dev) if want else None
    unc = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = B.load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.sbir_pairwise_topk_workspace_bytes(nq, ng, d, k, _dtype_id(q), metric, int(want)), dev)
        B.check(lib.sbir_pairwise_topk_shard(q.data_ptr(), nq, _ptr(g) if ng else None, _ptr(gsq), ng, d, _dtype_id(q), metric,
                                             k, int(index_offset), _ptr(pd), _ptr(pg), vals.data_ptr(), idx.data_ptr(),
                                             _ptr(cnt), unc.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                "sbir_pairwise_topk_shard")
    return vals, idx, cnt, unc


def topk_merge(dist: torch.Tensor, index: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge [L, Q, k] ascending lists into the k best per query (K4).  The L lists may be views into a
    larger buffer (e.g. the packed all-gather messages of sharded._exchange): only each [Q, k] block must be
    dense; the list stride is passed to the kernel, nothing is copied."""
    if dist.dtype != torch.float32 or index.dtype != torch.int64:
        dist, index = dist.float(), index.to(torch.int64)
    if not dist.is_cuda or not index.is_cuda:
        raise RuntimeError("topk_merge takes CUDA tensors: the sbir_b200 path has no CPU fallback")
    L, nq, k = dist.shape

    def dense_blocks(t):
        return t.stride(2) == 1 and t.stride(1) == k and (L <= 1 or t.stride(0) >= nq * k)
    if not dense_blocks(dist):
        dist = dist.contiguous()
    if not dense_blocks(index):
        index = index.contiguous()
    sd = dist.stride(0) if L > 1 else 0
    si = index.stride(0) if L > 1 else 0
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dist.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dist.device)
    with torch.cuda.device(dist.device):
        B.check(B.load().sbir_topk_merge(dist.data_ptr(), index.data_ptr(), L, sd, si, nq, k, out_d.data_ptr(),
                                         out_i.data_ptr(), _stream()), "sbir_topk_merge")
    return out_d, out_i


# --------------------------------------------------------------------------- H5 ----
def retrieval_metrics(rank0: torch.Tensor, k: int = 10) -> dict:
    """The reference's metrics dict (inference.py:95-98,113-134) from 0-based ranks:
    mean_reciprocal_rank, topk_acc (cumulative, K=1..k) and the pandas describe() fields
    of the 1-based rank (count, mean, std, min, 25%, 50%, 75%, max).  One D2H at the end."""
    r = _dev(rank0.to(torch.int64), "rank0")
    n = r.numel()
    out = torch.empty(k + 5, dtype=torch.float64, device=r.device)
    with torch.cuda.device(r.device):
        B.check(B.load().sbir_retrieval_metrics(r.data_ptr(), n, k, out.data_ptr(), _stream()),
                "sbir_retrieval_metrics")
    # quartiles (pandas' linear interpolation) need the order statistics: device sort, 6 scalars back
    srt = torch.sort(r + 1).values.to(torch.float64)
    qs = []
    for frac in (0.25, 0.5, 0.75):
        pos = frac * (n - 1)
        lo = int(pos)
        hi = min(lo + 1, n - 1)
        qs.append(srt[lo] + (srt[hi] - srt[lo]) * (pos - lo))
    host = torch.cat([out, torch.stack(qs)]).cpu().tolist()
    return {
        "mean_reciprocal_rank": host[0],
        "count": float(n),
        "mean": host[k + 1],
        "std": host[k + 2],
        "min": host[k + 3],
        "25%": host[k + 5],
        "50%": host[k + 6],
        "75%": host[k + 7],
        "max": host[k + 4],
        "topk_acc": host[1:k + 1],
    }


# ---------------------------------------------------------------------- H6 / H7 ----
class _TripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, p, n, margin, metric):
        a, p, n = (_dev(t.float(), "triplet input") for t in (a, p, n))
        if not (a.shape == p.shape == n.shape) or a.dim() != 2:
            raise ValueError("anchor, positive and negative must share a [batch, dim] shape")
        bsz, d = a.shape
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        per_row = torch.empty(bsz, dtype=torch.float32, device=a.device)
        need = [ctx.needs_input_grad[i] for i in range(3)]
        grads = [torch.empty_like(a) if nd else None for nd in need]
        with torch.cuda.device(a.device):
            B.check(B.load().sbir_triplet_margin_loss(a.data_ptr(), p.data_ptr(), n.data_ptr(), bsz, d, float(margin),
                                                      metric, loss.data_ptr(), per_row.data_ptr(), _ptr(grads[0]),
                                                      _ptr(grads[1]), _ptr(grads[2]), _stream()),
                    "sbir_triplet_margin_loss")
        ctx.save_for_backward(*[g for g in grads if g is not None])
        ctx.need = need
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        saved = list(ctx.saved_tensors)
        out = []
        for nd in ctx.need:
            out.append(saved.pop(0) * grad_loss if nd else None)
        return (*out, None, None)


def triplet_margin_loss(anchor, positive, negative, margin: float = 0.2, loss_type: str = "euclidean"):
    """mean_i max(0, margin + d(a_i,p_i) − d(a_i,n_i)) — nn.TripletMarginLoss (train.py:169) and
    nn.TripletMarginWithDistanceLoss (utils.py:56,69); forward and backward in one launch."""
    return _TripletFn.apply(anchor, positive, negative, margin, metric_id(loss_type))


# --------------------------------------------------------------------------- H8 ----
class _BatchHardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, p, n, margin, metric, anchor_label, cand_label):
        a, p, n = (_dev(t.float(), "triplet input") for t in (a, p, n))
        bsz, d = a.shape
        dev = a.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        hard = torch.empty((bsz, 2), dtype=torch.int64, device=dev)
        need = [ctx.needs_input_grad[i] for i in range(3)]
        grads = [torch.empty_like(a) if nd else None for nd in need]
        al = _dev(anchor_label.to(torch.int64), "anchor_label") if anchor_label is not None else None
        cl = _dev(cand_label.to(torch.int64), "cand_label") if cand_label is not None else None
        lib = B.load()
        with torch.cuda.device(dev):
            ws = _workspace(lib.sbir_batch_hard_workspace_bytes(bsz, d), dev)
            B.check(lib.sbir_batch_hard_triplet_loss(a.data_ptr(), p.data_ptr(), n.data_ptr(), bsz, d, float(margin),
                                                     metric, _ptr(al), _ptr(cl), loss.data_ptr(), hard.data_ptr(),
                                                     _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2]), ws.data_ptr(),
                                                     ws.numel(), _stream()), "sbir_batch_hard_triplet_loss")
        ctx.save_for_backward(*[g for g in grads if g is not None])
        ctx.need = need
        ctx.mark_non_differentiable(hard)
        return loss, hard

    @staticmethod
    def backward(ctx, grad_loss, _grad_hard):
        saved = list(ctx.saved_tensors)
        out = []
        for nd in ctx.need:
            out.append(saved.pop(0) * grad_loss if nd else None)
        return (*out, None, None, None, None)


def batch_hard_triplet_loss(anchor, positive, negative, margin: float = 0.2, loss_type: str = "euclidean",
                            labels: Optional[torch.Tensor] = None, return_indices: bool = False):
    """Batch-hard mining over X = cat(positive, negative): hardest positive
    (X_i, or every candidate sharing labels[i]) and hardest negative per anchor, margin loss,
    gradients through the selected pairs.  `labels` is int64 [batch]: the class of triplet i
    (its anchor and positive); negatives carry the label of their own row."""
    al = cl = None
    if labels is not None:
        al = labels
        # candidates: positives carry their triplet's label; a negative row j is a negative
        # for everyone whose label differs from the class it was drawn for (V2 datasets draw
        # negatives from the SAME class, data_preparation.py:214-222 — then it is only ever
        # a positive-set member if labels say so).  Callers pass negative labels explicitly
        # by concatenating; default: -1 - index (never equal to an anchor label).
        neg_lab = -1 - torch.arange(labels.numel(), device=labels.device, dtype=torch.int64)
        cl = torch.cat([labels.to(torch.int64), neg_lab])
    loss, hard = _BatchHardFn.apply(anchor, positive, negative, margin, metric_id(loss_type), al, cl)
    return (loss, hard) if return_indices else loss


# ------------------------------------------------------------------------ debug ----
def debug_dist_matrix(queries, gallery, loss_type="euclidean") -> torch.Tensor:
    """Raw tensor-core epilogue matrix (‖g‖²−2qg or −q·g/max(‖g‖,eps)); test-only."""
    q, g = _dev(queries, "queries"), _dev(gallery, "gallery")
    nq, ng, d = q.shape[0], g.shape[0], q.shape[1]
    out = torch.full((nq, ng), float("nan"), dtype=torch.float32, device=q.device)
    lib = B.load()
    with torch.cuda.device(q.device):
        ws = _workspace(lib.sbir_debug_dist_matrix_workspace_bytes(nq, ng, d, _dtype_id(q)), q.device)
        B.check(lib.sbir_debug_dist_matrix(q.data_ptr(), nq, g.data_ptr(), ng, d, _dtype_id(q), metric_id(loss_type),
                                           out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                "sbir_debug_dist_matrix")
    return out
