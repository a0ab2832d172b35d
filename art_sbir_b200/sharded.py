"""Gallery-sharded retrieval across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r holds the contiguous
gallery rows [offset_r, offset_r + n_r); queries are replicated.  The path has exactly one
exchange step:

    d(q, pos)      owner shard computes it, all-reduce(sum) of (value, has) pairs  [Q] fp64
    local K1       per-shard top-k with global indices + local count(d < d_pos)
    all-gather     ONE packed message per rank: [Q,k] dist + [Q,k] index + [Q] count — Q·(12k + 8) bytes
    K4 merge       k best of the P·k gathered candidates, ties by global index; counts summed locally

so the result is identical to the single-GPU result for any number of shards.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def shard_bounds(num_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`: the first (num_rows % world_size) ranks get one extra row."""
    base, rem = divmod(num_rows, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def weighted_shard_bounds(num_rows: int, weights, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous row range of `rank` when rank r gets a share of the gallery proportional to weights[r]
    (its measured scoring speed, see `rank_speed_weights`).  The GPUs of one chassis do not run a power-capped
    tensor kernel at the same clock, and the exchange step waits for the slowest shard, so equal shards leave the
    faster GPUs idle; cuts are rounded to `align` rows, cover [0, num_rows) exactly, and equal weights reproduce an
    even split.  Any partition gives the same result (module docstring), so this is purely a speed choice."""
    w = [max(float(x), 0.0) for x in weights]
    total = sum(w)
    if not w or total <= 0.0:
        raise ValueError("weights must hold one non-negative number per rank, not all zero")
    if not 0 <= rank < len(w):
        raise ValueError(f"rank {rank} outside the {len(w)} weights")
    align = max(1, int(align))
    cuts, acc = [0], 0.0
    for r in range(len(w) - 1):
        acc += w[r]
        cut = int(round(num_rows * acc / total / align)) * align
        cuts.append(min(max(cut, cuts[-1]), num_rows))
    cuts.append(num_rows)
    return cuts[rank], cuts[rank + 1]


def rebalanced_weights(rows, busy_ms, damping: float = 0.6):
    """Next shard weights from the current cut and the time every rank's GPU spent scoring it: each shard moves
    `damping` of the way (in log space) towards the size that would equalise the kernel times.  Damped because the
    measurement is biased — a GPU that finishes early idles until the exchange step and, under a power cap, clocks
    higher than it can sustain at full duty — so re-measure and repeat; 1.0 is the full correction."""
    rows = [float(r) for r in rows]
    ms = [max(float(t), 1e-9) for t in busy_ms]
    if len(rows) != len(ms) or not rows:
        raise ValueError("one row count and one time per rank")
    mean_ms = sum(ms) / len(ms)
    return [r * (mean_ms / t) ** float(damping) for r, t in zip(rows, ms)]


def rank_speed_weights(rows_local: int, busy_ms_local: float, device=None, group=None):
    """Gallery rows per millisecond of every rank (all-gathered, the same list everywhere): `busy_ms_local` is the
    time this rank's GPU spent scoring `rows_local` rows with nobody to wait for — the distance kernel's own time,
    not the step time, which the exchange step equalises across ranks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = torch.tensor([float(rows_local) / max(float(busy_ms_local), 1e-9)], dtype=torch.float64,
                        device=device if device is not None else "cpu")
    if world == 1:
        return [float(mine.item())]
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [float(t.item()) for t in out]


def _resolve_offsets(n_local: int, shard_offset: Optional[int], num_gallery_total: Optional[int], world: int,
                     rank: int, dev, group) -> Tuple[int, int]:
    """First global row of this rank's shard and the gallery length.  Whatever the caller left out is
    derived from an all-reduce of the shard sizes (prefix sum in rank order), so a missing
    `shard_offset` can never silently mean 0 on every rank; what it passed is range-checked."""
    if shard_offset is None or num_gallery_total is None:
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = n_local
        if world > 1:
            dist.all_reduce(sizes, group=group)
        sizes = sizes.cpu()
        if shard_offset is None:
            shard_offset = int(sizes[:rank].sum())
        if num_gallery_total is None:
            num_gallery_total = int(sizes.sum())
    if shard_offset < 0 or shard_offset + n_local > num_gallery_total:
        raise ValueError(f"shard rows [{shard_offset}, {shard_offset + n_local}) fall outside the gallery of "
                         f"{num_gallery_total} rows")
    return int(shard_offset), int(num_gallery_total)


def _cuda_local(queries, shard, k, loss_type, offset, pos_dist, pos_index_global=None):
    vals, idx, cnt, unc = ops.pairwise_topk_shard(queries, shard, k, loss_type, offset, pos_dist, pos_index_global)
    return vals, idx, cnt


def _cuda_pos_dist(queries, shard, pos_local, loss_type):
    return ops.positive_distance(queries, shard, pos_local, loss_type)


def sharded_pairwise_topk(queries: torch.Tensor, gallery_shard: torch.Tensor, k: int, loss_type: str = "euclidean",
                          pos_index: Optional[torch.Tensor] = None, shard_offset: Optional[int] = None,
                          num_gallery_total: Optional[int] = None, group=None,
                          local_fn: Callable = _cuda_local, pos_dist_fn: Callable = _cuda_pos_dist,
                          merge_fn: Callable = ops.topk_merge):
    """Top-k (and rank of the positive) of `queries` against a gallery sharded over `group`.

    pos_index holds GLOBAL gallery rows (<0 = no positive).  Returns (values [Q,k] fp32,
    indices [Q,k] int64 global, rank int64 [Q] or None) — the same on every rank.
    `local_fn` / `pos_dist_fn` / `merge_fn` exist so the collective plumbing can be exercised
    on CPU (gloo) with a stand-in scorer in tests; the defaults are the CUDA kernels."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_local = gallery_shard.shape[0]
    dev = queries.device
    shard_offset, num_gallery_total = _resolve_offsets(n_local, shard_offset, num_gallery_total, world, rank, dev, group)

    pos_dist = None
    if pos_index is not None:
        pos_index = pos_index.to(device=dev, dtype=torch.int64)
        mine = (pos_index >= shard_offset) & (pos_index < shard_offset + n_local)
        pos_local = torch.where(mine, pos_index - shard_offset, torch.full_like(pos_index, -1))
        d_local = pos_dist_fn(queries, gallery_shard, pos_local, loss_type).to(torch.float64)
        pair = torch.stack([torch.where(mine, d_local, torch.zeros_like(d_local)), mine.to(torch.float64)])
        if world > 1:
            dist.all_reduce(pair, group=group)  # exactly one owner per query: the sum is exact
        pos_dist = torch.where(pair[1] > 0, pair[0], torch.full_like(pair[0], float("nan")))

    vals, idx, cnt = local_fn(queries, gallery_shard, k, loss_type, shard_offset, pos_dist, pos_index)
    return _exchange(vals, idx, cnt, pos_dist, num_gallery_total, world, group, merge_fn)


def _exchange(vals, idx, cnt, pos_dist, num_gallery_total, world, group, merge_fn):
    """The path's one exchange step.  Every rank contributes ONE packed message — its [Q,k] distances (fp32),
    [Q,k] global indices (int64) and [Q] local counts (int64), Q·(12k + 8) bytes — to ONE all-gather; the
    gathered buffer is read in place by K4 (views of the [world, Q, k] blocks) and the counts are summed
    locally.  (Round 1 issued two all-gathers and an all-reduce here; with ≤ 13 MB per rank the step is
    latency-bound, so the number of collectives is what matters.)"""
    if world > 1:
        nq, k = vals.shape
        dev = vals.device
        has_cnt = cnt is not None
        nbytes = nq * k * 12 + (nq * 8 if has_cnt else 0)
        nbytes_pad = (nbytes + 15) // 16 * 16
        # int64 indices first (8-byte aligned views), then the counts, then the fp32 distances
        mine = torch.empty(nbytes_pad, dtype=torch.uint8, device=dev)
        o_idx, o_cnt, o_val = 0, nq * k * 8, nq * k * 8 + (nq * 8 if has_cnt else 0)
        mine[o_idx:o_idx + nq * k * 8].view(torch.int64).copy_(idx.reshape(-1))
        if has_cnt:
            mine[o_cnt:o_cnt + nq * 8].view(torch.int64).copy_(cnt)
        mine[o_val:o_val + nq * k * 4].view(torch.float32).copy_(vals.reshape(-1))
        gathered = torch.empty((world, nbytes_pad), dtype=torch.uint8, device=dev)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(gathered, mine, group=group)
        else:  # gloo (CPU tests of the plumbing): list form
            dist.all_gather(list(gathered.unbind(0)), mine, group=group)
        # strided [world, Q, k] views of the gathered messages (list stride = message size): K4 reads them in place
        all_idx = gathered[:, o_idx:o_idx + nq * k * 8].view(torch.int64).unflatten(1, (nq, k))
        all_vals = gathered[:, o_val:o_val + nq * k * 4].view(torch.float32).unflatten(1, (nq, k))
        vals, idx = merge_fn(all_vals, all_idx)
        if has_cnt:
            cnt = gathered[:, o_cnt:o_cnt + nq * 8].view(torch.int64).sum(dim=0)
    rank_out = None
    if pos_dist is not None:
        rank_out = torch.where(pos_dist != pos_dist, torch.full_like(cnt, num_gallery_total), cnt)
    return vals, idx, rank_out


_pinned_cache = {}


def _pinned_rows(rows: int, dim: int, dtype: torch.dtype) -> torch.Tensor:
    """Pinned host staging for gathered rows, kept between calls (pinning 100 MB costs more than the gather)."""
    key = (dim, dtype)
    buf = _pinned_cache.get(key)
    if buf is None or buf.shape[0] < rows:
        buf = torch.empty((rows, dim), dtype=dtype).pin_memory()
        _pinned_cache[key] = buf
    return buf


def _upload_replicated(x_host: torch.Tensor, dev, world: int, rank: int, group) -> torch.Tensor:
    """A host matrix every rank holds (the queries) → the device.  With several ranks on one box the ranks share the
    host's memory and PCIe bandwidth (8 ranks copying at once: ≈23 GB/s each), so uploading the same 100 MB eight times
    puts ≈4 ms in front of the first scoring launch; instead every rank uploads ITS 1/world of the rows and the slices
    are all-gathered over NVLink."""
    n = x_host.shape[0]
    if world == 1 or n < 4096 or dist.get_backend(group) != "nccl":
        return x_host.to(dev, non_blocking=True).contiguous()
    per = (n + world - 1) // world
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    mine = torch.zeros((per,) + tuple(x_host.shape[1:]), dtype=x_host.dtype, device=dev)
    if hi > lo:
        mine[:hi - lo].copy_(x_host[lo:hi], non_blocking=True)
    full = torch.empty((world * per,) + tuple(x_host.shape[1:]), dtype=x_host.dtype, device=dev)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n]


def sharded_retrieve_host(queries_host: torch.Tensor, gallery_shard_host: torch.Tensor, k: int,
                          loss_type: str = "euclidean", pos_index: Optional[torch.Tensor] = None,
                          shard_offset: Optional[int] = None, num_gallery_total: Optional[int] = None, group=None,
                          device: Optional[torch.device] = None):
    """`sharded_pairwise_topk` for embeddings that live in HOST memory (pinned for full PCIe speed):
    the rank's gallery shard is uploaded in chunks while earlier chunks are scored
    (sbir_retrieve_host_shard), the positives' rows are gathered on the host by their owner rank.
    Returns device tensors (values, global indices, rank or None), the same on every rank."""
    from . import _binding as B
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if queries_host.dim() != 2 or gallery_shard_host.dim() != 2 or queries_host.shape[1] != gallery_shard_host.shape[1]:
        raise ValueError(f"expected [Q,D] and [N,D], got {tuple(queries_host.shape)} and {tuple(gallery_shard_host.shape)}")
    if not 1 <= k <= B.MAX_K:
        raise ValueError(f"k must be in [1, {B.MAX_K}]")
    if pos_index is not None and (pos_index.dim() != 1 or pos_index.shape[0] != queries_host.shape[0]):
        raise ValueError("pos_index must have one entry per query")
    n_local = gallery_shard_host.shape[0]
    shard_offset, num_gallery_total = _resolve_offsets(n_local, shard_offset, num_gallery_total, world, rank, dev, group)
    if queries_host.dtype != gallery_shard_host.dtype or queries_host.dtype not in (torch.float32, torch.bfloat16):
        queries_host, gallery_shard_host = queries_host.float(), gallery_shard_host.float()
    g_host = gallery_shard_host.contiguous()
    q = _upload_replicated(queries_host, dev, world, rank, group)
    nq, d = q.shape
    pos_dist = pos_g = None
    if pos_index is not None:
        pos_cpu = pos_index.to("cpu", torch.int64).contiguous()
        # rows of the positives this rank owns, gathered from the host shard by a few C threads into pinned memory
        # (a zero row for every query whose positive lives elsewhere), then one upload
        local = pos_cpu - shard_offset
        mine = (local >= 0) & (local < n_local)
        own = local[mine].contiguous()                         # shard rows of the positives this rank owns
        n_own = own.numel()
        rows_host = _pinned_rows(max(n_own, 1), d, g_host.dtype)
        if n_own:
            B.check(B.load().sbir_gather_rows_host(g_host.data_ptr(), n_local, d * g_host.element_size(), own.data_ptr(), n_own,
                                                   rows_host.data_ptr(), 0), "sbir_gather_rows_host")
        rows = rows_host[:max(n_own, 1)].to(dev, non_blocking=True)
        ident = torch.where(mine, torch.cumsum(mine, 0) - 1, torch.full((nq,), -1, dtype=torch.int64)).to(dev)
        mine_d = mine.to(dev)
        d_local = ops.positive_distance(q, rows, ident, loss_type)
        pair = torch.stack([torch.where(mine_d, d_local, torch.zeros_like(d_local)), mine_d.to(torch.float64)])
        if world > 1:
            dist.all_reduce(pair, group=group)
        pos_dist = torch.where(pair[1] > 0, pair[0], torch.full_like(pair[0], float("nan"))).contiguous()
        pos_g = pos_cpu.to(dev)
    vals = torch.empty((nq, k), dtype=torch.float32, device=dev)
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    cnt = torch.zeros(nq, dtype=torch.int64, device=dev) if pos_dist is not None else None
    import ctypes
    unc = ctypes.c_int32(0)
    with torch.cuda.device(dev):
        B.check(B.load().sbir_retrieve_host_shard(
            q.data_ptr(), nq, g_host.data_ptr() if n_local else None, n_local, d, ops._dtype_id(q), ops.metric_id(loss_type), k,
            int(shard_offset), None if pos_dist is None else pos_dist.data_ptr(), None if pos_g is None else pos_g.data_ptr(),
            vals.data_ptr(), idx.data_ptr(), None if cnt is None else cnt.data_ptr(), ctypes.byref(unc),
            torch.cuda.current_stream().cuda_stream), "sbir_retrieve_host_shard")
    return _exchange(vals, idx, cnt, pos_dist, num_gallery_total, world, group, ops.topk_merge)
