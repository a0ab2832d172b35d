"""CPU suite, part 4: the multi-GPU (gallery-sharded) path's collective plumbing with
world_size 2 and 3 over gloo.  The per-shard scorer is the oracle here (tests may use it as a
stand-in checker); on the GPU box tests/test_gpu_parity.py runs the same function with the CUDA
kernels.  The sharded result must equal the unsharded oracle exactly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sbir_oracle as O


def _oracle_local(queries, shard, k, loss_type, offset, pos_dist, pos_index_global=None):
    n = shard.shape[0]
    kk = min(k, n)
    vals = torch.full((queries.shape[0], k), float("inf"))
    idx = torch.full((queries.shape[0], k), -1, dtype=torch.int64)
    cnt = torch.zeros(queries.shape[0], dtype=torch.int64) if pos_dist is not None else None
    for i in range(queries.shape[0]):
        if n == 0:
            continue
        d = O.distances(queries[i:i + 1], shard, loss_type)
        v, ix = d.topk(kk, largest=False)
        vals[i, :kk], idx[i, :kk] = v, ix + offset
        if pos_dist is not None and pos_dist[i] == pos_dist[i]:
            cnt[i] = int((d.double() < pos_dist[i]).sum())
    return vals, idx, cnt


def _oracle_pos_dist(queries, shard, pos_local, loss_type):
    out = torch.full((queries.shape[0],), float("nan"), dtype=torch.float64)
    for i in range(queries.shape[0]):
        if pos_local[i] >= 0:
            out[i] = O.distances(queries[i:i + 1], shard[pos_local[i]:pos_local[i] + 1], loss_type)[0].double()
    return out


def _cpu_merge(dist_lists, idx_lists):
    L, Q, k = dist_lists.shape
    d = dist_lists.permute(1, 0, 2).reshape(Q, L * k)
    ix = idx_lists.permute(1, 0, 2).reshape(Q, L * k)
    ix_sort = torch.where(ix < 0, torch.full_like(ix, 2 ** 62), ix)
    order = torch.argsort(ix_sort, dim=1, stable=True)           # ties by index ...
    d, ix = d.gather(1, order), ix.gather(1, order)
    order = torch.argsort(d, dim=1, stable=True)                 # ... within ascending distance
    return d.gather(1, order)[:, :k], ix.gather(1, order)[:, :k]


def _worker(rank, world, port, loss_type, ret, weighted=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from art_sbir_b200 import sharded
        Q, G, pos = O.synthetic_embeddings(12, 101, 32, seed=5, beta=0.3, num_classes=4)
        pos[3] = -1
        if weighted:
            # speed-weighted shards: every rank reports (rows, busy ms) of a calibration pass, all ranks derive the same cuts
            speeds = sharded.rank_speed_weights(100, 1.0 + 0.5 * rank)
            assert speeds == [100.0 / (1.0 + 0.5 * r) for r in range(world)]
            a, b = sharded.weighted_shard_bounds(G.shape[0], speeds, rank)
            ret[f"span{rank}"] = (a, b)
        else:
            a, b = sharded.shard_bounds(G.shape[0], world, rank)
        vals, idx, rk = sharded.sharded_pairwise_topk(Q, G[a:b], 5, loss_type, pos_index=pos, local_fn=_oracle_local,
                                                      pos_dist_fn=_oracle_pos_dist, merge_fn=_cpu_merge)
        ret[rank] = (vals, idx, rk)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("loss_type", ["euclidean", "cosine"])
def test_sharded_equals_unsharded(world, loss_type):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), loss_type, ret), nprocs=world, join=True)
    Q, G, pos = O.synthetic_embeddings(12, 101, 32, seed=5, beta=0.3, num_classes=4)
    pos[3] = -1
    want_v, want_i = O.pairwise_topk_batched(Q, G, 5, loss_type)
    want_r = O.rank_of_positive_batched(Q, G, pos, loss_type)
    for r in range(world):
        vals, idx, rk = ret[r]
        assert torch.equal(idx, want_i)
        assert torch.allclose(vals, want_v, rtol=1e-6)
        assert torch.equal(rk, want_r)
        assert rk[3].item() == G.shape[0]


def test_speed_weighted_shards_give_the_same_result():
    """Uneven shards cut from all-gathered per-rank speeds (what bench.py --gpus N does before timing):
    faster ranks hold more rows, the spans tile the gallery, and the result is the unsharded one."""
    world = 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), "euclidean", ret, True), nprocs=world, join=True)
    spans = [ret[f"span{r}"] for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == 101 and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in spans]
    assert sizes[0] > sizes[1] > sizes[2] > 0
    Q, G, pos = O.synthetic_embeddings(12, 101, 32, seed=5, beta=0.3, num_classes=4)
    pos[3] = -1
    want_v, want_i = O.pairwise_topk_batched(Q, G, 5, "euclidean")
    want_r = O.rank_of_positive_batched(Q, G, pos, "euclidean")
    for r in range(world):
        vals, idx, rk = ret[r]
        assert torch.equal(idx, want_i) and torch.equal(rk, want_r)
        assert torch.allclose(vals, want_v, rtol=1e-6)
