"""GPU suite, multi-rank part (-m gpu; skipped on a box with fewer than 2 GPUs): the CUDA kernels and real NCCL
together — `sharded_pairwise_topk` / `sharded_retrieve_host` over a 2-rank (and, when the box has them, 4-rank)
process group must return exactly what one GPU returns for the whole gallery (top-k values, indices, ranks),
on every rank.  The gloo tests (tests/test_sharded_gloo.py) cover the same plumbing on CPU with the oracle as scorer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sbir_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(dtype):
    Q, G, pos = O.synthetic_embeddings(700, 30011, 256, seed=33, beta=0.2)
    pos[::13] = -1
    return Q.to(dtype), G.to(dtype), pos


def _worker(rank, world, port, dtype_name, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from art_sbir_b200 import ops, sharded
        dtype = getattr(torch, dtype_name)
        Q, G, pos = _case(dtype)
        a, b = sharded.shard_bounds(G.shape[0], world, rank)
        q, shard, p = Q.to(dev), G[a:b].contiguous().to(dev), pos.to(dev)
        out = {}
        for lt, k in (("euclidean", 10), ("cosine", 40)):
            v, i, r = sharded.sharded_pairwise_topk(q, shard, k, lt, pos_index=p)          # offsets derived by all-reduce
            out[(lt, k, "device")] = (v.cpu(), i.cpu(), r.cpu())
        v, i, r = sharded.sharded_retrieve_host(Q.pin_memory(), G[a:b].contiguous().pin_memory(), 10, "euclidean", pos_index=pos,
                                                device=dev)                                    # shard streamed from host memory
        out[("euclidean", 10, "host")] = (v.cpu(), i.cpu(), r.cpu())
        if rank == 0:
            g = G.to(dev)
            for lt, k in (("euclidean", 10), ("cosine", 40)):
                v, i, r = ops.pairwise_topk(q, g, k, lt, pos_index=p)
                out[(lt, k, "single")] = (v.cpu(), i.cpu(), r.cpu())
        ret[rank] = out
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("dtype_name", ["bfloat16", "float32"])
def test_nccl_sharded_equals_single_gpu(world, dtype_name):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from art_sbir_b200 import _build
    _build.build()                      # once, before the ranks race for it
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), dtype_name, ret), nprocs=world, join=True)
    single = ret[0]
    Q, G, pos = _case(getattr(torch, dtype_name))
    for lt, k in (("euclidean", 10), ("cosine", 40)):
        want_v, want_i, want_r = single[(lt, k, "single")]
        assert (want_r[pos < 0] == G.shape[0]).all()
        for r in range(world):
            v, i, rk = ret[r][(lt, k, "device")]
            assert torch.equal(i, want_i) and torch.equal(v, want_v) and torch.equal(rk, want_r), (lt, k, r)
    want_v, want_i, want_r = single[("euclidean", 10, "single")]
    for r in range(world):
        v, i, rk = ret[r][("euclidean", 10, "host")]
        assert torch.equal(i, want_i) and torch.equal(v, want_v) and torch.equal(rk, want_r), ("host", r)
    # and the single-GPU answer is the oracle's (fp32 math on the inputs as given)
    ref_v, ref_i = O.pairwise_topk_batched(Q[:64].float(), G.float(), 10, "euclidean")
    assert torch.equal(want_i[:64], ref_i) and torch.allclose(want_v[:64], ref_v, rtol=1e-3)
