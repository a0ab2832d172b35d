"""Where a gallery-sharded step spends its time outside the distance kernel (multi-GPU).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/shard_breakdown.py [workload]

Re-states `sharded.sharded_pairwise_topk` segment by segment with CUDA events between the segments
(positives -> all-reduce -> local pass -> pack -> all-gather -> merge) and prints, per rank, the mean of each
segment over a few steps.  On the slowest rank the all-gather segment is the collective itself; on the others
it also holds the wait for the slowest shard.  A measuring tool: not part of the tests or of bench.py.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from art_sbir_b200 import ops, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    num_q, num_g, dim, dtype_name, k, _ = bench.WORKLOADS[name]
    dtype = getattr(torch, dtype_name)
    r0, r1 = sharded.shard_bounds(num_g, world, rank)
    Q, Gs, pos = bench.make_shard(num_q, num_g, dim, dtype, r0, r1, dev)
    n_local = r1 - r0
    names = ["positives", "all_reduce", "local_pass", "pack", "all_gather", "merge"]

    def step(ev):
        ev[0].record()
        mine = (pos >= r0) & (pos < r0 + n_local)
        pos_local = torch.where(mine, pos - r0, torch.full_like(pos, -1))
        d_local = ops.positive_distance(Q, Gs, pos_local, "euclidean").to(torch.float64)
        pair = torch.stack([torch.where(mine, d_local, torch.zeros_like(d_local)), mine.to(torch.float64)])
        ev[1].record()
        if world > 1:
            dist.all_reduce(pair)
        pos_dist = torch.where(pair[1] > 0, pair[0], torch.full_like(pair[0], float("nan")))
        ev[2].record()
        vals, idx, cnt, _ = ops.pairwise_topk_shard(Q, Gs, k, "euclidean", r0, pos_dist, pos)
        ev[3].record()
        nq = Q.shape[0]
        nbytes = nq * k * 12 + nq * 8
        msg = torch.empty((nbytes + 15) // 16 * 16, dtype=torch.uint8, device=dev)
        o_cnt, o_val = nq * k * 8, nq * k * 8 + nq * 8
        msg[:o_cnt].view(torch.int64).copy_(idx.reshape(-1))
        msg[o_cnt:o_val].view(torch.int64).copy_(cnt)
        msg[o_val:o_val + nq * k * 4].view(torch.float32).copy_(vals.reshape(-1))
        gathered = torch.empty((world, msg.numel()), dtype=torch.uint8, device=dev)
        ev[4].record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, msg)
        else:
            gathered.copy_(msg)
        ev[5].record()
        all_idx = gathered[:, :o_cnt].view(torch.int64).unflatten(1, (nq, k))
        all_vals = gathered[:, o_val:o_val + nq * k * 4].view(torch.float32).unflatten(1, (nq, k))
        v, i = ops.topk_merge(all_vals, all_idx)
        c = gathered[:, o_cnt:o_val].view(torch.int64).sum(dim=0)
        r = torch.where(pos_dist != pos_dist, torch.full_like(c, num_g), c)
        ev[6].record()
        return v, i, r

    def events():
        return [torch.cuda.Event(enable_timing=True) for _ in range(7)]

    for _ in range(3):
        step(events())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = 5
    evs = [events() for _ in range(steps)]
    for ev in evs:
        step(ev)
    torch.cuda.synchronize()
    seg = [sum(ev[j].elapsed_time(ev[j + 1]) for ev in evs) / steps for j in range(6)]
    total = sum(evs[s][0].elapsed_time(evs[s][6]) for s in range(steps)) / steps
    mine = torch.tensor(seg + [total], dtype=torch.float64, device=dev)
    every = [torch.empty_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(every, mine)
    else:
        every = [mine]
    if rank == 0:
        out = {"workload": name, "world": world, "segments_ms_per_rank": {}}
        for j, n in enumerate(names + ["step"]):
            out["segments_ms_per_rank"][n] = [round(t[j].item(), 3) for t in every]
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
