"""BASELINE config 5 — end-to-end evaluation: a PyTorch image encoder on synthetic 224x224 sketches
and artworks feeding the gallery-sharded distance / top-K / rank path.

    python tools/e2e_cfg5.py [--gallery 20000] [--queries 2000] [--batch 256]
    torchrun --nproc-per-node 8 tools/e2e_cfg5.py --gallery 200000 --queries 20000

The encoder is OUT OF SCOPE of this repo (it stays PyTorch); the reference's
ModifiedResNet(3,4,6,3; output_dim=1024) lives in its models.py and is not redistributed here, so a
stand-in with the same interface is used: a random-init ResNet-50-shaped torchvision model whose
head emits 1024-d embeddings (torchvision is a dependency of the reference too).  Each rank encodes
its slice of the gallery images straight into its shard of the [N, 1024] feature matrix
(DP over images yields the row-sharded layout for free), every rank encodes all
queries, then `sharded_pairwise_topk` scores them.  Reports encoder images/s and retrieval pairs/s
separately, as §8(d) asks.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build_encoder(out_dim=1024):
    try:
        import torchvision
        m = torchvision.models.resnet50(weights=None)
        m.fc = torch.nn.Linear(m.fc.in_features, out_dim)
        return m
    except Exception:  # torchvision missing: a small conv stack with the same interface
        return torch.nn.Sequential(torch.nn.Conv2d(3, 64, 7, 4, 3), torch.nn.ReLU(), torch.nn.Conv2d(64, 256, 3, 4, 1),
                                   torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(),
                                   torch.nn.Linear(256, out_dim))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gallery", type=int, default=20000)
    ap.add_argument("--queries", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--k", type=int, default=10)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from art_sbir_b200 import ops, sharded

    torch.manual_seed(0)  # identical weights on every rank
    enc = build_encoder().to(dev).eval().to(memory_format=torch.channels_last)
    a, b = sharded.shard_bounds(args.gallery, world, rank)
    feats = torch.empty(b - a, 1024, device=dev)           # preallocated shard, written in place
    qfeats = torch.empty(args.queries, 1024, device=dev)
    grid = torch.arange(3 * 224 * 224, device=dev, dtype=torch.float32).reshape(1, 3, 224, 224) * 1e-3

    def images(index):
        """Synthetic image i = a deterministic pattern of i (same on every rank, any batching)."""
        i = index.to(torch.float32).reshape(-1, 1, 1, 1)
        return torch.sin(0.37 * i + grid * (1.0 + 0.01 * (i % 17))).contiguous(memory_format=torch.channels_last)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        for lo in range(a, b, args.batch):
            hi = min(lo + args.batch, b)
            feats[lo - a:hi - a] = enc(images(torch.arange(lo, hi, device=dev))).float()
        for lo in range(0, args.queries, args.batch):
            hi = min(lo + args.batch, args.queries)
            # the "sketch" of gallery image p is that image plus noise, so the positive is meaningful
            base = images(torch.arange(lo, hi, device=dev) % args.gallery)
            qfeats[lo:hi] = enc(base + 0.05 * torch.randn_like(base)).float()
    torch.cuda.synchronize()
    t_enc = time.perf_counter() - t0
    pos = torch.arange(args.queries, device=dev) % args.gallery

    for _ in range(2):
        vals, idx, rk = sharded.sharded_pairwise_topk(qfeats, feats, args.k, "euclidean", pos_index=pos,
                                                      shard_offset=a, num_gallery_total=args.gallery)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    vals, idx, rk = sharded.sharded_pairwise_topk(qfeats, feats, args.k, "euclidean", pos_index=pos,
                                                  shard_offset=a, num_gallery_total=args.gallery)
    metrics = ops.retrieval_metrics(rk, args.k)
    torch.cuda.synchronize()
    t_ret = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"n_gpus": world, "gallery": args.gallery, "queries": args.queries, "dim": 1024,
                          "encoder_images_per_s": ((b - a) * world + args.queries * world) / t_enc,
                          "retrieval_pairs_per_s": args.queries * args.gallery / t_ret, "retrieval_ms": t_ret * 1e3,
                          "mrr": metrics["mean_reciprocal_rank"], "topk_acc": metrics["topk_acc"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
