"""BASELINE config 5 — end-to-end evaluation: a PyTorch image encoder on synthetic 224x224 sketches and artworks
feeding the gallery-sharded distance / top-K / rank path.

    python tools/e2e_cfg5.py [--gallery 20000] [--queries 2000] [--batch 256]
    torchrun --nproc-per-node 8 tools/e2e_cfg5.py --gallery 200000 --queries 20000

The encoder is OUT OF SCOPE of this repo (it stays PyTorch) and the reference's models.py is not
redistributed; `SketchEncoder` below re-declares the SHAPE of what the reference trains — a CLIP-style ResNet-50
(`ModifiedResNet(layers=(3,4,6,3), output_dim=1024)`, reference models.py:275-379): three-convolution stem, bottleneck
blocks that down-sample with average pooling, 2048 channels at 7x7, and an attention pool (32 heads) that emits a
1024-d embedding — with random weights (there is no network for checkpoints).  A randomly initialised deep ReLU network
maps every image to nearly the same point (the embeddings "collapse": a huge common component, tiny differences), which
says nothing about a trained encoder, so the FINAL AFFINE is calibrated once, load-free: the mean and scale of the
embedding are measured on 512 synthetic images (the same on every rank) and folded into the output — a reparametrisation of
the last projection's bias and gain, nothing is trained.  (The retrieval path is exact on collapsed embeddings too — its
centred escalation pass exists for them, tests/test_gpu_parity.py — but that is not what config 5 is about.)

Each rank encodes ITS slice of the gallery images straight into its shard through `compute_image_features`' append kernel
(data-parallel encoding yields the row-sharded layout for free), every rank encodes all queries, then
`sharded_pairwise_topk` scores them.  Reports encoder images/s and retrieval pairs/s separately (§8d), and CHECKS the
sharded result of 64 sampled queries against the CPU oracle on the gathered gallery."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch import nn  # noqa: E402


class _Block(nn.Module):
    """1x1 -> 3x3 -> (avg-pool) -> 1x1 bottleneck, expansion 4, shortcut pooled the same way."""

    def __init__(self, cin, planes, stride):
        super().__init__()
        cout = planes * 4
        self.a = nn.Sequential(nn.Conv2d(cin, planes, 1, bias=False), nn.BatchNorm2d(planes), nn.ReLU(inplace=True),
                               nn.Conv2d(planes, planes, 3, padding=1, bias=False), nn.BatchNorm2d(planes), nn.ReLU(inplace=True),
                               nn.AvgPool2d(stride) if stride > 1 else nn.Identity(),
                               nn.Conv2d(planes, cout, 1, bias=False), nn.BatchNorm2d(cout))
        self.skip = None
        if stride > 1 or cin != cout:
            self.skip = nn.Sequential(nn.AvgPool2d(stride) if stride > 1 else nn.Identity(),
                                      nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        return F.relu(self.a(x) + (x if self.skip is None else self.skip(x)))


class _AttentionPool(nn.Module):
    """Mean token + HxW tokens with a learned position table; one multi-head attention step queried by the mean token."""

    def __init__(self, side, width, heads, out_dim):
        super().__init__()
        self.pos = nn.Parameter(torch.randn(side * side + 1, width) / width ** 0.5)
        self.q, self.k, self.v = nn.Linear(width, width), nn.Linear(width, width), nn.Linear(width, width)
        self.out = nn.Linear(width, out_dim)
        self.heads = heads

    def forward(self, x):
        b, c, h, w = x.shape
        t = x.flatten(2).transpose(1, 2)                              # [B, HW, C]
        t = torch.cat([t.mean(1, keepdim=True), t], 1) + self.pos.to(t.dtype)
        hd = c // self.heads
        q = self.q(t[:, :1]).view(b, 1, self.heads, hd).transpose(1, 2)
        k = self.k(t).view(b, -1, self.heads, hd).transpose(1, 2)
        v = self.v(t).view(b, -1, self.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, c)
        return self.out(o)


class SketchEncoder(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), width=64, out_dim=1024, heads=32, resolution=224):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(3, width // 2, 3, 2, 1, bias=False), nn.BatchNorm2d(width // 2), nn.ReLU(inplace=True),
                                  nn.Conv2d(width // 2, width // 2, 3, padding=1, bias=False), nn.BatchNorm2d(width // 2), nn.ReLU(inplace=True),
                                  nn.Conv2d(width // 2, width, 3, padding=1, bias=False), nn.BatchNorm2d(width), nn.ReLU(inplace=True),
                                  nn.AvgPool2d(2))
        blocks, cin = [], width
        for stage, n in enumerate(layers):
            planes = width * 2 ** stage
            for i in range(n):
                blocks.append(_Block(cin, planes, 2 if (i == 0 and stage > 0) else 1))
                cin = planes * 4
        self.body = nn.Sequential(*blocks)
        self.pool = _AttentionPool(resolution // 32, cin, heads, out_dim)
        self.register_buffer("shift", torch.zeros(out_dim))
        self.register_buffer("gain", torch.ones(out_dim))

    def forward(self, x):
        return (self.pool(self.body(self.stem(x))).float() - self.shift) * self.gain


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gallery", type=int, default=20000)
    ap.add_argument("--queries", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--check", type=int, default=64, help="queries compared with the CPU oracle")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from art_sbir_b200 import inference as inf, ops, sharded

    torch.manual_seed(0)  # identical weights on every rank
    enc = SketchEncoder().to(dev).eval().to(memory_format=torch.channels_last)
    grid = torch.arange(3 * 224 * 224, device=dev, dtype=torch.float32).reshape(1, 3, 224, 224) * 1e-3

    def images(index):
        """Synthetic image i = a deterministic pattern of i (same on every rank, any batching)."""
        i = index.to(torch.float32).reshape(-1, 1, 1, 1)
        return torch.sin(0.37 * i + grid * (1.0 + 0.01 * (i % 17))).contiguous(memory_format=torch.channels_last)

    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        sample = enc(images(torch.arange(0, args.gallery, max(1, args.gallery // 512), device=dev)[:512]))
    enc.shift.copy_(sample.mean(0))
    enc.gain.copy_(1.0 / sample.std(0).clamp_min(1e-6))

    class Gallery(inf.InferenceDataset):          # gallery "files" are synthesised on the fly
        def __init__(self, n):
            self.image_paths, self.transform = list(range(n)), None

        def load_image(self, idx):
            return torch.tensor(idx)

    class Encode(nn.Module):                      # index -> image -> embedding, so compute_image_features drives the real path
        def forward(self, idx):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return enc(images(idx))

    a, b = sharded.shard_bounds(args.gallery, world, rank)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, gal, _ = inf.compute_image_features(Encode(), None, False, batch_size=args.batch, inference_dataset=Gallery(args.gallery),
                                           save=False, row_range=(a, b))
    qfeats = torch.empty(args.queries, 1024, device=dev)
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        for lo in range(0, args.queries, args.batch):
            hi = min(lo + args.batch, args.queries)
            # the "sketch" of gallery image p is that image plus noise, so the positive is meaningful
            base = images(torch.arange(lo, hi, device=dev) % args.gallery)
            qfeats[lo:hi] = enc(base + 0.05 * torch.randn_like(base))
    torch.cuda.synchronize()
    t_enc = time.perf_counter() - t0
    pos = torch.arange(args.queries, device=dev) % args.gallery
    feats = gal.rows

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

    # ---- correctness: sampled queries against the CPU oracle on the gathered gallery ----
    if world > 1:
        parts = [torch.empty(sharded.shard_bounds(args.gallery, world, r)[1] - sharded.shard_bounds(args.gallery, world, r)[0], 1024,
                             device=dev) for r in range(world)]
        dist.all_gather(parts, feats.contiguous()) if len({p.shape[0] for p in parts}) == 1 else None
        if len({p.shape[0] for p in parts}) != 1:   # ragged shards: gather through a padded buffer
            m = max(p.shape[0] for p in parts)
            padded = torch.zeros(m, 1024, device=dev)
            padded[:feats.shape[0]] = feats
            bufs = [torch.empty_like(padded) for _ in range(world)]
            dist.all_gather(bufs, padded)
            parts = [bufs[r][:parts[r].shape[0]] for r in range(world)]
        full = torch.cat(parts)
    else:
        full = feats
    check = None
    if rank == 0:
        from oracle import sbir_oracle as O
        sel = torch.linspace(0, args.queries - 1, min(args.check, args.queries)).round().long()
        Qc, Gc = qfeats[sel.to(dev)].cpu(), full.cpu()
        ref_v, ref_i = O.pairwise_topk_batched(Qc, Gc, args.k, "euclidean")
        ref_r = O.rank_of_positive_batched(Qc, Gc, pos[sel.to(dev)].cpu(), "euclidean")
        got_i, got_v, got_r = idx[sel.to(dev)].cpu(), vals[sel.to(dev)].cpu(), rk[sel.to(dev)].cpu()
        differ = got_i != ref_i
        tie = differ & ((got_v - ref_v).abs() <= 1e-4 * ref_v.abs().clamp_min(1e-30))
        check = {"queries": int(sel.numel()), "topk_index_mismatches": int((differ & ~tie).sum()), "tie_swaps": int(tie.sum()),
                 "max_rel_dist_err": float(((got_v - ref_v).abs() / ref_v.abs().clamp_min(1e-30)).max()),
                 "rank_mismatches": int((got_r != ref_r).sum()), "oracle": "CPU oracle (reference distance modules + topk) on the gathered gallery"}
        cen = full - full.mean(0)
        print(json.dumps({"n_gpus": world, "gallery": args.gallery, "queries": args.queries, "dim": 1024,
                          "encoder": "CLIP-style ResNet-50 (3,4,6,3) + 32-head attention pool -> 1024-d, random init, calibrated output affine, bf16 autocast",
                          "encoder_images_per_s": (args.gallery + args.queries * world) / t_enc,
                          "retrieval_pairs_per_s": args.queries * args.gallery / t_ret, "retrieval_ms": t_ret * 1e3,
                          "embedding_common_component_over_spread": float(full.mean(0).norm() / cen.norm(dim=1).mean()),
                          "mrr": metrics["mean_reciprocal_rank"], "topk_acc": metrics["topk_acc"], "oracle_check": check}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
