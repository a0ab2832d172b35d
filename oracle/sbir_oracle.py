"""TEST INFRASTRUCTURE — CPU oracle for the retrieval / triplet hot path of Peer222/art-sbir.

This is a restatement, on CPU torch, of what the reference computes on this path; every
function cites the reference file:line it follows (paths relative to the reference checkout).
The arithmetic of the path lives in a third-party dependency that is not vendored in the
reference: PyTorch ATen (`pairwise_distance`, `cosine_similarity`, `topk`, `triplet_margin_loss`,
`clamp_min`, `mean`).  The reference pins no torch version (README.md:6 "Required Packages:
TODO"); the oracle version is this image's torch 2.11.0 and it calls the same ATen ops the
reference calls, so "the reference's arithmetic" and "the oracle's arithmetic" are the same
code for H1–H7.  H8 (batch-hard) and H9 (L2 normalise) do not exist in the reference as ops;
they are defined in SURVEY.md §8a and restated here from the reference's own primitives.

Pinning: the reference ships no tests or golden vectors.  tests/golden/make_golden.py runs
the reference's UNMODIFIED functions (imported via oracle/ref_import.py in the build
container) on seeded inputs and freezes their outputs under tests/golden/; tests/test_oracle.py
checks this file against those fixtures and, when /root/reference is present, against the
live reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (art_sbir_b200/) never does.
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

MARGIN = 0.2  # utils.py:77

# --------------------------------------------------------------------------- H1 / H2 ----
# utils.py:42   euclidean_distance = nn.PairwiseDistance(p=2, keepdim=False)
euclidean_distance = nn.PairwiseDistance(p=2, keepdim=False)


class CosineLoss(nn.Module):
    """utils.py:31-38: (cosine_similarity * -1) + 1 with nn.CosineSimilarity(dim=1)."""

    def __init__(self) -> None:
        super().__init__()
        self.cosine_similarity = nn.CosineSimilarity(dim=1)

    def forward(self, sketch_tensor, image_tensor):
        return (self.cosine_similarity(sketch_tensor, image_tensor) * -1) + 1


cosine_distance = CosineLoss()  # utils.py:40


def distances(sketch_feature: torch.Tensor, image_features: torch.Tensor, loss_type: str) -> torch.Tensor:
    """inference.py:43-48 / 61-64: one [1,D] sketch embedding against the [N,D] gallery."""
    if loss_type == "euclidean":
        return euclidean_distance(sketch_feature, image_features)
    elif loss_type == "cosine":
        return cosine_distance(sketch_feature, image_features)
    raise Exception(f"loss type not correct {loss_type}")  # inference.py:48


def distances_fp64(q: np.ndarray, G: np.ndarray, loss_type: str) -> np.ndarray:
    """Real-number evaluation (numpy fp64) of the same formulas on the given inputs; used to
    account for tolerance: the CUDA path must sit at least as close to this as torch fp32 does."""
    q = np.asarray(q, dtype=np.float64).reshape(1, -1)
    G = np.asarray(G, dtype=np.float64)
    if loss_type == "euclidean":
        return np.sqrt(((q - G + 1e-6) ** 2).sum(1))
    qn = q / max(np.linalg.norm(q), 1e-8)
    Gn = G / np.maximum(np.linalg.norm(G, axis=1, keepdims=True), 1e-8)
    return 1.0 - (qn * Gn).sum(1)


# -------------------------------------------------------------------------------- H3 ----
def sketch_name_to_key(sketch_path, image_paths: Sequence[Path]) -> str:
    """inference.py:31-37 — the three file-name conventions that map a sketch to its photo."""
    if type(sketch_path) == str:
        sketch_path = Path(sketch_path)
    sketch_name = re.split("-", sketch_path.stem)
    if len(sketch_name) <= 2:
        if "artworks" in str(image_paths[0]):
            sketch_name = sketch_path.stem
        else:
            sketch_name = sketch_name[0]
    elif len(sketch_name) == 3:
        sketch_name = sketch_name[1]
    return sketch_name


def find_image_index(image_paths: Sequence[Path], sketch_name: str) -> int:
    """utils.py:22-25 — first gallery path whose stem equals the key, else -1."""
    for idx, path in enumerate(image_paths):
        if path.stem == sketch_name:
            return idx
    return -1


def ranking_position(sketch_feature: torch.Tensor, image_features: torch.Tensor, pos_img_index: int,
                     loss_type: str) -> int:
    """inference.py:38-57 with the positive's index already resolved: 0-based position of the
    positive in the full ascending sort (`topk(len(G), largest=False)`), len(G) if there is none."""
    if pos_img_index < 0:
        return len(image_features)  # inference.py:39-41
    d = distances(sketch_feature, image_features, loss_type)
    _, indices = d.topk(len(image_features), largest=False)  # inference.py:49
    hits = (indices == pos_img_index).nonzero().squeeze()  # inference.py:52,55
    return int(hits.item() if hits.dim() == 0 else hits[0].item())


def get_ranking_position(sketch_path, image_paths: List[Path], sketch_feature, image_features, loss_type) -> int:
    """inference.py:30-57, full signature."""
    key = sketch_name_to_key(sketch_path, image_paths)
    return ranking_position(sketch_feature, image_features, find_image_index(image_paths, key), loss_type)


# -------------------------------------------------------------------------------- H4 ----
def topk_images(k: int, sketch_feature, image_features, loss_type) -> Tuple[torch.Tensor, torch.Tensor]:
    """inference.py:60-65: (values, indices) of `distances.topk(k, largest=False)`."""
    d = distances(sketch_feature, image_features, loss_type)
    return d.topk(k, largest=False)


def get_topk_images(k: int, image_paths: List[Path], sketch_feature, image_features, loss_type):
    """inference.py:60-69, full signature → [(str(path), float(dist))]."""
    values, indices = topk_images(k, sketch_feature, image_features, loss_type)
    return list(zip([str(image_paths[i]) for i in indices], [v.item() for v in values]))


# -------------------------------------------------------------------------------- H5 ----
def retrieval_metrics(ranks0: Sequence[int], k: int = 10) -> Dict:
    """inference.py:95-98,113-118,123-133: MRR, cumulative top-k accuracy and the pandas
    describe() of the 1-based ranks, from 0-based ranks."""
    import pandas as pd
    ranks = []
    mrr = 0.0
    topk_acc = np.zeros(k)
    for rank in ranks0:
        rank = int(rank)
        ranks.append(rank + 1)
        mrr += 1 / (rank + 1)
        if rank < 10:
            topk_acc[rank:] += 1  # inference.py:118 (hard-coded 10 in the reference; k == 10 there)
    n = len(ranks)
    stats = {"mean_reciprocal_rank": mrr / n}
    stats.update(pd.DataFrame(ranks, columns=["rank"]).describe().to_dict()["rank"])
    stats["topk_acc"] = list(topk_acc / n)
    return stats


def process_inference(query_features: torch.Tensor, image_features: torch.Tensor, pos_index: Sequence[int],
                      loss_type: str, k: int = 10) -> Dict:
    """inference.py:94-136 with an identity encoder and resolved positives: the per-query loop."""
    ranks0 = [ranking_position(query_features[i:i + 1], image_features, int(pos_index[i]), loss_type)
              for i in range(len(query_features))]
    stats = retrieval_metrics(ranks0, k)
    stats["size"] = len(image_features)
    stats["ranks0"] = ranks0
    return stats


# -------------------------------------------------------------------------- H6 / H7 ----
def triplet_margin_loss(a, p, n, margin: float = MARGIN, loss_type: str = "euclidean") -> torch.Tensor:
    """train.py:169 nn.TripletMarginLoss(margin) (euclidean, no classifier) and train.py:175
    nn.TripletMarginWithDistanceLoss(margin, distance_function=cosine_distance)."""
    if loss_type == "euclidean":
        return nn.TripletMarginLoss(margin=margin)(a, p, n)
    return nn.TripletMarginWithDistanceLoss(margin=margin, distance_function=cosine_distance)(a, p, n)


class TripletMarginLoss_with_classification(nn.Module):
    """utils.py:49-60."""

    def __init__(self, margin, classification_weight=0.5, distance_f=euclidean_distance):
        super().__init__()
        self.classification_weight = classification_weight
        self.classification_weight2 = 0
        self.margin = margin
        self.triplet_loss = nn.TripletMarginWithDistanceLoss(margin=self.margin, distance_function=distance_f)
        self.classification_loss = nn.CrossEntropyLoss()

    def forward(self, s, p, n, cs, cp, labels):
        return self.triplet_loss(s, p, n) + self.classification_weight * (
            self.classification_loss(cs, labels) + self.classification_loss(cp, labels))


class TripletMarginLoss_with_classification2(nn.Module):
    """utils.py:62-75."""

    def __init__(self, margin, classification_weight=0.25, classification_weight2=0.5, distance_f=euclidean_distance):
        super().__init__()
        self.classification_weight = classification_weight
        self.classification_weight2 = classification_weight2
        self.margin = margin
        self.triplet_loss = nn.TripletMarginWithDistanceLoss(margin=self.margin, distance_function=distance_f)
        self.classification_loss = nn.CrossEntropyLoss()

    def forward(self, s, p, n, cs, cp, cs2, cp2, labels, labels2):
        c1 = self.classification_loss(cs, labels) + self.classification_loss(cp, labels)
        c2 = self.classification_loss(cs2, labels2) + self.classification_loss(cp2, labels2)
        return self.triplet_loss(s, p, n) + self.classification_weight * c1 + self.classification_weight2 * c2


# -------------------------------------------------------------------------------- H8 ----
def pairwise_matrix(a: torch.Tensor, x: torch.Tensor, loss_type: str) -> torch.Tensor:
    """D_ij = dist(a_i, x_j) with the reference's distance modules, row by row (differentiable)."""
    return torch.stack([distances(a[i:i + 1], x, loss_type) for i in range(a.shape[0])])


def batch_hard_triplet_loss(a, p, n, margin: float = MARGIN, loss_type: str = "euclidean",
                            labels: Optional[torch.Tensor] = None):
    """SURVEY.md §8a H8 (Hermans et al. 2017 adapted to the reference's triplet structure,
    data_preparation.py:67-69,214-222): candidates X = cat(p, n); positives of anchor i are
    {i} (or the candidates sharing labels[i]; negatives get unique negative labels);
    hp_i = max over positives, hn_i = min over the rest; mean_i max(0, margin + hp_i − hn_i).
    Returns (loss, hardest_pos_idx, hardest_neg_idx)."""
    B = a.shape[0]
    x = torch.cat([p, n])
    D = pairwise_matrix(a, x, loss_type)
    if labels is None:
        pos_mask = torch.zeros(B, 2 * B, dtype=torch.bool)
        pos_mask[torch.arange(B), torch.arange(B)] = True
    else:
        cand = torch.cat([labels.to(torch.int64), -1 - torch.arange(B, dtype=torch.int64)])
        pos_mask = cand[None, :] == labels.to(torch.int64)[:, None]
    hp, hpi = D.masked_fill(~pos_mask, -float("inf")).max(1)
    hn, hni = D.masked_fill(pos_mask, float("inf")).min(1)
    loss = torch.clamp_min(margin + hp - hn, 0).mean()
    return loss, hpi, hni


# -------------------------------------------------------------------------------- H9 ----
def l2_normalize(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """x / max(‖x‖₂, eps): the per-operand normalisation inside nn.CosineSimilarity (utils.py:34)."""
    return x / torch.linalg.vector_norm(x, 2, dim=1, keepdim=True).clamp_min(eps)


# ------------------------------------------------------ batched restatements (north_star) ----
def pairwise_topk_batched(Q: torch.Tensor, G: torch.Tensor, k: int, loss_type: str, fp64: bool = False):
    """The batched restatement of H1+H4 for all queries: per-query reference distance, then
    topk — identical arithmetic to the loop, just collected.  fp64=True evaluates the same
    formula in double (tolerance accounting; the reference itself runs in fp64 when the
    gallery was loaded from CSV, utils.py:258-263)."""
    if fp64:
        Q, G = Q.double(), G.double()
    vals, idxs = [], []
    for i in range(Q.shape[0]):
        v, ix = topk_images(min(k, G.shape[0]), Q[i:i + 1], G, loss_type)
        vals.append(v)
        idxs.append(ix)
    return torch.stack(vals), torch.stack(idxs)


def rank_of_positive_batched(Q: torch.Tensor, G: torch.Tensor, pos_index: torch.Tensor, loss_type: str,
                             fp64: bool = False) -> torch.Tensor:
    """count(d < d_pos) per query — equals ranking_position on tie-free data."""
    if fp64:
        Q, G = Q.double(), G.double()
    out = []
    for i in range(Q.shape[0]):
        pi = int(pos_index[i])
        if pi < 0:
            out.append(G.shape[0])
            continue
        d = distances(Q[i:i + 1], G, loss_type)
        out.append(int((d < d[pi]).sum()))
    return torch.tensor(out, dtype=torch.int64)


def cdist_topk(Q: torch.Tensor, G: torch.Tensor, k: int):
    """north_star's 'torch.cdist/topk path' (euclidean without the +1e-6)."""
    return torch.cdist(Q, G).topk(k, dim=1, largest=False)


# --------------------------------------------------------------------- synthetic inputs ----
def synthetic_embeddings(num_q: int, num_g: int, dim: int, seed: int = 1234, beta: Optional[float] = None,
                         num_classes: int = 125):
    """seeded clustered generator — 125 class centroids, gallery = centroid +
    noise, each query = centroid of a random gallery row + beta·(that row's noise) + noise, so
    recall@K is non-trivial (recall@1 ≈ 0.47, @10 ≈ 0.86 at 1k×10k).  Returns fp32 (Q, G, pos)."""
    if beta is None:
        beta = 0.06 if dim >= 2048 else 0.12 if dim >= 512 else 0.3
    g = torch.Generator().manual_seed(seed)
    cent = torch.randn(num_classes, dim, generator=g)
    cls = torch.arange(num_g) % num_classes
    noise = torch.randn(num_g, dim, generator=g)
    G = cent[cls] + noise
    if num_q <= num_g:
        pos = torch.randperm(num_g, generator=g)[:num_q]
    else:
        pos = torch.randint(0, num_g, (num_q,), generator=g)
    Q = cent[cls[pos]] + beta * noise[pos] + torch.randn(num_q, dim, generator=g)
    return Q.contiguous(), G.contiguous(), pos.to(torch.int64)
