"""Host-side mirror of the reference's `utils` module for the retrieval / triplet path.

Same names, constructor arguments, attributes and error behaviour as the reference
(utils.py:22-77, 258-284 of Peer222/art-sbir), so `inference.py` / `train.py` call sites keep
working; the arithmetic runs in libsbir_b200.so (CUDA tensors only — no CPU fallback).
"""
from __future__ import annotations

import csv
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from . import ops

MARGIN = 0.2  # utils.py:77 ("Sketching without Worrying"); train.py:133 overwrites it from --loss_margin


# ---------------------------------------------------------------------------- N3 ----
def build_stem_index(image_paths: Sequence[Path]) -> Dict[str, int]:
    """stem → FIRST gallery index (the reference's linear scan returns the first match,
    utils.py:22-25).  Built explicitly by the caller once per gallery and passed along — no
    hidden cache keyed on object identity, so two galleries of equal length can never alias."""
    m: Dict[str, int] = {}
    for idx, path in enumerate(image_paths):
        m.setdefault(Path(path).stem, idx)
    return m


def find_image_index(image_paths: List[Path], sketch_name, stem_index: Dict[str, int] = None) -> int:
    """utils.py:22-25: index of the first path whose stem equals `sketch_name`, else -1.
    Stateless like the reference; pass `stem_index=build_stem_index(image_paths)` to replace the
    O(N) scan by a dict lookup when many sketches are resolved against one gallery.  A name that
    is not a string (inference.py:33-37 leaves a LIST for stems with more than three '-' parts)
    matches nothing, as in the reference where `path.stem == [..]` is always False."""
    if not isinstance(sketch_name, str):
        return -1
    if stem_index is not None:
        return stem_index.get(sketch_name, -1)
    for idx, path in enumerate(image_paths):
        if Path(path).stem == sketch_name:
            return idx
    return -1


# ------------------------------------------------------------------------ H1 / H2 ----
class PairwiseDistance(nn.Module):
    """Stands in for `nn.PairwiseDistance(p=2, keepdim=False)` (utils.py:42): ‖x1 − x2 + eps‖₂
    with torch's single-row broadcasting; exposes the same attributes."""

    def __init__(self, p: float = 2.0, eps: float = 1e-6, keepdim: bool = False) -> None:
        super().__init__()
        if float(p) != 2.0 or keepdim or abs(eps - 1e-6) > 1e-12:
            raise NotImplementedError("the sbir_b200 path implements p=2, eps=1e-6, keepdim=False (the reference's use)")
        self.norm = float(p)
        self.eps = eps
        self.keepdim = keepdim

    def forward(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        return ops.pairwise_distance(x1, x2, "euclidean")


class CosineLoss(nn.Module):
    """utils.py:31-38: 1 − cosine similarity along dim 1 (per-operand norm clamp 1e-8)."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, sketch_tensor: torch.Tensor, image_tensor: torch.Tensor) -> torch.Tensor:
        return ops.pairwise_distance(sketch_tensor, image_tensor, "cosine")


cosine_distance = CosineLoss()            # utils.py:40
euclidean_distance = PairwiseDistance()   # utils.py:42


def _loss_type_of(distance_f) -> str:
    if isinstance(distance_f, PairwiseDistance) or isinstance(distance_f, nn.PairwiseDistance):
        return "euclidean"
    if isinstance(distance_f, CosineLoss) or type(distance_f).__name__ == "CosineLoss":
        return "cosine"
    raise TypeError("distance_f must be utils.euclidean_distance or utils.cosine_distance "
                    "(arbitrary distance callables have no fused kernel)")


# ------------------------------------------------------------------------ H6 / H7 ----
class TripletMarginLoss(nn.Module):
    """Replaces `nn.TripletMarginLoss(margin=utils.MARGIN)` (train.py:169)."""

    def __init__(self, margin: float = 1.0) -> None:
        super().__init__()
        self.margin = margin

    def forward(self, anchor, positive, negative):
        return ops.triplet_margin_loss(anchor, positive, negative, self.margin, "euclidean")


class TripletMarginWithDistanceLoss(nn.Module):
    """Replaces `nn.TripletMarginWithDistanceLoss(margin, distance_function)` (train.py:175,
    utils.py:56,69) for the two distance modules of the reference."""

    def __init__(self, *, distance_function=None, margin: float = 1.0) -> None:
        super().__init__()
        self.margin = margin
        self.distance_function = distance_function if distance_function is not None else euclidean_distance
        self._loss_type = _loss_type_of(self.distance_function)

    def forward(self, anchor, positive, negative):
        return ops.triplet_margin_loss(anchor, positive, negative, self.margin, self._loss_type)


class TripletMarginLoss_with_classification(nn.Module):
    """utils.py:49-60: triplet term (fused kernel) + weight·(CE(cs,l) + CE(cp,l)); the
    cross-entropy terms stay in torch."""

    def __init__(self, margin, classification_weight=0.5, distance_f=euclidean_distance):
        super().__init__()
        self.classification_weight = classification_weight
        self.classification_weight2 = 0
        self.margin = margin
        self.triplet_loss = TripletMarginWithDistanceLoss(margin=self.margin, distance_function=distance_f)
        self.classification_loss = nn.CrossEntropyLoss()

    def forward(self, s_logits, p_logits, n_logits, cs_logits, cp_logits, labels):
        return self.triplet_loss(s_logits, p_logits, n_logits) + self.classification_weight * (
            self.classification_loss(cs_logits, labels) + self.classification_loss(cp_logits, labels))


class TripletMarginLoss_with_classification2(nn.Module):
    """utils.py:62-75 (styles + genres heads)."""

    def __init__(self, margin, classification_weight=0.25, classification_weight2=0.5, distance_f=euclidean_distance):
        super().__init__()
        self.classification_weight = classification_weight
        self.classification_weight2 = classification_weight2
        self.margin = margin
        self.triplet_loss = TripletMarginWithDistanceLoss(margin=self.margin, distance_function=distance_f)
        self.classification_loss = nn.CrossEntropyLoss()

    def forward(self, s_logits, p_logits, n_logits, cs_logits, cp_logits, cs_logits2, cp_logits2, labels, labels2):
        classification_loss = self.classification_loss(cs_logits, labels) + self.classification_loss(cp_logits, labels)
        classification_loss2 = self.classification_loss(cs_logits2, labels2) + self.classification_loss(cp_logits2, labels2)
        return (self.triplet_loss(s_logits, p_logits, n_logits) + self.classification_weight * classification_loss
                + self.classification_weight2 * classification_loss2)


# -------------------------------------------------------------------------------- H8 ----
class BatchHardTripletLoss(nn.Module):
    """north_star extension: same call shape as the triplet losses above,
    but the hardest positive / negative are mined over cat(positive, negative) on the tensor
    cores.  `labels` (optional, int64 [batch]) widens the positive set to same-label candidates."""

    def __init__(self, margin: float = MARGIN, distance_f=euclidean_distance) -> None:
        super().__init__()
        self.margin = margin
        self._loss_type = _loss_type_of(distance_f)

    def forward(self, anchor, positive, negative, labels=None):
        return ops.batch_hard_triplet_loss(anchor, positive, negative, self.margin, self._loss_type, labels)


def make_loss_fn(loss_type: str, with_classification: bool, dataset_name: str = "", margin: float = MARGIN):
    """The dispatch of train.py:164-175 with the fused modules."""
    if loss_type not in ("euclidean", "cosine"):
        raise Exception(f"loss type not correct {loss_type}")
    dist = euclidean_distance if loss_type == "euclidean" else cosine_distance
    if with_classification:
        if "Sketchy" in dataset_name:
            return TripletMarginLoss_with_classification(margin=margin, distance_f=dist)
        if "Mixed" in dataset_name:
            w = {"classification_weight": 0.01} if loss_type == "euclidean" else {}
            return TripletMarginLoss_with_classification(margin=margin, distance_f=dist, **w)
        if "Kaggle" in dataset_name:
            w = {"classification_weight": 0, "classification_weight2": 0.2} if loss_type == "euclidean" else {}
            return TripletMarginLoss_with_classification2(margin=margin, distance_f=dist, **w)
    if loss_type == "euclidean":
        return TripletMarginLoss(margin=margin)
    return TripletMarginWithDistanceLoss(margin=margin, distance_function=cosine_distance)


# ---------------------------------------------------------------------------- N2 ----
# Feature store.  The reference writes two CSV files (utils.py:265-284: text floats, ≈20 bytes per value,
# float64 on reload, F8).  Beside them — the CSV pair stays readable by the reference — a binary sidecar:
#   image_features.f32.npy   or   image_features.bf16.npy (the bf16 bit patterns as uint16)
#   image_sqnorm.f32.npy     ‖row‖² of the rows as stored (what sbir_pairwise_topk takes as g_sqnorm)
# so a gallery is stored and reloaded in its own type (cfg4's 10.24 GB bf16 gallery stays 10.24 GB) and the
# scoring pass does not have to re-read it for the norms.
_SIDECAR_F32 = "image_features.f32.npy"
_SIDECAR_BF16 = "image_features.bf16.npy"
_SIDECAR_SQNORM = "image_sqnorm.f32.npy"


class GalleryFeatures:
    """Gallery rows + their stored ‖row‖² travelling together (feature store → scoring pass).  Quacks
    enough like the tensor the reference passes around (`.shape`, `.to(device)`, `len`) for
    process_inference / run_inference; `rows` is the [N, D] tensor, `sqnorm` fp32 [N] or None."""

    def __init__(self, rows: torch.Tensor, sqnorm: Optional[torch.Tensor] = None) -> None:
        self.rows, self.sqnorm = rows, sqnorm

    @property
    def shape(self):
        return self.rows.shape

    @property
    def dtype(self):
        return self.rows.dtype

    def __len__(self) -> int:
        return self.rows.shape[0]

    def to(self, *args, **kwargs) -> "GalleryFeatures":
        rows = self.rows.to(*args, **kwargs)
        sq = None
        if self.sqnorm is not None and rows.dtype == self.rows.dtype:   # a dtype change invalidates stored norms
            sq = self.sqnorm.to(device=rows.device)
        return GalleryFeatures(rows, sq)


def load_image_features(folder_name: str, root: Path = Path("data/image_features"), with_norms: bool = False):
    """utils.py:258-263 → (image_paths, image_features).  Reads the reference's CSV pair; if the binary
    sidecar written by save_image_features is present it is used instead (the gallery's own type, no text
    parsing).  The CSV route returns float64 exactly like the reference (pandas → torch.from_numpy).
    with_norms=True returns (image_paths, GalleryFeatures) carrying the stored ‖row‖² when the sidecar has them."""
    import pandas as pd
    path = Path(root) / folder_name
    image_paths = [Path(p[0]) for p in pd.read_csv(path / "image_paths.csv", header=None).values]
    feats = None
    if (path / _SIDECAR_BF16).is_file():
        feats = torch.from_numpy(np.load(path / _SIDECAR_BF16)).view(torch.bfloat16)
    elif (path / _SIDECAR_F32).is_file():
        feats = torch.from_numpy(np.load(path / _SIDECAR_F32))
    if feats is None:
        feats = torch.from_numpy(pd.read_csv(path / "image_features.csv", header=None).values)
        return (image_paths, GalleryFeatures(feats)) if with_norms else (image_paths, feats)
    if not with_norms:
        return image_paths, feats
    sq = torch.from_numpy(np.load(path / _SIDECAR_SQNORM)) if (path / _SIDECAR_SQNORM).is_file() else None
    if sq is not None and sq.shape[0] != feats.shape[0]:
        sq = None
    return image_paths, GalleryFeatures(feats, sq)


def save_image_features(model_name: str, dataset_name: str, inference_dataset, image_features,
                        root: Path = Path("data/image_features"), write_csv: bool = True, sqnorm: torch.Tensor = None) -> str:
    """utils.py:265-284: same folder naming and the same two CSV files; additionally the binary sidecar
    (rows in their own type, fp32 or bf16, and ‖row‖² when given or carried by a GalleryFeatures)."""
    if isinstance(image_features, GalleryFeatures):
        image_features, sqnorm = image_features.rows, (image_features.sqnorm if sqnorm is None else sqnorm)
    feature_path = Path(root)
    feature_path.mkdir(parents=True, exist_ok=True)
    date_time = datetime.now().strftime("%Y-%m-%d_%H-%M")
    feature_path = feature_path / f"{model_name}_{dataset_name}_{date_time}"
    feature_path.mkdir(parents=True, exist_ok=True)
    with open(feature_path / "image_paths.csv", "w") as f:
        csv.writer(f).writerows([[str(p)] for p in inference_dataset.image_paths])
    rows = image_features.detach().cpu().contiguous()
    if rows.dtype == torch.bfloat16:
        np.save(feature_path / _SIDECAR_BF16, rows.view(torch.uint16).numpy())
    else:
        rows = rows.float()
        np.save(feature_path / _SIDECAR_F32, rows.numpy())
    if sqnorm is not None:
        np.save(feature_path / _SIDECAR_SQNORM, sqnorm.detach().float().cpu().numpy())
    if write_csv:
        with open(feature_path / "image_features.csv", "w") as f:
            csv.writer(f).writerows(rows.float().numpy())
    return feature_path.name
