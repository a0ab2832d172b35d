"""Host-side mirror of the reference's `inference` module for the retrieval path.

Same function names, arguments, return types and error behaviour as inference.py:30-165 of
Peer222/art-sbir.  The reference scores one sketch at a time (PairwiseDistance + a full
`topk(len(G))` sort per query, inference.py:44-52); here all queries of the evaluation go
through ONE fused distance + top-k + rank-count pass on the GPU and the metrics dict is
assembled from device tensors with a single D2H.
"""
from __future__ import annotations

import random
import re
from pathlib import Path
from timeit import default_timer as timer
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops, utils

device = "cuda" if torch.cuda.is_available() else "cpu"  # inference.py:27


def _to_device(t: torch.Tensor) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise RuntimeError("art_sbir_b200.inference needs a CUDA device (no CPU fallback)")
    return t if t.is_cuda else t.cuda(non_blocking=True)


def _common_dtype(a: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    # CSV-loaded galleries are float64 in the reference (F8): the kernels score in fp32 and
    # re-score the survivors with fp64 accumulation.
    if a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16:
        return a, b
    return a.float(), b.float()


def sketch_key(sketch_path, image_paths: Sequence[Path]) -> str:
    """inference.py:31-37: the gallery stem a sketch file name points at (three conventions)."""
    if type(sketch_path) == str:
        sketch_path = Path(sketch_path)
    sketch_name = re.split("-", sketch_path.stem)
    if len(sketch_name) <= 2:
        if "artworks" in str(image_paths[0]):
            sketch_name = sketch_path.stem
        else:
            sketch_name = sketch_name[0]  # sketchy: id-number.png | kaggle: id.png
    elif len(sketch_name) == 3:
        sketch_name = sketch_name[1]  # sketchit: index-id-random_number.png
    return sketch_name


def positive_indices(sketch_paths: Sequence, image_paths: Sequence[Path], verbose: bool = True) -> torch.Tensor:
    """int64 [Q]: gallery index of each sketch's photo, -1 when there is none (N3: ONE explicit
    stem→index dict per call instead of a linear scan per query, utils.py:22-25)."""
    out = []
    stem_index = utils.build_stem_index(image_paths) if len(sketch_paths) > 1 else None
    for sp in sketch_paths:
        name = sketch_key(sp, image_paths)
        idx = utils.find_image_index(image_paths, name, stem_index)
        if idx < 0 and verbose:
            print(f"No image found: {sp} | {name}")  # inference.py:40
        out.append(idx)
    return torch.tensor(out, dtype=torch.int64)


# just one sketch per call - starts at 0
def get_ranking_position(sketch_path, image_paths: List[Path], sketch_feature: torch.Tensor,
                         image_features: torch.Tensor, loss_type) -> int:
    """inference.py:30-57."""
    if loss_type not in ("euclidean", "cosine"):
        raise Exception(f"loss type not correct {loss_type}")
    pos = positive_indices([sketch_path], image_paths)
    if int(pos[0]) < 0:
        return len(image_paths)
    q, g = _common_dtype(_to_device(sketch_feature.reshape(1, -1)), _to_device(image_features))
    rank = ops.rank_of_positive(q, g, _to_device(pos), loss_type)
    return int(rank.item())


# just one sketch per call
def get_topk_images(k: int, image_paths: List[Path], sketch_feature: torch.Tensor, image_features: torch.Tensor,
                    loss_type) -> List[Tuple[str, float]]:
    """inference.py:60-69.  Like `distances.topk(k)` there, k > len(image_features) raises
    (the batched `ops.pairwise_topk` pads with (+inf, -1) instead; a path must never be looked up
    with -1)."""
    if loss_type not in ("euclidean", "cosine"):
        raise Exception(f"loss type not correct {loss_type}")
    if k > image_features.shape[0]:
        raise RuntimeError("selected index k out of range")  # torch's message for topk(k > N)
    q, g = _common_dtype(_to_device(sketch_feature.reshape(1, -1)), _to_device(image_features))
    values, indices = ops.pairwise_topk(q, g, k, loss_type)
    values, indices = values[0].tolist(), indices[0].tolist()
    return [(str(image_paths[i]), float(v)) for i, v in zip(indices, values)]


def _split_gallery(image_features):
    """(rows, stored ‖row‖² or None) of a gallery given as a tensor or as utils.GalleryFeatures."""
    if isinstance(image_features, utils.GalleryFeatures):
        return image_features.rows, image_features.sqnorm
    return image_features, None


def evaluate_embeddings(sketch_features: torch.Tensor, image_features, pos_index: torch.Tensor,
                        loss_type: str = "euclidean", k: int = 10, sample_indices: Sequence[int] = ()) -> Dict:
    """The arithmetic of process_inference (inference.py:109-133) for already-encoded sketches:
    ranks, MRR, cumulative top-k accuracy, describe() stats, and the top-k lists of the
    sampled queries.  Returns the metrics dict plus 'ranks0' (device tensor) and 'samples'.
    `image_features` may be a utils.GalleryFeatures: its stored norms then spare the pass a read of the gallery."""
    rows, sqnorm = _split_gallery(image_features)
    rows = _to_device(rows)
    q, g = _common_dtype(_to_device(sketch_features), rows)
    if sqnorm is not None and g.dtype == rows.dtype:
        sqnorm = _to_device(sqnorm)
    else:
        sqnorm = None
    values, indices, rank0 = ops.pairwise_topk(q, g, k, loss_type, pos_index=_to_device(pos_index), gallery_sqnorm=sqnorm)
    stats = ops.retrieval_metrics(rank0, k)
    samples = {}
    if len(sample_indices):
        sel = torch.tensor(list(sample_indices), device=values.device, dtype=torch.int64)
        v = values.index_select(0, sel).cpu().tolist()
        ix = indices.index_select(0, sel).cpu().tolist()
        samples = {int(i): (ix_i, v_i) for i, ix_i, v_i in zip(sample_indices, ix, v)}
    stats["ranks0"] = rank0
    stats["samples"] = samples
    return stats


def process_inference(model, dataset, inference_dataset, dataloader, image_features, start_time,
                      with_classification, loss_type):
    """inference.py:94-136, same arguments and the same result dict
    {mean_reciprocal_rank, size, inference_time, count, mean, std, min, 25%, 50%, 75%, max,
     topk_acc, retrieval_samples}."""
    if loss_type not in ("euclidean", "cosine"):
        raise Exception(f"loss type not correct {loss_type}")
    k = 10
    random.seed(11)
    random_indices = [random.randrange(0, len(dataset)) for _ in range(10)]

    if not isinstance(image_features, utils.GalleryFeatures):
        image_features = _to_device(image_features)
    model.to(device)
    model.eval()
    feats = []
    with torch.inference_mode():
        # shuffle=False: the i-th encoded sketch is dataset.sketch_paths[i] (any batch size)
        for batch in dataloader:
            out = model(batch[0].to(device))
            feats.append(out[0] if with_classification else out)
    sketch_features = torch.cat([f.reshape(-1, f.shape[-1]) for f in feats]) if feats else \
        torch.empty((0, image_features.shape[1]), device=device)
    n = sketch_features.shape[0]
    sketch_paths = [dataset.sketch_paths[i] for i in range(n)]
    pos = positive_indices(sketch_paths, inference_dataset.image_paths)

    sample_ids = sorted({i for i in random_indices if i < n})
    if sample_ids and k > image_features.shape[0]:
        # get_topk_images(k=10, ...) of a sampled query (inference.py:120-121): topk(k > N) raises
        raise RuntimeError("selected index k out of range")
    ev = evaluate_embeddings(sketch_features, image_features, pos, loss_type, k, sample_ids)
    retrieval_samples = []
    image_paths = inference_dataset.image_paths
    for i in range(n):  # same order and multiplicity rule as inference.py:120-121
        if random_indices.count(i) > 0:
            ix, v = ev["samples"][i]
            retrieval_samples.append({str(dataset.sketch_paths[i]): [(str(image_paths[j]), float(d)) for j, d in zip(ix, v)]})

    # the reference divides by len(dataset) (inference.py:124-125)
    scale = n / len(dataset) if len(dataset) else 1.0
    time = timer() - start_time
    stats = {"mean_reciprocal_rank": ev["mean_reciprocal_rank"] * scale, "size": len(inference_dataset),
             "inference_time": time}
    for key in ("count", "mean", "std", "min", "25%", "50%", "75%", "max"):
        stats[key] = ev[key]
    stats["topk_acc"] = [a * scale for a in ev["topk_acc"]]
    stats["retrieval_samples"] = retrieval_samples
    return stats


def compute_image_features(model, dataset, with_classification: bool, batch_size: int = 50,
                           inference_dataset=None, save: bool = True, gallery_dtype: Optional[torch.dtype] = None,
                           row_range: Optional[Tuple[int, int]] = None, feature_root: Optional[Path] = None,
                           write_csv: bool = True):
    """inference.py:72-92 (N1): gallery embeddings for the de-duplicated, sorted photo paths.
    Every batch of encoder output goes through sbir_gallery_append straight into a preallocated [N, D]
    device buffer (ops.GalleryBuffer: storage type `gallery_dtype`, default the encoder's; ‖row‖² written
    by the same kernel) instead of the reference's torch.cat growth + .cpu(); the gallery stays on the GPU
    for the scoring pass.  `row_range=(a, b)` encodes only those gallery rows — with one encoder replica
    per GPU and sharded.shard_bounds this yields the row-sharded layout of the multi-GPU path directly.
    Returns (inference_dataset, utils.GalleryFeatures, feature folder name or None)."""
    from torch.utils.data import DataLoader, Subset
    if inference_dataset is None:
        inference_dataset = InferenceDataset(dataset.photo_paths, dataset.transform)
    a, b = (0, len(inference_dataset)) if row_range is None else row_range
    source = inference_dataset if row_range is None else Subset(inference_dataset, range(a, b))
    dataloader = DataLoader(dataset=source, batch_size=batch_size, num_workers=0, shuffle=False)
    model.to(device)
    model.eval()
    buf: Optional[ops.GalleryBuffer] = None
    with torch.inference_mode():
        for images in dataloader:
            out = model(images.to(device))
            out = (out[0] if with_classification else out)
            out = out.reshape(-1, out.shape[-1])
            if buf is None:
                store = gallery_dtype if gallery_dtype is not None else (torch.bfloat16 if out.dtype == torch.bfloat16 else torch.float32)
                buf = ops.GalleryBuffer(b - a, out.shape[1], store, device=out.device)
            buf.append(out)
    if buf is None:
        feats = utils.GalleryFeatures(torch.empty((0, 0), device=device), torch.empty(0, device=device))
    else:
        feats = utils.GalleryFeatures(buf.rows, buf.sqnorm)
    feature_path = None
    if save:
        kw = {} if feature_root is None else {"root": feature_root}
        feature_path = utils.save_image_features(model.__class__.__name__, dataset.state_dict["dataset"],
                                                 inference_dataset, feats, write_csv=write_csv, **kw)
    return inference_dataset, feats, feature_path


class InferenceDataset(torch.utils.data.Dataset):
    """data_preparation.py:24-41: duplicate-free, sorted gallery paths — the index space every
    ranked index refers to."""

    def __init__(self, image_paths: List[Path], transform=None):
        super().__init__()
        self.transform = transform
        self.image_paths = list(dict.fromkeys(image_paths))
        self.image_paths.sort()

    def load_image(self, idx: int):
        from PIL import Image
        return Image.open(self.image_paths[idx])

    def __len__(self) -> int:
        return len(self.image_paths)

    def __getitem__(self, idx: int):
        img = self.load_image(idx)
        return self.transform(img) if self.transform is not None else img


def run_inference(model, dataset, folder_name: str = None, loss_type="euclidean", second_dataset=None,
                  inference_dataset=None, feature_root: Optional[Path] = None) -> Dict:
    """inference.py:140-165.  `second_dataset` stands for the KaggleInferenceV1 sketches the
    reference loads for Kaggle/Mixed datasets (inference.py:157-160); dataset construction is
    outside this path, so the caller supplies it.  `inference_dataset` (optional) replaces the
    InferenceDataset built from dataset.photo_paths; `feature_root` the reference's data/image_features."""
    from torch.utils.data import DataLoader
    start_time = timer()
    with_classification = "with_classification" in type(model).__name__
    name = dataset.state_dict["dataset"] if hasattr(dataset, "state_dict") else ""
    two_sets = "Kaggle" in name or "Mixed" in name
    if two_sets and second_dataset is None:
        # the reference always returns the three-key dict for these datasets (inference.py:157-160);
        # returning the flat dict instead would silently change what visualization.visualize reads
        raise ValueError(f"run_inference on {name} evaluates a second sketch set (KaggleInferenceV1, "
                         "inference.py:157-160): pass it as second_dataset=")
    if folder_name:
        feature_folder = folder_name
        kw = {} if feature_root is None else {"root": feature_root}
        image_paths, image_features = utils.load_image_features(folder_name, with_norms=True, **kw)
        inference_dataset = InferenceDataset(image_paths, getattr(model, "transform", None))
        print("Image features loaded from file")
    else:
        inference_dataset, image_features, feature_folder = compute_image_features(
            model, dataset, with_classification, inference_dataset=inference_dataset, feature_root=feature_root)
    dataloader = DataLoader(dataset=dataset, batch_size=64, num_workers=0, shuffle=False)  # N4: batched queries
    inference_dict = process_inference(model, dataset, inference_dataset, dataloader, image_features, start_time,
                                       with_classification, loss_type)
    if two_sets:
        dataloader2 = DataLoader(dataset=second_dataset, batch_size=64, num_workers=0, shuffle=False)
        inference_dict2 = process_inference(model, second_dataset, inference_dataset, dataloader2, image_features,
                                            inference_dict["inference_time"], with_classification, loss_type)
        return {"image_features": feature_folder, "drawing_stats": inference_dict, "sketch_stats": inference_dict2}
    inference_dict["image_features"] = feature_folder
    return inference_dict
