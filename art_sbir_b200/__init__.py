"""art_sbir_b200 — B200-native (sm_100a) retrieval / triplet hot path of Peer222/art-sbir.

Layout (only what the path needs):
    csrc/         hand-written CUDA kernels + the extern "C" boundary (include/sbir_b200.h)
    _build.py     in-tree nvcc build of lib/libsbir_b200.so
    _binding.py   ctypes prototypes of the C ABI
    ops.py        batched device surface (pairwise_topk, rank_of_positive, losses, ...)
    utils.py      mirror of the reference's utils.py for this path (distance + loss modules)
    inference.py  mirror of the reference's inference.py (get_ranking_position, process_inference, ...)
    sharded.py    gallery sharding over torch.distributed (NCCL all-gather + merge kernel)
"""
from . import _binding  # noqa: F401
from .ops import (batch_hard_triplet_loss, l2_normalize, pairwise_distance, pairwise_topk,  # noqa: F401
                  rank_of_positive, retrieval_metrics, topk_merge, triplet_margin_loss)

__all__ = ["pairwise_topk", "rank_of_positive", "retrieval_metrics", "triplet_margin_loss",
           "batch_hard_triplet_loss", "l2_normalize", "pairwise_distance", "topk_merge"]
