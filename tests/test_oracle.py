"""CPU suite, part 1: the oracle (oracle/sbir_oracle.py) against the golden fixtures minted from
the reference's own functions, against the live reference when its checkout is present, and
against fp64 evaluation / edge cases.  No GPU, no product code."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_import, sbir_oracle as O


def _paths(arr):
    return [Path(str(p)) for p in arr]


def test_oracle_ranks_match_reference_golden(retrieval_golden):
    g = retrieval_golden
    Q, G = torch.from_numpy(g["Q"]), torch.from_numpy(g["G"])
    image_paths, sketch_paths = _paths(g["image_paths"]), _paths(g["sketch_paths"])
    ranks = [O.get_ranking_position(sketch_paths[i], image_paths, Q[i:i + 1], G, g["loss_type"]) for i in range(len(Q))]
    assert ranks == g["ranks"].tolist()
    assert ranks[5] == len(image_paths)  # the sketch without a photo (inference.py:39-41)
    # count-less-than restatement agrees on this tie-free data
    pos = torch.tensor([O.find_image_index(image_paths, O.sketch_name_to_key(s, image_paths)) for s in sketch_paths])
    assert O.rank_of_positive_batched(Q, G, pos, g["loss_type"]).tolist() == ranks


def test_oracle_topk_matches_reference_golden(retrieval_golden):
    g = retrieval_golden
    Q, G = torch.from_numpy(g["Q"]), torch.from_numpy(g["G"])
    vals, idx = O.pairwise_topk_batched(Q, G, 10, g["loss_type"])
    assert torch.equal(idx, torch.from_numpy(g["top_idx"]))
    assert torch.equal(vals, torch.from_numpy(g["top_val"]))  # same ATen ops → bit-equal
    image_paths = _paths(g["image_paths"])
    tk = O.get_topk_images(10, image_paths, Q[:1], G, g["loss_type"])
    assert [p for p, _ in tk] == [str(image_paths[i]) for i in g["top_idx"][0]]


@pytest.mark.parametrize("loss_type", ["euclidean", "cosine"])
def test_oracle_process_inference_matches_reference_golden(golden_dir, loss_type):
    z = np.load(golden_dir / f"retrieval_{loss_type}.npz")
    ref = json.load(open(golden_dir / f"process_inference_{loss_type}.json"))
    Q, G = torch.from_numpy(z["Q"]), torch.from_numpy(z["G"])
    image_paths, sketch_paths = _paths(z["image_paths"]), _paths(z["sketch_paths"])
    pos = [O.find_image_index(image_paths, O.sketch_name_to_key(s, image_paths)) for s in sketch_paths]
    got = O.process_inference(Q, G, pos, loss_type)
    for key in ("mean_reciprocal_rank", "size", "count", "mean", "std", "min", "25%", "50%", "75%", "max"):
        assert got[key] == pytest.approx(ref[key], rel=1e-12), key
    assert got["topk_acc"] == pytest.approx(ref["topk_acc"], rel=1e-12)


def test_oracle_triplet_matches_reference_golden(triplet_golden):
    t = triplet_golden
    a, p, n = (torch.from_numpy(t[k]) for k in ("a", "p", "n"))
    assert float(t["margin"]) == pytest.approx(O.MARGIN)
    m = O.MARGIN

    def run(fn):
        A, P, N = (x.clone().requires_grad_(True) for x in (a, p, n))
        loss = fn(A, P, N)
        loss.backward()
        return loss.detach().numpy(), A.grad.numpy(), P.grad.numpy(), N.grad.numpy()

    cs, cp, cs2, cp2 = (torch.from_numpy(t[k]) for k in ("cs", "cp", "cs2", "cp2"))
    l1, l2 = torch.from_numpy(t["l1"]), torch.from_numpy(t["l2"])
    cases = {
        "tml_euclid": lambda A, P, N: O.triplet_margin_loss(A, P, N, m, "euclidean"),
        "tmdl_cosine": lambda A, P, N: O.triplet_margin_loss(A, P, N, m, "cosine"),
        "wc_euclid": lambda A, P, N: O.TripletMarginLoss_with_classification(margin=m)(A, P, N, cs, cp, l1),
        "wc_cosine": lambda A, P, N: O.TripletMarginLoss_with_classification(margin=m, distance_f=O.cosine_distance)(A, P, N, cs, cp, l1),
        "wc2_euclid": lambda A, P, N: O.TripletMarginLoss_with_classification2(margin=m, classification_weight=0, classification_weight2=0.2)(A, P, N, cs, cp, cs2, cp2, l1, l2),
    }
    for name, fn in cases.items():
        loss, ga, gp, gn = run(fn)
        assert np.array_equal(loss, t[name + "_loss"]), name
        for got, key in ((ga, "_ga"), (gp, "_gp"), (gn, "_gn")):
            assert np.array_equal(got, t[name + key]), name + key
    # nn.TripletMarginLoss ≡ TripletMarginWithDistanceLoss(PairwiseDistance)
    assert np.array_equal(t["tml_euclid_loss"], t["tmdl_euclid_loss"])


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout only exists in the build container")
def test_oracle_matches_live_reference():
    ref_utils, ref_inference = ref_import.load()
    Q, G, pos = O.synthetic_embeddings(20, 300, 128, seed=7, beta=0.3, num_classes=6)
    paths = [Path(f"photos/n{i:05d}.jpg") for i in range(300)]
    for lt in ("euclidean", "cosine"):
        for i in range(20):
            sk = Path(f"sketches/n{int(pos[i]):05d}-1.png")
            assert ref_inference.get_ranking_position(sk, paths, Q[i:i + 1], G, lt) == \
                O.get_ranking_position(sk, paths, Q[i:i + 1], G, lt)
            assert ref_inference.get_topk_images(7, paths, Q[i:i + 1], G, lt) == O.get_topk_images(7, paths, Q[i:i + 1], G, lt)
    assert torch.equal(ref_utils.euclidean_distance(Q[:1], G), O.euclidean_distance(Q[:1], G))
    assert torch.equal(ref_utils.cosine_distance(Q[:1], G), O.cosine_distance(Q[:1], G))
    with pytest.raises(Exception, match="loss type not correct"):
        O.distances(Q[:1], G, "manhattan")
    with pytest.raises(Exception, match="loss type not correct"):
        ref_inference.get_ranking_position(Path("sketches/n00001-1.png"), paths, Q[:1], G, "manhattan")


def test_pairwise_distance_semantics():
    # F2: nn.PairwiseDistance is ||x1 - x2 + 1e-6||, not the plain norm
    torch.manual_seed(0)
    q, G = torch.randn(1, 256), torch.randn(50, 256)
    d = O.euclidean_distance(q, G)
    assert torch.equal(d, (q - G + 1e-6).norm(dim=1))
    assert np.allclose(d.numpy(), O.distances_fp64(q.numpy(), G.numpy(), "euclidean"), rtol=2e-6)
    # H2: per-operand clamp of the norms at 1e-8 (not of their product)
    tiny = torch.full((1, 4), 1e-10)
    one = torch.ones(1, 4)
    assert O.cosine_distance(tiny, one).item() == pytest.approx(1 - 0.02, abs=1e-3)
    assert O.cosine_distance(torch.zeros(1, 4), one).item() == 1.0
    c = O.cosine_distance(q, G)
    assert np.allclose(c.numpy(), O.distances_fp64(q.numpy(), G.numpy(), "cosine"), atol=5e-7)
    # F8: an fp64 gallery (CSV-loaded) promotes the computation to fp64
    assert O.euclidean_distance(q, G.double()).dtype == torch.float64


def test_fp32_and_fp64_rankings_agree_on_generator_data():
    Q, G, pos = O.synthetic_embeddings(64, 2000, 512, seed=1234)
    for lt in ("euclidean", "cosine"):
        r32 = O.rank_of_positive_batched(Q, G, pos, lt)
        r64 = O.rank_of_positive_batched(Q, G, pos, lt, fp64=True)
        assert torch.equal(r32, r64)
        _, i32 = O.pairwise_topk_batched(Q, G, 10, lt)
        _, i64 = O.pairwise_topk_batched(Q, G, 10, lt, fp64=True)
        assert torch.equal(i32, i64)
    # the cdist/topk restatement of north_star orders identically on this data
    _, ic = O.cdist_topk(Q, G, 10)
    assert torch.equal(ic, O.pairwise_topk_batched(Q, G, 10, "euclidean")[1])


def test_rank_edge_cases():
    G = torch.tensor([[0.0, 0.0], [1.0, 0.0], [1.0, 0.0], [3.0, 0.0]])
    q = torch.tensor([[0.9, 0.0]])
    assert O.ranking_position(q, G, -1, "euclidean") == 4          # missing positive → len(G)
    assert O.ranking_position(q, G, 3, "euclidean") == 3
    assert O.ranking_position(q, G, 0, "euclidean") == 2
    # duplicate rows tie: the reference's position is one of the tied slots, count-less-than is the first
    assert O.ranking_position(q, G, 1, "euclidean") in (0, 1)
    assert O.rank_of_positive_batched(q, G, torch.tensor([2]), "euclidean").tolist() == [0]
    m = O.retrieval_metrics([0, 0, 3, 12], k=10)
    assert m["topk_acc"][0] == 0.5 and m["topk_acc"][3] == 0.75 and m["topk_acc"][9] == 0.75
    assert m["mean_reciprocal_rank"] == pytest.approx((1 + 1 + 0.25 + 1 / 13) / 4)
    assert m["max"] == 13.0 and m["50%"] == 2.5


def test_name_parsing_conventions():
    photos = [Path("data/sketchy/photos/n01-5.jpg"), Path("data/sketchy/photos/n02.jpg")]
    assert O.sketch_name_to_key("s/n02-7.png", photos) == "n02"            # id-number.png
    assert O.sketch_name_to_key("s/n02.png", photos) == "n02"              # kaggle id.png
    assert O.sketch_name_to_key("s/12-n02-99887.png", photos) == "n02"     # index-id-random.png
    art = [Path("data/artworks/a-b.jpg")]
    assert O.sketch_name_to_key("s/a-b.png", art) == "a-b"                 # artworks keep the full stem
    assert O.find_image_index(photos, "n02") == 1 and O.find_image_index(photos, "zz") == -1


def test_batch_hard_definition():
    torch.manual_seed(3)
    a, p, n = torch.randn(6, 16), torch.randn(6, 16), torch.randn(6, 16)
    loss, hpi, hni = O.batch_hard_triplet_loss(a, p, n, 0.2, "euclidean")
    D = torch.cdist(a, torch.cat([p, n]))
    assert hpi.tolist() == list(range(6))                  # only X_i is positive for anchor i
    for i in range(6):
        row = D[i].clone()
        row[i] = float("inf")
        assert hni[i].item() == row.argmin().item()
    labels = torch.tensor([0, 0, 1, 1, 2, 2])
    loss_l, hpi_l, _ = O.batch_hard_triplet_loss(a, p, n, 0.2, "euclidean", labels)
    for i in range(6):
        assert hpi_l[i].item() in (i - (i % 2), i - (i % 2) + 1)
    assert loss_l.item() >= loss.item() - 1e-6             # a wider positive set cannot make hp smaller


def test_l2_normalize_matches_cosine_similarity_normalisation():
    x = torch.randn(10, 33)
    x[3] = 0
    y = O.l2_normalize(x)
    assert torch.allclose(y, torch.nn.functional.normalize(x, dim=1, eps=1e-8))
    assert torch.equal(y[3], torch.zeros(33))
    assert torch.allclose(1 - (O.l2_normalize(x[:1]) * O.l2_normalize(x)).sum(1), O.cosine_distance(x[:1], x), atol=1e-6)


def test_generator_is_seeded_and_nontrivial():
    Q1, G1, p1 = O.synthetic_embeddings(100, 1000, 512, seed=1234)
    Q2, G2, p2 = O.synthetic_embeddings(100, 1000, 512, seed=1234)
    assert torch.equal(Q1, Q2) and torch.equal(G1, G2) and torch.equal(p1, p2)
    r = O.rank_of_positive_batched(Q1, G1, p1, "euclidean")
    assert 0.2 < (r < 1).float().mean() < 0.95 and (r < 10).float().mean() > 0.6
