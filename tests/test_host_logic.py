"""CPU suite, part 3: host-side mirror of the reference interface (names, attributes, error
behaviour, file-name conventions, feature store) — everything that does not need a kernel."""
from pathlib import Path

import numpy as np
import pytest
import torch

from art_sbir_b200 import inference as inf
from art_sbir_b200 import sharded, utils as U
from oracle import sbir_oracle as O


def test_loss_module_surface_matches_reference():
    assert U.MARGIN == O.MARGIN == 0.2
    assert (U.euclidean_distance.norm, U.euclidean_distance.eps, U.euclidean_distance.keepdim) == (2.0, 1e-6, False)
    wc = U.TripletMarginLoss_with_classification(margin=0.2)
    assert (wc.margin, wc.classification_weight, wc.classification_weight2) == (0.2, 0.5, 0)
    wc2 = U.TripletMarginLoss_with_classification2(margin=0.3, classification_weight=0, classification_weight2=0.2)
    assert (wc2.margin, wc2.classification_weight, wc2.classification_weight2) == (0.3, 0, 0.2)
    # train.py:164-175 dispatch
    assert isinstance(U.make_loss_fn("euclidean", False), U.TripletMarginLoss)
    assert isinstance(U.make_loss_fn("cosine", False), U.TripletMarginWithDistanceLoss)
    assert U.make_loss_fn("euclidean", True, "MixedDatasetV2").classification_weight == 0.01
    assert U.make_loss_fn("cosine", True, "MixedDatasetV2").classification_weight == 0.5
    k = U.make_loss_fn("euclidean", True, "KaggleDatasetV2")
    assert isinstance(k, U.TripletMarginLoss_with_classification2) and (k.classification_weight, k.classification_weight2) == (0, 0.2)
    assert U.make_loss_fn("euclidean", True, "SketchyV2").margin == 0.2       # param_dict reads loss_fn.margin (train.py:178)
    with pytest.raises(Exception, match="loss type not correct"):
        U.make_loss_fn("manhattan", False)
    with pytest.raises(TypeError):
        U.TripletMarginLoss_with_classification(margin=0.2, distance_f=lambda a, b: (a - b).abs().sum(1))


def test_name_parsing_and_positive_lookup_match_oracle():
    photos = sorted(Path(f"data/sketchy/photos/n{i:04d}.jpg") for i in range(50))
    sketches = ["s/n0007-3.png", "s/n0049.png", "s/12-n0003-99887.png", "s/n9999-1.png"]
    got = inf.positive_indices(sketches, photos, verbose=False).tolist()
    want = [O.find_image_index(photos, O.sketch_name_to_key(s, photos)) for s in sketches]
    assert got == want == [7, 49, 3, -1]
    art = [Path("data/artworks/a-b.jpg"), Path("data/artworks/c.jpg")]
    assert inf.sketch_key("s/a-b.png", art) == O.sketch_name_to_key("s/a-b.png", art) == "a-b"
    # first match wins, like the reference's linear scan (utils.py:22-25)
    dup = [Path("x/a.jpg"), Path("y/a.jpg")]
    assert U.find_image_index(dup, "a") == 0


def test_inference_dataset_dedups_and_sorts():
    ds = inf.InferenceDataset([Path("b.jpg"), Path("a.jpg"), Path("b.jpg")])
    assert ds.image_paths == [Path("a.jpg"), Path("b.jpg")] and len(ds) == 2


def test_bad_loss_type_raises_like_reference():
    with pytest.raises(Exception, match="loss type not correct"):
        inf.get_ranking_position("s/n0001-1.png", [Path("p/n0001.jpg")], torch.zeros(1, 8), torch.zeros(1, 8), "manhattan")


def test_missing_positive_short_circuits_without_gpu(capsys):
    # inference.py:39-41: returns len(image_paths) and prints, before any distance is computed
    r = inf.get_ranking_position("s/zzz-1.png", [Path("p/n0001.jpg"), Path("p/n0002.jpg")], torch.zeros(1, 8), torch.zeros(2, 8), "euclidean")
    assert r == 2 and "No image found" in capsys.readouterr().out


def test_feature_store_roundtrip(tmp_path):
    class DS:
        image_paths = [Path("p/a.jpg"), Path("p/b.jpg"), Path("p/c.jpg")]
    feats = torch.randn(3, 16)
    name = U.save_image_features("ModifiedResNet", "SketchyV1", DS(), feats, root=tmp_path)
    assert name.startswith("ModifiedResNet_SketchyV1_")
    paths, loaded = U.load_image_features(name, root=tmp_path)
    assert paths == DS.image_paths and torch.equal(loaded, feats)           # sidecar: exact fp32
    (tmp_path / name / "image_features.f32.npy").unlink()
    paths, loaded = U.load_image_features(name, root=tmp_path)              # reference CSV route → float64 (F8)
    assert loaded.dtype == torch.float64 and torch.allclose(loaded.float(), feats, rtol=1e-6)


def test_feature_store_keeps_bf16_rows_and_norms(tmp_path):
    """N2: a bf16 gallery is stored and reloaded in its own type, with its stored ‖row‖² beside it; the
    reference's CSV pair is still written (and still what the reference itself would read)."""
    class DS:
        image_paths = [Path(f"p/{c}.jpg") for c in "abcd"]
    feats = torch.randn(4, 16).bfloat16()
    sq = (feats.float() ** 2).sum(1)
    name = U.save_image_features("ModifiedResNet", "KaggleV2", DS(), U.GalleryFeatures(feats, sq), root=tmp_path)
    paths, loaded = U.load_image_features(name, root=tmp_path)
    assert paths == DS.image_paths and loaded.dtype == torch.bfloat16 and torch.equal(loaded, feats)
    paths, gal = U.load_image_features(name, root=tmp_path, with_norms=True)
    assert isinstance(gal, U.GalleryFeatures) and torch.equal(gal.rows, feats) and torch.equal(gal.sqnorm, sq)
    assert gal.shape == (4, 16) and len(gal) == 4
    assert gal.to(torch.float32).sqnorm is None                     # norms belong to the rows as stored
    for side in ("image_features.bf16.npy", "image_sqnorm.f32.npy"):
        (tmp_path / name / side).unlink()
    paths, gal = U.load_image_features(name, root=tmp_path, with_norms=True)        # CSV route (F8): float64, no norms
    assert gal.rows.dtype == torch.float64 and gal.sqnorm is None and torch.equal(gal.rows.float(), feats.float())


def test_shard_bounds_partition_the_gallery():
    for n, w in ((10, 3), (75000, 8), (7, 8), (0, 2)):
        spans = [sharded.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_weighted_shard_bounds_follow_the_weights():
    """Speed-weighted sharding (sharded.weighted_shard_bounds): a partition of [0, N) in rank order whose shares
    follow the weights to within the alignment; equal weights reproduce the even split."""
    for n, w, align in ((10_000_000, [1.0, 1.1, 0.9, 1.0, 1.05, 0.95, 1.0, 1.0], 256), (101, [3, 1, 0], 1), (7, [1] * 8, 1),
                        (1000, [1e-3, 1e3], 16), (0, [1, 2], 1)):
        spans = [sharded.weighted_shard_bounds(n, w, r, align) for r in range(len(w))]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] and spans[i][0] <= spans[i][1] for i in range(len(w) - 1))
        tot = float(sum(w))
        for (a, b), wi in zip(spans, w):
            assert abs((b - a) - n * wi / tot) <= align + 1
    even = [sharded.weighted_shard_bounds(75_000, [2.5] * 8, r) for r in range(8)]
    assert even == [sharded.shard_bounds(75_000, 8, r) for r in range(8)]
    with pytest.raises(ValueError):
        sharded.weighted_shard_bounds(10, [0, 0], 0)
    with pytest.raises(ValueError):
        sharded.weighted_shard_bounds(10, [1, 1], 2)
    assert sharded.rank_speed_weights(1000, 4.0) == [250.0]      # no process group: this rank alone
    # damped re-cut towards equal kernel times: the slow rank gives rows away, less than the full correction would take
    rows, ms = [1000, 1000, 1000, 1000], [10.0, 10.0, 10.0, 12.0]
    w = sharded.rebalanced_weights(rows, ms, damping=0.6)
    full = sharded.rebalanced_weights(rows, ms, damping=1.0)
    assert w[3] < w[0] == w[1] == w[2] and full[3] < w[3] < rows[3]
    cut = [sharded.weighted_shard_bounds(4000, w, r) for r in range(4)]
    assert cut[-1][1] == 4000 and (cut[3][1] - cut[3][0]) < 1000 < (cut[0][1] - cut[0][0])
    assert sharded.rebalanced_weights([5, 7], [3.0, 3.0]) == [5.0, 7.0]      # equal times: nothing moves
    with pytest.raises(ValueError):
        sharded.rebalanced_weights([1, 2], [1.0])


def _plan(lib, nq, ng, d, k, dtype, sms=148):
    import ctypes
    out = (ctypes.c_int32 * 13)()
    assert lib.sbir_debug_plan(nq, ng, d, k, dtype, sms, out) == 0
    keys = ("cap", "lists", "q_tiles", "g_tiles", "parts", "tiles_per_part", "chunks", "tiles_per_chunk", "units",
            "part_fastest", "pair", "q_tile_stride", "tile_bf16")
    return dict(zip(keys, list(out)))


@pytest.mark.parametrize("sel_bf16", [-1, 0, 1])
@pytest.mark.parametrize("nq,ng,d,k,dtype", [
    (100_000, 10_000_000, 512, 10, 1), (12_500, 75_000, 2048, 100, 0), (12_500, 75_000, 2048, 10, 0),
    (1_000, 10_000, 2048, 10, 0), (1, 513, 1024, 10, 0), (300, 9001, 512, 10, 1), (5, 7, 64, 3, 0),
    (257, 3001, 512, 50, 1), (100_000, 1_250_000, 512, 10, 1), (40, 900, 36, 5, 0)])
def test_k1_plan_covers_every_tile_exactly_once(sbir_lib, nq, ng, d, k, dtype, sel_bf16):
    """The unit grid (query tile x partition x chunk) must tile the Q x G problem exactly: every
    (query tile, gallery tile) pair belongs to one unit, chunks of a partition are contiguous and
    ordered, capacities hold k plus slack, and the gallery chunk of a unit stays L2-sized.  fp32
    embeddings are planned for their bf16 selection copies (kind::f16 tiles) unless option
    k1_sel_bf16 = 0 keeps them on kind::tf32 (or their rows do not give bf16 copies a 16-byte pitch)."""
    from art_sbir_b200 import _binding
    try:
        _binding.set_debug_option("k1_sel_bf16", sel_bf16)
        p = _plan(sbir_lib, nq, ng, d, k, dtype)
        ws = sbir_lib.sbir_pairwise_topk_workspace_bytes(nq, ng, d, k, dtype, 0, 1)
    finally:
        _binding.set_debug_option("reset")
    # element type the tensor-core tiles read: fp32 embeddings go through bf16 selection copies when forced, or
    # (auto) when the problem is not launch-bound (kind::tf32 tiers behind the pass catch what its wider band cannot certify)
    auto = 2.0 * d * nq * ng >= 4e11
    tiles_bf16 = dtype == 1 or (d % 8 == 0 and (sel_bf16 == 1 or (sel_bf16 == -1 and auto)))
    assert p["tile_bf16"] == int(tiles_bf16)
    assert p["q_tiles"] == -(-nq // 128) and p["g_tiles"] == -(-ng // 256)
    slack = 6 if dtype == 1 else 16
    assert p["cap"] in (16, 32, 64, 128) and p["cap"] >= min(k + slack, 128) and p["cap"] * p["lists"] * p["parts"] <= 4096
    assert p["q_tile_stride"] >= p["q_tiles"] and p["q_tile_stride"] % 2 == 0
    # rows of the unit grid: query tiles, or PAIRS of them when the plan uses CTA pairs
    # (chosen for kind::tf32 tiles with the 64/128-entry lists of large k, and for bf16 tiles of rows of 4 KB and more —
    # the all-shared-memory form is L2-bound there and a pair moves a third fewer operand bytes)
    want_pair = (not tiles_bf16 and p["cap"] >= 64) or (tiles_bf16 and d * 2 >= 4096)
    assert p["pair"] == (2 if (want_pair and p["q_tiles"] >= 2) else 1)
    assert p["units"] == p["chunks"] * p["parts"] * -(-p["q_tiles"] // p["pair"])
    covered = []
    for part in range(p["parts"]):
        b = part * p["tiles_per_part"]
        e = min(b + p["tiles_per_part"], p["g_tiles"])
        for c in range(p["chunks"]):
            t0 = min(b + c * p["tiles_per_chunk"], e)
            t1 = min(t0 + p["tiles_per_chunk"], e)
            covered.extend(range(t0, t1))
    assert covered == list(range(p["g_tiles"]))                       # each gallery tile once, in order
    es = 2 if tiles_bf16 else 4
    # ~12 MB chunks; 48 MB for the resident-query form (bf16 tiles of rows of at most 1 KB: only gallery rows go through L2)
    # (96 MB with the 64/128-entry lists of large k: fewer, larger list hand-overs)
    limit = ((97 << 20) if p["cap"] >= 64 else (49 << 20)) if (tiles_bf16 and d * 2 <= 1024) else (13 << 20)
    assert p["tiles_per_chunk"] == 1 or p["tiles_per_chunk"] * 256 * d * es <= limit
    # few query tiles -> partitions supply the parallelism; many -> a single partition
    if p["q_tiles"] >= 2 * 148:
        assert p["parts"] == 1
    assert ws > 0 and ws >= p["parts"] * p["q_tile_stride"] * p["lists"] * p["cap"] * 128 * 8
    if dtype == 0 and tiles_bf16:
        assert ws >= (nq + ng) * d * 2                                # room for the bf16 selection copies


def test_bf16_selection_error_bound_from_residual_norms():
    """The certificate for fp32 embeddings selected on bf16 copies (common.cuh: e_margin) bounds
    |q·g − qh·gh| by ‖q‖·‖gl‖ + ‖ql‖·‖g‖ (+ ‖ql‖‖gl‖) with the MEASURED residual norms: checked here in fp64 on
    zero-mean, all-positive and wide-dynamic-range data.  It is about the size of the element-wise worst case
    2^-8·‖q‖‖g‖ (bf16 rounds each operand by at most 2^-9), i.e. roughly twice kind::tf32's band."""
    g = torch.Generator().manual_seed(1)
    for make in (lambda n: torch.randn(n, 512, generator=g), lambda n: torch.rand(n, 512, generator=g) * 3 + 1,
                 lambda n: torch.randn(n, 512, generator=g) * torch.logspace(-3, 3, 512)):
        Q, G = make(64).double(), make(512).double()
        Qh, Gh = Q.float().bfloat16().double(), G.float().bfloat16().double()
        err = (Q @ G.T - Qh @ Gh.T).abs()
        qn, gn = Q.norm(dim=1), G.norm(dim=1)
        qr, gr = (Q - Qh).norm(dim=1), (G - Gh).norm(dim=1)
        bound = 1.01 * qn[:, None] * gr.max() + qr[:, None] * gn.max() + qr[:, None] * gr.max()
        assert (err <= bound).all()
        assert 0.5 < (bound / (2.0 ** -8 * qn[:, None] * gn.max())).max() < 1.25   # vs the element-wise worst case


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's CPU arm: the oracle port of the reference's per-query loop on the host cores)
    prints ONE JSON line with the same metric / unit / config as the b200 arm, `impl: reference`, a cpu_baseline describing
    the run and an e2e object with zero copied bytes — and needs no GPU and no CUDA library."""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "cfg4" and d["config"]["num_g"] == 10_000_000 and d["config"]["k"] == 10
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
