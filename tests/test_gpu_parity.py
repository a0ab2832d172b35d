"""GPU suite (-m gpu): the CUDA path, called through the C ABI (art_sbir_b200.ops → ctypes →
libsbir_b200.so), against (1) the golden fixtures minted from the reference, (2) the oracle on
seeded inputs at sizes the oracle finishes in seconds, (3) size-independent properties at the
BASELINE.json sizes.  Tolerances are north_star's: recall@K / ranks bit-exact, ranked indices
identical except at distance ties within 1e-4 relative, distances and losses within 1e-3
relative in fp32 (bf16 inputs: oracle = fp32 math on the bf16-rounded inputs)."""
import ctypes
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sbir_oracle as O

pytestmark = pytest.mark.gpu

DIST_RTOL = 1e-3   # north_star: distances / loss within 1e-3 relative (fp32)
TIE_RTOL = 1e-4    # north_star: index swaps allowed only between distances closer than this


@pytest.fixture(scope="module")
def ops(sbir_lib):
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    assert sbir_lib.sbir_device_supported() == 1, "not an sm_100 device"
    from art_sbir_b200 import ops as _ops
    return _ops


@pytest.fixture
def dbg(sbir_lib):
    """Library tuning / test switches (sbir_debug_set_option), restored to their defaults afterwards."""
    from art_sbir_b200 import _binding
    yield _binding.set_debug_option
    _binding.set_debug_option("reset")


def assert_topk_matches(vals, idx, ref_vals, ref_idx, dist_rows):
    """dist_rows[i] = oracle distances of query i to every gallery row (for tie analysis)."""
    vals, idx = vals.cpu(), idx.cpu()
    assert vals.shape == ref_vals.shape and idx.shape == ref_idx.shape
    assert torch.allclose(vals, ref_vals.float(), rtol=DIST_RTOL, atol=1e-6)
    bad = (idx != ref_idx).nonzero()
    for i, j in bad.tolist():
        ours = dist_rows[i][idx[i, j]].item()
        theirs = ref_vals[i, j].item()
        assert abs(ours - theirs) <= TIE_RTOL * max(abs(theirs), 1e-6), (i, j, ours, theirs)
    return len(bad)


# What can separate two correct fp32 evaluations of the same distance formula: torch's own fp32 result
# deviates from the fp64 evaluation by up to 5·2^-24 relative (measured: 64..2048-d N(0,1) and
# cancellation-heavy data), a comparison d < d_pos involves two such values → 12·2^-24 ≈ 7e-7,
# 140× tighter than north_star's 1e-4 tie tolerance.
FP32_ULPS = 12 * 2.0 ** -24


def assert_ranks_match_up_to_fp32_ties(got, Qo, Go, pos, lt, ref_r32=None):
    """Rank parity with PROOF that every deviation from the torch-fp32 oracle is a tie at fp32 rounding level.

    For every query the fp64 evaluation of the reference formula gives the distances d64; a gallery row can
    only be counted differently by two correct fp32 evaluations if |d − d_pos| is within a few fp32 ulps of
    the quantity that was rounded (the distance itself for the euclidean metric, the similarity ≈ 1 for
    the cosine metric).  So rank ∈ [#{d64 < d_pos − τ}, #{d64 ≤ d_pos + τ} − 1] must hold for OUR rank and
    for the oracle's, and wherever that interval is a single value all three agree exactly.
    Returns the number of queries whose rank differs from the torch-fp32 oracle (all of them proven ties)."""
    got = got.cpu()
    if ref_r32 is None:
        ref_r32 = O.rank_of_positive_batched(Qo, Go, pos, lt)
    differing = 0
    for i in range(Qo.shape[0]):
        pi = int(pos[i])
        if pi < 0:
            assert int(got[i]) == Go.shape[0] == int(ref_r32[i])
            continue
        d64 = O.distances(Qo[i:i + 1].double(), Go.double(), lt)
        dp = d64[pi].item()
        tau = FP32_ULPS * (max(dp, 1e-30) if lt == "euclidean" else 1.0)
        lo = int((d64 < dp - tau).sum())
        hi = int((d64 <= dp + tau).sum()) - 1          # the positive itself is inside the band
        assert lo <= int(got[i]) <= hi, (i, lo, int(got[i]), hi)
        assert lo <= int(ref_r32[i]) <= hi, (i, lo, int(ref_r32[i]), hi)
        differing += int(got[i]) != int(ref_r32[i])
    return differing


# --------------------------------------------------------------------- golden fixtures ----
def test_golden_ranks_and_topk(ops, retrieval_golden):
    g = retrieval_golden
    lt = g["loss_type"]
    Q, G = torch.from_numpy(g["Q"]), torch.from_numpy(g["G"])
    from art_sbir_b200 import inference as inf
    image_paths = [Path(str(p)) for p in g["image_paths"]]
    sketch_paths = [Path(str(p)) for p in g["sketch_paths"]]
    pos = inf.positive_indices(sketch_paths, image_paths, verbose=False)
    vals, idx, rank, unc = ops.pairwise_topk(Q.cuda(), G.cuda(), 10, lt, pos_index=pos.cuda(), return_uncertified=True)
    assert rank.cpu().tolist() == g["ranks"].tolist()                       # bit-exact ranks → recall@K
    dist_rows = [O.distances(Q[i:i + 1], G, lt) for i in range(len(Q))]
    swaps = assert_topk_matches(vals, idx, torch.from_numpy(g["top_val"]), torch.from_numpy(g["top_idx"]), dist_rows)
    assert swaps == 0                                                        # this fixture is tie-free
    # the per-query drop-in functions (inference.py:30-69) give the same answers
    for i in (0, 5, 17):
        assert inf.get_ranking_position(sketch_paths[i], image_paths, Q[i:i + 1], G, lt) == int(g["ranks"][i])
    tk = inf.get_topk_images(10, image_paths, Q[2:3], G, lt)
    assert [p for p, _ in tk] == [str(image_paths[j]) for j in g["top_idx"][2]]
    assert np.allclose([v for _, v in tk], g["top_val"][2], rtol=DIST_RTOL)


@pytest.mark.parametrize("loss_type", ["euclidean", "cosine"])
def test_golden_process_inference(ops, golden_dir, loss_type):
    from art_sbir_b200 import inference as inf
    z = np.load(golden_dir / f"retrieval_{loss_type}.npz")
    ref = json.load(open(golden_dir / f"process_inference_{loss_type}.json"))
    Q, G = torch.from_numpy(z["Q"]), torch.from_numpy(z["G"])

    class DS:
        sketch_paths = [Path(str(p)) for p in z["sketch_paths"]]

        def __len__(self):
            return len(self.sketch_paths)

    class InfDS:
        image_paths = [Path(str(p)) for p in z["image_paths"]]

        def __len__(self):
            return len(self.image_paths)

    loader = [(Q[i:i + 16],) for i in range(0, len(Q), 16)]     # batched queries (N4)
    from timeit import default_timer as timer
    got = inf.process_inference(torch.nn.Identity(), DS(), InfDS(), loader, G, timer(), False, loss_type)
    assert set(got) == set(ref) | {"inference_time"}
    for key in ("size", "count", "min", "25%", "50%", "75%", "max"):
        assert got[key] == ref[key], key
    assert got["topk_acc"] == ref["topk_acc"]                                   # recall@1..10 bit-exact
    assert got["mean_reciprocal_rank"] == pytest.approx(ref["mean_reciprocal_rank"], rel=1e-12)
    assert got["mean"] == pytest.approx(ref["mean"], rel=1e-12) and got["std"] == pytest.approx(ref["std"], rel=1e-12)
    assert len(got["retrieval_samples"]) == len(ref["retrieval_samples"])
    for a, b in zip(got["retrieval_samples"], ref["retrieval_samples"]):
        assert list(a) == list(b)
        (ka, va), (kb, vb) = next(iter(a.items())), next(iter(b.items()))
        assert [p for p, _ in va] == [p for p, _ in vb]
        assert np.allclose([d for _, d in va], [d for _, d in vb], rtol=DIST_RTOL)
    json.dumps(got)                                                             # JSON-serialisable like the reference's


def test_golden_triplet_losses(ops, triplet_golden):
    from art_sbir_b200 import utils as U
    t = triplet_golden
    m = O.MARGIN
    dev = "cuda"
    cs, cp, cs2, cp2 = (torch.from_numpy(t[k]).to(dev) for k in ("cs", "cp", "cs2", "cp2"))
    l1, l2 = torch.from_numpy(t["l1"]).to(dev), torch.from_numpy(t["l2"]).to(dev)
    cases = {
        "tml_euclid": lambda A, P, N: U.TripletMarginLoss(margin=m)(A, P, N),
        "tmdl_cosine": lambda A, P, N: U.TripletMarginWithDistanceLoss(margin=m, distance_function=U.cosine_distance)(A, P, N),
        "tmdl_euclid": lambda A, P, N: U.TripletMarginWithDistanceLoss(margin=m, distance_function=U.euclidean_distance)(A, P, N),
        "wc_euclid": lambda A, P, N: U.TripletMarginLoss_with_classification(margin=m)(A, P, N, cs, cp, l1),
        "wc_cosine": lambda A, P, N: U.TripletMarginLoss_with_classification(margin=m, distance_f=U.cosine_distance)(A, P, N, cs, cp, l1),
        "wc2_euclid": lambda A, P, N: U.TripletMarginLoss_with_classification2(margin=m, classification_weight=0, classification_weight2=0.2)(A, P, N, cs, cp, cs2, cp2, l1, l2),
    }
    for name, fn in cases.items():
        A, P, N = (torch.from_numpy(t[k]).to(dev).requires_grad_(True) for k in ("a", "p", "n"))
        loss = fn(A, P, N)
        loss.backward()
        assert loss.item() == pytest.approx(float(t[name + "_loss"]), rel=DIST_RTOL), name
        for got, key in ((A.grad, "_ga"), (P.grad, "_gp"), (N.grad, "_gn")):
            ref = torch.from_numpy(t[name + key])
            assert torch.allclose(got.cpu(), ref, rtol=DIST_RTOL, atol=1e-6 * ref.abs().max().item()), name + key


# ------------------------------------------------------------- oracle on seeded inputs ----
SHAPES = [  # (Q, N, D, dtype, loss, k) — Q, N not multiples of the 128×256 tile; D in {64..2048}; k in {1,10,100}
    (37, 300, 64, "float32", "euclidean", 10),
    (130, 1000, 512, "float32", "euclidean", 1),
    (200, 2500, 1024, "float32", "cosine", 10),
    (64, 1500, 2048, "float32", "euclidean", 100),
    (257, 3001, 512, "bfloat16", "euclidean", 10),
    (100, 777, 256, "bfloat16", "cosine", 100),
    (1, 513, 1024, "float32", "euclidean", 10),
    (300, 5000, 96, "float32", "euclidean", 30),
]


@pytest.mark.parametrize("nq,ng,d,dtype,lt,k", SHAPES)
def test_topk_and_rank_match_oracle(ops, nq, ng, d, dtype, lt, k):
    Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=nq + ng, beta=0.3 if d < 512 else None, num_classes=max(4, ng // 80))
    pos[::9] = -1
    tdt = getattr(torch, dtype)
    Qd, Gd = Q.to(tdt), G.to(tdt)
    Qo, Go = Qd.float(), Gd.float()            # oracle = fp32 math on the (possibly bf16-rounded) inputs
    vals, idx, rank, unc = ops.pairwise_topk(Qd.cuda(), Gd.cuda(), k, lt, pos_index=pos.cuda(), return_uncertified=True)
    ref_v, ref_i = O.pairwise_topk_batched(Qo, Go, k, lt)
    dist_rows = [O.distances(Qo[i:i + 1], Go, lt) for i in range(nq)]
    assert_topk_matches(vals, idx, ref_v, ref_i, dist_rows)
    ref_r64 = O.rank_of_positive_batched(Qo, Go, pos, lt, fp64=True)
    ref_r32 = O.rank_of_positive_batched(Qo, Go, pos, lt)
    got = rank.cpu()
    # exact wherever the oracle itself is unambiguous (fp32 and fp64 evaluation agree)
    same = ref_r64 == ref_r32
    assert torch.equal(got[same], ref_r32[same])
    assert ((got == ref_r64) | (got == ref_r32)).all()
    assert (got[pos < 0] == ng).all()


@pytest.mark.parametrize("nq,ng,d,dtype,lt,k", [(40, 900, 66, "float32", "euclidean", 10), (3, 5, 66, "float32", "euclidean", 3),
                                                 (130, 2000, 7, "float32", "cosine", 5), (33, 500, 50, "bfloat16", "euclidean", 10),
                                                 (200, 3000, 1001, "float32", "euclidean", 20), (64, 700, 333, "bfloat16", "cosine", 100)])
def test_rows_of_any_width(ops, nq, ng, d, dtype, lt, k):
    """The reference accepts any embedding width (its distance modules are plain torch ops).  Rows whose byte length is
    not a multiple of 16 cannot be read by TMA in place: the tensor-core tiles read zero-padded copies kept in the
    workspace while every exact kernel reads the caller's rows.  Also a gallery VIEW that starts one (odd-width) row in."""
    Q, G, pos = O.synthetic_embeddings(nq, ng + 1, d, seed=nq + d, beta=0.3, num_classes=max(4, ng // 80))
    tdt = getattr(torch, dtype)
    Qd, Gd = Q.to(tdt), G.to(tdt)
    G_view = Gd.cuda()[1:]                                  # base pointer offset by one row: not 16-byte aligned for these widths
    pos = (pos - 1).clamp_min(-1)
    vals, idx, rank, unc = ops.pairwise_topk(Qd.cuda(), G_view, k, lt, pos_index=pos.cuda(), return_uncertified=True)
    Qo, Go = Qd.float(), Gd.float()[1:]
    kk = min(k, ng)
    ref_v, ref_i = O.pairwise_topk_batched(Qo, Go, kk, lt)
    dist_rows = [O.distances(Qo[i:i + 1], Go, lt) for i in range(nq)]
    assert_topk_matches(vals[:, :kk], idx[:, :kk], ref_v, ref_i, dist_rows)
    assert_ranks_match_up_to_fp32_ties(rank, Qo, Go, pos, lt)


def test_edge_cases(ops):
    dev = "cuda"
    torch.manual_seed(0)
    G = torch.randn(5, 64)
    Q = torch.randn(3, 64)
    # fewer gallery rows than k: tail padded with (+inf, -1)
    vals, idx = ops.pairwise_topk(Q.to(dev), G.to(dev), 10, "euclidean")
    rv, ri = O.pairwise_topk_batched(Q, G, 10, "euclidean")
    assert torch.equal(idx[:, :5].cpu(), ri) and torch.allclose(vals[:, :5].cpu(), rv, rtol=DIST_RTOL)
    assert (idx[:, 5:] == -1).all() and torch.isinf(vals[:, 5:]).all()
    # empty gallery / empty queries
    vals, idx, rank = ops.pairwise_topk(Q.to(dev), torch.empty(0, 64, device=dev), 3, "euclidean",
                                        pos_index=torch.tensor([-1, -1, -1], device=dev))
    assert (idx == -1).all() and (rank == 0).all()
    vals, idx = ops.pairwise_topk(torch.empty(0, 64, device=dev), G.to(dev), 3, "euclidean")
    assert vals.shape == (0, 3)
    # exact duplicates: ties resolved by ascending index; rank = count of strictly closer rows
    G2 = torch.cat([G, G[1:2], G[1:2]])            # rows 5, 6 duplicate row 1
    q = G[1:2] + 0.01
    vals, idx, rank = ops.pairwise_topk(q.to(dev), G2.to(dev), 3, "euclidean", pos_index=torch.tensor([6], device=dev))
    # canonical order (distance, index): row 6 sits behind its duplicates 1 and 5
    assert idx[0].tolist() == [1, 5, 6] and rank.item() == 2
    assert O.ranking_position(q, G2, 6, "euclidean") in (0, 1, 2)   # the reference lands on one of the tied slots
    # zero vectors under cosine: distance 1 to everything (per-operand clamp, utils.py:34)
    Gz = torch.cat([torch.zeros(1, 64), G])
    vals, idx = ops.pairwise_topk(torch.zeros(1, 64, device=dev), Gz.to(dev), 6, "cosine")
    assert torch.allclose(vals.cpu(), torch.ones(1, 6)) and idx[0].tolist() == [0, 1, 2, 3, 4, 5]
    ref = O.cosine_distance(Q[:1], Gz)
    vals, idx = ops.pairwise_topk(Q[:1].to(dev), Gz.to(dev), 6, "cosine")
    assert torch.allclose(vals.cpu()[0], ref.topk(6, largest=False).values, rtol=DIST_RTOL, atol=1e-6)
    with pytest.raises(Exception, match="loss type not correct"):
        ops.pairwise_topk(Q.to(dev), G.to(dev), 3, "manhattan")
    with pytest.raises(ValueError):
        ops.pairwise_topk(Q.to(dev), G.to(dev), 500, "euclidean")


@pytest.mark.parametrize("nq,ng,d,k", [(3, 5, 64, 10), (1, 1, 8, 1), (130, 257, 72, 4), (129, 129, 512, 16), (40, 3000, 200, 26)])
def test_bf16_small_and_ragged_shapes(ops, nq, ng, d, k):
    """bf16 rows of at most 1 KB take the resident-query form (query tile in tensor memory): fewer
    gallery rows than k, single rows, row lengths that do not fill the last 128-byte k-block, tile
    boundaries off by one.  Indices / ranks against the oracle on the bf16-rounded inputs."""
    g = torch.Generator().manual_seed(nq * 1000 + ng)
    Q = torch.randn(nq, d, generator=g).bfloat16()
    G = torch.randn(ng, d, generator=g).bfloat16()
    pos = torch.randint(0, ng, (nq,), generator=g)
    for lt in ("euclidean", "cosine"):
        vals, idx, rank = ops.pairwise_topk(Q.cuda(), G.cuda(), k, lt, pos_index=pos.cuda())
        kk = min(k, ng)
        ref_v, ref_i = O.pairwise_topk_batched(Q.float(), G.float(), kk, lt)
        ref_r = O.rank_of_positive_batched(Q.float(), G.float(), pos, lt)
        dist_rows = [O.distances(Q[i:i + 1].float(), G.float(), lt) for i in range(nq)]
        assert_topk_matches(vals[:, :kk], idx[:, :kk], ref_v, ref_i, dist_rows)
        assert (idx[:, kk:] == -1).all() and torch.isinf(vals[:, kk:]).all()
        # bf16 guarantee: ranks equal the torch-fp32 oracle on the bf16-rounded inputs except
        # where a gallery row ties with the positive within a few fp32 ulps — every deviation is proven a tie
        differing = assert_ranks_match_up_to_fp32_ties(rank, Q.float(), G.float(), pos, lt, ref_r)
        assert differing <= max(1, nq // 20)


@pytest.mark.parametrize("dtype,lt,k", [("bfloat16", "euclidean", 100), ("float32", "euclidean", 60), ("bfloat16", "cosine", 116),
                                        ("float32", "cosine", 100)])
def test_large_lists_on_hit_dense_data_with_duplicates(ops, dtype, lt, k):
    """Lists of 64/128 entries run the owner + feeder epilogue (8 warps, one list per row, hits of the
    second column half forwarded through a shared-memory queue, insertions batched per lane).  Stress it:
    unstructured data sorted so that later gallery rows are CLOSER (every tile keeps producing hits and
    the queues overflow), every row present three times (exact ties), several query tiles and gallery
    tiles that do not divide evenly.  Exact top-k sets, tie order by index, exact ranks."""
    g = torch.Generator().manual_seed(77)
    nq, base, d = 2600, 7000, 64
    tdt = getattr(torch, dtype)
    Q = torch.randn(nq, d, generator=g).to(tdt)
    G0 = torch.randn(base, d, generator=g)
    # gallery rows ordered by decreasing distance to the mean query: thresholds keep dropping during the scan
    order = (G0 - Q.float().mean(0)).norm(dim=1).argsort(descending=True)
    G = G0[order].repeat(3, 1).to(tdt)                      # rows j, j + base, j + 2·base are identical
    ng = G.shape[0]
    pos = torch.randint(0, ng, (nq,), generator=g)
    vals, idx, rank, unc = ops.pairwise_topk(Q.cuda(), G.cuda(), k, lt, pos_index=pos.cuda(), return_uncertified=True)
    Qf, Gf = Q.float(), G.float()
    ref_v, ref_i = O.pairwise_topk_batched(Qf, Gf, k, lt)
    dist_rows = [O.distances(Qf[i:i + 1], Gf, lt) for i in range(nq)]
    vals_c, idx_c = vals.cpu(), idx.cpu()
    assert torch.allclose(vals_c, ref_v.float(), rtol=DIST_RTOL, atol=1e-6)
    for i in range(nq):                                     # same multiset of distances; ties ordered by index
        got = dist_rows[i][idx_c[i]]
        assert torch.allclose(got, ref_v[i].float(), rtol=TIE_RTOL, atol=1e-6), i
        same = vals_c[i, 1:] == vals_c[i, :-1]
        assert (idx_c[i, 1:][same] > idx_c[i, :-1][same]).all(), i
    # rank = rows strictly closer + equally distant rows with a smaller index (canonical order); fp32-level
    # near-ties around d_pos may move it by a duplicate group
    dpos = torch.stack([dist_rows[i][pos[i]] for i in range(nq)])
    lower = torch.stack([(dist_rows[i] < dpos[i] * (1 - TIE_RTOL)).sum() for i in range(nq)])
    upper = torch.stack([(dist_rows[i] <= dpos[i] * (1 + TIE_RTOL)).sum() for i in range(nq)])
    r = rank.cpu()
    assert ((r >= lower) & (r < upper)).all()
    assert int(unc.item()) <= nq // 50 + 4


def test_resident_query_form_matches_default_kernel(ops, dbg):
    """Resident-query form (default for bf16 rows of at most 1 KB and small lists): the query tile lives
    in tensor memory and is the MMA's A operand from there (tcgen05.mma with A in TMEM), gallery
    half-tiles stream through shared memory.  option k1_qres = 0 selects the all-shared-memory form:
    identical results, ragged shapes included."""
    for nq, ng, d, lt, k in ((257, 3001, 512, "euclidean", 10), (1000, 20000, 192, "cosine", 20), (130, 700, 64, "euclidean", 1)):
        Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=nq, beta=0.3 if d < 512 else None)
        q, g, p = Q.bfloat16().cuda(), G.bfloat16().cuda(), pos.cuda()
        dbg("k1_qres", 0)
        v0, i0, r0 = ops.pairwise_topk(q, g, k, lt, pos_index=p)
        dbg("k1_qres", -1)
        v1, i1, r1 = ops.pairwise_topk(q, g, k, lt, pos_index=p)
        assert torch.equal(v0, v1) and torch.equal(i0, i1) and torch.equal(r0, r1)


def test_cta_pairs_match_single_cta_tiles_on_bf16_tiles(ops, dbg):
    """CTA pairs (cta_group::2, M = 256; the default for bf16 tiles of rows >= 4 KB, i.e. 2048-d fp32 embeddings selected on
    their bf16 copies) against single-CTA tiles: same candidates, same certificate, identical results — small and large
    lists, both metrics, an odd number of query tiles, bf16 embeddings and bf16-selected fp32 embeddings."""
    cases = ((700, 5000, 2048, torch.float32, "euclidean", 10), (300, 4000, 2048, torch.float32, "cosine", 100),
             (900, 6000, 256, torch.bfloat16, "euclidean", 10), (385, 3000, 2048, torch.bfloat16, "cosine", 30))
    for nq, ng, d, dtype, lt, k in cases:
        Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=nq + 7, beta=0.3 if d < 512 else None)
        q, g, p = Q.to(dtype).cuda(), G.to(dtype).cuda(), pos.cuda()
        dbg("reset", 0)
        if dtype == torch.float32:
            dbg("k1_sel_bf16", 1)
        dbg("k1_pair", 1)
        v0, i0, r0 = ops.pairwise_topk(q, g, k, lt, pos_index=p)
        dbg("k1_pair", 2)
        v1, i1, r1 = ops.pairwise_topk(q, g, k, lt, pos_index=p)
        assert torch.equal(v0, v1) and torch.equal(i0, i1) and torch.equal(r0, r1), (nq, ng, d, lt, k)
        want_v, want_i = O.pairwise_topk_batched(Q.to(dtype).float(), G.to(dtype).float(), k, lt)
        assert torch.equal(i1.cpu(), want_i) or (i1.cpu() != want_i).float().mean() < 0.01   # ties only (checked in detail elsewhere)
        assert torch.allclose(v1.cpu(), want_v.float(), rtol=2e-5, atol=1e-6)


def test_l2_bands_of_query_tiles_give_the_same_result(ops, dbg):
    """Unit order with the query tiles walked in L2 bands (option k1_bands; every band scans all chunk steps before the
    next one starts): a different schedule of the same units, so results are identical — also with chunk hand-overs,
    several gallery partitions, CTA pairs (static unit stride) and a last band that is shorter than the others."""
    cases = ((20000, 9000, 64, torch.bfloat16, "euclidean", 10, 1), (1500, 40000, 128, torch.float32, "cosine", 100, 1),
             (3000, 30000, 96, torch.float32, "euclidean", 10, 1), (700, 5000, 512, torch.bfloat16, "euclidean", 20, 0))
    for nq, ng, d, dtype, lt, k, chunk_mb in cases:
        Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=nq + 1, beta=0.3)
        q, g, p = Q.to(dtype).cuda(), G.to(dtype).cuda(), pos.cuda()
        dbg("reset", 0)
        if chunk_mb:
            dbg("k1_chunk_mb", chunk_mb)
        dbg("k1_bands", -1)
        v0, i0, r0 = ops.pairwise_topk(q, g, k, lt, pos_index=p)
        for bands in (2, 3, 5):
            dbg("k1_bands", bands)
            v1, i1, r1 = ops.pairwise_topk(q, g, k, lt, pos_index=p)
            assert torch.equal(v0, v1) and torch.equal(i0, i1) and torch.equal(r0, r1), (nq, ng, d, lt, k, bands)


@pytest.mark.parametrize("nq,ng,d,lt,k", [(130, 1000, 512, "euclidean", 1), (200, 2500, 1024, "cosine", 10), (64, 1500, 2048, "euclidean", 30),
                                           (300, 5000, 96, "euclidean", 10), (1000, 20000, 192, "cosine", 20), (37, 300, 64, "euclidean", 100)])
def test_fp32_selected_on_bf16_copies_matches_tf32_selection_and_oracle(ops, dbg, nq, ng, d, lt, k):
    """fp32 embeddings selected on their bf16-rounded copies (kind::f16 tiles at twice the kind::tf32 rate; certificate
    and rank band from measured rounding residuals).  Large problems take this path on their own; here it is forced
    on small shapes: results must be bit-identical to the kind::tf32 selection (both are re-scored exactly) and
    match the oracle — including k = 100 on a tiny gallery, where the wider band cannot be certified and the
    escalation / brute-force fallbacks have to deliver the exact answer."""
    Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=nq + ng + 1, beta=0.3 if d < 512 else None, num_classes=max(4, ng // 80))
    pos[::7] = -1
    q, g, p = Q.cuda(), G.cuda(), pos.cuda()
    dbg("k1_sel_bf16", 0)
    v0, i0, r0, u0 = ops.pairwise_topk(q, g, k, lt, pos_index=p, return_uncertified=True)
    dbg("k1_sel_bf16", 1)
    v1, i1, r1, u1 = ops.pairwise_topk(q, g, k, lt, pos_index=p, return_uncertified=True)
    assert torch.equal(v0, v1) and torch.equal(i0, i1) and torch.equal(r0, r1)
    ref_v, ref_i = O.pairwise_topk_batched(Q, G, k, lt)
    dist_rows = [O.distances(Q[i:i + 1], G, lt) for i in range(nq)]
    assert_topk_matches(v1, i1, ref_v, ref_i, dist_rows)
    assert_ranks_match_up_to_fp32_ties(r1, Q, G, pos, lt)
    if k <= 30:
        assert int(u1.item()) <= nq // 50 + 4                 # lists with room for the bf16 band certify (almost) everything


def test_bf16_selection_tiers_at_cfg3_top100(ops, dbg):
    """BASELINE cfg3 (12.5k x 75k x 2048 fp32, top-100): with 128-entry lists the bf16 selection band leaves ~0.6 % of the
    queries uncertified; the tier behind the pass re-selects exactly those on kind::tf32 tiles as a small batch.  Results
    must be bit-identical to the all-kind::tf32 pass, with nothing left to brute force."""
    Q, G, pos = _device_clustered(12500, 75000, 2048, torch.float32)
    dbg("k1_sel_bf16", 0)
    want = ops.pairwise_topk(Q, G, 100, "euclidean", pos_index=pos, return_uncertified=True)
    dbg("k1_sel_bf16", 1)
    got = ops.pairwise_topk(Q, G, 100, "euclidean", pos_index=pos, return_uncertified=True)
    for a, b in zip(want[:3], got[:3]):
        assert torch.equal(a, b)
    assert int(want[3].item()) == 0 and int(got[3].item()) == 0


def test_fp32_bf16_selection_on_collapsed_embeddings(ops, dbg):
    """Forced bf16 selection on embeddings with a large common component: the measured residual norms make the band
    cover everything, nothing certifies, and the centred 3xTF32 escalation pass must still produce exact results."""
    nq, ng, d = 200, 8000, 256
    Q0, G0, pos = O.synthetic_embeddings(nq, ng, d, seed=9, beta=0.12)
    base = 3.0 * torch.rand(1, d, generator=torch.Generator().manual_seed(1))
    Q, G = (base + 0.02 * Q0).contiguous(), (base + 0.02 * G0).contiguous()
    dbg("k1_sel_bf16", 0)
    want = ops.pairwise_topk(Q.cuda(), G.cuda(), 10, "euclidean", pos_index=pos.cuda())
    dbg("k1_sel_bf16", 1)
    got = ops.pairwise_topk(Q.cuda(), G.cuda(), 10, "euclidean", pos_index=pos.cuda(), return_uncertified=True)
    for a, b in zip(want, got[:3]):
        assert torch.equal(a, b)
    assert int(got[3].item()) <= nq // 50 + 4


def test_cancellation_heavy_fp32_escalates_to_3xtf32(ops):
    """Embeddings whose norms dwarf their distances (post-ReLU-like: a large common component) and
    positives unrelated to the queries: the TF32 error band covers much of the distance
    distribution, so the first pass cannot certify anything; the device-gated 3xTF32 pass must
    take over and the results must still be exact."""
    g = torch.Generator().manual_seed(5)
    nq, ng, d = 300, 6000, 256
    G = 8.0 + 0.5 * torch.randn(ng, d, generator=g)
    Q = 8.0 + 0.5 * torch.randn(nq, d, generator=g)
    pos = torch.randint(0, ng, (nq,), generator=g)
    for lt in ("euclidean", "cosine"):
        vals, idx, rank, unc = ops.pairwise_topk(Q.cuda(), G.cuda(), 10, lt, pos_index=pos.cuda(), return_uncertified=True)
        ref_v, ref_i = O.pairwise_topk_batched(Q, G, 10, lt, fp64=True)
        ref_r = O.rank_of_positive_batched(Q, G, pos, lt, fp64=True)
        dist_rows = [O.distances(Q[i:i + 1].double(), G.double(), lt) for i in range(nq)]
        vals_c, idx_c = vals.cpu(), idx.cpu()
        assert torch.allclose(vals_c.double(), ref_v, rtol=DIST_RTOL, atol=1e-6)
        for i, j in (idx_c != ref_i).nonzero().tolist():      # swaps only between fp32-level ties
            assert abs(dist_rows[i][idx_c[i, j]].item() - ref_v[i, j].item()) <= TIE_RTOL * max(abs(ref_v[i, j].item()), 1e-6)
        # against the torch-fp32 oracle (what north_star names): cosine distances here are ~4e-3 apart by
        # ~2e-7 while the fp32 quotients / sums the reference formula prescribes carry ~1e-7 of rounding, so
        # two correct fp32 evaluations legitimately differ — but ONLY on rows tied with the positive within
        # a few fp32 ulps of the similarity, which is what is asserted for every query
        assert_ranks_match_up_to_fp32_ties(rank, Q, G, pos, lt)
        assert int(unc.item()) <= nq // 50 + 4                   # the escalated pass certified (almost) everything


def test_collapsed_embeddings_use_the_centred_pass(ops):
    """Embeddings that share a large common component and differ only slightly (an untrained encoder,
    BASELINE cfg5; post-ReLU features): TF32 — and even 3xTF32 on the raw operands — cannot tell the
    rows apart, so the escalation pass centres both operands on the gallery mean (distances are
    translation-invariant).  Results must be exact and (almost) every query certified."""
    nq, ng, d = 200, 8000, 256
    Q0, G0, pos = O.synthetic_embeddings(nq, ng, d, seed=9, beta=0.12)
    base = 3.0 * torch.rand(1, d, generator=torch.Generator().manual_seed(1))
    Q, G = (base + 0.02 * Q0).contiguous(), (base + 0.02 * G0).contiguous()
    vals, idx, rank, unc = ops.pairwise_topk(Q.cuda(), G.cuda(), 10, "euclidean", pos_index=pos.cuda(), return_uncertified=True)
    # ground truth = the reference formula element by element in fp32, long sum in fp64
    rows = []
    for a in range(0, nq, 25):
        diff = (Q[a:a + 25, None, :] - G[None, :, :]) + torch.tensor(1e-6)
        rows.append((diff.double() ** 2).sum(-1).sqrt().float())
    dd = torch.cat(rows)
    dp = dd.gather(1, pos[:, None])
    ref_r = ((dd < dp) | ((dd == dp) & (torch.arange(ng)[None, :] < pos[:, None]))).sum(1)
    ref_v, ref_i = torch.topk(dd, 10, dim=1, largest=False)
    assert torch.allclose(vals.cpu(), ref_v, rtol=DIST_RTOL, atol=1e-7)
    idx_c = idx.cpu()
    for i, j in (idx_c != ref_i).nonzero().tolist():      # swaps only between fp32-level ties
        assert abs(dd[i, idx_c[i, j]].item() - ref_v[i, j].item()) <= TIE_RTOL * abs(ref_v[i, j].item())
    assert torch.equal(rank.cpu(), ref_r)
    assert int(unc.item()) <= nq // 50 + 4


def test_fp64_gallery_like_csv_features(ops):
    """F8: CSV-loaded galleries are float64 in the reference; ranks must still agree."""
    from art_sbir_b200 import inference as inf
    Q, G, pos = O.synthetic_embeddings(20, 400, 128, seed=3, beta=0.3, num_classes=5)
    G64 = G.double()
    paths = [Path(f"p/n{i:05d}.jpg") for i in range(400)]
    for i in range(0, 20, 7):
        sk = Path(f"s/n{int(pos[i]):05d}-1.png")
        assert inf.get_ranking_position(sk, paths, Q[i:i + 1], G64, "euclidean") == \
            O.get_ranking_position(sk, paths, Q[i:i + 1], G64, "euclidean")


def test_rowwise_distance_modules(ops):
    from art_sbir_b200 import utils as U
    torch.manual_seed(1)
    q, G = torch.randn(1, 777), torch.randn(300, 777)         # 777: scalar (unvectorised) row path
    for mod, ref in ((U.euclidean_distance, O.euclidean_distance), (U.cosine_distance, O.cosine_distance)):
        got = mod(q.cuda(), G.cuda()).cpu()
        assert torch.allclose(got, ref(q, G), rtol=1e-5, atol=1e-6)
        a, b = torch.randn(40, 512), torch.randn(40, 512)
        assert torch.allclose(mod(a.cuda(), b.cuda()).cpu(), ref(a, b), rtol=1e-5, atol=1e-6)
        # autograd through the drop-in distance module, broadcast and same-shape
        for x1, x2 in ((q[:, :64], G[:50, :64]), (a[:, :64], b[:, :64])):
            X1, X2 = x1.clone().cuda().requires_grad_(True), x2.clone().cuda().requires_grad_(True)
            mod(X1, X2).pow(2).sum().backward()
            Y1, Y2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
            ref(Y1, Y2).pow(2).sum().backward()
            assert torch.allclose(X1.grad.cpu(), Y1.grad, rtol=1e-3, atol=1e-5 * Y1.grad.abs().max().item())
            assert torch.allclose(X2.grad.cpu(), Y2.grad, rtol=1e-3, atol=1e-5 * Y2.grad.abs().max().item())


@pytest.mark.parametrize("in_dtype,store_dtype,normalize", [("float32", "float32", False), ("float32", "bfloat16", False),
                                                             ("bfloat16", "bfloat16", False), ("bfloat16", "float32", True),
                                                             ("float32", "bfloat16", True), ("float32", "float32", True)])
def test_gallery_append_builds_rows_and_norms_in_place(ops, in_dtype, store_dtype, normalize):
    """N1 (inference.py:72-92): blocks of encoder output appended straight into the preallocated gallery in
    its storage type, with ‖stored row‖²; ragged last block, a row length that is not a multiple of the
    vector width, and the scoring pass fed with the stored norms must give the same answer as without."""
    tin, tst = getattr(torch, in_dtype), getattr(torch, store_dtype)
    for n, d, bs in ((1037, 512, 50), (130, 100, 64), (7, 36, 7)):
        g = torch.Generator().manual_seed(n + d)
        X = (torch.randn(n, d, generator=g) * 3).to(tin)
        buf = ops.GalleryBuffer(n, d, tst, normalize=normalize)
        for lo in range(0, n, bs):
            buf.append(X[lo:lo + bs].cuda())
        assert buf.filled == n
        want = X.float()
        if normalize:
            want = O.l2_normalize(want)
        want = want.to(tst)
        got = buf.rows.cpu()
        if normalize and store_dtype == "bfloat16":      # reciprocal-multiply before the bf16 rounding: within one bf16 ulp
            assert torch.allclose(got.float(), want.float(), rtol=2 ** -7, atol=1e-30)
        elif normalize:
            assert torch.allclose(got.float(), want.float(), rtol=1e-6, atol=1e-9)    # norm and quotient: an fp32 ulp each
        else:
            assert torch.equal(got, want)
        sq = (got.double() ** 2).sum(1)
        assert torch.allclose(buf.sqnorm.cpu().double(), sq, rtol=2e-7)
        with pytest.raises(ValueError):
            buf.append(X[:1].cuda())                                             # full
    Q, G, pos = O.synthetic_embeddings(200, 3000, 256, seed=4)
    buf = ops.GalleryBuffer(3000, 256, tst)
    for lo in range(0, 3000, 500):
        buf.append(G[lo:lo + 500].to(tin).cuda())
    q = Q.to(tst).cuda()
    for lt in ("euclidean", "cosine"):
        a = ops.pairwise_topk(q, buf.rows, 10, lt, pos_index=pos.cuda())
        b = ops.pairwise_topk(q, buf.rows, 10, lt, pos_index=pos.cuda(), gallery_sqnorm=buf.sqnorm)
        for x, y in zip(a, b):
            assert torch.equal(x, y)


@pytest.mark.parametrize("loss_type", ["euclidean", "cosine"])
def test_run_inference_computed_and_stored_routes_match_golden(ops, golden_dir, loss_type, tmp_path):
    """inference.py:140-165 end to end with an identity encoder: (1) the gallery is BUILT by
    compute_image_features (append kernel → preallocated buffer → feature store with sidecar), (2) the
    second run LOADS it by folder name (sidecar rows + norms), (3) a third run reads the reference's own CSV
    pair (float64, F8).  All three must reproduce the dict the unmodified reference produced."""
    from art_sbir_b200 import inference as inf
    z = np.load(golden_dir / f"retrieval_{loss_type}.npz")
    ref = json.load(open(golden_dir / f"process_inference_{loss_type}.json"))
    Q, G = torch.from_numpy(z["Q"]), torch.from_numpy(z["G"])
    image_paths = [Path(str(p)) for p in z["image_paths"]]

    class Sketches(torch.utils.data.Dataset):
        sketch_paths = [Path(str(p)) for p in z["sketch_paths"]]
        photo_paths = image_paths
        transform = None
        state_dict = {"dataset": "SketchyV1"}

        def __len__(self):
            return len(self.sketch_paths)

        def __getitem__(self, i):
            return (Q[i],)

    class MemGallery(inf.InferenceDataset):
        def load_image(self, idx):
            return G[image_paths.index(self.image_paths[idx])]

    def check(got, folder):
        assert got["image_features"] == folder
        for key in ("size", "count", "min", "25%", "50%", "75%", "max"):
            assert got[key] == ref[key], key
        assert got["topk_acc"] == ref["topk_acc"]
        assert got["mean_reciprocal_rank"] == pytest.approx(ref["mean_reciprocal_rank"], rel=1e-12)
        assert len(got["retrieval_samples"]) == len(ref["retrieval_samples"])
        for a, b in zip(got["retrieval_samples"], ref["retrieval_samples"]):
            (ka, va), (kb, vb) = next(iter(a.items())), next(iter(b.items()))
            assert ka == kb and [p for p, _ in va] == [p for p, _ in vb]
            assert np.allclose([d for _, d in va], [d for _, d in vb], rtol=DIST_RTOL)
        json.dumps(got)

    ds = Sketches()
    got = inf.run_inference(torch.nn.Identity(), ds, None, loss_type, inference_dataset=MemGallery(image_paths), feature_root=tmp_path)
    folder = got["image_features"]
    assert folder.startswith("Identity_SketchyV1_") and (tmp_path / folder / "image_features.f32.npy").is_file()
    assert (tmp_path / folder / "image_sqnorm.f32.npy").is_file() and (tmp_path / folder / "image_features.csv").is_file()
    check(got, folder)
    check(inf.run_inference(torch.nn.Identity(), ds, folder, loss_type, feature_root=tmp_path), folder)       # sidecar route
    for side in ("image_features.f32.npy", "image_sqnorm.f32.npy"):
        (tmp_path / folder / side).unlink()
    check(inf.run_inference(torch.nn.Identity(), ds, folder, loss_type, feature_root=tmp_path), folder)       # reference CSV route


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_l2_normalize(ops, dtype):
    tdt = getattr(torch, dtype)
    x = torch.randn(1000, 520).to(tdt)
    x[7] = 0
    got = ops.l2_normalize(x.cuda()).cpu()
    ref = O.l2_normalize(x.float())
    tol = 1e-6 if dtype == "float32" else 4e-3
    assert torch.allclose(got.float(), ref, atol=tol, rtol=tol)
    assert (got[7] == 0).all()
    big = torch.randn(64, 4100)          # rows longer than the register-held part
    assert torch.allclose(ops.l2_normalize(big.cuda()).cpu(), O.l2_normalize(big), atol=1e-6)
    for rows, dim in ((1001, 512), (7, 64), (4099, 256)):   # rows <= 1 KB: four rows per warp
        s = torch.randn(rows, dim).to(tdt)
        s[rows // 2] = 0
        got = ops.l2_normalize(s.cuda()).cpu()
        assert torch.allclose(got.float(), O.l2_normalize(s.float()), atol=tol, rtol=tol)
        assert (got[rows // 2] == 0).all()


@pytest.mark.parametrize("lt", ["euclidean", "cosine"])
def test_triplet_and_batch_hard_cfg2(ops, lt):
    """BASELINE config 2: a/p/n [256, 2048] fp32, margin 0.2."""
    g = torch.Generator().manual_seed(11)
    a, p, n = (torch.randn(256, 2048, generator=g) for _ in range(3))
    p = a + 0.9 * p
    A, P, N = (t.clone().cuda().requires_grad_(True) for t in (a, p, n))
    loss = ops.triplet_margin_loss(A, P, N, 0.2, lt)
    loss.backward()
    Ar, Pr, Nr = (t.clone().requires_grad_(True) for t in (a, p, n))
    ref = O.triplet_margin_loss(Ar, Pr, Nr, 0.2, lt)
    ref.backward()
    assert loss.item() == pytest.approx(ref.item(), rel=DIST_RTOL)
    for got, want in ((A.grad, Ar.grad), (P.grad, Pr.grad), (N.grad, Nr.grad)):
        assert torch.allclose(got.cpu(), want, rtol=DIST_RTOL, atol=1e-6 * want.abs().max().item())

    for labels in (None, torch.arange(256) // 4):
        _check_batch_hard(ops, a, p, n, lt, labels)


def _check_batch_hard(ops, a, p, n, lt, labels=None, margin=0.2):
    """Batch-hard (H8) against the oracle: loss within 1e-3, the mined indices IDENTICAL to the oracle's
    wherever the oracle itself is unambiguous (its fp32 and fp64 evaluations pick the same candidate),
    and — unconditionally — all three gradients equal to autograd through the reference's distance
    modules at the selected pairs."""
    B = a.shape[0]
    A, P, N = (t.clone().cuda().requires_grad_(True) for t in (a, p, n))
    loss, hard = ops.batch_hard_triplet_loss(A, P, N, margin, lt, labels=None if labels is None else labels.cuda(),
                                             return_indices=True)
    loss.backward()
    hard = hard.cpu()
    Ar, Pr, Nr = (t.clone().requires_grad_(True) for t in (a, p, n))
    ref, hpi, hni = O.batch_hard_triplet_loss(Ar, Pr, Nr, margin, lt, labels)
    ref.backward()
    _, hpi64, hni64 = O.batch_hard_triplet_loss(a.double(), p.double(), n.double(), margin, lt, labels)
    assert loss.item() == pytest.approx(ref.item(), rel=DIST_RTOL, abs=1e-7)
    for got, o32, o64 in ((hard[:, 0], hpi, hpi64), (hard[:, 1], hni, hni64)):
        sure = o32 == o64
        assert torch.equal(got[sure], o32[sure])
        assert ((got == o32) | (got == o64)).all()
    # gradients: autograd of the reference formula evaluated at OUR selection (no condition on the indices)
    As, Ps, Ns = (t.clone().requires_grad_(True) for t in (a, p, n))
    X = torch.cat([Ps, Ns])
    dist = O.euclidean_distance if lt == "euclidean" else O.cosine_distance
    sel = torch.clamp_min(margin + dist(As, X[hard[:, 0]]) - dist(As, X[hard[:, 1]]), 0).mean()
    sel.backward()
    assert loss.item() == pytest.approx(sel.item(), rel=DIST_RTOL, abs=1e-7)
    for got, want in ((A.grad, As.grad), (P.grad, Ps.grad), (N.grad, Ns.grad)):
        assert torch.allclose(got.cpu(), want, rtol=DIST_RTOL, atol=2e-6 * max(want.abs().max().item(), 1e-30))
    if torch.equal(hard[:, 0], hpi) and torch.equal(hard[:, 1], hni):      # then they are the oracle's own gradients too
        for got, want in ((A.grad, Ar.grad), (P.grad, Pr.grad), (N.grad, Nr.grad)):
            assert torch.allclose(got.cpu(), want, rtol=DIST_RTOL, atol=2e-6 * max(want.abs().max().item(), 1e-30))
    return hard


@pytest.mark.parametrize("B,D", [(32, 1024), (100, 520), (1, 64), (129, 36), (300, 2048)])
@pytest.mark.parametrize("lt", ["euclidean", "cosine"])
def test_batch_hard_shapes(ops, B, D, lt):
    """The reference's own batch size (32, train.py:108) and embedding width (1024), ragged batches that do
    not fill the 128-anchor / 32-candidate tiles, row lengths that do not fill the last 128-byte k-block,
    a single triplet, more than two anchor tiles."""
    g = torch.Generator().manual_seed(B * 7 + D)
    a, p, n = (torch.randn(B, D, generator=g) for _ in range(3))
    p = a + 0.8 * p
    _check_batch_hard(ops, a, p, n, lt)
    if B > 4:
        _check_batch_hard(ops, a, p, n, lt, labels=torch.arange(B) // 3)


def test_batch_hard_near_ties_and_collapsed_embeddings_are_mined_exactly(ops):
    """Selection must not depend on tensor-core rounding: (1) many negatives almost equally hard (inside
    the TF32 error band of each other) — every one of them has to be re-scored exactly; (2) embeddings with
    a large common component, where the band covers the whole batch and mining degenerates to exact
    brute force.  Indices identical to the oracle, gradients unconditional."""
    g = torch.Generator().manual_seed(3)
    B, D = 96, 512
    a = torch.randn(B, D, generator=g)
    p = a + 0.5 * torch.randn(B, D, generator=g)
    base = torch.randn(1, D, generator=g)
    n = base + 1e-3 * torch.randn(B, D, generator=g)            # all negatives within 1e-3 of one point
    for lt in ("euclidean", "cosine"):
        _check_batch_hard(ops, a, p, n, lt)
    off = 30.0 * torch.rand(1, D, generator=g)
    _check_batch_hard(ops, off + 0.05 * a, off + 0.05 * p, off + 0.05 * torch.randn(B, D, generator=g), "euclidean")


def test_tensor_core_error_stays_inside_the_certified_bound(ops):
    """The selection certificate and the rank band rely on |e_tc − e_exact| <= kappa·(‖q‖²+‖g‖²) with
    kappa = 2^-9·1.01 (kind::tf32 operand truncation) + (k-steps + 16)·2^-23 (fp32 accumulation, one
    rounding per 32-byte k-step), see csrc/kernels.h.  Checked on zero-mean data and on all-positive
    data (post-ReLU-like), where accumulation rounding does not cancel."""
    D = 1024
    for dtype, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        kappa = (D * es / 32 + 16) * 2.0 ** -23 + (2.0 ** -9 * 1.01 if dtype == torch.float32 else 0.0)
        for offset in (0.0, 3.0):
            Q, G, _ = O.synthetic_embeddings(300, 2000, D, seed=9)
            Q, G = (Q + offset).to(dtype).cuda(), (G + offset).to(dtype).cuda()
            e = ops.debug_dist_matrix(Q, G, "euclidean").double()
            qd, gd = Q.double(), G.double()
            ref = (gd ** 2).sum(1)[None, :] - 2 * qd @ gd.T
            bound = kappa * ((qd ** 2).sum(1)[:, None] + (gd ** 2).sum(1)[None, :])
            assert not torch.isnan(e).any()
            worst = ((e - ref).abs() / bound).max().item()
            assert worst <= 1.0, (str(dtype), offset, worst)


# --------------------------------------------------- shard / merge / host-buffer identities ----
def test_sharded_merge_equals_single_pass(ops):
    Q, G, pos = O.synthetic_embeddings(300, 9001, 512, seed=21)
    Qc, Gc, pc = Q.cuda(), G.cuda(), pos.cuda()
    v1, i1, r1 = ops.pairwise_topk(Qc, Gc, 10, "euclidean", pos_index=pc)
    from art_sbir_b200 import sharded
    for world in (2, 3, 8):
        vs, is_, cnt = [], [], torch.zeros_like(r1)
        own = torch.full((300,), float("nan"), dtype=torch.float64, device="cuda")
        for r in range(world):
            a, b = sharded.shard_bounds(9001, world, r)
            mine = (pc >= a) & (pc < b)
            d = ops.positive_distance(Qc, Gc[a:b], torch.where(mine, pc - a, torch.full_like(pc, -1)))
            own = torch.where(mine, d, own)
        for r in range(world):
            a, b = sharded.shard_bounds(9001, world, r)
            v, i, c, _ = ops.pairwise_topk_shard(Qc, Gc[a:b].contiguous(), 10, "euclidean", a, own, pc)
            vs.append(v); is_.append(i); cnt += c
        vm, im = ops.topk_merge(torch.stack(vs), torch.stack(is_))
        assert torch.equal(im, i1) and torch.equal(vm, v1) and torch.equal(cnt, r1)


@pytest.mark.parametrize("lists,nq,k", [(8, 1000, 10), (2, 77, 100), (3, 5, 1), (8, 333, 116), (40, 9, 116)])
def test_topk_merge_against_a_sort(ops, lists, nq, k):
    """K4: k best of the gathered per-shard lists, ties by index, padding entries (index -1, +inf)
    of short lists ignored; (40, 9, 116) exceeds the shared-memory form and takes the generic kernel."""
    g = torch.Generator().manual_seed(lists * 1000 + k)
    d = torch.randint(0, 50, (lists, nq, k), generator=g).float() * 0.25     # many exact ties
    i = torch.randperm(lists * nq * k, generator=g).reshape(lists, nq, k) % 100000
    short = torch.randint(0, k + 1, (lists, nq), generator=g)                # valid entries per list
    short[0] = k if lists > 1 else short[0]
    pad = torch.arange(k)[None, None, :] >= short[:, :, None]
    d[pad] = float("inf")
    i[pad] = -1
    # each list ascending by (distance, index), padding last
    key = torch.where(pad, torch.full_like(d, 1e30), d).double() * 1e6 + i.clamp_min(0).double()
    order = key.argsort(dim=2)
    d, i = d.gather(2, order), i.gather(2, order)
    vm, im = ops.topk_merge(d.cuda(), i.cuda())
    dd = d.permute(1, 0, 2).reshape(nq, -1)
    ii = i.permute(1, 0, 2).reshape(nq, -1)
    key = torch.where(ii < 0, torch.full_like(dd, 1e30), dd).double() * 1e6 + ii.clamp_min(0).double()
    o = key.argsort(dim=1)[:, :k]
    ref_d, ref_i = dd.gather(1, o), ii.gather(1, o)
    assert torch.equal(vm.cpu(), ref_d) and torch.equal(im.cpu(), ref_i)


@pytest.mark.parametrize("chunk_rows", [0, 4096])
def test_retrieve_host_equals_device_path(ops, sbir_lib, chunk_rows, dbg):
    """Host-buffer entry point == device path, also when the gallery is uploaded and scored in
    several chunks (option host_chunk_rows forces 5 chunks here; production chunks are 1 GiB)."""
    from art_sbir_b200 import _binding as B
    if chunk_rows:
        dbg("host_chunk_rows", chunk_rows)
    nq, ng, d, k = 300, 20000, 512, 10
    Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=8)
    q, g = Q.bfloat16().pin_memory(), G.bfloat16().pin_memory()
    od = torch.empty(nq, k).pin_memory()
    oi = torch.empty(nq, k, dtype=torch.int64).pin_memory()
    orank = torch.empty(nq, dtype=torch.int64).pin_memory()
    unc = ctypes.c_int32(-1)
    B.check(sbir_lib.sbir_retrieve_host(q.data_ptr(), nq, g.data_ptr(), ng, d, B.SBIR_BF16, B.SBIR_EUCLIDEAN, k,
                                        pos.data_ptr(), od.data_ptr(), oi.data_ptr(), orank.data_ptr(), ctypes.byref(unc)),
            "sbir_retrieve_host")
    v, i, r = ops.pairwise_topk(q.cuda(), g.cuda(), k, "euclidean", pos_index=pos.cuda())
    assert torch.equal(i.cpu(), oi) and torch.equal(v.cpu(), od) and torch.equal(r.cpu(), orank) and unc.value == 0
    sbir_lib.sbir_release_host_staging()


def test_retrieve_host_shard_mode_with_a_short_last_chunk(ops, sbir_lib, dbg):
    """Several gallery partitions (few query tiles per worker) cannot be fed in pieces: the uploaded chunks are
    scored as shards and merged (K4).  The short last chunk is planned with MORE partitions than the long ones —
    the workspace must fit the largest layout (regression: it used to be sized for the longest chunk)."""
    from art_sbir_b200 import _binding as B
    dbg("host_chunk_rows", 4096)
    nq, ng, d, k = 20000, 9000, 64, 10
    Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=18, beta=0.3)
    q, g = Q.bfloat16().pin_memory(), G.bfloat16().pin_memory()
    od = torch.empty(nq, k).pin_memory()
    oi = torch.empty(nq, k, dtype=torch.int64).pin_memory()
    orank = torch.empty(nq, dtype=torch.int64).pin_memory()
    unc = ctypes.c_int32(-1)
    B.check(sbir_lib.sbir_retrieve_host(q.data_ptr(), nq, g.data_ptr(), ng, d, B.SBIR_BF16, B.SBIR_EUCLIDEAN, k, pos.data_ptr(),
                                        od.data_ptr(), oi.data_ptr(), orank.data_ptr(), ctypes.byref(unc)), "sbir_retrieve_host")
    v, i, r = ops.pairwise_topk(q.cuda(), g.cuda(), k, "euclidean", pos_index=pos.cuda())
    assert torch.equal(i.cpu(), oi) and torch.equal(v.cpu(), od) and torch.equal(r.cpu(), orank)
    sbir_lib.sbir_release_host_staging()


def test_retrieve_host_streams_chunks_into_one_pass(ops, sbir_lib, dbg):
    """With enough query tiles for a single gallery partition the uploaded chunks are FED to one
    retrieval pass (the distance kernel continues the same candidate lists from launch to launch).
    Small chunk steps (option k1_chunk_mb) and 8192-row uploads force four feeds here; the result must
    equal the device path bit for bit, ranks included, for top-10 and top-100."""
    from art_sbir_b200 import _binding as B
    dbg("k1_chunk_mb", 1)
    dbg("host_chunk_rows", 8192)
    nq, ng, d = 24000, 30000, 64
    Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=12, beta=0.3)
    pos[::97] = -1                                        # some queries without a positive
    q, g = Q.bfloat16().pin_memory(), G.bfloat16().pin_memory()
    for k in (10, 100):
        od = torch.empty(nq, k).pin_memory()
        oi = torch.empty(nq, k, dtype=torch.int64).pin_memory()
        orank = torch.empty(nq, dtype=torch.int64).pin_memory()
        unc = ctypes.c_int32(-1)
        B.check(sbir_lib.sbir_retrieve_host(q.data_ptr(), nq, g.data_ptr(), ng, d, B.SBIR_BF16, B.SBIR_EUCLIDEAN, k,
                                            pos.data_ptr(), od.data_ptr(), oi.data_ptr(), orank.data_ptr(), ctypes.byref(unc)),
                "sbir_retrieve_host")
        v, i, r = ops.pairwise_topk(q.cuda(), g.cuda(), k, "euclidean", pos_index=pos.cuda())
        assert torch.equal(i.cpu(), oi) and torch.equal(v.cpu(), od) and torch.equal(r.cpu(), orank) and unc.value == 0
        assert int((orank == ng).sum()) == int((pos < 0).sum())
    sbir_lib.sbir_release_host_staging()


def test_retrieve_host_streamed_fp32_with_escalation(ops, sbir_lib, dbg):
    """Streamed feeds followed by the device-gated escalation pass: collapsed fp32 embeddings (a large
    common component) uploaded in several chunks — the first pass (fed chunk by chunk) cannot certify
    them, the centred 3xTF32 pass over the now-resident gallery must, and the host entry point must
    return what the device path returns."""
    from art_sbir_b200 import _binding as B
    dbg("k1_chunk_mb", 1)
    dbg("host_chunk_rows", 4096)
    nq, ng, d, k = 24000, 14000, 64, 10
    Q0, G0, pos = O.synthetic_embeddings(nq, ng, d, seed=15, beta=0.3)
    base = 3.0 * torch.rand(1, d, generator=torch.Generator().manual_seed(2))
    q, g = (base + 0.02 * Q0).contiguous().pin_memory(), (base + 0.02 * G0).contiguous().pin_memory()
    od = torch.empty(nq, k).pin_memory()
    oi = torch.empty(nq, k, dtype=torch.int64).pin_memory()
    orank = torch.empty(nq, dtype=torch.int64).pin_memory()
    unc = ctypes.c_int32(-1)
    B.check(sbir_lib.sbir_retrieve_host(q.data_ptr(), nq, g.data_ptr(), ng, d, B.SBIR_F32, B.SBIR_EUCLIDEAN, k,
                                        pos.data_ptr(), od.data_ptr(), oi.data_ptr(), orank.data_ptr(), ctypes.byref(unc)),
            "sbir_retrieve_host")
    v, i, r, u = ops.pairwise_topk(q.cuda(), g.cuda(), k, "euclidean", pos_index=pos.cuda(), return_uncertified=True)
    assert torch.equal(i.cpu(), oi) and torch.equal(v.cpu(), od) and torch.equal(r.cpu(), orank)
    assert unc.value == int(u.item()) and unc.value <= nq // 50 + 4
    # and both agree with the reference formula evaluated directly (spot check on 200 queries)
    sel = torch.arange(0, nq, 120)
    dd = ((q[sel, None, :] - g[None, :, :]) + torch.tensor(1e-6)).double().pow(2).sum(-1).sqrt().float()
    ref_v, ref_i = torch.topk(dd, k, dim=1, largest=False)
    assert torch.allclose(od[sel], ref_v, rtol=DIST_RTOL, atol=1e-7)
    for a, b in (oi[sel] != ref_i).nonzero().tolist():
        assert abs(dd[a, oi[sel][a, b]].item() - ref_v[a, b].item()) <= TIE_RTOL * abs(ref_v[a, b].item())
    sbir_lib.sbir_release_host_staging()


def test_sharded_host_path_equals_single_pass(ops, sbir_lib, dbg):
    """sbir_retrieve_host_shard: each rank's shard comes from HOST memory in chunks fed to one pass.
    Two and three shards scored one after the other on this GPU + K4 merge must equal the
    single-GPU device path; sharded_retrieve_host (world size 1) as well."""
    from art_sbir_b200 import _binding as B, sharded
    dbg("k1_chunk_mb", 1)
    dbg("host_chunk_rows", 8192)
    nq, ng, d, k = 24000, 41000, 64, 10
    Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=14, beta=0.3)
    pos[::101] = -1
    qh, gh = Q.bfloat16().pin_memory(), G.bfloat16().pin_memory()
    qc, gc, pc = qh.cuda(), gh.cuda(), pos.cuda()
    v1, i1, r1 = ops.pairwise_topk(qc, gc, k, "euclidean", pos_index=pc)
    v0, i0, r0 = sharded.sharded_retrieve_host(qh, gh, k, "euclidean", pos_index=pos, shard_offset=0, num_gallery_total=ng)
    assert torch.equal(i0, i1) and torch.equal(v0, v1) and torch.equal(r0, r1)
    own = ops.positive_distance(qc, gc, pc)                      # NaN where there is no positive
    for world in (2, 3):
        vs, is_, total = [], [], torch.zeros(nq, dtype=torch.int64, device="cuda")
        for r in range(world):
            a, b = sharded.shard_bounds(ng, world, r)
            shard = gh[a:b].contiguous().pin_memory()
            v = torch.empty(nq, k, device="cuda")
            i = torch.empty(nq, k, dtype=torch.int64, device="cuda")
            c = torch.zeros(nq, dtype=torch.int64, device="cuda")
            unc = ctypes.c_int32(-1)
            B.check(sbir_lib.sbir_retrieve_host_shard(qc.data_ptr(), nq, shard.data_ptr(), b - a, d, B.SBIR_BF16, B.SBIR_EUCLIDEAN, k, a,
                                                      own.data_ptr(), pc.data_ptr(), v.data_ptr(), i.data_ptr(), c.data_ptr(),
                                                      ctypes.byref(unc), torch.cuda.current_stream().cuda_stream),
                    "sbir_retrieve_host_shard")
            assert unc.value == 0
            vs.append(v); is_.append(i); total += c
        vm, im = ops.topk_merge(torch.stack(vs), torch.stack(is_))
        rk = torch.where(own != own, torch.full_like(total, ng), total)
        assert torch.equal(im, i1) and torch.equal(vm, v1) and torch.equal(rk, r1)
    sbir_lib.sbir_release_host_staging()


# ---------------------------------------------------- buffer overrun guards (memcheck substitute) ----
@pytest.mark.parametrize("nq,ng,d,dtype,lt,k,opts", [
    (300, 6000, 128, "float32", "euclidean", 10, {"k1_sel_bf16": 0}),        # kind::tf32, 32-entry lists
    (300, 6000, 128, "float32", "euclidean", 100, {"k1_sel_bf16": 0}),       # CTA pairs, 128-entry lists
    (300, 6000, 128, "float32", "cosine", 10, {"k1_sel_bf16": 1}),           # fp32 selected on bf16 copies
    (300, 6000, 128, "bfloat16", "euclidean", 10, {}),                       # resident-query form
    (300, 6000, 128, "bfloat16", "cosine", 100, {}),                         # owner + feeder warps
    (20000, 9000, 64, "bfloat16", "euclidean", 10, {"k1_chunk_mb": 1}),      # chunk hand-overs
    (130, 2000, 66, "float32", "euclidean", 5, {}),                          # padded tile copies
    (200, 8000, 256, "float32", "euclidean", 10, {"collapsed": 1})])         # escalation pass + fallbacks
def test_no_kernel_writes_outside_its_buffers(ops, sbir_lib, dbg, nq, ng, d, dtype, lt, k, opts):
    """compute-sanitizer is closed on this GPU pool (profiles/r02_compute_sanitizer_closed.txt), so out-of-bounds WRITES
    are hunted with guard bands instead: the workspace and every output buffer of the C-ABI call sit between 64 KB
    bands of a known byte pattern inside one allocation, sized EXACTLY as the ABI reports; after the pass every band
    must be intact and the results must still be the reference's.  Covers every operand form of the distance kernel,
    both finalize forms, the rank pool, the padded-row copies and the escalation pass."""
    from art_sbir_b200 import _binding as B
    opts = dict(opts)
    collapsed = opts.pop("collapsed", 0)
    for key, val in opts.items():
        dbg(key, val)
    tdt = getattr(torch, dtype)
    Q, G, pos = O.synthetic_embeddings(nq, ng, d, seed=nq + k, beta=0.3 if d < 512 else None)
    if collapsed:
        base = 3.0 * torch.rand(1, d, generator=torch.Generator().manual_seed(1))
        Q, G = (base + 0.02 * Q).contiguous(), (base + 0.02 * G).contiguous()
    q, g, p = Q.to(tdt).cuda(), G.to(tdt).cuda(), pos.cuda()
    dt, metric = (B.SBIR_BF16 if tdt == torch.bfloat16 else B.SBIR_F32), ops.metric_id(lt)
    ws_bytes = sbir_lib.sbir_pairwise_topk_workspace_bytes(nq, ng, d, k, dt, metric, 1)
    GUARD, PAT = 1 << 16, 0xA5
    sizes = [ws_bytes, nq * k * 4, nq * k * 8, nq * 8, 4]                      # workspace, dist, index, rank, uncertified
    offs, o = [], GUARD
    for sz in sizes:
        offs.append(o)
        o += (sz + 255) // 256 * 256 + GUARD
    arena = torch.full((o,), PAT, dtype=torch.uint8, device="cuda")
    base_ptr = arena.data_ptr()
    assert base_ptr % 256 == 0
    B.check(sbir_lib.sbir_pairwise_topk(q.data_ptr(), nq, g.data_ptr(), None, ng, d, dt, metric, k, 0, p.data_ptr(),
                                        base_ptr + offs[1], base_ptr + offs[2], base_ptr + offs[3], base_ptr + offs[4],
                                        base_ptr + offs[0], ws_bytes, torch.cuda.current_stream().cuda_stream), "sbir_pairwise_topk")
    torch.cuda.synchronize()
    host = arena.cpu()
    prev_end = 0
    for off, sz in zip(offs, sizes):
        assert (host[prev_end:off] == PAT).all(), f"guard band before offset {off} was overwritten"
        prev_end = off + sz
    assert (host[prev_end:] == PAT).all(), "guard band behind the last buffer was overwritten"
    vals = host[offs[1]:offs[1] + nq * k * 4].view(torch.float32).reshape(nq, k)
    idx = host[offs[2]:offs[2] + nq * k * 8].view(torch.int64).reshape(nq, k)
    rank = host[offs[3]:offs[3] + nq * 8].view(torch.int64)
    v2, i2, r2 = ops.pairwise_topk(q, g, k, lt, pos_index=p)                    # the ordinary route: same bits
    assert torch.equal(vals, v2.cpu()) and torch.equal(idx, i2.cpu()) and torch.equal(rank, r2.cpu())
    sel = torch.arange(0, nq, max(1, nq // 40))
    Qo, Go = Q.to(tdt).float(), G.to(tdt).float()
    if collapsed:   # ground truth = the reference formula element by element in fp32, long sum in fp64
        dd = ((Qo[sel, None, :] - Go[None, :, :]) + torch.tensor(1e-6)).double().pow(2).sum(-1).sqrt().float()
        assert torch.equal(idx[sel], torch.topk(dd, k, dim=1, largest=False).indices)
    else:
        ref_v, ref_i = O.pairwise_topk_batched(Qo[sel], Go, k, lt)
        dist_rows = [O.distances(Qo[i:i + 1], Go, lt) for i in sel.tolist()]
        assert_topk_matches(vals[sel], idx[sel], ref_v, ref_i, dist_rows)


# ------------------------------------------------------------ BASELINE-size property checks ----
def _device_clustered(nq, ng, d, dtype, seed=1234):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    C = max(125, ng // 80)
    beta = 0.06 if d >= 2048 else 0.12
    cent = torch.randn(C, d, device="cuda", generator=gen)
    noise = torch.randn(ng, d, device="cuda", generator=gen)
    cls = torch.arange(ng, device="cuda") % C
    G = (cent[cls] + noise).to(dtype)
    pos = torch.randperm(ng, device="cuda", generator=gen)[:nq]
    Q = (cent[cls[pos]] + beta * noise[pos] + torch.randn(nq, d, device="cuda", generator=gen)).to(dtype)
    return Q.contiguous(), G.contiguous(), pos


@pytest.mark.parametrize("nq,ng,d,dtype,k", [(1000, 10000, 2048, torch.float32, 10),      # BASELINE cfg1
                                              (12500, 75000, 2048, torch.float32, 100),    # BASELINE cfg3 (kind::tf32 tiles)
                                              (12500, 75000, 2048, torch.float32, 10),     # cfg3 shape, top-10: fp32 selected on bf16 copies
                                              (4096, 400000, 512, torch.bfloat16, 10)])    # cfg4-shaped slice
def test_full_size_properties(ops, nq, ng, d, dtype, k):
    Q, G, pos = _device_clustered(nq, ng, d, dtype)
    vals, idx, rank, unc = ops.pairwise_topk(Q, G, k, "euclidean", pos_index=pos, return_uncertified=True)
    # (1) returned distances are the exact reference distances of the returned rows
    sub = torch.arange(0, nq, max(1, nq // 64), device="cuda")
    exact = ((Q[sub].double()[:, None, :] - G[idx[sub]].double() + 1e-6) ** 2).sum(-1).sqrt()
    assert torch.allclose(vals[sub].double(), exact, rtol=1e-5)
    # (2) ascending, unique indices
    assert (vals[:, 1:] >= vals[:, :-1]).all()
    assert (torch.sort(idx, dim=1).values[:, 1:] != torch.sort(idx, dim=1).values[:, :-1]).all()
    # (3) rank consistent with the top-k: positive inside top-k ⇔ rank < k, at the right slot
    inside = (idx == pos[:, None])
    assert torch.equal(inside.any(1), rank < k)
    assert torch.equal(inside.float().argmax(1)[rank < k], rank[rank < k])
    # (4) brute-force check of a query subsample against torch on the same device (fp64)
    sub = sub[:16]
    dm = ((Q[sub].double()[:, None, :] - G.double()[None, :, :] + 1e-6) ** 2).sum(-1).sqrt() if ng * d <= 2.1e8 else \
        torch.cdist(Q[sub].double(), G.double())
    ref_v, ref_i = dm.topk(k, dim=1, largest=False)
    mism = (ref_i != idx[sub])
    assert mism.float().mean() < 0.01
    assert torch.allclose(ref_v, vals[sub].double(), rtol=1e-5)
    dpos = dm.gather(1, pos[sub][:, None])
    assert ((dm < dpos).sum(1) - rank[sub]).abs().max() <= 1
    # (5) gallery-permutation invariance of the set of distances
    perm = torch.randperm(ng, device="cuda")
    v2, i2 = ops.pairwise_topk(Q[:256], G[perm].contiguous(), k, "euclidean")
    assert torch.equal(v2, vals[:256])

    def canon(v, i):   # exact fp32 ties are ordered by index, which the permutation changes
        o = torch.argsort(i, dim=1, stable=True)
        v, i = v.gather(1, o), i.gather(1, o)
        o = torch.argsort(v, dim=1, stable=True)
        return i.gather(1, o)
    same = canon(v2, perm[i2]) == canon(vals[:256], idx[:256])
    # a tie exactly at the k-th place may legitimately swap which of the tied rows is kept
    assert same[:, :-1].all() or (~same).sum() <= 2
