"""Mints the golden fixtures of tests/golden/ by running the UNMODIFIED reference
(Peer222/art-sbir at /root/reference, imported through oracle/ref_import.py) on seeded inputs.

    python tests/golden/make_golden.py        # only works where /root/reference exists

The reference ships no tests or known-answer vectors for this path, so these
files are the pin: tests/test_oracle.py checks oracle/sbir_oracle.py against them on every
box, tests/test_gpu_parity.py checks the CUDA path against them on the GPU.
Outputs (small, committed):
    retrieval_<loss>.npz   Q, G, sketch/gallery file names, ranks from get_ranking_position,
                           top-10 (index, value) from get_topk_images for every query
    process_inference_<loss>.json   the dict returned by process_inference (identity encoder)
    triplet.npz            a, p, n, logits, labels; losses + grads of nn.TripletMarginLoss,
                           TripletMarginWithDistanceLoss(cosine), TripletMarginLoss_with_classification{,2}
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_import, sbir_oracle  # noqa: E402

OUT = Path(__file__).resolve().parent


def retrieval_case(ref_utils, ref_inference, loss_type, nq=48, ng=400, dim=64):
    Q, G, pos = sbir_oracle.synthetic_embeddings(nq, ng, dim, seed=1234, beta=0.3, num_classes=8)
    # gallery: sorted unique photo paths (data_preparation.py:30-31); sketches named id-number.png
    image_paths = [Path(f"photos/class{(i % 7):02d}/n{i:06d}.jpg") for i in range(ng)]
    sketch_paths = [Path(f"sketches/n{int(pos[i]):06d}-{(i % 5) + 1}.png") for i in range(nq)]
    sketch_paths[5] = Path("sketches/n999999-1.png")  # no photo for this one (inference.py:39-41)
    ranks, top_idx, top_val = [], [], []
    path_to_idx = {str(p): i for i, p in enumerate(image_paths)}
    for i in range(nq):
        q = Q[i:i + 1]
        ranks.append(ref_inference.get_ranking_position(sketch_paths[i], image_paths, q, G, loss_type))
        tk = ref_inference.get_topk_images(10, image_paths, q, G, loss_type)
        top_idx.append([path_to_idx[p] for p, _ in tk])
        top_val.append([v for _, v in tk])
    np.savez_compressed(OUT / f"retrieval_{loss_type}.npz", Q=Q.numpy(), G=G.numpy(), pos=pos.numpy(),
                        image_paths=np.array([str(p) for p in image_paths]),
                        sketch_paths=np.array([str(p) for p in sketch_paths]),
                        ranks=np.array(ranks, dtype=np.int64), top_idx=np.array(top_idx, dtype=np.int64),
                        top_val=np.array(top_val, dtype=np.float32))

    # process_inference with an identity encoder
    class Identity(torch.nn.Module):
        def forward(self, x):
            return x

    class DS:
        pass
    ds = DS()
    ds.sketch_paths = sketch_paths
    ds.__class__.__len__ = lambda self: nq
    inf_ds = DS.__new__(DS)
    inf_ds.image_paths = image_paths
    loader = [(Q[i:i + 1],) for i in range(nq)]
    from timeit import default_timer as timer

    class InfDS:
        def __init__(self):
            self.image_paths = image_paths

        def __len__(self):
            return ng
    stats = ref_inference.process_inference(Identity(), ds, InfDS(), loader, G, timer(), False, loss_type)
    stats.pop("inference_time")
    with open(OUT / f"process_inference_{loss_type}.json", "w") as f:
        json.dump(stats, f, indent=1)


def triplet_case(ref_utils, B=24, D=64, C1=7, C2=5):
    g = torch.Generator().manual_seed(77)
    a, p, n = (torch.randn(B, D, generator=g) for _ in range(3))
    p = a + 0.7 * p  # positives correlated with anchors so some hinges are inactive
    cs, cp = torch.randn(B, C1, generator=g), torch.randn(B, C1, generator=g)
    cs2, cp2 = torch.randn(B, C2, generator=g), torch.randn(B, C2, generator=g)
    l1 = torch.randint(0, C1, (B,), generator=g)
    l2 = torch.randint(0, C2, (B,), generator=g)
    out = dict(a=a.numpy(), p=p.numpy(), n=n.numpy(), cs=cs.numpy(), cp=cp.numpy(), cs2=cs2.numpy(),
               cp2=cp2.numpy(), l1=l1.numpy(), l2=l2.numpy())

    def run(name, fn):
        A, P, N = (t.clone().requires_grad_(True) for t in (a, p, n))
        loss = fn(A, P, N)
        loss.backward()
        out[name + "_loss"] = loss.detach().numpy()
        out[name + "_ga"], out[name + "_gp"], out[name + "_gn"] = A.grad.numpy(), P.grad.numpy(), N.grad.numpy()

    m = ref_utils.MARGIN
    run("tml_euclid", torch.nn.TripletMarginLoss(margin=m))  # train.py:169
    run("tmdl_cosine", ref_utils.nn.TripletMarginWithDistanceLoss(margin=m, distance_function=ref_utils.cosine_distance))  # train.py:175
    run("tmdl_euclid", ref_utils.nn.TripletMarginWithDistanceLoss(margin=m, distance_function=ref_utils.euclidean_distance))
    wc = ref_utils.TripletMarginLoss_with_classification(margin=m)  # train.py:166
    run("wc_euclid", lambda A, P, N: wc(A, P, N, cs, cp, l1))
    wcc = ref_utils.TripletMarginLoss_with_classification(margin=m, distance_f=ref_utils.cosine_distance)  # train.py:172
    run("wc_cosine", lambda A, P, N: wcc(A, P, N, cs, cp, l1))
    wc2 = ref_utils.TripletMarginLoss_with_classification2(margin=m, classification_weight=0, classification_weight2=0.2)  # train.py:168
    run("wc2_euclid", lambda A, P, N: wc2(A, P, N, cs, cp, cs2, cp2, l1, l2))
    out["margin"] = np.float32(m)
    np.savez_compressed(OUT / "triplet.npz", **out)


if __name__ == "__main__":
    if not ref_import.available():
        sys.exit("reference checkout not found; golden fixtures can only be minted in the build container")
    torch.set_num_threads(4)
    ref_utils, ref_inference = ref_import.load()
    for lt in ("euclidean", "cosine"):
        retrieval_case(ref_utils, ref_inference, lt)
    triplet_case(ref_utils)
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")) + sorted(p.name for p in OUT.glob("*.json")))
