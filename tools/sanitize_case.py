"""Small-shape tour of every kernel form for compute-sanitizer (one --tool per gpurun call):

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py

Covers the three operand forms of the distance kernel (all-smem kind::tf32 and kind::f16, resident-query,
CTA pairs), small and large lists (two lists per row, owner + feeder warps), fp32 embeddings selected on
bf16 copies, chunk hand-overs inside one launch and across streamed feeds (host-buffer entry point), the
escalation pass, finalize (block / warp-per-query / radix-select forms), rank resolution, K4 merge, the row
kernels, the gallery append and the cooperative batch-hard kernel.  Results are compared with the oracle so a
"clean" report is about a program that computed the right thing.  The device-side watchdog is switched off
(sanitizers slow kernels down by orders of magnitude)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from art_sbir_b200 import _binding as B, ops  # noqa: E402
from oracle import sbir_oracle as O  # noqa: E402


def check_topk(name, Q, G, pos, k, lt, dtype, **opts):
    for key, val in opts.items():
        B.set_debug_option(key, val)
    try:
        q, g = Q.to(dtype).cuda(), G.to(dtype).cuda()
        v, i, r, u = ops.pairwise_topk(q, g, k, lt, pos_index=pos.cuda(), return_uncertified=True)
        torch.cuda.synchronize()
        ref_v, ref_i = O.pairwise_topk_batched(Q.to(dtype).float(), G.to(dtype).float(), k, lt)
        ref_r = O.rank_of_positive_batched(Q.to(dtype).float(), G.to(dtype).float(), pos, lt)
        same_i = (i.cpu() == ref_i).float().mean().item()
        same_r = (r.cpu() == ref_r).float().mean().item()
        assert same_i > 0.995 and same_r > 0.97 and torch.allclose(v.cpu(), ref_v, rtol=1e-3, atol=1e-6), (name, same_i, same_r)
        print(f"ok {name}: indices {same_i:.4f} ranks {same_r:.4f} uncertified {int(u.item())}", flush=True)
    finally:
        B.set_debug_option("reset")
        B.set_debug_option("watchdog_cycles", 0)


def main():
    lib = B.load()
    B.set_debug_option("watchdog_cycles", 0)
    Q, G, pos = O.synthetic_embeddings(300, 6000, 128, seed=7, beta=0.3)
    pos[::11] = -1
    f32, bf16 = torch.float32, torch.bfloat16
    check_topk("tf32 all-smem, 32-entry lists", Q, G, pos, 10, "euclidean", f32, k1_sel_bf16=0)
    check_topk("tf32 CTA pairs, 128-entry lists", Q, G, pos, 100, "euclidean", f32, k1_sel_bf16=0)
    check_topk("tf32 single CTAs, 128-entry lists, owner+feeder", Q, G, pos, 100, "cosine", f32, k1_sel_bf16=0, k1_pair=1)
    check_topk("fp32 selected on bf16 copies", Q, G, pos, 10, "euclidean", f32, k1_sel_bf16=1)
    check_topk("bf16 resident-query, two lists per row", Q, G, pos, 10, "euclidean", bf16)
    check_topk("bf16 all-smem, two lists per row", Q, G, pos, 10, "cosine", bf16, k1_qres=0)
    check_topk("bf16 128-entry lists, owner+feeder", Q, G, pos, 100, "euclidean", bf16)
    check_topk("bf16 CTA pairs", Q, G, pos, 30, "euclidean", bf16, k1_pair=2, k1_qres=0)
    # chunk hand-overs inside one launch: 1 MB chunk steps, a single partition (many query tiles)
    Q2, G2, pos2 = O.synthetic_embeddings(20000, 9000, 64, seed=8, beta=0.3)
    B.set_debug_option("k1_chunk_mb", 1)
    v, i, r = ops.pairwise_topk(Q2.bfloat16().cuda(), G2.bfloat16().cuda(), 10, "euclidean", pos_index=pos2.cuda())
    # ... and across streamed feeds of the host-buffer entry point
    B.set_debug_option("host_chunk_rows", 4096)
    qh, gh = Q2.bfloat16().pin_memory(), G2.bfloat16().pin_memory()
    od, oi = torch.empty(20000, 10).pin_memory(), torch.empty(20000, 10, dtype=torch.int64).pin_memory()
    orank = torch.empty(20000, dtype=torch.int64).pin_memory()
    unc = ctypes.c_int32(-1)
    B.check(lib.sbir_retrieve_host(qh.data_ptr(), 20000, gh.data_ptr(), 9000, 64, B.SBIR_BF16, B.SBIR_EUCLIDEAN, 10, pos2.data_ptr(),
                                   od.data_ptr(), oi.data_ptr(), orank.data_ptr(), ctypes.byref(unc)), "sbir_retrieve_host")
    assert torch.equal(oi, i.cpu()) and torch.equal(orank, r.cpu()) and torch.equal(od, v.cpu())
    lib.sbir_release_host_staging()
    B.set_debug_option("reset")
    B.set_debug_option("watchdog_cycles", 0)
    print("ok chunk hand-over in one launch == streamed feeds from host buffers", flush=True)
    # escalation pass (collapsed embeddings -> centred 3xTF32) and the brute-force fallbacks
    base = 3.0 * torch.rand(1, 128, generator=torch.Generator().manual_seed(1))
    Qc, Gc = (base + 0.02 * Q).contiguous(), (base + 0.02 * G).contiguous()
    v, i, r, u = ops.pairwise_topk(Qc.cuda(), Gc.cuda(), 10, "euclidean", pos_index=pos.cuda(), return_uncertified=True)
    dd = ((Qc[:, None, :] - Gc[None, :, :]) + torch.tensor(1e-6)).double().pow(2).sum(-1).sqrt().float()
    assert torch.equal(i.cpu(), torch.topk(dd, 10, dim=1, largest=False).indices)
    print(f"ok escalation pass on collapsed embeddings: uncertified {int(u.item())}", flush=True)
    # K4 merge, metrics, row kernels, gallery append
    d = torch.sort(torch.rand(8, 500, 10), dim=2).values
    ix = torch.randint(0, 100000, (8, 500, 10))
    vm, im = ops.topk_merge(d.cuda(), ix.cuda())
    assert torch.allclose(vm.cpu(), d.permute(1, 0, 2).reshape(500, -1).sort(dim=1).values[:, :10])
    ops.retrieval_metrics(r)
    x = torch.randn(1000, 520)
    assert torch.allclose(ops.l2_normalize(x.cuda()).cpu(), O.l2_normalize(x), atol=1e-6)
    assert torch.allclose(ops.l2_normalize(x.bfloat16().cuda()).float().cpu(), O.l2_normalize(x.bfloat16().float()), atol=4e-3)
    buf = ops.GalleryBuffer(1000, 520, torch.bfloat16, normalize=True)
    buf.append(x[:600].cuda()); buf.append(x[600:].cuda())
    assert torch.allclose(buf.sqnorm.cpu(), (buf.rows.cpu().double() ** 2).sum(1).float(), rtol=1e-6)
    assert torch.allclose(ops.pairwise_distance(x[:1].cuda(), x.cuda()).cpu(), O.euclidean_distance(x[:1], x), rtol=1e-5)
    print("ok merge / metrics / l2_normalize / gallery_append / row-wise distance", flush=True)
    # losses
    g = torch.Generator().manual_seed(3)
    a, p, n = (torch.randn(100, 520, generator=g) for _ in range(3))
    p = a + 0.8 * p
    for lt in ("euclidean", "cosine"):
        A, P, N = (t.clone().cuda().requires_grad_(True) for t in (a, p, n))
        loss = ops.triplet_margin_loss(A, P, N, 0.2, lt)
        loss.backward()
        assert abs(loss.item() - O.triplet_margin_loss(a, p, n, 0.2, lt).item()) < 1e-3
        A, P, N = (t.clone().cuda().requires_grad_(True) for t in (a, p, n))
        lb, hard = ops.batch_hard_triplet_loss(A, P, N, 0.2, lt, labels=(torch.arange(100) // 3).cuda(), return_indices=True)
        lb.backward()
        ref, hpi, hni = O.batch_hard_triplet_loss(a, p, n, 0.2, lt, torch.arange(100) // 3)
        assert abs(lb.item() - ref.item()) < 1e-3 and torch.equal(hard[:, 0].cpu(), hpi) and torch.equal(hard[:, 1].cpu(), hni)
    torch.cuda.synchronize()
    print("ok triplet / batch-hard (cooperative tcgen05 kernel)", flush=True)
    print("SANITIZE_CASE_DONE", flush=True)


if __name__ == "__main__":
    main()
