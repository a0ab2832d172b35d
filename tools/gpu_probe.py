"""GPU bring-up probe: runs each check in its own subprocess (a trapped kernel must not take
the other checks down) and writes gpurun_out/probe.log.  Usage on the GPU box:
    python tools/gpu_probe.py            # all cases
    python tools/gpu_probe.py <case>     # one case, in-process
Not part of the test-suite; torch is used here only as an on-device comparison.
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def B_set(name, value):
    """Library tuning / diagnostic switch (k1_flags needs a build with SBIR_BUILD_DIAG=1)."""
    from art_sbir_b200 import _binding
    _binding.set_debug_option(name, value)


def _data(nq, ng, d, dtype, seed=0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(nq, d, device="cuda", generator=g)
    x = torch.randn(ng, d, device="cuda", generator=g)
    c = torch.randn(8, d, device="cuda", generator=g)
    q = q + c[torch.arange(nq, device="cuda") % 8]
    x = x + c[torch.arange(ng, device="cuda") % 8]
    return q.to(dtype).contiguous(), x.to(dtype).contiguous()


def case_rowops():
    import torch
    from art_sbir_b200 import ops
    res = {}
    for dtype in (torch.float32, torch.bfloat16):
        x, y = _data(1000, 1000, 520 if dtype == torch.float32 else 520, dtype)
        n = ops.l2_normalize(x)
        ref = torch.nn.functional.normalize(x.float(), dim=1, eps=1e-8)
        res[f"l2norm_{dtype}"] = (n.float() - ref).abs().max().item()
        s = ops.row_sqnorm(x)
        res[f"sqnorm_rel_{dtype}"] = ((s - (x.double() ** 2).sum(1)).abs() / s).max().item()
        d = ops.pairwise_distance(x[:1], y, "euclidean")
        ref = torch.nn.PairwiseDistance(p=2)(x[:1].float(), y.float())
        res[f"pdist_rel_{dtype}"] = ((d - ref).abs() / ref).max().item()
        d = ops.pairwise_distance(x, y, "cosine")
        ref = 1 - torch.nn.CosineSimilarity(dim=1)(x.float(), y.float())
        res[f"cos_abs_{dtype}"] = (d - ref).abs().max().item()
    return res


def _dump_case(nq, ng, d, dtype_name, metric):
    import torch
    from art_sbir_b200 import ops
    dtype = getattr(torch, dtype_name)
    q, g = _data(nq, ng, d, dtype)
    e = ops.debug_dist_matrix(q, g, metric)
    torch.cuda.synchronize()
    qd, gd = q.double(), g.double()
    if metric == "euclidean":
        ref = (gd ** 2).sum(1)[None, :] - 2 * qd @ gd.T
        scale = ((qd ** 2).sum(1).max() + (gd ** 2).sum(1).max()).item()
    else:
        ref = -(qd @ gd.T) / gd.norm(dim=1).clamp_min(1e-8)[None, :]
        scale = qd.norm(dim=1).max().item()
    err = (e.double() - ref).abs()
    nan_frac = torch.isnan(e).float().mean().item()
    err = torch.nan_to_num(err, nan=0.0)
    worst = err.argmax().item()
    return {"shape": [nq, ng, d, dtype_name, metric], "max_abs_err": err.max().item(), "rel_to_scale": err.max().item() / scale,
            "mean_abs_err": err.mean().item(), "nan_frac": nan_frac, "worst_rc": [worst // ng, worst % ng],
            "sample": [e[0, 0].item(), ref[0, 0].item(), e[-1, -1].item(), ref[-1, -1].item()]}


def case_dump_small():
    return [_dump_case(128, 256, 32, "float32", "euclidean"), _dump_case(128, 256, 64, "float32", "euclidean"),
            _dump_case(128, 256, 64, "bfloat16", "euclidean")]


def case_dump_shapes():
    out = []
    for args in [(200, 700, 256, "float32", "euclidean"), (130, 1000, 2048, "float32", "euclidean"),
                 (300, 513, 512, "bfloat16", "euclidean"), (100, 300, 96, "float32", "cosine"),
                 (257, 2049, 1024, "bfloat16", "cosine"), (1000, 10000, 2048, "float32", "euclidean")]:
        out.append(_dump_case(*args))
    return out


def _topk_case(nq, ng, d, dtype_name, metric, k, with_rank=True):
    import torch
    from art_sbir_b200 import ops
    dtype = getattr(torch, dtype_name)
    q, g = _data(nq, ng, d, dtype, seed=1)
    pos = torch.randint(0, ng, (nq,), device="cuda")
    pos[::7] = -1
    t0 = time.time()
    if with_rank:
        vals, idx, rank, unc = ops.pairwise_topk(q, g, k, metric, pos_index=pos, return_uncertified=True)
    else:
        vals, idx, unc = ops.pairwise_topk(q, g, k, metric, return_uncertified=True)
        rank = None
    torch.cuda.synchronize()
    dt = time.time() - t0
    qd, gd = q.double(), g.double()
    if metric == "euclidean":
        dm = ((qd[:, None, :] - gd[None, :, :] + 1e-6) ** 2).sum(-1).sqrt() if nq * ng * d < 2e8 else torch.cdist(qd, gd)
    else:
        dm = 1 - (qd / qd.norm(dim=1, keepdim=True).clamp_min(1e-8)) @ (gd / gd.norm(dim=1, keepdim=True).clamp_min(1e-8)).T
    rv, ri = dm.topk(min(k, ng), dim=1, largest=False)
    kk = min(k, ng)
    idx_match = (idx[:, :kk] == ri).float().mean().item()
    val_err = ((vals[:, :kk].double() - rv).abs() / rv.abs().clamp_min(1e-12)).max().item()
    res = {"shape": [nq, ng, d, dtype_name, metric, k], "idx_match": idx_match, "val_rel_err": val_err,
           "uncertified": int(unc.item()), "first_call_s": dt}
    if with_rank:
        has = pos >= 0
        dpos = dm.gather(1, pos.clamp_min(0)[:, None])
        rr = (dm < dpos).sum(1)
        rr = torch.where(has, rr, torch.full_like(rr, ng))
        res["rank_match"] = (rank == rr).float().mean().item()
        res["rank_maxdiff"] = (rank - rr).abs().max().item()
    return res


def case_topk_small():
    return [_topk_case(100, 1000, 64, "float32", "euclidean", 10), _topk_case(128, 256, 64, "bfloat16", "euclidean", 10),
            _topk_case(5, 7, 64, "float32", "euclidean", 10), _topk_case(333, 5000, 512, "bfloat16", "cosine", 10),
            _topk_case(200, 3000, 128, "float32", "cosine", 5)]


def case_topk_mid():
    return [_topk_case(1000, 10000, 2048, "float32", "euclidean", 10), _topk_case(500, 20000, 1024, "float32", "euclidean", 100),
            _topk_case(1000, 50000, 512, "bfloat16", "euclidean", 10), _topk_case(300, 9000, 256, "bfloat16", "euclidean", 50)]


def case_triplet():
    import torch
    from art_sbir_b200 import ops
    res = {}
    for metric in ("euclidean", "cosine"):
        torch.manual_seed(0)
        a, p, n = (torch.randn(256, 2048, device="cuda", requires_grad=True) for _ in range(3))
        loss = ops.triplet_margin_loss(a, p, n, 0.2, metric)
        loss.backward()
        ga, gp, gn = a.grad.clone(), p.grad.clone(), n.grad.clone()
        a.grad = p.grad = n.grad = None
        if metric == "euclidean":
            ref = torch.nn.TripletMarginLoss(margin=0.2)(a, p, n)
        else:
            cosd = lambda x, y: 1 - torch.nn.CosineSimilarity(dim=1)(x, y)
            ref = torch.nn.TripletMarginWithDistanceLoss(margin=0.2, distance_function=cosd)(a, p, n)
        ref.backward()
        res[metric] = {"loss": loss.item(), "ref": ref.item(),
                       "grad_rel": max(((x - y.grad).abs().max() / y.grad.abs().max()).item() for x, y in ((ga, a), (gp, p), (gn, n)))}
    return res


def case_batch_hard():
    import torch
    from art_sbir_b200 import ops
    res = {}
    for metric in ("euclidean", "cosine"):
        torch.manual_seed(1)
        a, p, n = (torch.randn(256, 2048, device="cuda", requires_grad=True) for _ in range(3))
        loss, hard = ops.batch_hard_triplet_loss(a, p, n, 0.2, metric, return_indices=True)
        loss.backward()
        ga, gp, gn = a.grad.clone(), p.grad.clone(), n.grad.clone()
        a.grad = p.grad = n.grad = None
        x = torch.cat([p, n])
        if metric == "euclidean":
            dm = torch.cdist(a, x)
        else:
            dm = 1 - torch.nn.functional.normalize(a, dim=1, eps=1e-8) @ torch.nn.functional.normalize(x, dim=1, eps=1e-8).T
        posmask = torch.zeros_like(dm, dtype=torch.bool)
        posmask[torch.arange(256), torch.arange(256)] = True
        hp = dm.masked_fill(~posmask, -1e30).max(1)
        hn = dm.masked_fill(posmask, 1e30).min(1)
        ref = (0.2 + hp.values - hn.values).clamp_min(0).mean()
        ref.backward()
        res[metric] = {"loss": loss.item(), "ref": ref.item(), "hp_match": (hard[:, 0] == hp.indices).float().mean().item(),
                       "hn_match": (hard[:, 1] == hn.indices).float().mean().item(),
                       "grad_rel": max(((x_ - y.grad).abs().max() / y.grad.abs().max().clamp_min(1e-30)).item()
                                       for x_, y in ((ga, a), (gp, p), (gn, n)))}
    return res


def _clustered(nq, ng, d, dtype, seed=1234):
    """Device-side version of oracle.synthetic_embeddings (same construction, CUDA RNG)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    C = max(125, ng // 80)
    beta = 0.06 if d >= 2048 else 0.12
    cent = torch.randn(C, d, device="cuda", generator=g)
    G = torch.empty(ng, d, device="cuda", dtype=dtype)
    pos = torch.randint(0, ng, (nq,), device="cuda", generator=g)
    Q = torch.randn(nq, d, device="cuda", generator=g)
    step = 1 << 18
    for s0 in range(0, ng, step):
        n = min(step, ng - s0)
        cls = torch.arange(s0, s0 + n, device="cuda") % C
        noise = torch.randn(n, d, device="cuda", generator=g)
        G[s0:s0 + n] = (cent[cls] + noise).to(dtype)
        sel = (pos >= s0) & (pos < s0 + n)
        if sel.any():
            pi = pos[sel] - s0
            Q[sel] += cent[cls[pi]] + beta * noise[pi]
    return Q.to(dtype).contiguous(), G, pos


def _time_topk(nq, ng, d, dtype_name, k, iters=3, rank=False, clustered=True):
    import torch
    from art_sbir_b200 import ops
    dtype = getattr(torch, dtype_name)
    if clustered:
        q, g, pos = _clustered(nq, ng, d, dtype)
    else:
        q = torch.randn(nq, d, device="cuda").to(dtype)
        g = torch.randn(ng, d, device="cuda").to(dtype)
        pos = torch.randint(0, ng, (nq,), device="cuda")
    pos = pos if rank else None
    out = ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos, return_uncertified=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    pairs = nq * ng
    res = {"shape": [nq, ng, d, dtype_name, k, rank, clustered], "ms": ms, "pairs_per_s": pairs / ms * 1e3,
           "tflops": 2 * d * pairs / ms * 1e3 / 1e12, "uncertified": int(out[-1].item())}
    if rank:
        r = out[2]
        res["recall@1,10"] = [(r < 1).float().mean().item(), (r < 10).float().mean().item()]
        res["rank_max"] = int(r.max().item())
    return res


def case_time():
    return [_time_topk(1000, 10000, 2048, "float32", 10, rank=True),
            _time_topk(12500, 75000, 2048, "float32", 100),
            _time_topk(12500, 75000, 2048, "float32", 100, rank=True),
            _time_topk(12500, 75000, 2048, "float32", 10),
            _time_topk(12500, 75000, 2048, "float32", 10, rank=True),
            _time_topk(12500, 75000, 2048, "float32", 10, rank=True, clustered=False),
            _time_topk(20000, 1000000, 512, "bfloat16", 10),
            _time_topk(20000, 1000000, 512, "bfloat16", 10, rank=True),
            _time_topk(20000, 1000000, 512, "bfloat16", 100),
            _time_topk(12500, 75000, 2048, "bfloat16", 10, rank=True)]


def case_sel():
    """fp32 embeddings: selection on bf16 copies (kind::f16) vs kind::tf32 — time, uncertified queries."""
    out = []
    for shape in ((12500, 75000, 2048, 10), (12500, 75000, 2048, 100), (12500, 75000, 2048, 30), (1000, 10000, 2048, 10),
                  (20000, 200000, 1024, 10), (20000, 200000, 1024, 100), (20000, 1000000, 512, 10)):
        for sel in (0, -1):
            B_set("k1_sel_bf16", sel)
            r = _time_topk(shape[0], shape[1], shape[2], "float32", shape[3], rank=True)
            r["sel_bf16"] = sel
            out.append(r)
            r = _time_topk(shape[0], shape[1], shape[2], "float32", shape[3], rank=True, clustered=False) if shape[0] <= 1000 else None
            if r:
                r["sel_bf16"] = sel
                out.append(r)
    B_set("reset", 0)
    return out


def case_shard():
    """K1 efficiency on the per-GPU shard sizes of cfg4 (one GPU, no collectives): is the 8-GPU loss inherent to
    1.25M-row shards (list warm-up, tails), and does the chunk size matter there?"""
    out = []
    for ng in (1_250_000, 2_500_000):
        for mb in (0, 24, 96):
            B_set("k1_chunk_mb", mb)
            r = _time_topk(100_000, ng, 512, "bfloat16", 10, rank=True)
            r["chunk_mb"] = mb or 48
            out.append(r)
    B_set("reset", 0)
    return out


def _bench(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def case_bw():
    """Bandwidth kernels (K5 normalise, norms, K4 merge) and the cfg2 triplet step."""
    import torch
    from art_sbir_b200 import ops
    res = {}
    x = torch.randn(10_000_000, 512, device="cuda", dtype=torch.bfloat16)
    ms = _bench(lambda: ops.l2_normalize(x), 10)
    res["l2_normalize 10Mx512 bf16"] = {"ms": ms, "GB/s": 2 * x.numel() * 2 / ms / 1e6}
    ms = _bench(lambda: ops.row_sqnorm(x), 10)
    res["row_sqnorm 10Mx512 bf16"] = {"ms": ms, "GB/s": x.numel() * 2 / ms / 1e6}
    del x
    y = torch.randn(2_000_000, 1024, device="cuda")
    ms = _bench(lambda: ops.l2_normalize(y), 10)
    res["l2_normalize 2Mx1024 fp32"] = {"ms": ms, "GB/s": 2 * y.numel() * 4 / ms / 1e6}
    ms = _bench(lambda: torch.nn.functional.normalize(y, dim=1, eps=1e-8), 10)
    res["torch normalize 2Mx1024 fp32 (library, for reference)"] = {"ms": ms, "GB/s": 2 * y.numel() * 4 / ms / 1e6}
    del y
    d = torch.sort(torch.rand(8, 100_000, 10, device="cuda"), dim=2).values
    i = torch.randint(0, 10_000_000, (8, 100_000, 10), device="cuda")
    ms = _bench(lambda: ops.topk_merge(d, i), 20)
    res["topk_merge 8x100kx10"] = {"ms": ms, "GB/s": (d.numel() * 12 + 100_000 * 10 * 12) / ms / 1e6}
    a, p, n = (torch.randn(256, 2048, device="cuda", requires_grad=True) for _ in range(3))

    def ours():
        loss = ops.triplet_margin_loss(a, p, n, 0.2, "euclidean")
        loss.backward()

    def lib():
        loss = torch.nn.TripletMarginLoss(margin=0.2)(a, p, n)
        loss.backward()

    def bh():
        loss = ops.batch_hard_triplet_loss(a, p, n, 0.2, "euclidean")
        loss.backward()
    res["cfg2 triplet fwd+bwd 256x2048 (ours, us)"] = _bench(ours, 50) * 1e3
    res["cfg2 triplet fwd+bwd 256x2048 (torch library on the same GPU, us)"] = _bench(lib, 50) * 1e3
    res["cfg2 batch-hard fwd+bwd 256x(512)x2048 (ours, us)"] = _bench(bh, 50) * 1e3
    return res


def case_mainloop():
    """K1 with the epilogue switched off (SBIR_K1_FLAGS=8, results meaningless) vs the full kernel:
    how much of the pass is the TMA+MMA mainloop alone."""
    import ctypes
    import torch
    from art_sbir_b200 import _binding as B, ops
    lib = B.load()
    out = {}
    for nq, ng, d, dt, k in [(20000, 1000000, 512, "bfloat16", 10), (12500, 75000, 2048, "float32", 10),
                             (12500, 75000, 2048, "bfloat16", 10), (20000, 500000, 1024, "bfloat16", 10)]:
        q, g, pos = _clustered(nq, ng, d, getattr(torch, dt))
        res = {}
        for rep in range(2):
            for flags in (0, 8):
                B_set("k1_flags", int(flags))
                ops.pairwise_topk(q, g, k, "euclidean")
                torch.cuda.synchronize()
                lib.sbir_profile_enable(1)
                ms, n, l = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
                lib.sbir_profile_collect(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(l))
                for _ in range(5):
                    ops.pairwise_topk(q, g, k, "euclidean")
                torch.cuda.synchronize()
                lib.sbir_profile_collect(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(l))
                lib.sbir_profile_enable(0)
                k1 = ms.value / max(1, n.value)
                res.setdefault("full" if flags == 0 else "mainloop_only", []).append(
                    {"k1_ms": round(k1, 3), "tflops": round(2 * d * nq * ng / k1 / 1e9, 1)})
        out[f"{nq}x{ng}x{d} {dt}"] = res
    B_set("k1_flags", 0)
    return out


def case_ab():
    """Same-process A/B of K1 switches (SBIR_K1_FLAGS is read at every launch): 16 = no chunk screen."""
    import ctypes
    import torch
    from art_sbir_b200 import _binding as B, ops
    lib = B.load()
    out = {}
    for nq, ng, d, dt, k, rank in [(20000, 1000000, 512, "bfloat16", 10, True), (20000, 1000000, 512, "bfloat16", 10, False),
                                   (12500, 75000, 2048, "float32", 10, True), (12500, 75000, 2048, "float32", 100, True)]:
        q, g, pos = _clustered(nq, ng, d, getattr(torch, dt))
        p = pos if rank else None
        res = {}
        for rep in range(3):
            for flags in (0, 16):
                B_set("k1_flags", int(flags))
                ops.pairwise_topk(q, g, k, "euclidean", pos_index=p)
                torch.cuda.synchronize()
                lib.sbir_profile_enable(1)
                ms, n, l = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
                lib.sbir_profile_collect(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(l))
                for _ in range(5):
                    ops.pairwise_topk(q, g, k, "euclidean", pos_index=p)
                torch.cuda.synchronize()
                lib.sbir_profile_collect(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(l))
                lib.sbir_profile_enable(0)
                res.setdefault(flags, []).append(round(ms.value / max(1, n.value), 3))
        out[f"{nq}x{ng}x{d} {dt} k={k} rank={rank}"] = res
    B_set("k1_flags", 0)
    return out


def case_peaks():
    import torch
    res = {}
    for name, dtype, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True), ("fp32", torch.float32, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        n = 8192
        a = torch.randn(n, n, device="cuda", dtype=dtype)
        b = torch.randn(n, n, device="cuda", dtype=dtype)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name + "_tflops"] = 2 * n ** 3 / best * 1e3 / 1e12
    x = torch.empty(1 << 30, device="cuda", dtype=torch.bfloat16)
    y = torch.empty_like(x)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y.copy_(x)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res["copy_gbs"] = 2 * x.numel() * 2 / best * 1e3 / 1e9
    return res


def case_host():
    import torch
    import ctypes
    from art_sbir_b200 import _binding as B, ops
    nq, ng, d, k = 500, 20000, 512, 10
    q = torch.randn(nq, d).bfloat16().pin_memory()
    g = torch.randn(ng, d).bfloat16().pin_memory()
    pos = torch.randint(0, ng, (nq,), dtype=torch.int64)
    od = torch.empty(nq, k).pin_memory()
    oi = torch.empty(nq, k, dtype=torch.int64).pin_memory()
    orank = torch.empty(nq, dtype=torch.int64).pin_memory()
    unc = ctypes.c_int32(0)
    lib = B.load()
    B.check(lib.sbir_retrieve_host(q.data_ptr(), nq, g.data_ptr(), ng, d, B.SBIR_BF16, 0, k, pos.data_ptr(),
                                   od.data_ptr(), oi.data_ptr(), orank.data_ptr(), ctypes.byref(unc)), "retrieve_host")
    v, i, r = ops.pairwise_topk(q.cuda(), g.cuda(), k, "euclidean", pos_index=pos.cuda())
    return {"idx_equal": bool((i.cpu() == oi).all()), "val_equal": bool((v.cpu() == od).all()),
            "rank_equal": bool((r.cpu() == orank).all()), "unc": unc.value}


def case_escal_debug():
    """Why do ranks differ on cancellation-heavy data?  Compare against an on-device fp64 evaluation."""
    import torch
    from art_sbir_b200 import ops
    g = torch.Generator().manual_seed(5)
    nq, ng, d = 300, 6000, 256
    G = (8.0 + 0.5 * torch.randn(ng, d, generator=g)).cuda()
    Q = (8.0 + 0.5 * torch.randn(nq, d, generator=g)).cuda()
    pos = torch.randint(0, ng, (nq,), generator=g).cuda()
    vals, idx, rank, unc = ops.pairwise_topk(Q, G, 10, "euclidean", pos_index=pos, return_uncertified=True)
    # exact per-element fp32 arithmetic, fp64 accumulation (what the library's exact kernels do)
    diff = (Q[:, None, :] - G[None, :, :]) + torch.tensor(1e-6, device="cuda", dtype=torch.float32)
    d64 = (diff.double() ** 2).sum(-1).sqrt()
    dpos = d64.gather(1, pos[:, None])
    r64 = (d64 < dpos).sum(1)
    d32 = d64.float()
    dpos32 = d32.gather(1, pos[:, None])
    r32 = ((d32 < dpos32) | ((d32 == dpos32) & (torch.arange(ng, device="cuda")[None, :] < pos[:, None]))).sum(1)
    # real-number formula on fp64-cast inputs (the oracle's fp64 mode)
    dr = ((Q.double()[:, None, :] - G.double()[None, :, :] + 1e-6) ** 2).sum(-1).sqrt()
    rr = (dr < dr.gather(1, pos[:, None])).sum(1)
    out = {"unc": int(unc.item()), "match_r64": (rank == r64).float().mean().item(), "match_r32canon": (rank == r32).float().mean().item(),
           "match_oracle_fp64": (rank == rr).float().mean().item(), "r64_vs_oracle": (r64 == rr).float().mean().item(),
           "maxdiff_r64": (rank - r64).abs().max().item()}
    bad = (rank != r64).nonzero().flatten()[:5]
    det = []
    for i in bad.tolist():
        gap = (d64[i] - dpos[i]).abs()
        gap[pos[i]] = 1e9
        near = torch.topk(gap, 4, largest=False).values.tolist()
        det.append({"q": i, "ours": int(rank[i]), "r64": int(r64[i]), "nearest_gaps_in_d": near, "dpos": dpos[i].item()})
    out["detail"] = det
    return out


def case_centred():
    """Collapsed / non-centred fp32 embeddings (a large common component, small differences — what an
    untrained encoder or post-ReLU features give): exactness and time of the centred escalation pass."""
    import torch
    from art_sbir_b200 import ops
    out = []
    for (nq, ng, d, scale, structured) in ((2000, 25000, 1024, 0.02, True), (2000, 25000, 1024, 0.02, False),
                                           (12500, 75000, 2048, 0.05, True), (12500, 75000, 2048, 0.05, False)):
        q, g, pos = _clustered(nq, ng, d, torch.float32)
        if not structured:
            pos = torch.randint(0, ng, (nq,), device="cuda")
        base = 3.0 * torch.rand(1, d, device="cuda")
        q = (base + scale * q).contiguous()
        g = (base + scale * g).contiguous()
        for metric in ("euclidean", "cosine"):
            res = ops.pairwise_topk(q, g, 10, metric, pos_index=pos, return_uncertified=True)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            ops.pairwise_topk(q, g, 10, metric, pos_index=pos)
            ev[1].record()
            ops.pairwise_topk(q, g, 10, metric, pos_index=pos)
            ev[2].record()
            torch.cuda.synchronize()
            row = {"shape": [nq, ng, d, scale, structured, metric], "ms": min(ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])),
                   "uncertified": int(res[-1].item())}
            if nq * ng <= 60_000_000:   # exact check on device (fp32 elementwise, fp64 sum), chunked over queries
                bad_r = bad_i = 0
                for a in range(0, nq, 20):
                    qq = q[a:a + 20]
                    if metric == "euclidean":
                        diff = (qq[:, None, :] - g[None, :, :]) + torch.tensor(1e-6, device="cuda")
                        dd = (diff.double() ** 2).sum(-1).sqrt().float()
                    else:
                        qn = qq / qq.norm(dim=1, keepdim=True).clamp_min(1e-8)
                        gn = g / g.norm(dim=1, keepdim=True).clamp_min(1e-8)
                        dd = (1.0 - (qn[:, None, :] * gn[None, :, :]).double().sum(-1)).float()
                    pp = pos[a:a + 20]
                    dp = dd.gather(1, pp[:, None])
                    r = ((dd < dp) | ((dd == dp) & (torch.arange(ng, device="cuda")[None, :] < pp[:, None]))).sum(1)
                    bad_r += int((r != res[2][a:a + 20]).sum())
                    ti = torch.topk(dd, 10, dim=1, largest=False).indices
                    bad_i += int((ti.sort(1).values != res[1][a:a + 20].sort(1).values).any(1).sum())
                row["rank_mismatch"] = bad_r
                row["topk_set_mismatch"] = bad_i
            out.append(row)
    return out


def _k1_ms(q, g, k, pos=None, iters=4):
    import ctypes
    import torch
    from art_sbir_b200 import _binding as B, ops
    lib = B.load()
    ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos)
    torch.cuda.synchronize()
    lib.sbir_profile_enable(1)
    ms, n, l = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    lib.sbir_profile_collect(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(l))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos)
    ev[1].record()
    torch.cuda.synchronize()
    lib.sbir_profile_collect(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(l))
    lib.sbir_profile_enable(0)
    return round(ms.value / iters, 3), round(ev[0].elapsed_time(ev[1]) / iters, 3)


def case_k100():
    """Where does top-100 lose time against top-10?  K1 / whole-call ms for k=10 vs k=100 as the gallery
    grows (list warm-up is a fixed cost per (query tile, partition) chain; steady-state insertions grow
    with log N) and with one chunk per partition (no list park / restore between chunks)."""
    import torch
    out = {}
    for nq, ng, d, dt in [(12500, 75000, 2048, "float32"), (12500, 300000, 2048, "float32"), (12500, 75000, 2048, "bfloat16"),
                          (20000, 1000000, 512, "bfloat16")]:
        q, g, pos = _clustered(nq, ng, d, getattr(torch, dt))
        for chunk in ("12", "100000"):
            B_set("k1_chunk_mb", int(chunk))
            for k in (10, 100):
                out[f"{nq}x{ng}x{d} {dt} k={k} chunkMB={chunk}"] = _k1_ms(q, g, k)
    B_set("k1_chunk_mb", 0)
    return out


def _graph_us(fn, replays=50):
    """Device time of fn() in microseconds: captured once into a CUDA graph and replayed back to back,
    so host-side launch overhead (Python, ctypes, allocator) is not in the measurement."""
    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / replays * 1e3


def case_small():
    """Device time (CUDA-graph replay) of the small kernels: K4 merge, cfg2 triplet / batch-hard steps,
    retrieval metrics, cfg1-sized retrieval."""
    import torch
    from art_sbir_b200 import _binding as B, ops
    lib = B.load()
    st = lambda: torch.cuda.current_stream().cuda_stream
    res = {}
    for lists, nq, k in ((8, 100_000, 10), (8, 100_000, 100), (2, 100_000, 10)):
        d = torch.sort(torch.rand(lists, nq, k, device="cuda"), dim=2).values.contiguous()
        i = torch.randint(0, 10_000_000, (lists, nq, k), device="cuda")
        od, oi = torch.empty(nq, k, device="cuda"), torch.empty(nq, k, dtype=torch.int64, device="cuda")
        us = _graph_us(lambda: B.check(lib.sbir_topk_merge(d.data_ptr(), i.data_ptr(), lists, 0, 0, nq, k, od.data_ptr(), oi.data_ptr(), st()), "merge"))
        res[f"topk_merge {lists}x{nq}x{k}"] = {"us": us, "GB/s": (d.numel() * 12 + nq * k * 12) / us / 1e3}
    a, p, n = (torch.randn(256, 2048, device="cuda") for _ in range(3))
    loss, per_row = torch.empty((), device="cuda"), torch.empty(256, device="cuda")
    ga, gp, gn = (torch.empty_like(a) for _ in range(3))
    for metric, name in ((B.SBIR_EUCLIDEAN, "euclidean"), (B.SBIR_COSINE, "cosine")):
        us = _graph_us(lambda: B.check(lib.sbir_triplet_margin_loss(a.data_ptr(), p.data_ptr(), n.data_ptr(), 256, 2048, 0.2, metric,
                                                                   loss.data_ptr(), per_row.data_ptr(), ga.data_ptr(), gp.data_ptr(),
                                                                   gn.data_ptr(), st()), "triplet"))
        res[f"cfg2 triplet fwd+bwd 256x2048 {name}"] = {"us": us, "GB/s": 6 * a.numel() * 4 / us / 1e3}
        ws = torch.zeros(lib.sbir_batch_hard_workspace_bytes(256, 2048), dtype=torch.uint8, device="cuda")
        hard = torch.empty(256, 2, dtype=torch.int64, device="cuda")
        us = _graph_us(lambda: B.check(lib.sbir_batch_hard_triplet_loss(a.data_ptr(), p.data_ptr(), n.data_ptr(), 256, 2048, 0.2, metric, None, None,
                                                                       loss.data_ptr(), hard.data_ptr(), ga.data_ptr(), gp.data_ptr(), gn.data_ptr(),
                                                                       ws.data_ptr(), ws.numel(), st()), "batch_hard"))
        torch.cuda.synchronize()
        tw = ws[-(((8 + 8 * 512) * 8 + 255) // 256 * 256):].view(torch.int64).cpu().numpy()
        t = [int(x) for x in tw[:4]]                                         # phase boundaries seen by CTA 0 (globaltimer ns)
        if t[0] == 0:                                                        # product build: stamps exist only with SBIR_BUILD_DIAG=1
            res[f"cfg2 batch-hard fwd+bwd 256x512x2048 {name}"] = {"us": us}
            continue
        raw = tw[8:8 + 8 * 296].reshape(296, 8)
        stg = (raw - t[0]) / 1e3                                              # per-CTA stage stamps, us after CTA 0's start
        ok = stg[:256].copy()                                                # CTAs that own an anchor pass every stage
        import numpy as np
        ok[:, 5] = raw[:256, 5]                                              # slot 5 carries the band size, not a stamp
        slow = np.argsort(-(ok[:, 4] - ok[:, 0]))[:8]
        res[f"cfg2 batch-hard slowest CTAs {name} (cta, scan_us, band_us, exact_us, select+grad_us, band_entries)"] = [
            [int(c), round(ok[c, 1] - ok[c, 0], 2), round(ok[c, 2] - ok[c, 1], 2), round(ok[c, 3] - ok[c, 2], 2), round(ok[c, 4] - ok[c, 3], 2), int(ok[c, 5])] for c in slow]
        res[f"cfg2 batch-hard fwd+bwd 256x512x2048 {name}"] = {
            "us": us, "phase_us(mine,select,grad)": [(t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3],
            "stage_median_us": np.median(ok, 0).round(2).tolist(), "stage_max_us": ok.max(0).round(2).tolist(),
            "stage_min_us": ok.min(0).round(2).tolist()}
    ta, tp, tn = (t.clone().requires_grad_(True) for t in (a, p, n))

    def torch_step():
        ta.grad = tp.grad = tn.grad = None
        torch.nn.functional.triplet_margin_loss(ta, tp, tn, margin=0.2).backward()
    res["cfg2 triplet fwd+bwd 256x2048 torch library (same GPU)"] = {"us": _graph_us(torch_step)}

    eye = torch.zeros(256, 512, dtype=torch.bool, device="cuda")
    eye[torch.arange(256, device="cuda"), torch.arange(256, device="cuda")] = True

    def torch_batch_hard():
        # the same definition written with library ops: broadcast distance + masks + autograd
        ta.grad = tp.grad = tn.grad = None
        x = torch.cat([tp, tn])
        dm = (ta[:, None, :] - x[None, :, :] + 1e-6).norm(dim=2)
        hp = dm.masked_fill(~eye, float("-inf")).max(dim=1).values
        hn = dm.masked_fill(eye, float("inf")).min(dim=1).values
        torch.clamp_min(0.2 + hp - hn, 0).mean().backward()
    res["cfg2 batch-hard fwd+bwd 256x512x2048 torch library ops (same GPU)"] = {"us": _graph_us(torch_batch_hard, replays=20)}
    q, g, pos = _clustered(1000, 10000, 2048, torch.float32)
    res["cfg1 retrieval 1000x10000x2048 fp32 top-10+rank"] = {"us": _graph_us(lambda: ops.pairwise_topk(q, g, 10, "euclidean", pos_index=pos))}
    return res


def case_cfg1():
    """Small retrievals (the reference's own 1k x 10k evaluation): A/B of the tile / epilogue forms.
    Device time by CUDA-graph replay, K1's own time from the library's profiling hooks, results compared with the default form."""
    import ctypes
    import torch
    from art_sbir_b200 import _binding as B, ops
    lib = B.load()
    out = {}
    for nq, ng, d, k in ((1000, 10000, 2048, 10), (1000, 10000, 2048, 100), (2000, 40000, 2048, 10), (1000, 10000, 512, 10)):
        q, g, pos = _clustered(nq, ng, d, torch.float32)
        base = None
        for name, opts in (("default", {}), ("sel_bf16", {"k1_sel_bf16": 1}), ("sel_tf32", {"k1_sel_bf16": 0})):
            B_set("reset", 0)
            for o, v in opts.items():
                B_set(o, v)
            try:
                r = ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos, return_uncertified=True)
                torch.cuda.synchronize()
                us = _graph_us(lambda: ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos))
                lib.sbir_profile_enable(1)
                k1_ms, k1_n, launches = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
                lib.sbir_profile_collect(ctypes.byref(k1_ms), ctypes.byref(k1_n), ctypes.byref(launches))
                for _ in range(5):
                    ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos)
                torch.cuda.synchronize()
                lib.sbir_profile_collect(ctypes.byref(k1_ms), ctypes.byref(k1_n), ctypes.byref(launches))
                lib.sbir_profile_enable(0)
                rec = {"us": round(us, 1), "k1_us": round(k1_ms.value / 5 * 1e3, 1), "launches": launches.value // 5, "uncertified": int(r[3].item())}
                if base is None:
                    base = r
                else:
                    rec["same_as_default"] = bool(torch.equal(r[1], base[1]) and torch.equal(r[2], base[2]) and torch.equal(r[0], base[0]))
            except Exception as e:  # noqa: BLE001
                rec = {"error": str(e)[:200]}
            out[f"{nq}x{ng}x{d} k={k} {name}"] = rec
    B_set("reset", 0)
    return out


def case_qorder():
    """Does finalize's candidate re-scoring (gathers of 8 KB gallery rows) profit from L2 when queries with overlapping
    candidate sets run next to each other?  cfg3 with the queries in generator order vs grouped by class of their positive."""
    import torch
    from art_sbir_b200 import ops
    out = {}
    for k in (100, 10):
        q, g, pos = _clustered(12500, 75000, 2048, torch.float32)
        C = max(125, 75000 // 80)
        for name, order in (("generator order", None), ("grouped by class", torch.argsort(pos % C, stable=True)), ("random order", torch.randperm(12500, device="cuda"))):
            qq, pp = (q, pos) if order is None else (q[order].contiguous(), pos[order].contiguous())
            us = _graph_us(lambda: ops.pairwise_topk(qq, g, k, "euclidean", pos_index=pp), replays=20)
            out[f"cfg3 k={k} {name}"] = round(us, 1)
    return out


def case_bands():
    """cfg4 (and its 8-GPU shard) with the query tiles walked in L2 bands: time per pass for band count x chunk size."""
    import torch
    from art_sbir_b200 import ops
    out = []
    for ng in (10_000_000, 1_250_000):
        q, g, pos = _clustered(100_000, ng, 512, torch.bfloat16)
        base = None
        for bands, chunk in ((-1, 0), (0, 0), (2, 0), (3, 0), (4, 0), (3, 24), (4, 24), (6, 24), (2, 96)):
            B_set("reset", 0)
            B_set("k1_bands", bands)
            if chunk:
                B_set("k1_chunk_mb", chunk)
            r = ops.pairwise_topk(q, g, 10, "euclidean", pos_index=pos, return_uncertified=True)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            for i in range(3):
                ops.pairwise_topk(q, g, 10, "euclidean", pos_index=pos)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
            rec = {"gallery": ng, "bands": bands, "chunk_mb": chunk or 48, "ms": [round(x, 2) for x in ms], "uncertified": int(r[3].item())}
            if base is None:
                base = r
            else:
                rec["same_result"] = bool(torch.equal(r[0], base[0]) and torch.equal(r[1], base[1]) and torch.equal(r[2], base[2]))
            out.append(rec)
        del q, g, pos
    B_set("reset", 0)
    return out


def case_l2hints():
    """cfg4, its 8-GPU shard and cfg4 with K = 100: A/B of one library option inside one process (same box, same thermal
    state).  `python tools/gpu_probe.py l2hints` sweeps the L2 eviction hints of the resident-query form (k1_l2_hints bits);
    `python tools/gpu_probe.py l2hints <option>` alternates 0 / 1 of another switch, e.g. k1_q_early."""
    import torch
    from art_sbir_b200 import ops
    opt = sys.argv[2] if len(sys.argv) > 2 else "k1_l2_hints"
    values = (0, 1, 2, 4, 3, 5, 6, 7, 0, 1, 2, 4, 3, 5, 6, 7) if opt == "k1_l2_hints" else (0, 1, 0, 1)
    if len(sys.argv) > 3:                      # explicit values, e.g. `l2hints k1_chunk_mb 48,96`: alternated twice
        values = tuple(int(v) for v in sys.argv[3].split(",")) * 2
    out = []
    for ng, k in ((10_000_000, 10), (1_250_000, 10), (10_000_000, 100)):
        q, g, pos = _clustered(100_000, ng, 512, torch.bfloat16)
        base = None
        for value in values:
            B_set("reset", 0)
            B_set(opt, value)
            r = ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos, return_uncertified=True)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            for i in range(3):
                ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos)
                ev[i + 1].record()
            torch.cuda.synchronize()
            rec = {"gallery": ng, "k": k, opt: value, "ms": [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(3)], "uncertified": int(r[3].item())}
            if base is None:
                base = r
            else:
                rec["same_result"] = bool(torch.equal(r[0], base[0]) and torch.equal(r[1], base[1]) and torch.equal(r[2], base[2]))
            out.append(rec)
        del q, g, pos
    B_set("reset", 0)
    return out


def case_pairsel():
    """Wide fp32 rows selected on bf16 copies run the all-shared-memory form (the query tile does not fit tensor memory),
    which re-reads the query tile for every gallery tile: single-CTA tiles vs CTA pairs (each CTA loads half the gallery tile)."""
    import torch
    from art_sbir_b200 import ops
    out = []
    for nq, ng, d, k in ((12500, 75000, 2048, 10), (12500, 75000, 2048, 100), (20000, 200000, 1024, 10), (20000, 200000, 2048, 10), (12500, 75000, 2048, 30)):
        q, g, pos = _clustered(nq, ng, d, torch.float32)
        base = None
        for pair in (0, 2, 0, 2):
            B_set("reset", 0)
            B_set("k1_pair", pair)
            r = ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos, return_uncertified=True)
            torch.cuda.synchronize()
            us = _graph_us(lambda: ops.pairwise_topk(q, g, k, "euclidean", pos_index=pos), replays=20)
            k1 = _k1_ms(q, g, k, pos)
            rec = {"shape": [nq, ng, d, k], "k1_pair": pair, "us": round(us, 1), "k1": k1, "uncertified": int(r[3].item())}
            if base is None:
                base = r
            else:
                rec["same_result"] = bool(torch.equal(r[0], base[0]) and torch.equal(r[1], base[1]) and torch.equal(r[2], base[2]))
            out.append(rec)
        del q, g, pos
    B_set("reset", 0)
    return out


def case_k100_mainloop():
    """k=100 (cap 128: 3 operand stages) with the epilogue switched off: is it the mainloop?"""
    import torch
    out = {}
    for nq, ng, d, dt in [(20000, 1000000, 512, "bfloat16"), (12500, 75000, 2048, "float32")]:
        q, g, pos = _clustered(nq, ng, d, getattr(torch, dt))
        for k in (10, 100):
            for flags in ("0", "8"):
                B_set("k1_flags", int(flags))
                out[f"{nq}x{ng}x{d} {dt} k={k} flags={flags}"] = _k1_ms(q, g, k)
    B_set("k1_flags", 0)
    return out


def case_ab2():
    """Interleaved A/B of K1 variants in one process (SBIR_K1_FLAGS bits): median of several rounds."""
    import statistics
    import torch
    out = {}
    for nq, ng, d, dt, k in [(12500, 75000, 2048, "float32", 100), (20000, 1000000, 512, "bfloat16", 100),
                             (12500, 75000, 2048, "float32", 10), (20000, 1000000, 512, "bfloat16", 10)]:
        q, g, pos = _clustered(nq, ng, d, getattr(torch, dt))
        res = {}
        for rnd in range(5):
            for flags in ("0", "32", "8"):
                B_set("k1_flags", int(flags))
                res.setdefault(flags, []).append(_k1_ms(q, g, k, iters=3)[0])
        out[f"{nq}x{ng}x{d} {dt} k={k}"] = {f: [round(statistics.median(v), 3), round(min(v), 3)] for f, v in res.items()}
    B_set("k1_flags", 0)
    return out


def case_diag():
    """Where does the MMA issuer wait?  Per-CTA cycle counters (SBIR_K1_FLAGS=64) for the headline-like
    shapes: share of the MMA loop spent waiting for a free accumulator (epilogue-bound) or for operands
    (TMA / L2-bound)."""
    import ctypes
    import numpy as np
    import torch
    from art_sbir_b200 import _binding as B, ops
    lib = B.load()
    out = {}
    buf = (ctypes.c_uint64 * (148 * 8))()
    shapes = [(20000, 1000000, 512, "bfloat16", 10), (20000, 1000000, 512, "bfloat16", 100),
              (12500, 75000, 2048, "float32", 10), (12500, 75000, 2048, "float32", 100)]
    if os.environ.get("SBIR_DIAG_SMALL"):
        shapes = [(1000, 10000, 2048, "float32", 10), (1000, 10000, 512, "bfloat16", 10)]
    for nq, ng, d, dt, k in shapes:
        q, g, pos = _clustered(nq, ng, d, getattr(torch, dt))
        for flags in ("64", "72"):   # 72 = 64 + 8: epilogue switched off (mainloop alone)
            B_set("k1_flags", int(flags))
            ops.pairwise_topk(q, g, k, "euclidean")
            torch.cuda.synchronize()
            lib.sbir_debug_k1_diag(buf, 148 * 8)
            ops.pairwise_topk(q, g, k, "euclidean")
            torch.cuda.synchronize()
            lib.sbir_debug_k1_diag(buf, 148 * 8)
            a = np.array(list(buf), dtype=np.float64).reshape(148, 8)
            a = a[a[:, 2] > 0]                      # CTA pairs: only the leaders record
            loop = a[:, 2].mean()
            out[f"{nq}x{ng}x{d} {dt} k={k} flags={flags}"] = {
                "mma_loop_Mclk": round(loop / 1e6, 2), "wait_acc_frac": round(a[:, 0].mean() / loop, 3),
                "wait_operands_frac": round(a[:, 1].mean() / loop, 3), "epi_wait_acc_full_frac": round(a[:, 3].mean() / loop, 3),
                "per_kblock_clk": {"loop": round(loop / a[:, 7].mean(), 1), "wait_operands": round(a[:, 1].mean() / a[:, 7].mean(), 1),
                                   "issue_4_mma": round(a[:, 5].mean() / a[:, 7].mean(), 1), "commit": round(a[:, 6].mean() / a[:, 7].mean(), 1)}}
    B_set("k1_flags", 0)
    return out


def case_l2n():
    """l2_normalize of 10M x 512 bf16 (and 4M x 256 fp32): device time for the current SBIR_L2N_VARIANT."""
    import torch
    from art_sbir_b200 import ops
    res = {"variant": os.environ.get("SBIR_L2N_VARIANT", "default")}
    x = torch.randn(10_000_000, 512, device="cuda", dtype=torch.bfloat16)
    y = ops.l2_normalize(x)
    ref = torch.nn.functional.normalize(x[:1000].float(), dim=1)
    res["max_err"] = (y[:1000].float() - ref).abs().max().item()
    us = _graph_us(lambda: ops.l2_normalize(x), 10)
    res["10Mx512 bf16"] = {"us": us, "GB/s": 2 * x.numel() * 2 / us / 1e3}
    del x, y
    z = torch.randn(4_000_000, 256, device="cuda")
    us = _graph_us(lambda: ops.l2_normalize(z), 10)
    res["4Mx256 fp32"] = {"us": us, "GB/s": 2 * z.numel() * 4 / us / 1e3}
    return res


CASES = ["rowops", "dump_small", "dump_shapes", "topk_small", "topk_mid", "triplet", "batch_hard", "host", "peaks", "time"]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        out = globals()["case_" + sys.argv[1]]()
        print("RESULT " + json.dumps(out))
        sys.exit(0)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.log", "w") as log:
        for c in CASES:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True, timeout=400)
                tail = (r.stdout[-6000:] + "\n--stderr--\n" + r.stderr[-3000:]) if r.returncode else \
                    "\n".join(l for l in r.stdout.splitlines() if l.startswith("RESULT"))
                msg = f"=== {c} rc={r.returncode} {time.time() - t0:.1f}s\n{tail}\n"
            except subprocess.TimeoutExpired:
                msg = f"=== {c} TIMEOUT\n"
            log.write(msg)
            log.flush()
            print(msg)

