#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native retrieval hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Metric (BASELINE.json): query×gallery pairs/sec (distance + top-10) with rank / recall@K.
Default workload `cfg4`: 100 000 queries × 10 000 000 gallery embeddings, 512-d bf16, top-10 +
rank of the positive, synthetic clustered embeddings, gallery row-sharded over
the N GPUs (strong scaling: the problem is fixed, each rank scores N_g/N rows, one all-gather +
merge + all-reduce exchange).  A step is one full retrieval pass.  Other workloads
(`cfg1`, `cfg3`, `cfg3k10`) are the fp32 2048-d configs of BASELINE.json, single GPU.

One JSON line on stdout (rank 0).  `value` = pairs/s with inputs resident in HBM; `e2e` = the
same pass from pinned HOST buffers through the C ABI (H2D + compute + D2H in the timed region);
`roofline` = tensor-pipe fraction of the distance kernel; `cpu_baseline` = the reference's
per-query CPU path (oracle port) on a bounded sample, timed on this box's cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (num_q, num_g, dim, dtype, k, description)
    "cfg4": (100_000, 10_000_000, 512, "bfloat16", 10,
             "BASELINE cfg4: 100k queries x 10M gallery, 512-d bf16, top-10 + rank, gallery-sharded"),
    "cfg4k100": (100_000, 10_000_000, 512, "bfloat16", 100,
                 "BASELINE cfg4 with K=100: 100k queries x 10M gallery, 512-d bf16, top-100 + rank, gallery-sharded"),
    "cfg3": (12_500, 75_000, 2048, "float32", 100, "BASELINE cfg3: 12.5k x 75k, 2048-d fp32, top-100 + rank"),
    "cfg3k10": (12_500, 75_000, 2048, "float32", 10, "cfg3 shape with top-10 + rank (2048-d fp32 target line)"),
    "cfg1": (1_000, 10_000, 2048, "float32", 10, "BASELINE cfg1: 1k x 10k, 2048-d fp32, top-10 + rank"),
}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken in [t_begin, t_end] (wall clock); if the timed region was
        too short to catch one, of all samples since start() (warm-up included; flagged)."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for ts, r in self.rows if t_begin is None or (t_begin <= ts <= t_end + 0.15)]
        window = "timed region"
        if not rows:
            rows, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than the sampling period)"
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------ data ----
def make_shard(num_q, num_g, dim, dtype, row0, row1, device, seed=1234, centroids=None):
    """Rows [row0,row1) of the seeded clustered gallery + ALL queries: the SURVEY.md §8(d) generator
    (class centroids ~N(0,I); gallery row j = centroid[j mod C] + noise_j; positives = randperm(N)[:Q];
    query i = centroid of its positive + beta·noise_pos + N(0,I), beta 0.06 at 2048-d / 0.12 at 512-d),
    evaluated on device in chunks with per-chunk seeds so every rank sees the same global data.
    One stated deviation: §8(d) fixes C = 125 classes at its 1k × 10k calibration point (80 rows per class,
    recall@1/10 ≈ 0.47/0.86); here C = max(125, N/80) keeps those 80 rows per class — and that recall
    profile — at every gallery size.  With C = 125 at N = 10M every class has 80 000 rows, ~all positives rank in
    the thousands and recall@10 ≈ 0, which exercises nothing of the rank path; `--centroids 125` runs it."""
    import torch
    C = max(125, num_g // 80) if not centroids else int(centroids)
    beta = 0.06 if dim >= 2048 else 0.12
    gen = torch.Generator(device=device).manual_seed(seed)
    cent = torch.randn(C, dim, device=device, generator=gen)
    if num_q <= num_g:
        pos = torch.randperm(num_g, device=device, generator=gen)[:num_q].contiguous()
    else:
        pos = torch.randint(0, num_g, (num_q,), device=device, generator=gen)
    Q = torch.randn(num_q, dim, device=device, generator=gen)
    G = torch.empty(row1 - row0, dim, device=device, dtype=dtype)
    chunk = 1 << 18
    for c0 in range(0, num_g, chunk):
        c1 = min(c0 + chunk, num_g)
        sel = (pos >= c0) & (pos < c1)
        need_rows = not (c1 <= row0 or c0 >= row1)
        if not need_rows and not bool(sel.any()):
            continue
        cg = torch.Generator(device=device).manual_seed(seed + 1 + c0 // chunk)
        noise = torch.randn(c1 - c0, dim, device=device, generator=cg)
        cls = torch.arange(c0, c1, device=device) % C
        if need_rows:
            a, b = max(c0, row0), min(c1, row1)
            G[a - row0:b - row0] = (cent[cls[a - c0:b - c0]] + noise[a - c0:b - c0]).to(dtype)
        if bool(sel.any()):
            pi = pos[sel] - c0
            Q[sel] += cent[cls[pi]] + beta * noise[pi]
    return Q.to(dtype).contiguous(), G, pos


# --------------------------------------------------------- in-run parity (untimed) ----
def sampled_parity(Q, Gs, pos, r0, k, vals, idx, rank0, world, dist, n_sample=32):
    """§8(d): results compared to the oracle IN THE SAME RUN, at the full size.  A sample of queries is
    brute-forced on the device in fp64 with the reference formula ‖q − g + 1e-6‖₂ (utils.py:42) over ALL
    gallery rows (every rank scans its shard; lists / counts are gathered), and the returned top-k indices
    and ranks are compared.  Index differences are only accepted between distances closer than 1e-4 relative
    (north_star); ranks must lie inside the band of fp32-level ties around d(q, pos) (12·2^-24 relative,
    tests/test_gpu_parity.py) and are also counted for exact equality.  Torch is the CHECKER here, never timed."""
    import torch
    num_q, num_g_local = Q.shape[0], Gs.shape[0]
    dev = Q.device
    S = min(n_sample, num_q)
    sel = torch.linspace(0, num_q - 1, S, device=dev).round().long()
    qs = Q[sel].double()
    ps = pos[sel]
    # d(q, positive): the owner rank evaluates it, everyone gets it
    mine = (ps >= r0) & (ps < r0 + num_g_local)
    dpos = torch.zeros(S, dtype=torch.float64, device=dev)
    if bool(mine.any()):
        gp = Gs[(ps[mine] - r0)].double()
        dpos[mine] = ((qs[mine] - gp + 1e-6) ** 2).sum(1).sqrt()
    if world > 1:
        dist.all_reduce(dpos)
    tau = 12 * 2.0 ** -24 * dpos
    kk = min(k, max(num_g_local, 1))
    best_d = torch.full((S, k), float("inf"), dtype=torch.float64, device=dev)
    best_i = torch.full((S, k), -1, dtype=torch.int64, device=dev)
    lo = torch.zeros(S, dtype=torch.int64, device=dev)
    hi = torch.zeros(S, dtype=torch.int64, device=dev)
    chunk = 1 << 19
    for c0 in range(0, num_g_local, chunk):
        g = Gs[c0:c0 + chunk].double()
        for s_ in range(S):
            d = ((g - qs[s_] + 1e-6) ** 2).sum(1).sqrt()
            lo[s_] += (d < dpos[s_] - tau[s_]).sum()
            hi[s_] += (d <= dpos[s_] + tau[s_]).sum()
            v, ix = torch.topk(d, min(kk, d.numel()), largest=False)
            cat_d = torch.cat([best_d[s_], v])
            cat_i = torch.cat([best_i[s_], ix + c0 + r0])
            o = torch.argsort(cat_d, stable=True)[:k]
            best_d[s_], best_i[s_] = cat_d[o], cat_i[o]
        del g
    if world > 1:
        dist.all_reduce(lo)
        dist.all_reduce(hi)
        all_d = [torch.empty_like(best_d) for _ in range(world)]
        all_i = [torch.empty_like(best_i) for _ in range(world)]
        dist.all_gather(all_d, best_d)
        dist.all_gather(all_i, best_i)
        cat_d, cat_i = torch.cat(all_d, dim=1), torch.cat(all_i, dim=1)
        o = torch.argsort(cat_d, dim=1, stable=True)[:, :k]
        best_d, best_i = cat_d.gather(1, o), cat_i.gather(1, o)
    hi = hi - 1  # the positive itself lies inside the band
    ours_i, ours_v, ours_r = idx[sel], vals[sel].double(), rank0[sel]
    differ = ours_i != best_i
    rel = (ours_v - best_d).abs() / best_d.clamp_min(1e-30)
    tie = differ & (rel <= 1e-4)
    return {"checked": int(S), "oracle": "fp64 brute force of the reference formula over all gallery rows, on device, untimed",
            "topk_entries": int(S * k), "topk_index_mismatches": int((differ & ~tie).sum().item()),
            "topk_tie_swaps": int(tie.sum().item()), "max_rel_dist_err": float(rel[~differ].max().item()) if bool((~differ).any()) else None,
            "rank_mismatches": int(((ours_r < lo) | (ours_r > hi)).sum().item()),
            "rank_ambiguous_by_fp32_ties": int((lo != hi).sum().item()),
            "mismatches": int((differ & ~tie).sum().item() + ((ours_r < lo) | (ours_r > hi)).sum().item())}


# ----------------------------------------------------------------- CPU reference ----
def workload_config(name, world, centroids=0):
    """The `config` object of the JSON line: the WORKLOAD only (no results, no timing details), identical in the
    b200 arm and in the reference arm."""
    num_q, num_g, dim, dtype_name, k, desc = WORKLOADS[name]
    return {"workload": name, "description": desc, "num_q": num_q, "num_g": num_g, "dim": dim, "k": k, "metric_fn": "euclidean",
            "input_dtype": dtype_name, "sharding": f"gallery rows over {world} GPU(s)",
            "generator": "SURVEY 8(d) clustered generator, C = %s class centroids (80 gallery rows per class as at 8(d)'s 1k x 10k "
                         "calibration point), positives = randperm" % (centroids or max(125, num_g // 80))}


def cpu_reference_sample(num_g_sample, dim, nq_sample, seconds_hint=20.0, dtype_name="float32"):
    """The reference's per-query path (inference.py:44,49,52: PairwiseDistance broadcast + topk(len(G)) + nonzero) via
    the oracle port, all host threads, on a bounded SAMPLE OF THE SAME WORKLOAD: clustered embeddings from the same
    generator family (oracle.synthetic_embeddings, 80 gallery rows per class), rounded to the workload's input type and
    evaluated in fp32 as the reference would (its CPU ops take fp32; bf16 inputs = fp32 math on bf16-rounded values)."""
    import torch
    from oracle import sbir_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    Q, G, pos = O.synthetic_embeddings(nq_sample, num_g_sample, dim, seed=1234, num_classes=max(125, num_g_sample // 80))
    tdt = getattr(torch, dtype_name)
    Q, G = Q.to(tdt).float(), G.to(tdt).float()
    O.ranking_position(Q[:1], G, int(pos[0]), "euclidean")  # warm-up (allocator, threads)
    t0 = time.perf_counter()
    done = 0
    for i in range(nq_sample):
        # what the reference does for EVERY query (inference.py:113 -> :44,:49,:52); its full sort
        # subsumes the top-10 (get_topk_images, :62-65, only runs for 10 sampled queries)
        O.ranking_position(Q[i:i + 1], G, int(pos[i]), "euclidean")
        done += 1
        if time.perf_counter() - t0 > seconds_hint:
            break
    dt = time.perf_counter() - t0
    return done * num_g_sample / dt, done, dt, torch.get_num_threads()


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the
    reference is pure Python/torch and /root/reference does not exist on the GPU box), on this arm's
    `config`, `metric`, `unit`; every step is a bounded sample of that workload."""
    num_q, num_g, dim, dtype, k, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    ng_s = min(num_g, 500_000 if dim <= 512 else 75_000)
    times, pairs = [], 0
    per_step_q = 4
    for s in range(args.warmup + args.steps):
        v, done, dt, threads = cpu_reference_sample(ng_s, dim, per_step_q, seconds_hint=30.0, dtype_name=dtype)
        if s >= args.warmup:
            times.append(dt)
            pairs += done * ng_s
    value = pairs / sum(times)
    sample = (f"{per_step_q} queries x {ng_s} gallery rows per step of the same clustered workload ({dim}-d, {dtype}-rounded, fp32 math), "
              "PairwiseDistance + topk(N) + nonzero per query (inference.py:44,49,52)")
    line = {"impl": "reference", "metric": "query x gallery pairs/sec (distance + top-%d + rank)" % k, "value": value,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, world, args.centroids),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------- measurement ----
def measure_tf32_peak(dev):
    """Dense TF32 tensor throughput of THIS GPU, measured in the run (MEASURED_PEAKS.json carries bf16 only):
    cuBLAS fp32 matmul 8192^3 with TF32 allowed — best of 10 (burst: for a kernel timed alone) and back to back for
    ~1.5 s (sustained, under the power cap: for kernels timed inside a long loop).  TFLOP/s."""
    import torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(1500.0 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        return {"burst": flops / (best * 1e-3) / 1e12, "sustained": flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def time_retrieval(step, steps, warmup, barrier, lib, dev, flush, sampler=None):
    """W warm-up steps, then `steps` timed steps (CUDA events on the current stream, L2 flushed between steps
    when `flush`), K1 launch times and launch count from the library's profiling hooks.
    Returns (ms_per_step, k1_ms_per_step, launches_per_run, last output, clocks or None)."""
    import torch
    for _ in range(warmup):
        out = step()
    barrier()
    lib.sbir_profile_enable(1)
    k1_ms, k1_n, launches = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    lib.sbir_profile_collect(ctypes.byref(k1_ms), ctypes.byref(k1_n), ctypes.byref(launches))  # reset counters
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_begin = time.time()
    for a, b in evs:
        if flush is not None:
            flush.fill_(1)
        a.record()
        out = step()
        b.record()
    barrier()
    clocks = sampler.stop(t_begin, time.time()) if sampler is not None else None
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    lib.sbir_profile_collect(ctypes.byref(k1_ms), ctypes.byref(k1_n), ctypes.byref(launches))
    lib.sbir_profile_enable(0)
    # K1 device time per step: fp32 workloads enqueue a second, device-gated K1 launch (the 3xTF32
    # escalation pass) that returns at once when the first pass certified everything, so the sum of
    # the K1 launches of a step is the time of the one that did the work
    return ms_total / steps, k1_ms.value / max(1, steps), int(launches.value), out, clocks


def tiles_are_bf16(lib, num_q, num_g, dim, k, dtype_is_bf16):
    """Element type the distance kernel's tensor-core tiles read for this problem (bf16 embeddings, or fp32
    embeddings selected on their bf16 copies — sbir_debug_plan out[12]): decides which measured peak applies."""
    out = (ctypes.c_int32 * 13)()
    lib.sbir_debug_plan(num_q, num_g, dim, k, 1 if dtype_is_bf16 else 0, 148, out)
    return bool(out[12])


def k1_roofline(dim, num_q, rows, tile_bf16, k1_ms, ms_per_step, peaks, tf32_peak, traffic=None, long_run=None):
    """Tensor roofline of the distance kernel.  Denominator: the measured cuBLAS rate of the element type the tiles
    read — the SUSTAINED figure when the kernel was timed inside a long stretch of tensor work (one launch above
    100 ms, or a timed loop of about a second: the 1 kW power cap sets the clock), else the burst figure."""
    flops = 2.0 * dim * num_q * rows
    achieved = flops / (k1_ms * 1e-3) / 1e12 if k1_ms > 0 else None
    if long_run is None:
        long_run = k1_ms > 100
    if tile_bf16:
        peak = peaks["bf16_sustained"] if long_run else peaks["bf16"]
        note = ("bf16 dense (kind::f16 tiles), sustained, " if long_run else "bf16 dense (kind::f16 tiles), burst, ") + peaks["source"]
        burst = peaks["bf16"]
    else:
        peak = tf32_peak["sustained"] if long_run else tf32_peak["burst"]
        burst = tf32_peak["burst"]
        note = ("tf32 dense, cuBLAS fp32 8192^3 with TF32 allowed, measured in this run, "
                + ("back to back for ~1.5 s (sustained)" if long_run else "best of 10 (burst)") + "; kind::tf32 runs at half the bf16 rate")
    return {"bound": "tensor", "kernel": "dist_topk_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_note": note,
            "frac_of_burst_peak": (achieved / burst) if achieved else None,
            "k1_ms_per_launch": k1_ms, "k1_share_of_step": k1_ms / ms_per_step if ms_per_step else None}


def graph_us(fn, min_seconds=0.3):
    """Device time of fn() in microseconds: captured once into a CUDA graph and replayed back to back (host
    launch overhead excluded), long enough for the clock sampler to see it."""
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
    replays, total_ms, total_n = 50, 0.0, 0
    t0 = time.time()
    while True:
        e0.record()
        for _ in range(replays):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        total_n += replays
        if time.time() - t0 > min_seconds:
            break
    return total_ms / total_n * 1e3


def extra_retrieval_workload(name, lib, dev, local_rank, peaks, tf32_peak, centroids, reuse=None):
    """One of the other BASELINE configs on this GPU, briefly: ms, pairs/s, roofline, clocks, sampled parity."""
    import torch
    from art_sbir_b200 import ops
    num_q, num_g, dim, dtype_name, k, desc = WORKLOADS[name]
    dtype = getattr(torch, dtype_name)
    if reuse is not None:
        Q, G, pos = reuse
    else:
        Q, G, pos = make_shard(num_q, num_g, dim, dtype, 0, num_g, dev, centroids=centroids)
    torch.cuda.synchronize()
    in_bytes = (Q.numel() + G.numel()) * Q.element_size()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if in_bytes < 400e6 else None

    def step():
        return ops.pairwise_topk(Q, G, k, "euclidean", pos_index=pos, return_uncertified=True)

    def barrier():
        torch.cuda.synchronize()

    # enough steps for >= ~1 s of wall clock in the timed region, so the clock record is of THIS workload
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step(); step()
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1), 0.05) + (0.2 if flush is not None else 0.0)
    steps = int(min(2000, max(3, 1000.0 / est)))
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ms, k1_ms, launches, out, clocks = time_retrieval(step, steps, 3, barrier, lib, dev, flush, sampler)
    vals, idx, rank0, unc = out
    res = {"workload": name, "description": desc, "num_q": num_q, "num_g": num_g, "dim": dim, "k": k,
           "dtype": "bf16" if dtype == torch.bfloat16 else "f32 (selected on tensor-core tiles, exact fp32/fp64 re-score)",
           "steps": steps, "ms_per_step": ms, "value": num_q * num_g / (ms * 1e-3), "unit": "pairs/s",
           "l2": "inputs larger than L2" if flush is None else "L2 flushed (512 MiB write) between timed steps",
           "gpu_launches_per_step": launches / steps,
           "tensor_tiles": "bf16 (kind::f16)" if tiles_are_bf16(lib, num_q, num_g, dim, k, dtype == torch.bfloat16) else "tf32 (kind::tf32)",
           "roofline": k1_roofline(dim, num_q, num_g, tiles_are_bf16(lib, num_q, num_g, dim, k, dtype == torch.bfloat16), k1_ms, ms, peaks, tf32_peak,
                                   long_run=k1_ms * steps > 400),   # the timed loop keeps the tensor cores busy for >= 0.4 s: sustained peak
           "clocks": clocks, "uncertified_queries": int(unc.item()),
           **{f"recall@{kk}": float((rank0 < kk).float().mean().item()) for kk in (1, 5, 10)},
           "parity": sampled_parity(Q, G, pos, 0, k, vals, idx, rank0, 1, None)}
    return res


def cfg2_workload(lib, dev, local_rank):
    """BASELINE cfg2: a/p/n [256, 2048] fp32, margin 0.2 — the reference's triplet loss fwd+bwd (train.py:169) and
    batch-hard mining + loss + gradients, device time by CUDA-graph replay, next to torch's own kernels on the
    same GPU; loss / mined indices / gradients checked against the oracle in the same run."""
    import torch
    from art_sbir_b200 import _binding as B, ops
    from oracle import sbir_oracle as O
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    g = torch.Generator().manual_seed(11)
    a_h, p_h, n_h = (torch.randn(256, 2048, generator=g) for _ in range(3))
    p_h, n_h = a_h + 0.9 * p_h, a_h + 0.9 * n_h      # positives and negatives equally far: about half the hinges active
    a, p, n = a_h.to(dev), p_h.to(dev), n_h.to(dev)
    loss, per_row = torch.empty((), device=dev), torch.empty(256, device=dev)
    ga, gp, gn = (torch.empty_like(a) for _ in range(3))
    hard = torch.empty(256, 2, dtype=torch.int64, device=dev)
    ws = torch.empty(lib.sbir_batch_hard_workspace_bytes(256, 2048), dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    t_begin = time.time()
    res = {"workload": "cfg2", "description": "BASELINE cfg2: triplet step, batch 256 anchors/pos/neg 2048-d fp32, margin 0.2, 1 GPU",
           "timing": "device time per call, CUDA-graph replay (>= 0.3 s of back-to-back replays each)"}
    res["triplet_fwd_bwd_us"] = graph_us(lambda: B.check(lib.sbir_triplet_margin_loss(
        a.data_ptr(), p.data_ptr(), n.data_ptr(), 256, 2048, 0.2, B.SBIR_EUCLIDEAN, loss.data_ptr(), per_row.data_ptr(),
        ga.data_ptr(), gp.data_ptr(), gn.data_ptr(), st()), "triplet"))
    res["triplet_GBps"] = 6 * a.numel() * 4 / res["triplet_fwd_bwd_us"] / 1e3
    res["batch_hard_fwd_bwd_us"] = graph_us(lambda: B.check(lib.sbir_batch_hard_triplet_loss(
        a.data_ptr(), p.data_ptr(), n.data_ptr(), 256, 2048, 0.2, B.SBIR_EUCLIDEAN, None, None, loss.data_ptr(), hard.data_ptr(),
        ga.data_ptr(), gp.data_ptr(), gn.data_ptr(), ws.data_ptr(), ws.numel(), st()), "batch_hard"))
    ta, tp, tn = (t.clone().requires_grad_(True) for t in (a, p, n))

    def torch_triplet():
        ta.grad = tp.grad = tn.grad = None
        torch.nn.functional.triplet_margin_loss(ta, tp, tn, margin=0.2).backward()
    res["torch_triplet_fwd_bwd_us"] = graph_us(torch_triplet)
    eye = torch.zeros(256, 512, dtype=torch.bool, device=dev)
    eye[torch.arange(256, device=dev), torch.arange(256, device=dev)] = True

    def torch_batch_hard():  # the same definition with library ops: broadcast distance + masks + autograd
        ta.grad = tp.grad = tn.grad = None
        x = torch.cat([tp, tn])
        dm = (ta[:, None, :] - x[None, :, :] + 1e-6).norm(dim=2)
        hp = dm.masked_fill(~eye, float("-inf")).max(dim=1).values
        hn = dm.masked_fill(eye, float("inf")).min(dim=1).values
        torch.clamp_min(0.2 + hp - hn, 0).mean().backward()
    res["torch_batch_hard_fwd_bwd_us"] = graph_us(torch_batch_hard)
    res["clocks"] = sampler.stop(t_begin, time.time())
    # parity in the same run (oracle = the reference's loss modules on CPU)
    A, P, N = (t.clone().requires_grad_(True) for t in (a, p, n))
    l_t = ops.triplet_margin_loss(A, P, N, 0.2, "euclidean")
    l_t.backward()
    ref_t = O.triplet_margin_loss(a_h, p_h, n_h, 0.2, "euclidean")
    A2, P2, N2 = (t.clone().requires_grad_(True) for t in (a, p, n))
    l_b, hidx = ops.batch_hard_triplet_loss(A2, P2, N2, 0.2, "euclidean", return_indices=True)
    l_b.backward()
    Ar, Pr, Nr = (t.clone().requires_grad_(True) for t in (a_h, p_h, n_h))
    ref_b, hpi, hni = O.batch_hard_triplet_loss(Ar, Pr, Nr, 0.2, "euclidean")
    ref_b.backward()
    gerr = max(((x.grad.cpu() - y.grad).abs().max() / y.grad.abs().max()).item() for x, y in ((A2, Ar), (P2, Pr), (N2, Nr)))
    res["parity"] = {"triplet_loss_rel_err": abs(l_t.item() - ref_t.item()) / max(abs(ref_t.item()), 1e-12),
                     "batch_hard_loss_rel_err": abs(l_b.item() - ref_b.item()) / max(abs(ref_b.item()), 1e-12),
                     "batch_hard_index_mismatches": int((hidx[:, 0].cpu() != hpi).sum() + (hidx[:, 1].cpu() != hni).sum()),
                     "batch_hard_grad_max_err_rel_to_max": gerr}
    return res


# ------------------------------------------------------------------------ main ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (extra_workloads)")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled full-size oracle check")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep equal gallery shards (default: a few untimed calibration rounds cut the gallery by each GPU's "
                    "sustained distance-kernel speed, sharded.weighted_shard_bounds)")
    ap.add_argument("--balance-threshold", type=float, default=1.015, help="N > 1: stop re-cutting once max/min of the per-rank kernel times is below this")
    ap.add_argument("--centroids", type=int, default=0, help="class centroids of the generator (0: max(125, N/80), see make_shard)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from art_sbir_b200 import _binding as B, _build, ops, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device: the sbir_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _build.build()
    lib = B.load()
    if lib.sbir_device_supported() != 1:
        sys.exit("bench.py needs an sm_100 (B200) device")

    num_q, num_g, dim, dtype_name, k, desc = wl
    dtype = getattr(torch, dtype_name)
    r0, r1 = sharded.shard_bounds(num_g, world, rank)
    Q, Gs, pos = make_shard(num_q, num_g, dim, dtype, r0, r1, dev, centroids=args.centroids)
    torch.cuda.synchronize()

    def step():
        if world == 1:
            return ops.pairwise_topk(Q, Gs, k, "euclidean", pos_index=pos, return_uncertified=True)
        v, i, r = sharded.sharded_pairwise_topk(Q, Gs, k, "euclidean", pos_index=pos, shard_offset=r0, num_gallery_total=num_g)
        return v, i, r, None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- N > 1: speed-weighted shards.  The exchange step waits for the slowest shard, and under the power cap the GPUs of
    # one chassis do not sustain the same clock (8 GPUs, equal shards: per-rank kernel times 92-103 ms on one box, 93-99 ms
    # on another).  Untimed, before the measurement, every rank times its distance kernel and the gallery is re-cut towards
    # equal kernel times (sharded.weighted_shard_bounds) — damped and re-measured up to three times, because a GPU that
    # finishes early idles until the exchange and overstates what it sustains at full duty.  The data generator is
    # row-deterministic, so the global problem — and the result — is unchanged; `timing.shards` records every round.
    # Measured A/B on 8 GPUs (profiles/r02_bench_cfg4_n8_{equal,weighted}_shards*.json): 105.0 -> 101.3 ms on a box with
    # a persistent 12 % spread, 102.1 -> 102.1 ms on one whose 6 % spread was jitter.
    balance = None
    if world > 1 and not args.no_balance:
        balance = {"rounds": [], "damping": 0.6}
        for _ in range(3):
            _, cal_k1, _, _, _ = time_retrieval(step, 4, 3, barrier, lib, dev, None)
            mine = torch.tensor([float(r1 - r0), cal_k1], device=dev, dtype=torch.float64)
            every = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            rows = [t[0].item() for t in every]
            ms = [t[1].item() for t in every]
            balance["rounds"].append({"gallery_rows": [int(x) for x in rows], "k1_ms": [round(x, 3) for x in ms]})
            if max(ms) / min(ms) < args.balance_threshold:
                break
            # A GPU that finishes early idles (and cools) until the exchange step, so its measured speed overstates what it
            # sustains at full duty: move only part of the way towards equal kernel times, and re-measure.
            weights = sharded.rebalanced_weights(rows, ms, damping=0.6)
            n0, n1 = sharded.weighted_shard_bounds(num_g, weights, rank, align=256)
            del Q, Gs, pos
            torch.cuda.empty_cache()
            r0, r1 = n0, n1
            Q, Gs, pos = make_shard(num_q, num_g, dim, dtype, r0, r1, dev, centroids=args.centroids)
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()
        time.sleep(0.3)
    # inputs smaller than L2 are evicted between timed steps by writing a 512 MiB buffer (untimed)
    in_bytes = (Q.numel() + Gs.numel()) * Q.element_size()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if in_bytes < 400e6 else None
    ms_step, k1_ms_step, launches, out, clocks = time_retrieval(step, args.steps, args.warmup, barrier, lib, dev, flush, sampler)
    t = torch.tensor([ms_step, k1_ms_step, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_per_step, k1_ms_per_launch, total_launches = tmax[0].item(), tmax[1].item(), int(tsum[2].item())
        mine = torch.tensor([ms_step, k1_ms_step, float(r1 - r0)], device=dev, dtype=torch.float64)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = {"step_ms": [round(t[0].item(), 3) for t in every], "k1_ms": [round(t[1].item(), 3) for t in every],
                    "gallery_rows": [int(t[2].item()) for t in every]}
    else:
        ms_per_step, k1_ms_per_launch, total_launches = t[0].item(), t[1].item(), int(t[2].item())
        per_rank = None
    pairs = num_q * num_g
    value = pairs / (ms_per_step * 1e-3)

    vals, idx, rank0, unc = out
    recall = {f"recall@{kk}": float((rank0 < kk).float().mean().item()) for kk in (1, 5, 10)}
    uncert = int(unc.item()) if unc is not None else None
    parity = None if args.no_parity else sampled_parity(Q, Gs, pos, r0, k, vals, idx, rank0, world, dist)

    # ---- the other BASELINE configs, briefly, on rank 0's GPU (N = 1 only: they are single-GPU configs) ----
    extras = []
    peaks = measured_peaks()
    tf32_peak = None
    if world == 1 and not args.no_extra:
        tf32_peak = measure_tf32_peak(dev)
        if args.workload == "cfg4":
            extras.append(extra_retrieval_workload("cfg4k100", lib, dev, local_rank, peaks, tf32_peak, args.centroids, reuse=(Q, Gs, pos)))
        for name in ("cfg1", "cfg3", "cfg3k10"):
            if name != args.workload:
                extras.append(extra_retrieval_workload(name, lib, dev, local_rank, peaks, tf32_peak, args.centroids))
        extras.append(cfg2_workload(lib, dev, local_rank))

    # ---- e2e: pinned host buffers → C ABI / sharded API → host results ----
    e2e = None
    if not args.no_e2e:
        qh = Q.cpu().pin_memory()
        gh = Gs.cpu().pin_memory()
        ph = pos.cpu().pin_memory()
        od = torch.empty(num_q, k).pin_memory()
        oi = torch.empty(num_q, k, dtype=torch.int64).pin_memory()
        orank = torch.empty(num_q, dtype=torch.int64).pin_memory()
        h2d = qh.numel() * qh.element_size() + gh.numel() * gh.element_size() + ph.numel() * 8
        d2h = od.numel() * 4 + oi.numel() * 8 + orank.numel() * 8
        es = max(1, min(args.steps, 3))

        def e2e_step():
            if world == 1:
                c_unc = ctypes.c_int32(0)
                B.check(lib.sbir_retrieve_host(qh.data_ptr(), num_q, gh.data_ptr(), num_g, dim,
                                               B.SBIR_BF16 if dtype == torch.bfloat16 else B.SBIR_F32, B.SBIR_EUCLIDEAN, k,
                                               ph.data_ptr(), od.data_ptr(), oi.data_ptr(), orank.data_ptr(),
                                               ctypes.byref(c_unc)), "sbir_retrieve_host")
            else:
                # each rank streams its shard from pinned host memory (sbir_retrieve_host_shard)
                v, i, r = sharded.sharded_retrieve_host(qh, gh, k, "euclidean", pos_index=ph, shard_offset=r0,
                                                        num_gallery_total=num_g, device=dev)
                od.copy_(v, non_blocking=True)
                oi.copy_(i, non_blocking=True)
                orank.copy_(r, non_blocking=True)
                torch.cuda.synchronize()

        del Gs
        torch.cuda.empty_cache()
        e2e_step()  # warm-up (staging allocation)
        barrier()
        t0 = time.perf_counter()
        for _ in range(es):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / es], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(oi, idx.cpu()) and torch.equal(orank, rank0.cpu()))
        lib.sbir_release_host_staging()
        # context for the e2e number (untimed): how long the SAME host->device bytes take as plain pinned copies with all
        # ranks copying at once — the floor the host path cannot go below on this box, next to the device-only step time
        sink_q, sink_g = torch.empty_like(qh, device=dev), torch.empty_like(gh, device=dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            sink_q.copy_(qh, non_blocking=True)
            sink_g.copy_(gh, non_blocking=True)
        barrier()
        ct = torch.tensor([(time.perf_counter() - t0) / 2], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        del sink_q, sink_g
        e2e = {"value": pairs / dt.item(), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt.item() * 1e3, "steps": es, "matches_device_path": same,
               "h2d_copy_only_ms": ct.item() * 1e3,
               "h2d_copy_only_GBps_per_rank": (qh.numel() * qh.element_size() + gh.numel() * gh.element_size()) / ct.item() / 1e9,
               "note": "the upload is streamed into the pass: e2e ~ max(device step, plain pinned H2D copy of the same bytes with all ranks copying at once) + prologue"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic_file = ROOT / "profiles" / f"k1_traffic_{args.workload}.json"
    # ncu capture of the single-GPU launch of this workload (profiles/); per-rank launches at N > 1
    # cover a shard and were not captured separately
    traffic = json.loads(traffic_file.read_text()).get("dram_bytes_per_launch") if (traffic_file.is_file() and world == 1) else None
    tile_bf16 = tiles_are_bf16(lib, num_q, r1 - r0, dim, k, dtype == torch.bfloat16)
    if not tile_bf16 and tf32_peak is None:
        tf32_peak = measure_tf32_peak(dev)
    # N > 1: the roofline line describes the rank whose distance kernel ran longest (its own shard rows)
    slow_rows = r1 - r0 if per_rank is None else per_rank["gallery_rows"][max(range(world), key=lambda r: per_rank["k1_ms"][r])]
    roofline = k1_roofline(dim, num_q, slow_rows, tile_bf16, k1_ms_per_launch, ms_per_step, peaks, tf32_peak, traffic)
    if per_rank is not None:
        per_rank["k1_tflops"] = [round(2.0 * dim * num_q * n / (t * 1e-3) / 1e12, 1) if t > 0 else None
                                 for n, t in zip(per_rank["gallery_rows"], per_rank["k1_ms"])]

    cpu = None
    if not args.no_cpu:
        ng_s = min(num_g, 500_000 if dim <= 512 else 75_000)
        v, done, dt, threads = cpu_reference_sample(ng_s, dim, 4096, seconds_hint=12.0, dtype_name=dtype_name)
        cpu = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": f"{done} queries x {ng_s} gallery rows of the same clustered workload ({dim}-d, {dtype_name}-rounded, fp32 math), "
                         f"reference per-query loop (PairwiseDistance + topk(N) + nonzero, inference.py:44,49,52), {dt:.1f} s"}

    line = {"metric": "query x gallery pairs/sec (distance + top-%d + rank)" % k, "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if dtype == torch.bfloat16 else "f32 (selected on tensor-core tiles, exact fp32/fp64 re-score)",
            "data": "synthetic",
            "config": workload_config(args.workload, world, args.centroids),
            "timing": {"l2": "inputs larger than L2" if flush is None else "L2 flushed (512 MiB write) between timed steps",
                       "method": "CUDA events on the launching stream around each of K steps after W warm-up steps, barrier + synchronize on both sides, max over ranks",
                       "per_rank": per_rank,
                       "shards": None if world == 1 else ("equal" if balance is None else
                                                          {"rule": "gallery re-cut towards equal distance-kernel times in damped, re-measured rounds (untimed, before the measurement; stops below 1.5 % spread; same global problem, same result)",
                                                           **balance})},
            "results": {**recall, "uncertified_queries": uncert},
            "clocks": clocks, "e2e": e2e, "gpu_launches": total_launches, "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity, "tf32_peak_measured_tflops": tf32_peak, "extra_workloads": extras}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
