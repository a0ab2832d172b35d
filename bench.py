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
def make_shard(num_q, num_g, dim, dtype, row0, row1, device, seed=1234):
    """Rows [row0,row1) of the seeded clustered gallery + ALL queries (SURVEY.md §8d generator,
    evaluated on device in chunks with per-chunk seeds so every rank sees the same global data)."""
    import torch
    C = max(125, num_g // 80)
    beta = 0.06 if dim >= 2048 else 0.12
    gen = torch.Generator(device=device).manual_seed(seed)
    cent = torch.randn(C, dim, device=device, generator=gen)
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


# ----------------------------------------------------------------- CPU reference ----
def cpu_reference_sample(num_g_sample, dim, nq_sample, seconds_hint=20.0):
    """The reference's per-query path (inference.py:44,49: PairwiseDistance broadcast +
    topk(len(G))) via the oracle port, fp32, all host threads, on a bounded sample."""
    import torch
    from oracle import sbir_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    G = torch.randn(num_g_sample, dim, generator=g)
    Q = torch.randn(nq_sample, dim, generator=g)
    O.ranking_position(Q[:1], G, 0, "euclidean")  # warm-up (allocator, threads)
    t0 = time.perf_counter()
    done = 0
    for i in range(nq_sample):
        # what the reference does for EVERY query (inference.py:113 → :44,:49,:52); its full sort
        # subsumes the top-10 (get_topk_images, :62-65, only runs for 10 sampled queries)
        O.ranking_position(Q[i:i + 1], G, i % num_g_sample, "euclidean")
        done += 1
        if time.perf_counter() - t0 > seconds_hint:
            break
    dt = time.perf_counter() - t0
    return done * num_g_sample / dt, done, dt, torch.get_num_threads()


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the
    reference is pure Python/torch and /root/reference does not exist on the GPU box)."""
    num_q, num_g, dim, dtype, k, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ng_s = min(num_g, 500_000 if dim <= 512 else 75_000)
    times, pairs = [], 0
    per_step_q = 4
    for s in range(args.warmup + args.steps):
        v, done, dt, threads = cpu_reference_sample(ng_s, dim, per_step_q, seconds_hint=30.0)
        if s >= args.warmup:
            times.append(dt)
            pairs += done * ng_s
    value = pairs / sum(times)
    sample = f"{per_step_q} queries x {ng_s} gallery rows per step ({dim}-d fp32), PairwiseDistance + topk(N) per query (inference.py:44,49,52)"
    line = {"impl": "reference", "metric": "query x gallery pairs/sec (distance + top-%d + rank)" % k, "value": value,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "sampled": sample},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    Q, Gs, pos = make_shard(num_q, num_g, dim, dtype, r0, r1, dev)
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

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    for _ in range(args.warmup):
        out = step()
    barrier()
    lib.sbir_profile_enable(1)
    k1_ms, k1_n, launches = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    lib.sbir_profile_collect(ctypes.byref(k1_ms), ctypes.byref(k1_n), ctypes.byref(launches))  # reset counters
    # inputs smaller than L2 are evicted between timed steps by writing a 512 MiB buffer (untimed)
    in_bytes = (Q.numel() + Gs.numel()) * Q.element_size()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if in_bytes < 400e6 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_begin = time.time()
    for a, b in evs:
        if flush is not None:
            flush.fill_(1)
        a.record()
        out = step()
        b.record()
    barrier()
    clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    lib.sbir_profile_collect(ctypes.byref(k1_ms), ctypes.byref(k1_n), ctypes.byref(launches))
    lib.sbir_profile_enable(0)
    # K1 device time per step: fp32 workloads enqueue a second, device-gated K1 launch (the 3xTF32
    # escalation pass) that returns at once when the first pass certified everything, so the sum of
    # the K1 launches of a step is the time of the one that did the work
    t = torch.tensor([ms_total, k1_ms.value / max(1, args.steps), float(launches.value)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, k1_ms_per_launch, total_launches = tmax[0].item(), tmax[1].item(), int(tsum[2].item())
    else:
        ms_total, k1_ms_per_launch, total_launches = t[0].item(), t[1].item(), int(t[2].item())
    ms_per_step = ms_total / args.steps
    pairs = num_q * num_g
    value = pairs / (ms_per_step * 1e-3)

    vals, idx, rank0, unc = out
    recall = {f"recall@{kk}": float((rank0 < kk).float().mean().item()) for kk in (1, 5, 10)}
    uncert = int(unc.item()) if unc is not None else None

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
        e2e = {"value": pairs / dt.item(), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt.item() * 1e3, "steps": es, "matches_device_path": same}
        lib.sbir_release_host_staging()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    flops_per_launch = 2.0 * dim * num_q * (r1 - r0)
    achieved = flops_per_launch / (k1_ms_per_launch * 1e-3) / 1e12 if k1_ms_per_launch > 0 else None
    if dtype == torch.bfloat16:
        peak = peaks["bf16_sustained"] if k1_ms_per_launch > 100 else peaks["bf16"]
        peak_note = ("bf16 dense, sustained, " if k1_ms_per_launch > 100 else "bf16 dense, burst, ") + peaks["source"]
    else:
        peak = 746.8  # cuBLAS TF32 8192^3 measured on this pool (profiles/r01_probe2_shared_thr_pool.log), MEASURED_PEAKS has no tf32 entry
        peak_note = "tf32 dense, cuBLAS 8192^3 measured in round 1 (kind::tf32 runs at half the bf16 rate)"
    traffic_file = ROOT / "profiles" / f"k1_traffic_{args.workload}.json"
    # ncu capture of the single-GPU launch of this workload (profiles/); per-rank launches at N > 1
    # cover a shard and were not captured separately
    traffic = json.loads(traffic_file.read_text()).get("dram_bytes_per_launch") if (traffic_file.is_file() and world == 1) else None
    burst = peaks["bf16"] if dtype == torch.bfloat16 else peak
    roofline = {"bound": "tensor", "kernel": "dist_topk_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_note": peak_note,
                "frac_of_burst_peak": (achieved / burst) if achieved else None,
                "k1_ms_per_launch": k1_ms_per_launch, "k1_share_of_step": k1_ms_per_launch / ms_per_step}

    cpu = None
    if not args.no_cpu:
        ng_s = min(num_g, 500_000 if dim <= 512 else 75_000)
        v, done, dt, threads = cpu_reference_sample(ng_s, dim, 4096, seconds_hint=12.0)
        cpu = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": f"{done} queries x {ng_s} gallery rows, {dim}-d fp32, reference per-query loop "
                         f"(PairwiseDistance + topk(N), inference.py:44,49,52), {dt:.1f} s"}

    line = {"metric": "query x gallery pairs/sec (distance + top-%d + rank)" % k, "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if dtype == torch.bfloat16 else "tf32 (fp32 in, fp32 accumulate, exact fp32/fp64 re-score)",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "num_q": num_q, "num_g": num_g, "dim": dim, "k": k,
                       "sharding": f"gallery rows over {world} GPU(s)", "l2": "inputs larger than L2" if flush is None else "L2 flushed (512 MiB write) between timed steps",
                       **recall, "uncertified_queries": uncert},
            "clocks": clocks, "e2e": e2e, "gpu_launches": total_launches, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
