#!/usr/bin/env python
"""bench.py — VMIS-kNN predict_next throughput on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic evolving sessions.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm
  python bench.py --impl reference ...                           # CPU arm: restatement of the Rust path

Workload (config.workload "synthetic-60M-1.76M", BASELINE.json configs[2], the configuration the
metric is quoted on): 11.556 M synthetic training sessions = 60.0 M interactions over 1.76 M items
(seed 42), index built with m=1502, idf_weighting=2, max_len=34; queries are the last <= 4 items of
random prefixes of held-out sessions (seed 43 + step), k=288, m=1502, how_many=21, business logic off.
Every step uses a fresh query batch and the index (~0.6 GB) is far larger than L2.

Multi-GPU (--gpus N under torchrun): the path shards by query — every rank holds a replica of the
index and its own query stream; no data-path collective; scaling "weak".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_items, n_sessions)  — sessions chosen so that interactions hit the config's figure
    "synthetic-1M-50k": (50_000, 193_000),
    "synthetic-60M-1.76M": (1_760_000, 11_556_000),
    # configs 4-5: generated and indexed on the device (vmis_index_synth); no CPU arm at these sizes
    "synthetic-582M-6.5M": (6_500_000, 112_100_000),
    "synthetic-2.3B-6.5M": (6_500_000, 443_000_000),
}
DEVICE_BUILT = {"synthetic-582M-6.5M", "synthetic-2.3B-6.5M"}
K, M, HOW_MANY, MAX_ITEMS, IDF_W, MAX_LEN = 288, 1502, 21, 4, 2.0, 34
METRIC = "predict_next queries/sec @ k=288,m=1502"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi SM clock / throttle-reason sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower() == "active":
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_queries(sb, n_items, batch, n_batches, seed0):
    out = []
    for b in range(n_batches):
        out.append(sb.synth_queries(seed0 + b, n_items, batch, MAX_ITEMS))
    return out


def algorithmic_bytes(stats, q_off):
    """sum_j min(df_j,m)*8 + sum_{s in top-k}(8 + len_s*16) + L*8 + n_out*16 (reference widths)."""
    st = stats.astype(np.int64)
    L = np.diff(q_off.astype(np.int64))
    return int((st[:, 0] * 8 + st[:, 1] * 8 + st[:, 2] * 16 + st[:, 3] * 16).sum() + (L * 8).sum())


def cpu_reference(oracle_index, queries, threads, budget_s, batch_hint=4096):
    """Time the faithful CPU restatement on a bounded sample: grow the sample until ~budget_s."""
    q_items, q_off = queries
    n = min(batch_hint, len(q_off) - 1)
    r = oracle_index.predict_batch(q_items[:q_off[n]], q_off[:n + 1], K, M, HOW_MANY, False, mode=0,
                                   threads=threads, want_outputs=False)
    rate = n / max(r[3], 1e-9)
    n2 = int(min(len(q_off) - 1, max(n, rate * budget_s)))
    r = oracle_index.predict_batch(q_items[:q_off[n2]], q_off[:n2 + 1], K, M, HOW_MANY, False, mode=0,
                                   threads=threads, want_outputs=False)
    return n2 / r[3], n2, r[3]


def cpu_single_thread_latency(oracle_index, queries, n=2000):
    """evaluator.rs:57-89: per-call wall time of `predict` on ONE thread, percentiles in microseconds (the reference's
    README.md:16 quotes p90 < 1.7 ms for this index size on its own hardware)."""
    q_items, q_off = queries
    n = min(n, len(q_off) - 1)
    r = oracle_index.predict_batch(q_items[:q_off[n]], q_off[:n + 1], K, M, HOW_MANY, False, mode=0, threads=1,
                                   want_latency=True, want_outputs=False)
    lat = np.sort(r[4][n // 10:])                      # the first tenth warms the caches
    return {f"p{str(p).replace('.', '_')}": round(float(np.percentile(lat, p)), 1) for p in (25, 50, 75, 90, 95, 99.5)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="synthetic-60M-1.76M", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1 << 20, help="evolving sessions per step (per GPU)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-build", action="store_true", help="generate + index the workload on the device")
    ap.add_argument("--sharded", action="store_true",
                    help="multi-GPU only: item-shard the postings over the ranks (config 5 layout, remote lists read "
                         "over NVLink inside the kernel) instead of replicating the index")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: --warmup < 3 breaks the timing rules; use >= 3 for a reported number")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_items, n_sessions = WORKLOADS[args.workload]
    cfg = {"workload": args.workload, "interactions": None, "items": n_items, "sessions": n_sessions,
           "k": K, "m": M, "how_many": HOW_MANY, "max_items_in_session": MAX_ITEMS, "idf_weighting": IDF_W,
           "batch_per_gpu": args.batch, "parallelism": f"query-sharded replicas x{world}",
           "cache": "fresh query batch every step; index (>500 MB) larger than L2, no explicit flush"}

    import serenade_b200 as sb

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import vmis_oracle as vo
        t0 = time.time()
        items, off, ts = sb.synth_sessions(42, n_items, n_sessions)
        cfg["interactions"] = int(len(items))
        oix = vo.OracleIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W)
        log(f"[reference] data + CPU index in {time.time() - t0:.1f}s")
        threads = os.cpu_count() or 1
        steps = args.steps
        # each step: a bounded sample sized so that the whole run lasts about cpu-seconds * 2
        q = sb.synth_queries(43, n_items, 1 << 16, MAX_ITEMS)
        probe = oix.predict_batch(q[0][:q[1][2048]], q[1][:2049], K, M, HOW_MANY, False, mode=0, threads=threads,
                                  want_outputs=False)
        rate = 2048 / probe[3]
        per_step = int(max(512, min(1 << 16, rate * (2 * args.cpu_seconds) / max(1, steps + args.warmup))))
        tot_q, tot_t = 0, 0.0
        for s in range(args.warmup + steps):
            qs = sb.synth_queries(43 + s, n_items, per_step, MAX_ITEMS)
            r = oix.predict_batch(qs[0], qs[1], K, M, HOW_MANY, False, mode=0, threads=threads, want_outputs=False)
            if s >= args.warmup:
                tot_q += per_step
                tot_t += r[3]
        v = tot_q / tot_t
        sample = f"{per_step} queries/step x {steps} steps of the same generator (seed 43+step)"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/i32+f64",
                          "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "queries/s", "cores": threads, "kind": "port",
                                           "sample": sample,
                                           "single_thread_latency_us": cpu_single_thread_latency(oix, q)},
                          "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = sb.load_library()

    t0 = time.time()
    device_built = args.workload in DEVICE_BUILT or args.device_build
    shard_args = (rank, world) if (args.sharded and world > 1) else (0, 1)
    if device_built:
        items = off = ts = None
        t1 = time.time()
        gix = sb.VMISIndex.synth(42, n_items, n_sessions, M, MAX_LEN, IDF_W, local_rank, *shard_args)
        cfg["interactions"] = int(gix.stats()["n_pairs_kept"])
        cfg["index_build"] = "generated and indexed on the device (vmis_index_synth)"
    else:
        items, off, ts = sb.synth_sessions(42, n_items, n_sessions)
        cfg["interactions"] = int(len(items))
        t1 = time.time()
        if shard_args[1] > 1:
            gix = sb.VMISIndex.from_sessions_sharded(items, off, ts, M, MAX_LEN, IDF_W, local_rank, *shard_args)
        else:
            gix = sb.VMISIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W, device=local_rank)
    if shard_args[1] > 1:
        gix.connect_shards(rank, world)
        cfg["parallelism"] = f"item-sharded postings x{world} (peer HBM over NVLink, CUDA IPC), queries sharded"
    st = gix.stats()
    log(f"[rank {rank}] synth {t1 - t0:.1f}s, index build+upload {time.time() - t1:.1f}s, "
        f"{st['device_bytes'] / 1e6:.0f} MB in HBM, {st['n_items']} items, {st['n_postings']} postings")
    cfg["index_hbm_bytes"] = int(st["device_bytes"])

    n_batches = args.warmup + args.steps
    B, n = args.batch, HOW_MANY
    batches = make_queries(sb, n_items, B, n_batches, 43 + 1000 * rank)
    hist = np.bincount(np.diff(batches[0][1].astype(np.int64)), minlength=MAX_ITEMS + 1)
    cfg["session_len_hist"] = {str(i): int(c) for i, c in enumerate(hist) if c}

    # device-resident inputs for `value`
    d_batches = [(torch.from_numpy(qi.view(np.int64)).to(dev), torch.from_numpy(qo.view(np.int32)).to(dev))
                 for qi, qo in batches]
    d_ids = torch.zeros((B, n), dtype=torch.int64, device=dev)
    d_sc = torch.zeros((B, n), dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    d_st = torch.zeros((B, 4), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)          # kernels and timing events share this stream
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)

    def launch(b, stats=False):
        di, do = d_batches[b]
        rc = lib.vmis_predict_batch_device(gix.handle, di.data_ptr(), do.data_ptr(), B, K, M, n, 0, d_ids.data_ptr(),
                                           d_sc.data_ptr(), d_cnt.data_ptr(), d_st.data_ptr() if stats else None, sptr)
        if rc != 0:
            raise RuntimeError(lib.vmis_last_error().decode())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # untimed stats pass → exact algorithmic bytes of each timed batch (also part of warm-up)
    alg_bytes = []
    for b in range(n_batches):
        launch(b, stats=True)
        torch.cuda.synchronize()
        alg_bytes.append(algorithmic_bytes(d_st.cpu().numpy(), batches[b][1]))
    for b in range(args.warmup):
        launch(b)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for s in range(args.steps):
        launch(args.warmup + s)
        ev[s + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    from serenade_b200.shard import max_over_ranks
    total_ms_max = max_over_ranks(total_ms, device=dev)
    value = world * args.steps * B / (total_ms_max * 1e-3)

    # roofline of the (single) kernel: algorithmic bytes of the timed batches / their kernel time
    timed_bytes = sum(alg_bytes[args.warmup:])
    achieved = timed_bytes / (sum(step_ms) * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    else:
        peak, peak_src = 6650.0, "fallback"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("workload") == args.workload and tj.get("queries_per_launch") == args.batch:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "vmis_predict_kernel",
                "algorithmic_bytes_per_query": timed_bytes / (args.steps * B)}

    # e2e: same metric through the host-buffer C ABI call, pinned host memory, copies inside the timed region
    h_q = [(torch.from_numpy(qi.view(np.int64)).pin_memory(), torch.from_numpy(qo.view(np.int32)).pin_memory())
           for qi, qo in batches]
    h_ids = torch.zeros((B, n), dtype=torch.int64).pin_memory()
    h_sc = torch.zeros((B, n), dtype=torch.float64).pin_memory()
    h_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()
    u64p, u32p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_double)

    def e2e_call(b):
        qi, qo = h_q[b]
        rc = lib.vmis_predict_batch(gix.handle, C.cast(qi.data_ptr(), u64p), C.cast(qo.data_ptr(), u32p), B, K, M, n, 0,
                                    C.cast(h_ids.data_ptr(), u64p), C.cast(h_sc.data_ptr(), f64p),
                                    C.cast(h_cnt.data_ptr(), u32p), None)
        if rc != 0:
            raise RuntimeError(lib.vmis_last_error().decode())

    for b in range(args.warmup):
        e2e_call(b)
    barrier()
    t_e0 = time.perf_counter()
    for s in range(args.steps):
        e2e_call(args.warmup + s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t_e0
    e2e_value = world * args.steps * B / max_over_ranks(e2e_s, device=dev)
    h2d = int(np.mean([qi.numel() * 8 + qo.numel() * 4 for qi, qo in h_q[args.warmup:]]))
    d2h = B * n * 16 + B * 4

    # single-call latency of the reference's exact call shape (one evolving session, mod.rs:118-125), outside
    # every timed region; the reference's evaluator prints the same percentiles (evaluator.rs:82-89)
    lat = []
    qi0, qo0 = batches[0]
    ids1 = np.zeros(n, dtype=np.uint64); sc1 = np.zeros(n, dtype=np.float64)
    for q in range(300):
        ev = np.ascontiguousarray(qi0[qo0[q]:qo0[q + 1]])
        t_a = time.perf_counter()
        lib.vmis_predict(gix.handle, ev.ctypes.data_as(u64p), len(ev), K, M, n, 0, ids1.ctypes.data_as(u64p),
                         sc1.ctypes.data_as(f64p))
        lat.append((time.perf_counter() - t_a) * 1e6)
    lat = np.sort(np.array(lat[50:]))
    latency = {"single_call_us": {p: round(float(np.percentile(lat, float(p[1:]))), 1) for p in ("p50", "p90", "p99")},
               "note": "vmis_predict, one evolving session per call, host buffers"}

    out = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u32/i32+f64", "data": "synthetic", "config": cfg,
           "clocks": clocks, "gpu_launches": args.steps * 1,
           "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "launches_per_step": (B + (1 << 17) - 1) >> 17,
                   "note": "vmis_predict_batch pipelines the batch in chunks of 2^17 sessions over 3 streams"},
           "roofline": roofline, "step_ms": [round(x, 3) for x in step_ms], "latency": latency}

    if rank == 0 and world == 1 and not args.no_cpu_baseline and items is not None:
        from oracle import vmis_oracle as vo
        t2 = time.time()
        oix = vo.OracleIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W)
        threads = os.cpu_count() or 1
        qps, n_sample, secs = cpu_reference(oix, batches[args.warmup], threads, args.cpu_seconds)
        log(f"[cpu] oracle index {time.time() - t2 - secs:.1f}s; {n_sample} queries in {secs:.1f}s on {threads} threads")
        out["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                               "sample": f"first {n_sample} queries of the first timed batch, faithful mode, "
                                         f"{threads} threads sharing one index",
                               "single_thread_latency_us": cpu_single_thread_latency(oix, batches[args.warmup])}
        # spot parity inside the bench: the same sample through the canonical oracle vs the device result
        m_chk = min(2048, n_sample)
        qi, qo = batches[args.warmup]
        oids, osc, ocnt, _, _ = oix.predict_batch(qi[:qo[m_chk]], qo[:m_chk + 1], K, M, n, False, mode=1, threads=threads)
        launch(args.warmup)
        torch.cuda.synchronize()
        ok = (np.array_equal(d_ids[:m_chk].cpu().numpy().view(np.uint64), oids) and
              np.array_equal(d_sc[:m_chk].cpu().numpy(), osc) and
              np.array_equal(d_cnt[:m_chk].cpu().numpy().view(np.uint32), ocnt))
        out["parity_check"] = {"queries": m_chk, "bit_exact_vs_canonical_oracle": bool(ok)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
