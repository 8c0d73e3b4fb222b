#!/usr/bin/env python
"""bench.py — VMIS-kNN predict_next throughput on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic evolving sessions.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm
  python bench.py --impl reference ...                           # CPU arm: restatement of the Rust path

Headline workload (config.workload "synthetic-60M-1.76M", BASELINE.json configs[2], the configuration the
metric is quoted on): 11.556 M synthetic training sessions = 60.0 M interactions over 1.76 M items
(seed 42), index built with m=1502, idf_weighting=2, max_len=34; queries are the last <= 4 items of
random prefixes of held-out sessions (seed 43 + step), k=288, m=1502, how_many=21, business logic off.
Every step uses a fresh query batch and the index (~0.7 GB) is far larger than L2.

Besides the headline line the B200 arm reports, in the same JSON object:
  * N = 1: `sections.config2` (BASELINE configs[1]: 1 M interactions / 50 k items, launches of 1024 sessions),
    `sections.config4` (configs[3]: 582 M interactions / 6.5 M items, generated and indexed on the device) and
    `latency` (single call, and the micro-batcher under open-loop load);
  * N > 1 (torchrun): `sections.item_sharded` — the index of the headline workload with its postings sharded by
    item over the N GPUs (remote lists read over NVLink inside the kernel), checked bit for bit against the
    replica index, and at N = 8 `sections.config5` (configs[4]: 2.3 B interactions / 6.5 M items, item-sharded).
Multi-GPU headline (`value`): the path shards by query — every rank holds a replica of the index and its own query
stream; no data-path collective; scaling "weak".
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
if os.environ.get("VMIS_BENCH_KM"):      # tuning experiments only (never a reported number): k,m override
    K, M = (int(x) for x in os.environ["VMIS_BENCH_KM"].split(","))
METRIC = "predict_next queries/sec @ k=288,m=1502"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def base_config(workload, interactions, batch, world, len_hist):
    """The `config` object — identical keys and values in the B200 arm and the reference arm."""
    n_items, n_sessions = WORKLOADS[workload]
    return {"workload": workload, "interactions": interactions, "items": n_items, "sessions": n_sessions,
            "k": K, "m": M, "how_many": HOW_MANY, "max_items_in_session": MAX_ITEMS, "idf_weighting": IDF_W,
            "batch_per_gpu": batch, "parallelism": f"query-sharded replicas x{world}",
            "cache": "fresh query batch every step; index (>500 MB) larger than L2, no explicit flush",
            "session_len_hist": len_hist}


def length_hist(q_off):
    hist = np.bincount(np.diff(q_off.astype(np.int64)), minlength=MAX_ITEMS + 1)
    return {str(i): int(c) for i, c in enumerate(hist) if c}


class ClockSampler:
    """nvidia-smi SM clock / throttle-reason sampling during a timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def wait_for_samples(self, n, timeout=3.0):
        t0 = time.time()
        while self.proc and len(self.rows) < n and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
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


def faithful_agreement(oids, ocnt, fids, fcnt):
    """How far the canonical result (= the kernel's) sits from the faithful restatement of the Rust code on the same
    queries: at k < m the reference's k-boundary heuristic depends on hash-map iteration order
    (vmis_index.rs:394-412), so the two may pick different neighbours among equal-similarity sessions."""
    n = len(ocnt)
    same_set = same_list = 0
    overlap = 0.0
    for q in range(n):
        a, b = oids[q, :ocnt[q]], fids[q, :fcnt[q]]
        sa, sb_ = set(a.tolist()), set(b.tolist())
        same_set += sa == sb_
        same_list += len(a) == len(b) and bool(np.array_equal(a, b))
        overlap += len(sa & sb_) / max(1, max(len(sa), len(sb_)))
    return {"queries": n, "same_top_set_frac": same_set / n, "same_ranked_list_frac": same_list / n,
            "mean_set_overlap_at_21": overlap / n}


def bind_to_gpu_numa(local_rank):
    """Pin this process (and so its pinned staging buffers, first touch) to the NUMA node of its GPU."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return {"node": None, "note": "no NUMA information for the GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001 — topology files differ between boxes; binding is best effort
        return {"node": None, "note": f"not bound ({type(e).__name__})"}


class NvlinkCounters:
    """NVML NVLink data counters of one GPU (bytes, all links); `why` says what failed when the driver hides them."""

    def __init__(self, cuda_index):
        self.h, self.why = None, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            p = torch.cuda.get_device_properties(cuda_index)
            bus = f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            all_links = 0xFFFFFFFF
            self.fields = [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, all_links),
                           (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, all_links)]
        except Exception as e:  # noqa: BLE001
            self.h, self.why = None, f"{type(e).__name__}: {e}"

    def read(self):
        if self.h is None:
            return None
        try:
            vals = self.nv.nvmlDeviceGetFieldValues(self.h, self.fields)
            out = []
            for v in vals:
                if v.nvmlReturn != 0:
                    self.why = f"nvmlDeviceGetFieldValues: nvmlReturn {v.nvmlReturn}"
                    return None
                out.append(int(v.value.ullVal) * 1024)          # the counters tick in KiB
            return out          # [rx bytes, tx bytes]
        except Exception as e:  # noqa: BLE001
            self.why = f"{type(e).__name__}: {e}"
            return None


# ====================================================================================== reference (CPU) arm
def reference_arm(args, rank):
    """The reference's own CPU implementation of the path (C++ restatement, oracle/) on all host cores, on a prefix of
    the very batches the B200 arm times.  Maps only the generator library and the oracle — not libvmis_b200.so."""
    if rank != 0:
        return 0
    from serenade_b200 import synth
    from oracle import vmis_oracle as vo
    n_items, n_sessions = WORKLOADS[args.workload]
    if args.workload in DEVICE_BUILT:
        print(json.dumps({"impl": "reference", "unavailable": "no CPU arm at this size (index is generated on the device)"}))
        return 0
    t0 = time.time()
    items, off, ts = synth.synth_sessions(42, n_items, n_sessions)
    oix = vo.OracleIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W)
    log(f"[reference] data + CPU index in {time.time() - t0:.1f}s")
    threads = os.cpu_count() or 1
    first = synth.synth_queries(43, n_items, args.batch, MAX_ITEMS)       # batch 0 of rank 0 of the B200 arm
    cfg = base_config(args.workload, int(len(items)), args.batch, args.gpus, length_hist(first[1]))
    probe_n = min(2048, args.batch)
    probe = oix.predict_batch(first[0][:first[1][probe_n]], first[1][:probe_n + 1], K, M, HOW_MANY, False, mode=0,
                              threads=threads, want_outputs=False)
    rate = probe_n / probe[3]
    # each step: a prefix of the step's batch sized so that the whole run lasts about 2 x cpu-seconds
    per_step = int(max(512, min(args.batch, rate * (2 * args.cpu_seconds) / max(1, args.steps + args.warmup))))
    tot_q, tot_t = 0, 0.0
    for s in range(args.warmup + args.steps):
        qs = first if s == 0 else synth.synth_queries(43 + s, n_items, args.batch, MAX_ITEMS)
        r = oix.predict_batch(qs[0][:qs[1][per_step]], qs[1][:per_step + 1], K, M, HOW_MANY, False, mode=0,
                              threads=threads, want_outputs=False)
        if s >= args.warmup:
            tot_q += per_step
            tot_t += r[3]
    v = tot_q / tot_t
    sample = (f"first {per_step} queries of each of the {args.steps} timed batches of the B200 arm "
              f"(seed 43 + warmup + step, {args.batch} queries each)")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/i32+f64",
                      "data": "synthetic", "config": cfg,
                      "cpu_baseline": {"value": v, "unit": "queries/s", "cores": threads, "kind": "port",
                                       "sample": sample,
                                       "single_thread_latency_us": cpu_single_thread_latency(oix, first)},
                      "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


# ====================================================================================== B200 arm
class Runner:
    """Times vmis_predict_batch_device / vmis_predict_batch on one index handle."""

    def __init__(self, lib, torch, dist, dev, local_rank, world, n_out=HOW_MANY):
        self.lib, self.torch, self.dist, self.dev, self.local_rank, self.world, self.n = lib, torch, dist, dev, local_rank, world, n_out
        self.stream = torch.cuda.Stream(device=dev)         # kernels and timing events share this stream
        torch.cuda.set_stream(self.stream)
        self.sptr = C.c_void_p(self.stream.cuda_stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        from serenade_b200.shard import max_over_ranks
        return max_over_ranks(x, device=self.dev)

    def device_buffers(self, B):
        t = self.torch
        return (t.zeros((B, self.n), dtype=t.int64, device=self.dev), t.zeros((B, self.n), dtype=t.float64, device=self.dev),
                t.zeros(B, dtype=t.int32, device=self.dev), t.zeros((B, 4), dtype=t.int32, device=self.dev))

    def to_device(self, batches):
        t = self.torch
        return [(t.from_numpy(qi.view(np.int64)).to(self.dev), t.from_numpy(qo.view(np.int32)).to(self.dev)) for qi, qo in batches]

    def launch(self, gix, d_batch, B, out, stats=False):
        di, do = d_batch
        ids, sc, cnt, st = out
        rc = self.lib.vmis_predict_batch_device(gix.handle, di.data_ptr(), do.data_ptr(), B, K, M, self.n, 0, ids.data_ptr(),
                                                sc.data_ptr(), cnt.data_ptr(), st.data_ptr() if stats else None, self.sptr)
        if rc != 0:
            raise RuntimeError(self.lib.vmis_last_error().decode())

    def time_device(self, gix, d_batches, B, out, warmup, steps, launches_per_step=1):
        """`value`: inputs resident in HBM, CUDA events on the launch stream, max over ranks.  A step may be several
        launches (config 2: 1024 sessions per launch)."""
        t = self.torch
        for b in range(warmup):
            for l in range(launches_per_step):
                self.launch(gix, d_batches[b * launches_per_step + l], B, out)
        sampler = ClockSampler(self.local_rank)
        sampler.start()
        sampler.wait_for_samples(2)
        self.barrier()
        ev = [t.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(self.stream)
        for s in range(steps):
            for l in range(launches_per_step):
                self.launch(gix, d_batches[(warmup + s) * launches_per_step + l], B, out)
            ev[s + 1].record(self.stream)
        self.barrier()
        sampler.wait_for_samples(6)
        clocks = sampler.stop()
        step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(steps)]
        total_ms = self.max_over_ranks(ev[0].elapsed_time(ev[-1]))
        return {"qps": self.world * steps * launches_per_step * B / (total_ms * 1e-3), "ms_per_step": total_ms / steps,
                "step_ms": step_ms, "clocks": clocks}

    def time_e2e(self, gix, batches, B, warmup, steps):
        """`e2e`: the same metric through vmis_predict_batch with pinned HOST buffers, copies inside the timed region."""
        t, n = self.torch, self.n
        h_q = [(t.from_numpy(qi.view(np.int64)).pin_memory(), t.from_numpy(qo.view(np.int32)).pin_memory()) for qi, qo in batches]
        h_ids = t.zeros((B, n), dtype=t.int64).pin_memory()
        h_sc = t.zeros((B, n), dtype=t.float64).pin_memory()
        h_cnt = t.zeros(B, dtype=t.int32).pin_memory()
        u64p, u32p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_double)

        def call(b):
            qi, qo = h_q[b]
            rc = self.lib.vmis_predict_batch(gix.handle, C.cast(qi.data_ptr(), u64p), C.cast(qo.data_ptr(), u32p), B, K, M, n, 0,
                                             C.cast(h_ids.data_ptr(), u64p), C.cast(h_sc.data_ptr(), f64p),
                                             C.cast(h_cnt.data_ptr(), u32p), None)
            if rc != 0:
                raise RuntimeError(self.lib.vmis_last_error().decode())

        for b in range(warmup):
            call(b)
        self.barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            call(warmup + s)
        t.cuda.synchronize()
        mine = time.perf_counter() - t0
        worst = self.max_over_ranks(mine)
        h2d = int(np.mean([qi.numel() * 8 + qo.numel() * 4 for qi, qo in h_q[warmup:warmup + steps]]))
        return {"value": self.world * steps * B / worst, "unit": "queries/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": B * n * 16 + B * 4, "rank_value": steps * B / mine}


def b200_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import serenade_b200 as sb
    lib = sb.load_library()
    run = Runner(lib, torch, dist, dev, local_rank, world)
    n_items, n_sessions = WORKLOADS[args.workload]
    B, W, S = args.batch, args.warmup, args.steps
    sections_on = set() if args.sections == "none" else set(args.sections.split(","))
    want = lambda name: "all" in sections_on or name in sections_on  # noqa: E731

    # ------------------------------------------------------------------ headline: query-sharded replicas
    t0 = time.time()
    device_built = args.workload in DEVICE_BUILT or args.device_build
    index_info = {}
    if device_built:
        items = off = ts = None
        t1 = time.time()
        gix = sb.VMISIndex.synth(42, n_items, n_sessions, M, MAX_LEN, IDF_W, local_rank, 0, 1)
        interactions = int(gix.stats()["n_pairs_kept"])
        index_info["build"] = "generated and indexed on the device (vmis_index_synth)"
    else:
        items, off, ts = sb.synth_sessions(42, n_items, n_sessions)
        interactions = int(len(items))
        t1 = time.time()
        gix = sb.VMISIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W, device=local_rank)
    st = gix.stats()
    log(f"[rank {rank}] synth {t1 - t0:.1f}s, index build+upload {time.time() - t1:.1f}s, "
        f"{st['device_bytes'] / 1e6:.0f} MB in HBM, {st['n_items']} items, {st['n_postings']} postings")
    index_info.update({"hbm_bytes": int(st["device_bytes"]), "items": int(st["n_items"]), "postings": int(st["n_postings"]),
                       "build_s": round(time.time() - t1, 2)})

    n_batches = W + S
    batches = [sb.synth_queries(43 + 1000 * rank + b, n_items, B, MAX_ITEMS) for b in range(n_batches)]
    cfg = base_config(args.workload, interactions, B, world, length_hist(batches[0][1]))
    d_batches = run.to_device(batches)
    out = run.device_buffers(B)

    # untimed stats pass → exact algorithmic bytes of each timed batch (also part of warm-up)
    alg_bytes = []
    for b in range(n_batches):
        run.launch(gix, d_batches[b], B, out, stats=True)
        torch.cuda.synchronize()
        alg_bytes.append(algorithmic_bytes(out[3].cpu().numpy(), batches[b][1]))
    head = run.time_device(gix, d_batches, B, out, W, S)

    # roofline of the (single) kernel: algorithmic bytes of the timed batches / their kernel time
    timed_bytes = sum(alg_bytes[W:])
    achieved = timed_bytes / (sum(head["step_ms"]) * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    else:
        peak, peak_src = 6650.0, "fallback"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("workload") == args.workload and tj.get("queries_per_launch") == B:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "vmis_predict_kernel",
                "algorithmic_bytes_per_query": timed_bytes / (S * B)}

    e2e = run.time_e2e(gix, batches, B, W, S)
    e2e.update({"launches_per_step": (B + (1 << 15) - 1) >> 15,
                "note": "vmis_predict_batch pipelines the batch in chunks of 2^15 sessions over 3 streams"})
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, "e2e_qps": e2e["rank_value"], "numa": numa})
        e2e["per_rank"] = per_rank
    e2e.pop("rank_value")

    result = {"metric": METRIC, "value": head["qps"], "unit": "queries/s", "n_gpus": world, "steps": S, "warmup": W,
              "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "u32/i32+f64", "data": "synthetic", "config": cfg, "clocks": head["clocks"], "gpu_launches": S,
              "e2e": e2e, "roofline": roofline, "step_ms": [round(x, 3) for x in head["step_ms"]], "index": index_info,
              "numa": numa, "sections": {}}

    # The headline is complete.  The extra sections below must never cost it: each runs under try/except, and a
    # watchdog prints the line as it stands (and ends the process) if a section hangs, e.g. in a collective.
    emitted = threading.Event()

    def emit():
        if not emitted.is_set():
            emitted.set()
            if rank == 0:
                print(json.dumps(result), flush=True)

    def watchdog():
        if not emitted.wait(args.section_timeout):
            result["sections"]["watchdog"] = f"extra sections did not finish within {args.section_timeout:.0f}s; line printed as is"
            emit()
            os._exit(0)

    threading.Thread(target=watchdog, daemon=True).start()

    def guarded(name, fn, *a):
        try:
            return fn(*a)
        except Exception as e:  # noqa: BLE001 — an extra section must not take the headline down
            log(f"[rank {rank}] section {name} failed: {type(e).__name__}: {e}")
            return {"error": f"{type(e).__name__}: {e}"}

    # ------------------------------------------------------------------ N = 1 extras
    if world == 1 and want("latency"):
        result["latency"] = guarded("latency", latency_section, sb, lib, gix, batches[0])
    if world == 1 and not args.no_cpu_baseline and items is not None:
        from oracle import vmis_oracle as vo
        t2 = time.time()
        oix = vo.OracleIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W)
        threads = os.cpu_count() or 1
        qps, n_sample, secs = cpu_reference(oix, batches[W], threads, args.cpu_seconds)
        log(f"[cpu] oracle index {time.time() - t2 - secs:.1f}s; {n_sample} queries in {secs:.1f}s on {threads} threads")
        result["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                                  "sample": f"first {n_sample} queries of the first timed batch, faithful mode, "
                                            f"{threads} threads sharing one index",
                                  "single_thread_latency_us": cpu_single_thread_latency(oix, batches[W])}
        result["parity_check"] = parity_vs_oracle(run, gix, oix, batches[W], d_batches[W], B, out, args.parity_queries, threads)
        del oix
    if world == 1 and want("config2"):
        result["sections"]["config2"] = guarded("config2", section_config2, sb, run, args)
    if world == 1 and want("config4"):
        result["sections"]["config4"] = guarded("config4", section_device_built, sb, run, "synthetic-582M-6.5M", local_rank, rank, world, args)

    # ------------------------------------------------------------------ N > 1: item-sharded postings
    if world > 1 and want("item_sharded") and items is not None:
        result["sections"]["item_sharded"] = guarded("item_sharded", section_item_sharded, sb, run, gix, items, off, ts, batches,
                                                     d_batches, B, out, rank, world, local_rank, args)
    gix.close()
    del d_batches
    torch.cuda.empty_cache()
    if (world == 8 or args.config5_workload) and world > 1 and want("config5"):
        result["sections"]["config5"] = guarded("config5", section_device_built, sb, run,
                                                args.config5_workload or "synthetic-2.3B-6.5M", local_rank, rank, world, args)
    emit()
    if world > 1:
        dist.destroy_process_group()
    return 0


def parity_vs_oracle(run, gix, oix, batch, d_batch, B, out, n_chk, threads):
    """The same sample through the canonical oracle (must be bit-exact) and the faithful one (agreement measured)."""
    from oracle import vmis_oracle as vo
    qi, qo = batch
    n_chk = min(n_chk, B)
    oids, osc, ocnt, _, _ = oix.predict_batch(qi[:qo[n_chk]], qo[:n_chk + 1], K, M, HOW_MANY, False, mode=vo.CANONICAL, threads=threads)
    run.launch(gix, d_batch, B, out)
    run.torch.cuda.synchronize()
    g_ids = out[0][:n_chk].cpu().numpy().view(np.uint64)
    ok = (np.array_equal(g_ids, oids) and np.array_equal(out[1][:n_chk].cpu().numpy(), osc) and
          np.array_equal(out[2][:n_chk].cpu().numpy().view(np.uint32), ocnt))
    n_f = min(n_chk, 4096)
    fids, _, fcnt, _, _ = oix.predict_batch(qi[:qo[n_f]], qo[:n_f + 1], K, M, HOW_MANY, False, mode=vo.FAITHFUL, threads=threads)
    return {"queries": n_chk, "bit_exact_vs_canonical_oracle": bool(ok),
            "vs_faithful_restatement": faithful_agreement(oids[:n_f], ocnt[:n_f], fids, fcnt)}


def latency_section(sb, lib, gix, batch):
    """The reference's own call shape: one evolving session per predict() call (mod.rs:118-125), outside every timed
    region.  (a) a lone caller; (b) many callers through the micro-batcher under open-loop load — the reference's
    README quotes "< 1.7 ms p90" for the prediction and "1000 predictions/s on 2 vCPU" (README.md:16-17)."""
    u64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    qi0, qo0 = batch
    ids1 = np.zeros(HOW_MANY, dtype=np.uint64)
    sc1 = np.zeros(HOW_MANY, dtype=np.float64)
    lat = []
    for q in range(400):
        ev = np.ascontiguousarray(qi0[qo0[q]:qo0[q + 1]])
        t_a = time.perf_counter()
        lib.vmis_predict(gix.handle, ev.ctypes.data_as(u64p), len(ev), K, M, HOW_MANY, 0, ids1.ctypes.data_as(u64p),
                         sc1.ctypes.data_as(f64p))
        lat.append((time.perf_counter() - t_a) * 1e6)
    lat = np.sort(np.array(lat[100:]))
    pct = lambda v: {p: round(float(np.percentile(v, float(p[1:]))), 1) for p in ("p50", "p90", "p99")}
    # the same calls from a native loop (what a Rust / C++ host sees: no interpreter between the calls)
    n_nat = 2000
    nat = np.zeros(n_nat, dtype=np.float32)
    n_lat_q = min(len(qo0) - 1, 4096)
    rc = lib.vmis_predict_latency_test(gix.handle, qi0.ctypes.data_as(u64p), qo0.ctypes.data_as(C.POINTER(C.c_uint32)), n_lat_q,
                                       K, M, HOW_MANY, 0, n_nat, nat.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != n_nat:
        raise RuntimeError(f"vmis_predict_latency_test: {rc}")
    out = {"single_call_us": pct(nat[200:]), "single_call_via_python_us": pct(lat),
           "note": "vmis_predict, one evolving session per call, host buffers; native loop (vmis_predict_latency_test) and "
                   "the same call through ctypes",
           "reference_readme": "p90 < 1.7 ms per prediction; 1000 predictions/s on 2 vCPU (README.md:16-17)"}
    n_q = min(len(qo0) - 1, 1 << 16)
    sub = (qi0[:qo0[n_q]], qo0[:n_q + 1])
    threads = min(256, 8 * (os.cpu_count() or 8))
    b = sb.Batcher(gix, K, M, HOW_MANY, False, max_batch=4096, max_wait_us=50)
    rows = []
    try:
        b.load_test(sub, 20_000, 300, threads)                      # warm-up
        for rps in (10_000, 100_000, 300_000, 1_000_000):
            l, achieved = b.load_test(sub, rps, 1000, threads)
            if len(l) == 0:
                continue
            rows.append({"offered_rps": rps, "achieved_rps": round(achieved), "p50_us": round(float(np.percentile(l, 50)), 1),
                         "p90_us": round(float(np.percentile(l, 90)), 1), "p99_us": round(float(np.percentile(l, 99)), 1)})
        st = b.stats()
        out["batcher"] = {"caller_threads": threads, "max_batch": 4096, "max_wait_us": 50, "open_loop": rows,
                          "mean_batch": round(st["requests"] / max(1, st["batches"]), 1),
                          "note": "vmis_batcher_predict, one evolving session per call; latency counted from the request's "
                                  "due time (open loop), host callers on this box's cores"}
    finally:
        b.close()
    return out


def section_config2(sb, run, args):
    """BASELINE configs[1]: 1 M interactions / 50 k items (index ~17 MB, L2 resident), launches of 1024 evolving sessions."""
    torch = run.torch
    n_items, n_sessions = WORKLOADS["synthetic-1M-50k"]
    items, off, ts = sb.synth_sessions(42, n_items, n_sessions)
    gix = sb.VMISIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W, device=run.local_rank)
    Bs, per_step, W, S = 1024, 64, 3, 8
    batches = [sb.synth_queries(4300 + b, n_items, Bs, MAX_ITEMS) for b in range((W + S) * per_step)]
    d_batches = run.to_device(batches)
    out = run.device_buffers(Bs)
    r = run.time_device(gix, d_batches, Bs, out, W, S, launches_per_step=per_step)
    e2e = run.time_e2e(gix, batches, Bs, W * per_step, S * per_step)
    e2e.pop("rank_value")
    sec = {"workload": "synthetic-1M-50k", "interactions": int(len(items)), "batch": Bs, "launches_per_step": per_step,
           "value": r["qps"], "unit": "queries/s", "us_per_launch": 1e3 * r["ms_per_step"] / per_step, "clocks": r["clocks"],
           "e2e": {"value": e2e["value"], "unit": "queries/s", "note": "one vmis_predict_batch call per 1024 sessions"}}
    if not args.no_cpu_baseline:
        from oracle import vmis_oracle as vo
        oix = vo.OracleIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W)
        qi, qo = batches[W * per_step]
        oids, osc, ocnt, _, _ = oix.predict_batch(qi, qo, K, M, HOW_MANY, False, mode=vo.CANONICAL, threads=os.cpu_count() or 1)
        run.launch(gix, d_batches[W * per_step], Bs, out)
        torch.cuda.synchronize()
        sec["parity_check"] = {"queries": Bs, "bit_exact_vs_canonical_oracle": bool(
            np.array_equal(out[0].cpu().numpy().view(np.uint64), oids) and np.array_equal(out[1].cpu().numpy(), osc) and
            np.array_equal(out[2].cpu().numpy().view(np.uint32), ocnt))}
    gix.close()
    return sec


def section_device_built(sb, run, workload, local_rank, rank, world, args):
    """BASELINE configs[3] (N = 1: 582 M interactions, one GPU) and configs[4] (N = 8: 2.3 B interactions, postings
    item-sharded over the GPUs): generated and indexed on the device; the sharded result is checked bit for bit
    against the unsharded index of the same data built on rank 0."""
    torch, dist = run.torch, run.dist
    n_items, n_sessions = WORKLOADS[workload]
    t0 = time.time()
    shard = (rank, world) if world > 1 else (0, 1)
    gix = sb.VMISIndex.synth(42, n_items, n_sessions, M, MAX_LEN, IDF_W, local_rank, *shard)
    if world > 1:
        gix.connect_shards(rank, world)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    st = gix.stats()
    Bs, W, S = args.batch // 2, 3, 4
    batches = [sb.synth_queries(43 + 1000 * rank + b, n_items, Bs, MAX_ITEMS) for b in range(W + S)]
    d_batches = run.to_device(batches)
    out = run.device_buffers(Bs)
    r = run.time_device(gix, d_batches, Bs, out, W, S)
    e2e = run.time_e2e(gix, batches, Bs, W, S)
    sec = {"workload": workload, "interactions": int(st["n_pairs_kept"]), "items": int(st["n_items"]),
           "index_hbm_bytes_per_gpu": int(st["device_bytes"]), "build_s": round(build_s, 2), "batch_per_gpu": Bs,
           "steps": S, "warmup": W, "value": r["qps"], "unit": "queries/s", "ms_per_step": r["ms_per_step"],
           "clocks": r["clocks"], "e2e": {"value": e2e["value"], "unit": "queries/s"},
           "parallelism": (f"item-sharded postings x{world} (peer HBM over NVLink, CUDA IPC), queries sharded"
                           if world > 1 else "one GPU")}
    # size-independent checks on the full batch: counts in range, scores descending, ids distinct per row
    run.launch(gix, d_batches[W], Bs, out)
    torch.cuda.synchronize()
    ids, sc, cnt = out[0], out[1], out[2]
    col = torch.arange(HOW_MANY, device=run.dev)[None, :]
    valid = col < cnt[:, None]
    desc = bool(((sc[:, 1:] <= sc[:, :-1]) | ~valid[:, 1:]).all())
    srt = torch.sort(torch.where(valid, ids, -1 - col.expand_as(ids)), dim=1).values
    distinct = bool((srt[:, 1:] != srt[:, :-1]).all())
    sec["properties"] = {"queries": Bs, "scores_descending": desc, "ids_distinct": distinct,
                         "counts_in_range": bool((cnt <= HOW_MANY).all())}
    if world > 1:
        # rank 0 also builds the UNSHARDED index of the same data and must get identical rows
        ok = True
        if rank == 0:
            n_chk = min(Bs, 1 << 16)
            ref = sb.VMISIndex.synth(42, n_items, n_sessions, M, MAX_LEN, IDF_W, local_rank, 0, 1)
            out2 = run.device_buffers(Bs)
            run.launch(ref, d_batches[W], Bs, out2)
            torch.cuda.synchronize()
            ok = bool(torch.equal(out[0][:n_chk], out2[0][:n_chk]) and torch.equal(out[1][:n_chk], out2[1][:n_chk]) and
                      torch.equal(out[2][:n_chk], out2[2][:n_chk]))
            sec["parity_check"] = {"queries": n_chk, "bit_exact_vs_unsharded_index": ok}
            ref.close()
            del out2
        dist.barrier()
    gix.close()
    del d_batches, out
    torch.cuda.empty_cache()
    return sec


def section_item_sharded(sb, run, replica, items, off, ts, batches, d_batches, B, out, rank, world, local_rank, args):
    """The headline workload with the postings item-sharded over the N ranks (BASELINE configs[4] layout): the same
    kernel reads remote posting lists from peer HBM over NVLink.  Timed like the headline; every rank checks its whole
    first timed batch bit for bit against its replica index."""
    torch, dist = run.torch, run.dist
    W, S = args.warmup, args.steps
    t0 = time.time()
    gix = sb.VMISIndex.from_sessions_sharded(items, off, ts, M, MAX_LEN, IDF_W, local_rank, rank, world)
    gix.connect_shards(rank, world)
    build_s = time.time() - t0
    nvl = NvlinkCounters(local_rank)
    # statistics pass on the sharded handle: postings visited → algorithmic NVLink bytes
    run.launch(gix, d_batches[W], B, out, stats=True)
    torch.cuda.synchronize()
    postings = int(out[3][:, 0].to(torch.int64).sum().item())
    c0 = nvl.read()
    r = run.time_device(gix, d_batches, B, out, W, S)
    c1 = nvl.read()
    e2e = run.time_e2e(gix, batches, B, W, S)
    sec = {"workload": args.workload, "parallelism": f"item-sharded postings x{world} (peer HBM over NVLink, CUDA IPC), queries sharded",
           "index_hbm_bytes_per_gpu": int(gix.stats()["device_bytes"]), "build_s": round(build_s, 2), "batch_per_gpu": B,
           "steps": S, "warmup": W, "value": r["qps"], "unit": "queries/s", "ms_per_step": r["ms_per_step"],
           "clocks": r["clocks"], "e2e": {"value": e2e["value"], "unit": "queries/s"},
           "nvlink": {"algorithmic_remote_bytes_per_query": round(postings * 4 * (world - 1) / world / B, 1),
                      "note": "4-byte postings of the items whose shard lives on a peer: (N-1)/N of the postings visited"}}
    if c0 is not None and c1 is not None:
        n_q = B * (W + S)
        sec["nvlink"].update({"measured_rx_bytes_per_query": round((c1[0] - c0[0]) / n_q, 1),
                              "measured_tx_bytes_per_query": round((c1[1] - c0[1]) / n_q, 1),
                              "source": "NVML NVLINK_THROUGHPUT_DATA_RX/TX of this rank's GPU around warm-up + timed steps"})
    else:
        sec["nvlink"].update({"measured_rx_bytes_per_query": None, "why": nvl.why})
    # parity: sharded vs replica on the whole first timed batch, every rank
    out2 = run.device_buffers(B)
    run.launch(gix, d_batches[W], B, out)
    run.launch(replica, d_batches[W], B, out2)
    torch.cuda.synchronize()
    ok = torch.tensor([int(torch.equal(out[0], out2[0]) and torch.equal(out[1], out2[1]) and torch.equal(out[2], out2[2]))],
                      device=run.dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    sec["parity_check"] = {"queries_per_rank": B, "bit_exact_vs_replica_index_all_ranks": bool(ok.item())}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import vmis_oracle as vo
        oix = vo.OracleIndex.from_sessions(items, off, ts, M, MAX_LEN, IDF_W)
        n_chk = min(args.parity_queries, B)
        qi, qo = batches[W]
        oids, osc, ocnt, _, _ = oix.predict_batch(qi[:qo[n_chk]], qo[:n_chk + 1], K, M, HOW_MANY, False, mode=vo.CANONICAL,
                                                  threads=os.cpu_count() or 1)
        sec["parity_check"].update({"oracle_queries": n_chk, "bit_exact_vs_canonical_oracle": bool(
            np.array_equal(out[0][:n_chk].cpu().numpy().view(np.uint64), oids) and
            np.array_equal(out[1][:n_chk].cpu().numpy(), osc) and
            np.array_equal(out[2][:n_chk].cpu().numpy().view(np.uint32), ocnt))})
        del oix
    dist.barrier()
    gix.close()
    del out2
    return sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="synthetic-60M-1.76M", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1 << 20, help="evolving sessions per step (per GPU)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    ap.add_argument("--parity-queries", type=int, default=4096, help="queries of the in-bench oracle check")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip everything that runs the CPU oracle")
    ap.add_argument("--device-build", action="store_true", help="generate + index the workload on the device")
    ap.add_argument("--config5-workload", default=None, choices=sorted(DEVICE_BUILT),
                    help="run the config5 section (item-sharded, device-built) at any N > 1 on this workload")
    ap.add_argument("--section-timeout", type=float, default=420.0,
                    help="seconds the extra sections may take before the headline line is printed without them")
    ap.add_argument("--sections", default="all",
                    help="extra sections: all | none | comma list of latency,config2,config4,item_sharded,config5")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: --warmup < 3 breaks the timing rules; use >= 3 for a reported number")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank)
    return b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
