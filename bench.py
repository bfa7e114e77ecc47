#!/usr/bin/env python
"""bench.py — headline benchmark of the vecgo scan hot path on B200, plus every BASELINE config.

Headline workload (BASELINE.json configs[1]): SQ8 decode-and-scan, 10M x 768-d, 10k-query
batch, k=100.  A "step" is one pass of the hot path over one query batch: every
query against every row, fused top-k (+ for N>1 the NCCL all-gather of the
per-shard top-k and the device merge).  Rows are sharded across the N GPUs
(strong scaling: the database size is fixed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload sq8|int4] [--rows R] [--queries Q] [--k K] [--configs all|none|c1,c2b,c3,c4,c5]

Prints ONE JSON line (rank 0).  `value` = queries/s with inputs resident in HBM
(CUDA events, max over ranks); `e2e` = the same through the C ABI with host
buffers (H2D of the query batch and D2H of the results inside the timed region).
`configs` holds one time-boxed entry per BASELINE config (C1 Flat, C2b INT4, C3 PQ, C4 RaBitQ + rerank, C5 PQ
training), each with value, roofline, parity against the oracle and the CPU baseline on a bounded sample; under
--gpus N the sharded configs (C3: 25M rows per GPU = configs[2] at N=8; C4: 12.5M rows per GPU = configs[3] at N=8;
C5: subspaces split across the GPUs) run across the N ranks.
`--impl reference` times vecgo's own CPU path: the reference's AVX-512 C kernels
(oracle/_ref, built from /root/reference by oracle/Makefile) driven by the
flat.Search scan loop + CandidateHeap restatement, one query per worker thread
like Engine.BatchSearch, on a bounded row/query sample (linear extrapolation in
rows, stated in `cpu_baseline.sample`).

Data: SURVEY 8(d) seeds (NumPy PCG64: data 42, queries 43, benchmark_test/config_test.go:32,62,93).  Queries of every
config and the first SUB rows of every database are generated with those NumPy seeds on the host — that prefix is what
the CPU arm scans and what the oracle parity checks use, so both arms see the same rows; the remaining rows of the
multi-GB databases are drawn on the device (torch, same distribution; a host generator would take minutes).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DATA_SEED, QUERY_SEED = 42, 43  # benchmark_test/config_test.go:32,62,93
CHUNK = 262_144                 # generation / encode chunk (rows)
SUB = 200_000                   # NumPy-seeded prefix of every database: shared by the GPU arm, the CPU arm and the oracle checks
TRAIN_ROWS = 1_048_576          # sq.Train sample = the first 1M rows
F = np.float32


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="sq8", choices=["sq8", "int4"])
    p.add_argument("--rows", type=int, default=10_000_000)
    p.add_argument("--queries", type=int, default=10_000)
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--k", type=int, default=100)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=15.0)
    p.add_argument("--configs", default="all", help="all | none | comma list of c1,c2b,c3,c4,c5")
    p.add_argument("--config-cpu-seconds", type=float, default=6.0, help="CPU-baseline budget per config")
    p.add_argument("--c4-rows", type=int, default=12_500_000,
                   help="C4 rows per GPU (default: one shard of the 8-GPU layout of the 100M-row database; 25000000 on 4 GPUs is "
                        "the same database with 153.6 GB of float32 rows per GPU)")
    p.add_argument("--small", action="store_true", help="reduced config sizes (development only; the JSON says so)")
    return p.parse_args()


def peaks():
    """(HBM GB/s, dense bf16/fp16 TFLOP/s sustained, burst, source).  kind::f16 with fp16 operands runs at the bf16 rate."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
        return (float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), float(j["bf16_tflops"]),
                "measured (MEASURED_PEAKS.json: bf16_tflops_sustained for a kernel timed inside a long step, bf16_tflops (burst) for a "
                "kernel timed alone, hbm_gbs)")
    return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region, sampled in-process through NVML
    (nvidia_ml_py — the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`
    prints, without spawning a process that contends for the driver while kernels are timed)."""

    def __init__(self, index: int, period_s: float = 0.2):
        self.index, self.period, self.samples = index, period_s, []
        self._stop = threading.Event()
        self._thr = None
        self.h = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.h is not None and os.environ.get("BENCH_NO_CLOCKS") != "1":
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(s[2] & bit for s in self.samples))
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": float(max(s[1] for s in self.samples)),
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------ shared host data (NumPy PCG64 seeds)
def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_HOST_CACHE = {}


def np_rows(n, dim, seed, kind="normal"):
    """SURVEY 8(d) generators: NumPy PCG64 with the config's seed; N(0,1) or U[0,1) float32 (testutil.go:69-175)."""
    key = (n, dim, seed, kind)
    if key not in _HOST_CACHE:
        rng = np.random.default_rng(seed)
        _HOST_CACHE[key] = rng.standard_normal((n, dim), dtype=F) if kind == "normal" else rng.random((n, dim), dtype=F)
    return _HOST_CACHE[key]


def sq8_encode_host(x, mins, maxs, sc):
    return np.clip((np.clip(x, mins, maxs) - mins) * sc + F(0.5), 0, 255).astype(np.uint8)


def int4_encode_host(x, minv, diff):
    nrm = np.clip((x - minv) / diff, 0, 1)
    qn = np.floor(nrm.astype(np.float64) * 15 + 0.5).astype(np.uint8)
    return ((qn[:, 0::2] << 4) | qn[:, 1::2]).astype(np.uint8)


def make_cpu_sample(workload, rows, nq, dim):
    """The NumPy-seeded prefix of the GPU workload: the first `rows` rows and `nq` queries; codes from the oracle's encoder
    with parameters trained on those rows."""
    from oracle import oracle as o

    x = np_rows(SUB, dim, DATA_SEED)[:rows]
    q = np_rows(max(nq, 4096), dim, QUERY_SEED)[:nq]
    if workload == "sq8":
        mins, maxs, sc, inv = (np.zeros(dim, F) for _ in range(4))
        o.lib.vgo_sq8_train(o.fp(x), rows, dim, o.fp(mins), o.fp(maxs), o.fp(sc), o.fp(inv))
        return dict(q=q, codes=sq8_encode_host(x, mins, maxs, sc), mins=mins, inv=inv)
    minv, diff = np.zeros(dim, F), np.zeros(dim, F)
    o.lib.vgo_int4_train(o.fp(x), rows, dim, o.fp(minv), o.fp(diff))
    return dict(q=q, codes=int4_encode_host(x, minv, diff), minv=minv, diff=diff)


def cpu_arm(workload, dim, k, seconds, full_rows, steps=1, warmup=1):
    """Times the reference's CPU path on this host: `warmup` + `steps` passes over a bounded sample (each pass ~ seconds /
    steps of CPU work; successive passes take successive query slices).  Returns dict(value = QPS extrapolated to
    full_rows, step_ms = measured time of one sample pass, ...)."""
    from oracle import oracle as o

    threads = host_threads()
    kind = "reference" if o.ref is not None else "port"
    sample_rows = SUB
    s = make_cpu_sample(workload, sample_rows, 4096, dim)
    kern = o.ref_kernels() if o.ref is not None else o.oracle_kernels()
    seg = o.FlatOracle(dim=dim, metric=0, quant=1, codes=s["codes"], mins=s["mins"], inv=s["inv"], kernels=kern) if workload == "sq8" else None

    def run(q):
        t0 = time.perf_counter()
        if workload == "sq8":
            seg.search_batch(q, k, threads=threads)
        else:
            fn = o.fn_addr(o.ref.int4L2DistanceBatchAvx512) if o.ref is not None else o.fn_addr(o.lib.vgo_int4_l2_batch_a512)
            out = np.zeros((len(q), k), o.cand_dtype)
            cnt = np.zeros(len(q), np.int64)
            o.lib.vgo_int4_search_batch(o.fp(q), len(q), o.bp(s["codes"]), sample_rows, dim, o.fp(s["minv"]), o.fp(s["diff"]), k, fn,
                                        threads, out.ctypes.data_as(C.POINTER(o.Cand)), cnt.ctypes.data_as(o.i64p))
        return time.perf_counter() - t0

    nq0 = max(threads, 16)
    t = run(s["q"][:nq0])  # calibration pass (also warms caches)
    per_step = max(0.5, seconds / max(1, steps))
    nq_step = int(min(4096, max(nq0, nq0 * per_step / max(t, 1e-3))))
    nq_step = max(threads, nq_step // threads * threads)
    times = []
    for i in range(warmup + steps):
        off = (i * nq_step) % max(1, 4096 - nq_step + 1)
        dt = run(s["q"][off:off + nq_step])
        if i >= warmup:
            times.append(dt)
    t2 = float(np.sum(times))
    nq2 = nq_step * steps
    qps_sample = nq2 / t2
    qps_full = qps_sample * sample_rows / full_rows
    return {"value": qps_full, "unit": "queries/s", "cores": threads, "kind": kind,
            "sample": f"{steps} passes of {nq_step} queries x {sample_rows} rows x {dim}-d {workload} codes (the NumPy seed-{DATA_SEED}/{QUERY_SEED} prefix "
                      f"the GPU arm also holds) in {t2:.1f}s on {threads} threads "
                      f"({'reference AVX-512 C kernels (oracle/_ref)' if kind == 'reference' else 'oracle port'} + flat.Search loop/heap "
                      f"restatement, one query per worker); linear extrapolation x{sample_rows}/{full_rows} rows",
            "ns_per_row_per_thread": t2 * threads / (nq2 * sample_rows) * 1e9, "step_ms": t2 / steps * 1e3, "queries_per_step": nq_step,
            "rows_per_step": sample_rows}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, a.steps), max(0, a.warmup)
    cb = cpu_arm(a.workload, a.dim, a.k, max(5.0, min(60.0, 1.5 * steps)), a.rows, steps=steps, warmup=min(warm, 3))
    line = {
        "impl": "reference", "metric": f"batched QPS, {a.workload.upper()} decode-and-scan top-{a.k}", "value": cb["value"],
        "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["step_ms"],
        "ms_per_full_batch_extrapolated": a.queries / cb["value"] * 1e3,
        "step": f"one pass of the reference's scan over a bounded sample: {cb['queries_per_step']} queries x {cb['rows_per_step']} rows "
                f"(ms_per_step is the measured time of that pass; `value` extrapolates it linearly in rows to the {a.rows}-row workload)",
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a), "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(a):
    return {"workload": f"{a.workload.upper()} decode-and-scan, {a.rows} x {a.dim}-d codes, {a.queries}-query batch, k={a.k}, L2",
            "rows": a.rows, "dim": a.dim, "queries": a.queries, "k": a.k,
            "sharding": f"rows/{a.gpus} per GPU + NCCL all-gather top-k merge" if a.gpus > 1 else "single GPU",
            "l2_flush": "inputs larger than L2 (code matrix >= 0.9 GB per GPU vs 126 MB L2)"}


# ------------------------------------------------------------------ GPU arm: environment
class Env:
    """torch / torch.distributed / library handles of one rank."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist

        import vecgo_b200 as vg

        self.torch, self.dist, self.vg, self.L = torch, dist, vg, vg._lib
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            # NCCL's version / debug lines must not share stdout with the JSON line: fd 1 points at stderr while the
            # communicator is created (NCCL prints "NCCL version ..." to stdout at NCCL_DEBUG=VERSION and above)
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
            torch.cuda.set_device(self.local)
            dist.barrier()  # forces the communicator (and its banner) into existence now
            torch.cuda.synchronize()
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")
        self.L.call("vg_init", self.local)
        self.L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
        self.hbm_peak, self.tf_sustained, self.tf_burst, self.peak_src = peaks()
        self.flush_buf = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([float(v)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok: bool) -> bool:
        t = self.torch.tensor([1.0 if ok else 0.0], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def flush_l2(self):
        """Write a buffer twice the size of the 126 MB L2 (between timed iterations of workloads that fit in it)."""
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self.flush_buf.add_(1)

    def timed_steps(self, fn, steps, warmup, flush=False):
        """ms per step: `steps` back-to-back steps between one CUDA-event pair (barrier + synchronize on both sides, max
        over ranks); with flush=True every step is timed on its own after an L2 flush and the mean is returned."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            self.barrier()
            return self.max_over_ranks(e0.elapsed_time(e1)) / steps
        tot = 0.0
        for _ in range(steps):
            self.flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        self.barrier()
        return self.max_over_ranks(tot / steps)

    def out_bufs(self, nq, k):
        t = self.torch
        return (t.empty((nq, k), dtype=t.int32, device=self.dev), t.empty((nq, k), dtype=t.float32, device=self.dev),
                t.empty((nq,), dtype=t.int32, device=self.dev))

    def qtc_counters(self, enable=-1):
        ms, n, qn_, fb = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.L.call("vg_quant_tc_profile", enable, C.byref(ms), C.byref(n))
        self.L.call("vg_quant_tc_stats", C.byref(qn_), C.byref(fb))
        return ms.value, n.value, qn_.value, fb.value

    def flat_counters(self):
        qa, fb = C.c_uint64(), C.c_uint64()
        self.L.call("vg_flat_tc_stats", C.byref(qa), C.byref(fb))
        return qa.value, fb.value


def gen_chunk(env, chunk_idx, rows, dim, sub_dim=None):
    """Rows [chunk_idx*CHUNK, +rows) of a database: rows below SUB are the NumPy seed-42 prefix, the rest is drawn on the device."""
    torch = env.torch
    g = torch.Generator(device=env.dev).manual_seed(DATA_SEED * 1_000_003 + chunk_idx)
    x = torch.randn((rows, dim), dtype=torch.float32, device=env.dev, generator=g)
    r0 = chunk_idx * CHUNK
    if r0 < SUB:
        n = min(SUB - r0, rows)
        x[:n] = torch.from_numpy(np_rows(SUB, dim, DATA_SEED)[r0:r0 + n]).to(env.dev)
    return x


def same_results(torch, r1, s1, r2, s2):
    return bool(torch.equal(r1, r2)), bool(torch.equal(s1.view(torch.int32), s2.view(torch.int32)))


# ------------------------------------------------------------------ GPU arm: C2 (headline SQ8; C2b INT4)
class ScanDB:
    """An SQ8 / INT4 index over rows [lo, hi) of the synthetic database (quantizer trained on the first TRAIN_ROWS rows)."""

    def __init__(self, env, workload, rows, dim, lo, hi, keep_first_codes=0):
        vg, L, torch = env.vg, env.L, env.torch
        self.vg, self.L = vg, L
        self.workload, self.dim = workload, dim
        code_bytes = dim if workload == "sq8" else dim // 2
        train_rows = min(TRAIN_ROWS, rows)
        mins = np.full(dim, np.inf, F)
        maxs = np.full(dim, -np.inf, F)
        for c in range((train_rows + CHUNK - 1) // CHUNK):
            r = min(CHUNK, train_rows - c * CHUNK)
            x = gen_chunk(env, c, CHUNK, dim)[:r].contiguous()
            mn, mx = np.zeros(dim, F), np.zeros(dim, F)
            L.call("vg_minmax_dev", x.data_ptr(), r, dim, L.ptr(mn, L.f32p), L.ptr(mx, L.f32p))
            mins, maxs = np.minimum(mins, mn), np.maximum(maxs, mx)
        self.train_rows = train_rows
        if workload == "sq8":
            self.sq = vg.quantization.ScalarQuantizer(dim)
            self.sq.SetBounds(mins, maxs)  # min/max of the sample → scale = 255/(max-min) (Train's formulas for max>min)
            self.ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=hi - lo, row_base=lo, sq8=(self.sq.mins, self.sq.invScales))
        else:
            self.mins = mins
            self.diff = (maxs - mins).astype(F)
            self.diff[self.diff == 0] = 1.0
            self.ix = vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=dim, rows=hi - lo, row_base=lo, int4=(self.mins, self.diff))
        self.first_codes = None
        codes_chunk = torch.empty((CHUNK, code_bytes), dtype=torch.uint8, device=env.dev)
        for c in range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK):
            x = gen_chunk(env, c, CHUNK, dim)
            s, e = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
            xs = x[s - c * CHUNK:e - c * CHUNK].contiguous()
            self.encode(xs, e - s, codes_chunk)
            self.ix.upload_dev(e - s, d_codes=codes_chunk.data_ptr(), row0=s - lo)
            if keep_first_codes and self.first_codes is None and s == 0:
                self.first_codes = codes_chunk[:min(e - s, keep_first_codes)].cpu().numpy()
            del x, xs
        torch.cuda.synchronize()

    def encode(self, xs, r, out):
        L = self.L
        if self.workload == "sq8":
            L.call("vg_sq8_encode_dev", xs.data_ptr(), r, self.dim, L.ptr(self.sq.mins, L.f32p), L.ptr(self.sq.maxs, L.f32p),
                   L.ptr(self.sq.scales, L.f32p), out.data_ptr())
        else:
            L.call("vg_int4_encode_dev", xs.data_ptr(), r, self.dim, L.ptr(self.mins, L.f32p), L.ptr(self.diff, L.f32p), out.data_ptr())

    def sub_index(self, codes):
        vg, L = self.vg, self.L
        if self.workload == "sq8":
            ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=self.dim, rows=len(codes), sq8=(self.sq.mins, self.sq.invScales))
        else:
            ix = vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=self.dim, rows=len(codes), int4=(self.mins, self.diff))
        ix.upload(codes=codes)
        return ix

    def oracle_topk(self, codes, q, k):
        from oracle import oracle as o

        if self.workload == "sq8":
            seg = o.FlatOracle(dim=self.dim, metric=0, quant=1, codes=codes, mins=self.sq.mins, inv=self.sq.invScales)
            want, _ = seg.search_batch(q, k, threads=host_threads())
            return want["row"], want["score"]
        out = np.zeros((len(q), k), o.cand_dtype)
        cnt = np.zeros(len(q), np.int64)
        o.lib.vgo_int4_search_batch(o.fp(q), len(q), o.bp(codes), len(codes), self.dim, o.fp(self.mins), o.fp(self.diff), k,
                                    o.fn_addr(o.lib.vgo_int4_l2_batch_a512), host_threads(), out.ctypes.data_as(C.POINTER(o.Cand)),
                                    cnt.ctypes.data_as(o.i64p))
        return out["row"], out["score"]


def scan_parity(env, db, queries_t, rows, k):
    """Parity evidence of a scan config:
       (1) oracle: 64 queries x the first 100k rows (NumPy-seeded prefix) through the tensor-core filter vs the oracle;
       (2) multi-tile: >= 512 queries x ALL local rows, filter vs the repo's exact CUDA-core scan (ids + score bits).
       (N > 1: the caller adds the NCCL-merged result of 512 queries vs ONE index holding all rows on rank 0.)"""
    torch, L = env.torch, env.L
    out = {}
    if env.rank == 0 and db.first_codes is not None:
        try:
            hc = db.first_codes[:100_000]
            qh = queries_t[:64].cpu().numpy()
            with db.sub_index(hc) as pix:
                prow, psc, _ = pix.search(qh, k)
            wr, ws = db.oracle_topk(hc, qh, k)
            out.update(sample=f"{len(qh)} queries x first {len(hc)} rows (NumPy seed-{DATA_SEED} prefix) vs oracle",
                       topk_ids_identical=bool(np.array_equal(prow, wr)),
                       scores_bit_identical=bool(np.array_equal(psc.view(np.uint32), ws.view(np.uint32))))
        except Exception as ex:  # the oracle is a checker, never a dependency of the measurement
            out["error"] = repr(ex)
    nmt = min(512, queries_t.shape[0])
    r1, s1, c1 = env.out_bufs(nmt, k)
    r2, s2, c2 = env.out_bufs(nmt, k)
    st0 = env.qtc_counters()
    db.ix.search_dev(queries_t.data_ptr(), nmt, k, r1.data_ptr(), s1.data_ptr(), c1.data_ptr())
    st1 = env.qtc_counters()
    L.call("vg_flat_tc_enable", 0)
    try:
        db.ix.search_dev(queries_t.data_ptr(), nmt, k, r2.data_ptr(), s2.data_ptr(), c2.data_ptr())
    finally:
        L.call("vg_flat_tc_enable", 1)
    ids_ok, sc_ok = same_results(torch, r1, s1, r2, s2)
    mt_ok = env.all_true(ids_ok and sc_ok and st1[2] - st0[2] == nmt)
    if env.rank == 0:
        out["multi_tile"] = {"sample": f"{nmt} queries ({(nmt + 255) // 256} query tiles) x all {rows} rows: tensor-core filter vs exact CUDA-core scan, "
                                       "every rank's shard", "ids_and_score_bits_identical": mt_ok,
                             "filter_queries": int(st1[2] - st0[2]), "exact_rerun_queries": int(st1[3] - st0[3])}
    return out, (r1, s1)


def run_replicas(env, a, workload, whole, queries, rows, nq, k, steps, warmup, sh):
    """SURVEY 8(e), "a database that fits one GPU (C1, C2)": every GPU holds the WHOLE database and answers 1/W of the query
    batch; one all-gather (8-byte keys) gives every rank the whole batch's result.  Same total work as the row-sharded
    mode, but the per-query stages (selection, exact stage) shrink with the number of GPUs instead of staying fixed."""
    torch, dist, vg, L = env.torch, env.dist, env.vg, env.L
    from vecgo_b200.sharded import shard_range

    world, rank, dev = env.world, env.rank, env.dev
    dim = a.dim
    nql = (nq + world - 1) // world           # queries per rank (the last rank's slice is padded with repeats of its first query)
    qlo = min(rank * nql, nq - 1)
    idx = torch.clamp(torch.arange(qlo, qlo + nql, device=dev), max=nq - 1)
    myq = queries[idx].contiguous()
    main_stream = torch.cuda.current_stream()
    comm_stream = torch.cuda.Stream()
    bufs = [env.out_bufs(nql, k) for _ in range(2)]
    keyb = [torch.empty((nql, k), dtype=torch.int64, device=dev) for _ in range(2)]
    allk = [torch.empty((world * nql, k), dtype=torch.int64, device=dev) for _ in range(2)]
    outs = [env.out_bufs(world * nql, k) for _ in range(2)]
    pending = [None, None]
    # launch-only search (vg_index_search_dev_async): nothing waits for the device inside a step.  The certificate flags of
    # every rank travel with the keys (one more small all-gather), come back to pinned host memory asynchronously and are
    # looked at two steps later, when the buffer is reused; only if ANY rank flagged a query (never on this data) do all
    # ranks settle their flagged queries (vg_index_search_resolve) and repeat that step's exchange.
    flags = [torch.zeros((nql,), dtype=torch.int32, device=dev) for _ in range(2)]
    allflags = [torch.zeros((world * nql,), dtype=torch.int32, device=dev) for _ in range(2)]
    hflags = [torch.zeros((world * nql,), dtype=torch.int32).pin_memory() for _ in range(2)]
    settled = {"steps": 0}

    def exchange(b):
        r_, s_, c_ = bufs[b]
        L.call("vg_topk_pack_dev", r_.data_ptr(), s_.data_ptr(), nql * k, 0, keyb[b].data_ptr())
        dist.all_gather_into_tensor(allk[b], keyb[b])
        orow, osc, ocnt = outs[b]
        # one list per query: the "merge" only unpacks the gathered keys into (rows, scores, counts)
        L.call("vg_topk_merge_keys_dev", allk[b].data_ptr(), 1, world * nql, k, 0, k, orow.data_ptr(), osc.data_ptr(), ocnt.data_ptr())

    def settle(b):
        if pending[b] is None:
            return
        pending[b].synchronize()           # long complete: two steps old
        if bool(hflags[b].any()):          # same answer on every rank: the flags were all-gathered
            settled["steps"] += 1
            r_, s_, c_ = bufs[b]
            whole.ix.search_resolve(myq.data_ptr(), nql, k, r_.data_ptr(), s_.data_ptr(), c_.data_ptr(), flags[b].data_ptr())
            exchange(b)
            torch.cuda.synchronize()
        pending[b] = None

    def one_step(i):
        b = i & 1
        r_, s_, c_ = bufs[b]
        settle(b)
        whole.ix.search_dev_async(myq.data_ptr(), nql, k, r_.data_ptr(), s_.data_ptr(), c_.data_ptr(), flags[b].data_ptr())
        ready = torch.cuda.Event()
        ready.record(main_stream)
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ready)
            L.call("vg_set_stream", comm_stream.cuda_stream)
            try:
                exchange(b)
                dist.all_gather_into_tensor(allflags[b], flags[b])
                hflags[b].copy_(allflags[b], non_blocking=True)
            finally:
                L.call("vg_set_stream", main_stream.cuda_stream)
            done = torch.cuda.Event()
            done.record(comm_stream)
        pending[b] = done
        return outs[b]

    for i in range(warmup):
        one_step(i)
    main_stream.wait_stream(comm_stream)
    env.barrier()
    sampler = ClockSampler(env.local)
    if rank == 0:
        sampler.start()
    qtc0 = env.qtc_counters(1)
    launches0 = vg.launch_count()
    scan_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = None
    for i in range(steps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        res = one_step(i)
        eb.record()
        scan_ms.append((ea, eb))
    for b in range(2):
        settle(b)
    main_stream.wait_stream(comm_stream)
    e1.record()
    env.barrier()
    launches = vg.launch_count() - launches0
    qtc1 = env.qtc_counters(0)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = env.max_over_ranks(e0.elapsed_time(e1)) / steps
    search_ms = env.max_over_ranks(float(np.mean([x.elapsed_time(y) for x, y in scan_ms])))
    gl = int(qtc1[1])
    gemm_ms = env.max_over_ranks(qtc1[0] / gl if gl else 0.0)
    # parity: the gathered result of the whole batch equals the row-sharded, NCCL-merged result
    mr, ms_, mc = sh.search_dev(queries, nq, k)
    i_ok, s_ok = same_results(torch, res[0][:nq], res[1][:nq], mr, ms_)
    same = env.all_true(i_ok and s_ok)
    # e2e: every rank's query slice from ITS pinned host buffer, the whole batch's result back into pinned host memory
    hq = myq.cpu().pin_memory()
    dq = torch.empty_like(myq)
    hout = (torch.empty((world * nql, k), dtype=torch.int32).pin_memory(), torch.empty((world * nql, k), dtype=torch.float32).pin_memory())

    def e2e_step():
        dq.copy_(hq, non_blocking=True)
        r_, s_, c_ = bufs[0]
        whole.ix.search_dev(dq.data_ptr(), nql, k, r_.data_ptr(), s_.data_ptr(), c_.data_ptr())
        L.call("vg_topk_pack_dev", r_.data_ptr(), s_.data_ptr(), nql * k, 0, keyb[0].data_ptr())
        dist.all_gather_into_tensor(allk[0], keyb[0])
        orow, osc, ocnt = outs[0]
        L.call("vg_topk_merge_keys_dev", allk[0].data_ptr(), 1, world * nql, k, 0, k, orow.data_ptr(), osc.data_ptr(), ocnt.data_ptr())
        hout[0].copy_(orow, non_blocking=True)
        hout[1].copy_(osc, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(1, warmup)):
        e2e_step()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    env.barrier()
    e2e_s = env.max_over_ranks((time.perf_counter() - t0) / steps)
    flops = 2.0 * nql * rows * dim * steps / max(gl, 1)
    ach = flops / (gemm_ms / 1e3) / 1e12 if gl else 0.0
    st_i8 = C.c_int32(0)
    L.call("vg_quant_tc_i8_state", C.byref(st_i8))
    i8x = 2.0 if (bool(st_i8.value) and dim % 128 == 0 and (workload == "int4" or dim <= 1024) and (k <= 16 or 6 * k <= 2048)) else 1.0
    return {"value": nq / (ms_per_step / 1e3), "unit": "queries/s", "ms_per_step": ms_per_step, "queries_per_gpu": nql,
            "step_breakdown_ms": {"step": ms_per_step, "per_rank_search": search_ms, "gemm_kernel": gemm_ms, "search_minus_gemm": search_ms - gemm_ms,
                                  "note": "per-rank search of nq / W queries over ALL rows; the pack + all-gather + unpack of step i runs on a second "
                                          "stream under the scan of step i + 1"},
            "identical_to_row_sharded_result": same, "gpu_launches": int(launches), "clocks": clocks,
            "unproven_steps_settled": settled["steps"],
            "roofline": {"bound": "tensor", "achieved": ach, "peak": env.tf_sustained * i8x, "unit": "TOP/s" if i8x > 1 else "TFLOP/s",
                         "frac": ach / (env.tf_sustained * i8x), "frac_of_bf16_sustained_peak": ach / env.tf_sustained,
                         "frac_of_burst_peak": ach / (env.tf_burst * i8x), "kernel_ms": gemm_ms, "kernel_launches_in_timed_region": gl,
                         "share_of_step": gemm_ms * gl / steps / ms_per_step if gl else None, "algorithmic_flops_per_launch": flops,
                         "kernel": f"qtc2_kernel<{workload.upper()}> (CTA pair, tcgen05.mma cta_group::2 kind::f16, M=256 x N=256, fp32 accumulate in TMEM)",
                         "peak_source": env.peak_src, "traffic": None},
            "e2e": {"value": nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(nql * dim * 4) * world, "d2h_bytes_per_step": int(world * nql * k * 8),
                    "steps": steps, "note": "every rank copies ITS query slice from pinned host memory and reads the whole batch's result back into "
                                            "pinned host memory, every step; bytes are the sum over ranks of the inputs and one copy of the result"}}


def run_c2(env, a, workload, rows, nq, k, steps, warmup, headline):
    torch, dist, vg, L = env.torch, env.dist, env.vg, env.L
    from vecgo_b200.sharded import ShardedIndex, exchange_topk, shard_range

    dim = a.dim
    world, rank, dev = env.world, env.rank, env.dev
    code_bytes = dim if workload == "sq8" else dim // 2
    lo, hi = shard_range(rows, rank, world)
    nloc = hi - lo
    t_gen = time.time()
    db = ScanDB(env, workload, rows, dim, lo, hi, keep_first_codes=100_000)
    ix = db.ix
    t_gen = time.time() - t_gen
    queries = torch.from_numpy(np_rows(max(nq, 4096), dim, QUERY_SEED)[:nq]).to(dev)

    # ---- exact float32 ground truth for recall@10 on a few queries (chunk by chunk, never materialised)
    n_gt = min(64, nq)
    gt_rows, gt_scores = [], []
    for c in range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK):
        x = gen_chunk(env, c, CHUNK, dim)
        s, e = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        xs = x[s - c * CHUNK:e - c * CHUNK].contiguous()
        r = e - s
        kk = min(10, r)
        with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=r) as gix:
            gix.upload_dev(r, d_vectors=xs.data_ptr())
            rr, ss, cc = env.out_bufs(n_gt, kk)
            gix.search_dev(queries.data_ptr(), n_gt, kk, rr.data_ptr(), ss.data_ptr(), cc.data_ptr())
            gt_rows.append(rr.cpu().numpy().view(np.uint32).astype(np.int64) + s)
            gt_scores.append(ss.cpu().numpy())
        del x, xs
    gr, gs = np.concatenate(gt_rows, 1), np.concatenate(gt_scores, 1)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (gr, gs))
        gr, gs = np.concatenate([p[0] for p in parts], 1), np.concatenate([p[1] for p in parts], 1)
    order = np.argsort(gs, axis=1, kind="stable")[:, :10]
    gt10 = np.take_along_axis(gr, order, 1)

    sh = ShardedIndex(ix, descending=False)

    # ---- timed region: device-resident inputs.  N > 1: the exchange (pack -> ONE NCCL all-gather of 8-byte keys -> merge) of
    # step i runs on a second stream while the scan of step i+1 runs on the first; every step's merged result is complete
    # before the closing event (the comm stream is joined), so `value` is the steady-state rate of a stream of batches.
    overlap = world > 1 and os.environ.get("BENCH_NO_OVERLAP") != "1"
    main_stream = torch.cuda.current_stream()
    comm_stream = torch.cuda.Stream() if overlap else main_stream
    bufs = [env.out_bufs(nq, k) for _ in range(2)]
    merged = [None, None]

    def one_step(i):
        rows_b, sc_b, cn_b = bufs[i & 1]
        if overlap and merged[i & 1] is not None:
            main_stream.wait_event(merged[i & 1][3])   # the exchange that last read these buffers has finished
        ix.search_dev(queries.data_ptr(), nq, k, rows_b.data_ptr(), sc_b.data_ptr(), cn_b.data_ptr())
        if world == 1:
            return rows_b, sc_b, None
        ready = torch.cuda.Event()
        ready.record(main_stream)
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ready)
            L.call("vg_set_stream", comm_stream.cuda_stream)
            try:
                mr, ms_, mc = sh._exchange_merge(rows_b, sc_b, cn_b, k, k, False)
            finally:
                L.call("vg_set_stream", main_stream.cuda_stream)
            done = torch.cuda.Event()
            done.record(comm_stream)
        merged[i & 1] = (mr, ms_, mc, done)
        return mr, ms_, done

    for i in range(warmup):
        one_step(i)
    main_stream.wait_stream(comm_stream)
    env.barrier()
    sampler = ClockSampler(env.local)
    if rank == 0 and headline:
        sampler.start()
    qtc0 = env.qtc_counters(1)  # CUDA events around every GEMM launch of the filter, on the launching stream
    launches0 = vg.launch_count()
    scan_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    rows_t = sc_t = None
    for i in range(steps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        rows_t, sc_t, _ = one_step(i)
        eb.record()
        scan_ms.append((ea, eb))
    main_stream.wait_stream(comm_stream)
    e1.record()
    env.barrier()
    launches = vg.launch_count() - launches0
    qtc1 = env.qtc_counters(0)
    clocks = sampler.stop() if (rank == 0 and headline) else None
    ms_per_step = env.max_over_ranks(e0.elapsed_time(e1)) / steps
    kernel_ms = env.max_over_ranks(float(np.mean([x.elapsed_time(y) for x, y in scan_ms])))
    gemm_launches = int(qtc1[1])
    gemm_ms = env.max_over_ranks(qtc1[0] / gemm_launches if gemm_launches else 0.0)
    qps = nq / (ms_per_step / 1e3)
    # the exchange on its own (not overlapped), for the breakdown
    exchange_ms = 0.0
    if world > 1:
        rows_b, sc_b, cn_b = bufs[0]
        exchange_ms = env.timed_steps(lambda: sh._exchange_merge(rows_b, sc_b, cn_b, k, k, False), 5, 2)

    # ---- recall@10 of the (approximate) codes vs exact float32 brute force, on n_gt queries
    final_rows = rows_t[:n_gt, :10].cpu().numpy().view(np.uint32).astype(np.int64)
    recall = float(np.mean([len(set(final_rows[i]) & set(gt10[i])) / 10.0 for i in range(n_gt)]))

    # ---- parity
    parity, (mt_rows, mt_scores) = scan_parity(env, db, queries, nloc, k)
    merged_parity = None
    replicas = None
    if world > 1:
        # the NCCL-merged result of 512 queries against ONE index holding all rows (every rank holds one: it is also the
        # replica of the query-sharded mode below)
        nmt = min(512, nq)
        mr, ms_, mc = sh.search_dev(queries[:nmt].contiguous(), nmt, k)
        whole = ScanDB(env, workload, rows, dim, 0, rows)
        wr, ws, wc = env.out_bufs(nmt, k)
        whole.ix.search_dev(queries.data_ptr(), nmt, k, wr.data_ptr(), ws.data_ptr(), wc.data_ptr())
        i_ok, s_ok = same_results(torch, mr, ms_, wr, ws)
        merged_parity = env.all_true(i_ok and s_ok)
        if rank == 0:
            parity["merged"] = {"sample": f"{nmt} queries: {world} shards + NCCL all-gather + device merge vs ONE index holding all {rows} rows (on every rank)",
                                "ids_and_score_bits_identical": merged_parity}
        if headline:
            replicas = run_replicas(env, a, workload, whole, queries, rows, nq, k, steps, warmup, sh)
        whole.ix.close()
        del whole

    # ---- the same sharded search with the exchange INSIDE the C ABI (vg_shard_group_*: what a Go host binds); torch.distributed
    # only ships the 128-byte NCCL id.  Timed un-overlapped (the call returns when the merged result is complete).
    abi_group = None
    if world > 1:
        try:
            from vecgo_b200.sharded import ShardGroup

            grp = ShardGroup.from_torch_distributed(env.local)
            gr, gs_, gc = grp.search_dev(ix, queries, nq, k)
            mr, ms_, mc = sh.search_dev(queries, nq, k)
            i_ok, s_ok = same_results(torch, gr, gs_, mr, ms_)
            ok = env.all_true(i_ok and s_ok)
            for _ in range(2):
                grp.search_dev(ix, queries, nq, k)
            env.barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                grp.search_dev(ix, queries, nq, k)
            env.barrier()
            dt = env.max_over_ranks((time.perf_counter() - t0) / 5)
            grp.close()
            abi_group = {"value": nq / dt, "unit": "queries/s", "ms_per_step": dt * 1e3, "identical_to_torch_distributed_path": ok,
                         "note": "vg_shard_group_search_dev: per-shard scan, ONE ncclAllGather of 8-byte keys and the merge inside libvecgo_cuda "
                                 "(NCCL via dlopen); wall clock around the call, no overlap between batches"}
        except Exception as ex:  # noqa: BLE001
            abi_group = {"error": repr(ex)}

    # ---- e2e: host buffers through the C ABI (vg_index_search), H2D/D2H inside the timed region
    e2e = None
    if headline:
        e2e_steps = max(1, steps)   # same step count and warm-up as the device-timed loop: both run in the sustained (power-capped) regime
        hq_t = queries.cpu().pin_memory()   # the step's inputs come from pinned host memory (the library DMAs page-locked buffers directly)
        hq = hq_t.numpy()
        e2e_out_t = (torch.empty((nq, k), dtype=torch.int32).pin_memory(), torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                     torch.empty((nq,), dtype=torch.int32).pin_memory())
        e2e_out = (e2e_out_t[0].numpy().view(np.uint32), e2e_out_t[1].numpy(), e2e_out_t[2].numpy())
        h2d, d2h = hq.nbytes, nq * k * 8 + nq * 4
        dq = torch.empty((nq, dim), dtype=torch.float32, device=dev)

        def e2e_step():
            if world == 1:
                ix.search(hq, k, out=e2e_out)
            else:
                # every rank gets the whole query batch from ITS pinned host copy and reads the merged result back into
                # pinned host memory; the buffers are allocated once, the copies are inside the timed region
                dq.copy_(hq_t, non_blocking=True)
                r_, s_, c_ = sh.search_dev(dq, nq, k)
                e2e_out_t[0].copy_(r_, non_blocking=True)
                e2e_out_t[1].copy_(s_, non_blocking=True)
                e2e_out_t[2].copy_(c_, non_blocking=True)
                torch.cuda.synchronize()

        for _ in range(max(1, warmup)):  # untimed: first use of the host buffers, clocks back under load after the CPU-side checks
            e2e_step()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        env.barrier()
        e2e_s = env.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "note": "pinned host queries in, pinned host results out, every step; wall clock around the public call (vg_index_search: "
                       "results and certificate flags come back behind one synchronisation).  The device-timed loop above additionally "
                       "records a CUDA-event pair and synchronises after every GEMM launch (roofline.kernel_ms), which costs it ~0.4 ms per "
                       "step that this loop does not pay"}

    res = None
    if rank == 0:
        alg_bytes = float(nq) * nloc * code_bytes  # one (query,row) pair = the row's code bytes
        hbm_equiv = alg_bytes / (kernel_ms / 1e3) / 1e9
        traffic = traffic_i8 = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
                traffic = tj.get(f"qtc:{workload}:{nloc}x{dim}:q{nq}:k{k}")
                traffic_i8 = tj.get(f"qtc_i8:{workload}:{nloc}x{dim}:q{nq}:k{k}")
        if gemm_launches > 0:
            # dominant kernel: the decode-GEMM filter.  Algorithmic FLOPs per (query,row) pair = 2*dim; one launch
            # processes every pair of the batch (a 10k-query batch is one launch; longer batches are chunked).
            flops = 2.0 * nq * nloc * dim * steps / gemm_launches
            ach = flops / (gemm_ms / 1e3) / 1e12
            st_i8 = C.c_int32(0)
            L.call("vg_quant_tc_i8_state", C.byref(st_i8))
            i8 = bool(st_i8.value) and dim % 128 == 0 and (workload == "int4" or dim <= 1024) and (k <= 16 or 6 * k <= 2048)
            # kind::i8 issues at twice the kind::f16 rate (4.5 vs 2.25 P dense nominal); MEASURED_PEAKS.json holds no 8-bit figure,
            # so the denominator is twice the MEASURED sustained bf16 rate — stated in peak_source
            peak = env.tf_sustained * (2.0 if i8 else 1.0)
            roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TOP/s" if i8 else "TFLOP/s", "frac": ach / peak,
                        "traffic": traffic_i8 if i8 else traffic,
                        "kernel": (f"qtc_kernel<{workload.upper()}> (tcgen05.mma cta_group::1 kind::f16, M=128 x N=128)"
                                   if os.environ.get("VECGO_QTC_PAIR", "1")[:1] == "0" else
                                   "qtc2_kernel<SQ8I> (CTA pair, tcgen05.mma cta_group::2 kind::i8, M=256 x N=256: s8 query tile resident in shared "
                                   "memory x u8 code tiles by TMA, int32 accumulate in TMEM, two epilogue groups)" if i8 and workload == "sq8" else
                                   "qtc2_kernel<INT4I> (CTA pair, tcgen05.mma cta_group::2 kind::i8, M=256 x N=256: s8 query k-blocks by TMA x nibbles "
                                   "expanded to u8 by the decode warps, int32 accumulate in TMEM)" if i8 else
                                   f"qtc2_kernel<{workload.upper()}> (CTA pair, tcgen05.mma cta_group::2 kind::f16, M=256 x N=256, fp32 accumulate in TMEM)"),
                        "kernel_ms": gemm_ms, "kernel_launches_in_timed_region": gemm_launches, "share_of_step": gemm_ms * gemm_launches / steps / ms_per_step,
                        "algorithmic_flops_per_launch": flops,
                        "peak_source": (env.peak_src + "; kind::i8: 2 x bf16_tflops_sustained (no measured 8-bit peak; the 8-bit MMA issues at twice "
                                        "the 16-bit rate)") if i8 else env.peak_src,
                        "frac_of_bf16_sustained_peak": ach / env.tf_sustained, "frac_of_burst_peak": ach / (env.tf_burst * (2.0 if i8 else 1.0)),
                        "hbm_equivalent": {"achieved_gbs": hbm_equiv, "peak_gbs": env.hbm_peak, "frac": hbm_equiv / env.hbm_peak,
                                           "note": "queries x rows x code bytes per row / whole-search time: the per-query streaming bytes of "
                                                   "the reference (SURVEY 8d). Above 1 because one decoded code tile serves 256 queries."},
                        "note": "achieved = 2 x queries x rows x dim / GEMM kernel time (CUDA events on the launching stream around every "
                                "launch). " + ("The raw code bytes are the unsigned 8-bit B operand; the query tile is quantised to signed 8-bit per "
                                               "query and its measured quantisation error enters the certificate." if i8 and workload == "sq8" else
                                               "The 4-bit codes are expanded to unsigned bytes inside the kernel; the query tile is quantised to signed 8-bit "
                                               "per query and its measured quantisation error enters the certificate." if i8 else
                                               "Codes are decoded to exact fp16 integers inside the kernel; the binding limit is the tensor pipe, "
                                               "not HBM (DRAM traffic per launch in `traffic`).")}
            dtype = ("i8 tensor-core filter (u8 codes x s8 queries, int32 accumulate), then f32 exact re-check in the reference's AVX-512 order "
                     "(results bit-identical to the f32 scan)") if i8 else (
                     "f16 tensor-core filter over exact integer codes with f32 accumulate, then f32 exact re-check in the reference's "
                     "AVX-512 order (results bit-identical to the f32 scan)")
        else:
            kernel_name = "scan_topk_kernel<CodecSQ8Perm<16>>" if workload == "sq8" else "scan_topk_kernel<CodecINT4Perm>"
            roofline = {"bound": "hbm", "achieved": hbm_equiv, "peak": env.hbm_peak, "unit": "GB/s", "frac": hbm_equiv / env.hbm_peak, "traffic": traffic,
                        "kernel": kernel_name, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": env.peak_src,
                        "binding_limit": "fp32 fma pipe (exact CUDA-core scan; tensor-core filter disabled or shape unsupported)"}
            dtype = "f32 (codes decoded to f32, packed f32x2 FMA in the reference's AVX-512 order)"
        cb = None
        if world == 1 and not a.no_cpu_baseline:
            try:
                cb = cpu_arm(workload, dim, k, a.cpu_seconds if headline else a.config_cpu_seconds, rows, steps=3, warmup=1)
            except Exception as ex:
                cb = {"error": repr(ex)}
        res = {
            "metric": f"batched QPS, {workload.upper()} decode-and-scan top-{k}", "value": qps, "unit": "queries/s",
            "ms_per_step": ms_per_step, "dtype": dtype.split(" ")[0], "arithmetic": dtype,
            "data": f"synthetic N(0,1): rows [0,{SUB}) and all queries from NumPy PCG64 (seeds {DATA_SEED} / {QUERY_SEED}: the prefix the CPU arm and "
                    f"the oracle checks use), remaining rows drawn on the device (torch.randn, per-chunk seeds); quantizer trained on the first "
                    f"{db.train_rows} rows; generation+encode took {t_gen:.0f}s",
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "recall_at_10": recall, "recall_queries": n_gt, "parity": parity, "search_ms": kernel_ms, "scanned_gbs_per_gpu": hbm_equiv,
            "tensor_core_filter": {"queries": int(qtc1[2] - qtc0[2]), "exact_rerun_queries": int(qtc1[3] - qtc0[3])},
        }
        if world > 1:
            res["step_breakdown_ms"] = {"step": ms_per_step, "per_rank_search": kernel_ms, "gemm_kernel": gemm_ms,
                                        "search_minus_gemm": kernel_ms - gemm_ms, "exchange_and_merge_alone": exchange_ms,
                                        "exchange_overlapped_with_next_scan": bool(overlap),
                                        "note": "max over ranks of each part; search = query preparation + GEMM + group selection + exact stage "
                                                "(+ the host read of the certificate flags); exchange = pack to 8-byte keys + ONE NCCL all-gather + "
                                                "device merge, timed alone in a separate loop"}
            res["merged_parity"] = merged_parity
            res["c_abi_shard_group"] = abi_group
            res["replicas"] = replicas
    ix.close()
    del db
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------ C1: Flat exact L2, 100k x 128, 1k queries, k=10
def run_c1(env, a):
    torch, vg, L = env.torch, env.vg, env.L
    from oracle import oracle as o

    n, dim, nq, k = (20_000, 128, 256, 10) if a.small else (100_000, 128, 1000, 10)
    x = np_rows(n, dim, DATA_SEED, "uniform")   # FillUniform, config_test.go:62-91
    q = np_rows(nq, dim, QUERY_SEED, "uniform")
    ix = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n)
    ix.upload(vectors=x)
    dq = torch.from_numpy(q).to(env.dev)
    r, s, c = env.out_bufs(nq, k)
    flags = torch.zeros((nq,), dtype=torch.int32, device=env.dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ix.search_dev_async(dq.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr(), flags.data_ptr())

    f0 = env.flat_counters()
    launches0 = vg.launch_count()
    step()
    torch.cuda.synchronize()
    launches_per_step = vg.launch_count() - launches0
    steps, warm = max(20, a.steps), max(3, a.warmup)
    ms_cold = env.timed_steps(step, steps, warm, flush=True)     # L2 flushed before every step (the database fits in the 126 MB L2)
    ms_hot = env.timed_steps(step, steps, warm, flush=False)     # back to back: database and its fp16 shadow stay L2-resident
    unproven = ix.search_resolve(dq.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr(), flags.data_ptr())
    torch.cuda.synchronize()
    f1 = env.flat_counters()
    # parity: every query against the oracle (reference AVX-512 kernels when oracle/_ref is loaded)
    kern = o.ref_kernels() if o.ref is not None else o.oracle_kernels()
    t0 = time.perf_counter()
    want, _ = o.FlatOracle(dim=dim, metric=0, vectors=x, kernels=kern).search_batch(q, k, threads=host_threads())
    cpu_s = time.perf_counter() - t0
    got_r, got_s = r.cpu().numpy().view(np.uint32), s.cpu().numpy()
    parity = {"sample": f"all {nq} queries x all {n} rows vs oracle ({'reference AVX-512 squaredL2 (oracle/_ref)' if o.ref is not None else 'oracle port'} "
                        "+ flat.Search loop / CandidateHeap restatement)",
              "topk_ids_identical": bool(np.array_equal(got_r, want["row"])),
              "scores_bit_identical": bool(np.array_equal(got_s.view(np.uint32), want["score"].view(np.uint32)))}
    # e2e through vg_index_search with pinned host buffers
    hq_t = torch.from_numpy(q).pin_memory()
    out_t = (torch.empty((nq, k), dtype=torch.int32).pin_memory(), torch.empty((nq, k), dtype=torch.float32).pin_memory(),
             torch.empty((nq,), dtype=torch.int32).pin_memory())
    out = (out_t[0].numpy().view(np.uint32), out_t[1].numpy(), out_t[2].numpy())
    for _ in range(warm):
        ix.search(hq_t.numpy(), k, out=out)
    t0 = time.perf_counter()
    for _ in range(steps):
        ix.search(hq_t.numpy(), k, out=out)
    e2e_s = (time.perf_counter() - t0) / steps
    flops = 2.0 * nq * n * dim
    ach = flops / (ms_cold / 1e3) / 1e12
    res = {"workload": f"Flat exact L2, {n} x {dim}-d float32 U[0,1) (NumPy seed {DATA_SEED}), {nq} queries (seed {QUERY_SEED}), k={k}" + (" [--small]" if a.small else ""),
           "metric": "batched QPS, Flat exact L2 top-10", "value": nq / (ms_cold / 1e3), "unit": "queries/s", "ms_per_step": ms_cold,
           "ms_per_step_l2_resident": ms_hot, "qps_l2_resident": nq / (ms_hot / 1e3), "steps": steps,
           "l2_flush": "256 MB written before every timed step (database 51 MB + fp16 shadow 26 MB fit in the 126 MB L2); "
                       "ms_per_step_l2_resident is the same loop back to back without the flush",
           "dtype": "f16", "arithmetic": "f16 tensor-core filter (fp16 shadow of the rows, f32 accumulate) + exact float32 re-check in simd.SquaredL2 order",
           "roofline": {"bound": "tensor", "achieved": ach, "peak": env.tf_burst, "unit": "TFLOP/s", "frac": ach / env.tf_burst,
                        "kernel_ms": ms_cold, "launches_per_step": int(launches_per_step), "frac_l2_resident": flops / (ms_hot / 1e3) / 1e12 / env.tf_burst,
                        "peak_source": env.peak_src + " — burst: the step is a few tens of microseconds",
                        "note": "2 x queries x rows x dim / WHOLE step time (every launch of the search, not one kernel): the shape is launch-bound; "
                                "target 0.50 = 31.7 us (SURVEY 8d)", "traffic": None},
           "host_synchronisations_per_step": 0,
           "parity": parity, "certificate": {"filter_queries_in_timed_loops": int(f1[0] - f0[0]), "unproven_at_end": int(unproven)},
           "cpu_baseline": {"value": nq / cpu_s, "unit": "queries/s", "cores": host_threads(), "kind": "reference" if o.ref is not None else "port",
                            "sample": f"the whole config: {nq} queries x {n} rows in {cpu_s:.2f}s, one query per worker thread"},
           "e2e": {"value": nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(q.nbytes), "d2h_bytes_per_step": nq * k * 8 + nq * 4}}
    ix.close()
    return res


# ------------------------------------------------------------------ C3: PQ M=96 x 256 ADC scan, 25M rows per GPU
def run_c3(env, a):
    torch, vg, L = env.torch, env.vg, env.L
    from oracle import oracle as o

    n, dim, m, nq, k = (2_000_000, 768, 96, 1024, 100) if a.small else (25_000_000, 768, 96, 10_000, 100)
    world, rank, dev = env.world, env.rank, env.dev
    rng = np.random.default_rng(DATA_SEED)
    ds = dim // m
    cb = rng.integers(-128, 128, m * 256 * ds, dtype=np.int8)         # random int8 codebook / scale / offset
    sc = (0.01 + 0.002 * rng.random(m)).astype(F)
    of = (0.05 * rng.standard_normal(m)).astype(F)
    pq = (cb, sc, of, m, 256)
    sub = min(100_000, n)
    sub_codes = np.random.default_rng(DATA_SEED + 1).integers(0, 256, (sub, m), dtype=np.uint8)   # rows [0, sub) of shard 0: host copy for the oracle

    def build(row_base, shard):
        ix_ = vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=n, row_base=row_base, pq=pq)
        chunk = 1 << 22
        for r0 in range(0, n, chunk):
            mm = min(chunk, n - r0)
            g = torch.Generator(device=dev).manual_seed(DATA_SEED * 7919 + shard * 1009 + r0 // chunk)
            codes = torch.randint(0, 256, (mm, m), dtype=torch.uint8, device=dev, generator=g)   # uniform random bytes
            if shard == 0 and r0 == 0:
                codes[:sub] = torch.from_numpy(sub_codes).to(dev)
            ix_.upload_dev(mm, d_codes=codes.data_ptr(), row0=r0)
        return ix_

    ix = build(rank * n, rank)
    q_h = np_rows(max(nq, 4096), dim, QUERY_SEED)[:nq]
    queries = torch.from_numpy(q_h).to(dev)
    sh = vg.sharded.ShardedIndex(ix, descending=False)
    steps, warm = (3, 1)
    for _ in range(warm):
        sh.search_dev(queries, nq, k)
    env.barrier()
    qtc0 = env.qtc_counters(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = sh.search_dev(queries, nq, k)
    e1.record()
    env.barrier()
    qtc1 = env.qtc_counters(0)
    ms = env.max_over_ranks(e0.elapsed_time(e1)) / steps
    gl = int(qtc1[1])
    gemm_ms = env.max_over_ranks(qtc1[0] / gl if gl else 0.0)
    # parity (1): oracle on the host-held prefix
    parity = {}
    if rank == 0:
        with vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=sub, pq=pq) as pix:
            pix.upload(codes=sub_codes)
            prow, psc, _ = pix.search(q_h[:32], k)
        want, _ = o.FlatOracle(dim=dim, metric=0, quant=2, codes=sub_codes, pq=pq).search_batch(q_h[:32], k, threads=host_threads())
        parity = {"sample": f"32 queries x the first {sub} rows of shard 0 (host-held codes) through the tensor-core filter vs oracle "
                            "(generic BuildDistanceTableInt8 + pqAdcLookupAvx512 order)",
                  "topk_ids_identical": bool(np.array_equal(prow, want["row"])),
                  "scores_bit_identical": bool(np.array_equal(psc.view(np.uint32), want["score"].view(np.uint32)))}
    # parity (2): multi-tile filter vs exact scan on the whole shard
    nmt = min(512, nq)
    r1, s1, c1 = env.out_bufs(nmt, k)
    r2, s2, c2 = env.out_bufs(nmt, k)
    ix.search_dev(queries.data_ptr(), nmt, k, r1.data_ptr(), s1.data_ptr(), c1.data_ptr())
    L.call("vg_flat_tc_enable", 0)
    try:
        nex = min(64, nmt)   # the exact ADC scan runs at ~1k queries/s on 25M rows
        ix.search_dev(queries.data_ptr(), nex, k, r2.data_ptr(), s2.data_ptr(), c2.data_ptr())
    finally:
        L.call("vg_flat_tc_enable", 1)
    i_ok, s_ok = same_results(torch, r1[:nex], s1[:nex], r2[:nex], s2[:nex])
    mt_ok = env.all_true(i_ok and s_ok)
    merged_parity = None
    if world > 1:
        mr, ms_, _ = sh.search_dev(queries[:nmt].contiguous(), nmt, k)
        ok = True
        if rank == 0:
            # ONE index holding every shard's rows on rank 0 (world x 2.4 GB of codes)
            whole = vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=n * world, pq=pq)
            chunk = 1 << 22
            for shard in range(world):
                for r0 in range(0, n, chunk):
                    mm = min(chunk, n - r0)
                    g = torch.Generator(device=dev).manual_seed(DATA_SEED * 7919 + shard * 1009 + r0 // chunk)
                    codes = torch.randint(0, 256, (mm, m), dtype=torch.uint8, device=dev, generator=g)
                    if shard == 0 and r0 == 0:
                        codes[:sub] = torch.from_numpy(sub_codes).to(dev)
                    whole.upload_dev(mm, d_codes=codes.data_ptr(), row0=shard * n + r0)
            wr, ws, wc = env.out_bufs(nmt, k)
            whole.search_dev(queries.data_ptr(), nmt, k, wr.data_ptr(), ws.data_ptr(), wc.data_ptr())
            i2, s2_ = same_results(torch, mr, ms_, wr, ws)
            ok = i2 and s2_
            whole.close()
        merged_parity = env.all_true(ok)
    res = None
    if rank == 0:
        parity["multi_tile"] = {"sample": f"{nex} of {nmt} filtered queries x all {n} rows of every shard vs the exact CUDA-core ADC scan",
                                "ids_and_score_bits_identical": mt_ok}
        if world > 1:
            parity["merged"] = {"sample": f"{nmt} queries: {world} shards + NCCL all-gather + device merge vs ONE index holding all {n * world} rows on rank 0",
                                "ids_and_score_bits_identical": merged_parity}
        flops = 2.0 * nq * n * dim * steps / max(gl, 1)
        ach = flops / (gemm_ms / 1e3) / 1e12 if gl else 0.0
        hbm_equiv = float(nq) * n * world * m / (ms / 1e3) / 1e9
        cbase = None
        if world == 1 and not a.no_cpu_baseline:
            kern = o.ref_kernels() if o.ref is not None else o.oracle_kernels()
            seg = o.FlatOracle(dim=dim, metric=0, quant=2, codes=sub_codes, pq=pq, kernels=kern)
            th = host_threads()
            t0 = time.perf_counter()
            seg.search_batch(q_h[:th], k, threads=th)
            t1 = time.perf_counter() - t0
            nq2 = max(th, int(min(2048, th * a.config_cpu_seconds / max(t1, 1e-3))) // th * th)
            t0 = time.perf_counter()
            seg.search_batch(q_h[:nq2], k, threads=th)
            t2 = time.perf_counter() - t0
            cbase = {"value": nq2 / t2 * sub / (n * world), "unit": "queries/s", "cores": th, "kind": "reference" if o.ref is not None else "port",
                     "sample": f"{nq2} queries x {sub} rows (the host-held prefix) in {t2:.1f}s on {th} threads; linear extrapolation x{sub}/{n * world} rows"}
        st_i8 = C.c_int32(0)
        L.call("vg_quant_tc_i8_state", C.byref(st_i8))
        c3_i8 = bool(st_i8.value) and dim % 128 == 0 and dim // m == 8 and os.environ.get("VECGO_QTC_PQ_I8", "1")[:1] != "0" and 6 * k <= 2048
        res = {"workload": f"PQ M={m} x 256 ADC scan, {world} x {n} rows of {dim}-d codes (uniform random bytes, random int8 codebooks), {nq} queries, k={k}"
                           + (f"; {world} GPUs: NCCL all-gather of the per-shard top-k + device merge" if world > 1 else "")
                           + (" = BASELINE configs[2]" if world == 8 and not a.small else "") + (" [--small]" if a.small else ""),
               "metric": "batched QPS, PQ ADC scan top-100", "value": nq / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms, "steps": steps, "scaling": "weak",
               "dtype": "i8" if c3_i8 else "f16", "rows_total": n * world,
               "roofline": {"bound": "tensor", "achieved": ach, "peak": env.tf_sustained * (2.0 if c3_i8 else 1.0), "unit": "TOP/s" if c3_i8 else "TFLOP/s",
                            "frac": ach / (env.tf_sustained * (2.0 if c3_i8 else 1.0)), "frac_of_bf16_sustained_peak": ach / env.tf_sustained,
                            "frac_of_burst_peak": ach / (env.tf_burst * (2.0 if c3_i8 else 1.0)),
                            "kernel": ("qtc2_kernel<PQI> (kind::i8: int8 centroids gathered from a 32 KB shared-memory slice as signed bytes, 128 dims per "
                                       "k-block; peak = 2 x the measured sustained bf16 rate)" if c3_i8 else "qtc2_kernel<PQ>"),
                            "kernel_ms": gemm_ms, "kernel_launches_in_timed_region": gl,
                            "share_of_step": gemm_ms * gl / steps / ms if gl else None, "traffic": None, "peak_source": env.peak_src,
                            "hbm_equivalent": {"achieved_gbs": hbm_equiv, "peak_gbs": env.hbm_peak * world, "frac": hbm_equiv / (env.hbm_peak * world),
                                               "note": "queries x rows x 96 code bytes / step time (SURVEY 8d byte view) against the summed HBM peaks"}},
               "parity": parity, "merged_parity": merged_parity,
               "tensor_core_filter": {"queries": int(qtc1[2] - qtc0[2]), "exact_rerun_queries": int(qtc1[3] - qtc0[3])},
               "cpu_baseline": cbase}
    ix.close()
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------ C4: RaBitQ 1-bit scan + float32 rerank of the top-1000
def run_c4(env, a):
    torch, vg, L = env.torch, env.vg, env.L
    from oracle import oracle as o

    n, dim, nq, r_top, k = (1_000_000, 1536, 256, 1000, 100) if a.small else (a.c4_rows, 1536, 1000, 1000, 100)
    world, rank, dev = env.world, env.rank, env.dev
    code_bytes = dim // 8 + 4
    sub = min(50_000, n)
    x_sub = np_rows(sub, dim, DATA_SEED)   # rows [0, sub) of shard 0 (NumPy seed 42)
    ix = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n, row_base=rank * n)
    chunk = 1 << 18
    codes = torch.empty((chunk, code_bytes), dtype=torch.uint8, device=dev)
    sub_codes = None
    t0 = time.time()
    for r0 in range(0, n, chunk):
        mm = min(chunk, n - r0)
        g = torch.Generator(device=dev).manual_seed(DATA_SEED * 104729 + rank * 1009 + r0 // chunk)
        x = torch.randn((mm, dim), dtype=torch.float32, device=dev, generator=g)
        if rank == 0 and r0 == 0:
            x[:sub] = torch.from_numpy(x_sub).to(dev)
        L.call("vg_rabitq_encode_dev", x.data_ptr(), mm, dim, codes.data_ptr())
        ix.upload_dev(mm, d_codes=codes.data_ptr(), d_vectors=x.data_ptr(), row0=r0)
        if rank == 0 and r0 == 0:
            sub_codes = codes[:sub].cpu().numpy()
        del x
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    q_h = np_rows(max(nq, 1024), dim, QUERY_SEED)[:nq]
    queries = torch.from_numpy(q_h).to(dev)
    sh = vg.sharded.ShardedIndex(ix, descending=False)
    steps, warm = 5, 2
    rr, ss, cc = env.out_bufs(nq, r_top)
    ms_scan = env.timed_steps(lambda: ix.search_dev(queries.data_ptr(), nq, r_top, rr.data_ptr(), ss.data_ptr(), cc.data_ptr()), steps, warm)
    qtc0 = env.qtc_counters(1)
    ms = env.timed_steps(lambda: sh.search_rerank_dev(queries, nq, r_top, k), steps, 1)
    qtc1 = env.qtc_counters(0)
    gl = int(qtc1[1])
    gemm_ms = env.max_over_ranks(qtc1[0] / gl if gl else 0.0)
    parity = {}
    if rank == 0:
        # oracle: RaBitQ top-R over the prefix rows, exact float32 rerank, final top-k by (score, row)
        nqo, ro = 16, min(r_top, 1000)
        with vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=sub) as pix:
            pix.upload(codes=sub_codes, vectors=x_sub)
            arow, asc, _ = pix.search(q_h[:nqo], ro)
            frow, fsc, fcnt = pix.search_rerank(q_h[:nqo], ro, k)
        ids_ok = sc_ok = fin_ok = True
        for i in range(nqo):
            outc = np.zeros(ro, o.cand_dtype)
            cnt = o.lib.vgo_rabitq_search(o.fp(q_h[i]), o.bp(sub_codes), sub, dim, ro, None, outc.ctypes.data_as(C.POINTER(o.Cand)), None)
            ids_ok &= bool(np.array_equal(arow[i, :cnt], outc["row"][:cnt]))
            sc_ok &= bool(np.array_equal(asc[i, :cnt].view(np.uint32), outc["score"][:cnt].view(np.uint32)))
            ex = np.array([o.lib.vgo_sql2_a512(o.fp(q_h[i]), o.fp(x_sub[rw]), dim) for rw in outc["row"][:cnt]], F)
            order = np.lexsort((outc["row"][:cnt], ex))[:k]
            fin_ok &= bool(np.array_equal(frow[i, :fcnt[i]], outc["row"][:cnt][order]) and
                           np.array_equal(fsc[i, :fcnt[i]].view(np.uint32), ex[order].view(np.uint32)))
        parity = {"sample": f"{nqo} queries x the first {sub} rows of shard 0 (NumPy seed-{DATA_SEED} vectors): top-{ro} estimator scan, exact rerank, final top-{k} vs oracle",
                  "topk_ids_identical": ids_ok, "scores_bit_identical": sc_ok, "reranked_topk_identical": fin_ok}
    # multi-tile: filter vs exact popcount scan on the whole shard
    nmt = min(512, nq)
    nex = min(32, nmt)
    r1, s1, c1 = env.out_bufs(nmt, r_top)
    r2, s2, c2 = env.out_bufs(nex, r_top)
    ix.search_dev(queries.data_ptr(), nmt, r_top, r1.data_ptr(), s1.data_ptr(), c1.data_ptr())
    L.call("vg_flat_tc_enable", 0)
    try:
        ix.search_dev(queries.data_ptr(), nex, r_top, r2.data_ptr(), s2.data_ptr(), c2.data_ptr())
    finally:
        L.call("vg_flat_tc_enable", 1)
    i_ok, s_ok = same_results(torch, r1[:nex], s1[:nex], r2, s2)
    mt_ok = env.all_true(i_ok and s_ok)
    merged_parity = None
    merged_fits = n * (dim * 4 + code_bytes) + n * world * code_bytes < 170e9   # the shard's rows + codes and ALL codes on rank 0
    if world > 1 and merged_fits:
        # scan-only merged parity against ONE index holding every shard's codes on rank 0 (the float32 rows do not fit one GPU)
        mr, ms_, _ = sh.search_dev(queries[:nmt].contiguous(), nmt, r_top)
        ok = True
        if rank == 0:
            whole = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n * world)
            for shard in range(world):
                for r0 in range(0, n, chunk):
                    mm = min(chunk, n - r0)
                    g = torch.Generator(device=dev).manual_seed(DATA_SEED * 104729 + shard * 1009 + r0 // chunk)
                    x = torch.randn((mm, dim), dtype=torch.float32, device=dev, generator=g)
                    if shard == 0 and r0 == 0:
                        x[:sub] = torch.from_numpy(x_sub).to(dev)
                    L.call("vg_rabitq_encode_dev", x.data_ptr(), mm, dim, codes.data_ptr())
                    whole.upload_dev(mm, d_codes=codes.data_ptr(), row0=shard * n + r0)
                    del x
            wr, ws, wc = env.out_bufs(nmt, r_top)
            whole.search_dev(queries.data_ptr(), nmt, r_top, wr.data_ptr(), ws.data_ptr(), wc.data_ptr())
            i2, s2_ = same_results(torch, mr, ms_, wr, ws)
            ok = i2 and s2_
            whole.close()
        merged_parity = env.all_true(ok)
    res = None
    if rank == 0:
        parity["multi_tile"] = {"sample": f"{nex} of {nmt} filtered queries x all {n} rows of every shard, top-{r_top}: tensor-core filter vs exact popcount scan",
                                "ids_and_score_bits_identical": mt_ok}
        if world > 1 and merged_fits:
            parity["merged"] = {"sample": f"{nmt} queries, approximate top-{r_top}: {world} shards + NCCL all-gather + device merge vs ONE index holding "
                                          f"all {n * world} codes on rank 0", "ids_and_score_bits_identical": merged_parity}
        elif world > 1:
            parity["merged"] = {"skipped": f"{n * (dim * 4 + code_bytes) / 1e9:.1f} GB of shard + {n * world * code_bytes / 1e9:.1f} GB of all codes do not fit "
                                           "one GPU; the same exchange is checked at 12.5M rows per GPU"}
        flops = 2.0 * nq * n * dim * steps / max(gl, 1)
        ach = flops / (gemm_ms / 1e3) / 1e12 if gl else 0.0
        st_i8 = C.c_int32(0)
        L.call("vg_quant_tc_i8_state", C.byref(st_i8))
        c4_i8 = bool(st_i8.value) and dim % 128 == 0
        hbm_equiv = float(nq) * n * world * code_bytes / (ms / 1e3) / 1e9
        cbase = None
        if world == 1 and not a.no_cpu_baseline:
            th = host_threads()
            fn = o.fn_addr(o.ref.hammingAvx512) if o.ref is not None else None
            def cpu_pass(nq_):
                outc = np.zeros((nq_, r_top), o.cand_dtype)
                cnt = np.zeros(nq_, np.int64)
                t0 = time.perf_counter()
                o.lib.vgo_rabitq_search_batch(o.fp(q_h[:nq_]), nq_, o.bp(sub_codes), sub, dim, r_top, fn, th, outc.ctypes.data_as(C.POINTER(o.Cand)),
                                              cnt.ctypes.data_as(o.i64p))
                return time.perf_counter() - t0

            t1 = cpu_pass(th)
            nq2 = max(th, int(min(nq, th * a.config_cpu_seconds / max(t1, 1e-3))) // th * th)
            t2 = cpu_pass(nq2)
            cbase = {"value": nq2 / t2 * sub / (n * world), "unit": "queries/s", "cores": th, "kind": "reference" if o.ref is not None else "port",
                     "sample": f"{nq2} queries x {sub} rows, estimator scan to top-{r_top} (rerank excluded) in {t2:.1f}s on {th} threads; linear extrapolation "
                               f"x{sub}/{n * world} rows"}
        res = {"workload": f"RaBitQ 1-bit scan + float32 rerank of the global top-{r_top}, {world} x {n} x {dim}-d (N(0,1)), {nq} queries, final k={k}"
                           + (f"; {world} GPUs: two NCCL exchanges (global approximate top-R, then exact scores of the owned rows)" if world > 1 else "")
                           + (" = BASELINE configs[3]" if world == 8 and not a.small else "") + (" [--small]" if a.small else ""),
               "metric": "batched QPS, RaBitQ scan + rerank", "value": nq / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms, "steps": steps, "scaling": "weak",
               "scan_only_ms": ms_scan, "rerank_exchange_merge_ms": ms - ms_scan, "rows_total": n * world, "dtype": "f16",
               "generate_encode_upload_s": gen_s,
               "roofline": {"bound": "tensor", "achieved": ach, "peak": env.tf_burst * (2.0 if c4_i8 else 1.0), "unit": "TOP/s" if c4_i8 else "TFLOP/s",
                            "frac": ach / (env.tf_burst * (2.0 if c4_i8 else 1.0)), "frac_of_bf16_burst_peak": ach / env.tf_burst,
                            "kernel": ("qtc2_kernel<RABITQI> (kind::i8: sign bits as exact +-1 signed bytes, 128 dims per k-block; peak = 2 x the measured "
                                       "bf16 burst rate, MEASURED_PEAKS.json holds no 8-bit figure)" if c4_i8 else
                                       "qtc2_kernel<RABITQ> (sign bits as exact +-1 fp16 operands)"), "kernel_ms": gemm_ms, "kernel_launches_in_timed_region": gl,
                            "share_of_step": gemm_ms * gl / steps / ms if gl else None, "traffic": None, "peak_source": env.peak_src,
                            "hbm_equivalent": {"achieved_gbs": hbm_equiv, "peak_gbs": env.hbm_peak * world, "frac": hbm_equiv / (env.hbm_peak * world),
                                               "note": "queries x rows x 196 code bytes / step time (SURVEY 8d byte view) against the summed HBM peaks"}},
               "parity": parity, "merged_parity": merged_parity,
               "tensor_core_filter": {"queries": int(qtc1[2] - qtc0[2]), "exact_rerun_queries": int(qtc1[3] - qtc0[3])},
               "cpu_baseline": cbase}
    ix.close()
    del codes
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------ C5: PQ codebook training
def oracle_pq_train(v, m, k, iters, seed, threads):
    """ProductQuantizer.Train restatement: one worker per subspace (pq.go:79-140 runs a goroutine per subspace)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as o

    n, dim = v.shape
    ds = dim // m
    cents = np.zeros((m, k, ds), F)

    def one(s):
        cent = np.zeros((k, ds), F)
        o.lib.vgo_pq_kmeanspp_init(o.fp(v), n, dim, s * ds, ds, k, seed, s, o.fp(cent))
        assign = np.zeros(n, np.int32)
        o.lib.vgo_pq_lloyd(o.fp(v), n, dim, s * ds, ds, k, iters, seed, s, o.fp(cent), assign.ctypes.data_as(o.i32p))
        cents[s] = cent

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(m)))
    return cents


def run_c5(env, a):
    torch, vg, L = env.torch, env.vg, env.L
    n, dim, m, iters = (100_000, 768, 96, 25) if a.small else (1_000_000, 768, 96, 25)
    world, rank, dev = env.world, env.rank, env.dev
    ds = dim // m
    g = torch.Generator(device=dev).manual_seed(DATA_SEED)
    dx = torch.randn((n, dim), dtype=torch.float32, device=dev, generator=g)     # same on every rank
    sub = min(16_384, n)
    x_sub = np_rows(sub, dim, DATA_SEED)
    dx[:sub] = torch.from_numpy(x_sub).to(dev)
    cb, sc, of = np.zeros(m * 256 * ds, np.int8), np.zeros(m, F), np.zeros(m, F)
    cent = np.zeros((m, 256, ds), F)

    def pqa_stats():
        p, q_ = C.c_uint64(), C.c_uint64()
        L.call("vg_pq_assign_tc_stats", C.byref(p), C.byref(q_))
        return p.value, q_.value

    def train():
        if world == 1:
            L.call("vg_pq_train_dev", dx.data_ptr(), n, dim, m, 256, iters, 1, L.ptr(cb, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p),
                   L.ptr(cent, L.f32p))
        else:
            vg.sharded.pq_train_sharded(dx, n, dim, m, 256, iters, 1, cb, sc, of, cent)

    train()  # warm-up (first use of the scratch pool and the fp16 shadow)
    env.barrier()
    ps0 = pqa_stats()
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        train()
    env.barrier()
    sec = env.max_over_ranks((time.perf_counter() - t0) / reps)
    ps1 = pqa_stats()
    res = None
    # multi-GPU result must equal the single-GPU result bit for bit (subspaces are independent: pq.go:79-140)
    merged_parity = None
    if world > 1:
        cb1, sc1, of1 = np.zeros_like(cb), np.zeros_like(sc), np.zeros_like(of)
        ok = True
        if rank == 0:
            L.call("vg_pq_train_dev", dx.data_ptr(), n, dim, m, 256, iters, 1, L.ptr(cb1, L.i8p), L.ptr(sc1, L.f32p), L.ptr(of1, L.f32p), None)
            ok = bool(np.array_equal(cb, cb1) and np.array_equal(sc.view(np.uint32), sc1.view(np.uint32)) and
                      np.array_equal(of.view(np.uint32), of1.view(np.uint32)))
        merged_parity = env.all_true(ok)
    if rank == 0:
        # parity: the whole training on the NumPy-seeded sample against the oracle (float32 centroids bit for bit)
        pit = 5
        pq_ = vg.quantization.ProductQuantizer(dim, m, 256)
        pq_.Train(x_sub, iters=pit, seed=7)
        t0 = time.perf_counter()
        want = oracle_pq_train(x_sub, m, 256, pit, 7, host_threads())
        cpu_s = time.perf_counter() - t0
        parity = {"sample": f"ProductQuantizer.Train on the first {sub} samples (NumPy seed {DATA_SEED}), k-means++ + {pit} Lloyd iterations, all {m} subspaces vs oracle",
                  "centroids_bit_identical": bool(np.array_equal(pq_.centroids_f32.view(np.uint32), want.view(np.uint32)))}
        if world > 1:
            parity["merged"] = {"sample": f"codebooks / scales / offsets of the {world}-GPU training (subspaces split across the ranks, all-gathered) vs the "
                                          "single-GPU training on rank 0", "bit_identical": merged_parity}
        flops, byts = 2.0 * n * 256 * dim * iters, 4.0 * n * dim * iters   # the Lloyd assignment contraction, one pass over the samples per iteration
        # CPU: k-means++ cost ~ one assignment pass (n x 256 x ds) per pick... measured on the sample: (init + pit iterations); scaled per sample-iteration
        cpu_rate = sub * (pit + 1) / cpu_s
        res = {"workload": f"PQ codebook training (k-means++ init + {iters} Lloyd iterations), {n} x {dim}-d N(0,1), {m} subspaces x 256 centroids"
                           + (f"; subspaces split across {world} GPUs, codebooks all-gathered" if world > 1 else "") + (" [--small]" if a.small else ""),
               "metric": "PQ training throughput", "value": n * iters / sec, "unit": "sample-iterations/s", "seconds": sec, "ms_per_step": sec * 1e3,
               "steps": reps, "scaling": "strong", "dtype": "f32",
               "arithmetic": "f16 hi/lo-split tensor-core assignment + gap certificate + exact float32 re-evaluation; sample-order float32 sums; exact parallel k-means++ prefix",
               "roofline": {"bound": "hbm", "achieved": byts / sec / 1e9, "peak": env.hbm_peak * world, "unit": "GB/s", "frac": byts / sec / 1e9 / (env.hbm_peak * world),
                            "kernel_ms": sec * 1e3, "traffic": None, "peak_source": env.peak_src,
                            "tensor": {"achieved_tflops": flops / sec / 1e12, "peak": env.tf_burst * world, "frac": flops / sec / 1e12 / (env.tf_burst * world)},
                            "note": "SURVEY 8(d): algorithmic bytes = one pass over the 4-byte samples per Lloyd iteration, over the WHOLE training time "
                                    "(k-means++ initialisation included): 11.7 ms at the HBM peak"},
               "tensor_core_assignment": {"pairs": ps1[0] - ps0[0], "exact_reevaluated_pairs": ps1[1] - ps0[1]},
               "parity": parity, "merged_parity": merged_parity,
               "cpu_baseline": {"value": cpu_rate, "unit": "sample-iterations/s", "cores": host_threads(), "kind": "port",
                                "sample": f"oracle ProductQuantizer.Train restatement on {sub} samples, k-means++ + {pit} iterations, one subspace per worker "
                                          f"thread ({cpu_s:.1f}s); the k-means++ pass is counted as one iteration"}}
    del dx
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------ GPU arm: driver
def run_ours(a):
    env = Env(a)
    t_start = time.time()
    head = run_c2(env, a, a.workload, a.rows, a.queries, a.k, a.steps, a.warmup, headline=True)
    want = ["c1", "c2b", "c3", "c4", "c5"] if a.configs == "all" else ([] if a.configs == "none" else a.configs.split(","))
    configs = {}
    timing = {"headline_s": time.time() - t_start}
    for name in want:
        t0 = time.time()
        try:
            if name == "c1":
                if env.world > 1:
                    continue  # a 51 MB database: replicas only (one GPU answers a batch)
                res = run_c1(env, a)
            elif name == "c2b":
                if env.world > 1:
                    continue  # same path as the headline with 4-bit codes; sharded scaling is measured by the headline
                res = run_c2(env, a, "int4", a.rows if not a.small else 1_000_000, a.queries if not a.small else 2048, a.k, 5, 2, headline=False)
            elif name == "c3":
                res = run_c3(env, a)
            elif name == "c4":
                res = run_c4(env, a)
            elif name == "c5":
                if env.world > 1 and not hasattr(env.vg.sharded, "pq_train_sharded"):
                    continue
                res = run_c5(env, a)
            else:
                continue
        except Exception as ex:  # a failing side config must not take the headline line down
            import traceback

            res = {"error": repr(ex), "trace": traceback.format_exc()[-600:]}
            env.torch.cuda.empty_cache()
        if env.rank == 0 and res is not None:
            res["wall_s"] = time.time() - t0
            configs[{"c1": "C1", "c2b": "C2b", "c3": "C3", "c4": "C4", "c5": "C5"}[name]] = res
    if env.rank == 0:
        cfg = workload_config(a)
        rep = head.get("replicas") if env.world > 1 else None
        if rep and not rep.get("error") and os.environ.get("BENCH_ROW_SHARDED_VALUE") != "1":
            # N > 1: the headline database (7.68 GB of codes) fits one GPU, so the multi-GPU mode SURVEY 8(e) names for it is
            # replicas + a query-sharded batch; the row-sharded run (what configs[2] / [3] need, C3 / C4 below) is kept beside it
            row_sharded = {key: head[key] for key in ("value", "ms_per_step", "roofline", "e2e", "gpu_launches", "clocks", "step_breakdown_ms",
                                                     "merged_parity", "c_abi_shard_group") if key in head}
            row_sharded["sharding"] = cfg["sharding"]
            cfg["sharding"] = (f"replicas: every GPU holds all {a.rows} rows and answers {rep['queries_per_gpu']} of the {a.queries} queries; ONE NCCL "
                               "all-gather of 8-byte keys gives every rank the whole result (SURVEY 8e: a database that fits one GPU); the "
                               "row-sharded run of the same batch is under `row_sharded`")
            head = dict(head)
            for key in ("value", "ms_per_step", "roofline", "e2e", "gpu_launches", "clocks"):
                head[key] = rep[key]
            head["step_breakdown_ms"] = rep["step_breakdown_ms"]
            head["row_sharded"] = row_sharded
            head["replicas_identical_to_row_sharded"] = rep["identical_to_row_sharded_result"]
        line = {
            "metric": head["metric"], "value": head["value"], "unit": "queries/s",
            "n_gpus": env.world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": head["dtype"], "arithmetic": head["arithmetic"],
            "data": head["data"], "config": cfg,
            "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "recall_at_10": head["recall_at_10"],
            "recall_queries": head["recall_queries"], "parity": head["parity"], "search_ms": head["search_ms"],
            "scanned_gbs_per_gpu": head["scanned_gbs_per_gpu"], "tensor_core_filter": head["tensor_core_filter"],
        }
        for key in ("step_breakdown_ms", "merged_parity", "c_abi_shard_group", "row_sharded", "replicas_identical_to_row_sharded"):
            if key in head:
                line[key] = head[key]
        line["configs"] = configs
        line["bench_wall_s"] = time.time() - t_start
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
