"""Device-resident throughput of every BASELINE.json config shape on ONE GPU (per-GPU shard sizes for the sharded
configs), against the measured roofline.  One JSON line per config.  Not the driver's bench (bench.py is); its output
is committed under profiles/ as the per-config evidence SURVEY.md §8(d) asks for.

    python tools/bench_configs.py [c1 c1big c2a c2b c3 c4 c5] [--small]

Under torchrun (WORLD_SIZE > 1) only c3 and c4 run: every rank holds ONE shard of the 8-GPU layout (25M PQ rows /
12.5M RaBitQ rows, i.e. weak scaling: W x 25M rows at W GPUs), scans it for the whole query batch, and the per-shard
top-k lists go through the NCCL all-gather + device merge (C4: the two-exchange global-top-R rerank).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/bench_configs.py c3 c4
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = "--small" in sys.argv
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, "fallback"


HBM, TF, SRC = peaks()


def timed(fn, warm=2, reps=3):
    import torch.distributed as dist

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if WORLD > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if WORLD > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        ts.append(float(t.item()))
    return float(np.median(ts))


def qtc_stats():
    a, b = C.c_uint64(), C.c_uint64()
    L.call("vg_quant_tc_stats", C.byref(a), C.byref(b))
    return a.value, b.value


def out_bufs(nq, k, dev):
    return (torch.empty((nq, k), dtype=torch.int32, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev),
            torch.empty((nq,), dtype=torch.int32, device=dev))


def emit(name, workload, ms, nq, pairs, bytes_per_pair=None, flops=None, extra=None):
    if RANK != 0:
        return
    line = {"config": name, "n_gpus": WORLD, "workload": workload, "ms": ms, "qps": nq / ms * 1e3 if nq else None, "gpairs_per_s": pairs / ms / 1e6}
    if bytes_per_pair is not None:
        ach = pairs * bytes_per_pair / ms / 1e6
        line["roofline"] = {"bound": "hbm", "achieved_gbs": ach, "peak_gbs": HBM * WORLD, "frac": ach / (HBM * WORLD), "peak_source": SRC}
    if flops is not None:
        if "roofline" in line:
            line["hbm_equivalent"] = line.pop("roofline")  # per-query streaming bytes of the reference / time, vs the HBM peak
        ach = flops / ms / 1e9
        line["roofline"] = {"bound": "tensor", "achieved_tflops": ach, "peak_tflops_bf16": TF * WORLD, "frac": ach / (TF * WORLD), "peak_source": SRC,
                            "note": "useful FLOPs 2*Q*N*d of the filter GEMM (Flat: TF32, nominal dense peak half of bf16; quantized scans: fp16)"}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def flat(name, n, dim, nq, k):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(42)
    x = torch.rand((n, dim), dtype=torch.float32, device=dev, generator=g)
    q = torch.rand((nq, dim), dtype=torch.float32, device=dev, generator=torch.Generator(device=dev).manual_seed(43))
    ix = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n)
    ix.upload_dev(n, d_vectors=x.data_ptr())
    r, s, c = out_bufs(nq, k, dev)
    L.call("vg_flat_tc_enable", 1)
    qa, fb = C.c_uint64(), C.c_uint64()
    L.call("vg_flat_tc_stats", C.byref(qa), C.byref(fb))
    f0 = fb.value
    ms = timed(lambda: ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr()))
    L.call("vg_flat_tc_stats", C.byref(qa), C.byref(fb))
    r1, s1 = r.clone(), s.clone()
    L.call("vg_flat_tc_enable", 0)
    ms0 = timed(lambda: ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr()), warm=1, reps=1)
    L.call("vg_flat_tc_enable", 1)
    same = bool(torch.equal(r1, r) and torch.equal(s1.view(torch.int32), s.view(torch.int32)))
    emit(name, f"Flat exact L2, {n} x {dim} f32 U[0,1), {nq} queries, k={k} (tcgen05 filter: CTA-pair fp16 over an fp16 shadow + exact re-check)", ms, nq, n * nq,
         flops=2.0 * n * nq * dim,
         extra={"exact_cuda_core_scan_ms": ms0, "identical_to_exact_scan": same, "certificate_fallback_queries": fb.value - f0})
    ix.close()


def sq(name, codec, n, dim, nq, k):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(42)
    cb = dim if codec == "sq8" else dim // 2
    ix = (vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(np.full(dim, -4, np.float32), np.full(dim, 8 / 255, np.float32)))
          if codec == "sq8" else
          vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=dim, rows=n, int4=(np.full(dim, -4, np.float32), np.full(dim, 8, np.float32))))
    chunk = 1 << 20
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        codes = torch.randint(0, 256, (m, cb), dtype=torch.uint8, device=dev, generator=g)
        ix.upload_dev(m, d_codes=codes.data_ptr(), row0=r0)
    q = torch.randn((nq, dim), dtype=torch.float32, device=dev, generator=g)
    r, s, c = out_bufs(nq, k, dev)
    st0 = qtc_stats()
    ms = timed(lambda: ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr()), warm=1, reps=2)
    st1 = qtc_stats()
    emit(name, f"{codec.upper()} decode-and-scan, {n} x {dim}, {nq} queries, k={k} (uniform random codes)", ms, nq, n * nq, bytes_per_pair=cb,
         flops=2.0 * n * nq * dim, extra={"tensor_core_filter": {"queries": st1[0] - st0[0], "exact_rerun_queries": st1[1] - st0[1]}})
    ix.close()


def bq(name, n, dim, nq, k):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(42)
    cb = (dim + 63) // 64 * 8
    ix = vg.index.DeviceIndex(codec=L.CODEC_BQ, metric=0, dim=dim, rows=n, bq_threshold=0.0)
    chunk = 1 << 20
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        codes = torch.randint(0, 256, (m, cb), dtype=torch.uint8, device=dev, generator=g)
        ix.upload_dev(m, d_codes=codes.data_ptr(), row0=r0)
    q = torch.randn((nq, dim), dtype=torch.float32, device=dev, generator=g)
    r, s, c = out_bufs(nq, k, dev)
    st0 = qtc_stats()
    ms = timed(lambda: ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr()), warm=1, reps=2)
    st1 = qtc_stats()
    emit(name, f"BQ Hamming scan, {n} x {dim} bits, {nq} queries, k={k} (uniform random codes)", ms, nq, n * nq, bytes_per_pair=cb,
         flops=2.0 * n * nq * dim, extra={"tensor_core_filter": {"queries": st1[0] - st0[0], "exact_rerun_queries": st1[1] - st0[1]}})
    ix.close()


def pq(name, n, dim, m, nq, k):
    dev = torch.device(f"cuda:{LOCAL}")
    g = torch.Generator(device=dev).manual_seed(42 + RANK)
    rng = np.random.default_rng(0)
    ix = vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=n, row_base=RANK * n,
                              pq=(rng.integers(-128, 128, m * 256 * (dim // m), dtype=np.int8), np.full(m, 0.01, np.float32),
                                  np.zeros(m, np.float32), m, 256))
    chunk = 1 << 22
    for r0 in range(0, n, chunk):
        mm = min(chunk, n - r0)
        codes = torch.randint(0, 256, (mm, m), dtype=torch.uint8, device=dev, generator=g)
        ix.upload_dev(mm, d_codes=codes.data_ptr(), row0=r0)
    q = torch.randn((nq, dim), dtype=torch.float32, device=dev, generator=torch.Generator(device=dev).manual_seed(43))  # replicated
    r, s, c = out_bufs(nq, k, dev)
    sh = vg.sharded.ShardedIndex(ix, descending=False)
    st0 = qtc_stats()
    if WORLD > 1:
        ms = timed(lambda: sh.search_dev(q, nq, k), warm=1, reps=2)
    else:
        ms = timed(lambda: ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr()), warm=1, reps=2)
    st1 = qtc_stats()
    emit(name, f"PQ M={m} x 256 ADC scan, {WORLD} x {n} rows ({dim}-d), {nq} queries, k={k} (one 25M-row shard of the 8-GPU config per GPU"
         + (", NCCL all-gather + device merge of the per-shard top-k)" if WORLD > 1 else ")"), ms, nq, WORLD * n * nq,
         bytes_per_pair=m, flops=2.0 * WORLD * n * nq * dim,
         extra={"tensor_core_filter": {"queries": st1[0] - st0[0], "exact_rerun_queries": st1[1] - st0[1]}})
    ix.close()


def rabitq(name, n, dim, nq, r_top, k):
    dev = torch.device(f"cuda:{LOCAL}")
    g = torch.Generator(device=dev).manual_seed(42 + RANK)
    ix = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n, row_base=RANK * n)
    chunk = 1 << 18
    code_bytes = dim // 8 + 4
    codes = torch.empty((chunk, code_bytes), dtype=torch.uint8, device=dev)
    t0 = time.time()
    for r0 in range(0, n, chunk):
        mm = min(chunk, n - r0)
        x = torch.randn((mm, dim), dtype=torch.float32, device=dev, generator=g)
        L.call("vg_rabitq_encode_dev", x.data_ptr(), mm, dim, codes.data_ptr())
        ix.upload_dev(mm, d_codes=codes.data_ptr(), d_vectors=x.data_ptr(), row0=r0)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    q = torch.randn((nq, dim), dtype=torch.float32, device=dev, generator=torch.Generator(device=dev).manual_seed(43))  # replicated
    sh = vg.sharded.ShardedIndex(ix, descending=False)
    rr, ss, cc = out_bufs(nq, r_top, dev)
    st0 = qtc_stats()
    ms_scan = timed(lambda: ix.search_dev(q.data_ptr(), nq, r_top, rr.data_ptr(), ss.data_ptr(), cc.data_ptr()), warm=1, reps=2)
    ms = timed(lambda: sh.search_rerank_dev(q, nq, r_top, k), warm=1, reps=2)
    st1 = qtc_stats()
    emit(name, f"RaBitQ 1-bit scan + float32 rerank of top-{r_top}, {WORLD} x {n} x {dim}, {nq} queries, final k={k} (one 12.5M-row shard of "
         "the 8-GPU config per GPU" + (", two NCCL exchanges: global top-R, then exact scores)" if WORLD > 1 else ")"), ms, nq, WORLD * n * nq,
         bytes_per_pair=code_bytes, flops=2.0 * WORLD * n * nq * dim,
         extra={"scan_only_ms": ms_scan, "rerank_and_merge_ms": ms - ms_scan, "generate_encode_upload_s": gen_s,
                "tensor_core_filter": {"queries": st1[0] - st0[0], "exact_rerun_queries": st1[1] - st0[1]}})
    ix.close()


def pqa_stats():
    import ctypes as C

    a, b = C.c_uint64(), C.c_uint64()
    L.call("vg_pq_assign_tc_stats", C.byref(a), C.byref(b))
    return a.value, b.value


def pqtrain(name, n, dim, m, iters):
    rng = np.random.default_rng(42)
    x = rng.standard_normal((n, dim), dtype=np.float32)
    pq_ = vg.quantization.ProductQuantizer(dim, m, 256)
    torch.cuda.synchronize()
    t0 = time.time()
    pq_.Train(x, iters=iters, seed=1)
    torch.cuda.synchronize()
    s = time.time() - t0
    # the same training with the set already on the device (vg_pq_train_dev): what the kernels cost without the 3 GB host copy
    ds = dim // m
    dx = torch.from_numpy(x).cuda()
    cb, sc, of = np.zeros(m * 256 * ds, np.int8), np.zeros(m, np.float32), np.zeros(m, np.float32)
    torch.cuda.synchronize()
    ps0 = pqa_stats()
    t0 = time.time()
    L.call("vg_pq_train_dev", dx.data_ptr(), n, dim, m, 256, iters, 1, L.ptr(cb, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p), None)
    torch.cuda.synchronize()
    sd = time.time() - t0
    ps1 = pqa_stats()
    same = bool(np.array_equal(cb, pq_.codebooks) and np.array_equal(sc.view(np.uint32), pq_.scales.view(np.uint32)))
    hbm, tf = HBM, TF
    flops, byts = 2.0 * n * 256 * dim * iters, 4.0 * n * dim * iters  # assignment contraction, one pass over the samples per iteration
    print(json.dumps({"config": name, "workload": f"PQ codebook training (k-means++ init + {iters} Lloyd iterations), {n} x {dim}, {m} subspaces x 256 "
                      "centroids", "seconds": s, "samples_per_s": n * iters / s, "device_resident_seconds": sd,
                      "device_resident_samples_per_s": n * iters / sd, "identical_codebooks": same,
                      "tensor_core_assignment": {"pairs": ps1[0] - ps0[0], "exact_reevaluated_pairs": ps1[1] - ps0[1]},
                      "roofline": {"hbm": {"achieved_gbs": byts / sd / 1e9, "peak_gbs": hbm, "frac": byts / sd / 1e9 / hbm},
                                   "tensor": {"achieved_tflops": flops / sd / 1e12, "peak_tflops_bf16": tf, "frac": flops / sd / 1e12 / tf},
                                   "note": "algorithmic bytes / FLOPs of the Lloyd assignment only (SURVEY 8d) over the WHOLE training time (k-means++ init "
                                           "included); the order-exact float32 contract (sequential FMA distances, strict first-wins argmin, sample-order "
                                           "sums, sequential k-means++ prefix) bounds the rest: the assignment itself runs on the tensor cores "
                                           "(4.7 ms per pass, certificate + exact re-evaluation of the uncertain pairs)"},
                      "note": "`seconds` includes the host->device copy of the 3 GB training set from pageable memory; results identical to the "
                              "order-exact path: sequential-FMA distances, strict-< first-wins argmin, sample-order float32 centroid sums, "
                              "sequential k-means++ prefix sums"}), flush=True)


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c1big", "c2a", "c2b", "c3", "c4", "c5"]
    if WORLD > 1:
        import torch.distributed as dist

        which = [w for w in which if w in ("c3", "c4")]
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)  # NCCL banners go to stderr
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{LOCAL}"))
        torch.cuda.set_device(LOCAL)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    torch.cuda.set_device(LOCAL)
    L.call("vg_init", LOCAL)
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    for w in which:
        if w == "c1":
            flat("C1", 100_000, 128, 1000, 10)
        elif w == "c1big":
            flat("C1-scaled", 2_000_000 if SMALL else 10_000_000, 128, 1000 if SMALL else 4096, 10)
            flat("C1-768d", 1_000_000, 768, 2048, 10)
        elif w == "c2a":
            sq("C2a", "sq8", 1_000_000 if SMALL else 10_000_000, 768, 2048 if SMALL else 10_000, 100)
        elif w == "c2a8":  # the per-GPU work of the headline bench on eight GPUs: one 1.25M-row shard, the whole query batch
            sq("C2a-shard-of-8", "sq8", 1_250_000, 768, 10_000, 100)
        elif w == "c2b":
            sq("C2b", "int4", 1_000_000 if SMALL else 10_000_000, 768, 2048 if SMALL else 10_000, 100)
        elif w == "c3":
            pq("C3", 4_000_000 if SMALL else 25_000_000, 768, 96, 592 if SMALL else 10_000, 100)
        elif w == "c4":
            rabitq("C4", 2_000_000 if SMALL else 12_500_000, 1536, 512 if SMALL else 1000, 1000, 100)
        elif w == "bq":
            bq("BQ", 2_000_000 if SMALL else 12_500_000, 1536, 512 if SMALL else 1000, 1000)
        elif w == "c5":
            pqtrain("C5", 100_000 if SMALL else 1_000_000, 768, 96, 25)
    if WORLD > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
