#!/usr/bin/env python
"""bench.py -- LHub link-prediction rate (predicted edges/s) of the B200 path, next to the
reference's host-OpenMP path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rmat22] [--degree 16]
    python bench.py --impl reference ...      # the unmodified reference on the host cores

One *step* = one batch of the reference harness (main.cxx:208-221) at one hub threshold: the
graph with 10% of its edges removed is handed over, then all nine similarity measures are run
as LHub (MINDEGREE1 = --degree) asking for exactly the number of removed edges, like
PREDICT_LINKS does (main.cxx:50).  Metric: predicted edges per second of whole-job time.

  value : graph already resident in HBM, results left in HBM (device time, CUDA events).
  e2e   : the same step through the C ABI with HOST buffers -- nlp_set_graph() from pinned host
          memory and nlp_fetch() of every measure's (u, v, score) list inside the timed region.

N > 1 (torchrun, one rank per GPU, CSR replicated): by default every rank runs a whole step on its
own batch -- an independent random removal, as main.cxx:163 draws REPEAT_BATCH of them -- with no
data-path collective ("scaling": "weak"); --shard measures deals the nine predictions of ONE batch
to the ranks, --shard sources partitions the sources of every prediction (all-gather + merge).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np    # noqa: E402
import torch          # noqa: E402

WORKLOADS = {
    # name: (generator kwargs, removed fraction, description)
    "rmat22": dict(kind="rmat", scale=22, ef=16, seed=43, frac=0.1,
                   desc="R-MAT scale-22 ef-16 (0.57,0.19,0.19) ids permuted, 0.1|E| removed (BASELINE configs[1])"),
    "rmat20": dict(kind="rmat", scale=20, ef=16, seed=43, frac=0.1, desc="R-MAT scale-20 (smoke size)"),
    "rmat18": dict(kind="rmat", scale=18, ef=16, seed=42, frac=0.01, permute=False, desc="R-MAT scale-18 (ids as generated), 1e-2|E| removed (configs[0])"),
    "rmat16": dict(kind="rmat", scale=16, ef=16, seed=42, frac=0.1, desc="R-MAT scale-16 (tiny)"),
    # BASELINE configs[2..4]: parity/scale cases, run with --workload (not the default bench line)
    "road24m": dict(kind="road", side=4900, keep=0.6, seed=44, frac=0.1,
                    desc="road/mesh-shaped 4900x4900 lattice, edges kept with p=0.6 (~24M vertices, avg degree ~2.4), 0.1|E| removed (configs[2])"),
    "web50m": dict(kind="web", n=50_000_000, avg_out=150, seed=45, frac=0.1, leaves=True,
                   desc="web-crawl-shaped, 50M vertices, ~3e9 generated links (~2.2e9 undirected edges, > 2^32 directed entries), "
                        "power-law out-degree, host locality, 40% leaf pages of degree 1-4 (BASELINE configs[3], sk-2005 scale)"),
    "web50m_140": dict(kind="web", n=50_000_000, avg_out=140, seed=45, frac=0.1, leaves=True,
                       desc="web50m with avg_out = 140 (4.52e9 directed entries before, 4.18e9 after the removal): the graph of the first 1/4/8-GPU lines"),
    "web50m_r1": dict(kind="web", n=50_000_000, avg_out=19, seed=45, frac=0.1, desc="round 1's under-sized configs[3] stand-in (1.0e9 directed entries)"),
    "web25m": dict(kind="web", n=25_000_000, avg_out=19, seed=45, frac=0.1, desc="web-crawl-shaped, 25M vertices (configs[3] at half scale)"),
    "web6m": dict(kind="web", n=6_250_000, avg_out=19, seed=45, frac=0.1, desc="web-crawl-shaped, 6.25M vertices (configs[3] at 1/8 scale)"),
    "rmat24": dict(kind="rmat", scale=24, ef=16, seed=46, frac=0.01, desc="R-MAT scale-24, 1e-2|E| removed (configs[4]; use --degree 0)"),
    "rmat21": dict(kind="rmat", scale=21, ef=16, seed=46, frac=0.01, desc="R-MAT scale-21, 1e-2|E| removed (configs[4] at 1/8 scale; use --degree 0)"),
}
MEASURES = ["CN", "JC", "SI", "SC", "HP", "HD", "LHN", "AA", "RA"]
METRIC = "lhub_predicted_edges_per_s"
DTYPE = "u32 counts, f32 scores (f64 terms for AA/RA/Salton)"


def make_config(info, degree, measures, S, M):
    """The workload description -- identical in both arms (the driver compares the dicts);
    everything run-specific lives in the line's "run" object."""
    return dict(info, min_degree1=degree, measures=list(measures),
                l2="inputs larger than L2 (CSR %.0f MB)" % ((8 * (S + 1) + 4 * M) / 1e6))


BASE_GRAPH = {}           # the last workload's graph before the removal (the e2e leg applies the batch on the device)
REMOVAL_SEED = 12345      # default_random_engine(12345) for every parity run


def build_workload(name, device, batch=0, pred=None):
    """The graph of workload `name` with its edges removed the way the reference removes them
    (main.cxx:166-169): generateEdgeDeletions draws `frac * |E| / 2` times "random vertex, then a
    random entry of its row" from std::default_random_engine(12345 + batch), tidyBatchUpdateU makes
    the list unique, applyBatchUpdateOmpU deletes both directions.  `batch` picks the random
    removal: the reference draws REPEAT_BATCH = 5 independent ones per fraction (main.cxx:26-28,163).
    With a Predictor the draw runs on the GPU (nlp_generate_deletions, draw for draw the reference's
    sequence); without one (the --impl reference arm) the oracle's restatement of the generator
    runs on the host -- tests/test_batch_oracle.py pins the two to each other and to the reference.
    Returns (offsets, keys, K, info, (del_u, del_v)): K = removed undirected edges = the prediction
    count PREDICT_LINKS asks for (main.cxx:50); del_* = main.cxx's sorted directed `deletions0`."""
    import nlp_b200 as N
    w = WORKLOADS[name]
    t0 = time.time()
    if w["kind"] == "rmat":
        off, keys = N.graphs.rmat(w["scale"], w["ef"], w["seed"], permute=w.get("permute", True), device=device)
    elif w["kind"] == "road":
        off, keys = N.graphs.road_lattice(w["side"], w["keep"], w["seed"], device=device)
    else:
        off, keys = N.graphs.web_crawl(w["n"], w["avg_out"], seed=w["seed"], device=device, leaves=w.get("leaves", False))
    S = int(off.numel() - 1)
    if device != "cpu":
        torch.cuda.synchronize()
        torch.cuda.empty_cache()        # the library allocates with cudaMalloc: hand the generator's cached blocks back
    batch_size = int(w["frac"] * int(keys.numel()) / 2)            # size_t(d * x.size() / 2), main.cxx:166
    seed = REMOVAL_SEED + batch
    base = (off, keys)
    if pred is not None:
        # on the device: draw (nlp_generate_deletions), apply (nlp_apply_deletions), copy the rebuilt CSR out
        pred.set_graph_pointers(off.data_ptr(), keys.data_ptr(), S, device=True, keep=(off, keys))
        n, _ = pred.generate_deletions(seed, batch_size, fetch=False)
        du = torch.empty(n, dtype=torch.int32, device=device); dv = torch.empty(n, dtype=torch.int32, device=device)
        if n:
            pred.fetch_deletions_into(du.data_ptr(), dv.data_ptr(), n)
        pred.apply_deletions(pointers=(du.data_ptr(), dv.data_ptr(), n))
        S2, M2 = pred.graph_size()
        off = torch.empty(S2 + 1, dtype=torch.int64, device=device); keys = torch.empty(M2, dtype=torch.int32, device=device)
        pred._check(pred.lib.nlp_fetch_graph(pred.h, off.data_ptr(), keys.data_ptr() if M2 else None))
    else:
        from oracle import oracle_py as O
        offn, keysn = N.graphs.to_numpy(off, keys)
        u, v, _ = O.oracle_edge_deletions(offn, keysn, seed, batch_size)
        du = torch.from_numpy(u.astype(np.int32)).to(device); dv = torch.from_numpy(v.astype(np.int32)).to(device)
        del offn, keysn
        off, keys = N.graphs.apply_deletions(off, keys, du, dv)
    K = int(du.numel()) // 2
    if device != "cpu":
        torch.cuda.synchronize()
        torch.cuda.empty_cache()        # the library allocates with cudaMalloc: hand torch's cached blocks back
    info = {"workload": name, "description": w["desc"], "span": S, "entries": int(keys.numel()),
            "predict_count_K": K, "removal": "reference sampler (inc/batch.hxx:99-112), default_random_engine(%d)" % seed}
    BASE_GRAPH["graph"] = base
    return off, keys, K, info, (du, dv), round(time.time() - t0, 2)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:   # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:   # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:   # noqa: BLE001
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_step(R, K, degree, measures, threads, omp=True):
    """One step on the unmodified reference.  Returns (edges, ref_ms, wall_s).  omp=True (the
    OpenMP templates, inc/predict.hxx:409-467) is only safe when at least K candidates exist: the
    reference's merge reads an empty vector otherwise (inc/predict.hxx:424,452-453; observed
    segfault).  omp=False runs the sequential templates (inc/predict.hxx:358-374) on one core."""
    edges, ref_ms = 0, 0.0
    t0 = time.perf_counter()
    for m in measures:
        u, v, s, tm, ts = R.predict(m, degree, max_edges=K, omp=omp, threads=threads, canonical=False)
        edges += len(u)
        ref_ms += tm
    return edges, ref_ms, time.perf_counter() - t0


class PortGraph:
    """Stand-in with RefGraph's predict() signature over the plain-C oracle (OpenMP over sources,
    safe on shortfall); only used when oracle/_ref/libnlpref.so is absent."""

    def __init__(self, O, offsets, keys):
        self.O, self.off, self.keys = O, offsets, keys

    def predict(self, measure, min_degree1, max_edges, omp=True, threads=0, canonical=False):
        t0 = time.perf_counter()
        u, v, s, _ = self.O.oracle_predict(self.off, self.keys, measure, min_degree1, max_edges=max_edges, threads=threads)
        ms = (time.perf_counter() - t0) * 1e3
        return u, v, s, ms, ms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle_py as O
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    off, keys, K, info, _, build_s = build_workload(args.workload, dev)
    import nlp_b200 as N
    offn, keysn = N.graphs.to_numpy(off, keys)
    S, M = int(off.numel() - 1), int(keys.numel())
    del off, keys
    threads = os.cpu_count() or 1
    # How many pairs qualify at all?  The C oracle (OpenMP, shortfall-safe) answers that before the
    # reference runs: with fewer than K of them the reference's OpenMP merge reads an empty vector
    # (inc/predict.hxx:424,452-453; observed segfault), so its sequential templates are timed instead.
    _, _, _, st = O.oracle_predict(offn, keysn, "JC", args.degree, max_edges=1, threads=threads)   # torchrun sets OMP_NUM_THREADS=1
    use_omp = st["kept"] >= K
    if O.ref_available():
        kind = "reference"
        R = O.RefGraph(offn, keysn)
    else:
        # oracle/_ref did not travel: the plain-C port of the same algorithm stands in
        kind = "port"
        R = PortGraph(O, offn, keysn)
        use_omp = True
    if not use_omp:
        threads = 1
    e1, ms1, w1 = reference_step(R, K, args.degree, ["JC"], threads, omp=use_omp)
    total_steps = args.steps + args.warmup
    per_measure = max(w1, 1e-3)
    nm = int(max(1, min(len(MEASURES), 150.0 / (per_measure * total_steps))))
    order = ["JC", "AA", "CN", "SC", "RA", "SI", "HP", "HD", "LHN"]
    all_measures = [m for m in args.measures.split(",") if m] if args.measures else MEASURES
    sample = [m for m in all_measures if m in order[:nm]] or all_measures[:1]
    for _ in range(args.warmup):
        reference_step(R, K, args.degree, sample, threads, omp=use_omp)
    edges = 0; ref_ms = 0.0; wall = 0.0
    for _ in range(args.steps):
        e, ms, w = reference_step(R, K, args.degree, sample, threads, omp=use_omp)
        edges += e; ref_ms += ms; wall += w
    value = edges / (ref_ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": make_config(info, args.degree, all_measures, S, M),
        "run": {"build_s": build_s, "measures_run": sample, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": threads, "kind": kind,
                         "sample": "%d of 9 measures per step (%s), full graph, reference's own `time` field, %s"
                                   % (len(sample), ",".join(sample),
                                      "OpenMP templates" if use_omp else "SEQUENTIAL templates (fewer candidates than K: the OpenMP merge is undefined)")},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_edges_per_s": edges / wall,
    }
    print(json.dumps(line))
    return 0


def run_b200(args):
    import torch.distributed as dist
    import nlp_b200 as N
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = "cuda:%d" % local
    all_measures = [m for m in args.measures.split(",") if m] if args.measures else MEASURES
    # N > 1 (CSR replicated on every GPU in all modes):
    #  comm     (default) the north star's design: the sources of EVERY prediction are partitioned by
    #           wedge work, the library merges the ranks' candidates itself over its own NCCL
    #           communicator (nlp_comm_init: all-reduced select histograms for the global cutoff, one
    #           all-gather, final sort) and every rank ends with the full result: strong scaling
    #  batches  every rank runs a whole step on ITS OWN batch -- the reference draws REPEAT_BATCH
    #           independent random removals per fraction (main.cxx:163); independent units, no
    #           collective: weak scaling.  Also measured (shortly) in comm mode as the "replicas" key.
    #  measures the nine predictions of ONE batch dealt to the ranks (strong scaling, no collective)
    #  sources  round-1 plumbing: nlp_set_partition + torch.distributed all-gather + nlp_merge
    shard = args.shard
    if shard == "auto":
        shard = "comm"
    if world == 1:
        shard = "none"
    pred = N.Predictor(local)
    wl = build_workload(args.workload, dev, batch=rank if shard == "batches" else 0, pred=pred)
    off, keys, K, info, (du, dv), build_s = wl
    S = int(off.numel() - 1); M = int(keys.numel())
    if shard == "comm":
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(N.Predictor.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        torch.cuda.synchronize()
        pred.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    elif shard == "sources":
        pred.set_partition(rank, world)
    stream = torch.cuda.ExternalStream(pred.lib.nlp_stream(pred.h), device=dev)
    measures = all_measures[rank::world] if shard == "measures" else all_measures
    D = args.degree
    do_e2e = not args.no_e2e
    base = BASE_GRAPH.get("graph")                 # (off0, keys0) before the removal, still on the device
    if do_e2e:
        # host side of a step: the batch update (main.cxx's sorted directed deletions0) in pinned memory,
        # two sets of pinned result buffers (transfers are double buffered, nlp_fetch_async)
        h_du = du.cpu().pin_memory(); h_dv = dv.cpu().pin_memory()
        h_out = [[torch.empty(K, dtype=torch.int32).pin_memory() for _ in range(2)] + [torch.empty(K, dtype=torch.float32).pin_memory()]
                 for _ in range(2)]
        if args.e2e_upload:
            h_off = off.cpu().pin_memory(); h_keys = keys.cpu().pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def predict(m):
        if shard == "sources":
            r, n, ms = N.distributed.predict_distributed(pred, m, D, K)
            return r, n
        r = pred.predict(m, D, max_edges=K)
        return r, r["count"]

    def one_step(collect=None):
        edges = 0
        for m in measures:
            r, n = predict(m)
            edges += n
            if collect is not None:
                collect.append(r)
        return edges

    def fetch_results(i, n):
        if shard in ("comm", "sources") and rank != 0:
            return                                  # every rank holds the full result; rank 0 hands it to the host
        if shard == "sources":
            pred.fetch_into(h_out[0][0].data_ptr(), h_out[0][1].data_ptr(), h_out[0][2].data_ptr(), n)
        else:    # the device -> host transfer of this result overlaps the next prediction
            o = h_out[i % 2]
            pred.fetch_async(o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), n)

    e2e_marks = []

    def one_step_e2e():
        # One batch of the harness, host buffers in, host results out (main.cxx:164-169 + 208-221): back to
        # the base graph (resident since it was loaded, as main.cxx's x; nlp_graph_rollback = duplicate(x)),
        # send this batch's deletions from pinned host memory, apply them on the device
        # (nlp_apply_deletions), predict, fetch.
        pred.graph_rollback()
        pred.apply_deletions(pointers=(h_du.data_ptr(), h_dv.data_ptr(), int(h_du.numel())))
        edges = 0
        for i, m in enumerate(measures):
            r, n = predict(m)
            fetch_results(i, n)
            edges += n
        pred.fetch_wait()
        e2e_marks.append(time.perf_counter())
        return edges

    def one_step_upload():
        # the same with the whole CSR of the batch's graph uploaded from pinned host memory (round 1's e2e)
        pred.set_graph_pointers(h_off.data_ptr(), h_keys.data_ptr(), S, device=False, keep=(h_off, h_keys))
        edges = 0
        for i, m in enumerate(measures):
            r, n = predict(m)
            fetch_results(i, n)
            edges += n
        pred.fetch_wait()
        return edges

    def timed(fn, steps, collect=False):
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        results = [] if collect else None
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        edges = 0
        for _ in range(steps):
            edges += fn(results) if collect else fn()
        ev1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
            if shard in ("measures", "batches"):     # every rank predicted something different: whole-job edge count
                e = torch.tensor([edges], device=dev, dtype=torch.int64)
                dist.all_reduce(e, op=dist.ReduceOp.SUM)
                edges = int(e[0])
        return edges, ms, wall, results

    # ---- value: graph resident in HBM ---------------------------------------------------------
    pred.set_graph_pointers(off.data_ptr(), keys.data_ptr(), S, device=True, keep=(off, keys))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs ~0.1 s to come up: start it before the warm-up
    for _ in range(args.warmup):
        one_step()
    launches0 = pred.launch_count()
    comm0 = pred.comm_bytes()
    edges, ms, wall, results = timed(one_step, args.steps, collect=True)
    launches = pred.launch_count() - launches0
    comm_bytes = pred.comm_bytes() - comm0
    value = edges / (ms / 1e3)
    # The sampler covers the warm-up and the timed region of `value`.  It is stopped here: every
    # nvidia-smi query holds a driver lock for tens of milliseconds, which the device-timed region above
    # does not see (launches queue up) but the host-synchronous e2e steps below do (one 55 ms step in 20,
    # `e2e.host_ms_per_step_min_median_max`).  BENCH_SAMPLER_ALL=1 keeps it running to the end.
    clocks = None
    if rank == 0 and not os.environ.get("BENCH_SAMPLER_ALL"):
        clocks = sampler.stop()

    # in comm mode: is every rank's result the single-GPU result?  (checked outside the timed region)
    identical = None
    job = None
    if shard == "comm" and not args.no_identical:
        import hashlib

        def digest(m):
            r = pred.predict(m, D, max_edges=K)
            u, v, s_ = pred.fetch(r["count"])
            return hashlib.sha256(u.tobytes() + v.tobytes() + s_.tobytes()).hexdigest() + ":%d" % r["count"]
        check = [m for m in ("JC", "CN", "AA") if m in measures] or measures[:1]
        with_comm = [digest(m) for m in check]
        pred.comm_destroy()                          # leaves the handle on (0, 1): the whole prediction on this GPU
        alone = [digest(m) for m in check]
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.frombuffer(bytearray(N.Predictor.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(uid, 0)
        torch.cuda.synchronize()
        pred.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        t = torch.tensor([int(with_comm == alone)], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        identical = bool(int(t[0]))
    if world > 1:
        # job-wide counters: the source-centric kernels count per rank, the bucket path's plan counters
        # (first hop, eligible first hop, wedges) are already those of the whole graph
        mine = [sum(r["wedges"] for r in results), sum(r["candidates"] for r in results),
                sum(r["eligible_first_hop"] for r, m in zip(results, measures * args.steps) if m in ("AA", "RA")),
                sum(r["count"] for r in results)]
        t = torch.tensor(mine, device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = [int(x) for x in t.tolist()]
        global_stats = shard in ("comm", "sources") and results and results[0]["path"] == 2
        job = {"W": mine[0] if global_stats else tot[0], "C": tot[1], "flt_elig": mine[2] if global_stats else tot[2],
               "Kout": mine[3] if shard in ("comm", "sources") else tot[3],
               "nrun": len(results) * (world if shard in ("batches", "measures") else 1)}

    # ---- e2e: host buffers in, host results out ----------------------------------------------
    e2e = None
    e2e_upload = None
    if do_e2e:
        pred.set_graph_pointers(base[0].data_ptr(), base[1].data_ptr(), S, device=True, keep=base)
        pred.graph_checkpoint()
        for _ in range(max(1, min(args.warmup, 2))):
            one_step_e2e()
        # PCIe probe (explains run-to-run spread of the end-to-end figure: the step moves 0.5 GB to the host)
        pe0 = torch.cuda.Event(enable_timing=True); pe1 = torch.cuda.Event(enable_timing=True)
        probe_src = torch.empty(K, dtype=torch.int32, device=dev)
        h_out[0][0].copy_(probe_src, non_blocking=True); torch.cuda.synchronize()
        pe0.record(); h_out[0][0].copy_(probe_src, non_blocking=True); pe1.record(); torch.cuda.synchronize()
        d2h_gbps = 4 * K / (pe0.elapsed_time(pe1) * 1e-3) / 1e9
        del probe_src
        del e2e_marks[:]
        e2e_marks.append(time.perf_counter())
        e_edges, e_ms, e_wall, _ = timed(one_step_e2e, args.steps)
        gaps = sorted((b - a) * 1e3 for a, b in zip(e2e_marks[-args.steps - 1:-1], e2e_marks[-args.steps:]))
        copies = 1 if shard in ("none",) else world      # every rank receives the batch
        e2e = {"value": e_edges / (e_ms / 1e3), "unit": "edges/s", "h2d_bytes_per_step": 8 * int(h_du.numel()) * copies,
               "d2h_bytes_per_step": int(e_edges / args.steps) * 12, "ms_per_step": e_ms / args.steps,
               "host_ms_per_step_min_median_max": [round(gaps[0], 2), round(gaps[len(gaps) // 2], 2), round(gaps[-1], 2)] if gaps else None,
               "pinned_d2h_probe_gbps": round(d2h_gbps, 1),
               "step": "nlp_graph_rollback to the resident base graph, H2D of the batch's %d directed deletions from pinned memory, nlp_apply_deletions "
                       "(CSR rebuilt on the device), %d predictions, every (u, v, score) list fetched to pinned host memory" % (int(h_du.numel()), len(measures))}
        if args.e2e_upload:
            for _ in range(max(1, min(args.warmup, 2))):
                one_step_upload()
            u_edges, u_ms, u_wall, _ = timed(one_step_upload, args.steps)
            e2e_upload = {"value": u_edges / (u_ms / 1e3), "unit": "edges/s", "h2d_bytes_per_step": ((S + 1) * 8 + M * 4) * copies,
                          "d2h_bytes_per_step": int(u_edges / args.steps) * 12, "ms_per_step": u_ms / args.steps,
                          "step": "the whole CSR of the batch's graph uploaded from pinned host memory (nlp_set_graph) instead of the batch update"}
        pred.set_graph_pointers(off.data_ptr(), keys.data_ptr(), S, device=True, keep=(off, keys))
    # ---- extra: the same step with reuse across measures (nlp_set_reuse) -----
    sweep = None
    if results and results[0]["path"] in (2, 3) and shard == "none" and args.sweep_reuse:
        def one_step_reuse():
            pred.set_reuse(True)
            return one_step()
        for _ in range(max(1, min(args.warmup, 2))):
            one_step_reuse()
        r_edges, r_ms, r_wall, _ = timed(one_step_reuse, args.steps)
        pred.set_reuse(False)
        sweep = {"value": r_edges / (r_ms / 1e3), "unit": "edges/s", "ms_per_step": r_ms / args.steps,
                 "note": "sorted wedge records shared by the measures of a step; store emptied every step"}
    # ---- extra (comm mode): N independent replicas, one whole step per GPU on the same batch ----------
    replicas = None
    if shard == "comm" and not args.no_replicas:
        pred.comm_destroy()
        for _ in range(2):
            one_step()
        barrier()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        rep_edges = 0
        for _ in range(args.steps):
            rep_edges += one_step()
        ev1.record(stream)
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        replicas = {"value": rep_edges * world / (float(t[0]) / 1e3), "unit": "edges/s", "ms_per_step": float(t[0]) / args.steps,
                    "note": "%d independent replicas (one whole step per GPU, no collective): weak scaling, for comparison" % world}
    if rank == 0 and clocks is None:
        clocks = sampler.stop()                          # BENCH_SAMPLER_ALL: sampled over warm-up + all timed regions
    if shard == "comm":
        pred.comm_destroy()             # rank 0 goes on alone (parity check against the oracle)
    elif shard == "sources":
        pred.set_partition(0, 1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant phase --------------------------------------------------------
    path = results[0]["path"]
    if path == 2:
        names = ["bucket: plan lookup (built once per graph and D)", "bucket: k_bucket (gather + shared-memory sort + reduce + score)",
                 "bucket: big sources (k_pair_emit + global radix sort + k_pair_reduce + k_big_place)",
                 "bucket: k_score (exclusion + scoring, one thread per slot)", "-", "-", "-",
                 "select+sort (radix top-K)"]
    elif path == 3:
        names = ["pair: eligible rows + item descriptors + scans", "pair: k_pair_emit (wedge records)",
                 "pair: radix sort by (u,v) (k_tilehist+k_rowscan+k_scatter per digit)", "pair: k_pair_reduce (run count + exclusion + score)",
                 "-", "-", "-", "select+sort (radix top-K)"]
    else:
        names = ["frontier(k_elig+k_work+k_bin)", "hub-heavy sources (k_range / k_range_flt + k_dense)", "k_hash(16K)", "k_hash(4K)", "k_hash(1K)",
                 "k_tiny<32>", "k_tiny<8>", "select+sort (radix top-K)"]
    phase = [sum(r["phase_ms"][i] for r in results) for i in range(8)]
    nrun = len(results)
    peak, peak_src = peaks()
    W = sum(r["wedges"] for r in results); C = sum(r["candidates"] for r in results)
    E = sum(r["emitted"] for r in results); Kout = sum(r["count"] for r in results)
    if path == 2:
        # own traffic of the bucket path: item descriptors (28 B per eligible
        # first-hop entry), one key per wedge record gathered, 4 B of aligned score per record, 8 B per kept pair
        P = sum(r["pair_records"] for r in results)
        elig = sum(r["eligible_first_hop"] for r in results)
        cands = [(names[1], phase[1], 28 * elig + 4 * P + 4 * P + 8 * C),
                 (names[2], phase[2], 0.0),
                 (names[3], phase[3], 4 * P + 8 * C + 4 * P),
                 (names[7], phase[7], 4 * 4 * P + 12 * E + 12 * Kout)]
    elif path == 3:
        # per-phase algorithmic bytes of the pair path
        idbits = max(1, (S - 1).bit_length())
        digits = 2 * ((idbits + 7) // 8)
        P = sum(r["pair_records"] for r in results)
        Pflt = sum(r["pair_records"] for r, m in zip(results, measures * args.steps) if m in ("AA", "RA"))
        rec_bytes = 8 * P + 4 * Pflt
        elig = sum(r["eligible_first_hop"] for r in results)
        cands = [(names[0], phase[0], nrun * 12 * S / max(1, world) + 4 * elig),
                 (names[1], phase[1], 4 * P + rec_bytes),
                 (names[2], phase[2], digits * 2 * rec_bytes),
                 (names[3], phase[3], rec_bytes + 12 * E),
                 (names[7], phase[7], 12 * E + 12 * Kout)]
    else:
        wedge_ms = sum(phase[1:7])
        if wedge_ms == 0:   # pruned candidate buffer (passes > 1, e.g. IHub): per-phase events are off
            wedge_ms = sum(r["scoring_ms"] - r["frontier_ms"] for r in results)
        # algorithmic bytes, split by the phase that moves them
        cands = [("frontier: k_work (first-hop scan, hub test, work sums)", phase[0], nrun * (8 * (S + 1) + 4 * M + 4 * M) / max(1, world)),
                 ("wedge kernels (k_dense/k_hash/k_tiny: 2-hop scan + count + score)", wedge_ms, 4 * W + 4 * C + 12 * E),
                 (names[7], phase[7], 12 * E + 12 * Kout)]
    dom = max(cands, key=lambda c: c[1])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from `ncu --set full` (see profiles/README.md)
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload + ":" + str(D), {}).get(dom[0])
    kernel_gbs = dom[2] / (dom[1] / 1e3) / 1e9 if dom[1] > 0 else 0.0
    # The roofline SURVEY.md section 8(d) defines: the reference ALGORITHM's bytes
    #   B_alg = 8(S+1) + 4M + 4M  (offsets, first-hop adjacency, one degree word per first-hop entry)
    #         + 4 W(D)            (one key per enumerated wedge)
    #         + 4 C               (deg(v) per scored candidate)          + 12 K_out (result)
    #         [+ 4 M_elig for the float measures: the term per eligible first-hop entry]
    # per prediction (W, C, K_out and M_elig are the kernels' own counters), summed over the
    # predictions of the timed region, over the device time of that region, against the measured
    # HBM copy bandwidth.  `kernel_*` is the dominant phase against the traffic this implementation
    # chose to move through it (the figure round 1 called `frac`).
    flt_elig = sum(r["eligible_first_hop"] for r, m in zip(results, measures * args.steps) if m in ("AA", "RA"))
    jW, jC, jF, jK, jn = (job["W"], job["C"], job["flt_elig"], job["Kout"], job["nrun"]) if job else (W, C, flt_elig, Kout, nrun)
    total_alg = jn * (8 * (S + 1) + 4 * M + 4 * M) + 4 * jW + 4 * jC + 12 * jK + 4 * jF       # the whole job
    achieved = total_alg / (ms / 1e3) / 1e9
    peak_job = peak * world
    step_frac = achieved / peak_job
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_job, "unit": "GB/s", "frac": step_frac,
                "traffic": traffic, "peak_source": peak_src,
                "scope": "whole step: SURVEY.md 8(d) algorithmic bytes of all predictions of the job / device time of the step / (GPUs x measured HBM copy bandwidth)",
                "algorithmic_bytes_per_step": total_alg / args.steps,
                "kernel": dom[0], "kernel_gbs": kernel_gbs, "kernel_frac": kernel_gbs / peak,
                "kernel_share_of_step": dom[1] / max(1e-9, sum(phase)),
                "kernel_own_bytes_per_step": dom[2] / args.steps}

    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample -------
    # ---- parity: the full (u, v, score bits) lists of the sample against the C oracle ------------
    cpu = None
    parity = {"status": "skipped", "detail": "--no-cpu-baseline / --no-verify"}
    try:
        from oracle import oracle_py as O
        threads = os.cpu_count() or 1
        sample = [m for m in ("JC", "AA") if m in measures] or measures[:1]
        offn = keysn = None
        if not (args.no_cpu_baseline and args.no_verify):
            offn, keysn = N.graphs.to_numpy(off, keys)
        if args.no_cpu_baseline:
            cpu = {"value": None, "unit": "edges/s", "cores": 0, "kind": "reference", "sample": "skipped (--no-cpu-baseline)"}
        elif O.ref_available():
            R = O.RefGraph(offn, keysn)
            enough = all(r["count"] == K for r in results)     # fewer candidates than K: OpenMP merge undefined
            rthreads = threads if enough else 1
            e, rms, w = reference_step(R, K, D, sample, rthreads, omp=enough)
            cpu = {"value": e / (rms / 1e3), "unit": "edges/s", "cores": rthreads, "kind": "reference",
                   "sample": "one pass of %s (%d of %d measures) on the full graph, reference %s templates, "
                             "reference's own `time` field (%.0f ms wall)" % ("+".join(sample), len(sample), len(measures),
                                                                              "OpenMP" if enough else "SEQUENTIAL (fewer candidates than K: OpenMP merge undefined)", w * 1e3),
                   "edges": e}
            R.close()
        else:
            t0 = time.perf_counter()
            u, v, s_, st = O.oracle_predict(offn, keysn, "JC", D, max_edges=K)
            w = time.perf_counter() - t0
            cpu = {"value": len(u) / w, "unit": "edges/s", "cores": threads, "kind": "port",
                   "sample": "one JC pass of the C oracle (OpenMP) on the full graph"}
        if not args.no_verify:
            # The reference breaks score ties by heap accident (inc/predict.hxx:332), so its list is
            # canonicalised by the oracle: the plain-C restatement of the same loop (pinned against the
            # compiled reference, tests/test_oracle_vs_reference.py) with the (score desc, u, v) order.
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import parity as PT
            bad, rows = [], 0
            for m in sample:
                r = pred.predict(m, D, max_edges=K)
                got = pred.fetch(r["count"])
                wu, wv, ws, st = O.oracle_predict(offn, keysn, m, D, max_edges=K, threads=threads)
                err = PT.compare(got, (wu, wv, ws), "%s D=%d K=%d" % (m, D, K))
                if err is None:
                    for k in ("first_hop", "eligible_first_hop", "wedges", "candidates", "kept"):
                        if r[k] != st[k]:
                            err = "%s: counter %s = %d, oracle %d" % (m, k, r[k], st[k])
                            break
                rows += len(wu)
                if err:
                    bad.append(err)
            parity = {"status": "bit-exact" if not bad else "MISMATCH", "measures": sample, "rows_compared": rows,
                      "against": "oracle/nlp_oracle.c (C restatement of inc/predict.hxx:214-265, canonical tie order), full (u, v, score bits) lists + wedge/candidate counters",
                      "detail": bad}
    except Exception as ex:   # noqa: BLE001
        if cpu is None:
            cpu = {"value": None, "unit": "edges/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}
        else:
            parity = {"status": "failed", "detail": repr(ex)}

    line = {
        "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if shard in ("none", "batches") else "strong",
        "identical_to_single_gpu": identical,
        "comm_bytes_per_step": comm_bytes / args.steps if world > 1 else 0,
        "replicas": replicas,
        "e2e_upload": e2e_upload,
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": make_config(info, D, all_measures, S, M),
        "run": {"build_s": build_s, "shard": shard,
                "parallelism": {"none": "1 GPU",
                                "comm": "sources of every prediction partitioned by wedge work over %d GPUs (contiguous source ranges of equal wedge-record count; CSR replicated), merged inside the library over NCCL: all-reduced select histograms (global cutoff), one all-gather of the survivors, final on-device sort; every rank holds the result" % world,
                                "batches": "%d batches (independent random removals of the same graph, main.cxx:163), one whole step per GPU, CSR replicated, no collective" % world,
                                "measures": "the %d predictions of a step dealt to %d GPUs (independent units, no collective), CSR replicated" % (len(all_measures), world),
                                "sources": "sources of every prediction partitioned over %d GPUs, CSR replicated, one all-gather + on-device merge" % world}[shard]},
        "parity": parity,
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "sweep_reuse": sweep,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wedges_per_s": jW / (sum(r["scoring_ms"] for r in results) / 1e3),
        "step_hbm_frac": step_frac,
        "phase_ms_per_step": {n: p / args.steps for n, p in zip(names, phase)},
        "wall_ms_per_step": wall * 1e3 / args.steps,
        "bins": results[0]["bin_sources"][:7],
        "path": {1: "source-centric", 2: "bucket (pair)", 3: "pair, global sort"}.get(path, path),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rmat22", choices=sorted(WORKLOADS))
    ap.add_argument("--degree", type=int, default=16, help="MINDEGREE1 of the LHub runs (0 = IHub)")
    ap.add_argument("--measures", default="", help="comma list (default: all nine)")
    ap.add_argument("--shard", default="auto", choices=["auto", "comm", "batches", "measures", "sources"],
                    help="N > 1: comm (default) = sources of every prediction partitioned by wedge work, merged inside the library "
                         "over its NCCL communicator; batches = one batch per rank (weak scaling); measures = the predictions of one "
                         "batch dealt to the ranks; sources = round-1 plumbing (torch.distributed all-gather + nlp_merge)")
    ap.add_argument("--no-replicas", action="store_true", help="comm mode: skip the extra independent-replicas measurement")
    ap.add_argument("--no-identical", action="store_true", help="comm mode: skip the check that every rank's result equals the single-GPU result "
                                                              "(it runs the sample on ONE GPU: minutes for IHub at scale 24)")
    ap.add_argument("--e2e-upload", action="store_true", help="also time the e2e step with the whole CSR uploaded per step (round 1's e2e)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (large workloads)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference run on the host cores")
    ap.add_argument("--sweep-reuse", action="store_true", help="also time the step with nlp_set_reuse (sorted-record store of the global-sort pair path)")
    ap.add_argument("--no-verify", action="store_true", help="skip the full-list parity check against the C oracle")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
