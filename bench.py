#!/usr/bin/env python
"""Benchmark of the RandomForestMC hot path on B200 (and of the CPU reference arm beside it).

Metric (BASELINE.json): fit candidate-trees/s and predict rows*trees/s.  One JSON line on stdout.

Workload: synthetic 1M rows x 64 mixed numeric/categorical features,
4 classes, n_trees=256, max_discard_trees=4.  Two regimes of the same config:
  c2b (default)  batch_train_pclass=2000, batch_val_pclass=500  -> 8000 train + 2000 hold-out rows
                 per candidate: the regime in which the grower does real work;
  c2a            the reference's default 10/10 rows per class (40 + 40 rows per candidate).

A "step" = one pass of the hot path over one batch: growing and scoring every candidate of
``n_trees`` slots (fit), and voting 1M rows through the fitted forest (predict).
  value  device time only (CUDA events on the launch stream), inputs resident in HBM;
  e2e    the public API (model.fit / model.testForest) with host buffers: host RNG draws, H2D
         copies, kernels, D2H of scores / survivors / labels inside the timed region.
``--impl reference`` times the CPU port of the reference (oracle/rfmc_oracle.py, the same numpy
calls per node as the reference) on all host cores, one process per core, one slot per process.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REGIMES = {
    "c2b": dict(batch_train_pclass=2000, batch_val_pclass=500),
    "c2a": dict(batch_train_pclass=10, batch_val_pclass=10),
}


# ------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------
def make_frame(n_rows: int, n_feat: int = 64, n_classes: int = 4, seed: int = 0) -> pd.DataFrame:
    g = np.random.default_rng(seed)
    q = n_feat // 4
    cols = {}
    for j in range(q):
        cols[f"f64_{j}"] = np.round(g.normal(size=n_rows), 3)
    for j in range(q):
        cols[f"i64_{j}"] = g.integers(0, 1000, size=n_rows)
    for j in range(q):
        cols[f"f32_{j}"] = np.round(g.normal(size=n_rows), 2).astype(np.float32)
    for j in range(n_feat - 3 * q):
        levels = 3 + (j * 7) % 10
        cols[f"cat_{j}"] = g.choice(np.array([f"v{k}" for k in range(levels)], dtype=object), size=n_rows)
    df = pd.DataFrame(cols)
    score = np.zeros(n_rows)
    for j in range(0, q, 2):
        score += df[f"f64_{j}"].to_numpy() + df[f"i64_{j}"].to_numpy() / 400.0 + df[f"f32_{j}"].to_numpy()
    score += (df["cat_0"].to_numpy() == "v1") * 1.5 + g.normal(size=n_rows) * 2.0
    edges = np.quantile(score, np.linspace(0, 1, n_classes + 1)[1:-1])
    df["target"] = np.searchsorted(edges, score).astype(str)
    return df


def model_kwargs(args):
    kw = dict(n_trees=args.n_trees, target_col="target", max_discard_trees=args.max_discard_trees)
    kw.update(REGIMES[args.regime])
    return kw


def workload_name(args):
    r = REGIMES[args.regime]
    return (
        f"C2{args.regime[-1]}: synthetic {args.rows}x{args.features} mixed numeric/categorical, {args.classes} classes, "
        f"n_trees={args.n_trees}, max_discard_trees={args.max_discard_trees}, batch_train_pclass={r['batch_train_pclass']}, "
        f"batch_val_pclass={r['batch_val_pclass']}; predict {args.predict_rows} rows x {args.n_trees} trees"
    )


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        top = sorted(sm)[len(sm) // 2 :] if sm else []
        return {
            "sm_mhz": (sorted(top)[len(top) // 2] if top else None),
            "sm_max_mhz": mx,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU port on all host cores
# ------------------------------------------------------------------------------------------------
_REF = {}
_emit = print


def _ref_worker(task):
    from oracle import rfmc_oracle as orc

    seed, policy = task
    np.random.seed(seed)
    random.seed(seed)
    t0 = time.perf_counter()
    _, _, grown = orc.select_tree(_REF["ds"], **policy)
    return grown, time.perf_counter() - t0


def reference_fit_rate(frame, args, steps, warmup, cores):
    """candidate-trees/s of the CPU port: `cores` processes, one tree slot each per step."""
    import multiprocessing as mp

    from oracle import rfmc_oracle as orc

    kw = model_kwargs(args)
    ds = orc.ingest(frame, target_col="target", batch_train_pclass=kw["batch_train_pclass"], batch_val_pclass=kw["batch_val_pclass"])
    _REF["ds"] = ds
    policy = dict(max_discard_trees=kw["max_discard_trees"])
    ctx = mp.get_context("fork")
    times, cands = [], []
    with ctx.Pool(cores) as pool:
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            out = pool.map(_ref_worker, [(10_000 + s * cores + c, policy) for c in range(cores)], chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                cands.append(sum(o[0] for o in out))
    return sum(cands) / sum(times), sum(times) / len(times) * 1e3, sum(cands) / len(cands)


def reference_predict_rate(frame, args, trees, scores, class_vals, rows=256):
    """rows*trees/s of the CPU port's testForest on a small slice (one process, like the reference)."""
    from oracle import rfmc_oracle as orc

    part = frame.drop(columns=["target"]).head(rows)
    t0 = time.perf_counter()
    orc.forest_labels(trees, scores, class_vals, part, False, False)
    dt = time.perf_counter() - t0
    return rows * len(trees) / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = args.ref_cores or os.cpu_count() or 1
    frame = make_frame(args.rows, args.features, args.classes, seed=0)
    rate, ms, per_step = reference_fit_rate(frame, args, args.steps, args.warmup, cores)
    line = {
        "impl": "reference",
        "metric": "fit_candidate_trees_per_s",
        "value": rate,
        "unit": "candidate-trees/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args)},
        "cpu_baseline": {
            "value": rate,
            "unit": "candidate-trees/s",
            "cores": cores,
            "kind": "port",
            "sample": f"{cores} processes x 1 tree slot ({per_step:.0f} candidates) per step; oracle/rfmc_oracle.py = the reference's numpy calls per node",
        },
        "e2e": {"value": rate, "unit": "candidate-trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, workload):
    """dram bytes per launch from the committed ncu capture -- only for the workload it was taken on."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if d.get("workload") == workload:
            return d.get("bytes_per_launch", {}).get(kernel)
    return None


def run_b200(args):
    import torch
    import torch.distributed as td

    from random_forest_mc_b200 import dist, engine
    from random_forest_mc_b200.flat import encode_frame
    from random_forest_mc_b200.model import RandomForestMC

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    engine.require_gpu()
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    # ---- data + model (not timed: process_dataset is outside the metric) ----
    frame = make_frame(args.rows, args.features, args.classes, seed=0)
    kw = model_kwargs(args)
    model = RandomForestMC(**kw)
    model.device = local
    import logging

    logging.disable(logging.WARNING)
    t0 = time.perf_counter()
    model.process_dataset(frame.copy())
    t_process = time.perf_counter() - t0
    enc = model.encoded
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush.zero_()

    # ---- fit, device-only: a fixed plan (drawn once with the reference's RNGs) resident in HBM ----
    np.random.seed(1234 + rank)
    random.seed(1234 + rank)
    K = kw["max_discard_trees"]
    plan = [model._draw_slot(K) for _ in range(args.n_trees)]
    batch, _ = None, None
    idx_T = np.stack([p[0] for p in plan])
    idx_V = np.stack([p[1] for p in plan])
    slot_of, feats = [], []
    for s, p in enumerate(plan):
        for f in p[2]:
            slot_of.append(s)
            feats.append(f)
    batch = engine.GrowBatch(model._dev_dataset, idx_T, idx_V, slot_of, feats, model.max_depth, model.min_samples_split, stream)
    n_cand = batch.n_cand
    for _ in range(args.warmup):
        batch.run(stream)
    torch.cuda.synchronize()
    correct, n_nodes = batch.results()
    sampler = ClockSampler(local).start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    batch.set_timing(True)  # events between the two kernels of a run (slot gather, grower), on the launching stream
    k_ms = []
    for a, b in ev:
        flush_l2()
        a.record(stream)
        batch.run(stream)
        b.record(stream)
        launches += batch.launches
        k_ms.append(batch.kernel_ms())
    batch.set_timing(False)
    barrier()
    fit_ms = [a.elapsed_time(b) for a, b in ev]
    gather_ms = float(np.mean([g for g, _ in k_ms]))
    grow_ms = float(np.mean([k for _, k in k_ms]))
    fit_total_s = max_over_ranks(sum(fit_ms) / 1e3)
    fit_value = world * n_cand * args.steps / fit_total_s
    # algorithmic HBM bytes of the grower per launch (DESIGN.md "roofline"): every sampled row's
    # codes for the candidate's features + its row id and label, plus the node table written
    w = enc.width
    C = enc.n_classes
    nf = np.array([len(f) for f in feats], dtype=np.float64)
    algo_bytes = float(((batch.n_train + batch.n_val) * (nf * w + 4 + 1)).sum() + (n_nodes.astype(np.float64) * (24 + 4 * C + 1)).sum())
    hbm_peak, peak_src = peaks()
    grow_gbs = algo_bytes / (grow_ms / 1e3) / 1e9
    # slot gather: every bootstrap row read once (its row id, codes and label) and written once, transposed
    gather_bytes = float(args.n_trees * batch.n_train * (4 + 2 * (args.features * w + 1)))
    gather_gbs = gather_bytes / (gather_ms / 1e3) / 1e9
    total_nodes = int(n_nodes.sum())

    # ---- fit, end to end through the public API ----
    # Two numbers.  `e2e` (headline) is fitParallel: the slots shared over k independent host RNG
    # streams per rank -- the semantics of the reference's process pool and of the CPU arm, which
    # also runs one independent stream per core.  `e2e_fit_serial` is fit(): ONE stream, every draw
    # in the serial reference's order (its floor is the host walk over that stream).
    cores = os.cpu_count() or 1
    k_streams = max(1, min(args.host_streams or max(1, min(16, cores // world)), cores, args.n_trees))

    def e2e_fit(k):
        model.data, model.survived_scores = [], []
        np.random.seed(4321 + rank)
        random.seed(4321 + rank)
        seeds = [9000 + 100 * rank + r for r in range(k)]
        torch.cuda.synchronize()
        t = time.perf_counter()
        if world > 1:
            forest = dist.fit_sharded_local(model, args.n_trees, k, seeds)
        elif k > 1:
            forest = model._fit_slots_streams(args.n_trees, seeds)
        else:
            forest = model._fit_slots(args.n_trees, stream)
        torch.cuda.synchronize()
        return time.perf_counter() - t, forest

    def e2e_run(k):
        e2e_fit(k)  # warm-up: device block cache, pinned staging pool, CUDA streams of the host threads
        e2e_fit(k)
        barrier()
        total, cands, each, forest = 0.0, 0, [], None
        for _ in range(args.e2e_steps):
            dt, forest = e2e_fit(k)
            total += dt
            each.append(round(dt * 1e3, 2))
            print("e2e step", k, "streams", round(dt * 1e3, 1), "ms", {q: (round(v, 4) if isinstance(v, float) else v) for q, v in model.fit_stats.items()}, file=sys.stderr)
            cands += model.fit_stats["candidates_grown"]
        barrier()
        return max_over_ranks(total), cands, each, forest, dict(model.fit_stats)

    ser_s, ser_cands, ser_each, forest, ser_stats = e2e_run(1)
    if k_streams > 1:
        e2e_s, e2e_cands, e2e_each, forest, stats = e2e_run(k_streams)
    else:
        e2e_s, e2e_cands, e2e_each, stats = ser_s, ser_cands, ser_each, ser_stats
    surv_nodes = sum(t.soa.n_nodes for t in forest[: args.n_trees])
    h2d = int(args.n_trees * (batch.n_train + batch.n_val) * 4 + sum(len(f) for f in feats) * 2 + n_cand * 8)
    d2h = int(n_cand * 8 + surv_nodes * (24 + 4 * C))
    e2e_fit_value = world * e2e_cands / e2e_s

    # ---- predict: vote predict_rows rows through the fitted forest ----
    model.data = list(forest[: args.n_trees])
    model.survived_scores = [t.survived_score for t in model.data]
    dev = model._device_forest()
    tables = model._forest_tables
    pframe = frame.drop(columns=["target"]).head(args.predict_rows)
    t0 = time.perf_counter()
    codes, absent = encode_frame(tables, pframe)
    t_encode = time.perf_counter() - t0
    codes_t = torch.from_numpy(codes).cuda()
    labels_t = torch.empty(len(codes), dtype=torch.int32, device="cuda")
    R, T = len(codes), tables.n_trees

    def pred_dev():
        dev.predict_device(codes_t.data_ptr(), tables.width, R, tables.stride, None, model.soft_voting, model.weighted_tree, 0, labels_t.data_ptr(), stream)

    for _ in range(args.warmup):
        pred_dev()
    barrier()
    pev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in pev:
        flush_l2()
        a.record(stream)
        pred_dev()
        b.record(stream)
    barrier()
    clocks = sampler.stop()
    pred_ms = [a.elapsed_time(b) for a, b in pev]
    pred_total_s = max_over_ranks(sum(pred_ms) / 1e3)
    pred_value = world * R * T * args.steps / pred_total_s
    pred_algo = float(R * (args.features * tables.width + 4) + len(tables.pnodes) * 16 + tables.probs.nbytes * 2)
    pred_gbs = pred_algo / (np.mean(pred_ms) / 1e3) / 1e9
    # e2e predict: host codes -> H2D -> kernel -> D2H labels (through DeviceForest.predict)
    dev.predict(codes, absent, False, False, False, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, lab = dev.predict(codes, absent, False, False, False, True)
    torch.cuda.synchronize()
    pe2e_s = max_over_ranks(time.perf_counter() - t0)
    pred_e2e = world * R * T * args.steps / pe2e_s
    # the public call on a RAW frame: model.testForest(df) = host categorical mapping + pinned staging
    # + GPU encode + vote + D2H labels
    model.testForest(pframe.head(4096))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lab_api = model.testForest(pframe)
    torch.cuda.synchronize()
    frame_s = max_over_ranks(time.perf_counter() - t0)
    frame_e2e = world * R * T / frame_s
    raw_bytes = int(sum(pframe[c].to_numpy().nbytes for c in pframe.columns if pframe[c].dtype != object))

    # ---- optional: BASELINE config 5, one GPU's shard (12.5M of the 100M rows) x 4096 merged trees ----
    c5 = None
    if args.c5_rows > 0:
        m5 = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=args.max_discard_trees)  # C2a defaults
        m5.device = local
        m5.process_dataset(frame.copy())
        for seed in range(16):  # 16 fits appended == 16 models merged with mergeForest(by="add") (model.py:369, forest.py:123)
            np.random.seed(seed)
            random.seed(seed)
            m5.fit(disable_progress_bar=True)
        m5.setSoftVoting(True)
        m5.setWeightedTrees(True)
        dev5 = m5._device_forest()
        t5 = m5._forest_tables
        R5 = args.c5_rows
        top = torch.tensor([c.n_unique + 1 for c in m5.encoded.columns] + [1] * (t5.stride - len(m5.encoded.columns)), device="cuda", dtype=torch.int32)
        assert int(top.max()) < 32768 and t5.width == 2
        codes5 = torch.empty((R5, t5.stride), dtype=torch.int16, device="cuda")
        g5 = torch.Generator(device="cuda")
        g5.manual_seed(5 + rank)
        for lo in range(0, R5, 1 << 20):  # rows drawn on the device, code 0 (missing) included
            hi = min(R5, lo + (1 << 20))
            codes5[lo:hi] = (torch.randint(0, 1 << 30, (hi - lo, t5.stride), device="cuda", dtype=torch.int32, generator=g5) % top).to(torch.int16)
        probs5 = torch.empty((R5, C), dtype=torch.float64, device="cuda")

        def pred5():
            dev5.predict_device(codes5.data_ptr(), t5.width, R5, t5.stride, None, True, True, probs5.data_ptr(), 0, stream)

        for _ in range(args.warmup):
            pred5()
        barrier()
        ev5 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in ev5:
            a.record(stream)
            pred5()
            b.record(stream)
        barrier()
        ms5 = [a.elapsed_time(b) for a, b in ev5]
        s5 = max_over_ranks(sum(ms5) / 1e3)
        algo5 = float(R5 * (args.features * t5.width + 8 * C) + len(t5.pnodes) * 16 + t5.probs.nbytes * 2)
        sums = probs5[: 1 << 16].sum(dim=1)
        assert bool(((sums - 1.0).abs() < 1e-12).all())
        c5 = {
            "workload": f"C5 shard: {R5} rows x {t5.n_trees} trees (16 fits x 256 slots at the default 10+10 rows per class, appended = mergeForest by add), weighted soft voting, float64 probabilities out",
            "value": world * R5 * t5.n_trees * args.steps / s5,
            "unit": "rows*trees/s",
            "ms_per_step": float(np.mean(ms5)),
            "forest_nodes": int(len(t5.pnodes)),
            "data": "code rows drawn on the device (torch.randint per column, missing code included); inputs (1.65 GB) larger than L2",
            "roofline": {"kernel": "rfmc::predict_kernel", "bound": "hbm", "achieved": algo5 / (np.mean(ms5) / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": algo5 / (np.mean(ms5) / 1e3) / 1e9 / hbm_peak, "traffic": None, "algorithmic_bytes_per_launch": algo5},
        }
        del codes5, probs5
        dev5.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = args.ref_cores or os.cpu_count() or 1
        rate, ms, per_step = reference_fit_rate(frame.copy(), args, 1 if args.regime == "c2b" else 3, 0, cores)
        from oracle import rfmc_oracle as orc

        small = [t.data for t in model.data[:8]]
        prate = reference_predict_rate(frame, args, small, model.survived_scores[:8], model.class_vals, rows=64)
        cpu_baseline = {
            "value": rate,
            "unit": "candidate-trees/s",
            "cores": cores,
            "kind": "port",
            "sample": f"{cores} processes x 1 tree slot ({per_step:.0f} candidates) per round of the same workload; oracle/rfmc_oracle.py (the reference's numpy calls per node)",
            "predict_rows_trees_per_s": prate,
            "predict_sample": "testForest port, 64 rows x 8 trees of the fitted forest, 1 process (the reference's row loop is serial)",
        }

    if rank == 0:
        line = {
            "metric": "fit_candidate_trees_per_s",
            "value": fit_value,
            "unit": "candidate-trees/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": float(np.mean(fit_ms)),
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": {1: "u8", 2: "u16", 4: "u32"}[w],
            "data": "synthetic",
            "config": {
                "workload": workload_name(args),
                "candidates_per_step": n_cand,
                "nodes_per_step": total_nodes,
                "l2": "flushed between timed iterations (256 MiB memset)",
                "parallelism": f"slots sharded over {world} rank(s), matrix replicated, no data-path collective",
            },
            "clocks": clocks,
            "e2e": {
                "value": e2e_fit_value,
                "unit": "candidate-trees/s",
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "api": f"RandomForestMC.fitParallel: {k_streams} independent seeded host RNG stream(s) per rank (the semantics of the "
                "reference's process pool and of the CPU arm): host RNG draws + H2D + grow/score kernel + D2H scores + survivor tables"
                + ("; + all-gather of survivors" if world > 1 else ""),
                "host_streams_per_rank": k_streams,
                "ms_per_step": e2e_s / args.e2e_steps * 1e3,
                "ms_each_rank0": e2e_each,
                "fit_stats": stats,
            },
            "e2e_fit_serial": {
                "value": world * ser_cands / ser_s,
                "unit": "candidate-trees/s",
                "api": "RandomForestMC.fit loop (_fit_slots): ONE RNG stream per rank, every draw in the serial reference's order"
                + ("; + all-gather of survivors" if world > 1 else ""),
                "ms_per_step": ser_s / args.e2e_steps * 1e3,
                "ms_each_rank0": ser_each,
                "fit_stats": ser_stats,
            },
            "gpu_launches": int(launches) + args.steps,
            "roofline": {
                "kernel": "rfmc::grow_kernel",
                "bound": "hbm",
                "achieved": grow_gbs,
                "peak": hbm_peak,
                "unit": "GB/s",
                "frac": grow_gbs / hbm_peak,
                "traffic": ncu_traffic("grow_kernel", workload_name(args)),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "ms_per_launch": grow_ms,
                "share_of_step": grow_ms / float(np.mean(fit_ms)),
            },
            "roofline_slot_gather": {
                "kernel": "rfmc::slot_gather_kernel",
                "bound": "hbm",
                "achieved": gather_gbs,
                "peak": hbm_peak,
                "unit": "GB/s",
                "frac": gather_gbs / hbm_peak,
                "traffic": ncu_traffic("slot_gather_kernel", workload_name(args)),
                "algorithmic_bytes_per_launch": gather_bytes,
                "ms_per_launch": gather_ms,
                "share_of_step": gather_ms / float(np.mean(fit_ms)),
            },
            "predict": {
                "metric": "predict_rows_trees_per_s",
                "value": pred_value,
                "unit": "rows*trees/s",
                "ms_per_step": float(np.mean(pred_ms)),
                "rows": R,
                "trees": T,
                "forest_nodes": int(len(tables.pnodes)),
                "e2e": {
                    "value": pred_e2e,
                    "unit": "rows*trees/s",
                    "h2d_bytes_per_step": int(codes.nbytes),
                    "d2h_bytes_per_step": int(R * 4),
                    "api": "DeviceForest.predict on already-encoded rows (host codes -> H2D -> vote kernel -> D2H labels)",
                },
                "frame_api": {
                    "value": frame_e2e,
                    "unit": "rows*trees/s",
                    "seconds": frame_s,
                    "h2d_bytes_per_step": raw_bytes,
                    "api": "model.testForest(DataFrame): host categorical mapping + pinned staging + GPU encode + vote + D2H labels",
                },
                "host_encode_s": t_encode,
                "roofline": {
                    "kernel": "rfmc::predict_kernel",
                    "bound": "hbm",
                    "achieved": pred_gbs,
                    "peak": hbm_peak,
                    "unit": "GB/s",
                    "frac": pred_gbs / hbm_peak,
                    "traffic": ncu_traffic("predict_kernel", workload_name(args)),
                    "algorithmic_bytes_per_launch": pred_algo,
                },
            },
            "process_dataset_s": t_process,
            "process_dataset": {
                "seconds": t_process,
                "api": "RandomForestMC.process_dataset(DataFrame): pandas dropna/typing + strings ranked in Arrow on the host + raw numeric columns H2D + device rank coding (sort, distinct, rank, interleave)",
                "h2d_bytes": int(sum(model.dataset_numpy[c].nbytes for c in model.numeric_cols) + 4 * enc.n_rows * (len(model.feature_cols) - len(model.numeric_cols)) + enc.n_rows),
            },
            **({"predict_c5": c5} if c5 is not None else {}),
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        _emit(json.dumps(line))
    batch.close()
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--regime", default="c2b", choices=list(REGIMES))
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--features", type=int, default=64)
    ap.add_argument("--classes", type=int, default=4)
    ap.add_argument("--n-trees", type=int, default=256, dest="n_trees")
    ap.add_argument("--max-discard-trees", type=int, default=4, dest="max_discard_trees")
    ap.add_argument("--predict-rows", type=int, default=1_000_000, dest="predict_rows")
    ap.add_argument("--e2e-steps", type=int, default=4, dest="e2e_steps")
    ap.add_argument("--ref-cores", type=int, default=0, dest="ref_cores")
    ap.add_argument("--host-streams", type=int, default=0, dest="host_streams", help="host RNG streams per rank for the e2e fit (0 = min(16, cores / ranks))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c5-rows", type=int, default=0, dest="c5_rows", help="also vote this many device-generated rows through a 4096-tree merged forest (BASELINE config 5; 12500000 = one GPU's shard)")
    args = ap.parse_args()
    args.predict_rows = min(args.predict_rows, args.rows)
    # stdout carries exactly ONE JSON line: anything a library prints (NCCL's version banner ...)
    # goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit
    _emit = lambda text: os.write(real_stdout, (text + "\n").encode())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
