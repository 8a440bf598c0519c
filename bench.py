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
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def stored_ncu(kernel, workload, any_workload=False):
    """Figures of the committed ncu capture of `kernel` (profiles/ncu_traffic.json) -- only for the workload it
    was taken on.  They are STORED captures (ncu replays kernels; nothing is profiled inside a bench run):
    the dict names the capture and the commit it was taken at."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if any_workload or d.get("workload") == workload:
            k = d.get("kernels", {}).get(kernel)
            if k is None and kernel in d.get("bytes_per_launch", {}):
                k = {"dram_bytes_per_launch": d["bytes_per_launch"][kernel]}
            if k is not None:
                return dict(k, source="stored ncu capture", capture=k.get("capture", d.get("capture")), commit=k.get("commit", d.get("commit")))
    return None


def make_model(args, regime, n_trees, local):
    from random_forest_mc_b200.model import RandomForestMC

    kw = dict(n_trees=n_trees, target_col="target", max_discard_trees=args.max_discard_trees)
    kw.update(REGIMES[regime])
    m = RandomForestMC(**kw)
    m.device = local
    return m


def algorithmic_fit_bytes(batch, feats, n_nodes, width, n_classes):
    """per candidate (n_T + n_V) * (|F_c| * w + 4 + 1) + nodes_c * 16; and what the grower
    really writes per node today (24-byte node + 4 C counts + vote)."""
    nf = np.array([len(f) for f in feats], dtype=np.float64)
    rows = float(((batch.n_train + batch.n_val) * (nf * width + 4 + 1)).sum())
    nodes = float(n_nodes.astype(np.float64).sum())
    return rows + nodes * 16, rows + nodes * (24 + 4 * n_classes + 1)


def best_per_slot(correct, n_slots, K):
    """The candidate survivedTree keeps when no candidate of a slot meets th_start = 1.0: the first best score."""
    c = np.asarray(correct).reshape(n_slots, K)
    return (np.arange(n_slots) * K + c.argmax(axis=1)).astype(np.int32)


def run_b200(args):
    import logging

    import torch
    import torch.distributed as td

    from random_forest_mc_b200 import dist, engine
    from random_forest_mc_b200.flat import NODE_DTYPE, encode_frame

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    engine.require_gpu()
    stream = torch.cuda.current_stream()
    logging.disable(logging.WARNING)
    hbm_peak, peak_src, sm_mhz = peaks()
    smem_peak = 148 * 128 * sm_mhz * 1e6 / 1e9  # GB/s: 148 SMs x 128 B/clk
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)  # taskset / cpuset aware
    K = args.max_discard_trees

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

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return float(t.item())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush.zero_()

    # -------------------------------------------------------------------------------------------------
    # one fit configuration: device-timed launch (+ survivor exchange at N > 1), then the public API
    # -------------------------------------------------------------------------------------------------
    def device_fit(model, n_slots, steps, warmup, label):
        """`steps` launches of one plan of `n_slots` slots x K candidates resident in HBM, CUDA events on the
        launching stream.  At N > 1 the timed region also packs every slot's surviving candidate on the device
        and all-gathers the node tables over NCCL (rfmc_allgather_survivors) -- the path's one collective."""
        enc = model.encoded
        np.random.seed(1234 + rank)
        random.seed(1234 + rank)
        plan = [model._draw_slot(K) for _ in range(n_slots)]
        idx_T = np.stack([p[0] for p in plan])
        idx_V = np.stack([p[1] for p in plan])
        slot_of, feats = [], []
        for s_, p in enumerate(plan):
            for f in p[2]:
                slot_of.append(s_)
                feats.append(f)
        batch = engine.GrowBatch(model._dev_dataset, idx_T, idx_V, slot_of, feats, model.max_depth, model.min_samples_split, stream)
        n_cand = batch.n_cand
        C = enc.n_classes
        comm = dist.nccl_comm_ptr() if world > 1 else 0
        gathered_bytes, parity = 0, None

        def one_step(timed):
            nonlocal gathered_bytes
            batch.run(stream)
            if world == 1:
                return None
            correct, n_nodes = batch.results()
            ids = best_per_slot(correct, n_slots, K)
            sizes = n_nodes[ids].astype(np.int64)
            total = int(sizes.sum())
            tn = torch.empty(total * NODE_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            tc = torch.empty(total * C * 4, dtype=torch.uint8, device="cuda")
            batch.pack_to_device(ids, tn.data_ptr(), tc.data_ptr())
            g = engine.GatheredSurvivors(comm, local, tn.data_ptr(), tc.data_ptr(), sizes, correct[ids] / max(1, batch.n_val), None, C, stream)
            gathered_bytes = int(g.nodes_per_rank.sum()) * (NODE_DTYPE.itemsize + 4 * C)
            return g, ids

        for _ in range(warmup):
            out = one_step(False)
            if out is not None:
                out[0].close()
        torch.cuda.synchronize()
        correct, n_nodes = batch.results()
        barrier()
        batch.set_timing(True)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        k_ms, launches, last = [], 0, None
        for a_, b_ in ev:
            flush_l2()
            a_.record(stream)
            out = one_step(True)
            b_.record(stream)
            launches += batch.launches + (5 if world > 1 else 0)  # + pack kernel and four all-gathers
            k_ms.append(batch.kernel_ms())
            if last is not None:
                last[0].close()
            last = out
        batch.set_timing(False)
        barrier()
        ms = [a_.elapsed_time(b_) for a_, b_ in ev]
        total_s = max_over_ranks(sum(ms) / 1e3)
        if last is not None:
            # in-run parity of the exchange: (1) every rank holds the same gathered tables; (2) rank 0 grows rank
            # 1's first two slots itself, from rank 1's seed, and compares them with what came over NVLink
            g, ids = last
            h = torch.tensor([g.checksum() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
            lo_, hi_ = h.clone(), h.clone()
            td.all_reduce(lo_, op=td.ReduceOp.MIN)
            td.all_reduce(hi_, op=td.ReduceOp.MAX)
            parity = {"gathered_checksum_equal_on_all_ranks": bool(lo_.item() == hi_.item())}
            if rank == 0:
                np.random.seed(1234 + 1)
                random.seed(1234 + 1)
                p2 = [model._draw_slot(K) for _ in range(2)]
                so, ff = [], []
                for s_, p in enumerate(p2):
                    for f in p[2]:
                        so.append(s_)
                        ff.append(f)
                b2 = engine.GrowBatch(model._dev_dataset, np.stack([p[0] for p in p2]), np.stack([p[1] for p in p2]), so, ff,
                                      model.max_depth, model.min_samples_split, stream).run(stream)
                c2, _ = b2.results()
                mine = b2.fetch(best_per_slot(c2, 2, K))
                sizes1, _, _ = g.meta(1)
                nodes1, counts1 = g.tables(1)
                at, same = 0, True
                for t_, sz in zip(mine, sizes1[:2]):
                    same &= int(sz) == t_.n_nodes and np.array_equal(nodes1[at : at + int(sz)], t_.nodes) and np.array_equal(counts1[at : at + int(sz)], t_.counts)
                    at += int(sz)
                parity["rank1_first_two_slots_regrown_on_rank0_equal"] = bool(same)
                b2.close()
            g.close()
        gather_ms = float(np.mean([g_ for g_, _ in k_ms]))
        grow_ms = float(np.mean([k_ for _, k_ in k_ms]))
        algo16, algo_written = algorithmic_fit_bytes(batch, feats, n_nodes, enc.width, C)
        gather_bytes = float(n_slots * batch.n_train * (4 + 2 * (len(model.feature_cols) * enc.width + 1)))
        stored = stored_ncu("grow_kernel", label)
        res = {
            "value": world * n_cand * steps / total_s,
            "ms_per_step": float(np.mean(ms)),
            "candidates_per_step": n_cand,
            "nodes_per_step": int(n_nodes.sum()),
            "launches": launches,
            "grow_ms": grow_ms,
            "gather_ms": gather_ms,
            "survivor_exchange": None if world == 1 else {"bytes_gathered_per_step": gathered_bytes, "parity": parity,
                                                           "ms_per_step_beyond_the_kernels": float(np.mean(ms)) - grow_ms - gather_ms},
            "roofline": {
                "kernel": "rfmc::grow_kernel",
                "bound": "hbm",
                "achieved": algo16 / (grow_ms / 1e3) / 1e9,
                "peak": hbm_peak,
                "unit": "GB/s",
                "frac": algo16 / (grow_ms / 1e3) / 1e9 / hbm_peak,
                "traffic": None if stored is None else stored.get("dram_bytes_per_launch"),
                "traffic_source": None if stored is None else {k_: stored.get(k_) for k_ in ("source", "capture", "commit")},
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo16,
                "algorithmic_bytes_formula": "SURVEY 8(d): sum over candidates of (n_T + n_V) * (|F_c| * w + 4 + 1) + nodes_c * 16",
                "bytes_the_kernel_writes_per_launch": algo_written,
                "frac_on_written_bytes": algo_written / (grow_ms / 1e3) / 1e9 / hbm_peak,
                "ms_per_launch": grow_ms,
                "share_of_step": grow_ms / float(np.mean(ms)),
                "secondary": None if stored is None else {
                    "bound": "shared memory / issue slots (what binds the kernel in practice, SURVEY 8d)",
                    "shared_wavefront_bytes_per_launch": stored.get("shared_wavefront_bytes_per_launch"),
                    "shared_frac": None if not stored.get("shared_wavefront_bytes_per_launch") else
                    stored["shared_wavefront_bytes_per_launch"] / (grow_ms / 1e3) / 1e9 / smem_peak,
                    "shared_peak_gbs": smem_peak,
                    "issue_slot_utilisation": stored.get("issue_slot_utilisation"),
                    "source": stored.get("source"), "capture": stored.get("capture"), "commit": stored.get("commit"),
                },
            },
            "roofline_slot_gather": {
                "kernel": "rfmc::slot_gather_kernel", "bound": "hbm", "achieved": gather_bytes / (gather_ms / 1e3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": gather_bytes / (gather_ms / 1e3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": gather_bytes,
                "ms_per_launch": gather_ms, "share_of_step": gather_ms / float(np.mean(ms)),
            },
            "h2d_plan_bytes": int(n_slots * (batch.n_train + batch.n_val) * 4 + sum(len(f) for f in feats) * 2 + n_cand * 8),
            "h2d_idx_bytes": int(n_slots * (batch.n_train + batch.n_val) * 4),
        }
        batch.close()
        return res

    def api_fit(model, n_trees_total, k, steps, rank_seed_base):
        """The public call: model.fit() (k == 0) or model.fitParallel(max_workers=k).  At N > 1 fitParallel shards
        the slots over the ranks and all-gathers the survivors; every rank ends with the whole forest."""
        model.n_trees = n_trees_total

        def once():
            model.reset_forest()
            np.random.seed(4321 + rank)
            random.seed(4321 + rank)
            model.rank_seeds = [rank_seed_base + 100 * r for r in range(world)]
            model.worker_seeds = [rank_seed_base + 100 * rank + w for w in range(max(1, k))]
            torch.cuda.synchronize()
            t = time.perf_counter()
            if k == 0 and world == 1:
                model.fit(disable_progress_bar=True)
            else:
                model.fitParallel(disable_progress_bar=True, max_workers=max(1, k))
            torch.cuda.synchronize()
            return time.perf_counter() - t

        once()
        once()  # warm-ups: device block cache, pinned staging pool, CUDA streams of the host threads
        barrier()
        total, cands, each, stats = 0.0, 0.0, [], None
        for _ in range(steps):
            dt = once()
            total += dt
            each.append(round(dt * 1e3, 2))
            stats = dict(model.fit_stats)
            print("e2e step", "fit" if k == 0 else f"fitParallel({k})", round(dt * 1e3, 1), "ms",
                  {q: (round(v, 4) if isinstance(v, float) else v) for q, v in stats.items()}, file=sys.stderr)
            cands += sum_over_ranks(float(stats["candidates_grown"]))
        barrier()
        total = max_over_ranks(total)
        # what model2dict() / tree.data cost after the fit: the nested dicts of the kept forest are built lazily
        t0 = time.perf_counter()
        n_dict_nodes = 0
        for t_ in model.data[: min(len(model.data), 64)]:
            _ = t_.data
            n_dict_nodes += t_.soa.n_nodes if t_.soa is not None else 0
        dict_s = (time.perf_counter() - t0) * (len(model.data) / max(1, min(len(model.data), 64)))
        surv_nodes = sum((t_.soa.n_nodes if t_.soa is not None else 0) for t_ in model.data)
        host_share = (stats["t_draw"] / max(1, stats.get("host_streams", 1))) / (total / steps) if stats else None
        return {
            "value": cands / total,
            "ms_per_step": total / steps * 1e3,
            "ms_each_rank0": each,
            "fit_stats": stats,
            "host_rng_share": host_share,
            "dict_materialise_s": dict_s,
            "survivor_nodes": int(surv_nodes),
        }

    def fit_config(frame, regime, n_slots_local, label, steps, warmup, e2e_steps, streams, cpu_rounds, strong_total=None):
        model = make_model(args, regime, n_slots_local, local)
        t0 = time.perf_counter()
        model.process_dataset(frame.copy())
        t_process = time.perf_counter() - t0
        # With several ranks on one box a rank has only its share of the host cores for the walk of numpy's stream; below
        # six cores per rank the model draws with the DEVICE walk of the same stream (model.device_draws_on(): same rows,
        # counts and stream states; 2 ranks x 4 cores: 61.9 against 85.8 ms per step, profiles/r2b_experiments.md).
        # RFMC_DEVICE_DRAWS=0/1 overrides the model's choice.
        use_dev = model.device_draws_on()
        if use_dev and not args.host_streams:
            streams = min(streams, 4)  # four draw plans in flight per GPU: more only queue behind each other's kernels
        dev = device_fit(model, n_slots_local, steps, warmup, label)
        total_slots = strong_total if strong_total is not None else n_slots_local * world
        par = api_fit(model, total_slots, streams, e2e_steps, 9000)
        ser = api_fit(model, total_slots, 0 if world == 1 else 1, max(1, e2e_steps // 2), 9000)
        C = model.encoded.n_classes
        d2h = int(par["fit_stats"]["candidates_grown"] * 8 + par["survivor_nodes"] * (24 + 4 * C))
        block = {
            "workload": label,
            "metric": "fit_candidate_trees_per_s",
            "unit": "candidate-trees/s",
            "value": dev["value"],
            "ms_per_step": dev["ms_per_step"],
            "scaling": "strong" if strong_total is not None else "weak",
            "candidates_per_step": dev["candidates_per_step"],
            "nodes_per_step": dev["nodes_per_step"],
            "survivor_exchange": dev["survivor_exchange"],
            "roofline": dev["roofline"],
            "roofline_slot_gather": dev["roofline_slot_gather"],
            "e2e": {
                "value": par["value"], "unit": "candidate-trees/s", "h2d_bytes_per_step": (dev["h2d_plan_bytes"] - (dev["h2d_idx_bytes"] if use_dev else 0)) * max(1, total_slots // max(1, n_slots_local * world)),
                "d2h_bytes_per_step": d2h,
                "api": f"RandomForestMC.fitParallel(max_workers={streams}): " + (f"slots sharded over {world} ranks, " if world > 1 else "") +
                f"{streams} independent seeded RNG stream(s) per rank (the semantics of the reference's process pool and of the CPU arm): " +
                ("numpy's stream walked on the DEVICE (rfmc_draw_plan_device, row ids stay in HBM) + " if use_dev else "host RNG draws + H2D + ") +
                "grow/score kernels + D2H scores + survivor tables" + ("; survivors all-gathered over NCCL (rfmc_allgather_survivors)" if world > 1 else ""),
                "draws": "device" if use_dev else "host",
                "host_streams_per_rank": streams, "ms_per_step": par["ms_per_step"], "ms_each_rank0": par["ms_each_rank0"], "fit_stats": par["fit_stats"],
                "host_rng_share": par["host_rng_share"], "dict_materialise_s": par["dict_materialise_s"],
                "dict_materialise_note": "trees stay SoA after a fit; tree.data / model2dict() build the reference's nested dicts on first access, outside the timed fit",
            },
            "e2e_fit_serial": {
                "value": ser["value"], "unit": "candidate-trees/s",
                "api": "RandomForestMC.fit(): ONE RNG stream, every draw in the serial reference's order" if world == 1 else
                "RandomForestMC.fitParallel(max_workers=1): one RNG stream per rank",
                "ms_per_step": ser["ms_per_step"], "ms_each_rank0": ser["ms_each_rank0"], "fit_stats": ser["fit_stats"], "host_rng_share": ser["host_rng_share"],
            },
            "process_dataset_s": t_process,
            "gpu_launches": dev["launches"],
        }
        if rank == 0 and world == 1 and not args.no_cpu_baseline and cpu_rounds > 0:
            ncores = args.ref_cores or cores
            saved = args.regime
            args.regime = regime
            rate, ms_, per_step = reference_fit_rate(frame, args, cpu_rounds, 0, ncores)
            args.regime = saved
            block["cpu_baseline"] = {
                "value": rate, "unit": "candidate-trees/s", "cores": ncores, "kind": "port",
                "sample": f"{cpu_rounds} round(s) of {ncores} processes x 1 tree slot ({per_step:.0f} candidates per round) of the same workload; oracle/rfmc_oracle.py (the reference's numpy calls per node)",
            }
        return model, block

    # ---- C2b (default) / C2a: the headline configuration ----
    frame = make_frame(args.rows, args.features, args.classes, seed=0)
    # host RNG streams per rank: one per core at N = 1; with several ranks on the box twice the rank's share of the
    # cores (the streams of a rank alternate between drawing and waiting for their launch)
    k_streams = max(1, min(args.host_streams or max(1, min(16, cores if world == 1 else max(4, 2 * cores // world))), cores, args.n_trees))
    sampler = ClockSampler(local).start()
    model, main = fit_config(frame, args.regime, args.n_trees, workload_name(args), args.steps, args.warmup, args.e2e_steps, k_streams,
                             1 if args.regime == "c2b" else 3)
    enc = model.encoded
    w, C = enc.width, enc.n_classes

    # ---- predict: vote predict_rows rows through the fitted forest (n_trees of it: at N > 1 every rank holds all ranks' trees) ----
    model.data = model.data[: args.n_trees]
    model.survived_scores = model.survived_scores[: args.n_trees]
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
    for a_, b_ in pev:
        flush_l2()
        a_.record(stream)
        pred_dev()
        b_.record(stream)
    barrier()
    clocks = sampler.stop()
    pred_ms = [a_.elapsed_time(b_) for a_, b_ in pev]
    pred_total_s = max_over_ranks(sum(pred_ms) / 1e3)
    pred_value = world * R * T * args.steps / pred_total_s
    pred_algo = float(R * (args.features * tables.width + 4) + len(tables.pnodes) * 16 + tables.probs.nbytes * 2)
    pred_gbs = pred_algo / (np.mean(pred_ms) / 1e3) / 1e9
    dev.predict(codes, absent, False, False, False, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, lab = dev.predict(codes, absent, False, False, False, True)
    torch.cuda.synchronize()
    pe2e_s = max_over_ranks(time.perf_counter() - t0)
    pred_e2e = world * R * T * args.steps / pe2e_s
    model.testForest(pframe.head(4096))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.testForest(pframe)
    torch.cuda.synchronize()
    frame_s = max_over_ranks(time.perf_counter() - t0)
    frame_e2e = world * R * T / frame_s
    raw_bytes = int(sum(pframe[c].to_numpy().nbytes for c in pframe.columns if pframe[c].dtype != object))
    # mean root-to-leaf path of the voted rows: the shared-memory bound of SURVEY 8(d) is rows * trees * path * 8 bytes
    sample_payload, _ = dev.leaves(codes[:2048], absent)
    stored_p = stored_ncu("predict_kernel", workload_name(args))
    predict_block = {
        "metric": "predict_rows_trees_per_s", "value": pred_value, "unit": "rows*trees/s", "ms_per_step": float(np.mean(pred_ms)),
        "rows": R, "trees": T, "forest_nodes": int(len(tables.pnodes)),
        "e2e": {"value": pred_e2e, "unit": "rows*trees/s", "h2d_bytes_per_step": int(codes.nbytes), "d2h_bytes_per_step": int(R * 4),
                "api": "DeviceForest.predict on already-encoded rows (host codes -> H2D -> vote kernel -> D2H labels)"},
        "frame_api": {"value": frame_e2e, "unit": "rows*trees/s", "seconds": frame_s, "h2d_bytes_per_step": raw_bytes,
                      "api": "model.testForest(DataFrame): host categorical mapping + pinned staging + GPU encode + vote + D2H labels"},
        "host_encode_s": t_encode,
        "roofline": {"kernel": dev.vote_kernel, "bound": "hbm", "achieved": pred_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": pred_gbs / hbm_peak,
                     "traffic": None if stored_p is None else stored_p.get("dram_bytes_per_launch"),
                     "traffic_source": None if stored_p is None else {k_: stored_p.get(k_) for k_ in ("source", "capture", "commit")},
                     "algorithmic_bytes_per_launch": pred_algo,
                     "secondary": {"bound": "dependent node loads (SURVEY 8d: rows * trees * path_len * 8 B against 148 SM x 128 B/clk)",
                                   "shared_peak_gbs": smem_peak, "rows_trees_per_s_at_that_peak_for_path_7": smem_peak * 1e9 / (7 * 8),
                                   "frac_of_that_bound": pred_value / world / (smem_peak * 1e9 / (7 * 8))}},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        small = [t_.data for t_ in model.data[:8]]
        predict_block["cpu_baseline"] = {
            "value": reference_predict_rate(frame, args, small, model.survived_scores[:8], model.class_vals, rows=64), "unit": "rows*trees/s", "cores": 1,
            "kind": "port", "sample": "testForest port, 64 rows x 8 trees of the fitted forest, 1 process (the reference's row loop is serial)"}

    # ---- the other named configurations, as sub-objects ----
    extra = {}
    if not args.no_c2a and args.regime != "c2a":
        saved = args.regime
        args.regime = "c2a"
        label = workload_name(args)
        args.regime = saved
        _, extra["c2a"] = fit_config(frame, "c2a", args.n_trees, label, args.steps, args.warmup, args.e2e_steps, k_streams, 3)
    if args.c5_rows > 0:
        extra["c5"] = predict_c5(args, frame, local, rank, world, stream, barrier, max_over_ranks, hbm_peak, smem_peak)
    del frame, codes_t
    if args.c4_rows > 0:
        # BASELINE config 4: 8,192 candidates (2,048 slots x 4) on 10M x 128, the slots split over the ranks (strong scaling)
        frame4 = make_frame(args.c4_rows, 128, args.classes, seed=1)
        total_slots = args.c4_slots
        n_local = max(1, total_slots // world)
        label4 = (f"C4: synthetic {args.c4_rows}x128 mixed numeric/categorical, {args.classes} classes, {total_slots} slots x {K} candidates "
                  f"({total_slots * K} candidate trees) split over {world} rank(s), batch_train_pclass=2000, batch_val_pclass=500; device-timed value on one launch "
                  f"of {min(n_local, 256)} slots per rank")
        saved_trees = args.n_trees
        _, extra["c4"] = fit_config(frame4, "c2b", min(n_local, 256), label4, max(1, args.steps), max(1, args.warmup - 1), 2, k_streams, 1, strong_total=total_slots)
        args.n_trees = saved_trees
        del frame4

    if rank == 0:
        line = {
            "metric": "fit_candidate_trees_per_s",
            "value": main["value"],
            "unit": "candidate-trees/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"],
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": {1: "u8", 2: "u16", 4: "u32"}[w],
            "data": "synthetic",
            "config": {"workload": workload_name(args)},
            "details": {
                "candidates_per_step": main["candidates_per_step"],
                "nodes_per_step": main["nodes_per_step"],
                "l2": "flushed between timed iterations (256 MiB memset)",
                "parallelism": f"slots sharded over {world} rank(s), matrix replicated" + ("; the timed step ends with the survivor all-gather over NCCL" if world > 1 else ""),
                "survivor_exchange": main["survivor_exchange"],
            },
            "clocks": clocks,
            "e2e": main["e2e"],
            "e2e_fit_serial": main["e2e_fit_serial"],
            "gpu_launches": int(main["gpu_launches"]) + args.steps,
            "roofline": main["roofline"],
            "roofline_slot_gather": main["roofline_slot_gather"],
            "predict": predict_block,
            "process_dataset_s": main["process_dataset_s"],
            **extra,
        }
        if "cpu_baseline" in main:
            line["cpu_baseline"] = dict(main["cpu_baseline"])
            if "cpu_baseline" in predict_block:
                line["cpu_baseline"]["predict_rows_trees_per_s"] = predict_block["cpu_baseline"]["value"]
                line["cpu_baseline"]["predict_sample"] = predict_block["cpu_baseline"]["sample"]
        _emit(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def predict_c5(args, frame, local, rank, world, stream, barrier, max_over_ranks, hbm_peak, smem_peak):
    """BASELINE config 5, one GPU's shard: c5_rows device-generated rows x 4,096 merged trees, weighted soft voting."""
    import torch

    m5 = make_model(args, "c2a", 256, local)
    m5.process_dataset(frame.copy())
    for seed in range(16):  # 16 fits appended == 16 models merged with mergeForest(by="add") (model.py:369, forest.py:123)
        np.random.seed(seed)
        random.seed(seed)
        m5.fit(disable_progress_bar=True)
    m5.setSoftVoting(True)
    m5.setWeightedTrees(True)
    dev5 = m5._device_forest()
    t5 = m5._forest_tables
    C = m5.encoded.n_classes
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
    for a_, b_ in ev5:
        a_.record(stream)
        pred5()
        b_.record(stream)
    barrier()
    ms5 = [a_.elapsed_time(b_) for a_, b_ in ev5]
    s5 = max_over_ranks(sum(ms5) / 1e3)
    algo5 = float(R5 * (args.features * t5.width + 8 * C) + len(t5.pnodes) * 16 + t5.probs.nbytes * 2)
    sums = probs5[: 1 << 16].sum(dim=1)
    assert bool(((sums - 1.0).abs() < 1e-12).all())
    # end to end for the same shard: host code rows -> pinned chunks -> vote -> D2H probabilities (DeviceForest.predict)
    n_e2e = min(R5, 1 << 20)
    host_codes = codes5[:n_e2e].cpu().numpy().view(np.uint16)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev5.predict(host_codes, None, True, True, True, False)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    value = world * R5 * t5.n_trees * args.steps / s5
    stored_small = stored_ncu(dev5.vote_kernel.split("::")[-1], None, any_workload=True)  # taken on 1,048,576 rows of this forest, not on R5
    block = {
        "workload": f"C5 shard: {R5} rows x {t5.n_trees} trees per GPU (16 fits x 256 slots at the default 10+10 rows per class, appended = mergeForest by add), weighted soft voting, float64 probabilities out",
        "metric": "predict_rows_trees_per_s", "value": value, "unit": "rows*trees/s", "ms_per_step": float(np.mean(ms5)), "forest_nodes": int(len(t5.pnodes)),
        "scaling": "weak", "data": "code rows drawn on the device (torch.randint per column, missing code included); inputs (1.65 GB per GPU) larger than L2",
        "e2e": {"value": world * n_e2e * t5.n_trees / e2e_s, "unit": "rows*trees/s", "h2d_bytes_per_step": int(host_codes.nbytes), "d2h_bytes_per_step": int(n_e2e * C * 8),
                "api": f"DeviceForest.predict on {n_e2e} host code rows of the shard: pinned double-buffered chunks, vote kernel, float64 probabilities back"},
        "roofline": {"kernel": dev5.vote_kernel, "bound": "hbm", "achieved": algo5 / (np.mean(ms5) / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": algo5 / (np.mean(ms5) / 1e3) / 1e9 / hbm_peak, "traffic": None, "traffic_stored_capture": stored_small,
                     "algorithmic_bytes_per_launch": algo5,
                     "secondary": {"bound": "shared-memory wavefronts of the walk (SURVEY 8d): per tree level one 8-byte node record + one 2-byte row code, "
                                            "both from shared memory in predict_small_kernel (bank conflicts of 32 lanes in 32 different records are what it pays)",
                                   "shared_peak_gbs": smem_peak, "small_tree_stages": dev5.small_tree_stages,
                                   "frac_of_that_bound_for_path_7": value / world / (smem_peak * 1e9 / (7 * 8)),
                                   "frac_of_that_bound_for_path_7_incl_row_codes": value / world / (smem_peak * 1e9 / (7 * 10))}},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        small = [t_.data for t_ in m5.data[:16]]
        block["cpu_baseline"] = {"value": reference_predict_rate(frame, args, small, m5.survived_scores[:16], m5.class_vals, rows=128), "unit": "rows*trees/s",
                                 "cores": 1, "kind": "port", "sample": "testForest port, 128 rows x 16 trees of the merged forest, 1 process"}
    del codes5, probs5
    dev5.close()
    return block


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
    ap.add_argument("--c5-rows", type=int, default=12_500_000, dest="c5_rows", help="BASELINE config 5: vote this many device-generated rows per GPU through a 4096-tree merged forest (12500000 = one GPU's shard of the 100M rows; 0 = skip)")
    ap.add_argument("--c4-rows", type=int, default=10_000_000, dest="c4_rows", help="BASELINE config 4: rows of the 128-feature frame whose 2,048 slots are split over the ranks (0 = skip)")
    ap.add_argument("--c4-slots", type=int, default=2048, dest="c4_slots")
    ap.add_argument("--no-c2a", action="store_true", dest="no_c2a")
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
