"""r2b probe (one process per grower variant, env knobs read at first launch):
  python profiles/r2b_probe.py grow            -> gather / grower ms of 1,024 C2b candidates (CUDA events, L2 flushed)
  python profiles/r2b_probe.py streams 12 16.. -> fitParallel(k) step times for the host-stream counts given"""
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from random_forest_mc_b200 import engine  # noqa: E402
from random_forest_mc_b200.model import RandomForestMC  # noqa: E402
import torch  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "grow"
frame = bench.make_frame(1_000_000, 64, 4, seed=0)
model = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4, **bench.REGIMES["c2b"])
model.process_dataset(frame)
K = 4
if mode == "grow":
    np.random.seed(1234)
    random.seed(1234)
    plan = [model._draw_slot(K) for _ in range(256)]
    idx_T = np.stack([p[0] for p in plan])
    idx_V = np.stack([p[1] for p in plan])
    slot_of, feats = [], []
    for s_, p in enumerate(plan):
        for f in p[2]:
            slot_of.append(s_)
            feats.append(f)
    stream = torch.cuda.current_stream().cuda_stream
    batch = engine.GrowBatch(model._dev_dataset, idx_T, idx_V, slot_of, feats, model.max_depth, model.min_samples_split, stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        batch.run(stream)
    torch.cuda.synchronize()
    correct, n_nodes = batch.results()
    batch.set_timing(True)
    ms = []
    for _ in range(8):
        flush.zero_()
        batch.run(stream)
        ms.append(batch.kernel_ms())
    tag = {k: v for k, v in os.environ.items() if k.startswith("RFMC_")}
    print(f"grow probe {tag}: gather {np.mean([g for g, _ in ms]):.4f} ms, grower {np.mean([k for _, k in ms]):.4f} ms (min {min(k for _, k in ms):.4f}), "
          f"nodes {int(n_nodes.sum())}, correct checksum {int(correct.astype(np.int64).sum())}", flush=True)
    batch.close()
else:
    ks = [int(x) for x in sys.argv[2:]] or [16]
    for k in ks:
        times = []
        for rep in range(6):
            model.reset_forest()
            np.random.seed(4321)
            random.seed(4321)
            model.worker_seeds = [9000 + w for w in range(k)]
            torch.cuda.synchronize()
            t = time.perf_counter()
            model.fitParallel(max_workers=k)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t) * 1e3)
        st = model.fit_stats
        print(f"streams {k}: steps {[round(x, 1) for x in times]} ms | best {min(times[2:]):.1f} mean(last 4) {np.mean(times[2:]):.1f} | "
              f"per-thread draw {st['t_draw'] / k * 1e3:.1f} launch {st['t_launch'] / k * 1e3:.2f} wait {st['t_wait'] / k * 1e3:.2f} fetch {st['t_fetch'] / k * 1e3:.2f}", flush=True)
