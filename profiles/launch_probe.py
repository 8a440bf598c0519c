"""Where does the per-stream launch time of a 16-stream fit go?  16 host threads, each with its own CUDA stream and a
fixed 16-slot x 4-candidate plan (C2b), timing GrowBatch create / run / results / fetch / close per thread."""
import logging
import random
import sys
import threading
import time

sys.path.insert(0, ".")
import numpy as np

import bench
from random_forest_mc_b200 import engine
from random_forest_mc_b200.model import RandomForestMC

logging.disable(logging.WARNING)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
if len(sys.argv) > 2:
    sys.setswitchinterval(float(sys.argv[2]))
frame = bench.make_frame(1_000_000, 64, 4, seed=0)
m = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4, **bench.REGIMES["c2b"])
m.process_dataset(frame)
np.random.seed(1)
random.seed(1)
plans = [m._draw_plan(256 // K, 4) for _ in range(K)]
rows = {}


def work(r, reps):
    st = engine.stream_acquire(0)
    plan = plans[r]
    idx_T = np.stack([p[0] for p in plan])
    idx_V = np.stack([p[1] for p in plan])
    slot_of, feats = [], []
    for s, p in enumerate(plan):
        for f in p[2]:
            slot_of.append(s)
            feats.append(f)
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        b = engine.GrowBatch(m._dev_dataset, idx_T, idx_V, slot_of, feats, m.max_depth, m.min_samples_split, st)
        t1 = time.perf_counter()
        b.run()
        t2 = time.perf_counter()
        b.results()
        t3 = time.perf_counter()
        b.fetch(list(range(0, b.n_cand, 4)))
        t4 = time.perf_counter()
        b.close()
        t5 = time.perf_counter()
        out.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4))
    rows[r] = np.array(out) * 1e3
    engine.stream_release(0, st)


for rnd in range(3):
    th = [threading.Thread(target=work, args=(r, 4)) for r in range(K)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = (time.perf_counter() - t0) * 1e3
    a = np.stack([rows[r] for r in range(K)])  # [thread, rep, phase]
    print(f"round {rnd}: wall {wall:.1f} ms for 4 reps x {K} threads; per call mean/max ms: "
          + ", ".join(f"{n} {a[:, :, i].mean():.2f}/{a[:, :, i].max():.2f}" for i, n in enumerate(["create", "run", "results", "fetch", "close"])))
