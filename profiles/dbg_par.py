"""Host-side timeline of fitParallel over k host streams (debug aid): python profiles/dbg_par.py [c2a|c2b] [k]"""
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
regime = sys.argv[1] if len(sys.argv) > 1 else "c2b"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
frame = bench.make_frame(1_000_000, 64, 4, 0)
m = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4, **bench.REGIMES[regime])
m.process_dataset(frame)
m.worker_seeds = list(range(100, 100 + k))

events = []
T0 = [0.0]


def timed(obj, name, label):
    fn = getattr(obj, name)

    def wrap(*a, **kw):
        t = time.perf_counter()
        try:
            return fn(*a, **kw)
        finally:
            events.append((threading.get_ident(), label, t - T0[0], time.perf_counter() - T0[0]))

    setattr(obj, name, wrap)


timed(engine, "stream_create", "stream_create")
timed(engine, "stream_destroy", "stream_destroy")
timed(RandomForestMC, "_draw_plan", "draw")
timed(RandomForestMC, "_grow_plan", "grow_plan")
timed(engine.GrowBatch, "results", "results")
timed(engine.GrowBatch, "fetch", "fetch")
timed(engine.GrowBatch, "close", "close")
timed(RandomForestMC, "_collect", "collect")

for i in range(5):
    m.data, m.survived_scores = [], []
    events.clear()
    T0[0] = time.perf_counter()
    m.fitParallel(disable_progress_bar=True, max_workers=k)
    wall = time.perf_counter() - T0[0]
    print(f"step {i}: {wall*1e3:.1f} ms")
tids = sorted({e[0] for e in events})
for j, tid in enumerate(tids):
    line = " ".join(f"{lab}[{a*1e3:.0f}-{b*1e3:.0f}]" for t, lab, a, b in events if t == tid)
    print(f"thread {j}: {line}")
