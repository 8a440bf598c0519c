"""Host-side profile of the serial fit loop (debug aid): python profiles/dbg_e2e.py [c2a|c2b]"""
import cProfile
import logging
import pstats
import random
import sys
import time

sys.path.insert(0, ".")
import numpy as np

import bench
from random_forest_mc_b200.model import RandomForestMC

logging.disable(logging.WARNING)
regime = sys.argv[1] if len(sys.argv) > 1 else "c2a"
frame = bench.make_frame(1_000_000, 64, 4, 0)
m = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4, **bench.REGIMES[regime])
m.process_dataset(frame)
for i in range(6):
    m.data, m.survived_scores = [], []
    np.random.seed(1)
    random.seed(1)
    t = time.perf_counter()
    m._fit_slots(256)
    print("step", i, round((time.perf_counter() - t) * 1e3, 1), "ms", {k: round(v, 4) if isinstance(v, float) else v for k, v in m.fit_stats.items()})
np.random.seed(1)
random.seed(1)
cProfile.run("m._fit_slots(256)", "/tmp/fit.prof")
pstats.Stats("/tmp/fit.prof").sort_stats("cumtime").print_stats(25)
