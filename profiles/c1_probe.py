"""BASELINE config 1 (titanic-shaped frame of tests/golden/titanic_c1, n_trees = 8 / 16 / 32, max_discard_trees = 4):
wall seconds of fit() -- the quantity the reference publishes (linkedin_post.md:24-26: 0.0635 / 0.1439 / 0.4631 s,
hardware not stated) -- through the B200 engine and through the CPU port of the reference on this box."""
import logging
import random
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from oracle import rfmc_oracle as orc
from random_forest_mc_b200.model import RandomForestMC
from tests.helpers import load_golden

logging.disable(logging.WARNING)
g = load_golden("titanic_c1")
frame = g["frame"]
test = frame.drop(columns=["Survived"])
for n_trees in (8, 16, 32):
    best_gpu, best_cpu, best_vote = 1e9, 1e9, 1e9
    for rep in range(5):
        m = RandomForestMC(n_trees=n_trees, target_col="Survived", max_discard_trees=4)
        m.process_dataset(frame.copy())
        np.random.seed(rep)
        random.seed(rep)
        t0 = time.perf_counter()
        m.fit(disable_progress_bar=True)
        best_gpu = min(best_gpu, time.perf_counter() - t0)
        t0 = time.perf_counter()
        labels = m.testForest(test)
        best_vote = min(best_vote, time.perf_counter() - t0)
        ds = orc.ingest(frame.copy(), target_col="Survived")
        np.random.seed(rep)
        random.seed(rep)
        t0 = time.perf_counter()
        trees, scores, _ = orc.fit(ds, n_trees, max_discard_trees=4)
        best_cpu = min(best_cpu, time.perf_counter() - t0)
        assert [float(s) for s in scores] == [float(s) for s in m.survived_scores]
    acc = float(np.mean(np.array(labels) == frame["Survived"].astype(str).to_numpy()))
    print(f"n_trees={n_trees}: fit B200 engine {best_gpu * 1e3:.2f} ms, CPU port {best_cpu * 1e3:.1f} ms; "
          f"testForest(712 rows) {best_vote * 1e3:.2f} ms, train accuracy {acc:.3f}")
