"""process_dataset timing probe (C2 shape): total, and the device encode call alone.
    python profiles/encode_probe.py [rows] [features]"""
import logging
import sys
import time

sys.path.insert(0, ".")
import bench
from random_forest_mc_b200 import engine
from random_forest_mc_b200.model import RandomForestMC

logging.disable(logging.WARNING)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
feats = int(sys.argv[2]) if len(sys.argv) > 2 else 64
frame = bench.make_frame(rows, feats, 4, seed=0)
m = RandomForestMC(n_trees=4, target_col="target", max_discard_trees=2)
for it in range(3):
    f = frame.copy()
    t0 = time.perf_counter()
    m.process_dataset(f)
    t1 = time.perf_counter()
    enc, dev = engine.encode_dataset_device(m.dataset_numpy, m.feature_cols, m.numeric_cols, "target", m.class_vals)
    t2 = time.perf_counter()
    dev.close()
    print(f"iter {it}: process_dataset {t1 - t0:.3f} s, of which encode_dataset_device {t2 - t1:.3f} s; width {enc.width}, stride {enc.stride}")
