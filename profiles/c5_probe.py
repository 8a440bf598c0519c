"""C5-shaped vote probe: R rows (device-generated codes) x 4096 merged default-size trees, weighted soft voting.
    python profiles/c5_probe.py [rows] [reps]"""
import logging
import random
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

import bench
from random_forest_mc_b200.model import RandomForestMC

logging.disable(logging.WARNING)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frame = bench.make_frame(1_000_000, 64, 4, seed=0)
m = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4)
m.process_dataset(frame)
for seed in range(16):
    np.random.seed(seed)
    random.seed(seed)
    m.fit(disable_progress_bar=True)
m.setSoftVoting(True)
m.setWeightedTrees(True)
dev = m._device_forest()
t = m._forest_tables
top = torch.tensor([c.n_unique + 1 for c in m.encoded.columns] + [1] * (t.stride - len(m.encoded.columns)), device="cuda", dtype=torch.int32)
g = torch.Generator(device="cuda")
g.manual_seed(5)
codes = torch.empty((R, t.stride), dtype=torch.int16, device="cuda")
for lo in range(0, R, 1 << 20):
    hi = min(R, lo + (1 << 20))
    codes[lo:hi] = (torch.randint(0, 1 << 30, (hi - lo, t.stride), device="cuda", dtype=torch.int32, generator=g) % top).to(torch.int16)
probs = torch.empty((R, 4), dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream()
for mode in ((True, True), (False, False)):
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        dev.predict_device(codes.data_ptr(), t.width, R, t.stride, None, mode[0], mode[1], probs.data_ptr(), 0, stream)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"soft={mode[0]} weighted={mode[1]}: {ms:.2f} ms  {R * t.n_trees / ms / 1e6:.1f} G rows*trees/s  nodes {len(t.pnodes)}")
print("checksum", float(probs.sum()))
