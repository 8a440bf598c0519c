"""BASELINE config 3 as a vote-kernel measurement: C2b forest (256 trees of ~14k nodes), 1M rows with 20 % of the cells
missing (code 0 -> '<'), then additionally 13 of the 64 columns absent ('>=' + renormalised first leaf), weighted soft
voting, float64 probabilities out.  Device time per launch (CUDA events), inputs resident in HBM."""
import logging
import random
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

import bench
from random_forest_mc_b200.model import RandomForestMC

logging.disable(logging.WARNING)
frame = bench.make_frame(1_000_000, 64, 4, seed=0)
m = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4, **bench.REGIMES["c2b"])
m.process_dataset(frame)
np.random.seed(0)
random.seed(0)
m.fitParallel(disable_progress_bar=True)
dev = m._device_forest()
t = m._forest_tables
from random_forest_mc_b200.flat import encode_frame

codes, _ = encode_frame(t, frame.drop(columns=["target"]))
g = np.random.default_rng(3)
holes = codes.copy()
holes[g.random(holes.shape) < 0.2] = 0
absent = np.zeros(t.n_features, dtype=np.uint8)
absent[g.choice(t.n_features, size=13, replace=False)] = 1
R = len(codes)
stream = torch.cuda.current_stream()
probs = torch.empty((R, 4), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, arr, ab in (("clean, hard unweighted", codes, None), ("clean, weighted soft", codes, None),
                      ("20% cells missing, weighted soft", holes, None), ("20% cells missing + 13 absent columns, weighted soft", holes, absent)):
    soft = weighted = "soft" in name
    d = torch.from_numpy(arr).cuda()
    ms = []
    for rep in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        dev.predict_device(d.data_ptr(), t.width, R, t.stride, ab, soft, weighted, probs.data_ptr(), 0, stream)
        b.record(stream)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    best = float(np.median(ms[1:]))
    print(f"{name}: {best:.2f} ms  {R * t.n_trees / best / 1e6:.1f} G rows*trees/s   row sums ok: {bool(((probs[:4096].sum(dim=1) - 1).abs() < 1e-12).all())}")
