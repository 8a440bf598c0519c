"""Vote kernel: device time with the tree tops staged in shared memory vs the all-global blocked walk.
usage: python profiles/vote_probe.py   (C2b forest: 256 trees grown on 8,000 rows; C5-like forest: 4,096 default trees)"""
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from random_forest_mc_b200.flat import encode_frame  # noqa: E402
from random_forest_mc_b200.model import RandomForestMC  # noqa: E402

import logging

logging.disable(logging.WARNING)
frame = bench.make_frame(1_000_000, 64, 4, seed=0)
stream = torch.cuda.current_stream()


def time_votes(model, label, rows, soft, weighted):
    tables = None
    out = {}
    for env in ("0", "1"):
        if env == "0":
            os.environ["RFMC_PRED_TOPS"] = "1"
        else:
            os.environ.pop("RFMC_PRED_TOPS", None)
        dev = model._device_forest()
        tables = model._forest_tables
        codes, absent = encode_frame(tables, frame.drop(columns=["target"]).head(rows))
        codes_t = torch.from_numpy(codes).cuda()
        probs_t = torch.empty((rows, 4), dtype=torch.float64, device="cuda")
        labels_t = torch.empty(rows, dtype=torch.int32, device="cuda")
        run = lambda: dev.predict_device(codes_t.data_ptr(), tables.width, rows, tables.stride, None, soft, weighted, probs_t.data_ptr(), labels_t.data_ptr(), stream)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5):
            run()
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        out[env] = (ms, probs_t.cpu().numpy().copy(), labels_t.cpu().numpy().copy())
    same = np.array_equal(out["0"][1], out["1"][1]) and np.array_equal(out["0"][2], out["1"][2])
    T = tables.n_trees
    print(f"{label}: staged tops {out['0'][0]:.3f} ms = {rows * T / out['0'][0] / 1e-3:.3e} rows*trees/s | all-global {out['1'][0]:.3f} ms = {rows * T / out['1'][0] / 1e-3:.3e} | bit-equal {same} | nodes {len(tables.pnodes)}", flush=True)


m = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4, batch_train_pclass=2000, batch_val_pclass=500)
m.process_dataset(frame.copy())
np.random.seed(0), random.seed(0)
m.fit()
time_votes(m, "C2b forest, hard votes, 1M rows", 1_000_000, False, False)
time_votes(m, "C2b forest, weighted soft, 1M rows", 1_000_000, True, True)
m5 = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4)
m5.process_dataset(frame.copy())
for seed in range(16):
    np.random.seed(seed), random.seed(seed)
    m5.fit()
time_votes(m5, "C5 forest (4096 default trees), weighted soft, 1M rows", 1_000_000, True, True)
