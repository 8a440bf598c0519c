"""Vote kernel for forests of small trees (predict_small_kernel) against the generic blocked walk: device time
and bit-equality of labels and probabilities on a C5-like forest (4,096 default-configuration trees) and on a
256-tree default forest, all four voting modes.   usage: python profiles/vote_small_probe.py [rows]"""
import logging
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from random_forest_mc_b200.flat import encode_frame  # noqa: E402
from random_forest_mc_b200.model import RandomForestMC  # noqa: E402

logging.disable(logging.WARNING)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
frame = bench.make_frame(1_000_000, 64, 4, seed=0)
stream = torch.cuda.current_stream()


def time_votes(model, label, soft, weighted):
    dev = model._device_forest()
    tables = model._forest_tables
    codes, absent = encode_frame(tables, frame.drop(columns=["target"]).head(rows))
    codes_t = torch.from_numpy(codes).cuda()
    out = {}
    for small in ("0", "r8", "1"):
        os.environ["RFMC_PRED_SMALL"] = "0" if small == "0" else "1"
        os.environ["RFMC_PRED_SMALL_REC4"] = "0" if small == "r8" else "1"
        probs_t = torch.empty((rows, 4), dtype=torch.float64, device="cuda")
        labels_t = torch.empty(rows, dtype=torch.int32, device="cuda")
        run = lambda: dev.predict_device(codes_t.data_ptr(), tables.width, rows, tables.stride, None, soft, weighted, probs_t.data_ptr(), labels_t.data_ptr(), stream)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(3):
            run()
        b.record(stream)
        torch.cuda.synchronize()
        out[small] = (a.elapsed_time(b) / 3, probs_t.cpu().numpy().copy(), labels_t.cpu().numpy().copy())
    os.environ.pop("RFMC_PRED_SMALL", None)
    os.environ.pop("RFMC_PRED_SMALL_REC4", None)
    same = all(np.array_equal(out["0"][1], out[k][1]) and np.array_equal(out["0"][2], out[k][2]) for k in ("r8", "1"))
    T = tables.n_trees
    print(f"{label} soft={soft} weighted={weighted}: generic {out['0'][0]:.3f} ms = {rows * T / out['0'][0] / 1e-3:.3e} rows*trees/s | small-tree kernel, 8-byte records "
          f"{out['r8'][0]:.3f} ms | 4-byte records {out['1'][0]:.3f} ms = {rows * T / out['1'][0] / 1e-3:.3e} "
          f"| bit-equal {same} | trees {T} nodes {len(tables.pnodes)}", flush=True)


m5 = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4)
m5.process_dataset(frame.copy())
np.random.seed(0), random.seed(0)
m5.fit()
for soft, weighted in ((False, False), (True, True)):
    time_votes(m5, "default forest (256 trees)", soft, weighted)
for seed in range(1, 16):
    np.random.seed(seed), random.seed(seed)
    m5.fit()
for soft, weighted in ((False, False), (False, True), (True, False), (True, True)):
    time_votes(m5, "C5 forest (4096 default trees)", soft, weighted)
