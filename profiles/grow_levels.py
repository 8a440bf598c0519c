"""Per-level cycle profile of the grower (debug build):
    RFMC_NVCC_DEFS=-DRFMC_GROW_TIMING python random-forest-mc_b200/build.py --force && python profiles/grow_levels.py"""
import ctypes as C
import logging
import random
import sys

sys.path.insert(0, ".")
import numpy as np

import bench
from random_forest_mc_b200 import engine
from random_forest_mc_b200.model import RandomForestMC

logging.disable(logging.WARNING)
frame = bench.make_frame(1_000_000, 64, 4, 0)
m = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4, **bench.REGIMES["c2b"])
m.process_dataset(frame)
np.random.seed(1)
random.seed(1)
plan = m._draw_plan(256, 4)
for _ in range(2):
    batch, _ = m._grow_plan(plan)
    batch.results()
    out = np.zeros(8192, dtype=np.uint64)
    lib = engine.load_library()
    lib.rfmc_debug_read.argtypes = [C.c_void_p, C.c_int]
    rc = lib.rfmc_debug_read(out.ctypes.data_as(C.c_void_p), 8192)
    batch.close()
assert rc == 0, "build with -DRFMC_GROW_TIMING"
t = out.astype(np.int64)
levels = int(t[4])
print(f"gather {t[1]-t[0]:>9d} cyc | levels {t[2]-t[1]:>9d} | scoring {t[3]-t[2]:>9d} | total {t[3]-t[0]}")
print("lvl   tiny small medium large huge |   phase1   scan   emit")
for L in range(levels):
    b = 16 + 8 * L
    print(f"{L:3d} {t[b+4]:6d} {t[b+5]:5d} {t[b+6]:6d} {t[b+7] & 0xFFFF:5d} {t[b+7] >> 16:4d} | {t[b+1]-t[b]:8d} {t[b+2]-t[b+1]:6d} {t[b+3]-t[b+2]:6d}")
