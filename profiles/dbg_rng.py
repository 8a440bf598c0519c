"""Host RNG walk: ms per tree slot of rfmc_legacy_draw_plan by worker / producer configuration (debug aid)."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from random_forest_mc_b200 import engine

n, C = 1_000_000, 4
rows = [np.arange(c, n, C, dtype=np.int64) for c in range(C)]


def run(flags, slots=64):
    st = np.random.RandomState(5).get_state()
    key = st[1].copy()
    t = time.perf_counter()
    engine.legacy_draw_plan(key, int(st[2]), slots, rows, 2500, 2000, 4, 2, 65, flags)
    return (time.perf_counter() - t) / slots * 1e3


for fl in (0, 1, 2, 3, 4, 6, 8, 0x100, 0x102, 0x104, 0x106):
    run(fl, 4)
    print(hex(fl), "ms/slot", round(min(run(fl) for _ in range(3)), 3))
