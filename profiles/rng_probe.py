"""Device draw plan at the bench shapes: kernel times (CUDA events) and parity with the host walk.
usage: python profiles/rng_probe.py [rows_per_class] [n_slots ...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from random_forest_mc_b200 import engine  # noqa: E402

per = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
slots = [int(x) for x in sys.argv[2:]] or [16, 64, 256]
sizes = [per, per + 1, per - 1, per]
g = np.random.default_rng(0)
labels = np.concatenate([np.full(s, str(c)) for c, s in enumerate(sizes)])
g.shuffle(labels)
cols = {"x": np.zeros(len(labels)), "target": labels.astype(object)}
enc, ds = engine.encode_dataset_device(cols, ["x"], ["x"], "target", list(dict.fromkeys(labels.tolist())))
engine.drawplan_timing(True)
for per_class, n_train in ((20, 10), (2500, 2000)):
    for n_slots in slots:
        np.random.seed(1)
        st = np.random.get_state()
        for rep in range(3):
            t0 = time.perf_counter()
            plan = engine.DrawPlan(ds, enc.class_rows, st[1], int(st[2]), n_slots, per_class, n_train, 4, 2, 65)
            counts, key_end, pos_end = plan.finish()
            wall = time.perf_counter() - t0
            ms = plan.kernel_ms()
            info = plan.info()
            if rep < 2:
                plan.close()
        key = st[1].copy()
        t0 = time.perf_counter()
        T, V, K, ks, ps, pos = engine.legacy_draw_plan(key, int(st[2]), n_slots, enc.class_rows, per_class, n_train, 4, 2, 65, 2)
        host = time.perf_counter() - t0
        gT, gV = plan.indices()
        ok = np.array_equal(gT, T) and np.array_equal(gV, V) and np.array_equal(counts, K) and np.array_equal(key_end, key) and pos_end == pos
        plan.close()
        print(f"rows/class {per} per_class {per_class} slots {n_slots}: fill {ms[0]:.3f} rounds {ms[1]:.3f} final {ms[2]:.3f} resolve {ms[3]:.3f} ms"
              f" | wall {wall*1e3:.2f} ms | host walk {host*1e3:.1f} ms | words {info['words_consumed']/1e6:.1f}M pending {info['words_pending']/1e6:.1f}M launches {info['launches']} redone {info['chunk_fallbacks']} | parity {ok}", flush=True)
