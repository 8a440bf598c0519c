import sys
import numpy as np
sys.path.insert(0, ".")
from random_forest_mc_b200 import engine
per = 250_000
sizes = [per, per + 1, per - 1, per]
g = np.random.default_rng(0)
labels = np.concatenate([np.full(s, str(c)) for c, s in enumerate(sizes)])
g.shuffle(labels)
cols = {"x": np.zeros(len(labels)), "target": labels.astype(object)}
enc, ds = engine.encode_dataset_device(cols, ["x"], ["x"], "target", list(dict.fromkeys(labels.tolist())))
n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for per_class, n_train in ((2500, 2000), (600, 500), (100, 50)):
    np.random.seed(1)
    st = np.random.get_state()
    plan = engine.DrawPlan(ds, enc.class_rows, st[1], int(st[2]), n_slots, per_class, n_train, 4, 2, 65)
    counts, key_end, pos_end = plan.finish()
    gT, gV = plan.indices()
    key = st[1].copy()
    T, V, K, ks, ps, pos = engine.legacy_draw_plan(key, int(st[2]), n_slots, enc.class_rows, per_class, n_train, 4, 2, 65, 0)
    bad = 0
    for s in range(n_slots):
        for c in range(4):
            a = np.concatenate([gT[s, c * n_train:(c + 1) * n_train], gV[s, c * (per_class - n_train):(c + 1) * (per_class - n_train)]])
            b = np.concatenate([T[s, c * n_train:(c + 1) * n_train], V[s, c * (per_class - n_train):(c + 1) * (per_class - n_train)]])
            d = np.flatnonzero(a != b)
            if len(d):
                bad += 1
                if bad <= 6:
                    print(f"per_class {per_class} slot {s} class {c}: {len(d)} differ, first at p={d[:8]}, got {a[d[:4]]} want {b[d[:4]]}; same multiset {np.array_equal(np.sort(a), np.sort(b))}")
    print("per_class", per_class, "bad draws", bad, "of", n_slots * 4, "counts ok", np.array_equal(counts, K), "state ok", np.array_equal(key_end, key) and pos_end == pos, plan.info())
    plan.close()
