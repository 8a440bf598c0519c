"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (ysraell/random-forest-mc v1.4.0) ships no seeded known-answer tests for this path
(SURVEY.md section 4 / 8c), so the vectors are recorded from the reference itself under
``np.random.seed(s); random.seed(s)``.  For every case we store the input frame, the constructor
arguments, the seed, one record per candidate tree (bootstrap rows, hold-out rows, ordered feature
list, the nested tree dict exactly as the reference built it incl. numpy value types, and its
hold-out score), the kept forest + survived scores, and testForest / testForestProbs outputs in the
four voting modes on three predict frames (clean, NaN cells, dropped columns).

The GPU box has no /root/reference; tests read only the committed ``*.pkl.gz`` files.
"""

import gzip
import os
import pickle
import random
import sys

import numpy as np
import pandas as pd

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def titanic_like(seed=0, n=712):
    """Titanic-shaped synthetic frame (the real CSV is a Git-LFS pointer)."""
    g = np.random.default_rng(seed)
    pclass = g.choice(["1", "2", "3"], size=n, p=[0.25, 0.2, 0.55])
    sex = g.choice(["male", "female"], size=n, p=[0.64, 0.36])
    age = g.integers(0, 80, size=n).astype(np.uint8)
    sibsp = g.integers(0, 8, size=n).astype(np.uint8)
    fare = g.integers(0, 512, size=n).astype(np.uint32)
    emb = g.choice(["S", "C", "Q"], size=n, p=[0.7, 0.2, 0.1])
    z = 1.2 * (sex == "female") - 0.8 * (pclass == "3") + 0.6 * (pclass == "1") - 0.01 * age.astype(float) + g.normal(size=n) * 0.7
    surv = (z > np.median(z)).astype(int).astype(str)
    return pd.DataFrame({"Pclass": pclass, "Sex": sex, "Age": age, "SibSp": sibsp, "Fare": fare, "Embarked": emb, "Survived": surv})


def mixed(seed=1, n=2500):
    """All column dtypes the encode path distinguishes + labels whose lexicographic order differs
    from first-appearance order ('10' < '2' < 'b')."""
    g = np.random.default_rng(seed)
    df = pd.DataFrame(
        {
            "f64": np.round(g.normal(size=n), 2),
            "i64": g.integers(-20, 50, size=n),
            "f32": np.round(g.normal(size=n), 1).astype(np.float32),
            "u8": g.integers(0, 9, size=n).astype(np.uint8),
            "i16": g.integers(-300, 300, size=n).astype(np.int16),
            "cat": g.choice(["a", "b", "c", "dd", "10", "2"], size=n),
            "flag": g.choice([True, False], size=n),
            "wide": g.normal(size=n),
        }
    )
    s = df.f64 + (df.cat == "a") * 1.0 + df.i64 / 40 + 0.5 * df.f32 + g.normal(size=n) * 0.3
    df["target"] = pd.qcut(s, 3, labels=["b", "10", "2"]).astype(str)
    return df


def ties(seed=2, n=900):
    """Few distinct values everywhere: exercises the '>' fallback, constant columns, equal values
    in 2-row nodes, mode ties and full-rotation leaves."""
    g = np.random.default_rng(seed)
    df = pd.DataFrame(
        {
            "a_1": g.integers(0, 3, size=n),
            "b_2": g.choice([0.0, 0.5, 0.5, 1.0], size=n),
            "c_3": np.zeros(n, dtype=np.float32),
            "d_4": g.choice(["x", "y"], size=n),
            "e_5": g.choice(["k"], size=n),
            "f_6": g.integers(0, 2, size=n).astype(np.int8),
            "g_7": g.choice([-1.5, 2.25, 2.25, 2.25, 7.0], size=n).astype(np.float32),
        }
    )
    s = df.a_1 + (df.d_4 == "x") * 1.0 + df.b_2 + g.normal(size=n) * 0.8
    df["target"] = pd.qcut(s, 4, labels=["c3", "c0", "c2", "c1"]).astype(str)
    return df


def nan_cells(frame, target, seed, frac=0.2):
    g = np.random.default_rng(seed)
    out = frame.copy()
    for c in out.columns:
        if c == target:
            continue
        hit = g.random(len(out)) < frac
        if pd.api.types.is_numeric_dtype(out[c]) and not pd.api.types.is_bool_dtype(out[c]):
            col = out[c].astype("float64" if out[c].dtype != np.float32 else "float32")
            col[hit] = np.nan
            out[c] = col
        else:
            col = out[c].astype(object)
            col[hit] = np.nan
            out[c] = col
    return out


CASES = [
    # name, frame builder, ctor kwargs, seed, n predict rows
    ("titanic_c1", titanic_like, dict(n_trees=8, target_col="Survived", max_discard_trees=4), 0, 60),
    ("mixed_default", mixed, dict(n_trees=6, target_col="target", max_discard_trees=3), 1, 50),
    ("mixed_big", mixed, dict(n_trees=3, target_col="target", max_discard_trees=2, batch_train_pclass=150, batch_val_pclass=40), 2, 40),
    ("mixed_mss3_d6", mixed, dict(n_trees=4, target_col="target", max_discard_trees=2, batch_train_pclass=60, batch_val_pclass=20, min_samples_split=3, max_depth=6), 3, 40),
    ("mixed_nobest", mixed, dict(n_trees=4, target_col="target", max_discard_trees=2, get_best_tree=False, th_start=0.9, delta_th=0.15, batch_train_pclass=25, batch_val_pclass=12), 4, 30),
    ("mixed_feat", mixed, dict(n_trees=4, target_col="target", max_discard_trees=3, min_feature=3, max_feature=5, th_start=0.6), 5, 30),
    ("ties_default", ties, dict(n_trees=6, target_col="target", max_discard_trees=3, batch_train_pclass=40, batch_val_pclass=15), 6, 50),
    ("ties_temporal", ties, dict(n_trees=4, target_col="target", max_discard_trees=2, temporal_features=True, batch_train_pclass=30, batch_val_pclass=10, min_samples_split=2), 7, 30),
]


def rng_fingerprint():
    """Position-sensitive fingerprint of both global RNG states (numpy legacy MT19937 + CPython)."""
    import hashlib

    st = np.random.get_state()
    h = hashlib.sha256()
    h.update(st[1].tobytes())
    h.update(repr(st[2:]).encode())
    h.update(repr(random.getstate()).encode())
    return h.hexdigest()


def record(name, build, kwargs, seed, n_pred):
    from random_forest_mc.model import RandomForestMC  # the real reference

    frame = build()
    model = RandomForestMC(**kwargs)
    trace = []
    state = {}
    o_split, o_feats, o_plant, o_val = model.split_train_val, model.sampleFeats, model.plantTree, model.validationTree

    def split():
        a, b = o_split()
        state["T"], state["V"] = a.copy(), b.copy()
        return a, b

    def feats(cols):
        out = o_feats(cols)
        state["F"] = list(out)
        return out

    def plant(idx, F):
        t = o_plant(idx, F)
        state["tree"] = t.data
        return t

    def val(tree, idx):
        s = o_val(tree, idx)
        trace.append({"idx_T": state["T"], "idx_V": state["V"], "F": state["F"], "tree": state["tree"], "th_val": s})
        return s

    model.split_train_val, model.sampleFeats, model.plantTree, model.validationTree = split, feats, plant, val
    np.random.seed(seed)
    random.seed(seed)
    train = frame.copy()
    model.fit(train, disable_progress_bar=True)
    rng_after = rng_fingerprint()

    tcol = kwargs["target_col"]
    base = frame.drop(columns=[tcol]).iloc[:: max(1, len(frame) // n_pred)].head(n_pred).reset_index(drop=True)
    feats_all = list(base.columns)
    drop = feats_all[1::3]
    frames = {
        "clean": base,
        "nan": nan_cells(base, tcol, seed + 100),
        "absent": base.drop(columns=drop),
        "absent_nan": nan_cells(base.drop(columns=feats_all[:1]), tcol, seed + 200),
    }
    preds = {}
    for fname, fr in frames.items():
        for soft in (False, True):
            for weighted in (False, True):
                model.setSoftVoting(soft)
                model.setWeightedTrees(weighted)
                preds[(fname, soft, weighted)] = {"labels": model.testForest(fr), "probs": model.testForestProbs(fr)}
    return {
        "name": name,
        "frame": frame,
        "kwargs": kwargs,
        "seed": seed,
        "class_vals": model.class_vals,
        "feature_cols": model.feature_cols,
        "numeric_cols": model.numeric_cols,
        "type_of_cols": model.type_of_cols,
        "rows_per_class": model._N,
        "min_feature": model.min_feature,
        "max_feature": model.max_feature,
        "temporal_features": model.temporal_features,
        "n_samples": model.n_samples,
        "trace": trace,
        "forest": [t.data for t in model.data],
        "survived_scores": list(model.survived_scores),
        "md5": [t.md5hexdigest for t in model.data],
        "frames": frames,
        "preds": preds,
        "rng_after": rng_after,
    }


def main():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("make_golden.py needs the reference checkout at /root/reference (build container only)")
    sys.path.insert(0, REF_SRC)
    import logging

    logging.disable(logging.WARNING)
    for name, build, kwargs, seed, n_pred in CASES:
        rec = record(name, build, kwargs, seed, n_pred)
        path = os.path.join(HERE, name + ".pkl.gz")
        with gzip.open(path, "wb") as f:
            pickle.dump(rec, f, protocol=4)
        print(f"{name}: {len(rec['trace'])} candidates, {len(rec['forest'])} kept, {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
