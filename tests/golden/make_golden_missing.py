"""Golden vectors for predictMissingValues, recorded from the REAL reference (build container only).
    python tests/golden/make_golden_missing.py"""
import gzip
import os
import pickle
import random
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import HERE, REF_SRC, titanic_like  # noqa: E402


def main():
    sys.path.insert(0, REF_SRC)
    import logging

    logging.disable(logging.WARNING)
    from random_forest_mc.model import RandomForestMC

    frame = titanic_like()
    kwargs = dict(n_trees=8, target_col="Survived", max_discard_trees=4)
    m = RandomForestMC(**kwargs)
    np.random.seed(21)
    random.seed(21)
    m.fit(frame.copy(), disable_progress_bar=True)
    cols = [c for c in frame.columns if c != "Survived"]
    g = np.random.default_rng(5)
    part = frame.sample(n=12, random_state=3).reset_index(drop=True)
    keep = g.choice([True, False], size=part[cols].shape, p=[0.7, 0.3])
    holes = part[cols].mask(~keep)
    holes["Survived"] = part["Survived"]
    dict_values = {c: frame[c].unique().tolist() for c in cols}
    row = next(r for _, r in holes.iterrows() if r.isna().any())
    out = {
        "frame": frame, "kwargs": kwargs, "seed": 21, "holes": holes, "row": row, "dict_values": dict_values,
        "by_row": m.predictMissingValues(row, dict_values),
        "by_frame": m.predictMissingValues(holes, dict_values),
        "survived_scores": list(m.survived_scores),
    }
    path = os.path.join(HERE, "..", "golden_missing", "missing_titanic.pkl.gz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with gzip.open(path, "wb") as f:
        pickle.dump(out, f, protocol=4)
    print("rows:", len(out["by_row"]), len(out["by_frame"]), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
