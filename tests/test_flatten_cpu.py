"""CPU: forest flattening + frame encoding + the CPU statement of the voting kernel reproduce the
reference's recorded testForest / testForestProbs outputs bit-for-bit (4 modes x 4 frames)."""
import numpy as np
import pytest

from oracle import code_model as cm
from random_forest_mc_b200.flat import encode_frame, forest_tables_from_dicts
from tests.helpers import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_flattened_voting_equals_reference(name):
    g = load_golden(name)
    t = forest_tables_from_dicts(g["forest"], g["survived_scores"], g["class_vals"], g["feature_cols"])
    assert (t.stride * t.width) % 4 == 0
    for (fname, soft, weighted), want in g["preds"].items():
        codes, absent = encode_frame(t, g["frames"][fname])
        probs, labels = cm.vote_tables(t, codes, absent, soft, weighted)
        ref = np.array([[d[c] for c in g["class_vals"]] for d in want["probs"]], dtype=np.float64)
        assert np.array_equal(probs, ref), (fname, soft, weighted)
        assert [g["class_vals"][k] for k in labels] == want["labels"], (fname, soft, weighted)
