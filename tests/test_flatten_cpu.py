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


def test_flat_feature_lists_behave_like_the_list_of_arrays_they_replace():
    """engine.FlatFeatureLists (feature ids of a plan's candidates back to back, what rfmc_batch_create takes) is a
    read-only sequence of per-candidate arrays: length, indexing, negative indices, slices and iteration."""
    import numpy as np

    from random_forest_mc_b200.engine import FlatFeatureLists

    lists = [np.array([3, 1, 2], dtype=np.int32), np.array([7], dtype=np.int32), np.array([], dtype=np.int32), np.array([5, 0], dtype=np.int32)]
    flat = FlatFeatureLists(np.concatenate(lists), np.array([[3, 1], [0, 2]]))
    assert len(flat) == 4 and flat.off.tolist() == [0, 3, 4, 4, 6]
    for k, want in enumerate(lists):
        assert flat[k].tolist() == want.tolist() and flat[k - 4].tolist() == want.tolist()
    assert [x.tolist() for x in flat] == [x.tolist() for x in lists]
    assert [x.tolist() for x in flat[1:3]] == [x.tolist() for x in lists[1:3]]
    assert [x.tolist() for x in list(flat)[2:]] == [[], [5, 0]]
