import json

import numpy as np

from random_forest_mc_b200 import utils


def test_json_helpers_roundtrip(tmp_path):
    p = tmp_path / "a.json"
    utils.dump_file_json(p, {"x": np.float64(1.5), "y": [np.int64(2)]})
    assert utils.load_file_json(p) == {"x": 1.5, "y": [2]}
    d = utils.LoadDicts(tmp_path)
    assert len(d) == 1 and d["a"]["x"] == 1.5 and d.a["y"] == [2]
    assert utils.flat([[1], [2, 3]]) == [1, 2, 3]
    assert utils.flatten_nested_list([1, [2, [3, [4]]], 5]) == [1, 2, 3, 4, 5]
    assert json.dumps(np.int8(3), default=utils.json_encoder) == "3"
