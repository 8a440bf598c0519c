"""CPU, world_size 2, gloo: the rank-sharded fit (fitParallel analogue) and row-sharded predict.
Contract: rank r == the serial algorithm on its share of the slots under
np.random.seed(s_r); random.seed(s_r); the forest is the concatenation in rank order."""
import os
import pickle
import random
import socket
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, pickle, random, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as td
from tests import fake_engine
fake_engine.install_plain()
from tests.helpers import load_golden
from random_forest_mc_b200.model import RandomForestMC
from random_forest_mc_b200 import dist

td.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
g = load_golden("mixed_default")
kw = dict(g["kwargs"]); kw["n_trees"] = 5
m = RandomForestMC(**kw)
m.rank_seeds = [1000, 1001]
m.fitParallel(g["frame"].copy(), disable_progress_bar=True)
(lo, hi), labels = dist.predict_sharded(m, g["frames"]["nan"], want_probs=False)
out = dict(trees=[t.data for t in m.data], scores=[float(s) for s in m.survived_scores], lo=lo, hi=hi, labels=labels)
with open(os.path.join({tmp!r}, "rank%d.pkl" % td.get_rank()), "wb") as f:
    pickle.dump(out, f)
td.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_fit_and_predict_world2():
    from oracle import rfmc_oracle as orc
    from tests.helpers import load_golden, tree_diff

    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "worker.py")
        with open(script, "w") as f:
            f.write(WORKER.format(root=ROOT, tmp=tmp))
        port = _free_port()
        procs = []
        for r in range(2):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            procs.append(subprocess.Popen([sys.executable, script], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
        for p in procs:
            out, _ = p.communicate(timeout=300)
            assert p.returncode == 0, out.decode()[-3000:]
        res = [pickle.load(open(os.path.join(tmp, f"rank{r}.pkl"), "rb")) for r in range(2)]
    # every rank holds the same, complete forest
    assert len(res[0]["trees"]) == 5
    for a, b in zip(res[0]["trees"], res[1]["trees"]):
        assert tree_diff(a, b) is None
    assert res[0]["scores"] == res[1]["scores"]
    # == serial oracle runs per rank seed, concatenated in rank order (3 + 2 slots)
    g = load_golden("mixed_default")
    kw = g["kwargs"]
    want_trees, want_scores = [], []
    for seed, n in ((1000, 3), (1001, 2)):
        ds = orc.ingest(g["frame"].copy(), target_col=kw["target_col"])
        np.random.seed(seed)
        random.seed(seed)
        t, s, _ = orc.fit(ds, n, max_discard_trees=kw["max_discard_trees"])
        want_trees += t
        want_scores += [float(x) for x in s]
    assert res[0]["scores"] == want_scores
    for a, b in zip(res[0]["trees"], want_trees):
        assert tree_diff(a, b) is None
    # row-sharded predict: the two slices tile the frame and agree with the oracle
    fr = g["frames"]["nan"]
    assert (res[0]["lo"], res[1]["hi"]) == (0, len(fr)) and res[0]["hi"] == res[1]["lo"]
    want = orc.forest_labels(want_trees, want_scores, g["class_vals"], fr, False, False)
    assert res[0]["labels"] + res[1]["labels"] == want
