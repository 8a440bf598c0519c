"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the rank-sharded fit with the real engine.
Same contract as tests/test_dist_gloo.py: rank r == the serial algorithm under seed s_r, forest =
concatenation in rank order; row-sharded voting tiles the frame."""
import os
import pickle
import random
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, pickle, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as td
from tests.helpers import load_golden
from random_forest_mc_b200.model import RandomForestMC
from random_forest_mc_b200 import dist
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
g = load_golden("mixed_default")
kw = dict(g["kwargs"]); kw["n_trees"] = 5
m = RandomForestMC(**kw)
m.device = local
m.rank_seeds = [1000, 1001]
m.fitParallel(g["frame"].copy(), disable_progress_bar=True)
(lo, hi), labels = dist.predict_sharded(m, g["frames"]["nan"], want_probs=False)
used = [sorted(t.used_features) for t in m.data]  # from the gathered masks, before any table is materialised
# the exchange went through the C ABI (rfmc_allgather_survivors): every rank holds the same gathered tables
h = torch.tensor([m._last_gather.checksum() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
hs = [torch.zeros_like(h) for _ in range(td.get_world_size())]
td.all_gather(hs, h)
# one slot for two ranks: rank 1 contributes nothing
m1 = RandomForestMC(**dict(kw, n_trees=1))
m1.device = local
m1.rank_seeds = [1000, 1001]
m1.fitParallel(g["frame"].copy(), disable_progress_bar=True)
# two host streams per rank (the bench's layout): worker order inside a rank, rank order across
m2 = RandomForestMC(**dict(kw, n_trees=4))
m2.device = local
m2.process_dataset(g["frame"].copy())
f2 = dist.fit_sharded_local(m2, 4, 2, [2000 + 10 * td.get_rank(), 2001 + 10 * td.get_rank()])
out = dict(used=used, trees=[t.data for t in m.data], scores=[float(s) for s in m.survived_scores], lo=lo, hi=hi, labels=labels,
           checksums=[int(x.item()) for x in hs], one=[t.data for t in m1.data], streams=[t.data for t in f2],
           stream_scores=[float(t.survived_score) for t in f2])
with open(os.path.join({tmp!r}, "rank%d.pkl" % td.get_rank()), "wb") as f:
    pickle.dump(out, f)
td.destroy_process_group()
"""


def _split_features(tree):
    seen, todo = set(), [tree]
    while todo:
        node = todo.pop()
        name = next(iter(node))
        if name != "leaf":
            seen.add(name)
            todo += [node[name]["split"]["<"], node[name]["split"][">="]]
    return seen


def test_sharded_fit_over_nccl_world2():
    from random_forest_mc_b200 import engine

    if engine.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import rfmc_oracle as orc
    from tests.helpers import load_golden, tree_diff

    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "worker.py")
        with open(script, "w") as f:
            f.write(WORKER.format(root=ROOT, tmp=tmp))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29513", script]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
        out = [pickle.load(open(os.path.join(tmp, f"rank{r}.pkl"), "rb")) for r in range(2)]
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
    for r in range(2):
        assert out[r]["scores"] == want_scores
        for a, b in zip(out[r]["trees"], want_trees):
            assert tree_diff(a, b) is None
        for u, b in zip(out[r]["used"], want_trees):
            assert u == sorted(_split_features(b))
    fr = g["frames"]["nan"]
    want = orc.forest_labels(want_trees, want_scores, g["class_vals"], fr, False, False)
    assert out[0]["labels"] + out[1]["labels"] == want
    for r in range(2):
        assert out[r]["checksums"][0] == out[r]["checksums"][1]
        assert len(out[r]["one"]) == 1 and tree_diff(out[r]["one"][0], want_trees[0]) is None
    # host streams: rank r stream w == the serial algorithm on 2 slots under seed 2000 + 10 r + w
    want_streams = []
    for seed in (2000, 2001, 2010, 2011):
        ds = orc.ingest(g["frame"].copy(), target_col=kw["target_col"])
        np.random.seed(seed)
        random.seed(seed)
        t, s, _ = orc.fit(ds, 2, max_discard_trees=kw["max_discard_trees"])
        want_streams += t
    for r in range(2):
        assert len(out[r]["streams"]) == 8
        for a, b in zip(out[r]["streams"], want_streams):
            assert tree_diff(a, b) is None
