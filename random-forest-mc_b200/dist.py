"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the box,
gloo in the CPU tests).

Fit (the ``fitParallel`` analogue, model.py:372-404): tree slots are independent given their RNG
draws, so rank r grows slots ``[r*n/R, (r+1)*n/R)`` against its own replica of the encoded matrix
with NO data-path collective; only the survivors' node tables are exchanged, by one all-gather
after the last slot (sizes first, then the padded payload).  The reference's own ``fitParallel`` is
not reproducible (forked workers reseed ``random`` from OS entropy), so the checkable
contract is: rank r == the serial algorithm on ``n_r`` slots under ``np.random.seed(s_r);
random.seed(s_r)``, and the forest is the concatenation in rank order (model.py:403).

Predict: rows are sharded contiguously over ranks, the forest is replicated, no collective.
"""

from __future__ import annotations

import random as _pyrandom
from typing import List, Optional, Sequence

import numpy as np

from .flat import NODE_DTYPE, TreeSoA


def _td():
    import torch.distributed as td

    return td


def is_distributed() -> bool:
    try:
        td = _td()
    except Exception:
        return False
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def rank_world():
    if not is_distributed():
        return 0, 1
    td = _td()
    return td.get_rank(), td.get_world_size()


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced split of ``n`` units: the first ``n % world`` ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _comm_device():
    import torch

    td = _td()
    return torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")


def all_gather_trees(trees: Sequence[TreeSoA], scores: Sequence[float], n_classes: int):
    """All-gather variable-length survivor tables.  Returns [(TreeSoA, score)] in rank order."""
    import torch

    td = _td()
    world = td.get_world_size()
    dev = _comm_device()
    sizes = np.array([t.n_nodes for t in trees], dtype=np.int64)
    nodes = np.concatenate([t.nodes for t in trees]) if len(trees) else np.zeros(0, dtype=NODE_DTYPE)
    counts = np.concatenate([t.counts for t in trees]) if len(trees) else np.zeros((0, n_classes), dtype=np.int32)
    sc = np.array([float(s) for s in scores], dtype=np.float64)
    meta = torch.tensor([len(trees), len(nodes)], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    td.all_gather(metas, meta)
    metas = [m.cpu().numpy() for m in metas]
    max_trees = max(int(m[0]) for m in metas)
    max_nodes = max(int(m[1]) for m in metas)

    def padded(raw: bytes, nbytes: int):
        buf = np.zeros(nbytes, dtype=np.uint8)
        buf[: len(raw)] = np.frombuffer(raw, dtype=np.uint8)
        return torch.from_numpy(buf).to(dev)

    def gather(raw: bytes, nbytes: int) -> List[np.ndarray]:
        mine = padded(raw, max(nbytes, 1))
        outs = [torch.zeros_like(mine) for _ in range(world)]
        td.all_gather(outs, mine)
        return [o.cpu().numpy() for o in outs]

    g_sizes = gather(sizes.tobytes(), max_trees * 8)
    g_scores = gather(sc.tobytes(), max_trees * 8)
    g_nodes = gather(nodes.tobytes(), max_nodes * NODE_DTYPE.itemsize)
    g_counts = gather(np.ascontiguousarray(counts).tobytes(), max_nodes * n_classes * 4)
    out = []
    for r in range(world):
        nt, nn = int(metas[r][0]), int(metas[r][1])
        r_sizes = g_sizes[r][: nt * 8].view(np.int64)
        r_scores = g_scores[r][: nt * 8].view(np.float64)
        r_nodes = g_nodes[r][: nn * NODE_DTYPE.itemsize].view(NODE_DTYPE)
        r_counts = g_counts[r][: nn * n_classes * 4].view(np.int32).reshape(nn, n_classes)
        at = 0
        for k in range(nt):
            s = int(r_sizes[k])
            out.append((TreeSoA(r_nodes[at : at + s].copy(), r_counts[at : at + s].copy()), np.float64(r_scores[k])))
            at += s
    return out


def fit_sharded(model, rank_seeds: Optional[Sequence[int]] = None):
    """Rank-sharded fit; every rank returns the full forest (list of DecisionTreeMC, rank order)."""
    rank, world = rank_world()
    lo, hi = shard_bounds(model.n_trees, rank, world)
    seeds = rank_seeds if rank_seeds is not None else getattr(model, "rank_seeds", None)
    if seeds is not None:
        np.random.seed(seeds[rank])
        _pyrandom.seed(seeds[rank])
    local = model._fit_slots(hi - lo) if hi > lo else []
    for t in local:
        if t.soa is None:
            raise RuntimeError("sharded fit expects SoA-backed trees")
    gathered = all_gather_trees([t.soa for t in local], [t.survived_score for t in local], len(model.class_vals))
    forest = []
    for soa, score in gathered:
        t = model._wrap(soa, score)
        t.features = model.feature_cols
        t.used_features = model.tree2feats(t)
        forest.append(t)
    return forest


def fit_sharded_local(model, n_local: int):
    """Weak-scaling variant used by bench.py: every rank grows ``n_local`` slots from its current
    RNG state, then the survivors are all-gathered (rank order)."""
    local = model._fit_slots(n_local)
    if not is_distributed():
        return local
    gathered = all_gather_trees([t.soa for t in local], [t.survived_score for t in local], len(model.class_vals))
    forest = []
    for soa, score in gathered:
        t = model._wrap(soa, score)
        t.features = model.feature_cols
        t.used_features = model.tree2feats(t)
        forest.append(t)
    return forest


def predict_sharded(model, frame, want_probs: bool):
    """Row-sharded voting: this rank's contiguous slice of ``frame``; no collective."""
    rank, world = rank_world()
    lo, hi = shard_bounds(len(frame), rank, world)
    part = frame.iloc[lo:hi]
    return (lo, hi), (model.testForestProbs(part) if want_probs else model.testForest(part))
