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
from typing import Optional, Sequence

import numpy as np

from .flat import NODE_DTYPE, TreeSoA, reference_used_features


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


class GatheredBlock:
    """The all-gathered survivor tables of every rank, left where the collective put them (HBM on
    the box) until some tree of the block is actually read on the host."""

    def __init__(self, nodes_t, counts_t, metas, n_classes):
        self._nodes_t, self._counts_t = nodes_t, counts_t
        self.metas, self.n_classes = metas, n_classes
        self._host = None

    def host(self):
        if self._host is None:
            nodes = self._nodes_t.cpu().numpy()
            counts = self._counts_t.cpu().numpy()
            per_rank = []
            for r, (nt, nn) in enumerate(self.metas):
                n = nodes[r][: nn * NODE_DTYPE.itemsize].view(NODE_DTYPE)
                c = counts[r][: nn * self.n_classes * 4].view(np.int32).reshape(nn, self.n_classes)
                per_rank.append((n, c))
            self._host = per_rank
            self._nodes_t = self._counts_t = None
        return self._host


class AbiGatheredBlock:
    """Same role as GatheredBlock, over the C ABI's handle (rfmc_allgather_survivors): rank r's tables
    are copied to the host the first time one of its trees is read."""

    def __init__(self, handle):
        self.handle = handle
        self._host = {}

    def host(self):
        return self

    def __getitem__(self, r):
        if r not in self._host:
            self._host[r] = self.handle.tables(r)
        return self._host[r]


def nccl_comm_ptr() -> int:
    """The default process group's ncclComm_t (what rfmc_allgather_survivors takes)."""
    import torch

    td = _td()
    backend = td.distributed_c10d._get_default_group()._get_backend(torch.device("cuda", torch.cuda.current_device()))
    try:
        return int(backend._comm_ptr())
    except Exception:
        td.barrier()  # the communicator is created lazily by the first collective
        return int(backend._comm_ptr())


def all_gather_trees_abi(trees, scores, n_classes, used_masks, dev_payload, device: int, stream=None):
    """The survivor exchange through the C ABI (NCCL path): same result layout as all_gather_trees."""
    from . import engine

    sizes = np.array([t.n_nodes for t in trees], dtype=np.int64)
    sc = np.array([float(s) for s in scores], dtype=np.float64)
    nodes_ptr = dev_payload[0].data_ptr() if dev_payload is not None and int(sizes.sum()) else 0
    counts_ptr = dev_payload[1].data_ptr() if dev_payload is not None and int(sizes.sum()) else 0
    g = engine.GatheredSurvivors(nccl_comm_ptr(), device, nodes_ptr, counts_ptr, sizes, sc, used_masks, n_classes, stream)
    block = AbiGatheredBlock(g)
    out = []
    for r in range(g.world):
        r_sizes, r_scores, r_masks = g.meta(r)
        at = 0
        for k in range(len(r_sizes)):
            s = int(r_sizes[k])
            soa = LazyTreeSoA(block, r, at, s) if (r != g.rank or dev_payload is not None) else trees[k]
            out.append((soa, np.float64(r_scores[k]), r_masks[k] if g.mask_bytes else None))
            at += s
    return out, g


class LazyTreeSoA:
    """TreeSoA interface over a slice of a GatheredBlock (materialised on first read)."""

    def __init__(self, block: GatheredBlock, rank: int, start: int, size: int):
        self._block, self._rank, self._start, self._size = block, rank, start, size

    @property
    def n_nodes(self) -> int:
        return self._size

    @property
    def nodes(self):
        return self._block.host()[self._rank][0][self._start : self._start + self._size]

    @property
    def counts(self):
        return self._block.host()[self._rank][1][self._start : self._start + self._size]


class PackedTreeSoA:
    """TreeSoA interface over a slice of a PackRecorder's device buffers (copied to the host on first read)."""

    def __init__(self, nodes_t, counts_t, start: int, size: int, n_classes: int):
        self._nodes_t, self._counts_t, self._start, self._size, self._C = nodes_t, counts_t, start, size, n_classes
        self._host = None

    @property
    def n_nodes(self) -> int:
        return self._size

    def _fetch(self):
        if self._host is None:
            nb, cb = NODE_DTYPE.itemsize, self._C * 4
            n = self._nodes_t[self._start * nb : (self._start + self._size) * nb].cpu().numpy().view(NODE_DTYPE)
            c = self._counts_t[self._start * cb : (self._start + self._size) * cb].cpu().numpy().view(np.int32).reshape(self._size, self._C)
            self._host = (n, c)
            self._nodes_t = self._counts_t = None
        return self._host

    @property
    def nodes(self):
        return self._fetch()[0]

    @property
    def counts(self):
        return self._fetch()[1]


class PackRecorder:
    """NCCL path: the survivors of every launch are packed into a device tensor (by the engine's
    pack kernel), copied to the host from there, and the device copy is kept as the all-gather
    send buffer -- the node tables never make a host -> device trip."""

    def __init__(self, n_classes: int, device: int = 0):
        self.n_classes = n_classes
        self.device = device
        self.nodes, self.counts, self.sizes, self.n_trees = [], [], [], 0

    def __call__(self, batch, ids):
        import torch

        if batch.n_nodes is None:
            batch.results()
        sizes = batch.n_nodes[np.asarray(ids, dtype=np.int64)].astype(np.int64)
        total, C = int(sizes.sum()), self.n_classes
        dev = torch.device("cuda", self.device)
        tn = torch.empty(total * NODE_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        tc = torch.empty(total * C * 4, dtype=torch.uint8, device=dev)
        batch.pack_to_device(ids, tn.data_ptr(), tc.data_ptr())
        self.nodes.append(tn)
        self.counts.append(tc)
        self.sizes.append(sizes)
        self.n_trees += len(ids)
        # no host copy here: after the exchange this rank's trees are read, like everybody else's, from the
        # gathered block; until then a tree knows its size and where its tables sit on the device
        out, at = [], 0
        for s in sizes:
            out.append(PackedTreeSoA(tn, tc, at, int(s), self.n_classes))
            at += int(s)
        return out


def all_gather_trees(trees: Sequence[TreeSoA], scores: Sequence[float], n_classes: int, used_masks=None, dev_payload=None):
    """All-gather variable-length survivor tables (sizes first, then the padded payload).
    Returns [(soa, score, used_mask)] in rank order; this rank's own trees are returned as they
    are, the others as lazy views of the gathered (device-resident) block."""
    import torch

    td = _td()
    world, rank = td.get_world_size(), td.get_rank()
    dev = _comm_device()
    sizes = np.array([t.n_nodes for t in trees], dtype=np.int64)
    n_local_nodes = int(sizes.sum())
    if dev_payload is None:
        nodes = np.concatenate([t.nodes for t in trees]) if len(trees) else np.zeros(0, dtype=NODE_DTYPE)
        counts = np.concatenate([t.counts for t in trees]) if len(trees) else np.zeros((0, n_classes), dtype=np.int32)
    sc = np.array([float(s) for s in scores], dtype=np.float64)
    nmask = 0 if used_masks is None else used_masks.shape[1]
    meta = torch.tensor([len(trees), n_local_nodes], dtype=torch.int64, device=dev)
    metas_t = [torch.zeros_like(meta) for _ in range(world)]
    td.all_gather(metas_t, meta)
    metas = [tuple(int(x) for x in m.cpu().numpy()) for m in metas_t]
    max_trees = max(m[0] for m in metas)
    max_nodes = max(m[1] for m in metas)

    def padded(arr: np.ndarray, nbytes: int):
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        buf = torch.zeros(max(nbytes, 1), dtype=torch.uint8, device=dev)
        if len(raw):
            buf[: len(raw)] = torch.from_numpy(raw).to(dev)
        return buf

    def gather(arr: np.ndarray, nbytes: int):
        mine = padded(arr, nbytes)
        out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=dev)
        td.all_gather_into_tensor(out, mine)
        return out.view(world, mine.numel())

    # small per-tree metadata travels packed per field: [sizes | scores | masks], each padded to max_trees
    def field(arr, width):
        buf = np.zeros(max_trees * width, dtype=np.uint8)
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        buf[: len(raw)] = raw
        return buf

    packed = np.concatenate([field(sizes, 8), field(sc, 8)] + ([field(used_masks, nmask)] if nmask else [np.zeros(0, dtype=np.uint8)]))
    def gather_dev(t, nbytes: int):
        mine = torch.zeros(max(nbytes, 1), dtype=torch.uint8, device=dev)
        mine[: t.numel()] = t
        out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=dev)
        td.all_gather_into_tensor(out, mine)
        return out.view(world, mine.numel())

    g_small = gather(packed, len(packed)).cpu().numpy()
    if dev_payload is not None:
        g_nodes = gather_dev(dev_payload[0], max_nodes * NODE_DTYPE.itemsize)
        g_counts = gather_dev(dev_payload[1], max_nodes * n_classes * 4)
    else:
        g_nodes = gather(nodes, max_nodes * NODE_DTYPE.itemsize)
        g_counts = gather(counts, max_nodes * n_classes * 4)
    block = GatheredBlock(g_nodes, g_counts, metas, n_classes)
    out = []
    for r in range(world):
        nt = metas[r][0]
        row = g_small[r]
        r_sizes = row[: max_trees * 8].view(np.int64)[:nt]
        r_scores = row[max_trees * 8 : max_trees * 16].view(np.float64)[:nt]
        r_masks = row[max_trees * 16 :].reshape(max_trees, nmask)[:nt] if nmask else [None] * nt
        at = 0
        for k in range(nt):
            s = int(r_sizes[k])
            soa = trees[k] if r == rank else LazyTreeSoA(block, r, at, s)
            out.append((soa, np.float64(r_scores[k]), r_masks[k]))
            at += s
    return out


def _gather_forest(model, local, recorder=None):
    nf = len(model.feature_cols)
    payload = None
    if recorder is not None and recorder.n_trees == len(local) and recorder.nodes:
        import torch

        payload = (torch.cat(recorder.nodes), torch.cat(recorder.counts))
        # which features each tree splits on, read off the device copy of the node tables
        feat = payload[0].view(torch.int32).view(-1, NODE_DTYPE.itemsize // 4)[:, 0].long()
        sizes = torch.from_numpy(np.concatenate(recorder.sizes)).to(feat.device)
        owner = torch.repeat_interleave(torch.arange(len(local), device=feat.device), sizes)
        hit = torch.zeros(len(local) * nf, dtype=torch.uint8, device=feat.device)
        sel = feat >= 0
        hit[owner[sel] * nf + feat[sel]] = 1
        masks = hit.view(len(local), nf).cpu().numpy()
    else:
        masks = np.zeros((len(local), nf), dtype=np.uint8)
        pos = {name: j for j, name in enumerate(model.feature_cols)}
        for i, t in enumerate(local):
            for f in t.used_features:
                masks[i, pos[f]] = 1
    if _td().get_backend() == "nccl":
        # NCCL: the exchange is the C ABI's (rfmc_allgather_survivors), device to device.  Every rank takes
        # this branch; a rank whose tables are only on the host uploads them first.
        import torch

        dev_id = int(getattr(model, "device", 0) or 0)
        if payload is None and len(local):
            where = torch.device("cuda", dev_id)
            nodes = np.concatenate([t.soa.nodes for t in local])
            counts = np.concatenate([t.soa.counts for t in local])
            payload = (torch.from_numpy(np.ascontiguousarray(nodes).view(np.uint8).reshape(-1)).to(where),
                       torch.from_numpy(np.ascontiguousarray(counts).view(np.uint8).reshape(-1)).to(where))
        gathered, handle = all_gather_trees_abi([t.soa for t in local], [t.survived_score for t in local], len(model.class_vals),
                                                masks, payload, dev_id)
        model._last_gather = handle
    else:
        gathered = all_gather_trees([t.soa for t in local], [t.survived_score for t in local], len(model.class_vals), masks, payload)
    forest = []
    for soa, score, mask in gathered:
        t = model._wrap(soa, score)
        t.features = model.feature_cols
        # model.py:413-417 semantics (flat.reference_used_features); every class of the balanced bootstrap
        # ends up in some leaf, so the leaf labels of a fitted tree are the class labels
        if mask is not None:
            t._used_mask = (mask, model.feature_cols, model.class_vals)  # used_features is derived on first access
        else:
            t.used_features = reference_used_features(model.feature_cols, [], model.class_vals, False)
        forest.append(t)
    return forest


def _install_recorder(model):
    if _td().get_backend() != "nccl":
        return None
    rec = PackRecorder(len(model.class_vals), int(getattr(model, "device", 0) or 0))
    model._tls.pack_hook = rec
    return rec


def _merge_recorders(recs):
    """Recorders of the host streams of one rank, in worker order (the order of the local forest)."""
    recs = [r for r in recs if r is not None]
    if not recs:
        return None
    out = PackRecorder(recs[0].n_classes, recs[0].device)
    for r in recs:
        out.nodes += r.nodes
        out.counts += r.counts
        out.sizes += r.sizes
        out.n_trees += r.n_trees
    return out


def _fit_local(model, n_local: int, host_streams: int, seeds):
    """This rank's slots: the serial loop, or ``host_streams`` independent host RNG streams."""
    if host_streams <= 1 or n_local < 2:
        recorder = _install_recorder(model)
        try:
            local = model._fit_slots(n_local) if n_local > 0 else []
        finally:
            model._tls.pack_hook = None
        return local, recorder
    nccl = _td().get_backend() == "nccl"
    dev = int(getattr(model, "device", 0) or 0)
    factory = (lambda: PackRecorder(len(model.class_vals), dev)) if nccl else None
    local = model._fit_slots_streams(n_local, list(seeds), factory)
    return local, _merge_recorders(model._stream_hooks)


def _agree_or_raise(error: Optional[BaseException]):
    """Every rank learns whether some rank failed before the collective: a rank that died in its local
    fit must not leave the others blocked in the all-gather."""
    import torch

    td = _td()
    flag = torch.tensor([1 if error is not None else 0], dtype=torch.int32, device=_comm_device())
    td.all_reduce(flag, op=td.ReduceOp.MAX)
    if error is not None:
        raise error
    if int(flag.item()):
        raise RuntimeError("rank-sharded fit: another rank failed in its local fit")


def stream_seeds(rank_seed: int, host_streams: int):
    """Seeds of the host RNG streams of one rank (worker w of rank r == the serial algorithm under it)."""
    return [(int(rank_seed) * 1009 + 7919 * w) % (2**31 - 1) for w in range(host_streams)]


def fit_sharded(model, rank_seeds: Optional[Sequence[int]] = None, host_streams: int = 1):
    """Rank-sharded fit; every rank returns the full forest (list of DecisionTreeMC, rank order).
    Without explicit ``rank_seeds`` the ranks seed themselves ``base + rank`` with ``base`` drawn from the
    global numpy stream (ranks started under one np.random.seed would otherwise grow R copies of the
    same trees).  ``host_streams`` > 1: the rank's slots are shared over that many independent host RNG
    streams (seeds: ``stream_seeds``), worker order inside the rank."""
    rank, world = rank_world()
    lo, hi = shard_bounds(model.n_trees, rank, world)
    seeds = rank_seeds if rank_seeds is not None else getattr(model, "rank_seeds", None)
    if seeds is None:
        base = int(np.random.randint(0, 2**31 - 1 - world))
        seeds = [base + r for r in range(world)]
    np.random.seed(seeds[rank])
    _pyrandom.seed(seeds[rank])
    host_streams = max(1, min(int(host_streams or 1), hi - lo))
    local, recorder, error = [], None, None
    try:
        local, recorder = _fit_local(model, hi - lo, host_streams, stream_seeds(seeds[rank], host_streams))
    except BaseException as e:  # reported to every rank below
        error = e
    _agree_or_raise(error)
    from time import perf_counter

    t0 = perf_counter()
    forest = _gather_forest(model, local, recorder)
    if isinstance(getattr(model, "fit_stats", None), dict):
        model.fit_stats["t_gather"] = perf_counter() - t0
    return forest


def fit_sharded_local(model, n_local: int, host_streams: int = 1, seeds=None):
    """Weak-scaling variant used by bench.py: every rank grows ``n_local`` slots -- from its current
    RNG state, or over ``host_streams`` independent host streams seeded by ``seeds`` -- then the
    survivors are all-gathered (rank order, worker order inside a rank)."""
    if not is_distributed():
        if host_streams > 1:
            return model._fit_slots_streams(n_local, list(seeds))
        return model._fit_slots(n_local)
    local, recorder, error = [], None, None
    try:
        local, recorder = _fit_local(model, n_local, host_streams, seeds)
    except BaseException as e:
        error = e
    _agree_or_raise(error)
    from time import perf_counter

    t0 = perf_counter()
    forest = _gather_forest(model, local, recorder)
    if isinstance(getattr(model, "fit_stats", None), dict):
        model.fit_stats["t_gather"] = perf_counter() - t0
    return forest


def predict_sharded(model, frame, want_probs: bool):
    """Row-sharded voting: this rank's contiguous slice of ``frame``; no collective."""
    rank, world = rank_world()
    lo, hi = shard_bounds(len(frame), rank, world)
    part = frame.iloc[lo:hi]
    return (lo, hi), (model.testForestProbs(part) if want_probs else model.testForest(part))
