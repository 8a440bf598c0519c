"""``RandomForestMC``: the reference's training API (model.py:90-404) driving the CUDA engine.

What stays on the host, bit-for-bit as in the reference:
  * ``process_dataset`` semantics (model.py:162-195) -- plus the one-time rank-code encode + upload;
  * every random draw: ``split_train_val`` (model.py:198-210, numpy's global legacy RNG) and
    ``sampleFeats`` (model.py:213-219, numpy + CPython ``random``), in the reference's call order;
  * the Monte-Carlo keep/discard policy of ``survivedTree`` (model.py:317-348).

What moves to the GPU: growing every candidate (``plantTree``) and scoring it on its hold-out rows
(``validationTree``), for many slots and candidates per launch.

Batching without changing the RNG streams: how many ``sampleFeats`` draws a slot
consumes depends on the scores, so the planner draws a batch of slots *speculatively* (assuming
each slot uses all ``max_discard_trees`` candidates, or one when trees mostly pass), snapshots both
RNG states after every draw, lets the GPU grow and score everything, replays the reference's
policy over the returned scores and -- at the first slot that stopped earlier or later than
assumed -- rewinds both RNGs to the snapshot after that slot's last consumed draw and replans from
there.  The committed sequence of draws is therefore exactly the serial reference's.
"""

from __future__ import annotations

import logging as log
import os
import sys
import threading
import random as _pyrandom
from sys import getrecursionlimit
from typing import List, Optional, Tuple

import numpy as np
import pandas as pd

from . import engine
from .encode import EncodedDataset, LazyColumns
from .flat import encode_frame, forest_tables_from_dicts
from .forest import BaseRandomForestMC
from .tree import DecisionTreeMC


class MissingValuesNotFound(Exception):
    def __init__(
        self,
        message="Dataset or row without missing values! Please, give a dataset or row with missing values using 'NaN'.",
    ):
        super().__init__(message)


class DatasetNotFound(Exception):
    def __init__(self, message="Dataset not found! Please, give a dataset for functions fit() or process_dataset()."):
        super().__init__(message)


class DictValuesAllFeaturesMissing(Exception):
    def __init__(
        self,
        message="All features in the given dictionary 'dict_values' are not found int he trained model (forest).",
    ):
        super().__init__(message)


def _usable_cores() -> int:
    """Cores this process may run on (taskset / cpuset aware), not the cores of the box."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


class _SlotJudge:
    """The keep/discard state machine of one ``survivedTree`` call (model.py:319-342)."""

    def __init__(self, th_start, delta_th, max_discard_trees, get_best_tree):
        self.threshold = th_start
        self.delta_th = delta_th
        self.max_discard = max_discard_trees
        self.get_best = get_best_tree
        self.failures = 0
        self.best = None
        self.best_score = 0.0
        self.choice = None
        self.done = False
        self.used = 0

    def feed(self, cand, score) -> bool:
        self.used += 1
        if score < self.threshold:
            if self.get_best and self.best_score < score:
                self.best, self.best_score = cand, score
        else:
            self.best_score = score
            self.choice = cand
            self.done = True
            return True
        self.failures += 1
        if self.failures >= self.max_discard:
            if self.get_best:
                self.choice = self.best  # None when every score was 0.0: the reference raises here
                self.done = True
                return True
            self.threshold -= self.delta_th
            log.info("New threshold for drop: {:.4f}".format(self.threshold))
        return False


class _DeviceSlotState:
    """numpy's stream state before one slot of a device draw plan; read back (624 words) only when a
    mis-speculated plan has to be rewound to that slot."""

    def __init__(self, dev, slot, template):
        self.dev, self.slot, self.template = dev, slot, template

    def resolve(self):
        key, pos = self.dev.state_at(self.slot)
        t = self.template
        return (t[0], key, pos, t[3], t[4])


class _PlanList(list):
    """The slots of one plan; ``dev`` is the device draw plan that owns their row ids (or None).  ``rows`` /
    ``feats`` / ``spec`` hold the whole plan's host arrays as the draw calls returned them (row ids [n_slots, n],
    feature ids back to back, candidates per slot), so a launch needs no per-slot Python work."""

    dev = None
    rows = None
    feats = None
    spec = 0


class RandomForestMC(BaseRandomForestMC):
    def __init__(
        self,
        n_trees: int = 16,
        target_col: str = "target",
        batch_train_pclass: int = 10,
        batch_val_pclass: int = 10,
        max_discard_trees: int = 10,
        delta_th: float = 0.1,
        th_start: float = 1.0,
        get_best_tree: bool = True,
        min_feature: Optional[int] = None,
        max_feature: Optional[int] = None,
        th_decease_verbose: bool = False,
        temporal_features: bool = False,
        split_with_replace: bool = False,
        max_depth: Optional[int] = None,
        min_samples_split: int = 1,
        got_best_tree_verbose: bool = False,
        threaded_fit: bool = False,
    ) -> None:
        super().__init__(
            n_trees=n_trees,
            target_col=target_col,
            min_feature=min_feature,
            max_feature=max_feature,
            temporal_features=temporal_features,
        )
        self.split_with_replace = split_with_replace
        if th_decease_verbose:
            log.basicConfig(level=log.INFO)
        self.batch_train_pclass = batch_train_pclass
        self.batch_val_pclass = batch_val_pclass
        self._N = batch_train_pclass + batch_val_pclass
        self.th_start = th_start
        self.delta_th = delta_th
        self.max_discard_trees = max_discard_trees
        self.dataset = None
        self.dataset_numpy = None
        self.max_depth = getrecursionlimit() if max_depth is None else int(max_depth)
        self.min_samples_split = int(min_samples_split)
        self.got_best_tree_verbose = got_best_tree_verbose
        self.get_best_tree = get_best_tree
        self.threaded_fit = threaded_fit
        # engine state
        self.encoded: Optional[EncodedDataset] = None
        self._dev_dataset: Optional[engine.DeviceDataset] = None
        self.fit_stats = {}
        self._tls = threading.local()
        self.max_candidates_per_launch = 4096
        self.max_batch_bytes = 24 << 30
        # resolve helpers of the host RNG walk: the sequential part (MT19937 blocks + rejection) is the
        # floor; two helpers keep up with it.  A producer thread for the blocks is opt-in only: since
        # the blocks are generated with AVX-512 the cache-to-cache hand-off costs more than it saves.
        ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        self.rng_threads = max(0, min(2, _usable_cores() // ranks_here - 1))
        self.rng_producer = os.environ.get("RFMC_RNG_PRODUCER", "0") == "1"
        self.gil_switch_interval = float(os.environ.get("RFMC_GIL_SWITCH_INTERVAL", "1e-4"))  # seconds, host-stream fits only
        # numpy's stream can be walked on the GPU (csrc/rng_walk.cu: same rows, counts and stream states, row
        # ids never leave HBM).  Measured on a B200 next to 32 AVX-512 host cores the host walk
        # (csrc/host_rng.cpp) is still the faster one, so the device walk is opt-in;
        # replays, single-slot calls and draws wider than the device resolve always use the host walk.
        # True / False, or None = decide per fit (device_draws_on): on when this process is one of several ranks on
        # the box and its share of the host cores is below six -- there the device walk is the faster one (2 ranks x 4
        # cores: 61.9 against 85.8 ms per 1,024 C2b candidates, profiles/r2b_experiments.md) -- off otherwise.
        env_dd = os.environ.get("RFMC_DEVICE_DRAWS")
        self.device_draws = (env_dd == "1") if env_dd in ("0", "1") else None
        self.device_draw_auto_max_rows = 2_000_000  # frames the automatic choice has been measured on
        self.device_draw_min_slots = 2
        # software pipelining of the fit loop (_fit_slots): plans of more than this many slots are cut into
        # sub-plans so that the kernels and copies of one hide behind the draws of the next; 0 = one plan at a time
        self.pipeline_slots = int(os.environ.get("RFMC_PIPELINE_SLOTS", "0"))
        self.pipeline_min_words = int(os.environ.get("RFMC_PIPELINE_MIN_WORDS", "2000000"))

    # ---------------------------------------------------------------------------------------------
    # process_dataset (model.py:162-195)
    # ---------------------------------------------------------------------------------------------
    def process_dataset(self, dataset: pd.DataFrame) -> None:
        dataset[self.target_col] = dataset[self.target_col].astype(str)
        feature_cols = [col for col in dataset.columns if col != self.target_col]
        numeric_cols = dataset.select_dtypes([np.number]).columns.to_list()
        categorical_cols = list(set(feature_cols) - set(numeric_cols))
        type_of_cols = {col: "numeric" for col in numeric_cols}
        type_of_cols.update({col: "categorical" for col in categorical_cols})
        if self.min_feature is None:
            self.min_feature = 2
        if self.max_feature is None:
            self.max_feature = len(feature_cols)
        self.numeric_cols = numeric_cols
        self.feature_cols = feature_cols
        self.type_of_cols = type_of_cols
        dataset = dataset.dropna()
        log.warning("Rows with missing values were dropped from the dataset.")
        self.class_vals = dataset[self.target_col].unique().tolist()
        if self.temporal_features and (not self.validFeaturesTemporal()):
            self.temporal_features = False
            log.warning("Temporal features ordering disable: you do not have all orderable features!")
        min_class = dataset[self.target_col].value_counts().min()
        self._N = min(self._N, min_class)
        self.dataset_numpy = LazyColumns(dataset, numeric_cols)
        self.n_samples = len(dataset)
        # engine: encode once, upload once
        if self.min_samples_split < 1:
            raise ValueError("min_samples_split must be >= 1")
        if self._dev_dataset is not None:
            self._dev_dataset.close()
            self._dev_dataset = None
        self.encoded = None
        if len(dataset) == 0 or len(feature_cols) == 0:
            # Nothing to encode.  The reference accepts such a frame here and only trips inside its first
            # tree slot; _reference_preflight reproduces where and how.
            return
        self.encoded, self._dev_dataset = engine.encode_dataset_device(
            self.dataset_numpy, feature_cols, numeric_cols, self.target_col, self.class_vals, self.device
        )

    # ---------------------------------------------------------------------------------------------
    # host RNG draws, exactly the reference's (model.py:198-219)
    # ---------------------------------------------------------------------------------------------
    def _np_rng(self):
        """numpy's global legacy generator (the reference's), or this thread's own RandomState when
        fitParallel runs several host streams."""
        return getattr(self._tls, "np_rng", None) or np.random

    def _py_rng(self):
        return getattr(self._tls, "py_rng", None) or _pyrandom

    def split_train_val(self) -> Tuple[np.ndarray, np.ndarray]:
        """model.py:198-210.  ``np.random.choice(class_indices, _N, replace=False)`` per class in
        ``class_vals`` order, drawn from numpy's global legacy stream by the host restatement in
        csrc/host_rng.cpp (same indices, same end state as numpy: tests/test_host_rng.py).  The
        per-class row lists ``indices[target_vals == val]`` are cached: building them uses no RNG."""
        st = self._np_rng().get_state()
        key, pos = st[1].copy(), int(st[2])
        train, val = [], []
        for class_indices in self.encoded.class_rows:
            idx, pos = engine.legacy_choice_indices(key, pos, len(class_indices), self._N)
            selected = class_indices[idx]
            train.append(selected[: self.batch_train_pclass])
            val.append(selected[self.batch_train_pclass :])
        self._np_rng().set_state((st[0], key, pos, st[3], st[4]))
        return np.concatenate(train), np.concatenate(val)

    def sampleFeats(self, feature_cols: List[str]) -> List[str]:
        n_samples = self._np_rng().randint(self.min_feature, self.max_feature + 1)
        out = self._py_rng().sample(feature_cols, min(len(feature_cols), n_samples))
        if not self.temporal_features:
            return out
        return sorted(out, key=lambda x: int(x.split("_")[-1]))

    # ---------------------------------------------------------------------------------------------
    # single-candidate API (model.py:264-314), one launch each
    # ---------------------------------------------------------------------------------------------
    def _reference_preflight(self):
        """Configurations the reference only trips over inside its first tree slot (model.py:318-326:
        split_train_val -> sampleFeats -> plantTree -> validationTree).  Same exception, raised after
        the same draws, so both RNG streams end where the reference leaves them:
          * min_feature > max_feature (e.g. a frame with a single feature column and the default
            min_feature = 2): np.random.randint raises ValueError inside sampleFeats, after the row draw;
          * no training rows (batch_train_pclass <= 0, or a frame with no rows left after dropna) or no
            hold-out rows (batch_val_pclass == 0, or a class smaller than batch_train_pclass):
            split_train_val returns np.array([]) -- float64 -- and plantTree / validationTree die
            indexing with it (model.py:210, 267, 306): IndexError."""
        if self.dataset_numpy is None:
            return
        no_rows = self.encoded is None and self.n_samples == 0
        lo, hi = int(self.min_feature), int(self.max_feature) + 1
        no_train = no_rows or int(self.batch_train_pclass) <= 0
        no_val = not no_rows and int(self._N) <= int(self.batch_train_pclass)
        if not (lo >= hi or no_train or no_val or self.encoded is None):
            return
        if self.encoded is not None:
            self.split_train_val()
        elif not no_rows:  # rows but no feature columns: the row draw still happens (on the label column only)
            st = self._np_rng()
            target = self.dataset_numpy[self.target_col]
            for val in self.class_vals:
                st.choice(np.flatnonzero(target == val), int(self._N), replace=False)
        self.sampleFeats(self.feature_cols[:])  # numpy's own ValueError when min_feature > max_feature
        raise IndexError("arrays used as indices must be of integer (or boolean) type")

    def _require_dataset(self):
        if self.dataset_numpy is not None and self._dev_dataset is None and self.n_samples and self.feature_cols:
            # a model that was pickled / deep-copied travels without its device handles: encode again
            self.encoded, self._dev_dataset = engine.encode_dataset_device(
                self.dataset_numpy, self.feature_cols, self.numeric_cols, self.target_col, self.class_vals, self.device
            )
        if self.dataset_numpy is None or self._dev_dataset is None:
            raise DatasetNotFound

    # -- pickling / copy.deepcopy (the reference model is a plain picklable object; joblib.dump(model) is common) --
    def __getstate__(self):
        state = self.__dict__.copy()
        for key in ("_tls", "_dev_dataset", "_forest_dev", "_forest_f32", "_stream_hooks", "_forest_mutex", "_last_gather"):
            state.pop(key, None)
        state["_forest_key"] = None
        state["_forest_tables"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._tls = threading.local()
        self._dev_dataset = None
        self._forest_dev = None
        self._forest_f32 = None

    def _feat_ids(self, feature_list: List[str]) -> List[int]:
        pos = {name: j for j, name in enumerate(self.feature_cols)}
        return [pos[f] for f in feature_list]

    def device_draws_on(self) -> bool:
        """Whether plans are drawn by the device walk of numpy's stream (``device_draws``; None = automatic)."""
        if self.device_draws is not None:
            return bool(self.device_draws)
        ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        if ranks_here < 2 or self.encoded is None:
            return False
        rows = sum(len(r) for r in self.encoded.class_rows)
        return _usable_cores() // ranks_here < 6 and rows <= self.device_draw_auto_max_rows

    def _draw_plan(self, n_plan: int, spec: int, stream=None):
        """All draws of ``n_plan`` slots with ``spec`` candidates each, in the reference's order.
        numpy's stream (rows per class, then one feature count per candidate) is walked in ONE call:
        on the GPU (``rfmc_draw_plan_device``; the row ids stay in HBM for the grower) or, for plans
        the device resolve does not take, on the host (``rfmc_legacy_draw_plan``).  CPython's
        ``random.sample`` -- a separate stream -- then picks and orders the features.  Entry layout as
        ``_draw_slot``, plus the device plan and slot that hold the row ids."""
        if self.min_feature > self.max_feature:
            raise ValueError("low >= high")  # what np.random.randint raises in the reference
        st = self._np_rng().get_state()
        dev = None
        if (self.device_draws_on() and n_plan >= self.device_draw_min_slots and int(self._N) <= engine.DRAW_MAX_PER_CLASS
                and int(self.max_feature) - int(self.min_feature) < 2**31 - 2):
            try:
                dev = engine.DrawPlan(self._dev_dataset, self.encoded.class_rows, st[1], int(st[2]), n_plan, int(self._N),
                                      int(self.batch_train_pclass), spec, int(self.min_feature), int(self.max_feature) + 1, stream)
                try:
                    K, key, pos = dev.finish()
                except BaseException:
                    dev.close()
                    raise
            except engine.EngineError:
                # Both walks restate the same stream and nothing has been committed yet (numpy's state is set below):
                # when the device walk was only the automatic choice, a plan it gives up on (it verifies itself and
                # fails loudly) is drawn by the host walk instead -- same rows, same counts, same stream state.
                if self.device_draws is not None:
                    raise
                log.warning("device draw plan failed; drawing this plan on the host", exc_info=True)
                dev = None
                self.device_draw_fallbacks = getattr(self, "device_draw_fallbacks", 0) + 1
            T = V = ksnap = psnap = None
        if dev is None:
            key = st[1].copy()
            T, V, K, ksnap, psnap, pos = engine.legacy_draw_plan(
                key, int(st[2]), n_plan, self.encoded.class_rows, int(self._N), int(self.batch_train_pclass), spec,
                int(self.min_feature), int(self.max_feature) + 1, getattr(self._tls, "rng_flags", self.rng_threads | (0x100 if self.rng_producer else 0)),
            )
        self._np_rng().set_state((st[0], key, pos, st[3], st[4]))
        cols = self.feature_cols
        py_rng = self._py_rng()
        py_state0 = py_rng.getstate()
        ks = np.minimum(K, len(cols)).astype(np.int32)
        if not self.temporal_features:
            # CPython's stream walked in one host call as well (engine.pyrandom_sample_plan)
            flat = engine.pyrandom_sample_plan(py_rng, len(cols), ks)
            per_cand = engine.FlatFeatureLists(flat, ks)
        else:
            per_cand = []
            where = {name: j for j, name in enumerate(cols)}
            for k in ks.reshape(-1):
                out = sorted(py_rng.sample(cols, int(k)), key=lambda x: int(x.split("_")[-1]))
                per_cand.append(np.array([where[f] for f in out], dtype=np.int32))
        plan = _PlanList()
        plan.dev = dev
        plan.rows = (T, V)
        plan.feats = per_cand if isinstance(per_cand, engine.FlatFeatureLists) else None
        plan.spec = spec
        per_cand = list(per_cand)
        for s in range(n_plan):
            if dev is not None:
                start = (_DeviceSlotState(dev, s, st), ("replay", py_state0, K, spec, s))
                plan.append((None, None, per_cand[s * spec : (s + 1) * spec], (start, True), (dev, s)))
            else:
                start = ((st[0], ksnap[s], int(psnap[s]), st[3], st[4]), ("replay", py_state0, K, spec, s))
                plan.append((T[s], V[s], per_cand[s * spec : (s + 1) * spec], (start, True), None))
        return plan

    @staticmethod
    def _slot_rows(p):
        """(idx_T, idx_V) of a plan entry; a device plan's rows are copied to the host only here."""
        if p[0] is None and len(p) > 4 and p[4] is not None:
            dev, s = p[4]
            T, V = dev.indices(s, 1)
            return T[0], V[0]
        return p[0], p[1]

    def _wrap(self, soa, score=None) -> DecisionTreeMC:
        t = DecisionTreeMC.from_soa(soa, self.encoded, self.type_of_cols, self.class_vals)
        t.survived_score = score
        return t

    def plantTree(self, indices: np.ndarray, feature_list: List[str]) -> DecisionTreeMC:
        self._require_dataset()
        idx = np.asarray(indices).reshape(1, -1)
        b = engine.GrowBatch(
            self._dev_dataset, idx, np.zeros((1, 0), dtype=np.int32), [0], [self._feat_ids(feature_list)],
            self.max_depth, self.min_samples_split,
        ).run()
        try:
            return self._wrap(b.fetch([0])[0])
        finally:
            b.close()

    def validationTree(self, Tree: DecisionTreeMC, indices: np.ndarray):
        """model.py:296-314: hold-out accuracy of one tree, through the voting kernel."""
        self._require_dataset()
        indices = np.asarray(indices)
        tables = forest_tables_from_dicts([Tree.data], [1.0], self.class_vals, self.feature_cols)
        frame = pd.DataFrame({c: self.dataset_numpy[c][indices] for c in self.feature_cols})
        dev = engine.DeviceForest(tables, self.device)
        try:
            codes, absent = encode_frame(tables, frame)
            _, labels = dev.predict(codes, absent, False, False, False, True) if len(indices) else (None, np.zeros(0, dtype=np.int32))
        finally:
            dev.close()
        y_val = self.dataset_numpy[self.target_col][indices]
        y_pred = [self.class_vals[k] for k in labels]
        return np.mean(y_val == y_pred)

    def tree2feats(self, Tree) -> List[str]:
        """model.py:413-417, quirks included: see flat.reference_used_features."""
        from .flat import reference_used_features_of_dict, reference_used_features_of_soa

        soa = getattr(Tree, "soa", None)
        if soa is not None and Tree._data is None and Tree._enc is not None:
            return reference_used_features_of_soa(self.feature_cols, soa, Tree._enc.class_sorted)
        return reference_used_features_of_dict(self.feature_cols, Tree.data)

    # ---------------------------------------------------------------------------------------------
    # the batched, speculative fit loop
    # ---------------------------------------------------------------------------------------------
    def _snapshot(self):
        return self._np_rng().get_state(), self._py_rng().getstate()

    def _restore(self, snap):
        np_state = snap[0].resolve() if isinstance(snap[0], _DeviceSlotState) else snap[0]
        self._np_rng().set_state(np_state)
        py_rng = self._py_rng()
        py = snap[1]
        if isinstance(py, tuple) and len(py) == 5 and py[0] == "replay":
            # CPython's stream is snapshotted once per plan; a slot's state is recovered by replaying
            # the plan's random.sample calls up to that slot (they are pure functions of the stream)
            _, state0, counts, spec, upto = py
            py_rng.setstate(state0)
            cols = self.feature_cols
            for t in range(upto):
                for k in range(spec):
                    py_rng.sample(cols, min(len(cols), int(counts[t, k])))
        else:
            py_rng.setstate(py)

    def _draw_slot(self, n_cands: int, idx=None):
        """One slot's draws: rows once (unless continuing a slot), then ``n_cands`` feature lists.
        Returns (idx_T, idx_V, [feature id lists], RNG snapshot taken BEFORE the slot's draws)."""
        start = self._snapshot()
        if idx is None:
            idx = self.split_train_val()
        feats = [self._feat_ids(self.sampleFeats(self.feature_cols[:])) for _ in range(n_cands)]
        return idx[0], idx[1], feats, (start, idx is None)

    def _rewind_into_slot(self, p, n_used: int, fresh_rows: bool = True):
        """Put both RNGs where the serial reference is after ``n_used`` sampleFeats calls of slot
        ``p``: restore the snapshot taken before the slot and replay its draws (they are pure
        functions of the stream, so the replay reproduces them)."""
        start, _ = p[3]
        self._restore(start)
        if fresh_rows:
            self.split_train_val()
        for _ in range(n_used):
            self.sampleFeats(self.feature_cols[:])

    def _candidate_bytes(self) -> int:
        n_train = self.batch_train_pclass * len(self.class_vals) if self._N > self.batch_train_pclass else self._N * len(self.class_vals)
        # node table + counts + votes, plus the slot's gathered [feature][row] matrix (a slot may end up
        # with a single candidate)
        return max(1, n_train) * (2 * (24 + 4 * len(self.class_vals) + 1) + len(self.feature_cols) * self.encoded.width)

    def _grow_plan(self, plan, stream=None):
        """One launch for every (slot, candidate) of the plan.  Returns (batch, first cand id per slot)."""
        dev = getattr(plan, "dev", None)
        whole = getattr(plan, "feats", None)
        if whole is not None and plan.spec > 0 and len(whole) == len(plan) * plan.spec:
            # a plan as _draw_plan made it: the arrays of the draw calls go to the launch as they are
            spec = plan.spec
            slot_of = np.repeat(np.arange(len(plan), dtype=np.int32), spec)
            feats = whole
            first = list(range(0, len(plan) * spec, spec))
            rows = plan.rows
        else:
            slot_of, feats, first, rows = [], [], [], None
            for s, p in enumerate(plan):
                first.append(len(feats))
                for f in p[2]:
                    slot_of.append(s)
                    feats.append(f)
        if dev is not None:  # the row ids are already in HBM
            b = engine.GrowBatch(self._dev_dataset, None, None, slot_of, feats, self.max_depth, self.min_samples_split, stream, plan=dev)
        else:
            if rows is not None and rows[0] is not None:
                idx_T, idx_V = rows
            else:
                idx_T = np.stack([p[0] for p in plan])
                idx_V = np.stack([p[1] for p in plan])
            b = engine.GrowBatch(self._dev_dataset, idx_T, idx_V, slot_of, feats, self.max_depth, self.min_samples_split, stream)
        b.run()
        return b, first

    def _score(self, correct, n_val):
        """th_val = np.mean(y_val == y_pred) (model.py:314); NaN for an empty hold-out set."""
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.float64(correct) / np.float64(n_val) if n_val else np.float64("nan")

    def _scores(self, correct, n_val):
        """``_score`` for a whole launch: float64 division element by element, np.float64 scalars out."""
        c = np.asarray(correct).astype(np.float64)
        if not n_val:
            return np.full(len(c), np.nan, dtype=np.float64)
        return c / np.float64(n_val)

    def _fit_slots(self, n_slots: int, stream=None):
        """Grow ``n_slots`` surviving trees; the committed RNG stream equals the serial reference's.

        Plans are software-pipelined: while the GPU grows and scores plan A, the host already draws (and
        launches) plan B from the stream state A leaves behind if its speculation holds; A is then judged
        and its survivors fetched while the GPU works on B.  A plan that needs a rewind (a slot that used
        fewer or more draws than assumed) voids its successor: B is dropped, the streams are restored from
        A's snapshots, and drawing resumes there -- the committed draws are the reference's either way."""
        if n_slots > 0:
            self._reference_preflight()
        self._require_dataset()
        trees: List[DecisionTreeMC] = []
        stats = dict(candidates_grown=0, candidates_used=0, gpu_launches=0, rewinds=0, batches=0, plans_dropped=0,
                     t_draw=0.0, t_launch=0.0, t_wait=0.0, t_fetch=0.0)
        from time import perf_counter as _now

        K = max(1, int(self.max_discard_trees))
        cap = max(1, min(self.max_candidates_per_launch, self.max_batch_bytes // max(1, self._candidate_bytes())))
        state = dict(assume_pass=False, avg_commit=float("inf"))  # hypothesis on the candidates a slot consumes: K, or 1 when trees mostly pass
        recent_used: List[int] = []
        # pipelining pays when a plan's draws take about as long as its kernels (numpy's stream costs one word per
        # dataset row and slot); below that a second launch only adds latency (titanic-sized fits)
        pipe = max(0, int(self.pipeline_slots))
        if sum(len(r) for r in self.encoded.class_rows) * max(1, n_slots) < self.pipeline_min_words:
            pipe = 0

        def start(ahead: int):
            """Draw and launch the next plan, ``ahead`` slots being in flight already (0: none)."""
            spec = 1 if state["assume_pass"] else K
            left = n_slots - len(trees) - ahead
            n_plan = min(left, max(1, cap // spec))
            if state["avg_commit"] != float("inf"):
                n_plan = min(n_plan, max(4, int(4 * state["avg_commit"]) + 1))
            if pipe and n_plan > pipe:
                # sub-plans: about four per call, so that all but the last one's kernels and copies hide behind draws
                n_plan = max(pipe, min(n_plan, 8 * pipe, -(-(n_slots) // 4)))
            t0 = _now()
            plan = self._draw_plan(n_plan, spec, stream)
            t1 = _now()
            try:
                batch, first = self._grow_plan(plan, stream)
            except BaseException:
                if getattr(plan, "dev", None) is not None:
                    plan.dev.close()
                raise
            stats["t_draw"] += t1 - t0
            stats["t_launch"] += _now() - t1
            return plan, batch, first, spec

        def drop(rec):
            if rec is None:
                return
            plan, batch = rec[0], rec[1]
            batch.close()
            if getattr(plan, "dev", None) is not None:
                plan.dev.close()

        def finish(rec) -> bool:
            """Judge one plan, collect its survivors; True when every slot used exactly the draws assumed."""
            plan, batch, first, spec = rec
            t2 = _now()
            correct, _ = batch.results()
            stats["t_wait"] += _now() - t2
            stats["batches"] += 1
            stats["gpu_launches"] += batch.launches
            stats["candidates_grown"] += batch.n_cand
            chosen: List[Tuple[int, float]] = []
            resume = extend = None
            scores = self._scores(correct, batch.n_val)
            for s, p in enumerate(plan):
                judge = _SlotJudge(self.th_start, self.delta_th, self.max_discard_trees, self.get_best_tree)
                for k in range(spec):
                    if judge.feed(first[s] + k, scores[first[s] + k]):
                        break
                if not judge.done:
                    # needs more candidates than were drawn: everything drawn after this slot is
                    # off-stream; continue the slot from the state right after its last draw
                    resume, extend = (p, spec), (p, judge)
                    break
                if judge.choice is None:
                    # every candidate of the slot scored 0.0: the reference dies on Tree.survived_score
                    # (model.py:345) with both streams where this slot left them
                    self._rewind_into_slot(p, judge.used)
                    raise AttributeError("'NoneType' object has no attribute 'survived_score'")
                chosen.append((judge.choice, judge.best_score))
                stats["candidates_used"] += judge.used
                recent_used.append(judge.used)
                if judge.used != spec:
                    resume = (p, judge.used)
                    break
            t3 = _now()
            self._collect(batch, chosen, trees)
            stats["t_fetch"] += _now() - t3
            committed = len(chosen)
            clean = resume is None and extend is None
            if resume is not None and not (extend is None and resume[0] is plan[-1] and resume[1] == spec):
                stats["rewinds"] += 1
                self._rewind_into_slot(resume[0], resume[1])
            if extend is not None:
                trees.append(self._extend_slot(extend[0], extend[1], batch, stream, stats))
                recent_used.append(extend[1].used)
                committed += 1
            state["avg_commit"] = committed if state["avg_commit"] == float("inf") else 0.5 * state["avg_commit"] + 0.5 * committed
            if len(recent_used) >= 4:
                tail = recent_used[-16:]
                state["assume_pass"] = K > 1 and (sum(1 for u in tail if u == 1) * 2 > len(tail))
            return clean

        cur = nxt = None
        last_clean = True
        try:
            if n_slots > 0:
                cur = start(0)
            while cur is not None:
                # the successor is drawn before this plan is judged only while plans keep their promise
                if pipe and last_clean and len(trees) + len(cur[0]) < n_slots:
                    nxt = start(len(cur[0]))
                try:
                    last_clean = finish(cur)
                finally:
                    drop(cur)
                    cur = None
                if nxt is not None and not last_clean:
                    stats["plans_dropped"] += 1
                    drop(nxt)
                    nxt = None
                cur, nxt = nxt, None
                if cur is None and len(trees) < n_slots:
                    cur = start(0)
        finally:
            drop(cur)
            drop(nxt)
        for t in trees:
            t.features = self.feature_cols  # used_features is derived from the node table on first access
        self.fit_stats = stats
        self._tls.last_stats = stats
        return trees

    def _fit_slots_streams(self, n_slots: int, seeds, hook_factory=None):
        """``n_slots`` shared over ``len(seeds)`` independent host RNG streams, one thread each, all
        feeding this GPU.  Worker r is exactly the serial algorithm on its share of the slots under
        ``np.random.seed(seeds[r]); random.seed(seeds[r])`` (its own RandomState / Random objects);
        the forest is the concatenation in worker order -- the same contract as the rank-sharded
        fit, and the semantics of the reference's process pool (independent streams per worker)."""
        from .dist import shard_bounds

        k = len(seeds)
        parts, stats, errors = [None] * k, [None] * k, []
        self._stream_hooks = [hook_factory() if hook_factory is not None else None for _ in range(k)]

        def work(r):
            stream = 0
            try:
                stream = engine.stream_acquire(self.device)  # each host stream keeps its own launches in flight
                self._tls.pack_hook = self._stream_hooks[r]
                self._tls.np_rng = np.random.RandomState(int(seeds[r]))
                self._tls.py_rng = _pyrandom.Random(int(seeds[r]))
                # one resolve helper per stream while there are cores for it (the box's cores are shared by the
                # ranks of a multi-GPU run); no producer thread
                ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
                self._tls.rng_flags = 1 if 2 * k * ranks_here <= _usable_cores() else 0
                lo, hi = shard_bounds(n_slots, r, k)
                parts[r] = self._fit_slots(hi - lo, stream) if hi > lo else []
                stats[r] = getattr(self._tls, "last_stats", None)
            except BaseException as e:  # surfaced on the calling thread
                errors.append(e)
            finally:
                engine.stream_release(self.device, stream)

        # Every C call of a worker (draws, launch, wait, fetch) releases the GIL and has to win it back
        # on return; with CPython's default 5 ms switch interval a returning worker can wait that long
        # behind another worker's numpy glue, several times per batch.  A short interval while the
        # workers run turns those hand-offs into microseconds.
        switch = sys.getswitchinterval()
        sys.setswitchinterval(min(switch, self.gil_switch_interval))
        try:
            threads = [threading.Thread(target=work, args=(r,)) for r in range(k)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        finally:
            sys.setswitchinterval(switch)
        if errors:
            raise errors[0]
        total = {}
        for st in stats:
            for key, v in (st or {}).items():
                total[key] = total.get(key, 0) + v
        total["host_streams"] = k
        self.fit_stats = total
        return [t for part in parts for t in part]

    def _collect(self, batch, chosen, trees_out):
        ids = []
        for cid, _ in chosen:
            if cid is None:
                raise AttributeError("'NoneType' object has no attribute 'survived_score'")  # model.py:345
            ids.append(cid)
        hook = getattr(self._tls, "pack_hook", None)  # rank-sharded fit: keep the packed tables on the device too
        soas = hook(batch, ids) if (hook is not None and ids) else batch.fetch(ids)
        for soa, (_, score) in zip(soas, chosen):
            trees_out.append(self._wrap(soa, score))

    def _extend_slot(self, p, judge, batch, stream, stats) -> DecisionTreeMC:
        """A slot that outlived its speculation (a decaying threshold with get_best_tree=False, or a
        first candidate that failed under the trees-mostly-pass hypothesis): keep drawing feature
        lists for the same rows until the reference's policy stops."""
        K = max(1, int(self.max_discard_trees))
        kept = {}
        if judge.best is not None:
            kept[judge.best] = batch.fetch([judge.best])[0]
        it = 0
        rows = self._slot_rows(p)
        while True:
            it += 1
            _, _, feats, snap = self._draw_slot(K, idx=rows)
            ext = (rows[0], rows[1], feats, snap)
            b, _ = self._grow_plan([ext], stream)
            try:
                correct, _ = b.results()
                stats["gpu_launches"] += b.launches
                stats["candidates_grown"] += b.n_cand
                stop = None
                for k in range(K):
                    ref = ("ext", it, k)
                    done = judge.feed(ref, self._score(correct[k], b.n_val))
                    if judge.best == ref and ref not in kept:
                        kept = {ref: b.fetch([k])[0]}
                    if done:
                        stop = k
                        break
                if stop is None:
                    continue
                choice = judge.choice
                if choice is None:
                    self._rewind_into_slot(ext, stop + 1, fresh_rows=False)
                    raise AttributeError("'NoneType' object has no attribute 'survived_score'")  # model.py:345
                soa = kept.get(choice)
                if soa is None:
                    soa = b.fetch([choice[2]])[0]  # the candidate that met the threshold, in this launch
                if stop != K - 1:
                    stats["rewinds"] += 1
                    self._rewind_into_slot(ext, stop + 1, fresh_rows=False)
                stats["candidates_used"] += judge.used
                return self._wrap(soa, judge.best_score)
            finally:
                b.close()

    # ---------------------------------------------------------------------------------------------
    # public training API (model.py:317-404)
    # ---------------------------------------------------------------------------------------------
    def survivedTree(self, _=None) -> DecisionTreeMC:
        tree = self._fit_slots(1)[0]
        if self.got_best_tree_verbose:
            log.info("Got best tree: {:.4f}".format(tree.survived_score))
        return tree

    def fit(self, dataset: Optional[pd.DataFrame] = None, disable_progress_bar: bool = False) -> None:
        if dataset is not None:
            self.process_dataset(dataset)
        if self.dataset_numpy is None:
            raise DatasetNotFound
        Forest = self._fit_slots(self.n_trees)
        self.data.extend(Forest)
        self.survived_scores.extend([Tree.survived_score for Tree in Forest])

    def fitParallel(
        self,
        dataset: Optional[pd.DataFrame] = None,
        disable_progress_bar: bool = False,
        max_workers: Optional[int] = None,
    ):
        """The analogue of the reference's process pool (model.py:372-404): tree slots are sharded
        over independent RNG streams.  With an initialised torch.distributed group: one rank per GPU,
        survivors all-gathered.  Otherwise: ``max_workers`` host threads, each with its own seeded
        streams, all feeding this GPU (the draws -- the serial floor of ``fit`` -- run in parallel).
        Either way worker r == the serial algorithm under its seed, forest = concatenation."""
        if dataset is not None:
            self.process_dataset(dataset)
        if self.dataset_numpy is None:
            raise DatasetNotFound
        from . import dist

        if dist.is_distributed():
            Forest = dist.fit_sharded(self, host_streams=max_workers or 1)
        else:
            k = max_workers or min(16, _usable_cores())  # the reference's pool defaults to one worker per core
            k = max(1, min(int(k), self.n_trees))
            seeds = getattr(self, "worker_seeds", None)
            if seeds is None:  # derived from the global stream: reproducible under np.random.seed(s)
                seeds = [int(x) for x in self._np_rng().randint(0, 2**31 - 1, size=k)]
            Forest = self._fit_slots_streams(self.n_trees, list(seeds)[:k] if len(seeds) >= k else list(seeds))
        self.data.extend(Forest)
        self.survived_scores.extend([Tree.survived_score for Tree in Forest])

    # ---------------------------------------------------------------------------------------------
    # explainability helpers on top of used_features / per-tree votes (model.py:406-493)
    # ---------------------------------------------------------------------------------------------
    def sampleClass2trees(self, row, Class) -> List[DecisionTreeMC]:
        """Trees whose own vote for ``row`` is ``Class`` (model.py:406-407), via the leaf-export kernel."""
        if isinstance(row, dict):
            row = pd.Series(row)
        dev = self._device_forest()
        frame = pd.DataFrame([row.to_numpy(dtype=object)], columns=list(row.index))
        codes, absent = encode_frame(self._forest_tables, frame)
        payload, seen = dev.leaves(codes, absent)
        t = self._forest_tables
        votes = np.where(seen[0], t.vote_absent[payload[0]], t.vote[payload[0]])
        return [Tree for Tree, v in zip(self.data, votes) if self.class_vals[int(v)] == Class]

    @property
    def trees2depths(self) -> List[List[int]]:
        return [tree.depths for tree in self.data]

    def featCount(self, Forest=None):
        if Forest is None:
            Forest = self.data
        out = [len(Tree.used_features) for Tree in Forest]
        return (np.mean(out), np.std(out), min(out), max(out)), out

    def featImportance(self, Forest=None):
        if Forest is None:
            Forest = self.data
        n, seen = len(Forest), {}
        for tree in Forest:
            for feat in tree.used_features:
                seen[feat] = seen.get(feat, 0) + 1
        return {feat: c / n for feat, c in seen.items()}

    def featScoreMean(self, Forest=None):
        if Forest is None:
            Forest = self.data
        acc = {}
        for Tree, score in zip(Forest, self.survived_scores):
            for feat in Tree.used_features:
                acc.setdefault(feat, []).append(score)
        return {feat: np.mean(s) for feat, s in acc.items()}

    def featPairImportance(self, disable_progress_bar=False, Forest=None):
        from itertools import combinations

        if Forest is None:
            Forest = self.data
        n, out = len(Forest), {}
        for Tree in Forest:
            used = set(Tree.used_features)
            for pair in combinations(self.feature_cols, 2):
                if pair[0] in used and pair[1] in used:
                    out[pair] = out.get(pair, 0) + 1 / n
        return out

    def featCorrDataFrame(self, Forest=None) -> pd.DataFrame:
        N = len(self.feature_cols)
        matrix = np.zeros((N, N), dtype=np.float16)
        for feat, count in self.featImportance(Forest=Forest).items():
            i = self.feature_cols.index(feat)
            matrix[i][i] = count
        for pair, count in self.featPairImportance(Forest=Forest).items():
            a, b = self.feature_cols.index(pair[0]), self.feature_cols.index(pair[1])
            matrix[a][b], matrix[b][a] = count, count
        return pd.DataFrame(matrix, index=self.feature_cols, columns=self.feature_cols)

    # missing-value exploration (model.py:495-568): pandas reshaping around one predict_proba call
    @staticmethod
    def _fill_row_missing(row, dict_values):
        from .missing import fill_row

        return fill_row(row, dict_values)

    def _validationMissingValues(self, dict_values) -> None:
        from .missing import check_dict_values

        check_dict_values(self, dict_values)

    def _genFilledDataMissing(self, row_or_matrix, dict_values):
        from .missing import expand

        return expand(row_or_matrix, dict_values)

    def predictMissingValues(self, row_or_matrix, dict_values):
        from .missing import predict_missing_values

        return predict_missing_values(self, row_or_matrix, dict_values)

    def sampleClassFeatCount(self, row, Class):
        return self.featCount(self.sampleClass2trees(row=row, Class=Class))

    def sampleClassFeatImportance(self, row, Class):
        return self.featImportance(self.sampleClass2trees(row=row, Class=Class))

    def sampleClassFeatScoreMean(self, row, Class):
        return self.featScoreMean(self.sampleClass2trees(row=row, Class=Class))

    def sampleClassFeatPairImportance(self, row, Class):
        # As in the reference (model.py:473-476) the tree list is passed POSITIONALLY, i.e. as
        # `disable_progress_bar`: the result is the pair importance of the whole forest.
        return self.featPairImportance(self.sampleClass2trees(row=row, Class=Class))

    def sampleClassFeatCorrDataFrame(self, row, Class):
        return self.featCorrDataFrame(self.sampleClass2trees(row=row, Class=Class))
