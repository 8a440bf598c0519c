"""``BaseRandomForestMC``: the reference's forest container and inference API (forest.py:28-274)
with the row loops replaced by one launch of the voting kernel.

Public surface kept byte-compatible: ``predict`` / ``predict_proba`` /
``useForest`` / ``testForest`` / ``testForestProbs`` (+ the ``*Parallel`` spellings),
``setSoftVoting`` / ``setWeightedTrees``, ``mergeForest``, ``model2dict`` / ``dict2model``,
``drop_duplicated_trees``, ``reset_forest``, ``maxProbClas``, ``Forest`` / ``Forest_size``.

The forest is flattened once per distinct tree list (``flat.forest_tables_*``), uploaded to HBM
(``engine.DeviceForest``) and cached until the list or the scores change.
"""

from __future__ import annotations

from collections import UserList
from random import shuffle
from typing import List, Optional

import numpy as np
import pandas as pd

from . import __version__
from . import engine
from .flat import ForestTables, encode_frame, forest_tables_from_dicts
from .tree import DecisionTreeMC


def _reaches_marked(t: ForestTables, row_codes, absent, marked) -> bool:
    """Does any tree's evaluation of this row touch a node that tests one of the ``marked`` features,
    given that nodes testing an absent column are evaluated on both sides (tree.py:166-175)?"""
    ff, thr, child = t.pnodes["feat_flags"], t.pnodes["thr"], t.pnodes["child"]
    marked = set(int(j) for j in marked)
    for root in t.roots:
        todo = [int(root)]
        while todo:
            i = todo.pop()
            if ff[i] & engine.PNODE_LEAF:
                continue
            f = int(ff[i] & 0xFFFF)
            if absent[f]:
                todo += [int(child[i]), int(child[i]) + 1]
                continue
            if f in marked:
                return True
            c = int(row_codes[f])
            go = (c == thr[i]) if (ff[i] & engine.PNODE_CATEGORICAL) else (c >= thr[i])
            todo.append(int(child[i]) + (0 if go else 1))
    return False


class _ForestKey:
    """What the device copy of a forest was flattened from.  The trees themselves are held (an id() of a
    collected tree can come back on a new one), together with the version of each tree's dict: a forest
    rebuilt by reset_forest() + fit(), dict2model() or an assignment to ``tree.data`` is a different key."""

    def __init__(self, trees, scores, class_vals):
        self.trees = list(trees)
        self.versions = [getattr(t, "_version", 0) for t in self.trees]
        self.materialised = [getattr(t, "_data", True) is not None for t in self.trees]
        self.scores = tuple(float(s) for s in scores)
        self.class_vals = tuple(class_vals)

    def same(self, other: "_ForestKey") -> bool:
        return (
            len(self.trees) == len(other.trees)
            and all(a is b for a, b in zip(self.trees, other.trees))
            and self.versions == other.versions
            and self.scores == other.scores
            and self.class_vals == other.class_vals
        )


class BaseRandomForestMC(UserList):
    typer_error_msg = "Both objects must be instances of 'RandomForestMC' class."

    def __init__(
        self,
        n_trees: int = 16,
        target_col: str = "target",
        min_feature: Optional[int] = None,
        max_feature: Optional[int] = None,
        temporal_features: bool = False,
    ) -> None:
        self.__version__ = __version__
        self.version = __version__
        self.model_version = __version__
        self.target_col = target_col
        self.min_feature = min_feature
        self.max_feature = max_feature
        self.temporal_features = temporal_features
        self.n_trees = n_trees
        self.feat_types = ["numeric", "categorical"]
        self.numeric_cols = None
        self.feature_cols = None
        self.type_of_cols = None
        self.class_vals = None
        self.data = []
        self.survived_scores = []
        self.attr_to_save = [
            "min_feature",
            "max_feature",
            "n_trees",
            "class_vals",
            "survived_scores",
            "version",
            "numeric_cols",
            "feature_cols",
            "type_of_cols",
            "target_col",
            "class_vals",
        ]
        self.soft_voting = False
        self.weighted_tree = False
        self.device = 0
        self._forest_key = None
        self._forest_dev: Optional[engine.DeviceForest] = None
        self._forest_tables: Optional[ForestTables] = None
        self._forest_f32 = None  # ((forest key, float32 features), DeviceForest, ForestTables): see _device_forest

    def __repr__(self) -> str:
        return "{}(len(Forest)={},n_trees={},model_version={},module_version={})".format(
            self.__class__.__name__, len(self.data), self.n_trees, self.model_version, self.version
        )

    def __eq__(self, other):
        if not isinstance(other, BaseRandomForestMC):
            raise TypeError(self.typer_error_msg)
        return all([getattr(self, att) == getattr(other, att) for att in self.attr_to_save])

    __hash__ = None

    # -- switches (forest.py:151-155) ------------------------------------------------------------
    def setSoftVoting(self, set: bool = True) -> None:
        self.soft_voting = set

    def setWeightedTrees(self, set: bool = True) -> None:
        self.weighted_tree = set

    def reset_forest(self) -> None:
        self.data = []
        self.survived_scores = []

    def validFeaturesTemporal(self):
        return all([x.split("_")[-1].isnumeric() for x in self.feature_cols])

    @property
    def Forest_size(self) -> int:
        return len(self)

    @property
    def Forest(self) -> List[DecisionTreeMC]:
        return self.data

    @staticmethod
    def maxProbClas(leaf: dict):
        """forest.py:200-202: stable descending sort -> first key among equal maxima."""
        return sorted(leaf.items(), key=lambda x: x[1], reverse=True)[0][0]

    # -- persistence / merge (forest.py:114-190) ---------------------------------------------------
    def model2dict(self) -> dict:
        out = {attr: getattr(self, attr) for attr in self.attr_to_save}
        out["Forest"] = [Tree.tree2dict() for Tree in self.data]
        return out

    def dict2model(self, dict_model: dict, add: bool = False) -> None:
        new_trees = [
            DecisionTreeMC(t["data"], t["class_vals"], t["survived_score"], t["features"], t["used_features"])
            for t in dict_model["Forest"]
        ]
        if add:
            self.data.extend(new_trees)
            self.survived_scores.extend(dict_model["survived_scores"])
        for attr in self.attr_to_save:
            setattr(self, attr, dict_model[attr])
        self.model_version = self.version
        self.version = self.__version__
        self.data = new_trees  # forest.py:184 -- also on add=True, kept as is

    def mergeForest(self, otherForest, N: int = -1, by: str = "add"):
        if not isinstance(otherForest, BaseRandomForestMC):
            raise TypeError(self.typer_error_msg)
        same = (
            all(c in otherForest.class_vals for c in self.class_vals)
            and all(c in self.class_vals for c in otherForest.class_vals)
            and all(f in otherForest.feature_cols for f in self.feature_cols)
            and all(f in self.feature_cols for f in otherForest.feature_cols)
        )
        if not same:
            raise ValueError("Both forests must have the same set of features and classes.")
        if by == "add":
            self.data.extend(otherForest.data)
        if by == "score":
            self.data = sorted(self.data + otherForest.data)[::-1][:N]
        if by == "random":
            data = self.data + otherForest.data
            shuffle(data)
            self.data = data[:N]
        self.survived_scores = [Tree.survived_score for Tree in self.data]

    def drop_duplicated_trees(self) -> int:
        seen, keep = set(), []
        for Tree in self.data:
            h = Tree.md5hexdigest
            keep.append(h not in seen)
            seen.add(h)
        self.data = [Tree for Tree, k in zip(self.data, keep) if k]
        self.survived_scores = [s for s, k in zip(self.survived_scores, keep) if k]
        return sum(keep)

    # -- the GPU forest ------------------------------------------------------------------------------
    def _device_forest(self, f32_feats=frozenset()) -> engine.DeviceForest:
        """The forest resident on the device.  ``f32_feats`` (normally empty) selects the variant whose
        Python-float split values of those features are rounded to float32, for frames that hand over
        np.float32 cells (``_float32_features``); it is flattened on first use and cached beside the
        default one."""
        if len(self.data) == 0:
            raise ZeroDivisionError("the forest is empty")
        with self._forest_lock():
            key = _ForestKey(self.data, self.survived_scores, self.class_vals)
            if f32_feats:
                hit = self._forest_f32
                if hit is None or hit[0][1] != f32_feats or not key.same(hit[0][0]):
                    if hit is not None:
                        hit[1].close()
                    tables = self._flatten(f32_feats)
                    hit = self._forest_f32 = ((key, f32_feats), engine.DeviceForest(tables, self.device), tables)
                return hit[1]
            if self._forest_dev is None or self._forest_key is None or not key.same(self._forest_key):
                if self._forest_dev is not None:
                    self._forest_dev.close()
                tables = self._flatten()
                self._forest_dev = engine.DeviceForest(tables, self.device)
                self._forest_tables = tables
                self._forest_key = key
            return self._forest_dev

    def _forest_lock(self):
        lock = self.__dict__.get("_forest_mutex")
        if lock is None:
            import threading

            lock = self.__dict__.setdefault("_forest_mutex", threading.RLock())
        return lock

    def _flatten(self, f32_feats=frozenset()) -> ForestTables:
        from .flat import forest_tables_from_soa

        scores = list(self.survived_scores)
        if len(scores) != len(self.data):  # dict2model(add=True) leaves them out of step (forest.py:177-184)
            scores = scores[: len(self.data)] + [0.0] * max(0, len(self.data) - len(scores))
        tables = None
        if all(getattr(t, "soa", None) is not None and t._data is None for t in self.data):
            enc = self.data[0]._enc
            if all(t._enc is enc for t in self.data):
                tables = forest_tables_from_soa([t.soa for t in self.data], scores, enc, self.class_vals, self.feature_cols, f32_feats)
        if tables is None:
            tables = forest_tables_from_dicts([t.data for t in self.data], scores, self.class_vals, self.feature_cols, f32_feats)
        if len(self.survived_scores) != len(self.data):
            # weighted votes pair trees with scores by zip() but divide by fsum(self.survived_scores) -- ALL of
            # them, also the ones without a tree (forest.py:208-212, 224-231)
            from math import fsum

            tables.score_fsum = fsum(float(s) for s in self.survived_scores)
        return tables

    def _float32_features(self, frame: pd.DataFrame, cells=None) -> frozenset:
        """Features whose cells reach the comparison of tree.py:177-180 as np.float32 scalars.  The
        reference walks ``ds.iterrows()``: a frame whose columns are all plain numbers with a float32
        common dtype yields float32 rows (mixed frames yield object rows of Python numbers, wider frames
        float64 rows).  A np.float32 cell meets a Python-float split value in float32 (NEP 50).
        ``cells``: the row itself when a single Series / dict row is voted (its cells keep their types)."""
        if cells is not None:
            return frozenset(j for j, name in enumerate(self.feature_cols) if type(cells.get(name)) is np.float32)
        kinds = [getattr(dt, "kind", "O") for dt in frame.dtypes]
        if len(kinds) == 0 or any(k not in "fiu" for k in kinds) or np.result_type(*frame.dtypes) != np.float32:
            return frozenset()
        return frozenset(j for j, name in enumerate(self.feature_cols) if name in frame.columns)

    def _vote_frame(self, frame: pd.DataFrame, want_probs: bool, want_labels: bool, cells=None):
        if len(self.data) == 0:
            # no tree votes (forest.py:204-233): the unweighted modes divide by len([]) -> ZeroDivisionError; the
            # weighted ones divide an empty sum by fsum(self.survived_scores), which may still hold scores
            from math import fsum

            if not self.weighted_tree:
                raise ZeroDivisionError("division by zero")
            share = 0.0 / fsum(float(s) for s in self.survived_scores)
            n, C = len(frame), len(self.class_vals)
            return (np.full((n, C), share) if want_probs else None), (np.zeros(n, dtype=np.int64) if want_labels else None)
        f32 = self._float32_features(frame, cells)
        dev = self._device_forest(f32)
        self._raise_for_unorderable_cells(dev, frame)
        return dev.predict_frame(frame, self.soft_voting, self.weighted_tree, want_probs, want_labels)

    def _raise_for_unorderable_cells(self, dev, frame: pd.DataFrame) -> None:
        """tree.py:177-180 tests ``val == split_val or (numeric and val > split_val)``: a cell of a numeric
        feature that is neither a number nor NaN (None, pd.NA, a string) makes ``>`` raise TypeError -- but
        only when some tree's walk reaches a node that tests it.  Such cells can only live in object /
        string / nullable columns; the rows that hold one are walked once more through the leaf-export
        kernel with those features marked: a walk that crosses a marked feature is a walk the reference
        would have died on (up to its first marked node the walk is the true one)."""
        t = dev.tables
        marks = {}
        for j, name in enumerate(t.feature_cols):
            if t.thresholds[j] is None or name not in frame.columns:
                continue
            col = frame[name]
            if isinstance(col.dtype, np.dtype) and col.dtype != object:
                continue  # plain numeric / bool columns hold numbers only
            cells = col.astype(object).to_numpy()  # what ds.iterrows() hands to the comparison, cell by cell
            bad = np.fromiter((not isinstance(v, (int, float, np.number, np.bool_)) for v in cells), dtype=bool, count=len(cells))
            if bad.any():
                marks[j] = bad
        if not marks:
            return
        feats = sorted(marks)
        pattern = np.stack([marks[j] for j in feats], axis=1)  # [rows, marked features]
        rows = np.flatnonzero(pattern.any(axis=1))
        codes, absent = encode_frame(t, frame.iloc[rows])
        groups = {}
        for k, r in enumerate(rows):
            groups.setdefault(pattern[r].tobytes(), []).append(k)
        message = "'>' not supported between a non-numeric cell and a numeric split value (tree.py:177-180)"
        for key, members in groups.items():
            marked = [j for j, hit in zip(feats, np.frombuffer(key, dtype=bool)) if hit]
            if absent.any():
                # a column missing from the frame makes the reference evaluate BOTH subtrees (tree.py:166-175),
                # so a marked node anywhere below counts; rare enough to walk on the host, row by row
                if any(_reaches_marked(t, codes[k], absent, marked) for k in members):
                    raise TypeError(message)
                continue
            ab = absent.copy()
            ab[marked] = 1
            _, crossed = dev.leaves(codes[members], ab)
            if np.any(crossed):
                raise TypeError(message)

    # -- inference API (forest.py:100-112, 204-274) ----------------------------------------------
    def useForest(self, row) -> dict:
        if isinstance(row, dict):
            row = pd.Series(row)
        frame = pd.DataFrame([row.to_numpy(dtype=object)], columns=list(row.index))
        cells = row if row.dtype == object else ({k: row.dtype.type() for k in row.index} if row.dtype == np.float32 else {})
        probs, _ = self._vote_frame(frame, True, False, cells=cells)
        return {c: float(p) for c, p in zip(self.class_vals, probs[0])}

    def testForest(self, ds: pd.DataFrame) -> list:
        if len(ds) == 0:
            return []
        _, labels = self._vote_frame(ds, False, True)
        return np.array(self.class_vals, dtype=object)[labels].tolist()

    def testForestProbs(self, ds: pd.DataFrame) -> List[dict]:
        if len(ds) == 0:
            return []
        probs, _ = self._vote_frame(ds, True, False)
        cv = self.class_vals
        return [dict(zip(cv, row)) for row in probs.tolist()]

    def _testForest_func(self, row):
        return self.maxProbClas(self.useForest(row))

    def testForestParallel(self, ds: pd.DataFrame, max_workers: Optional[int] = None, chunksize: Optional[int] = None):
        """Same result as ``testForest``; the row parallelism lives in the kernel."""
        return self.testForest(ds)

    def testForestProbsParallel(self, ds: pd.DataFrame, max_workers: Optional[int] = None, chunksize: Optional[int] = None):
        return self.testForestProbs(ds)

    def predict_proba(self, row_or_matrix, prob_output: bool = True):
        return self.predict(row_or_matrix, prob_output)

    def predict(self, row_or_matrix, prob_output: bool = False):
        if isinstance(row_or_matrix, pd.Series):
            return self.useForest(row_or_matrix)
        if isinstance(row_or_matrix, pd.DataFrame):
            if prob_output:
                return self.testForestProbs(row_or_matrix)
            return self.testForest(row_or_matrix)
        raise TypeError("The input argument must be a Pandas Series or a Pandas DataFrame.")


def _single_tree_output(tree: DecisionTreeMC, row) -> dict:
    """``Tree(row)`` (tree.py:185-212): a one-tree forest through the leaf-export kernel; the
    payload index maps back to the leaf dict, or to the renormalised all-classes dict when the
    walk crossed a column the row does not have (tree.py:201-212)."""
    if isinstance(row, dict):
        row = pd.Series(row)
    feats = tree.features
    if feats is None:
        feats = sorted(_features_of(tree.data))
    # cells that arrive as np.float32 meet Python-float split values in float32 (see _float32_features)
    if row.dtype == np.float32:
        f32 = frozenset(j for j, name in enumerate(feats) if name in row.index)
    elif row.dtype == object:
        f32 = frozenset(j for j, name in enumerate(feats) if name in row.index and type(row[name]) is np.float32)
    else:
        f32 = frozenset()
    tables = forest_tables_from_dicts([tree.data], [1.0], tree.class_vals, feats, f32)
    frame = pd.DataFrame([row.to_numpy(dtype=object)], columns=list(row.index))
    codes, absent = encode_frame(tables, frame)
    marked = [j for j, name in enumerate(feats) if tables.thresholds[j] is not None and name in row.index
              and not isinstance(row[name], (int, float, np.number, np.bool_))]
    if marked and _reaches_marked(tables, codes[0], absent, marked):
        raise TypeError("'>' not supported between a non-numeric cell and a numeric split value (tree.py:177-180)")
    dev = engine.DeviceForest(tables)
    try:
        payload, seen = dev.leaves(codes, absent)
    finally:
        dev.close()
    pid = int(payload[0, 0])
    if seen[0, 0]:
        return {c: float(p) for c, p in zip(tree.class_vals, tables.probs_absent[pid])}
    return tables.payload_leaf[pid]


def _features_of(node, out=None):
    out = set() if out is None else out
    todo = [node]
    while todo:
        n = todo.pop()
        name = next(iter(n))
        if name != "leaf":
            out.add(name)
            todo.append(n[name]["split"][">="])
            todo.append(n[name]["split"]["<"])
    return out
