"""``DecisionTreeMC``: the reference's tree object (tree.py:44-212) backed by the engine's SoA table.

A tree grown on the GPU arrives as a flat node table.  The nested dict the reference exposes as
``tree.data`` is materialised on first access (``flat.soa_to_dict``) and cached; a tree built from a
dict (``dict2model``, user code) simply holds the dict.  Either way the public surface is the
reference's: mapping behaviour, ordering by ``survived_score``, ``tree2dict``, ``md5hexdigest``,
``depths``.
"""

from __future__ import annotations

import json
from collections import UserDict
from hashlib import md5
from typing import Any, List, Optional

import numpy as np

from . import __version__
from .flat import TreeSoA, soa_to_dict


def json_encoder(obj: Any) -> Any:
    """utils.py:34-38 of the reference: numpy scalars -> Python numbers, dates -> ISO strings."""
    import datetime

    if isinstance(obj, np.generic):
        return obj.item()
    if isinstance(obj, (datetime.date, datetime.datetime)):
        return obj.isoformat()


class DecisionTreeMC(UserDict):
    typer_error_msg = "Both objects must be instances of 'DecisionTreeMC' class."

    def __init__(
        self,
        data: Optional[dict],
        class_vals: List[Any],
        survived_score=None,
        features: Optional[List[str]] = None,
        used_features: Optional[List[str]] = None,
    ) -> None:
        # UserDict.__init__ is bypassed on purpose, as in the reference (tree.py:86)
        self._data = data
        self._soa: Optional[TreeSoA] = None
        self._used_features = None
        self._enc = None
        self._type_of_cols = None
        self.class_vals = class_vals
        self.survived_score = survived_score
        self.features = features
        self.used_features = used_features
        self.module_version = __version__
        self.attr_to_save = ["data", "class_vals", "survived_score", "features", "used_features", "module_version"]

    @classmethod
    def from_soa(cls, soa: TreeSoA, enc, type_of_cols, class_vals, **kw) -> "DecisionTreeMC":
        t = cls(None, class_vals, **kw)
        t._soa, t._enc, t._type_of_cols = soa, enc, type_of_cols
        return t

    # -- the nested dict, built lazily ---------------------------------------------------------
    @property
    def data(self) -> dict:
        if self._data is None and self._soa is not None:
            self._data = soa_to_dict(self._soa, self._enc, self._type_of_cols)
        return self._data

    @data.setter
    def data(self, value) -> None:
        self._data = value
        self._soa = None
        self._version = getattr(self, "_version", 0) + 1  # the device copy of a forest notices the new dict

    @property
    def soa(self) -> Optional[TreeSoA]:
        return self._soa

    @property
    def used_features(self):
        """Features the tree splits on (model.py:413-417).  For a tree that came from the grower the
        list is read off the node table on first access."""
        if self._used_features is None and getattr(self, "_used_mask", None) is not None:
            # a tree that came over the survivor exchange: the mask of its split features travelled with it
            from .flat import reference_used_features

            mask, feature_cols, class_vals = self._used_mask
            split_names = [feature_cols[j] for j, hit in enumerate(mask) if hit]
            self._used_features = reference_used_features(feature_cols, split_names, class_vals, len(split_names) > 0)
        if self._used_features is None and self._soa is not None and self._enc is not None:
            from .flat import reference_used_features_of_soa

            self._used_features = reference_used_features_of_soa(self._enc.feature_cols, self._soa, self._enc.class_sorted)
        return self._used_features

    @used_features.setter
    def used_features(self, value) -> None:
        self._used_features = value

    def _check_format(self, other):
        if not isinstance(other, DecisionTreeMC):
            raise TypeError(self.typer_error_msg)

    def __str__(self) -> str:
        return str(self.data)

    def __repr__(self) -> str:
        return "DecisionTreeMC(survived_score={},module_version={})".format(self.survived_score, self.module_version)

    def __eq__(self, other) -> bool:
        self._check_format(other)
        return self.survived_score == other.survived_score

    __hash__ = None

    def __gt__(self, other) -> bool:
        self._check_format(other)
        return self.survived_score > other.survived_score

    def __ge__(self, other) -> bool:
        self._check_format(other)
        return self.survived_score >= other.survived_score

    def __lt__(self, other) -> bool:
        self._check_format(other)
        return self.survived_score < other.survived_score

    def __le__(self, other) -> bool:
        self._check_format(other)
        return self.survived_score <= other.survived_score

    def tree2dict(self) -> dict:
        return {attr: getattr(self, attr) for attr in self.attr_to_save}

    @property
    def md5hexdigest(self) -> str:
        """tree.py:138-143: md5 over the key-sorted JSON of the tree dict."""
        return md5(json.dumps(self.data, sort_keys=True, default=json_encoder).encode("utf-8")).hexdigest()

    @property
    def depths(self) -> List[int]:
        """Depth of every leaf (tree.py:145-153 parses them out of ``str(tree)``)."""
        if self._soa is not None:
            nodes = self._soa.nodes
            depth = np.zeros(len(nodes), dtype=np.int64)
            depth[0] = 1
            out = []
            # BFS numbering: parents precede children
            for i in range(len(nodes)):
                if nodes["feat"][i] < 0:
                    continue
                c = int(nodes["child"][i])
                depth[c] = depth[c + 1] = depth[i] + 1
            # the reference lists leaves in dict (pre-order, '>=' first) order
            todo = [0]
            while todo:
                i = todo.pop()
                if nodes["feat"][i] < 0:
                    out.append(int(depth[i]))
                else:
                    c = int(nodes["child"][i])
                    todo.append(c + 1)
                    todo.append(c)
            return out
        out = []
        todo = [self.data]
        while todo:
            node = todo.pop()
            name = next(iter(node))
            if name == "leaf":
                out.append(int(str(node["depth"]).split("#")[0]))
            else:
                s = node[name]["split"]
                todo.append(s["<"])
                todo.append(s[">="])
        return out

    # -- single-row traversal --------------------------------------------------------------------
    def __call__(self, row):
        return self.useTree(row)

    def useTree(self, row):
        """tree.py:185-212 for one row: runs the voting kernel on a one-tree forest and maps the
        leaf it reached back to the leaf dict (absent-column variant included)."""
        from .forest import _single_tree_output

        return _single_tree_output(self, row)
