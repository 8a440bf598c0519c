"""I/O helpers of the reference package (utils.py:11-98): JSON encoding of numpy scalars / dates,
file load/dump, a directory-of-JSON loader and two list flatteners.  No role on the hot path; kept so
that ``from random_forest_mc.utils import ...`` has a drop-in counterpart."""

from __future__ import annotations

import json
from keyword import iskeyword
from pathlib import Path
from typing import Any, Iterator, Tuple

from .tree import json_encoder  # noqa: F401  (same function the tree's md5 uses)

DEFAULT_DICT_PATH = "./data"
JSON_EXTENSION = ".json"


def flat(a):
    out = []
    for part in a:
        out += part
    return out


def flatten_nested_list(lst):
    out, todo = [], [iter(lst)]
    while todo:
        try:
            item = next(todo[-1])
        except StopIteration:
            todo.pop()
            continue
        if isinstance(item, list):
            todo.append(iter(item))
        else:
            out.append(item)
    return out


def load_file_json(path):
    with open(Path(path), "r") as f:
        return json.load(f)


def dump_file_json(path, var: Any):
    with open(Path(path), "w") as f:
        return json.dump(var, f, indent=4, default=json_encoder)


class LoadDicts:
    def __init__(self, dict_path=DEFAULT_DICT_PATH, ignore_errors: bool = False):
        self.List, self.Dict, self.not_attr = [], {}, []
        for path_json in Path(dict_path).glob("*" + JSON_EXTENSION):
            try:
                name = path_json.as_posix().split("/")[-1].replace(JSON_EXTENSION, "")
                self.List.append(name)
                self.Dict[name] = load_file_json(path_json)
                if name.isidentifier() and not iskeyword(name):
                    setattr(self, name, self.Dict[name])
                else:
                    self.not_attr.append(name)
            except Exception as e:
                print(f"Error trying to load the file: {path_json.absolute()}: ")
                if not ignore_errors:
                    raise e
                print(e)

    def __repr__(self) -> str:
        return "LoadDicts: {}".format(", ".join(self.List))

    def __len__(self) -> int:
        return len(self.List)

    def __iter__(self) -> Iterator[Any]:
        for item in self.List:
            yield self.Dict[item]

    def __getitem__(self, key: str) -> Any:
        return self.Dict[key]

    def items(self) -> Iterator[Tuple[str, Any]]:
        for item in self.List:
            yield item, self.Dict[item]

    def add(self, other: "LoadDicts") -> None:
        for name in other.List:
            self.Dict[name] = other.Dict[name]
        self.List.extend(other.List)
