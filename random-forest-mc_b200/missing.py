"""``predictMissingValues`` and its helpers (model.py:495-568 of the reference): enumerate candidate
fill-ins for the NaN cells of a row / frame, vote every filled row in ONE ``predict_proba`` call
(the GPU path), and join the probabilities back to the rows they came from.  Pandas reshaping
around the hot path; the behaviour (including which columns take part in the join) is the
reference's."""

from __future__ import annotations

import logging as log

import pandas as pd


def fill_row(row: pd.Series, dict_values: dict):
    """One copy of ``row`` per (missing column, candidate value).  model.py:496-507."""
    filled = []
    for col, candidates in dict_values.items():
        if pd.isna(row[col]):
            for value in candidates:
                clone = row.copy()
                clone[col] = value
                filled.append(clone)
    if not filled:
        log.warning("Filling rows process: found row without missing data!")
        return None
    return pd.concat(filled, axis=1).transpose().reset_index(drop=True)


def check_dict_values(model, dict_values: dict) -> None:
    """model.py:509-518."""
    from .model import DictValuesAllFeaturesMissing

    used = set()
    for tree in model:
        used |= set(tree.used_features)
    unknown = set(dict_values.keys()) - used
    if unknown:
        log.warning(f"The Forest model have not the following feature(s): [{', '.join(unknown)}].")
    if len(set(dict_values.keys())) == len(unknown):
        raise DictValuesAllFeaturesMissing


def expand(row_or_matrix, dict_values: dict):
    """model.py:520-542: (input as a frame, all filled rows)."""
    from .model import MissingValuesNotFound

    filled = None
    if isinstance(row_or_matrix, pd.Series):
        filled = fill_row(row_or_matrix, dict_values)
        if filled is None:
            raise MissingValuesNotFound
        row_or_matrix = pd.DataFrame(row_or_matrix).transpose().reset_index(drop=True)
    elif isinstance(row_or_matrix, pd.DataFrame):
        row_or_matrix = row_or_matrix.reset_index(drop=True)
        parts = [p for p in (fill_row(r, dict_values) for _, r in row_or_matrix.iterrows()) if p is not None]
        if not parts:
            raise MissingValuesNotFound
        filled = pd.concat(parts).reset_index(drop=True)
    return row_or_matrix, filled


def predict_missing_values(model, row_or_matrix, dict_values: dict) -> pd.DataFrame:
    """model.py:544-568."""
    check_dict_values(model, dict_values)
    frame, filled = expand(row_or_matrix, dict_values)
    scored = pd.concat([filled, pd.DataFrame.from_dict(model.predict_proba(filled))], axis=1)
    keys = list(dict_values.keys())
    out = []
    for i, row in frame.reset_index(drop=True).iterrows():
        match = filled[keys[0]] == row[keys[0]]
        for col in keys[1:]:
            if not pd.isna(row[col]):
                match = match & (filled[col] == row[col])
        block = pd.concat([pd.DataFrame(row).transpose(), scored.loc[match]]).drop_duplicates().reset_index(drop=True)
        block["row_id"] = i
        out.append(block)
    return pd.concat(out).reset_index(drop=True)
