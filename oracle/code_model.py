"""CPU ORACLE, code domain -- TEST INFRASTRUCTURE ONLY (see oracle/rfmc_oracle.py for the rules).

Restates the reference's grower and candidate scorer (src/random_forest_mc/model.py:222-314) on the
dense rank codes of ``random-forest-mc_b200/encode.py`` and emits the same flat node table the CUDA
grower emits: breadth-first numbering, root = 0, children of the split nodes of one level numbered
consecutively in level order ('>=' child first, '<' child = '>=' child + 1).  With that numbering
the GPU output can be compared array-for-array, not only through the dict conversion.

Pinned two ways in tests/test_code_model.py: (1) dict conversion of its output equals the tree
dicts recorded from the real reference (tests/golden), types and key order included; (2) its
correct-counts equal the recorded hold-out scores.

Semantics restated:
  numeric, n > 2 : a, b = the two middle order statistics; split_val = numpy's lerp in the column
                   dtype; mask ``v >= split_val``; if a side is empty, mask ``v > split_val``;
                   the stored test is always ``code >= lower_bound(uniq, split_val) + 1``
  categorical    : mode (ties -> lowest code = np.unique order); mask ``v == mode``
  n == 2         : stable argsort; the larger value (second row on ties) goes to '>='
  accept         : both sides >= min_samples_split, else rotate to the next feature of the
                   candidate's list; a full rotation without success makes a leaf
"""

from __future__ import annotations

import importlib
import numpy as np

_enc = importlib.import_module("random_forest_mc_b200.encode")
_flat = importlib.import_module("random_forest_mc_b200.flat")

KIND_CATEGORICAL = _enc.KIND_CATEGORICAL
NODE_DTYPE = _flat.NODE_DTYPE
FLAG_CATEGORICAL = _flat.FLAG_CATEGORICAL
FLAG_TWO_ROW = _flat.FLAG_TWO_ROW


def split_codes(enc, f, codes_node):
    """One split attempt on feature index ``f``.  Returns (mask for '>=', thr stored in the node,
    flags, split value as float64 or nan)."""
    col = enc.columns[f]
    n = len(codes_node)
    cat = col.kind == KIND_CATEGORICAL
    if n <= 2:  # model.py:258-262
        c0, c1 = int(codes_node[0]), int(codes_node[1])
        second_is_hi = c0 <= c1
        mask = np.array([not second_is_hi, second_is_hi])
        flags = FLAG_TWO_ROW | (FLAG_CATEGORICAL if cat else 0)
        return mask, max(c0, c1), flags, 0.0
    if cat:  # model.py:249-254
        hist = np.bincount(codes_node)
        mode = int(np.argmax(hist))
        return codes_node == mode, mode, FLAG_CATEGORICAL, 0.0
    s = np.sort(codes_node)  # model.py:237
    odd = n % 2 == 1
    k0 = (n - 1) // 2
    ca, cb = int(s[k0]), int(s[k0 + 1])
    sv = _enc.median_value(col.table[ca - 1], col.table[cb - 1], odd, col.tag)
    if (col.tag & 0xFF) == _enc.TAG_F32:
        key = np.float32(sv)
        table = col.table.astype(np.float32)
    else:
        key, table = sv, col.table
    thr_ge = int(np.searchsorted(table, key, "left")) + 1
    mask = codes_node >= thr_ge  # model.py:238
    if not mask.any() or mask.all():  # model.py:242-245
        thr_gt = int(np.searchsorted(table, key, "right")) + 1
        mask = codes_node >= thr_gt
    return mask, thr_ge, 0, sv


def grow_codes(enc, idx_T, feats, max_depth, min_samples_split):
    """Grow one candidate.  ``feats`` = ordered feature INDICES.  Returns (nodes, counts)."""
    idx_T = np.asarray(idx_T, dtype=np.int64)
    C = enc.n_classes
    nf = len(feats)
    nodes = [None]
    counts = [None]
    level = [(idx_T, 0, 0)]  # (rows, rotation offset, node id)
    depth = 1
    while level:
        nxt = []
        for rows, off, nid in level:
            hist = np.bincount(enc.labels[rows], minlength=C).astype(np.int32)
            counts[nid] = hist
            done = None
            if not (depth >= max_depth or np.count_nonzero(hist) == 1):
                for k in range(nf):
                    fl = (off + k) % nf
                    f = feats[fl]
                    mask, thr, flags, sv = split_codes(enc, f, enc.codes[rows, f])
                    na = int(mask.sum())
                    if na >= min_samples_split and len(rows) - na >= min_samples_split:
                        done = (f, thr, flags, sv, rows[mask], rows[~mask], (off + k + 1) % nf)
                        break
            if done is None:
                nodes[nid] = (-1, 0, -1, 0, 0.0)
                continue
            f, thr, flags, sv, hi, lo, off2 = done
            child = len(nodes)
            nodes.extend([None, None])
            counts.extend([None, None])
            nodes[nid] = (f, thr, child, flags, sv)
            nxt.append((hi, off2, child))
            nxt.append((lo, off2, child + 1))
        level = nxt
        depth += 1
    return np.array(nodes, dtype=NODE_DTYPE), np.stack(counts).astype(np.int32)


def leaf_votes(counts):
    """First maximum in sorted-label order == maxProbClas on a genLeaf dict (forest.py:200-202)."""
    return np.argmax(counts, axis=1)


def score_codes(enc, nodes, counts, idx_V):
    """model.py:296-314 on codes: number of hold-out rows whose leaf vote equals their label."""
    votes = leaf_votes(counts)
    correct = 0
    for r in np.asarray(idx_V, dtype=np.int64):
        i = 0
        while nodes["feat"][i] >= 0:
            c = int(enc.codes[r, nodes["feat"][i]])
            t = int(nodes["thr"][i])
            go = (c == t) if (nodes["flags"][i] & FLAG_CATEGORICAL) else (c >= t)
            i = int(nodes["child"][i]) + (0 if go else 1)
        correct += int(votes[i] == enc.labels[r])
    return correct


def vote_tables(t, codes, absent, soft, weighted):
    """forest.py:204-236 on a flattened forest (``flat.ForestTables``) and an encoded frame: the
    CPU statement of what the CUDA voting kernel computes.  Returns (probs [n, C], labels [n])."""
    LEAF, CAT = _flat.PNODE_LEAF, _flat.PNODE_CATEGORICAL
    n, C = codes.shape[0], t.n_classes
    ff, thr, child = t.pnodes["feat_flags"], t.pnodes["thr"], t.pnodes["child"]
    probs = np.zeros((n, C), dtype=np.float64)
    labels = np.zeros(n, dtype=np.int32)
    T = len(t.roots)
    for r in range(n):
        acc = [0.0] * C
        cnt = [0] * C
        for ti in range(T):
            i = int(t.roots[ti])
            seen = False
            while not (ff[i] & LEAF):
                f = int(ff[i] & 0xFFFF)
                c = int(codes[r, f])
                go = (c == thr[i]) if (ff[i] & CAT) else (c >= thr[i])
                if absent is not None and absent[f]:
                    go, seen = True, True
                i = int(child[i]) + (0 if go else 1)
            pid = int(thr[i])
            if not soft:
                v = int(t.vote_absent[pid] if seen else t.vote[pid])
                if weighted:
                    acc[v] = acc[v] + float(t.scores[ti])
                else:
                    cnt[v] += 1
            else:
                p = t.probs_absent[pid] if seen else t.probs[pid]
                for k in range(C):
                    acc[k] = acc[k] + (float(p[k]) * float(t.scores[ti]) if weighted else float(p[k]))
        for k in range(C):
            num = float(cnt[k]) if (not soft and not weighted) else acc[k]
            den = float(T) if not weighted else float(t.score_fsum)
            probs[r, k] = num / den
        labels[r] = int(np.argmax(probs[r]))
    return probs, labels
