"""TEST INFRASTRUCTURE: an oracle-backed stand-in for the three engine handles, so the host-side
logic (speculative batching + RNG rewind, keep/discard policy, forest flattening, persistence,
rank sharding) can be exercised on a machine without a GPU.  The product never imports this."""
import numpy as np

from oracle import code_model as cm
from random_forest_mc_b200.flat import TreeSoA


class FakeDataset:
    def __init__(self, enc, device=0):
        self.enc = enc
        self.device = device

    def close(self):
        pass


class FakeBatch:
    grown_total = 0

    def __init__(self, ds, idx_train, idx_val, cand_slot, feat_lists, max_depth, min_samples_split, stream=None, plan=None):
        self.ds = ds
        if plan is not None:  # rows of a (fake) device draw plan
            assert idx_train is None and idx_val is None and not plan.closed
            idx_train, idx_val = plan.T, plan.V
        self.idx_train = np.asarray(idx_train)
        self.idx_val = np.asarray(idx_val)
        self.cand_slot = list(cand_slot)
        self.feat_lists = [list(f) for f in feat_lists]
        self.max_depth, self.mss = max_depth, min_samples_split
        self.n_cand = len(self.cand_slot)
        self.n_val = self.idx_val.shape[1]
        self.n_train = self.idx_train.shape[1]
        self.launches = 0
        self._trees = None

    def run(self, stream=None):
        enc = self.ds.enc
        self._trees = []
        for s, f in zip(self.cand_slot, self.feat_lists):
            self._trees.append(cm.grow_codes(enc, self.idx_train[s], f, self.max_depth, self.mss))
        FakeBatch.grown_total += self.n_cand
        self.launches = 1
        return self

    def results(self):
        enc = self.ds.enc
        correct = np.array(
            [cm.score_codes(enc, n, c, self.idx_val[s]) for (n, c), s in zip(self._trees, self.cand_slot)], dtype=np.int32
        )
        n_nodes = np.array([len(n) for n, _ in self._trees], dtype=np.int32)
        self.correct, self.n_nodes = correct, n_nodes
        return correct, n_nodes

    def fetch(self, cand_ids):
        return [TreeSoA(self._trees[i][0].copy(), self._trees[i][1].copy()) for i in cand_ids]

    def close(self):
        pass


class FakeDrawPlan:
    """engine.DrawPlan's interface on top of the HOST walk of numpy's stream (product code,
    csrc/host_rng.cpp): the device-plan branch of the fit loop -- row ids that never visit the host,
    stream states read back only for a rewind -- runs on a machine without a GPU."""

    created = 0

    def __init__(self, ds, class_rows, key, pos, n_slots, per_class, n_train_pclass, n_cands, low, high, stream=None):
        from random_forest_mc_b200 import engine

        FakeDrawPlan.created += 1
        self.n_slots, self.closed = n_slots, False
        self.key = np.array(key, dtype=np.uint32).copy()
        self.T, self.V, self.K, self.ksnap, self.psnap, self.pos = engine.legacy_draw_plan(
            self.key, int(pos), n_slots, class_rows, per_class, n_train_pclass, n_cands, low, high, 0)

    def finish(self):
        return self.K, self.key, self.pos

    def state_at(self, slot):
        assert not self.closed, "stream state asked of a closed plan"
        return (self.key, self.pos) if slot == self.n_slots else (self.ksnap[slot], int(self.psnap[slot]))

    def indices(self, slot0=0, n_slots=None):
        assert not self.closed, "rows asked of a closed plan"
        n = self.n_slots - slot0 if n_slots is None else n_slots
        return self.T[slot0 : slot0 + n], self.V[slot0 : slot0 + n]

    def close(self):
        self.closed = True


class FakeForest:
    def __init__(self, tables, device=0):
        self.tables = tables
        self.n_classes = tables.n_classes

    def predict(self, codes, absent, soft, weighted, want_probs=True, want_labels=True, stream=None):
        p, l = cm.vote_tables(self.tables, codes, absent, soft, weighted)
        return (p if want_probs else None), (l if want_labels else None)

    def predict_frame(self, frame, soft, weighted, want_probs=True, want_labels=True):
        from random_forest_mc_b200.flat import encode_frame

        codes, absent = encode_frame(self.tables, frame)
        return self.predict(codes, absent, soft, weighted, want_probs, want_labels)

    def leaves(self, codes, absent, stream=None):
        t = self.tables
        LEAF, CAT = 0x80000000, 0x40000000
        ff, thr, child = t.pnodes["feat_flags"], t.pnodes["thr"], t.pnodes["child"]
        n, T = codes.shape[0], len(t.roots)
        payload = np.zeros((n, T), dtype=np.int32)
        seen = np.zeros((n, T), dtype=bool)
        for r in range(n):
            for ti in range(T):
                i = int(t.roots[ti])
                while not (ff[i] & LEAF):
                    f = int(ff[i] & 0xFFFF)
                    c = int(codes[r, f])
                    go = (c == thr[i]) if (ff[i] & CAT) else (c >= thr[i])
                    if absent is not None and absent[f]:
                        go, seen[r, ti] = True, True
                    i = int(child[i]) + (0 if go else 1)
                payload[r, ti] = int(thr[i])
        return payload, seen

    def close(self):
        pass


def fake_encode_dataset_device(dataset_numpy, feature_cols, numeric_cols, target_col, class_vals, device=0):
    from oracle import encode_ref

    enc = encode_ref.encode_dataset(dataset_numpy, feature_cols, numeric_cols, target_col, class_vals)
    return enc, FakeDataset(enc, device)


def install(monkeypatch):
    from random_forest_mc_b200 import engine

    monkeypatch.setattr(engine, "stream_create", lambda device=0: 0)
    monkeypatch.setattr(engine, "stream_destroy", lambda device, stream: None)
    monkeypatch.setattr(engine, "DeviceDataset", FakeDataset)
    monkeypatch.setattr(engine, "encode_dataset_device", fake_encode_dataset_device)
    monkeypatch.setattr(engine, "GrowBatch", FakeBatch)
    monkeypatch.setattr(engine, "DrawPlan", FakeDrawPlan)
    monkeypatch.setattr(engine, "DeviceForest", FakeForest)


def install_plain():
    from random_forest_mc_b200 import engine

    engine.DeviceDataset, engine.GrowBatch, engine.DeviceForest = FakeDataset, FakeBatch, FakeForest
    engine.DrawPlan = FakeDrawPlan
    engine.encode_dataset_device = fake_encode_dataset_device
    engine.stream_create = lambda device=0: 0
    engine.stream_destroy = lambda device, stream: None
