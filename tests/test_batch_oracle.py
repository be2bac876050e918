"""oracle/batch_oracle.py (the checker of the on-GPU batch builder) pinned to what the reference's own reader +
`DatasetBuilder.build_data` produced from dataset/corpus.txt (tests/golden/builder_corpus.npz, oracle/gen_golden.py)."""
import os
import sys
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import batch_oracle as bo  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "builder_corpus.npz"))
OFF, CTX = G["offsets"], G["contexts"]
MT, QT, L = int(G["method_token"]), int(G["question_token"]), int(G["max_path_length"])
N = len(OFF) - 1


def _rewritten(item):
    c = CTX[OFF[item]:OFF[item + 1]].astype(np.int64).copy()
    c[c[:, 0] == MT, 0] = QT
    c[c[:, 2] == MT, 2] = QT
    return c


def _bag(s, p, e):
    n = int((p != 0).sum()) if (p != 0).any() else 0
    n = max(n, int((s != 0).sum()), int((e != 0).sum()))
    return Counter(map(tuple, np.stack([s[:n], p[:n], e[:n]], 1).tolist())), n


def test_fixture_covers_long_and_short_methods():
    n = np.diff(OFF)
    assert (n > L).sum() >= 6 and (n <= L).sum() >= 40 and n.max() > 3000


def test_oracle_against_the_reference_builder():
    """dataset_builder.py:112-150: same bag sizes, zero-padded suffix, @method_0 -> @question, every bag a sub-multiset
    of the item's contexts, and -- whenever the method has <= max_path_length contexts, where the reference's shuffle
    cannot change the content -- exactly the reference's multiset."""
    s, p, e = bo.build_batch(OFF, CTX, np.arange(N), L, seed=1234, method_token=MT, question_token=QT)
    for i in range(N):
        n = int(OFF[i + 1] - OFF[i])
        mine, n_mine = _bag(s[i], p[i], e[i])
        ref, n_ref = _bag(G["ref_starts"][i], G["ref_paths"][i], G["ref_ends"][i])
        assert n_mine == n_ref == min(n, L)
        assert (s[i, n_mine:] == 0).all() and (p[i, n_mine:] == 0).all() and (e[i, n_mine:] == 0).all()
        assert MT not in s[i] and MT not in e[i]
        pool = Counter(map(tuple, _rewritten(i).tolist()))
        assert not (mine - pool) and not (ref - pool)            # sub-multisets of the item's (rewritten) contexts
        if n <= L:
            assert mine == ref
    assert np.array_equal(G["item_label"], G["ref_label"])


def test_selection_is_a_uniform_subset():
    """shuffle-then-truncate = uniform subset: inclusion frequency of every context of a 50-context method at L = 10
    over 4000 seeds is 0.2 +- 5 sigma, and no index is favoured by position."""
    n, l, trials = 50, 10, 4000
    cnt = np.zeros(n)
    for seed in range(trials):
        sel = bo.select(seed, 7, n, l)
        assert len(sel) == l and len(set(sel.tolist())) == l and (np.diff(sel) > 0).all()
        cnt[sel] += 1
    p = l / n
    sigma = np.sqrt(trials * p * (1 - p))
    assert np.abs(cnt - trials * p).max() < 5 * sigma
    assert abs(np.corrcoef(np.arange(n), cnt)[0, 1]) < 0.5


def test_different_seeds_and_items_give_different_bags():
    i = int(np.argmax(np.diff(OFF)))
    a = bo.select(1, i, int(OFF[i + 1] - OFF[i]), L)
    b = bo.select(2, i, int(OFF[i + 1] - OFF[i]), L)
    c = bo.select(1, i + 1, int(OFF[i + 1] - OFF[i]), L)
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.array_equal(a, bo.select(1, i, int(OFF[i + 1] - OFF[i]), L))


# ---- variable-name task (dataset_builder.py:152-204) pinned to tests/golden/builder_vars.npz ------------------------
GV = np.load(os.path.join(ROOT, "tests", "golden", "builder_vars.npz"))


def _triples(s, p, e):
    n = int(max((s != 0).sum(), (p != 0).sum(), (e != 0).sum()))
    return Counter(map(tuple, np.stack([s[:n], p[:n], e[:n]], 1).tolist())), n


import pytest  # noqa: E402


@pytest.mark.parametrize("tag", ["synth", "real"])
def test_variable_task_oracle_against_the_reference_builder(tag):
    off, ctx, units = GV[f"{tag}_offsets"], GV[f"{tag}_contexts"], GV[f"{tag}_units"]
    Lv, q = int(GV[f"{tag}_L"]), int(GV[f"{tag}_question"])
    rs, rp, re_, rl = (GV[f"{tag}_ref_{k}"] for k in ("starts", "paths", "ends", "label"))
    assert len(units) == len(rs) and (units[:, 2] == rl).all()            # one bag per (item, @var alias), same order
    assert (GV[f"{tag}_item_ids"][units[:, 0]] == GV[f"{tag}_ref_ids"]).all()
    s, p, e = bo.build_batch_vars(off, ctx, units[:, 0], units[:, 1], np.arange(len(units)), Lv, 99, q)
    n_long = 0
    for u, (item, v, _) in enumerate(units):
        c = ctx[off[item]:off[item + 1]].astype(np.int64)
        c = c[(c[:, 0] == v) | (c[:, 2] == v)].copy()
        c[c[:, 0] == v, 0] = q; c[c[:, 2] == v, 2] = q
        pool = Counter(map(tuple, c.tolist()))
        mine, n_mine = _triples(s[u], p[u], e[u])
        ref, n_ref = _triples(rs[u], rp[u], re_[u])
        assert n_mine == n_ref == min(len(c), Lv)
        assert not (mine - pool) and not (ref - pool)
        assert (s[u, n_mine:] == 0).all() and (p[u, n_mine:] == 0).all() and (e[u, n_mine:] == 0).all()
        assert v not in s[u] and v not in e[u]
        if len(c) <= Lv:
            assert mine == ref                                                # the shuffle cannot change the content
        else:
            n_long += 1
    assert n_long >= 1


def test_variable_permutation_is_a_permutation_and_only_touches_variables():
    off, ctx, units = GV["synth_offsets"], GV["synth_contexts"], GV["synth_units"]
    var = GV["synth_variable_indexes"]
    Lv, q = int(GV["synth_L"]), int(GV["synth_question"])
    ids = np.arange(len(units))
    s0, p0, e0 = bo.build_batch_vars(off, ctx, units[:, 0], units[:, 1], ids, Lv, 7, q)
    s1, p1, e1 = bo.build_batch_vars(off, ctx, units[:, 0], units[:, 1], ids, Lv, 7, q, var, True)
    assert (p0 == p1).all()
    changed = 0
    for u, (item, v, _) in enumerate(units):
        sigma = bo.var_permutation(7, int(item), len(var))
        assert sorted(sigma.tolist()) == list(range(len(var)))
        mp = {int(var[i]): int(var[sigma[i]]) for i in range(len(var))}
        for a0, a1 in ((s0[u], s1[u]), (e0[u], e1[u])):
            want = np.asarray([t if t == q or t == 0 else mp.get(int(t), int(t)) for t in a0])
            assert (want == a1).all()
            changed += int((a0 != a1).sum())
    assert changed > 0
