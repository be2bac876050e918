"""c2v_build_batch (on-GPU DatasetBuilder.build_data, dataset_builder.py:112-150) against oracle/batch_oracle.py,
bit for bit, through the C ABI; and the model's indifference to the order inside a bag."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu

G = np.load(os.path.join(ROOT, "tests", "golden", "builder_corpus.npz"))


def _corpus():
    from code2vec_b200.batch_builder import DeviceCorpus
    return DeviceCorpus(G["offsets"], G["contexts"], G["item_label"], int(G["method_token"]), int(G["question_token"]), "cuda:0")


@pytest.mark.parametrize("L", [200, 31, 1, 4000])
@pytest.mark.parametrize("seed", [0, 1234567890123])
def test_kernel_equals_oracle_on_the_reference_corpus(L, seed):
    from oracle import batch_oracle as bo
    c = _corpus()
    n = c.n_items
    ids = np.concatenate([np.arange(n), np.array([0, 0, n - 1, 3])])          # every method, some twice
    s, p, e, lab = c.build(torch.from_numpy(ids), L, seed, check=True)
    rs, rp, re = bo.build_batch(G["offsets"], G["contexts"], ids, L, seed, c.method_token, c.question_token)
    assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(p.cpu().numpy(), rp) and np.array_equal(e.cpu().numpy(), re)
    assert np.array_equal(lab.cpu().numpy(), G["item_label"][ids])
    assert torch.equal(s[0], s[n]) and torch.equal(p[0], p[n])                  # same (seed, item) -> same bag


def test_very_long_method_and_bad_ids():
    """top11's longest method has 60,810 contexts: radix select over a synthetic one, ties included."""
    from code2vec_b200.batch_builder import DeviceCorpus
    from oracle import batch_oracle as bo
    rng = np.random.default_rng(0)
    ns = [60810, 201, 200, 0, 5]
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    ctx = rng.integers(1, 1000, (off[-1], 3)).astype(np.int32)
    ctx[::7, 0] = 2; ctx[::11, 2] = 2                                          # @method_0 = 2 -> @question = 1
    c = DeviceCorpus(off, ctx, np.arange(len(ns)), 2, 1, "cuda:0")
    ids = np.array([0, 1, 2, 3, 4, 0])
    for seed in (3, 4):
        s, p, e, lab = c.build(torch.from_numpy(ids), 200, seed)
        rs, rp, re = bo.build_batch(off, ctx, ids, 200, seed, 2, 1)
        assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(p.cpu().numpy(), rp) and np.array_equal(e.cpu().numpy(), re)
        assert (s[3] == 0).all() and (s.cpu().numpy() != 2).all()
    s, p, e, lab = c.build(torch.tensor([7, -1, 1]), 200, 1)                     # not methods of this corpus: all-pad rows
    assert (s[:2] == 0).all() and (p[:2] == 0).all() and (s[2] != 0).any()
    with pytest.raises(IndexError):
        c.build(torch.tensor([7]), 200, 1, check=True)


def test_model_output_does_not_depend_on_the_order_inside_a_bag():
    """The reference shuffles the contexts of a method; the on-GPU builder keeps the stored order.  For methods with
    <= max_path_length contexts both bags hold the same multiset, and forward() must give the same code vector."""
    from gpu_util import cuda, random_params
    from code2vec_b200 import functional as CF, _lib
    c = _corpus()
    L = int(G["max_path_length"])
    n = np.diff(G["offsets"])
    ids = np.nonzero((n <= L) & (n > 0))[0][:24]
    s, p, e, lab = c.build(torch.from_numpy(ids), L, 5)
    T = int(max(G["contexts"][:, 0].max(), G["contexts"][:, 2].max())) + 1; P = int(G["contexts"][:, 1].max()) + 1
    rng = np.random.default_rng(1)
    prm = random_params(rng, T, P, 8, 100, 100, 100)
    tp = {k: cuda(v) for k, v in prm.items()}
    dims = CF.make_dims(T, P, 8, 100, 100, 100)
    params = CF.make_params(tp["terminal_embedding.weight"], tp["path_embedding.weight"], tp["input_linear.weight"],
                            tp["input_layer_norm.weight"], tp["input_layer_norm.bias"], tp["attention_parameter"])
    cv, _ = CF.encode_forward(dims, params, s, p, e, algo=_lib.ALGO_AUTO, check_indices=True)
    rs, rp, re = (cuda(G[k][ids]) for k in ("ref_starts", "ref_paths", "ref_ends"))     # the reference's shuffled bags
    cv_ref, _ = CF.encode_forward(dims, params, rs, rp, re, algo=_lib.ALGO_AUTO, check_indices=True)
    assert (cv - cv_ref).abs().max().item() <= 2e-6


def test_epoch_covers_every_method_once_and_shards_by_rank():
    c = _corpus()
    seen = []
    for s, p, e, lab in c.epoch(batch_size=10, max_path_length=200, seed=3):
        assert s.shape[1] == 200 and s.shape[0] <= 10
        seen.append(lab)
    assert sum(x.numel() for x in seen) == c.n_items
    a = [lab for *_, lab in c.epoch(16, 200, 3, rank=0, world=2)]
    b = [lab for *_, lab in c.epoch(16, 200, 3, rank=1, world=2)]
    assert sum(x.numel() for x in a) + sum(x.numel() for x in b) == c.n_items


# ---- variable-name task (dataset_builder.py:152-204): c2v_build_batch_vars vs the oracle, bit for bit ---------------
GV = np.load(os.path.join(ROOT, "tests", "golden", "builder_vars.npz"))


def _var_corpus(tag, shuffle):
    from code2vec_b200.batch_builder import DeviceCorpus
    units = GV[f"{tag}_units"]
    T = int(max(GV[f"{tag}_contexts"][:, [0, 2]].max(), GV[f"{tag}_variable_indexes"].max())) + 1
    c = DeviceCorpus(GV[f"{tag}_offsets"], GV[f"{tag}_contexts"], None, -1, int(GV[f"{tag}_question"]), "cuda:0")
    c.set_variable_units(units[:, 0], units[:, 1], units[:, 2], GV[f"{tag}_variable_indexes"], T, shuffle)
    return c, units


@pytest.mark.parametrize("tag", ["synth", "real"])
@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("L", [200, 5, 1])
def test_variable_task_kernel_equals_oracle(tag, shuffle, L):
    from oracle import batch_oracle as bo
    c, units = _var_corpus(tag, shuffle)
    ids = np.concatenate([np.arange(len(units)), np.array([0, len(units) - 1])])
    for seed in (3, 98765432123456789):
        s, p, e, lab = c.build_vars(torch.from_numpy(ids), L, seed)
        rs, rp, re = bo.build_batch_vars(GV[f"{tag}_offsets"], GV[f"{tag}_contexts"], units[:, 0], units[:, 1], ids, L, seed,
                                         int(GV[f"{tag}_question"]), GV[f"{tag}_variable_indexes"], shuffle)
        assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(p.cpu().numpy(), rp) and np.array_equal(e.cpu().numpy(), re)
        assert np.array_equal(lab.cpu().numpy(), units[ids, 2])


def test_variable_task_matches_the_reference_builder_where_the_shuffle_cannot_matter():
    """bags with <= max_path_length matching contexts: the same multiset as DatasetBuilder.build_data produced"""
    from collections import Counter
    c, units = _var_corpus("real", False)
    L = int(GV["real_L"])
    s, p, e, lab = c.build_vars(torch.arange(len(units)), L, 1)
    s, p, e = s.cpu().numpy(), p.cpu().numpy(), e.cpu().numpy()
    assert np.array_equal(lab.cpu().numpy(), GV["real_ref_label"])
    same = 0
    for u in range(len(units)):
        n = int((GV["real_ref_paths"][u] != 0).sum())
        if n < L:
            ref = Counter(map(tuple, np.stack([GV["real_ref_starts"][u, :n], GV["real_ref_paths"][u, :n], GV["real_ref_ends"][u, :n]], 1).tolist()))
            got = Counter(map(tuple, np.stack([s[u, :n], p[u, :n], e[u, :n]], 1).tolist()))
            assert ref == got and (p[u, n:] == 0).all()
            same += 1
    assert same > 100
