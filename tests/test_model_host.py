"""Host-side mirror of the reference interface (no GPU): parameter names / shapes / init order
(SURVEY.md 8a row 1, 8b), star-import surface (main.py:22,130,261), and loud failure on CPU."""
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
import code2vec_b200.model as M
from code2vec_b200 import _lib


def make_option(T, P, C, Et, Ep, H, dropout=0.0, angular=False):
    o = types.SimpleNamespace()
    o.terminal_count, o.path_count, o.label_count = T, P, C
    o.terminal_embed_size, o.path_embed_size, o.encode_size = Et, Ep, H
    o.dropout_prob = dropout
    o.angular_margin_loss, o.angular_margin, o.inverse_temp = angular, 0.5, 30.0
    o.device = torch.device("cpu")
    return o


def test_initial_weights_identical_to_reference_under_same_seed():
    z = np.load(f"{GOLDEN}/init_seed123.npz")
    torch.manual_seed(123)
    m = M.Code2Vec(make_option(50, 40, 9, 6, 10, 8, dropout=0.25))
    sd = m.state_dict()
    ref = {k[6:]: z[k] for k in z.files if k.startswith("param.")}
    assert sorted(sd) == sorted(ref)
    for k in ref:
        assert tuple(sd[k].shape) == ref[k].shape, k
        assert np.array_equal(sd[k].numpy(), ref[k]), k
    torch.manual_seed(123)
    m = M.Code2Vec(make_option(50, 40, 9, 6, 10, 8, angular=True))
    sd = m.state_dict()
    ref = {k[8:]: z[k] for k in z.files if k.startswith("angular.")}
    assert sorted(sd) == sorted(ref)
    for k in ref:
        assert np.array_equal(sd[k].numpy(), ref[k]), k


def test_state_dict_round_trip_with_reference_checkpoint_keys():
    rec = load_golden("tiny")
    o = rec["opt"]
    m = M.Code2Vec(make_option(o["T"], o["P"], o["C"], o["Et"], o["Ep"], o["H"]))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in rec["params"].items()}, strict=True)
    assert [n for n, _ in m.named_parameters()] == [
        "attention_parameter", "terminal_embedding.weight", "path_embedding.weight", "input_linear.weight",
        "input_layer_norm.weight", "input_layer_norm.bias", "output_linear.weight", "output_linear.bias"]


def test_star_import_surface():
    ns = {}
    exec("from code2vec_b200.model import *", ns)
    for name in ("Code2Vec", "nn", "F", "torch", "NINF"):
        assert name in ns
    assert ns["NINF"] == pytest.approx(-3.4e38)


def test_dropout_module_presence_follows_reference_rule():
    assert M.Code2Vec(make_option(5, 5, 3, 4, 4, 4, dropout=0.25)).input_dropout is not None
    assert M.Code2Vec(make_option(5, 5, 3, 4, 4, 4, dropout=0.0)).input_dropout is None
    assert M.Code2Vec(make_option(5, 5, 3, 4, 4, 4, dropout=1.0)).input_dropout is None


def test_cpu_tensors_fail_loudly_no_fallback():
    m = M.Code2Vec(make_option(7, 5, 3, 4, 4, 8))
    s = torch.zeros((2, 3), dtype=torch.int64)
    with pytest.raises(_lib.C2VError, match="no CPU fallback"):
        m.forward(s, s, s, torch.zeros(2, dtype=torch.int64))


def test_product_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(M.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "c2v_oracle" not in src, f
