"""helpers shared by the -m gpu tests"""
import types

import numpy as np
import torch

from code2vec_b200 import _lib
from code2vec_b200 import functional as CF
from code2vec_b200.model import Code2Vec

ALGOS = {"ffma": _lib.ALGO_FFMA, "tcgen05": _lib.ALGO_TCGEN05}


def option_from(o, dropout=0.0):
    opt = types.SimpleNamespace()
    opt.terminal_count, opt.path_count, opt.label_count = o["T"], o["P"], o["C"]
    opt.terminal_embed_size, opt.path_embed_size, opt.encode_size = o["Et"], o["Ep"], o["H"]
    opt.dropout_prob = dropout
    opt.angular_margin_loss = bool(o.get("angular", 0))
    opt.angular_margin, opt.inverse_temp = o.get("margin", 0.5), o.get("inverse_temp", 30.0)
    opt.device = torch.device("cuda:0")
    return opt


def model_from_golden(rec, algo="auto", dropout=0.0):
    m = Code2Vec(option_from(rec["opt"], dropout), algo=algo)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in rec["params"].items()}, strict=True)
    return m.to("cuda:0")


def supports_tcgen05(o):
    import ctypes
    d = CF.make_dims(o["T"], o["P"], o["C"], o["Et"], o["Ep"], o["H"])
    return bool(_lib.load().c2v_encode_supports_tcgen05(ctypes.byref(d)))


def algos_for(o):
    return ["ffma", "tcgen05"] if supports_tcgen05(o) else ["ffma"]


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def random_params(rng, T, P, C, Et, Ep, H, scale=1.0):
    D = 2 * Et + Ep
    return {
        "terminal_embedding.weight": (rng.standard_normal((T, Et)) * scale).astype(np.float32),
        "path_embedding.weight": (rng.standard_normal((P, Ep)) * scale).astype(np.float32),
        "input_linear.weight": (rng.uniform(-1, 1, (H, D)) / np.sqrt(D)).astype(np.float32),
        "input_layer_norm.weight": (1.0 + 0.2 * rng.standard_normal(H)).astype(np.float32),
        "input_layer_norm.bias": (0.1 * rng.standard_normal(H)).astype(np.float32),
        "attention_parameter": (rng.standard_normal(H) * np.sqrt(2.0 / (H + 1))).astype(np.float32),
        "output_linear.weight": (rng.uniform(-1, 1, (C, H)) / np.sqrt(H)).astype(np.float32),
        "output_linear.bias": (0.05 * rng.standard_normal(C)).astype(np.float32),
    }


def random_batch(rng, B, L, T, P, C, ragged=True):
    starts = rng.integers(1, T, (B, L)); paths = rng.integers(1, P, (B, L)); ends = rng.integers(1, T, (B, L))
    if ragged:
        n = rng.integers(1, L + 1, B)
        for b in range(B):
            starts[b, n[b]:] = 0; paths[b, n[b]:] = 0; ends[b, n[b]:] = 0
    return starts.astype(np.int64), paths.astype(np.int64), ends.astype(np.int64), rng.integers(0, C, B).astype(np.int64)
