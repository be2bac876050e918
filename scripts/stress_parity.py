"""stress: tcgen05 vs FFMA encode on many full batches (catches rare races / stale smem reads)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from bench import WORKLOADS, synth_params, synth_pool
from code2vec_b200 import _lib, functional as CF
w = dict(WORKLOADS["cfg2"]); dev = torch.device("cuda:0")
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 24
p = synth_params(w, dev); s, pth, e, lab = synth_pool(w, nb, dev, 99)
B, L = w["B"], w["L"]
dims = CF.make_dims(w["T"], w["P"], w["C"], w["Et"], w["Ep"], w["H"])
params = CF.make_params(p["terminal_embedding.weight"], p["path_embedding.weight"], p["input_linear.weight"],
                        p["input_layer_norm.weight"], p["input_layer_norm.bias"], p["attention_parameter"])
worst = 0.0
for i in range(nb):
    o = i * B
    a = CF.encode_forward(dims, params, s[o:o+B], pth[o:o+B], e[o:o+B], algo=_lib.ALGO_TCGEN05)
    b = CF.encode_forward(dims, params, s[o:o+B], pth[o:o+B], e[o:o+B], algo=_lib.ALGO_FFMA)
    for rep in range(3):   # back-to-back relaunches of the tensor-core kernel must be bit-identical
        c = CF.encode_forward(dims, params, s[o:o+B], pth[o:o+B], e[o:o+B], algo=_lib.ALGO_TCGEN05)
        assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]), f"nondeterministic at batch {i}"
    d = max((a[0] - b[0]).abs().max().item(), (a[1] - b[1]).abs().max().item())
    worst = max(worst, d)
print(f"stress: {nb} batches x 1024 bags, max |tcgen05 - ffma| = {worst:.3e}")
assert worst < 5e-6
