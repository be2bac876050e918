import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from bench import WORKLOADS, synth_params, synth_pool
from code2vec_b200 import _lib, functional as CF
w = dict(WORKLOADS["cfg2"]); dev = torch.device("cuda:0")
p = synth_params(w, dev); s, pth, e, lab = synth_pool(w, 2, dev, 1234)
B = w["B"]
dims = CF.make_dims(w["T"], w["P"], w["C"], w["Et"], w["Ep"], w["H"])
params = CF.make_params(p["terminal_embedding.weight"], p["path_embedding.weight"], p["input_linear.weight"],
                        p["input_layer_norm.weight"], p["input_layer_norm.bias"], p["attention_parameter"])
a = CF.encode_forward(dims, params, s[:B], pth[:B], e[:B], algo=_lib.ALGO_TCGEN05)
a2 = CF.encode_forward(dims, params, s[:B], pth[:B], e[:B], algo=_lib.ALGO_TCGEN05)
b = CF.encode_forward(dims, params, s[:B], pth[:B], e[:B], algo=_lib.ALGO_FFMA)
d1 = max((a[0]-b[0]).abs().max().item(), (a[1]-b[1]).abs().max().item())
d2 = max((a2[0]-b[0]).abs().max().item(), (a2[1]-b[1]).abs().max().item())
nbad = ((a[0]-b[0]).abs().max(1).values > 2e-6).sum().item()
print(f"first {d1:.2e} second {d2:.2e} bags_bad_first {nbad} fence={os.environ.get('C2V_PRODUCER_FENCE','0')}")
