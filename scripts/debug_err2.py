import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from bench import WORKLOADS, synth_params, synth_pool
from code2vec_b200 import _lib, functional as CF
from oracle import oracle
w = dict(WORKLOADS["cfg2"]); dev = torch.device("cuda:0")
p = synth_params(w, dev); s, pth, e, lab = synth_pool(w, 64, dev, 1234)
B, L = w["B"], w["L"]
dims = CF.make_dims(w["T"], w["P"], w["C"], w["Et"], w["Ep"], w["H"])
params = CF.make_params(p["terminal_embedding.weight"], p["path_embedding.weight"], p["input_linear.weight"],
                        p["input_layer_norm.weight"], p["input_layer_norm.bias"], p["attention_parameter"])
cv_t, at_t = CF.encode_forward(dims, params, s[:B], pth[:B], e[:B], algo=_lib.ALGO_AUTO)   # FIRST launch in the process
torch.cuda.synchronize()
n = 64
cp = {k: v.cpu() for k, v in p.items()}
npar = {k: v.numpy() for k, v in cp.items()}
o_cv, o_at = oracle.encode_forward(s[:n].cpu().numpy(), pth[:n].cpu().numpy(), e[:n].cpu().numpy(),
                                   npar["terminal_embedding.weight"], npar["path_embedding.weight"], npar["input_linear.weight"],
                                   npar["input_layer_norm.weight"], npar["input_layer_norm.bias"], npar["attention_parameter"])
with torch.no_grad():
    _, t_cv, t_at = oracle.torch_forward(cp, s[:n].cpu(), pth[:n].cpu(), e[:n].cpu(), lab[:n].cpu())
print("threads", torch.get_num_threads())
print(f"torchCPU-vs-Coracle (64 bags) cv {np.abs(t_cv.numpy()-o_cv).max():.2e} att {np.abs(t_at.numpy()-o_at).max():.2e}")
print(f"GPU first launch-vs-Coracle cv {np.abs(cv_t[:n].cpu().numpy()-o_cv).max():.2e} att {np.abs(at_t[:n].cpu().numpy()-o_at).max():.2e}")
d = np.abs(t_cv.numpy()-o_cv).max(1); print("per-bag torchCPU err:", np.round(d[:64]*1e7).astype(int).tolist())
