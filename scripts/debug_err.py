"""debug: where do tcgen05-vs-FFMA differences sit? (dense numerics vs sparse staleness)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from bench import WORKLOADS, synth_params, synth_pool
from code2vec_b200 import _lib, functional as CF
w = dict(WORKLOADS["cfg2"]); dev = torch.device("cuda:0")
p = synth_params(w, dev); s, pth, e, lab = synth_pool(w, 2, dev, 1234)
B, L = w["B"], w["L"]
dims = CF.make_dims(w["T"], w["P"], w["C"], w["Et"], w["Ep"], w["H"])
params = CF.make_params(p["terminal_embedding.weight"], p["path_embedding.weight"], p["input_linear.weight"],
                        p["input_layer_norm.weight"], p["input_layer_norm.bias"], p["attention_parameter"])
cv_f, at_f = CF.encode_forward(dims, params, s[:B], pth[:B], e[:B], algo=_lib.ALGO_FFMA)
outs = []
for rep in range(3):
    cv_t, at_t = CF.encode_forward(dims, params, s[:B], pth[:B], e[:B], algo=_lib.ALGO_TCGEN05)
    outs.append((cv_t.clone(), at_t.clone()))
    dcv = (cv_t - cv_f).abs(); dat = (at_t - at_f).abs()
    rel = dat / at_f.clamp_min(1e-12)
    print(f"rep{rep}: cv max {dcv.max().item():.3e} mean {dcv.mean().item():.3e} | att max {dat.max().item():.3e} mean {dat.mean().item():.3e} "
          f"| att rel max {rel.max().item():.3e} rel mean {rel.mean().item():.3e} | bags with cv err>2e-6: {(dcv.max(1).values>2e-6).sum().item()} "
          f"| rows with rel att err>1e-4: {(rel>1e-4).sum().item()} of {rel.numel()}")
print("deterministic:", all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:]))
rel = ((outs[0][1] - at_f).abs() / at_f.clamp_min(1e-12)).flatten()
top = torch.topk(rel, 10)
print("top rel att err rows:", [(int(i) // L, int(i) % L, float(v)) for v, i in zip(top.values, top.indices)])
h = torch.histc(torch.log10(rel.clamp_min(1e-9)), bins=8, min=-9, max=-1)
print("log10 rel err hist (-9..-1):", h.tolist())
# --- which CPU reference is off?  C oracle (double accumulation) vs torch CPU vs GPU on 16 bags
from oracle import oracle
import subprocess
print(subprocess.run("lscpu | grep -E 'Model name|Flags' | cut -c1-400", shell=True, capture_output=True, text=True).stdout[:600])
n = 16
cp = {k: v.cpu() for k, v in p.items()}
npar = {k: v.numpy() for k, v in cp.items()}
o_cv, o_at = oracle.encode_forward(s[:n].cpu().numpy(), pth[:n].cpu().numpy(), e[:n].cpu().numpy(),
                                   npar["terminal_embedding.weight"], npar["path_embedding.weight"], npar["input_linear.weight"],
                                   npar["input_layer_norm.weight"], npar["input_layer_norm.bias"], npar["attention_parameter"])
for nt in (1, 8, os.cpu_count()):
    torch.set_num_threads(nt)
    with torch.no_grad():
        _, t_cv, t_at = oracle.torch_forward(cp, s[:n].cpu(), pth[:n].cpu(), e[:n].cpu(), lab[:n].cpu())
    print(f"threads={nt}: torchCPU-vs-Coracle cv {np.abs(t_cv.numpy()-o_cv).max():.2e} att {np.abs(t_at.numpy()-o_at).max():.2e}")
print(f"GPU tcgen05-vs-Coracle cv {np.abs(outs[0][0][:n].cpu().numpy()-o_cv).max():.2e} att {np.abs(outs[0][1][:n].cpu().numpy()-o_at).max():.2e}")
print(f"GPU ffma-vs-Coracle cv {np.abs(cv_f[:n].cpu().numpy()-o_cv).max():.2e} att {np.abs(at_f[:n].cpu().numpy()-o_at).max():.2e}")
print("mkldnn", torch.backends.mkldnn.is_available(), "matmul precision", torch.get_float32_matmul_precision())
