"""Stage-by-stage version of tests/test_backward_parity_gpu.py::test_gradients_with_attention_in_the_loss, first GPU work
of a fresh process: every intermediate of the CUDA path (x stash, code vector, attention, logits, d_cv, each gradient) is
compared with an fp64 torch-CPU evaluation of the same formulas.  Prints ONE line: 'ok' or 'BAD {...}' (exit 1).
Environment switches of the library (C2V_BACKWARD_DC/DW=ffma, C2V_NO_STASH=1) and FWD_ALGO=ffma select the components."""
import os, sys, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, torch.nn.functional as F
from gpu_util import cuda, random_params, random_batch
from code2vec_b200 import functional as CF, _lib

rng = np.random.default_rng(5)
T, P, C, E, H, B, L = 300, 200, 11, 128, 128, 7, 90
p = random_params(rng, T, P, C, E, E, H)
starts, paths, ends, label = random_batch(rng, B, L, T, P, C)
starts[3, :] = 0
wa = rng.standard_normal((B, L)).astype(np.float32); wc = rng.standard_normal((B, H)).astype(np.float32)
# ---- fp64 reference with every intermediate
dt = torch.float64
tp = {k: torch.from_numpy(v).to(dt).requires_grad_(True) for k, v in p.items()}
s, pp, e = (torch.from_numpy(a) for a in (starts, paths, ends))
c = torch.cat((F.embedding(s, tp["terminal_embedding.weight"]), F.embedding(pp, tp["path_embedding.weight"]),
               F.embedding(e, tp["terminal_embedding.weight"])), 2)
x = F.linear(c, tp["input_linear.weight"]); x.retain_grad()
xv = x.view(-1, H)
mu = xv.mean(1, keepdim=True); var = xv.var(1, unbiased=False, keepdim=True)
y = (xv - mu) / torch.sqrt(var + 1e-5) * tp["input_layer_norm.weight"] + tp["input_layer_norm.bias"]
h = torch.tanh(y).view(B, L, H)
mask = (s > 0).to(dt)
z = (h * tp["attention_parameter"]).sum(2) * mask + (1 - mask) * (-3.4e38)
att = F.softmax(z, 1); att.retain_grad()
cv = (h * att.unsqueeze(-1)).sum(1); cv.retain_grad()
out = F.linear(cv, tp["output_linear.weight"], tp["output_linear.bias"])
((att * torch.from_numpy(wa).to(dt)).sum() + (cv * torch.from_numpy(wc).to(dt)).sum() + 0.1 * out.square().sum()).backward()

# ---- CUDA path through the functional layer
g = {k: cuda(v) for k, v in p.items()}
dims = CF.make_dims(T, P, C, E, E, H)
params = CF.make_params(g["terminal_embedding.weight"], g["path_embedding.weight"], g["input_linear.weight"],
                        g["input_layer_norm.weight"], g["input_layer_norm.bias"], g["attention_parameter"],
                        g["output_linear.weight"], g["output_linear.bias"])
ds, dp_, de, dl = cuda(starts), cuda(paths), cuda(ends), cuda(label)
algo = {"ffma": _lib.ALGO_FFMA, "tcgen05": _lib.ALGO_TCGEN05}.get(os.environ.get("FWD_ALGO", "tcgen05"))
stash = os.environ.get("C2V_NO_STASH", "0") != "1"
res = CF.encode_forward(dims, params, ds, dp_, de, algo=algo, stash=stash)
cv2, att2 = res[0], res[1]
xs = res[2] if stash else None
out2 = CF.label_logits(dims, params, cv2)
d_out = 0.2 * out2
d_cv_lab, d_w, d_b = CF.label_backward(dims, params, cv2, d_out)
d_cv = d_cv_lab + cuda(wc)
d_att = cuda(wa)
shapes = {"terminal_embedding": (T, E), "path_embedding": (P, E), "input_linear": (H, 3 * E), "ln_weight": (H,),
          "ln_bias": (H,), "attention": (H,)}
gr = CF.encode_backward(dims, params, ds, dp_, de, cv2, att2, d_cv, d_att, shapes, x_stash=xs)
torch.cuda.synchronize()

def rel(a, b):
    a = a.detach().cpu().double().numpy() if hasattr(a, "detach") else np.asarray(a, np.float64)
    b = b.detach().double().numpy() if hasattr(b, "detach") else np.asarray(b, np.float64)
    d = np.abs(a - b)
    i = np.unravel_index(np.nanargmax(np.where(np.isnan(d), np.inf, d)), d.shape)
    return float(d[i]) / max(1e-30, float(np.abs(b).max())), [int(v) for v in i], int(np.isnan(a).sum())

checks = {}
if xs is not None:
    checks["x_stash"] = rel(xs.view(B, L, H), x)
checks["code_vector"] = rel(cv2, cv); checks["attention"] = rel(att2, att); checks["outputs"] = rel(out2, out)
checks["d_cv"] = rel(d_cv, cv.grad)
names = {"terminal_embedding": "terminal_embedding.weight", "path_embedding": "path_embedding.weight",
         "input_linear": "input_linear.weight", "ln_weight": "input_layer_norm.weight", "ln_bias": "input_layer_norm.bias",
         "attention": "attention_parameter"}
for k, rk in names.items():
    checks["g_" + k] = rel(gr[k], tp[rk].grad)
checks["g_out_w"] = rel(d_w, tp["output_linear.weight"].grad); checks["g_out_b"] = rel(d_b, tp["output_linear.bias"].grad)
LIM = {"x_stash": 3e-6, "code_vector": 5e-6, "attention": 5e-6, "outputs": 5e-6, "d_cv": 5e-6}
bad = {k: v for k, v in checks.items() if not (v[0] <= LIM.get(k, 1.2e-5) and v[2] == 0)}
if bad:
    print("BAD", json.dumps({k: [float("%.3g" % v[0]), v[1], v[2]] for k, v in checks.items()}), "FIRST:", list(bad)[:3], flush=True)
    sys.exit(1)
print("ok", json.dumps({k: float("%.2g" % v[0]) for k, v in checks.items()}))
