"""one fresh-process execution of tests/test_backward_parity_gpu.py::test_gradients_with_attention_in_the_loss through the
MODEL path (Code2Vec + autograd), the first GPU work of the process, with every tensor that enters / leaves the backward
kernels recorded and compared with an fp64 torch-CPU evaluation.  'ok' or 'MISMATCH {...}' (exit 1)."""
import os, sys, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, torch.nn.functional as F
from gpu_util import cuda, model_from_golden, random_params, random_batch
from code2vec_b200 import functional as CF

rng = np.random.default_rng(5)
T, P, C, E, H, B, L = 300, 200, 11, 128, 128, 7, 90
p = random_params(rng, T, P, C, E, E, H)
starts, paths, ends, label = random_batch(rng, B, L, T, P, C)
starts[3, :] = 0
wa = rng.standard_normal((B, L)).astype(np.float32); wc = rng.standard_normal((B, H)).astype(np.float32)
dt = torch.float64
tp = {k: torch.from_numpy(v).to(dt).requires_grad_(True) for k, v in p.items()}
s, pp, e = (torch.from_numpy(a) for a in (starts, paths, ends))
c = torch.cat((F.embedding(s, tp["terminal_embedding.weight"]), F.embedding(pp, tp["path_embedding.weight"]),
               F.embedding(e, tp["terminal_embedding.weight"])), 2)
x = F.linear(c, tp["input_linear.weight"])
xv = x.view(-1, H)
mu = xv.mean(1, keepdim=True); var = xv.var(1, unbiased=False, keepdim=True)
y = (xv - mu) / torch.sqrt(var + 1e-5) * tp["input_layer_norm.weight"] + tp["input_layer_norm.bias"]
h = torch.tanh(y).view(B, L, H)
mask = (s > 0).to(dt)
z = (h * tp["attention_parameter"]).sum(2) * mask + (1 - mask) * (-3.4e38)
att = F.softmax(z, 1); att.retain_grad()
cv = (h * att.unsqueeze(-1)).sum(1); cv.retain_grad()
out = F.linear(cv, tp["output_linear.weight"], tp["output_linear.bias"]); out.retain_grad()
((att * torch.from_numpy(wa).to(dt)).sum() + (cv * torch.from_numpy(wc).to(dt)).sum() + 0.1 * out.square().sum()).backward()

rec = {}
_eb, _lb = CF.encode_backward, CF.label_backward
def eb(dims, params, starts_, paths_, ends_, cv_, att_, d_cv, d_att, shapes, *a, **k):
    rec.update(x_stash=k.get("x_stash"), cv_in=cv_.clone(), att_in=att_.clone(), d_cv=d_cv.clone(),
               d_att=None if d_att is None else d_att.clone())
    return _eb(dims, params, starts_, paths_, ends_, cv_, att_, d_cv, d_att, shapes, *a, **k)
def lb(dims, params, cv_, d_out, *a, **k):
    rec.update(d_out=d_out.clone())
    r = _lb(dims, params, cv_, d_out, *a, **k)
    rec.update(d_cv_lab=r[0].clone())
    return r
CF.encode_backward, CF.label_backward = eb, lb
import code2vec_b200.model as M

m = model_from_golden({"opt": {"T": T, "P": P, "C": C, "Et": E, "Ep": E, "H": H}, "params": p}).train()
out2, cv2, att2 = m.forward(cuda(starts), cuda(paths), cuda(ends), cuda(label))
((att2 * cuda(wa)).sum() + (cv2 * cuda(wc)).sum() + 0.1 * out2.square().sum()).backward()
torch.cuda.synchronize()

def rel(a, b):
    a = a.detach().cpu().double().numpy(); b = b.detach().double().numpy()
    d = np.abs(a - b)
    i = np.unravel_index(np.nanargmax(np.where(np.isnan(d), np.inf, d)), d.shape)
    return [float("%.3g" % (float(d[i]) / max(1e-30, float(np.abs(b).max())))), [int(v) for v in i], int(np.isnan(a).sum())]

KEYS = ["terminal_embedding.weight", "path_embedding.weight", "input_linear.weight", "input_layer_norm.weight",
        "input_layer_norm.bias", "attention_parameter", "output_linear.weight", "output_linear.bias"]
got = dict(m.named_parameters())
bad = {}
for k in KEYS:
    ref = tp[k].grad.numpy(); g = got[k].grad.cpu().numpy()
    tol = 2e-5 * max(1.0, float(np.abs(ref).max()))
    err = float(np.nanmax(np.abs(g - ref)))
    if not err <= tol:
        bad[k] = round(err / tol, 2)
if bad:
    stages = {"outputs": rel(out2, out), "code_vector": rel(cv2, cv), "attention": rel(att2, att)}
    if rec.get("x_stash") is not None:
        stages["x_stash"] = rel(rec["x_stash"].view(B, L, H), x)
    stages["bw.cv_in"] = rel(rec["cv_in"], cv); stages["bw.att_in"] = rel(rec["att_in"], att)
    stages["bw.d_out"] = rel(rec["d_out"], out.grad); stages["bw.d_cv_lab"] = rel(rec["d_cv_lab"], cv.grad - torch.from_numpy(wc).to(dt))
    stages["bw.d_cv"] = rel(rec["d_cv"], cv.grad); stages["bw.d_att"] = rel(rec["d_att"], torch.from_numpy(wa).to(dt))
    for k in KEYS:
        stages["g." + k] = rel(got[k].grad, tp[k].grad)
    print("MISMATCH", json.dumps(bad), "STAGES", json.dumps(stages), flush=True)
    sys.exit(1)
print("ok")
