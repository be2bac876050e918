"""tests/test_backward_parity_gpu.py::test_gradients_with_attention_in_the_loss exactly as round 1 wrote it (fp32 torch-CPU
reference, model path, no instrumentation), first GPU work of a fresh process.  On a miss BOTH sides are re-judged against
an fp64 evaluation: 'gpu/64' = CUDA gradients vs fp64, 'cpu32/64' = the test's own fp32 CPU reference vs fp64
(both in units of the test tolerance).  Tells a wrong kernel from a noisy reference."""
import os, sys, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from gpu_util import cuda, model_from_golden, random_params, random_batch
from oracle import oracle
KEYS = ["terminal_embedding.weight", "path_embedding.weight", "input_linear.weight", "input_layer_norm.weight",
        "input_layer_norm.bias", "attention_parameter", "output_linear.weight", "output_linear.bias"]
rng = np.random.default_rng(5)
T, P, C, E, H, B, L = 300, 200, 11, 128, 128, 7, 90
p = random_params(rng, T, P, C, E, E, H)
starts, paths, ends, label = random_batch(rng, B, L, T, P, C)
starts[3, :] = 0
wa = rng.standard_normal((B, L)).astype(np.float32); wc = rng.standard_normal((B, H)).astype(np.float32)

def cpu_grads(dt):
    tp = {k: torch.from_numpy(v).to(dt).clone().requires_grad_(True) for k, v in p.items()}
    out, cv, att = oracle.torch_forward(tp, torch.from_numpy(starts), torch.from_numpy(paths), torch.from_numpy(ends),
                                        torch.from_numpy(label))
    ((att * torch.from_numpy(wa).to(dt)).sum() + (cv * torch.from_numpy(wc).to(dt)).sum() + 0.1 * out.square().sum()).backward()
    return {k: tp[k].grad.double().numpy() for k in KEYS}

ref32 = cpu_grads(torch.float32)                       # what the round-1 test compared against
m = model_from_golden({"opt": {"T": T, "P": P, "C": C, "Et": E, "Ep": E, "H": H}, "params": p}).train()
out2, cv2, att2 = m.forward(cuda(starts), cuda(paths), cuda(ends), cuda(label))
((att2 * cuda(wa)).sum() + (cv2 * cuda(wc)).sum() + 0.1 * out2.square().sum()).backward()
got = {k: v.grad.cpu().double().numpy() for k, v in m.named_parameters()}
tol = {k: 2e-5 * max(1.0, float(np.abs(ref32[k]).max())) for k in KEYS}
bad = {k: round(float(np.abs(got[k] - ref32[k]).max()) / tol[k], 2) for k in KEYS if not np.abs(got[k] - ref32[k]).max() <= tol[k]}
if bad:
    ref64 = cpu_grads(torch.float64)
    again32 = cpu_grads(torch.float32)
    rep = {k: {"gpu/32": bad.get(k, 0), "gpu/64": round(float(np.abs(got[k] - ref64[k]).max()) / tol[k], 2),
               "cpu32/64": round(float(np.abs(ref32[k] - ref64[k]).max()) / tol[k], 2),
               "cpu32again/64": round(float(np.abs(again32[k] - ref64[k]).max()) / tol[k], 2)} for k in bad}
    print("MISMATCH", json.dumps(rep), "threads", torch.get_num_threads(), flush=True)
    sys.exit(1)
print("ok")
