"""stress: the model-path gradient comparison of tests/test_backward_parity_gpu.py::test_gradients_with_attention_in_the_loss
repeated in one process; prints every parameter whose gradient leaves the tolerance and the worst ratio seen."""
import os, sys
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
tp = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in p.items()}
out, cv, att = oracle.torch_forward(tp, torch.from_numpy(starts), torch.from_numpy(paths), torch.from_numpy(ends), torch.from_numpy(label))
((att * torch.from_numpy(wa)).sum() + (cv * torch.from_numpy(wc)).sum() + 0.1 * out.square().sum()).backward()
ref = {k: tp[k].grad.numpy() for k in KEYS}
rec = {"opt": {"T": T, "P": P, "C": C, "Et": E, "Ep": E, "H": H}, "params": p}
worst = {k: 0.0 for k in KEYS}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for it in range(n):
    m = model_from_golden(rec).train()
    out2, cv2, att2 = m.forward(cuda(starts), cuda(paths), cuda(ends), cuda(label))
    ((att2 * cuda(wa)).sum() + (cv2 * cuda(wc)).sum() + 0.1 * out2.square().sum()).backward()
    got = dict(m.named_parameters())
    for k in KEYS:
        tol = 2e-5 * max(1.0, float(np.abs(ref[k]).max()))
        err = float(np.abs(got[k].grad.cpu().numpy() - ref[k]).max())
        worst[k] = max(worst[k], err / tol)
        if err > tol:
            d = np.abs(got[k].grad.cpu().numpy() - ref[k])
            idx = np.unravel_index(d.argmax(), d.shape)
            print(f"iter {it}: {k} err {err:.3e} tol {tol:.3e} at {idx} ref {ref[k][idx]:.4e} n_bad {(d > tol).sum()}", flush=True)
print("worst err/tol per key:", {k: round(v, 3) for k, v in worst.items()})
