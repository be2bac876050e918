"""Kernel-level timing of the label head (K2 and its loss / backward variants), CUDA events, whole calls.
    B=1024 C=8192 H=128 python scripts/time_label.py          (cfg2)      C=195299 H=100 ... (cfg3, the top11 label count)"""
import sys, os, ctypes
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from code2vec_b200 import _lib, functional as CF
lib = _lib.load()
B, C, H = int(os.environ.get("B", 1024)), int(os.environ.get("C", 8192)), int(os.environ.get("H", 128))
ONLY = os.environ.get("ONLY")
dev = torch.device("cuda:0")
cv = torch.tanh(torch.randn(B, H, device=dev)); w = torch.randn(C, H, device=dev) * 0.1; b = torch.zeros(C, device=dev)
lab = torch.randint(0, C, (B,), device=dev)
dims = CF.make_dims(10, 10, C, H, H, H)
params = CF.make_params(None, None, None, None, None, None, w, b)
out = torch.empty(B, C, device=dev); am = torch.empty(B, dtype=torch.int64, device=dev); mx = torch.empty(B, device=dev)
loss = torch.empty((), device=dev); lse = torch.empty(B, device=dev)
dcv = torch.empty(B, H, device=dev); dw = torch.empty(C, H, device=dev); db = torch.empty(C, device=dev)
n = lib.c2v_label_workspace_bytes(ctypes.byref(dims), B); ws = torch.empty(n, dtype=torch.uint8, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
D, PR = ctypes.byref(dims), ctypes.byref(params)
def run(kind, reuse=True):
    a = (0x100 if reuse else 0)
    if kind == "logits": return lib.c2v_label_logits(D, PR, P(cv), B, P(out), P(ws), n, a, st)
    if kind == "logits+argmax": return lib.c2v_label_logits_argmax(D, PR, P(cv), B, P(out), P(am), P(mx), P(ws), n, a, st)
    if kind == "loss+argmax, no logits": return lib.c2v_label_loss_argmax(D, PR, P(cv), P(lab), B, None, P(loss), P(lse), P(am), P(mx), P(ws), n, a, st)
    if kind == "loss+argmax+logits": return lib.c2v_label_loss_argmax(D, PR, P(cv), P(lab), B, P(out), P(loss), P(lse), P(am), P(mx), P(ws), n, a, st)
    if kind == "loss pass over stored logits": return lib.c2v_loss_argmax(P(out), P(lab), B, C, P(loss), P(am), P(mx), None, st)
    if kind == "dlogits": return lib.c2v_label_dlogits(D, PR, P(cv), P(lab), P(lse), B, 1.0 / B, None, P(out), P(ws), n, a, st)
    if kind == "backward tensor cores": return lib.c2v_label_backward_ws(D, PR, P(cv), P(out), B, P(dcv), P(dw), P(db), P(ws), n, a, st)
    if kind == "backward tensor cores after dlogits":      # C2V_FLAG_GRAD_ABSMAX_READY: max |G| and the cv image come from the workspace
        return lib.c2v_label_backward_ws(D, PR, P(cv), P(out), B, P(dcv), P(dw), P(db), P(ws), n, a | 0x400, st)
    if kind == "backward cuda cores": return lib.c2v_label_backward(D, PR, P(cv), P(out), B, P(dcv), P(dw), P(db), st)
assert run("logits", False) == 0; torch.cuda.synchronize()
kinds = ["logits", "logits+argmax", "loss+argmax, no logits", "loss+argmax+logits", "loss pass over stored logits", "dlogits",
         "backward tensor cores", "backward tensor cores after dlogits", "backward cuda cores"]
for kind in kinds:
    if ONLY and ONLY not in kind: continue
    reps = 20 if C > 50000 else 50
    if kind.endswith("after dlogits"): assert run("dlogits") == 0
    for _ in range(3): assert run(kind) == 0, lib.c2v_last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): run(kind)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} C={C} H={H}  {kind:32s} {e0.elapsed_time(e1)/reps*1000:9.1f} us per call", flush=True)
