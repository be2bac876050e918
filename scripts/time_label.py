import sys, os, ctypes
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from code2vec_b200 import _lib, functional as CF
lib = _lib.load()
B, C, H = 1024, 8192, 128
dev = torch.device("cuda:0")
cv = torch.tanh(torch.randn(B, H, device=dev)); w = torch.randn(C, H, device=dev) * 0.1; b = torch.zeros(C, device=dev)
dims = CF.make_dims(10, 10, C, H, H, H)
params = CF.make_params(None, None, None, None, None, None, w, b)
out = torch.empty(B, C, device=dev); am = torch.empty(B, dtype=torch.int64, device=dev); mx = torch.empty(B, device=dev)
n = lib.c2v_label_workspace_bytes(ctypes.byref(dims), B); ws = torch.empty(n, dtype=torch.uint8, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr()); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(kind, reuse):
    a = (0x100 if reuse else 0)
    if kind == "logits": return lib.c2v_label_logits(ctypes.byref(dims), ctypes.byref(params), P(cv), B, P(out), P(ws), n, a, st)
    if kind == "fused": return lib.c2v_label_logits_argmax(ctypes.byref(dims), ctypes.byref(params), P(cv), B, P(out), P(am), P(mx), P(ws), n, a, st)
    if kind == "ffma": return lib.c2v_label_logits(ctypes.byref(dims), ctypes.byref(params), P(cv), B, P(out), P(ws), n, 1, st)
    if kind == "argmax": return lib.c2v_loss_argmax(P(out), None, B, C, None, P(am), P(mx), None, st)
run("logits", False); torch.cuda.synchronize()
for kind in ("logits", "fused", "argmax", "ffma"):
    for _ in range(5): run(kind, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): run(kind, True)
    e1.record(); torch.cuda.synchronize()
    print(kind, f"{e0.elapsed_time(e1)/50*1000:.1f} us per call")
