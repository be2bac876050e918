import sys, os, ctypes
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from bench import WORKLOADS, synth_params, synth_pool
from code2vec_b200 import _lib, functional as CF
w = dict(WORKLOADS["cfg2"]); dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib = _lib.load()
p = synth_params(w, dev); s, pth, e, lab = synth_pool(w, 64, dev, 1234)
B, L, H, C = w["B"], w["L"], w["H"], w["C"]
dims = CF.make_dims(w["T"], w["P"], C, w["Et"], w["Ep"], H)
params = CF.make_params(p["terminal_embedding.weight"], p["path_embedding.weight"], p["input_linear.weight"],
                        p["input_layer_norm.weight"], p["input_layer_norm.bias"], p["attention_parameter"],
                        p["output_linear.weight"], p["output_linear.bias"])
cv = torch.empty((B, H), dtype=torch.float32, device=dev); att = torch.empty((B, L), dtype=torch.float32, device=dev)
ws_n = lib.c2v_encode_workspace_bytes(ctypes.byref(dims), B, L)
ws = torch.empty((ws_n,), dtype=torch.uint8, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
_lib.check(lib.c2v_encode_forward(ctypes.byref(dims), ctypes.byref(params), P(s[0:B]), P(pth[0:B]), P(e[0:B]), B, L, None,
                                  P(cv), P(att), P(ws), ws_n, _lib.ALGO_AUTO, st), "encode")
torch.cuda.synchronize()
a = (cv.clone(), att.clone())
b = CF.encode_forward(dims, params, s[:B], pth[:B], e[:B], algo=_lib.ALGO_FFMA)
dcv = (a[0]-b[0]).abs().max(1).values; dat = ((a[1]-b[1]).abs()/b[1].clamp_min(1e-12))
bad = (dcv > 2e-6).nonzero().flatten().tolist()
badrows = (dat > 1e-4).nonzero()
print(f"first cv {dcv.max().item():.2e} att_rel {dat.max().item():.2e} bad bags {len(bad)}: {bad[:40]}")
if len(badrows):
    rows = (badrows[:,0]*L + badrows[:,1]).tolist()
    tiles = sorted(set(r//128 for r in rows))
    print("bad rows", len(rows), "tiles", tiles[:40], "rows in tile", sorted(set(r%128 for r in rows))[:64])
