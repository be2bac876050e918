"""Launch-level timeline of the forward step bench.py measures (torch profiler = CUPTI activity records, kernels run
back to back as in the bench, not serialised like under ncu): per-kernel average duration, and the idle gaps
between consecutive kernels of a step.

    python scripts/profile_step.py [workload]
"""
import ctypes
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from torch.profiler import ProfilerActivity, profile

from bench import WORKLOADS, synth_params, synth_pool
from code2vec_b200 import _lib, functional as CF

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = dict(WORKLOADS[wl]); dev = torch.device("cuda:0")
nb = 16
p = synth_params(w, dev); s, pth, e, lab = synth_pool(w, nb, dev, 99)
B, L = w["B"], w["L"]
dims = CF.make_dims(w["T"], w["P"], w["C"], w["Et"], w["Ep"], w["H"])
params = CF.make_params(p["terminal_embedding.weight"], p["path_embedding.weight"], p["input_linear.weight"],
                        p["input_layer_norm.weight"], p["input_layer_norm.bias"], p["attention_parameter"],
                        p["output_linear.weight"], p["output_linear.bias"])
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
cv = torch.empty((B, w["H"]), device=dev); att = torch.empty((B, L), device=dev)
out = torch.empty((B, w["C"]), device=dev); am = torch.empty((B,), dtype=torch.int64, device=dev); mx = torch.empty((B,), device=dev)
ws_n = lib.c2v_encode_workspace_bytes(ctypes.byref(dims), B, L); ws = torch.empty((ws_n,), dtype=torch.uint8, device=dev)
wl_n = lib.c2v_label_workspace_bytes(ctypes.byref(dims), B); wlb = torch.empty((wl_n,), dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def step(i, reuse=CF.REUSE_PREP):
    o = (i % nb) * B
    _lib.check(lib.c2v_encode_forward(ctypes.byref(dims), ctypes.byref(params), P(s[o:o + B]), P(pth[o:o + B]), P(e[o:o + B]),
                                      B, L, None, P(cv), P(att), P(ws), ws_n, _lib.ALGO_AUTO | reuse, st), "encode")
    _lib.check(lib.c2v_label_logits_argmax(ctypes.byref(dims), ctypes.byref(params), P(cv), B, P(out), P(am), P(mx),
                                           P(wlb), wl_n, _lib.ALGO_AUTO | reuse, st), "label")


step(0, 0)
for i in range(20):
    step(i)
torch.cuda.synchronize()
N = 50
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        step(i)
    torch.cuda.synchronize()
ev = sorted([e_ for e_ in prof.events() if e_.device_type.name == "CUDA"], key=lambda e_: e_.time_range.start)
tot = {}
for e_ in ev:
    d = tot.setdefault(e_.name[:60], [0, 0.0]); d[0] += 1; d[1] += e_.time_range.end - e_.time_range.start
span = ev[-1].time_range.end - ev[0].time_range.start
busy = sum(v[1] for v in tot.values())
print(f"{wl}: {span / N:.1f} us/step wall on the GPU, {busy / N:.1f} us/step inside kernels+memsets, {100 * (1 - busy / span):.1f} % idle gaps")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:60s} n/step={n / N:4.1f}  {t / N:7.2f} us/step  ({t / n:6.2f} us each)")
