"""K2 timing: label logits (+ fused arg-max) for B x C x H given on the command line:  B:C:H ..."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
from code2vec_b200 import _lib, functional as CF
dev = torch.device("cuda:0")
for spec in sys.argv[1:]:
    B, C, H = (int(x) for x in spec.split(":"))
    g = torch.Generator(device=dev).manual_seed(1)
    cv = torch.tanh(torch.randn(B, H, generator=g, device=dev))
    w = torch.randn(C, H, generator=g, device=dev) * 0.1
    b = torch.zeros(C, device=dev)
    dims = CF.make_dims(10, 10, C, H, H, H)
    params = CF.make_params(None, None, None, None, None, None, w, b)
    cache = CF.PrepCache()
    for arg in (True, False):
        f = (lambda: CF.label_logits_argmax(dims, params, cv, cache=cache, weight=w)) if arg else (lambda: CF.label_logits(dims, params, cv, cache=cache, weight=w))
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f"B={B} C={C} H={H} argmax={arg}: {us:.1f} us  logits write {B * C * 4 / us / 1e3:.0f} GB/s", flush=True)
