"""Throughput of c2v_build_batch (on-GPU DatasetBuilder.build_data) on a synthetic corpus with top11's shape
(131 contexts per method on average, 16 % of the methods longer than 200, the longest 60,810), next to
the numpy restatement (oracle/batch_oracle.py) on one host core -- the reference's own builder is pure-Python list
handling (dataset_builder.py:112-150) and slower still."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch
from code2vec_b200.batch_builder import DeviceCorpus

rng = np.random.default_rng(0)
n_items = 200_000
n = np.minimum(np.maximum(rng.lognormal(4.3, 1.0, n_items).astype(np.int64), 1), 60810)
n[:3] = (60810, 20000, 5000)
off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
ctx = rng.integers(1, 300_000, (off[-1], 3), dtype=np.int64).astype(np.int32)
print(f"corpus: {n_items} methods, {off[-1]} contexts ({ctx.nbytes / 1e6:.0f} MB), mean {n.mean():.1f}, {100 * (n > 200).mean():.1f} % > 200")
c = DeviceCorpus(off, ctx, np.zeros(n_items, np.int64), 2, 1, "cuda:0")
B, L = 1024, 200
g = torch.Generator(device="cuda:0").manual_seed(0)
ids = [torch.randint(0, n_items, (B,), generator=g, device="cuda:0") for _ in range(32)]
for i in range(5):
    c.build(ids[i], L, i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200):
    c.build(ids[i % 32], L, i)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 200 * 1e3
valid = float(np.minimum(n, L).mean()) * B
print(f"GPU: {us:.1f} us per {B} x {L} batch = {B * L / us:.1f} M slots/s ({3 * 8 * B * L / us / 1e3:.1f} GB/s of int64 indices written)")
from oracle import batch_oracle as bo
t0 = time.perf_counter()
for i in range(3):
    bo.build_batch(off, ctx, ids[i].cpu().numpy(), L, i, 2, 1)
dt = (time.perf_counter() - t0) / 3
print(f"CPU (numpy restatement, 1 core): {dt * 1e3:.1f} ms per batch -> GPU/CPU = {dt * 1e6 / us:.0f}x")
