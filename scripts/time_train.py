"""launch-level breakdown of one training step (torch profiler, CUDA time per kernel);  WORKLOAD=cfg2|cfg3|cfg4|cfg5"""
import sys, os, types
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from bench import WORKLOADS, synth_params, synth_pool
from code2vec_b200.model import Code2Vec
from code2vec_b200.distributed import ShardedFlatAdam, ddp_step
w = dict(WORKLOADS[os.environ.get("WORKLOAD", "cfg2")]); dev = torch.device("cuda:0")
p = synth_params(w, dev); s, pth, e, lab = synth_pool(w, 8, dev, 1)
B = w["B"]
o = types.SimpleNamespace(terminal_count=w["T"], path_count=w["P"], label_count=w["C"], terminal_embed_size=w["Et"],
                          path_embed_size=w["Ep"], encode_size=w["H"], dropout_prob=0.25, angular_margin_loss=False,
                          angular_margin=0.5, inverse_temp=30.0, device=dev)
m = Code2Vec(o); m.load_state_dict(p); m = m.to(dev).train()
opt = ShardedFlatAdam(m.parameters(), lr=0.01); bucket = None
lf = None if os.environ.get("FUSED_LOSS", "1") == "1" else (lambda o_, l_: F.nll_loss(F.log_softmax(o_, dim=1), l_))
for i in range(3): ddp_step(m, opt, bucket, s[:B], pth[:B], e[:B], lab[:B], lf)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(4): ddp_step(m, opt, bucket, s[i*B:(i+1)*B], pth[i*B:(i+1)*B], e[i*B:(i+1)*B], lab[i*B:(i+1)*B], lf)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda k: -k.device_time_total)[:22]
for k in rows: print(f"{k.key[:70]:70s} n={k.count:4d} total={k.device_time_total/4:9.1f} us/step")
