"""-m gpu: the drop-in module under the reference's own loop shape (INTEGRATION.md "one-line switch"):
/root/reference/main.py:160-177 (DataLoader of [b, L] int64 CPU tensors, ragged last batch, `.to(device)`, zero_grad,
model.forward, calculate_loss = log_softmax + NLLLoss(weight = 1), backward, torch.optim.Adam.step, loss.item()) and
main.py:267-297 (eval under no_grad, torch.max) -- its loss curve against the pinned torch-CPU restatement
(oracle.torch_forward) driven the same way, dropout off; and the deferred IndexError of the error surface."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset

from gpu_util import option_from, random_params
from code2vec_b200.model import Code2Vec

pytestmark = pytest.mark.gpu


class _Items(Dataset):                       # CodeDataset of model/dataset.py:14-38
    def __init__(self, ids, s, p, e, lab):
        self.ids, self.s, self.p, self.e, self.lab = ids, s, p, e, lab

    def __len__(self):
        return len(self.s)

    def __getitem__(self, i):
        return {"id": self.ids[i], "starts": self.s[i], "paths": self.p[i], "ends": self.e[i], "label": self.lab[i]}


def _dataset(rng, n, L, T, P, C):
    s = rng.integers(1, T, (n, L)); p = rng.integers(1, P, (n, L)); e = rng.integers(1, T, (n, L))
    cnt = rng.integers(1, L + 1, n)
    for i in range(n):                                         # zero-padded suffix, dataset_builder.py:145-147
        s[i, cnt[i]:] = 0; p[i, cnt[i]:] = 0; e[i, cnt[i]:] = 0
    lab = (s[:, 0] + p[:, 0]) % C                              # learnable from the first context
    t = lambda a: torch.tensor(a, dtype=torch.long)
    return _Items(list(range(n)), t(s), t(p), t(e), t(lab))


@pytest.mark.parametrize("E,H", [(128, 128), (100, 100), (12, 20)])
def test_reference_training_and_eval_loop_shape(E, H):
    from oracle import oracle
    rng = np.random.default_rng(E)
    T, P, C, L, n, bs = 300, 200, 17, 40, 150, 32            # 150 % 32 = 22: ragged last batch (drop_last unset, main.py:162)
    ds = _dataset(rng, n, L, T, P, C)
    params = random_params(rng, T, P, C, E, E, H)
    device = torch.device("cuda:0")
    opt = option_from({"T": T, "P": P, "C": C, "Et": E, "Ep": E, "H": H})
    model = Code2Vec(opt)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    model = model.to(device)                                   # main.py:132
    criterion = nn.NLLLoss(weight=torch.ones(C)).to(device)    # main.py:129-130: every label frequency is 1
    optimizer = torch.optim.Adam(model.parameters(), lr=0.01, betas=(0.9, 0.999), weight_decay=0.0)    # main.py:138
    ref = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in params.items()}
    ref_opt = torch.optim.Adam(list(ref.values()), lr=0.01, betas=(0.9, 0.999), weight_decay=0.0)
    ref_crit = nn.NLLLoss(weight=torch.ones(C))

    losses, ref_losses = [], []
    for epoch in range(5):                                     # 5 epochs x 5 batches = 25 optimizer steps
        g = torch.Generator().manual_seed(epoch)
        loader = DataLoader(ds, batch_size=bs, shuffle=True, generator=g, num_workers=0)            # main.py:161-162
        model.train()                                          # main.py:164
        for sample_batched in loader:
            starts = sample_batched["starts"].to(device); paths = sample_batched["paths"].to(device)
            ends = sample_batched["ends"].to(device); label = sample_batched["label"].to(device)
            optimizer.zero_grad()                              # main.py:171
            preds, _, _ = model.forward(starts, paths, ends, label)                                # main.py:172
            loss = criterion(F.log_softmax(preds, dim=1), label)                                    # main.py:251-264
            loss.backward()                                    # main.py:174
            optimizer.step()                                   # main.py:175
            losses.append(loss.item())                         # main.py:177
            ref_opt.zero_grad()
            rp, _, _ = oracle.torch_forward(ref, sample_batched["starts"], sample_batched["paths"], sample_batched["ends"],
                                            sample_batched["label"])
            rl = ref_crit(F.log_softmax(rp, dim=1), sample_batched["label"])
            rl.backward()
            ref_opt.step()
            ref_losses.append(rl.item())
    assert len(losses) == 25
    # Adam amplifies rounding differences (every step moves each weight by ~lr): the curves must agree closely at first
    # and stay together as training proceeds
    assert np.abs(np.array(losses[:5]) - np.array(ref_losses[:5])).max() <= 2e-4, (losses[:5], ref_losses[:5])
    assert np.abs(np.array(losses) - np.array(ref_losses)).max() <= 2e-2, (losses, ref_losses)
    assert losses[-1] < losses[0]

    # ---- main.py:267-297: eval pass, shuffled test loader, torch.max
    model.eval()
    with torch.no_grad():
        loader = DataLoader(ds, batch_size=bs, shuffle=True, generator=torch.Generator().manual_seed(9))
        test_loss, agree, total, rtest = 0.0, 0, 0, 0.0
        for sample_batched in loader:
            starts = sample_batched["starts"].to(device); paths = sample_batched["paths"].to(device)
            ends = sample_batched["ends"].to(device); label = sample_batched["label"].to(device)
            preds, code_vector, attention = model.forward(starts, paths, ends, label)
            test_loss += criterion(F.log_softmax(preds, dim=1), label).item()
            _, preds_label = torch.max(preds, dim=1)
            pl, cv2, at2, lab2 = model.predict(starts, paths, ends) if hasattr(model, "predict") else (preds_label,) * 4
            assert torch.equal(pl, preds_label)
            # the reference side, with the weights the drop-in has NOW (isolates the forward from the training drift)
            cur = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            rp, rcv, ratt = oracle.torch_forward(cur, sample_batched["starts"], sample_batched["paths"],
                                                 sample_batched["ends"], sample_batched["label"])
            assert (preds.cpu() - rp).abs().max() <= 1e-4 and (code_vector.cpu() - rcv).abs().max() <= 1e-4
            assert (attention.cpu() - ratt).abs().max() <= 1e-4
            rtest += ref_crit(F.log_softmax(rp, dim=1), sample_batched["label"]).item()
            agree += int((preds_label.cpu() == rp.max(dim=1)[1]).sum()); total += len(label)
        assert abs(test_loss - rtest) <= 1e-3 and agree >= total - 1


def test_out_of_range_indices_raise_index_error_one_call_late():
    """the reference: IndexError on CPU, a device assert -- surfacing at a later sync -- on CUDA (SURVEY 8b "Call")"""
    rng = np.random.default_rng(0)
    T, P, C, E, H = 50, 40, 5, 128, 128
    model = Code2Vec(option_from({"T": T, "P": P, "C": C, "Et": E, "Ep": E, "H": H}))
    model.load_state_dict({k: torch.from_numpy(v) for k, v in random_params(rng, T, P, C, E, E, H).items()})
    model = model.to("cuda:0").eval()
    good = torch.randint(1, 40, (4, 9), device="cuda:0")
    lab = torch.zeros(4, dtype=torch.long, device="cuda:0")
    with torch.no_grad():
        model.forward(good, good, good, lab)
        model.forward(good, good, good, lab)                   # nothing pending
        model.check_indices()
        bad = good.clone(); bad[2, 3] = T + 7                  # one start index outside the terminal table
        out, cv, att = model.forward(bad, good, good, lab)     # clamped to row 0, counted
        torch.cuda.synchronize()
        with pytest.raises(IndexError):
            model.forward(good, good, good, lab)               # raised by the NEXT call, without a sync of its own
        model.forward(good, good, good, lab)                   # the error is consumed
        bad2 = good.clone(); bad2[0, 0] = -1
        model.forward(good, bad2, good, lab)
        with pytest.raises(IndexError):
            model.check_indices()                              # explicit, synchronising form
        model.check_indices()
    # the same through the training path: the backward of the offending forward raises
    model.train()
    out, _, _ = model.forward(bad, good, good, lab)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        out.sum().backward()
