"""-m gpu: backward of the CUDA path (c2v_encode_backward / c2v_label_backward through autograd)
against (1) the gradients the unmodified reference's autograd produced (tests/golden/grad_*.npz),
(2) torch-CPU autograd over the pinned restatement for a loss that also uses `attention`,
(3) the C oracle with the kernel's own dropout mask in training mode."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden_names, load_golden
from gpu_util import cuda, model_from_golden, random_batch, random_params
from code2vec_b200 import functional as CF

pytestmark = pytest.mark.gpu

KEYS = ["terminal_embedding.weight", "path_embedding.weight", "input_linear.weight", "input_layer_norm.weight",
        "input_layer_norm.bias", "attention_parameter", "output_linear.weight", "output_linear.bias"]


def _tol(ref):
    return 2e-5 * max(1.0, float(np.abs(ref).max()))      # fp32 sums of O(1e3) atomically-added terms


@pytest.mark.parametrize("name", golden_names("grad_"))
@pytest.mark.parametrize("algo", ["ffma", "auto"])
def test_gradients_match_reference_autograd(name, algo):
    rec = load_golden(name)
    m = model_from_golden(rec, algo=algo).train()         # dropout_prob == 0 in these fixtures
    out, cv, att = m.forward(cuda(rec["starts"]), cuda(rec["paths"]), cuda(rec["ends"]), cuda(rec["label"]))
    loss = F.nll_loss(F.log_softmax(out, dim=1), cuda(rec["label"]))       # main.py:251-264
    assert abs(loss.item() - float(rec["loss"])) <= 1e-5 * max(1.0, abs(float(rec["loss"])))
    loss.backward()
    got = dict(m.named_parameters())
    for k in KEYS:
        g = got[k].grad.cpu().numpy()
        assert np.abs(g - rec["grads"][k]).max() <= _tol(rec["grads"][k]), k


def _fp64_reference_gradients(p, starts, paths, ends, label, wa, wc):
    """fp64 evaluation of model.py:44-88 + the test loss with LayerNorm written out in elementwise ops.
    Why not `oracle.torch_forward` in fp32 here (what round 1 did): on the 128-thread GPU-box host the FIRST fp32
    ATen-CPU evaluation in a fresh process is occasionally off by ~1e-4 relative (always in LayerNorm's weight gradient,
    sometimes in everything downstream); recomputed in the same process it is right, and the CUDA gradients agreed with
    fp64 to 0.08 x tolerance in every such run (profiles/r2_flake_root_cause.txt: 7 of 240 fresh processes).  That --
    not a kernel -- was the "intermittent gradient mismatch" of round 1."""
    dt = torch.float64
    tp = {k: torch.from_numpy(v).to(dt).requires_grad_(True) for k, v in p.items()}
    s, pp, e = (torch.from_numpy(a) for a in (starts, paths, ends))
    H = p["input_linear.weight"].shape[0]
    c = torch.cat((F.embedding(s, tp["terminal_embedding.weight"]), F.embedding(pp, tp["path_embedding.weight"]),
                   F.embedding(e, tp["terminal_embedding.weight"])), 2)                       # model.py:48-51
    x = F.linear(c, tp["input_linear.weight"])                                               # model.py:54
    xv = x.view(-1, H)
    mu = xv.mean(1, keepdim=True); var = xv.var(1, unbiased=False, keepdim=True)
    y = (xv - mu) / torch.sqrt(var + 1e-5) * tp["input_layer_norm.weight"] + tp["input_layer_norm.bias"]   # model.py:55-56
    h = torch.tanh(y).view(x.shape)                                                          # model.py:57
    mask = (s > 0).to(dt)                                                                    # model.py:64
    z = (h * tp["attention_parameter"]).sum(2) * mask + (1 - mask) * (-3.4e38)               # model.py:92-93
    att = F.softmax(z, 1)                                                                    # model.py:96
    cv = (h * att.unsqueeze(-1)).sum(1)                                                      # model.py:68-69
    out = F.linear(cv, tp["output_linear.weight"], tp["output_linear.bias"])                 # model.py:83
    ((att * torch.from_numpy(wa).to(dt)).sum() + (cv * torch.from_numpy(wc).to(dt)).sum() + 0.1 * out.square().sum()).backward()
    return {k: v.grad.numpy() for k, v in tp.items()}, (out.detach().numpy(), cv.detach().numpy(), att.detach().numpy())


def test_gradients_with_attention_in_the_loss():
    from oracle import oracle
    rng = np.random.default_rng(5)
    T, P, C, E, H, B, L = 300, 200, 11, 128, 128, 7, 90
    p = random_params(rng, T, P, C, E, E, H)
    starts, paths, ends, label = random_batch(rng, B, L, T, P, C)
    starts[3, :] = 0                                       # all-pad bag: PAD row 0 does get gradient
    wa = rng.standard_normal((B, L)).astype(np.float32); wc = rng.standard_normal((B, H)).astype(np.float32)
    ref_grads, (ref_out, ref_cv, ref_att) = _fp64_reference_gradients(p, starts, paths, ends, label, wa, wc)
    # the fp64 evaluation is the pinned restatement: same forward as oracle.torch_forward (which is pinned to the reference)
    o32 = oracle.torch_forward({k: torch.from_numpy(v) for k, v in p.items()}, torch.from_numpy(starts),
                               torch.from_numpy(paths), torch.from_numpy(ends), torch.from_numpy(label))
    assert np.abs(o32[0].numpy() - ref_out).max() <= 2e-5 and np.abs(o32[2].numpy() - ref_att).max() <= 2e-6
    rec = {"opt": {"T": T, "P": P, "C": C, "Et": E, "Ep": E, "H": H}, "params": p}

    m = model_from_golden(rec).train()
    out2, cv2, att2 = m.forward(cuda(starts), cuda(paths), cuda(ends), cuda(label))
    ((att2 * cuda(wa)).sum() + (cv2 * cuda(wc)).sum() + 0.1 * out2.square().sum()).backward()
    got = dict(m.named_parameters())
    bad = {}
    assert np.abs(cv2.detach().cpu().numpy() - ref_cv).max() <= 1e-5 and np.abs(att2.detach().cpu().numpy() - ref_att).max() <= 1e-5
    for k in KEYS:
        ref = ref_grads[k]
        g = got[k].grad.cpu().numpy()
        d = np.abs(g - ref)
        err = float(np.nanmax(d)) if not np.isnan(d).all() else float("nan")
        if not err <= _tol(ref):                           # (also catches NaN)
            idx = np.unravel_index(np.nanargmax(np.where(np.isnan(d), np.inf, d)), d.shape)
            bad[k] = {"err": err, "tol": _tol(ref), "n_bad": int((~(d <= _tol(ref))).sum()), "n_nan": int(np.isnan(g).sum()),
                      "at": tuple(int(i) for i in idx), "got": float(g[idx]), "ref": float(ref[idx])}
    pad_grad = float(np.abs(got["terminal_embedding.weight"].grad[0].cpu().numpy()).max())
    # strict, no retry: the round-1 intermittent miss was the fp32 CPU reference, not the kernels
    assert not bad, bad
    assert pad_grad > 0


def test_training_mode_backward_regenerates_the_same_dropout_mask():
    from oracle import oracle
    from philox_ref import dropout_mask
    rng = np.random.default_rng(9)
    T, P, C, E, H, B, L = 200, 150, 9, 128, 128, 5, 64
    p = random_params(rng, T, P, C, E, E, H)
    starts, paths, ends, label = random_batch(rng, B, L, T, P, C)
    dims = CF.make_dims(T, P, C, E, E, H)
    tp = {k: cuda(v) for k, v in p.items()}
    params = CF.make_params(tp["terminal_embedding.weight"], tp["path_embedding.weight"], tp["input_linear.weight"],
                            tp["input_layer_norm.weight"], tp["input_layer_norm.bias"], tp["attention_parameter"],
                            tp["output_linear.weight"], tp["output_linear.bias"])
    seed, prob = 77, 0.25
    s, pp, e = cuda(starts), cuda(paths), cuda(ends)
    cv, att = CF.encode_forward(dims, params, s, pp, e, drop_p=prob, training=True, seed=seed)
    out = CF.label_logits(dims, params, cv)
    _, _, _, dout = CF.loss_argmax(out, cuda(label), want_grad=True)
    d_cv, d_w, d_b = CF.label_backward(dims, params, cv, dout)
    shapes = {"terminal_embedding": (T, E), "path_embedding": (P, E), "input_linear": (H, 3 * E), "ln_weight": (H,),
              "ln_bias": (H,), "attention": (H,)}
    g = CF.encode_backward(dims, params, s, pp, e, cv, att, d_cv, None, shapes, drop_p=prob, training=True, seed=seed)
    mask = dropout_mask(seed, B * L, H, prob).reshape(B, L, H)
    ref = oracle.backward(p, starts, paths, ends, dout.cpu().numpy(), dropmask=mask)
    names = {"terminal_embedding": "terminal_embedding.weight", "path_embedding": "path_embedding.weight",
             "input_linear": "input_linear.weight", "ln_weight": "input_layer_norm.weight",
             "ln_bias": "input_layer_norm.bias", "attention": "attention_parameter"}
    for k, rk in names.items():
        assert np.abs(g[k].cpu().numpy() - ref[rk]).max() <= _tol(ref[rk]), k
    assert np.abs(d_w.cpu().numpy() - ref["output_linear.weight"]).max() <= _tol(ref["output_linear.weight"])
    assert np.abs(d_b.cpu().numpy() - ref["output_linear.bias"]).max() <= _tol(ref["output_linear.bias"])


def test_one_adam_step_matches_reference_training_semantics():
    """main.py:171-175: zero_grad, forward, loss, backward, Adam.step -- parameters after one step equal the
    torch-CPU restatement's (same init, dropout off)."""
    from oracle import oracle
    rec = load_golden("grad_cfg2")
    m = model_from_golden(rec).train()
    opt = torch.optim.Adam(m.parameters(), lr=0.01, betas=(0.9, 0.999))
    s, pth, e, lab = (cuda(rec[k]) for k in ("starts", "paths", "ends", "label"))
    opt.zero_grad()
    out, _, _ = m.forward(s, pth, e, lab)
    F.nll_loss(F.log_softmax(out, dim=1), lab).backward()
    opt.step()
    tp = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in rec["params"].items()}
    ropt = torch.optim.Adam(list(tp.values()), lr=0.01, betas=(0.9, 0.999))
    out, _, _ = oracle.torch_forward(tp, *(torch.from_numpy(rec[k]) for k in ("starts", "paths", "ends", "label")))
    F.nll_loss(F.log_softmax(out, dim=1), torch.from_numpy(rec["label"])).backward()
    ropt.step()
    sd = m.state_dict()
    for k in KEYS:
        # Adam's first step moves every touched weight by ~lr regardless of gradient size, so compare loosely
        # where the gradient is ~0 and tightly elsewhere
        d = np.abs(sd[k].cpu().numpy() - tp[k].detach().numpy())
        assert np.quantile(d, 0.999) <= 2e-4, k


@pytest.mark.parametrize("E,H,B", [(128, 128, 256), (256, 256, 160), (200, 192, 40), (128, 256, 40), (256, 128, 40), (132, 100, 24)])
def test_tensor_core_dw_matches_the_cuda_core_gemm_at_many_tiles_per_cta(E, H, B):
    """K3b (dW = dX^T . C on tcgen05, MN-major operands) and K3c (dC = dX . W + scatter into the embedding gradients)
    against the CUDA-core kernels on up to 51,200 context rows = 400 tiles (every CTA walks several tiles: operand stage /
    ring reuse and the TMEM-resident partial), ragged bags and an all-pad bag included.  Sizes above 128 run as 128-wide
    windows of h and d (dW: one window pair per blockIdx.y; dC: the contraction over h in blocks through tensor memory)."""
    import os
    rng = np.random.default_rng(13)
    T, P, C, L = 3000, 2000, 16, 200
    p = random_params(rng, T, P, C, E, E, H)
    starts, paths, ends, label = random_batch(rng, B, L, T, P, C)
    starts[7, :] = 0
    dims = CF.make_dims(T, P, C, E, E, H)
    tp = {k: cuda(v) for k, v in p.items()}
    params = CF.make_params(tp["terminal_embedding.weight"], tp["path_embedding.weight"], tp["input_linear.weight"],
                            tp["input_layer_norm.weight"], tp["input_layer_norm.bias"], tp["attention_parameter"],
                            tp["output_linear.weight"], tp["output_linear.bias"])
    s, pp, e = cuda(starts), cuda(paths), cuda(ends)
    cv, att = CF.encode_forward(dims, params, s, pp, e)
    # gradients as small as a mean loss over a large batch makes them (|dx| ~ 1e-7: far below fp16's normal range;
    # the kernel rescales dX by a power of two found from max |dx|)
    d_cv = cuda((1e-6 * rng.standard_normal((B, H))).astype(np.float32))
    shapes = {"terminal_embedding": (T, E), "path_embedding": (P, E), "input_linear": (H, 3 * E), "ln_weight": (H,),
              "ln_bias": (H,), "attention": (H,)}
    os.environ["C2V_BACKWARD_DW"] = "ffma"; os.environ["C2V_BACKWARD_DC"] = "ffma"
    try:
        g_ffma = CF.encode_backward(dims, params, s, pp, e, cv, att, d_cv, None, shapes)
    finally:
        del os.environ["C2V_BACKWARD_DW"]; del os.environ["C2V_BACKWARD_DC"]
    g_tc = CF.encode_backward(dims, params, s, pp, e, cv, att, d_cv, None, shapes)
    a, b = g_tc["input_linear"].cpu().numpy(), g_ffma["input_linear"].cpu().numpy()
    assert 1e-8 < np.abs(b).max() < 1e-3
    assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()
    for k in ("terminal_embedding", "path_embedding", "ln_weight", "ln_bias", "attention"):
        assert torch.allclose(g_tc[k], g_ffma[k], rtol=0, atol=2e-5 * float(g_ffma[k].abs().max()))


@pytest.mark.parametrize("E,H,algo", [(128, 128, "tcgen05"), (100, 100, "tcgen05"), (256, 256, "tcgen05"), (128, 128, "ffma"), (36, 52, "ffma")])
def test_stashed_x_equals_the_input_linear_output_and_gives_the_same_gradients(E, H, algo):
    """c2v_encode_forward_stash keeps x = c . W^T (model.py:54) per context row; c2v_encode_backward_stashed must give
    the gradients of the recomputing backward (same kernels after x)."""
    from gpu_util import ALGOS
    rng = np.random.default_rng(E + H)
    T, P, C, B, L = 400, 300, 9, 11, 77
    p = random_params(rng, T, P, C, E, E, H)
    starts, paths, ends, label = random_batch(rng, B, L, T, P, C)
    starts[2, :] = 0
    dims = CF.make_dims(T, P, C, E, E, H)
    tp = {k: cuda(v) for k, v in p.items()}
    params = CF.make_params(tp["terminal_embedding.weight"], tp["path_embedding.weight"], tp["input_linear.weight"],
                            tp["input_layer_norm.weight"], tp["input_layer_norm.bias"], tp["attention_parameter"])
    s, pp, e = cuda(starts), cuda(paths), cuda(ends)
    cv, att, xs = CF.encode_forward(dims, params, s, pp, e, drop_p=0.25, training=True, seed=5, algo=ALGOS[algo], stash=True)
    c = np.concatenate([p["terminal_embedding.weight"][starts], p["path_embedding.weight"][paths],
                        p["terminal_embedding.weight"][ends]], axis=2).reshape(B * L, 3 * E).astype(np.float64)
    x_ref = c @ p["input_linear.weight"].astype(np.float64).T
    err = float(np.abs(xs.cpu().numpy() - x_ref).max())
    assert err <= 1e-5 * max(1.0, np.abs(x_ref).max()), err          # fp32 accumulation over K = 3E products (<= 768)
    d_cv = cuda(rng.standard_normal((B, H)).astype(np.float32))
    d_att = cuda(rng.standard_normal((B, L)).astype(np.float32))
    shapes = {"terminal_embedding": (T, E), "path_embedding": (P, E), "input_linear": (H, 3 * E), "ln_weight": (H,),
              "ln_bias": (H,), "attention": (H,)}
    g0 = CF.encode_backward(dims, params, s, pp, e, cv, att, d_cv, d_att, shapes, drop_p=0.25, training=True, seed=5)
    g1 = CF.encode_backward(dims, params, s, pp, e, cv, att, d_cv, d_att, shapes, drop_p=0.25, training=True, seed=5, x_stash=xs)
    for k in shapes:
        ref = g0[k].cpu().numpy()
        assert np.abs(g1[k].cpu().numpy() - ref).max() <= _tol(ref), k


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_flat_adam_matches_torch_adam_over_several_steps(wd):
    """c2v_adam_step (one launch: Adam + zero_grad + 1/world) against torch.optim.Adam (main.py:138) driving the same
    model through the same batches: parameters and losses agree step by step, the gradient bucket is left zeroed, and the
    model's weight-image caches notice the update."""
    from code2vec_b200.distributed import FlatAdam, FlatGradBucket, ddp_step
    rec = load_golden("grad_cfg2")
    s, pth, e, lab = (cuda(rec[k]) for k in ("starts", "paths", "ends", "label"))
    lf = lambda o_, l_: F.nll_loss(F.log_softmax(o_, dim=1), l_)
    m1 = model_from_golden(rec).train(); m2 = model_from_golden(rec).train()
    b1 = FlatGradBucket(m1.parameters()); o1 = FlatAdam(b1, lr=0.01, betas=(0.9, 0.999), weight_decay=wd)
    b2 = FlatGradBucket(m2.parameters()); o2 = torch.optim.Adam(m2.parameters(), lr=0.01, betas=(0.9, 0.999), weight_decay=wd)
    for step in range(4):
        l1 = ddp_step(m1, o1, b1, s, pth, e, lab, lf)
        l2 = ddp_step(m2, o2, b2, s, pth, e, lab, lf)
        assert abs(l1.item() - l2.item()) <= 2e-5 * max(1.0, abs(l2.item())), (step, l1.item(), l2.item())
        assert float(b1.flat.abs().max()) == 0.0
        for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
            assert n1 == n2
            assert (p1 - p2).abs().max().item() <= 3e-6 * max(1.0, p2.abs().max().item()), (step, n1)
    assert l1.item() < float(rec["loss"])            # it does learn


@pytest.mark.parametrize("name", ["angular_grad", "angular_grad128"])
def test_angular_head_gradients_match_reference_autograd(name):
    """training through the angular-margin head (model.py:71-80): c2v_angular_forward_train / c2v_angular_backward
    against the gradients the unmodified reference's autograd produced (oracle/gen_golden.py)"""
    rec = load_golden(name)
    m = model_from_golden(rec).train()
    out, cv, att = m.forward(cuda(rec["starts"]), cuda(rec["paths"]), cuda(rec["ends"]), cuda(rec["label"]))
    assert np.abs(out.detach().cpu().numpy() - rec["outputs"]).max() <= 1e-4
    loss = F.nll_loss(F.log_softmax(out, dim=1), cuda(rec["label"]))
    assert abs(loss.item() - float(rec["loss"])) <= 1e-5 * max(1.0, abs(float(rec["loss"])))
    loss.backward()
    for k, p in m.named_parameters():
        ref = rec["grads"][k]
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= _tol(ref), k


@pytest.mark.parametrize("name", ["grad_cfg2", "grad_cfg1", "grad_odd", "grad_wide", "grad_e200"])
def test_fused_grad_accumulation_equals_autograd_accumulation(name):
    """Code2Vec.fuse_grad_accumulation: the table / input_linear gradients are added straight into existing .grad buffers
    (what ddp_step uses with the flat optimizers) -- same result as letting autograd accumulate fresh gradient tensors."""
    rec = load_golden(name)
    s, p, e, lab = (cuda(rec[k]) for k in ("starts", "paths", "ends", "label"))
    res = []
    for fuse in (False, True):
        torch.manual_seed(3)
        m = model_from_golden(rec).train()
        for prm in m.parameters():                       # pre-existing, non-zero gradients: accumulation, not overwrite
            prm.grad = torch.randn_like(prm) * 0.01
        before = {k: v.grad.clone() for k, v in m.named_parameters()}
        m.fuse_grad_accumulation = fuse
        calls = []
        if fuse:                                         # the two-phase backward: hook runs once path_embedding.grad is complete
            path_before = m.path_embedding.weight.grad.clone()
            def hook():
                torch.cuda.synchronize()
                calls.append((m.path_embedding.weight.grad - path_before).cpu().numpy())
            m.on_path_grads_ready = hook
        ptrs = {k: v.grad.data_ptr() for k, v in m.named_parameters()}
        out, cv, att = m.forward(s, p, e, lab)
        F.nll_loss(F.log_softmax(out, dim=1), lab).backward()
        for k, v in m.named_parameters():
            assert v.grad.data_ptr() == ptrs[k], k       # .grad buffers stay where they are (flat-bucket views survive)
        res.append({k: (v.grad - before[k]).cpu().numpy() for k, v in m.named_parameters()})
        if fuse:
            assert len(calls) == 1
            ref = rec["grads"]["path_embedding.weight"]      # already complete when the hook ran, untouched afterwards
            assert np.abs(calls[0] - ref).max() <= _tol(ref) + 2e-8
            assert np.array_equal(calls[0], res[-1]["path_embedding.weight"]) or \
                np.abs(calls[0] - res[-1]["path_embedding.weight"]).max() <= 1e-9
    for k in res[0]:
        ref = rec["grads"][k]
        assert np.abs(res[1][k] - ref).max() <= _tol(ref) + 2e-8, k          # fused path vs the reference's autograd
        assert np.abs(res[1][k] - res[0][k]).max() <= 2e-6 * max(1.0, float(np.abs(ref).max())) + 2e-8, k
