"""-m gpu: the loss fused into the label GEMM (model.py:83 + main.py:251-264 + main.py:285 in one
pass, logits never written) against oracle.loss_argmax on the oracle's own logits, its backward against torch-CPU
autograd over the pinned restatement, and the module's forward_loss() against forward() + calculate_loss."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from gpu_util import cuda, model_from_golden, random_batch, random_params
from code2vec_b200 import functional as CF

pytestmark = pytest.mark.gpu


def _case(rng, B, C, H, scale=1.0):
    cv = (rng.standard_normal((B, H)) * 0.5).astype(np.float32)
    W = (rng.uniform(-1, 1, (C, H)) / np.sqrt(H) * scale).astype(np.float32)
    b = (0.05 * rng.standard_normal(C)).astype(np.float32)
    lab = rng.integers(0, C, B).astype(np.int64)
    return cv, W, b, lab


@pytest.mark.parametrize("B,C,H,scale", [(7, 11, 128, 1.0), (130, 1000, 128, 8.0), (64, 4097, 100, 4.0), (33, 260, 256, 2.0),
                                          (1024, 8192, 128, 6.0), (5, 3, 4, 1.0)])
@pytest.mark.parametrize("want_logits", [False, True])
def test_fused_loss_matches_oracle(B, C, H, scale, want_logits):
    from oracle import oracle
    rng = np.random.default_rng(B * 31 + C)
    cv, W, b, lab = _case(rng, B, C, H, scale)
    lab[0] = C - 1; lab[-1] = 0                                   # first / last column as targets
    dims = CF.make_dims(10, 10, C, H, H, H)
    params = CF.make_params(None, None, None, None, None, None, cuda(W), cuda(b))
    assert CF.label_loss_supported(dims, B)
    loss, lse, am, mx, out = CF.label_loss(dims, params, cuda(cv), cuda(lab), want_logits=want_logits)
    ref_out = oracle.label_logits(cv, W, b)
    ref_loss, ref_am, ref_mx = oracle.loss_argmax(ref_out, lab)
    ref_lse = torch.logsumexp(torch.from_numpy(ref_out).double(), dim=1).numpy()
    assert abs(loss.item() - ref_loss) <= 2e-5 * max(1.0, abs(ref_loss)), (loss.item(), ref_loss)
    assert np.abs(lse.cpu().numpy() - ref_lse).max() <= 2e-5 * max(1.0, np.abs(ref_lse).max())
    got_mx = mx.cpu().numpy()
    assert np.abs(got_mx - ref_mx).max() <= 1e-4
    am_ = am.cpu().numpy()                                        # ties / near-ties: the kernel's pick must be a maximum
    assert np.all(ref_out[np.arange(B), am_] >= ref_mx - 1e-5)
    if want_logits:
        assert np.abs(out.cpu().numpy() - ref_out).max() <= 1e-4
    else:
        assert out is None


def test_fused_loss_at_the_top11_label_count():
    """C = 195,299 (top11_dataset's label vocabulary): 153 MB of logits at B = 200 that are never written"""
    from oracle import oracle
    rng = np.random.default_rng(7)
    B, C, H = 200, 195299, 100
    cv, W, b, lab = _case(rng, B, C, H, 5.0)
    dims = CF.make_dims(10, 10, C, H, H, H)
    params = CF.make_params(None, None, None, None, None, None, cuda(W), cuda(b))
    loss, lse, am, mx, out = CF.label_loss(dims, params, cuda(cv), cuda(lab))
    ref_out = oracle.label_logits(cv, W, b)
    ref_loss, ref_am, ref_mx = oracle.loss_argmax(ref_out, lab)
    assert out is None
    assert abs(loss.item() - ref_loss) <= 2e-5 * abs(ref_loss), (loss.item(), ref_loss)
    assert np.abs(mx.cpu().numpy() - ref_mx).max() <= 1e-4
    assert (am.cpu().numpy() == ref_am).mean() >= 0.99
    # backward: d loss / d logits recomputed tile by tile == softmax - onehot over B
    dout = CF.label_dlogits(dims, params, cuda(cv), cuda(lab), lse, 1.0 / B)
    t = torch.from_numpy(ref_out[:16]).double()
    ref_d = (torch.softmax(t, 1) - F.one_hot(torch.from_numpy(lab[:16]), C)) / B
    assert np.abs(dout[:16].cpu().numpy() - ref_d.numpy()).max() <= 1e-7


@pytest.mark.parametrize("name", ["grad_cfg2", "grad_cfg1", "grad_tiny"])
def test_forward_loss_equals_forward_plus_calculate_loss_and_its_gradients(name):
    """Code2Vec.forward_loss (fused) vs the reference's autograd gradients of mean NLL (tests/golden/grad_*.npz)"""
    rec = load_golden(name)
    m = model_from_golden(rec).train()
    s, p, e, lab = (cuda(rec[k]) for k in ("starts", "paths", "ends", "label"))
    loss, am, mx, cv, att = m.forward_loss(s, p, e, lab)
    assert abs(loss.item() - float(rec["loss"])) <= 1e-5 * max(1.0, abs(float(rec["loss"])))
    ref_out = rec["outputs"]
    assert np.abs(mx.cpu().numpy() - ref_out.max(1)).max() <= 1e-4
    assert np.abs(cv.detach().cpu().numpy() - rec["code_vector"]).max() <= 1e-4
    loss.backward()
    for k, g in m.named_parameters():
        ref = rec["grads"][k]
        tol = 2e-5 * max(1.0, float(np.abs(ref).max()))
        assert np.abs(g.grad.cpu().numpy() - ref).max() <= tol, k


def test_an_out_of_range_label_makes_the_loss_nan():
    rng = np.random.default_rng(1)
    cv, W, b, lab = _case(rng, 6, 40, 128)
    lab[2] = 40                                                   # the reference's NLLLoss raises "Target out of bounds"
    dims = CF.make_dims(10, 10, 40, 128, 128, 128)
    params = CF.make_params(None, None, None, None, None, None, cuda(W), cuda(b))
    loss = CF.label_loss(dims, params, cuda(cv), cuda(lab))[0]
    assert np.isnan(loss.item())


# ---- label backward on the tensor cores (c2v_label_backward_ws) vs fp64 matrix products ------------------------------
@pytest.mark.parametrize("B,C,H", [(7, 11, 128), (130, 1000, 128), (64, 4097, 100), (33, 260, 256), (1024, 8192, 128),
                                    (200, 777, 64), (5, 3, 4)])
@pytest.mark.parametrize("gscale", [1.0, 1e-6])
def test_label_backward_tensor_cores(B, C, H, gscale):
    rng = np.random.default_rng(B + C)
    cv, W, b, lab = _case(rng, B, C, H, 3.0)
    G = (rng.standard_normal((B, C)) * gscale).astype(np.float32)
    G[rng.random((B, C)) < 0.3] = 0.0
    dims = CF.make_dims(10, 10, C, H, H, H)
    w_t = cuda(W)
    params = CF.make_params(None, None, None, None, None, None, w_t, cuda(b))
    cache = CF.PrepCache()
    ref_dcv = G.astype(np.float64) @ W.astype(np.float64)
    ref_dw = G.astype(np.float64).T @ cv.astype(np.float64)
    ref_db = G.astype(np.float64).sum(0)
    for attempt in ("fresh image", "image reused"):               # second call: C2V_FLAG_REUSE_PREP path
        d_cv, d_w, d_b = CF.label_backward(dims, params, cuda(cv), cuda(G), cache=cache, weight=w_t)
        for got, ref, name in ((d_cv, ref_dcv, "d_cv"), (d_w, ref_dw, "d_w"), (d_b, ref_db, "d_b")):
            tol = 2e-5 * max(float(np.abs(ref).max()), 1e-30)
            assert np.abs(got.cpu().numpy() - ref).max() <= tol, (name, attempt)
    # and the CUDA-core path gives the same answer (no workspace)
    d_cv2, d_w2, d_b2 = CF.label_backward(dims, params, cuda(cv), cuda(G))
    assert np.abs(d_cv2.cpu().numpy() - ref_dcv).max() <= 2e-5 * max(float(np.abs(ref_dcv).max()), 1e-30)


@pytest.mark.parametrize("B,C,H", [(130, 1000, 128), (64, 4097, 100), (33, 260, 256), (1024, 8192, 128)])
def test_dlogits_leaves_the_gradient_maximum_for_the_label_backward(B, C, H):
    """c2v_label_dlogits keeps max |d logit| in the label workspace while it writes the gradient;
    c2v_label_backward_ws(C2V_FLAG_GRAD_ABSMAX_READY) then skips its own pass over [B, C]: same result as without the
    flag, and right against fp64 products of the same gradient."""
    rng = np.random.default_rng(B * 3 + C)
    cv, W, b, lab = _case(rng, B, C, H, 4.0)
    dims = CF.make_dims(10, 10, C, H, H, H)
    w_t = cuda(W)
    params = CF.make_params(None, None, None, None, None, None, w_t, cuda(b))
    cache = CF.PrepCache()
    cv_d, lab_d = cuda(cv), cuda(lab)       # ONE device tensor: the library honours the flag only for the code_vector pointer dlogits saw
    loss, lse, am, mx, out = CF.label_loss(dims, params, cv_d, lab_d, cache=cache, weight=w_t)
    G = CF.label_dlogits(dims, params, cv_d, lab_d, lse, 1.0 / B, cache=cache, weight=w_t)
    fast = CF.label_backward(dims, params, cv_d, G, cache=cache, weight=w_t, absmax_ready=True)
    slow = CF.label_backward(dims, params, cv_d, G, cache=cache, weight=w_t)
    # a different code_vector buffer (same values): the flag is ignored, the result is the same
    other = CF.label_backward(dims, params, cv_d.clone(), G, cache=cache, weight=w_t, absmax_ready=True)
    assert torch.equal(other[1], slow[1]) and torch.equal(other[2], slow[2])
    # d_w / d_b: one CTA owns a label tile (deterministic) -> identical bits; d_cv is a split-K sum through atomics
    assert torch.equal(fast[1], slow[1]) and torch.equal(fast[2], slow[2])
    assert (fast[0] - slow[0]).abs().max().item() <= 1e-6 * slow[0].abs().max().item()
    g64 = G.cpu().numpy().astype(np.float64)
    for got, ref in ((fast[0], g64 @ W.astype(np.float64)), (fast[1], g64.T @ cv.astype(np.float64)), (fast[2], g64.sum(0))):
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-5 * max(float(np.abs(ref).max()), 1e-30)
