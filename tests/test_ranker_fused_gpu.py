"""Fused ranker tail + loss: `ltr_triple_approx_ndcg_fwdbwd` (TMA + tcgen05 tf32 MMAs + the approxNDCG
pair loop in one kernel) against the stock torch.nn TripleLayerNet in fp32 followed by the oracle's approxNDCG loss and
torch autograd (architeture/tripleLayer.py:5-17, losses/approxNDCG.py:7-53, main_batch_execution.py:128-170).
Tolerance: tf32 operands (10-bit mantissa), fp32 accumulation -> 2e-3 relative on scores / loss, 1e-2 of the largest
entry on parameter gradients; the stated bar for the tensor-core path is 1e-3 (bf16-class)... measured errors are printed."""
import numpy as np
import pytest
import torch

from ltr_testing import make_slates

pytestmark = pytest.mark.gpu


def _reference_step(net_cpu, X, y, alpha):
    from oracle import ltr_oracle as O
    for prm in net_cpu.parameters():
        prm.grad = None
    scores = net_cpu(X, None, None).squeeze(-1)
    loss = O.approx_ndcg_loss(scores, y, alpha=alpha)
    loss.backward()
    return scores.detach(), loss.detach(), {k: v.grad.clone() for k, v in net_cpu.named_parameters()}


@pytest.mark.timeout(300)
@pytest.mark.parametrize("B,seed", [(1, 0), (5, 1), (300, 2)])
def test_fused_triple_layer_step_matches_torch_modules(B, seed):
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    torch.manual_seed(seed)
    net = TripleLayerNet(136)
    g = torch.Generator().manual_seed(100 + seed)
    X = torch.rand(B, 128, 136, generator=g)                     # MSLR "Norm" features are min-max scaled to [0, 1]
    _, y = make_slates(B, 128, seed=seed, min_len=64)
    X[y == -1] = 0.0
    alpha = 10.0
    s_ref, l_ref, g_ref = _reference_step(net, X, y, alpha)

    dev_net = TripleLayerNet(136).cuda()
    dev_net.load_state_dict(net.state_dict())
    from ltr_b200.architeture.tripleLayer import fused_triple_approx_ndcg
    loss, scores = fused_triple_approx_ndcg(dev_net, X.cuda(), y.cuda(), alpha=alpha, return_scores=True)
    loss.backward()
    torch.cuda.synchronize()
    s = scores.cpu()
    valid = y != -1
    s_err = float((s - s_ref)[valid].abs().max() / s_ref[valid].abs().max())
    l_err = abs(float(loss) - float(l_ref)) / abs(float(l_ref))
    print(f"B={B}: score err {s_err:.2e}, loss rel err {l_err:.2e}")
    assert s_err < 2e-3
    assert l_err < 2e-3
    for name, prm in dev_net.named_parameters():
        gr = g_ref[name]
        # l3.bias: d loss / d b3 = sum of d loss / d score = 0 up to rounding (approxNDCG is shift invariant): measure it
        # against the scale of the neighbouring l3.weight gradient instead of against its own ~1e-9
        scale = g_ref["l3.weight"].abs().max() if name == "l3.bias" else gr.abs().max().clamp_min(1e-30)
        err = float((prm.grad.cpu() - gr).abs().max() / scale)
        print(f"   grad {name}: rel-to-max err {err:.2e}")
        assert err < 1e-2, name
    # the module's own method and the stock forward agree with each other
    again = dev_net.approx_ndcg_loss(X.cuda(), y.cuda(), alpha=alpha)
    assert torch.equal(again.detach(), loss.detach()), "bitwise repeatable"
    fwd = dev_net(X.cuda(), None, None).squeeze(-1).cpu()
    assert float((fwd - s).abs().max()) < 2e-3 * float(s.abs().max())


@pytest.mark.timeout(300)
def test_fused_step_full_config2_batch_properties():
    """BASELINE configs[1] size: 4096 slates x 128 docs x 136 features.  The loss equals the standalone approxNDCG kernel
    on the fused kernel's own scores (same arithmetic family: 1e-5), padded docs change nothing, no NaN anywhere."""
    from ltr_b200.architeture.tripleLayer import TripleLayerNet, fused_triple_approx_ndcg
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    torch.manual_seed(3)
    net = TripleLayerNet(136).cuda()
    g = torch.Generator().manual_seed(7)
    B = 4096
    X = torch.rand(B, 128, 136, generator=g).cuda()
    _, y = make_slates(B, 128, seed=11, min_len=64)
    y = y.cuda()
    loss, scores = fused_triple_approx_ndcg(net, X, y, alpha=10.0, return_scores=True)
    loss.backward()
    ref = approxNDCGLoss(scores.clone().requires_grad_(True), y, alpha=10.0)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    grads = [p.grad.clone() for p in net.parameters()]
    assert all(bool(torch.isfinite(t).all()) for t in grads)
    X2 = X.clone()
    X2[y == -1] = 123.0                                          # features of padded docs must not matter
    net.zero_grad()
    loss2 = fused_triple_approx_ndcg(net, X2, y, alpha=10.0)
    loss2.backward()
    assert abs(float(loss2) - float(loss)) <= 1e-6 * abs(float(loss))
    for a, b in zip(grads, [p.grad for p in net.parameters()]):
        assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max())


def _fused_vs_reference(net, X, y, alpha, eps=1e-10, s_tol=2e-3, l_tol=2e-3, g_tol=1e-2):
    """Run the fused step and the fp32 torch modules + oracle loss on the same inputs; return the measured errors."""
    from ltr_b200.architeture.tripleLayer import TripleLayerNet, fused_triple_approx_ndcg
    from oracle import ltr_oracle as O
    for prm in net.parameters():
        prm.grad = None
    s_ref = net(X, None, None).squeeze(-1)
    l_ref = O.approx_ndcg_loss(s_ref, y, eps=eps, alpha=alpha)
    l_ref.backward()
    g_ref = {k: v.grad.clone() for k, v in net.named_parameters()}
    dev_net = TripleLayerNet(X.shape[2]).cuda()
    dev_net.load_state_dict(net.state_dict())
    loss, scores = fused_triple_approx_ndcg(dev_net, X.cuda(), y.cuda(), eps=eps, alpha=alpha, return_scores=True)
    loss.backward()
    torch.cuda.synchronize()
    valid = y != -1
    s = scores.cpu()
    if bool(valid.any()):
        assert float((s - s_ref.detach())[valid].abs().max()) <= s_tol * float(s_ref.detach()[valid].abs().max())
    assert abs(float(loss) - float(l_ref)) <= l_tol * abs(float(l_ref)) + 1e-7, (float(loss), float(l_ref))
    for name, prm in dev_net.named_parameters():
        gr = g_ref[name]
        scale = g_ref["l3.weight"].abs().max() if name == "l3.bias" else gr.abs().max()
        err = float((prm.grad.cpu() - gr).abs().max())
        assert err <= g_tol * float(scale) + 1e-9, (name, err, float(scale))
    return float(loss)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("F", [8, 64, 132, 160])
def test_fused_step_feature_widths(F):
    """1, 2 and 5 feature chunks, a width that is not a multiple of 8 (the last k step is half zero fill)."""
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    torch.manual_seed(F)
    net = TripleLayerNet(F)
    g = torch.Generator().manual_seed(F)
    B = 37
    X = torch.rand(B, 128, F, generator=g)
    _, y = make_slates(B, 128, seed=F, min_len=70)
    _fused_vs_reference(net, X, y, alpha=10.0)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("B", [2, 3, 149, 600, 1500])
def test_fused_step_slate_counts(B):
    """Fewer slates than compute groups, one CTA with two slates, several rounds of every group; short slates
    (2 and 3 docs per lane), single-doc and fully padded slates in the batch."""
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    torch.manual_seed(B)
    net = TripleLayerNet(136)
    g = torch.Generator().manual_seed(B)
    X = torch.rand(B, 128, 136, generator=g)
    _, y = make_slates(B, 128, seed=B, min_len=1)
    y[0, 1:] = -1.0                                              # one doc
    if B > 2:
        y[B // 2] = -1.0                                         # no doc at all: slate value 0, no gradient
    _fused_vs_reference(net, X, y, alpha=10.0)


@pytest.mark.timeout(300)
def test_fused_step_padding_anywhere_and_fractional_labels():
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    torch.manual_seed(5)
    net = TripleLayerNet(136)
    g = torch.Generator().manual_seed(5)
    B = 64
    X = torch.rand(B, 128, 136, generator=g)
    _, y = make_slates(B, 128, seed=5, min_len=30, pad_anywhere=True)
    frac = torch.rand(B, 128, generator=g) * 4.0
    y[::2] = torch.where(y[::2] == -1, y[::2], frac[::2])        # every other slate: labels off the 0..4 grid
    y[1, :10] = 7.0                                              # integer labels above 4
    _fused_vs_reference(net, X, y, alpha=10.0)


@pytest.mark.timeout(300)
def test_fused_step_wide_score_range_takes_direct_form():
    """alpha * log2(e) * (max - min score) >= 126: the factorised pair sigmoid would overflow; the direct form runs."""
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    torch.manual_seed(6)
    net = TripleLayerNet(136)
    with torch.no_grad():
        net.l3.weight.mul_(150.0)
    g = torch.Generator().manual_seed(6)
    B = 48
    X = torch.rand(B, 128, 136, generator=g)
    _, y = make_slates(B, 128, seed=6, min_len=64)
    s = net(X, None, None).squeeze(-1)
    assert float((s.max(dim=1).values - s.min(dim=1).values).max()) * 10.0 * 1.4427 > 126.0
    # saturated sigmoids: gradients of both implementations are tiny differences of O(1) terms
    _fused_vs_reference(net, X, y, alpha=10.0, g_tol=3e-2)


@pytest.mark.timeout(300)
def test_fused_step_large_eps_clamps_pairs():
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    torch.manual_seed(8)
    net = TripleLayerNet(136)
    with torch.no_grad():
        net.l3.weight.mul_(4.0)
    g = torch.Generator().manual_seed(8)
    B = 40
    X = torch.rand(B, 128, 136, generator=g)
    _, y = make_slates(B, 128, seed=8, min_len=64)
    _fused_vs_reference(net, X, y, alpha=10.0, eps=1e-3, g_tol=3e-2)


@pytest.mark.timeout(300)
def test_folded_entry_equals_whole_step_entry():
    """`ltr_triple_approx_ndcg_fwdbwd` (caller folds l1 / l2 itself) against `ltr_triple_approx_ndcg_step` (fold and unfold on
    the device): same scores bit for bit when fed the device-folded weights' torch equivalent to ~1e-6, and gradients that
    unfold to the same six tensors."""
    from ltr_b200 import _lib, _runtime as rt
    from ltr_b200.architeture.tripleLayer import TripleLayerNet, fused_triple_approx_ndcg
    lib = _lib.load()
    torch.manual_seed(21)
    dev = torch.device("cuda", 0)
    net = TripleLayerNet(136).cuda()
    B, F = 333, 136
    g = torch.Generator().manual_seed(21)
    X = torch.rand(B, 128, F, generator=g).cuda()
    _, y = make_slates(B, 128, seed=21, min_len=40)
    y = y.cuda()
    loss, scores = fused_triple_approx_ndcg(net, X, y, alpha=10.0, return_scores=True)
    loss.backward()
    W1, b1, W2, b2, W3, b3 = [p.detach() for p in (net.l1.weight, net.l1.bias, net.l2.weight, net.l2.bias, net.l3.weight, net.l3.bias)]
    Wc = (W2 @ W1).contiguous()
    bc = (W2 @ b1 + b2).contiguous()
    w3 = W3.reshape(-1).contiguous()
    sc2 = torch.empty(B, 128, device=dev)
    slate = torch.empty(B, device=dev)
    out = torch.zeros(1 + 32 * F + 65, device=dev)
    nscr = lib.ltr_triple_scratch_bytes(B, 128, F)
    scr = torch.zeros(nscr, dtype=torch.uint8, device=dev)
    tab = rt.disc_prefix(dev)
    rc = lib.ltr_triple_approx_ndcg_fwdbwd(
        X.data_ptr(), y.data_ptr(), Wc.data_ptr(), bc.data_ptr(), w3.data_ptr(), float(b3), B, 128, F, 10.0, 1e-10, -1.0,
        tab.data_ptr(), tab.numel(), sc2.data_ptr(), slate.data_ptr(), out[0:1].data_ptr(), out[1:1 + 32 * F].data_ptr(),
        out[1 + 32 * F:33 + 32 * F].data_ptr(), out[33 + 32 * F:65 + 32 * F].data_ptr(), out[65 + 32 * F:].data_ptr(),
        scr.data_ptr(), nscr, rt.workspace(dev).data_ptr(), rt.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    assert float((sc2 - scores).abs().max()) <= 2e-5 * float(scores.abs().max())      # cuBLAS fold vs the fold kernel: fp32 rounding
    assert abs(float(out[0]) - float(loss)) <= 1e-5 * abs(float(loss))
    gWc, gbc, gw3 = out[1:1 + 32 * F].view(32, F), out[1 + 32 * F:33 + 32 * F], out[33 + 32 * F:65 + 32 * F]
    want = {"l1.weight": W2.t() @ gWc, "l1.bias": W2.t() @ gbc, "l2.weight": gWc @ W1.t() + torch.outer(gbc, b1), "l2.bias": gbc,
            "l3.weight": gw3.view(1, 32)}
    for name, prm in net.named_parameters():
        if name in want:
            scale = float(want[name].abs().max())
            assert float((prm.grad - want[name]).abs().max()) <= 2e-4 * scale, name


@pytest.mark.timeout(300)
@pytest.mark.parametrize("B,F", [(1, 136), (7, 136), (450, 136), (64, 64)])
def test_fused_listnet_step_matches_torch_modules(B, F):
    """`ltr_triple_listnet_step` against the stock modules + the oracle's listnetLoss (losses/listnet.py:5-16: a SUM over the
    batch, every doc of a slate counts, padding labels included) + torch autograd."""
    from ltr_b200.architeture.tripleLayer import TripleLayerNet, fused_triple_listnet
    from oracle import ltr_oracle as O
    torch.manual_seed(B + F)
    net = TripleLayerNet(F)
    g = torch.Generator().manual_seed(B)
    X = torch.rand(B, 128, F, generator=g)
    _, y = make_slates(B, 128, seed=B, min_len=64)
    s_ref = net(X, None, None).squeeze(-1)
    l_ref = O.listnet_loss(y, s_ref)
    l_ref.backward()
    dev_net = TripleLayerNet(F).cuda()
    dev_net.load_state_dict(net.state_dict())
    loss, scores = fused_triple_listnet(dev_net, X.cuda(), y.cuda(), return_scores=True)
    loss.backward()
    torch.cuda.synchronize()
    assert float((scores.cpu() - s_ref.detach()).abs().max()) <= 2e-3 * float(s_ref.detach().abs().max())
    assert abs(float(loss) - float(l_ref)) <= 2e-3 * abs(float(l_ref))
    for (name, prm), ref in zip(dev_net.named_parameters(), net.parameters()):
        scale = net.l3.weight.grad.abs().max() if name == "l3.bias" else ref.grad.abs().max()
        assert float((prm.grad.cpu() - ref.grad).abs().max()) <= 1e-2 * float(scale) + 1e-9, name
    again = dev_net.listnet_loss(X.cuda(), y.cuda())
    assert torch.equal(again.detach(), loss.detach()), "bitwise repeatable"


@pytest.mark.timeout(300)
def test_fused_step_propagates_nan_like_the_modules():
    """A NaN feature of one doc makes that slate's scores NaN in the stock modules and the batch loss NaN in approxNDCGLoss;
    the fused step must not hide it (and must not hang or trap: the direct-form path takes such a slate)."""
    from ltr_b200.architeture.tripleLayer import TripleLayerNet, fused_triple_approx_ndcg
    torch.manual_seed(9)
    net = TripleLayerNet(136).cuda()
    g = torch.Generator().manual_seed(9)
    X = torch.rand(20, 128, 136, generator=g).cuda()
    _, y = make_slates(20, 128, seed=9, min_len=64)
    X[3, 5, 17] = float("nan")
    loss, scores = fused_triple_approx_ndcg(net, X, y.cuda(), alpha=10.0, return_scores=True)
    torch.cuda.synchronize()
    assert bool(torch.isnan(loss)) and bool(torch.isnan(scores[3, 5]))
    assert bool(torch.isfinite(scores[0]).all()) and bool(torch.isfinite(scores[4]).all())
    # and the next call on clean data is clean again
    X[3, 5, 17] = 0.5
    loss2 = fused_triple_approx_ndcg(net, X, y.cuda(), alpha=10.0)
    assert bool(torch.isfinite(loss2))
