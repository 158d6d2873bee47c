"""approxNDCG: CUDA path (through the C-ABI) vs golden vectors from the reference and vs the oracle."""
import numpy as np
import pytest
import torch

from conftest import golden
from ltr_testing import assert_grad_close, assert_loss_close, make_slates

pytestmark = pytest.mark.gpu
G = golden()


def _run_gpu(s, y, **kw):
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    sd = s.cuda().requires_grad_(True)
    loss = approxNDCGLoss(sd, y.cuda(), **kw)
    loss.backward()
    return loss.detach().cpu().numpy(), sd.grad.cpu().numpy()


def _run_oracle(s, y, dtype=torch.float32, **kw):
    from oracle import ltr_oracle as O
    sc = s.to(dtype).clone().requires_grad_(True)
    loss = O.approx_ndcg_loss(sc, y.to(dtype), **kw)
    loss.backward()
    return loss.detach().numpy(), sc.grad.numpy()


def _check(s, y, **kw):
    """GPU vs the fp32 oracle, with the oracle's own fp32-vs-fp64 distance as extra slack (see ltr_testing)."""
    loss, grad = _run_gpu(s, y, **kw)
    rloss, rgrad = _run_oracle(s, y, **kw)
    _, rgrad64 = _run_oracle(s, y, dtype=torch.float64, **kw)
    assert_loss_close(loss, rloss)
    assert_grad_close(grad, rgrad, ref64=rgrad64)
    return loss, grad


@pytest.mark.parametrize("name", G.names("approx/"))
def test_golden(name):
    ins = G.inputs(name)
    loss, grad = _run_gpu(torch.from_numpy(ins["y_pred"]), torch.from_numpy(ins["y_true"]), **G.kwargs(name))
    assert loss.shape == G.out(name).shape == ()
    assert_loss_close(loss, G.out(name))
    _, g64 = _run_oracle(torch.from_numpy(ins["y_pred"]), torch.from_numpy(ins["y_true"]), dtype=torch.float64,
                         **G.kwargs(name))
    assert_grad_close(grad, G.grad(name, "y_pred"), ref64=g64)


@pytest.mark.parametrize("B,n,min_len,alpha,kw", [
    (64, 128, 64, 10.0, {}),
    (64, 128, 64, 1.0, {}),
    (33, 100, 1, 10.0, dict(pad_anywhere=True)),
    (16, 300, 150, 10.0, {}),
    (6, 1024, 512, 10.0, {}),
    (4, 1500, 700, 1.0, {}),
    (32, 128, 64, 10.0, dict(ties=True)),
    (32, 128, 64, 10.0, dict(score_scale=25.0)),     # range too wide for the factorised form
    (8, 7, 1, 10.0, {}),
])
def test_vs_oracle(B, n, min_len, alpha, kw):
    s, y = make_slates(B, n, seed=7 + n, min_len=min_len, **kw)
    loss, grad = _run_gpu(s, y, alpha=alpha)
    rloss, rgrad = _run_oracle(s, y, alpha=alpha)
    assert_loss_close(loss, rloss)
    assert_grad_close(grad, rgrad)
    assert np.all(grad[(y == -1).numpy()] == 0.0), "padded docs must get exactly zero gradient"


def test_every_valid_length_1_to_128():
    """One slate per valid length 0..128: every (R, T) shape of the block-circulant pair schedule,
    odd and even T, ragged last blocks."""
    n = 128
    s, y = make_slates(n + 1, n, seed=99)
    for b in range(n + 1):
        y[b, b:] = -1.0                      # slate b has b valid docs
    for alpha in (1.0, 10.0):
        _check(s, y, alpha=alpha)
    # per-slate values, not only their mean
    for b in list(range(1, 12)) + [31, 32, 33, 63, 64, 65, 66, 95, 96, 97, 127, 128]:
        _check(s[b:b + 1], y[b:b + 1], alpha=10.0)
        _check(s[b:b + 1], y[b:b + 1], alpha=1.0)


def test_wide_score_range_mixed_batch():
    """Slates beyond the factorised form's exponent budget (direct form) mixed with ordinary ones."""
    s, y = make_slates(40, 128, seed=17, min_len=20, pad_anywhere=True)
    s[::3] *= 40.0
    s[1::3] += 1000.0                        # large common offset, ordinary spread
    _check(s, y, alpha=10.0)


def test_fractional_labels_take_the_sort_path():
    s, y = make_slates(16, 96, seed=3, min_len=40)
    y = torch.where(y >= 0, y * 0.37, y)
    loss, grad = _run_gpu(s, y, alpha=5.0)
    rloss, rgrad = _run_oracle(s, y, alpha=5.0)
    assert_loss_close(loss, rloss)
    assert_grad_close(grad, rgrad)


def test_large_eps_clamp_is_exact():
    s, y = make_slates(16, 64, seed=5, min_len=30)
    for eps in (1e-3, 1e-6):
        loss, grad = _run_gpu(s, y, alpha=10.0, eps=eps)
        rloss, rgrad = _run_oracle(s, y, alpha=10.0, eps=eps)
        assert_loss_close(loss, rloss)
        assert_grad_close(grad, rgrad)


def test_full_size_properties():
    """BASELINE config 2 (4096 x 128): permutation invariance, padding invariance, oracle on a sample."""
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    from ltr_b200 import ops
    B, n = 4096, 128
    s, y = make_slates(B, n, seed=2020, min_len=64)
    sd, yd = s.cuda(), y.cuda()
    base = approxNDCGLoss(sd, yd, alpha=10.0)
    # determinism: bitwise-equal on repeat
    assert torch.equal(base, approxNDCGLoss(sd, yd, alpha=10.0))
    # permutation of docs inside every slate leaves the loss unchanged
    perm = torch.argsort(torch.rand(B, n, device="cuda"), dim=1)
    shuffled = approxNDCGLoss(sd.gather(1, perm), yd.gather(1, perm), alpha=10.0)
    assert abs(float(shuffled) - float(base)) <= 2e-6 * abs(float(base))
    # appending padded docs changes nothing
    sp = torch.cat([sd, torch.randn(B, 32, device="cuda")], dim=1)
    yp = torch.cat([yd, torch.full((B, 32), -1.0, device="cuda")], dim=1)
    padded = approxNDCGLoss(sp, yp, alpha=10.0)
    assert abs(float(padded) - float(base)) <= 1e-6 * abs(float(base))
    # mean of slates == loss; a 256-slate sample agrees with the oracle
    sub = slice(1000, 1256)
    loss_sub, grad_sub = _run_gpu(s[sub], y[sub], alpha=10.0)
    rloss, rgrad = _run_oracle(s[sub], y[sub], alpha=10.0)
    assert_loss_close(loss_sub, rloss)
    assert_grad_close(grad_sub, rgrad)


def test_cpu_tensors_in_cpu_loss_out():
    """Reference calling convention: CPU tensors in, CPU scalar out, .grad on the CPU leaf."""
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    s, y = make_slates(8, 40, seed=11, min_len=20)
    y0 = y.clone()
    sc = s.clone().requires_grad_(True)
    loss = approxNDCGLoss(sc, y, alpha=10.0)
    assert loss.device.type == "cpu" and loss.dim() == 0
    loss.backward(retain_graph=True)
    loss.backward()          # stash survives a retained graph (main_batch_execution.py:170)
    rloss, rgrad = _run_oracle(s, y, alpha=10.0)
    assert_loss_close(loss.detach().numpy(), rloss)
    assert_grad_close(sc.grad.numpy() / 2.0, rgrad)
    assert torch.equal(y, y0), "inputs must not be mutated"


def test_forward_only_and_errors():
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    from ltr_b200._lib import LtrError
    s, y = make_slates(4, 30, seed=13, min_len=10)
    loss = approxNDCGLoss(s.cuda(), y.cuda(), alpha=10.0)
    assert not loss.requires_grad
    rloss, _ = _run_oracle(s, y, alpha=10.0)
    assert_loss_close(loss.cpu().numpy(), rloss)
    with pytest.raises(LtrError):
        approxNDCGLoss(torch.zeros(1, 100000, device="cuda"), torch.zeros(1, 100000, device="cuda"))


# ---- batches of >= 16,384 slates take the one-warp-per-slate kernel (approx_ndcg_warp.cuh) ---------------------
@pytest.mark.parametrize("n,min_len,kw", [
    (128, 64, {}),
    (128, 1, dict(pad_anywhere=True)),
    (100, 30, {}),
    (37, 1, {}),
    (128, 64, dict(ties=True)),
    (128, 64, dict(score_scale=25.0)),       # direct form for the slates whose range is too wide
])
def test_large_batch_kernel_matches_small_batch_kernel_and_oracle(n, min_len, kw):
    """The same slates through the large-batch kernel (one call of 16,640 slates) and through the four-warps-per-slate
    kernel (calls of 4,160 slates): per-slate gradients agree to a few ulp (they are scaled by 1/B, so the small
    calls are rescaled), the batch loss is the mean; a slice is checked against the oracle itself."""
    B, parts = 16640, 4
    s, y = make_slates(B, n, seed=77, min_len=min_len, **kw)
    if min_len == 1:
        y[5] = -1.0                                       # a fully padded slate
    loss, grad = _run_gpu(s, y, alpha=10.0)
    step = B // parts
    losses, grads = [], []
    for i in range(parts):
        l, g = _run_gpu(s[i * step:(i + 1) * step], y[i * step:(i + 1) * step], alpha=10.0)
        losses.append(float(l))
        grads.append(g / parts)
    assert_loss_close(loss, np.float32(np.mean(losses)), rtol=2e-6)
    ref = np.concatenate(grads)
    assert np.all(np.isfinite(grad))
    assert np.abs(grad - ref).max() <= 2e-5 * np.abs(ref).max()
    # oracle on the first 256 slates (its gradient is scaled by 1/256 instead of 1/B)
    _, rgrad = _run_oracle(s[:256], y[:256], alpha=10.0)
    _, rgrad64 = _run_oracle(s[:256], y[:256], dtype=torch.float64, alpha=10.0)
    assert_grad_close(grad[:256] * (B / 256.0), rgrad, ref64=rgrad64)


def test_large_batch_kernel_is_bitwise_repeatable_and_flags_nan():
    B, n = 16384, 128
    s, y = make_slates(B, n, seed=78, min_len=64)
    l1, g1 = _run_gpu(s, y, alpha=10.0)
    l2, g2 = _run_gpu(s, y, alpha=10.0)
    assert l1.tobytes() == l2.tobytes() and g1.tobytes() == g2.tobytes()
    s[123, 7] = float("nan")
    l3, _ = _run_gpu(s, y, alpha=10.0)
    assert np.isnan(l3)


# ---- host-buffer entry (ltr_approx_ndcg_fwdbwd_host): pinned buffers are read / written by the kernel over PCIe,
#      pageable ones take the staged copies; both must give what the device-resident call gives ---------------------
@pytest.mark.parametrize("pinned", [True, False])
def test_host_buffer_entry_matches_device_entry(pinned):
    from ltr_b200.host import ApproxNDCGHost
    B, n = 512, 128
    s, y = make_slates(B, n, seed=31, min_len=40)
    loss, grad = _run_gpu(s, y, alpha=10.0)
    plan = ApproxNDCGHost(B, n, alpha=10.0, depth=2)
    hs, hy = (s.pin_memory(), y.pin_memory()) if pinned else (s.clone(), y.clone())
    for _ in range(3):                                    # slots are reused: results must not depend on the slot
        slot, hloss, hgrad = plan.submit(hs, hy)
        plan.wait(slot)
        assert hloss.numpy().tobytes() == np.float32(loss).tobytes()
        assert np.array_equal(hgrad.numpy(), grad)
    assert np.array_equal(hs.numpy(), s.numpy()) and np.array_equal(hy.numpy(), y.numpy())     # inputs untouched


# ---- 128 < n <= 1024: one cluster of 8 CTAs per slate (approx_ndcg_cluster.cuh) -------------------------------
@pytest.mark.parametrize("B,n,min_len,alpha,kw", [
    (5, 129, 129, 10.0, {}),                              # one full group + one doc
    (7, 132, 1, 10.0, dict(pad_anywhere=True)),
    (9, 200, 100, 10.0, {}),
    (4, 256, 256, 1.0, {}),
    (6, 516, 300, 10.0, dict(pad_anywhere=True)),
    (3, 1000, 900, 10.0, {}),
    (3, 1024, 1024, 10.0, {}),                            # 8 full groups
    (5, 640, 320, 10.0, dict(ties=True)),
    (4, 512, 256, 10.0, dict(score_scale=25.0)),          # direct form
    (300, 260, 130, 5.0, {}),                             # more clusters than fit the GPU at once
])
def test_cluster_kernel_vs_oracle(B, n, min_len, alpha, kw):
    s, y = make_slates(B, n, seed=5000 + n, min_len=min_len, **kw)
    if B >= 5:
        y[1] = -1.0                                       # a fully padded slate
        y[2, 3:] = -1.0                                   # three docs
        y[3] = torch.where(y[3] >= 0, y[3] * 1.5, y[3])  # fractional labels: ideal ranks by counting
    _check(s, y, alpha=alpha)


def test_cluster_kernel_is_bitwise_repeatable_and_propagates_nan():
    s, y = make_slates(40, 1024, seed=6001, min_len=512)
    l1, g1 = _run_gpu(s, y, alpha=10.0)
    l2, g2 = _run_gpu(s, y, alpha=10.0)
    assert l1.tobytes() == l2.tobytes() and g1.tobytes() == g2.tobytes()
    s[7, 11] = float("nan")
    l3, _ = _run_gpu(s, y, alpha=10.0)
    assert np.isnan(l3)
