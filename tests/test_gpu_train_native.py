"""GPU parity of the native training step (BASELINE config 4): train-mode forward with batch
statistics, WeightedBCE, and every parameter gradient, against
  * one iteration of the VERBATIM reference (train.py:96-103) on CPU in float64 -- tests/golden/train_reference.npz,
    written by tests/golden/make_train_golden.py -- and
  * float64 autograd of the same network at the config-4 shape (64 crops x 30 frames).
Tolerance: relative L2 error per gradient tensor <= 1e-3.

ReLU ties: a gradient is discontinuous where a BatchNorm output crosses zero.  An element with |y| below the forward
rounding error may be masked either way by ANY float32 implementation (the reference's own float32 run differs from its
float64 run in the same way), and one flip moves a tensor's relative error to ~1e-2.  So
  * the strict golden cases are small and seeded tie-free (every |BN output| >= 2e-5 in float64): full 1e-3 bar;
  * the larger golden cases carry a flip-sized tolerance;
  * the full-size check runs float64 autograd with THIS implementation's ReLU masks forced (decoded from its conv outputs
    through sedb_debug_train_layout), which tests every kernel to the 1e-3 bar at any size, and separately asserts that
    the masks differ from the float64 ones only at near-ties."""
import copy
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import sed_b200
from sed_b200 import _ext
from sed_b200.models._native import aligned_ptr
from sed_b200.models.spectogram_models import interpolate
from sed_b200.utils.common import WeightedBCE
from oracle import cnn_ref
import refmodels

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL_GRAD = 1e-3
TOL_GRAD_WITH_FLIPS = 5e-2


def seeded_train_model(cfg=refmodels.MAIN_CFG):
    from sed_b200.models.spectogram_models import Cnn_AvgPooling
    torch.manual_seed(0)
    m = Cnn_AvgPooling(1, model_config=cfg)
    m.load_state_dict(cnn_ref.randomize_bn_({k: v.clone() for k, v in m.state_dict().items()}, seed=7))
    return m


def case_inputs(B, T, Tt, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, T, 64, generator=g) * 1.5
    y = (torch.rand(B, Tt, 1, generator=g) > 0.8).float()
    return x, y


@pytest.mark.parametrize("name,tol", [("B2_T16", TOL_GRAD), ("B4_T8", TOL_GRAD), ("B3_T13", TOL_GRAD),
                                      ("B4_T30", TOL_GRAD_WITH_FLIPS), ("B2_T182", TOL_GRAD_WITH_FLIPS)])
def test_train_step_vs_verbatim_reference(name, tol):
    gold = np.load(os.path.join(GOLD, "train_reference.npz"))
    B, T, Tt, seed = (int(v) for v in gold[f"{name}_shape"])
    m = seeded_train_model().cuda()
    m.train()
    x, y = case_inputs(B, T, Tt, seed)
    out = m(x.cuda())
    assert out.grad_fn is not None and type(out.grad_fn).__name__.startswith("_NativeTrainFunction")
    loss = WeightedBCE(recall_factor=5, multi_frame=True)(out, y.cuda())
    loss.backward()
    assert np.abs(out.detach().cpu().numpy() - gold[f"{name}_logits"]).max() < 1e-4
    assert abs(loss.item() - float(gold[f"{name}_loss"])) < 1e-5 * float(gold[f"{name}_loss"])
    names = [n for n, _ in m.named_parameters()]
    assert names == list(gold["param_names"])
    worst, bad = 0.0, []
    for i, (n, p) in enumerate(m.named_parameters()):
        g = p.grad.detach().double().cpu().numpy().reshape(-1)
        idx, val = gold[f"{name}_grad_{i}_idx"], gold[f"{name}_grad_{i}_val"]
        rel = np.linalg.norm(g[idx] - val) / np.linalg.norm(val)
        rel_norm = abs(np.linalg.norm(g) - float(gold[f"{name}_grad_{i}_norm"])) / float(gold[f"{name}_grad_{i}_norm"])
        if not (rel < tol and rel_norm < tol):
            bad.append((n, float(rel), float(rel_norm)))
        worst = max(worst, rel)
    print(f"{name}: worst relative gradient error {worst:.2e} (tolerance {tol:g})")
    assert not bad, bad
    sd = m.state_dict()
    for k in sd:
        if "running" in k:
            assert np.allclose(sd[k].cpu().numpy(), gold[f"{name}_{k}"], rtol=2e-5, atol=2e-6), k
        if "num_batches_tracked" in k:
            assert int(sd[k]) == 1


def native_conv_outputs(m, x):
    """The conv outputs Z_l (pre-BatchNorm, float32) of the last native train-mode forward, decoded from the workspace."""
    lib = _ext.load()
    B, _, T, _ = x.shape
    h = m._train_handle(x.device)
    lay = (ctypes.c_longlong * 256)()
    _ext.check(lib.sedb_debug_train_layout(h, B, T, lay, 256))
    ws = m._native.workspace(x.device, ("train", B, T), 0)
    ptr, _ = aligned_ptr(ws)
    raw = ws[ptr.value - ws.data_ptr():]
    stats = raw[:int(lay[4 + 4]) // 8 * 8].view(torch.float64)        # the statistics precede the first plane buffer
    zs = []
    for l in range(lay[0]):
        C, H, W, _, zo, zS = [int(lay[4 + 12 * l + i]) for i in range(6)]
        so = int(lay[4 + 12 * l + 10])
        t = raw[zo:zo + B * (C // 8) * zS * 32].view(torch.float32).view(B, C // 8, zS, 8)[:, :, 8:8 + (H + 2) * (W + 2)]
        t = t.reshape(B, C // 8, H + 2, W + 2, 8)[:, :, 1:H + 1, 1:W + 1]
        zs.append((t.permute(0, 1, 4, 2, 3).reshape(B, C, H, W).clone(), stats[so:so + 2 * C].clone()))
    return zs


class _MaskedRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return x * mask

    @staticmethod
    def backward(ctx, g):
        return g * ctx.saved_tensors[0], None


def float64_step_with_masks(ref, x, y, crit, zs_native):
    """One float64 iteration of the network with the ReLU masks of the native run: mask = fmaf(z, a, b) > 0 with a, b
    formed from the run's own batch sums exactly as csrc/cnn_train.cuh::bn_consts does.  Returns the loss, the number of
    elements whose mask differs from the float64 one and the largest |float64 BN output| among them."""
    h = x.double()
    flips, worst_tie = 0, 0.0
    k = 0
    for blk in ref.conv_blocks:
        for conv, bn, pool in ((blk.conv1, blk.bn1, 1), (blk.conv2, blk.bn2, blk.pool_size)):
            z = conv(h)
            yb = bn(z)
            zn, sums = zs_native[k]
            C = zn.shape[1]
            n = zn.numel() // C
            mean = sums[:C] * (1.0 / n)
            var = (sums[C:] * (1.0 / n) - mean * mean).clamp_min(0.0)
            rstd = (1.0 / torch.sqrt(var + float(np.float32(1e-5)))).float()
            a = bn.weight.float() * rstd                                            # __fmul_rn
            b = (bn.bias.double() - mean.float().double() * a.double()).float()     # __fmaf_rn(-mean, a, beta)
            mask = (zn.double() * a.double()[None, :, None, None] + b.double()[None, :, None, None]) > 0    # sign of the fma
            diff = mask != (yb > 0)
            flips += int(diff.sum())
            if diff.any():
                worst_tie = max(worst_tie, float(yb[diff].abs().max()))
            h = _MaskedRelu.apply(yb, mask.double())
            if pool != 1:
                h = F.avg_pool2d(h, pool)
            k += 1
    o = torch.mean(h, dim=3).transpose(1, 2)
    o = interpolate(ref.event_fc(o), 2 ** ref.num_pools)
    return crit(o, y.double()), flips, worst_tie


@pytest.mark.parametrize("cfg,B,T", [(refmodels.MAIN_CFG, 64, 30), (refmodels.MAIN_CFG, 5, 61), ([(32, 2), (64, 1)], 7, 9),
                                     (refmodels.DEFAULT_CFG, 3, 16)])
def test_train_step_vs_float64_autograd_same_masks(cfg, B, T):
    """BASELINE config 4 shape (64 crops x 30 frames) and others: every gradient within 1e-3 of float64 autograd once the
    (measure-zero, but at this size inevitable) ReLU ties are taken out of the comparison."""
    m = seeded_train_model(cfg).cuda()
    ref = copy.deepcopy(m).double()
    ref.native_training = False
    m.train(); ref.train()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(B, 1, T, 64, device="cuda", generator=g) * 1.5
    y = (torch.rand(B, T, 1, device="cuda", generator=g) > 0.8).float()
    crit = WeightedBCE(recall_factor=5, multi_frame=True)
    loss = crit(m(x), y)
    loss.backward()
    loss_ref, flips, worst_tie = float64_step_with_masks(ref, x, y, crit, native_conv_outputs(m, x))
    loss_ref.backward()
    n_elems = sum(int(z.numel()) for z, _ in native_conv_outputs(m, x))
    print(f"B={B} T={T}: {flips} of {n_elems} ReLU masks differ from float64 (largest |y| among them {worst_tie:.1e})")
    assert flips <= max(4, n_elems // 50000) and worst_tie < 1e-4
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * abs(loss_ref.item())
    worst = 0.0
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        rel = float((p.grad.double() - q.grad).norm() / q.grad.norm())
        worst = max(worst, rel)
        assert rel < TOL_GRAD, (n, rel)
    print(f"worst relative gradient error {worst:.2e}")
    for (k, v), (_, w) in zip(m.state_dict().items(), ref.state_dict().items()):
        if "running" in k:
            assert torch.allclose(v.double(), w, rtol=2e-5, atol=2e-6), k
    # a second iteration reuses plan and workspace; eval afterwards sees the updated running statistics
    m.zero_grad()
    crit(m(x), y).backward()
    m.eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        want = cnn_ref.cnn_avgpooling_forward(sd, x[:2].cpu(), [p for _, p in cfg])
        assert (m(x[:2]).cpu() - want).abs().max() < 5e-3


def test_bce_with_logits_kernel_matches_torch():
    lib = _ext.load()
    _ext.context()
    g = torch.Generator().manual_seed(0)
    for (B, Fo, Ft, K) in ((64, 24, 30, 1), (3, 176, 182, 2), (2, 8, 5, 1)):
        x = (torch.randn(B, Fo, K, generator=g) * 3).cuda().requires_grad_(True)
        y = (torch.rand(B, Ft, K, generator=g) > 0.7).float().cuda()
        n = min(Fo, Ft)
        ref = torch.nn.functional.binary_cross_entropy_with_logits(x[:, :n], y[:, :n], pos_weight=torch.tensor([5.0]).cuda())
        ref.backward()
        loss = torch.zeros(1, device="cuda")
        dx = torch.empty_like(x)
        p = lambda t: ctypes.c_void_p(t.data_ptr())     # noqa: E731
        _ext.check(lib.sedb_bce_with_logits(p(x.detach()), p(y), B, Fo, Ft, K, 5.0, 1.0, p(loss), p(dx), None))
        assert abs(float(loss) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
        assert torch.allclose(dx, x.grad, atol=1e-9, rtol=1e-5)


def test_adam_dev_matches_torch_and_counts_steps():
    lib = _ext.load()
    _ext.context()
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(5003, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True)
    p = p0.clone().cuda()
    m, v, vm = (torch.zeros_like(p) for _ in range(3))
    state = torch.tensor([0.0, 1e-3], device="cuda")
    hyper = torch.zeros(2, device="cuda")
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())      # noqa: E731
    for step in range(1, 5):
        grad = torch.randn(5003, generator=g) * (0.1 if step % 2 else 3.0)
        ref.grad = grad.clone()
        opt.step()
        _ext.check(lib.sedb_adam_amsgrad_step_dev(ptr(p), ptr((grad * 2).cuda()), ptr(m), ptr(v), ptr(vm), p.numel(),
                                                  ptr(state), ptr(hyper), 0.9, 0.999, 1e-8, 0.0, 0.5, None))
        assert torch.allclose(p.cpu(), ref.data, atol=2e-7, rtol=1e-5)
    assert float(state[0]) == 4.0
