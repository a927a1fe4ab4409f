"""GPU edge cases of the drop-in surface: several classes, long inputs, workspace reuse across shapes, unaligned clips,
non-default streams."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import sed_b200
from sed_b200.dataset.spectogram import preprocess as P
from sed_b200.models.spectogram_models import Cnn_AvgPooling
from sed_b200.models.waveform_models import M5
from oracle import cnn_ref, logmel_ref as R
import refmodels
import signals


def _randomised(module, seed):
    sd = cnn_ref.randomize_bn_({k: v.clone() for k, v in module.state_dict().items()}, seed=seed)
    module.load_state_dict(sd)
    module.eval()
    return sd


def test_cnn_multiclass_and_small_config():
    torch.manual_seed(5)
    cfg = [(16, 2), (32, 2), (32, 1)]
    m = Cnn_AvgPooling(3, model_config=cfg)
    sd = _randomised(m, 21)
    x = refmodels.cnn_inputs(45, 77, batch=3)
    with torch.no_grad():
        ref = cnn_ref.cnn_avgpooling_forward(sd, x, [2, 2, 1])
    out = m.cuda()(x.cuda()).cpu()
    assert out.shape == ref.shape == (3, 44, 3)
    assert float((out - ref).abs().max()) < 2e-3
    assert float((torch.sigmoid(out) - torch.sigmoid(ref)).abs().max()) < 1e-3


def test_cnn_long_clip_and_workspace_reuse_across_shapes():
    m, sd = refmodels.seeded_cnn(refmodels.MAIN_CFG, seed=4, bn_seed=5)
    m = m.cuda()
    shapes = [(2, 30), (1, 1819), (3, 61), (2, 30)]            # 1819 frames = a 10-minute clip
    for (b, t) in shapes:
        x = refmodels.cnn_inputs(t, 1000 + t + b, batch=b)
        with torch.no_grad():
            ref = torch.sigmoid(cnn_ref.cnn_avgpooling_forward(sd, x, [2, 2, 2, 1])).numpy()
        p = m.logits(x.cuda()).cpu().numpy()
        assert p.shape == ref.shape
        assert np.abs(p - ref).max() < 1e-3, (b, t)


def test_m5_multiclass():
    torch.manual_seed(3)
    m = M5(4)
    sd = _randomised(m, 9)
    x = refmodels.m5_inputs(5, seed=2)
    with torch.no_grad():
        ref = cnn_ref.m5_forward(sd, x)
    out = m.cuda()(x.cuda()).cpu()
    assert out.shape == ref.shape == (5, 4)
    assert float((out - ref).abs().max()) < 2e-3


def test_logmel_unaligned_rows_and_stream():
    n = 100001                                                  # odd stride: every second clip starts unaligned
    ys = np.stack([signals.hdr(n, 60 + i) for i in range(3)])
    w = torch.from_numpy(ys).float().cuda()
    ref = R.waveform_to_log_mel(ys)
    out = P.waveform_to_log_mel(w).cpu().numpy()
    assert np.abs(out - ref).max() < 1e-2
    view = torch.zeros(3 * n + 3, device="cuda")[1:1 + 3 * n].view(3, n)     # base pointer off by 4 bytes
    view.copy_(w)
    assert np.abs(P.waveform_to_log_mel(view).cpu().numpy() - ref).max() < 1e-2
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out_s = P.waveform_to_log_mel(w)
    s.synchronize()
    assert torch.equal(out_s.cpu(), torch.from_numpy(out))
