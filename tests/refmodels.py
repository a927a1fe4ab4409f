"""Seeded model weights shared by tests, smoke() and bench.py (same recipe as tests/golden/make_golden.py)."""
import torch

from oracle import cnn_ref

MAIN_CFG = [(32, 2), (64, 2), (128, 2), (128, 1)]            # reference main.py:35
DEFAULT_CFG = [(64, 2), (128, 2), (256, 2), (512, 1)]        # reference spectogram_models.py:7 (infer.py:21)


def seeded_cnn(cfg=MAIN_CFG, classes=1, seed=0, bn_seed=7):
    """Drop-in Cnn_AvgPooling with reference init under `seed` and randomised BN statistics; returns (module, sd)."""
    from sed_b200.models.spectogram_models import Cnn_AvgPooling
    torch.manual_seed(seed)
    m = Cnn_AvgPooling(classes, model_config=cfg)
    sd = cnn_ref.randomize_bn_({k: v.clone() for k, v in m.state_dict().items()}, seed=bn_seed)
    m.load_state_dict(sd)
    m.eval()
    return m, sd


def seeded_m5(classes=1, seed=0, bn_seed=11):
    from sed_b200.models.waveform_models import M5
    torch.manual_seed(seed)
    m = M5(classes)
    sd = m5_test_weights(m.state_dict(), bn_seed)
    m.load_state_dict(sd)
    m.eval()
    return m, sd


def m5_test_weights(sd, bn_seed=11):
    """Randomised BN statistics plus a 40x head gain so that frame logits spread over O(0.1) instead of O(0.005)."""
    sd = cnn_ref.randomize_bn_({k: v.clone() for k, v in sd.items()}, seed=bn_seed)
    sd["fc.weight"] = sd["fc.weight"] * 40.0
    return sd


def cnn_inputs(T, seed, batch=2):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 1, T, 64, generator=g) * 1.5


def m5_inputs(n, seed=5):
    """Waveform frames with per-frame level and spectral content so the logits differ between frames."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(31680, dtype=torch.float32) / 48000.0
    frames = []
    for i in range(n):
        amp = 0.02 * (4.0 ** (i % 5))
        tone = torch.sin(2 * torch.pi * (110.0 * (i + 1)) * t) * (0.5 if i % 2 else 0.0)
        frames.append(amp * (torch.randn(31680, generator=g) + tone))
    return torch.stack(frames)[:, None, :]
