"""Generates the golden fixtures in this directory.  Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

* ``cnn_reference.npz`` / ``m5_reference.npz``: outputs of the VERBATIM reference modules
  (models/spectogram_models.py, models/waveform_models.py; matplotlib stubbed because utils/common.py:3 imports it)
  for seeded weights and inputs.  The weights are not stored: ``torch.manual_seed(seed)`` + module construction is
  reproducible, and this script asserts that the drop-in modules of this repo draw identical initial weights.
* ``metrics_reference.npz``: utils/metric_utils.calculate_metrics on seeded probabilities/targets.
* ``logmel_oracle.npz``: log-mel of seeded signals from oracle/logmel_ref.py (librosa itself is not installable
  here, so these vectors pin the oracle against regressions; the oracle is pinned against torch.stft/torchaudio in
  tests/test_oracle_logmel.py).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import cnn_ref, logmel_ref, metrics_ref  # noqa: E402
import signals  # noqa: E402

REF = "/root/reference"
MAIN_CFG = [(32, 2), (64, 2), (128, 2), (128, 1)]          # main.py:35


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        m = types.ModuleType(name)
        m.use = lambda *a, **k: None
        sys.modules.setdefault(name, m)
    sys.path.insert(0, REF)
    # our own tests/ and repo root also hold packages called `models`/`utils`/`dataset`? no: ours live under sed_b200
    from models.spectogram_models import Cnn_AvgPooling
    from models.waveform_models import M5
    from utils.metric_utils import calculate_metrics, f_score
    return Cnn_AvgPooling, M5, calculate_metrics, f_score


def cnn_inputs(T, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(2, 1, T, 64, generator=g) * 1.5


def main():
    RefCnn, RefM5, ref_metrics, ref_fscore = import_reference()
    import sed_b200  # noqa: F401
    from sed_b200.models.spectogram_models import Cnn_AvgPooling
    from sed_b200.models.waveform_models import M5
    torch.set_num_threads(4)

    # ---------------- Cnn_AvgPooling
    out = {}
    for cfg_name, cfg in (("main", MAIN_CFG), ("default", None)):
        torch.manual_seed(0)
        ref = RefCnn(1, model_config=cfg) if cfg else RefCnn(1)
        torch.manual_seed(0)
        ours = Cnn_AvgPooling(1, model_config=cfg) if cfg else Cnn_AvgPooling(1)
        sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
        assert list(sd_ref.keys()) == list(sd_ours.keys()), "state_dict keys differ"
        for k in sd_ref:
            assert torch.equal(sd_ref[k], sd_ours[k]), f"seeded init differs at {k}"
        sd = cnn_ref.randomize_bn_({k: v.clone() for k, v in sd_ref.items()}, seed=7)
        ref.load_state_dict(sd)
        ref.eval()
        pools = [p for _, p in (cfg or [(64, 2), (128, 2), (256, 2), (512, 1)])]
        for T in ((30, 181, 182, 183, 184) if cfg_name == "main" else (30, 182)):
            x = cnn_inputs(T, 100 + T)
            with torch.no_grad():
                y_ref = ref(x)
                p_ref = ref.logits(x)
                y_or = cnn_ref.cnn_avgpooling_forward(sd, x, pools)
            assert torch.allclose(y_ref, y_or, atol=1e-5, rtol=1e-5), (cfg_name, T, (y_ref - y_or).abs().max())
            out[f"{cfg_name}_T{T}_logits"] = y_ref.numpy()
            out[f"{cfg_name}_T{T}_probs"] = p_ref.numpy()
            print("cnn", cfg_name, T, tuple(y_ref.shape), float((y_ref - y_or).abs().max()))
    np.savez_compressed(os.path.join(HERE, "cnn_reference.npz"), **out)

    # ---------------- M5
    torch.manual_seed(0)
    ref = RefM5(1)
    torch.manual_seed(0)
    ours = M5(1)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys())
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ours[k]), f"seeded init differs at {k}"
    import refmodels
    sd = refmodels.m5_test_weights(sd_ref, 11)
    ref.load_state_dict(sd)
    ref.eval()
    x = refmodels.m5_inputs(10)
    with torch.no_grad():
        y_ref = ref(x)
        y_or = cnn_ref.m5_forward(sd, x)
    assert torch.allclose(y_ref, y_or, atol=1e-5, rtol=1e-5), (y_ref - y_or).abs().max()
    print("m5", tuple(y_ref.shape), float((y_ref - y_or).abs().max()))
    np.savez_compressed(os.path.join(HERE, "m5_reference.npz"), logits=y_ref.numpy())

    # ---------------- metrics
    rng = np.random.default_rng(3)
    mo = {}
    for i in range(4):
        probs = rng.random((176, 1)).astype(np.float32)
        tgt = metrics_ref.create_event_matrix(182, [5.0 + 7 * i, 30.0], [5.66 + 7 * i, 30.66])
        if i == 3:
            tgt[:] = 0
        r, p, ap = ref_metrics(probs, tgt)
        r2, p2, ap2 = metrics_ref.calculate_metrics(probs, tgt)
        assert np.array_equal(r, r2) and np.array_equal(p, p2) and ap == ap2
        mo[f"probs{i}"], mo[f"target{i}"], mo[f"recall{i}"], mo[f"precision{i}"], mo[f"ap{i}"] = probs, tgt, r, p, ap
        mo[f"f1_{i}"] = ref_fscore(r, p)
    np.savez_compressed(os.path.join(HERE, "metrics_reference.npz"), **mo)
    print("metrics ok")

    # ---------------- log-mel (oracle vectors)
    lm = {}
    for name, fn in signals.ALL.items():
        lm[f"{name}_100000"] = logmel_ref.waveform_to_log_mel(fn(100000, 3))
    lm["tone1k_48000"] = logmel_ref.waveform_to_log_mel(signals.tone(48000))
    lm["impulse0_31680"] = logmel_ref.waveform_to_log_mel(signals.impulse(31680, 0))
    np.savez_compressed(os.path.join(HERE, "logmel_oracle.npz"), **lm)
    print("logmel ok", {k: v.shape for k, v in lm.items()})


if __name__ == "__main__":
    main()
