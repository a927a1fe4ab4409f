"""Host-side logic added in round 2, no GPU needed: the conv planner (decomposition per layer and batch size), the
training-golden fixture's self-consistency, the bench helpers that tie ncu captures to kernel sources, and the native
null-argument paths of the training entry points."""
import ctypes
import json
import os

import numpy as np
import pytest

import sed_b200
from sed_b200 import _ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAIN_LAYERS = [(32, 32, 2, 182, 64), (32, 64, 1, 91, 32), (64, 64, 2, 91, 32), (64, 128, 1, 45, 16), (128, 128, 2, 45, 16),
               (128, 128, 1, 22, 8), (128, 128, 1, 22, 8)]           # (cin, cout, pool, H, W) of main.py:35's model at T = 182


def plan(cin, cout, pool, H, W, amode, n_img, mode=0, ntaps=9, sms=148):
    o = (ctypes.c_int * 8)()
    rc = _ext.load().sedb_debug_plan_layer(cin, cout, pool, mode, ntaps, H, W, amode, n_img, sms, o)
    assert rc == 0, _ext.load().sedb_last_error()
    return dict(tiles=o[0], nsub=o[1], fuse=o[2], bands=o[3], R=o[4], smem=o[5], S=o[6], kpb=o[7] // 100)


@pytest.mark.parametrize("amode", [0, 1])
@pytest.mark.parametrize("n_img", [1, 16, 64, 256])
def test_plans_fit_the_hardware(amode, n_img):
    for (cin, cout, pool, H, W) in MAIN_LAYERS:
        p = plan(cin, cout, 1 if amode else pool, H, W, amode, n_img)
        assert p["smem"] <= 227 * 1024
        assert 1 <= p["tiles"] <= 4 and p["nsub"] in (1, 2, 4)
        cols = (2 * cout if p["fuse"] else cout // p["nsub"]) * p["tiles"]
        assert cols <= 256                                            # accumulators are double buffered in 512 TMEM columns
        assert p["R"] * (W + 2) <= 128 * p["tiles"] and p["bands"] * p["R"] >= H
        if pool == 2 and not amode:
            assert p["R"] % 2 == 0                                    # pooling windows never straddle bands
        assert p["S"] >= 8 + (H + 2) * (W + 2)                        # the input planes hold the padded image behind the lead
        assert p["kpb"] in (1, 3, 9)


def test_fusing_is_a_property_of_the_layer_not_of_the_batch():
    """a clip's result must not depend on the batch it arrives in: [wH | wL] fusion changes the rounding"""
    for (cin, cout, pool, H, W) in MAIN_LAYERS:
        fuse = {plan(cin, cout, pool, H, W, 0, n)["fuse"] for n in (1, 3, 16, 70, 128, 256)}
        assert len(fuse) == 1 and fuse == {1 if cout <= 64 else 0}


def test_small_batches_get_smaller_items():
    big = plan(128, 128, 1, 22, 8, 0, 256)
    small = plan(128, 128, 1, 22, 8, 0, 16)
    assert small["tiles"] * small["bands"] >= big["tiles"] * big["bands"] or small["tiles"] < big["tiles"]
    assert small["tiles"] <= big["tiles"]


def test_one_dimensional_layers_plan():
    for (cin, cout, pool, L) in ((64, 64, 1, 1980), (64, 64, 4, 1980), (128, 256, 1, 30), (256, 256, 1, 30)):
        p = plan(cin, cout, pool, 1, L, 0, 128, mode=1, ntaps=3)
        assert p["smem"] <= 227 * 1024 and p["kpb"] in (1, 3)


def test_unplannable_layer_is_an_error_not_a_crash():
    o = (ctypes.c_int * 8)()
    assert _ext.load().sedb_debug_plan_layer(32, 32, 2, 0, 9, 1, 1000, 0, 4, 148, o) != 0      # a band of one 1002-pixel row does not fit


def test_training_entry_points_reject_null_arguments():
    lib = _ext.load()
    assert lib.sedb_cnn_train_workspace_bytes(None, 4, 30) == 0
    assert lib.sedb_cnn_train_forward(None, None, 0, None, 4, 30, 0.1, None, None, 0, None) != 0
    assert lib.sedb_cnn_train_backward(None, None, 0, None, None, 4, 30, None, 0, None, 0, None) != 0
    assert lib.sedb_bce_with_logits(None, None, 4, 24, 30, 1, 5.0, 1.0, None, None, None) != 0
    assert lib.sedb_adam_amsgrad_step_dev(None, None, None, None, None, 10, None, None, 0.9, 0.999, 1e-8, 0.0, 1.0, None) != 0
    assert lib.sedb_cnn_workspace_invalidate(None, None) != 0 and lib.sedb_m5_workspace_invalidate(None, None) != 0


def test_training_golden_fixture_is_self_consistent():
    g = np.load(os.path.join(ROOT, "tests", "golden", "train_reference.npz"))
    names = list(g["param_names"])
    assert len(names) == 26 and names[0] == "conv_blocks.0.conv1.weight" and names[-1] == "event_fc.bias"
    for case in ("B2_T16", "B4_T8", "B3_T13"):
        assert float(g[f"{case}_relu_margin"]) >= 2e-5                 # tie-free by construction (make_train_golden.py)
        B, T, Tt, _ = (int(v) for v in g[f"{case}_shape"])
        assert g[f"{case}_logits"].shape == (B, 8 * (T // 8), 1)
        for i in range(26):
            assert float(g[f"{case}_grad_{i}_ref32_relerr"]) < 1e-4    # the reference's own float32 run agrees with float64
            assert g[f"{case}_grad_{i}_idx"].shape == g[f"{case}_grad_{i}_val"].shape
    # the loose cases exist and keep the reference's float32 noise beside them
    assert "B4_T30_grad_0_ref32_relerr" in g and "B2_T182_loss" in g


def test_bench_ties_captures_to_kernel_sources():
    import bench
    sha = bench.kernel_source_sha()
    assert len(sha) == 16 and sha == bench.kernel_source_sha()
    prof = bench.profiled_logmel(256)
    assert prof is not None and prof["capture"].startswith("profiles/r") and prof["traffic"] > 2.9e9
    d = json.load(open(os.path.join(ROOT, prof["capture"])))
    assert prof["stale"] == (d.get("source_sha") != sha)
    assert bench.profiled_logmel(7) is None                            # no capture at that clip count
