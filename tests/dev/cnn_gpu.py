"""Development driver: CNN / M5 parity vs oracle + golden, rough timing."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from oracle import cnn_ref
import refmodels
res = {}
gold = np.load(os.path.join(ROOT, "tests/golden/cnn_reference.npz"))
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "cnn"):
    for cfg_name, cfg in (("main", refmodels.MAIN_CFG), ("default", refmodels.DEFAULT_CFG)):
        m, sd = refmodels.seeded_cnn(cfg)
        m = m.cuda()
        for T in ((30, 181, 182, 183, 184) if cfg_name == "main" else (30, 182)):
            x = refmodels.cnn_inputs(T, 100 + T)
            with torch.no_grad():
                y = m(x.cuda()).cpu().numpy()
                p = m.logits(x.cuda()).cpu().numpy()
            yg = gold[f"{cfg_name}_T{T}_logits"]; pg = gold[f"{cfg_name}_T{T}_probs"]
            r = (float(np.abs(y - yg).max()), float(np.abs(p - pg).max()), float(np.abs(yg).max()))
            res[f"cnn_{cfg_name}_T{T}"] = r
            print("cnn", cfg_name, T, y.shape, r, flush=True)
if what in ("all", "m5"):
    m, sd = refmodels.seeded_m5()
    m = m.cuda()
    x = refmodels.m5_inputs(10)
    with torch.no_grad():
        y = m(x.cuda()).cpu().numpy()
    yg = np.load(os.path.join(ROOT, "tests/golden/m5_reference.npz"))["logits"]
    res["m5"] = (float(np.abs(y - yg).max()), float(np.abs(yg).max()))
    print("m5", res["m5"], y.ravel(), yg.ravel(), flush=True)
if what in ("all", "time"):
    m, sd = refmodels.seeded_cnn(refmodels.MAIN_CFG); m = m.cuda()
    for B in (16, 128, 256):
        x = torch.randn(B, 1, 182, 64, device="cuda")
        with torch.no_grad():
            for _ in range(3): m.logits(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): m.logits(x)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[f"cnn_time_B{B}"] = ms
        print(f"cnn B={B}: {ms:.3f} ms -> {B*60/3600/(ms*1e-3):.0f} audio-h/s, {B*0.9658*3/ms:.1f} TFLOP/s(mma x3)", flush=True)
    m5, _ = refmodels.seeded_m5(); m5 = m5.cuda()
    x = torch.randn(128, 1, 31680, device="cuda") * 0.1
    with torch.no_grad():
        for _ in range(3): m5(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m5(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res["m5_time_B128"] = ms
    print(f"m5 B=128: {ms:.3f} ms", flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"cnn_gpu_{what}.json"), "w"), indent=1)
