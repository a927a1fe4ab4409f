"""Development driver for the first GPU runs: tcgen05 probe matrix, log-mel parity, rough timing."""
import ctypes, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200 import _ext
from sed_b200.dataset.spectogram import preprocess as P
from oracle import logmel_ref as R
import signals

lib = _ext.load()
res = {}
print(torch.cuda.get_device_name(0), flush=True)

def probe(N, K, a_major, b_major, pad, neg_b, swap):
    g = torch.Generator(device="cpu").manual_seed(N * 7 + K)
    a = torch.randn(128, K, generator=g).cuda(); b = torch.randn(K, N, generator=g).cuda()
    d = torch.zeros(128, N, device="cuda")
    rc = lib.sedb_debug_umma_probe(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(d.data_ptr()),
                                   N, K, a_major, b_major, pad, neg_b, swap, None)
    if rc: return "rc:" + lib.sedb_last_error().decode()
    torch.cuda.synchronize()
    ref = a.bfloat16().float() @ b.bfloat16().float()
    if neg_b: ref = -ref
    return float((d - ref).abs().max() / ref.abs().max())

if "--skip-probe" not in sys.argv:
    for swap in (0, 1):
        for (am, bm) in ((0, 0), (0, 1), (1, 0), (1, 1)):
            for (N, K, pad, neg) in ((128, 16, 0, 0), (128, 64, 0, 0), (256, 64, 0, 0), (128, 32, 16, 0), (128, 32, 0, 1)):
                key = f"swap{swap}_a{am}_b{bm}_N{N}_K{K}_pad{pad}_neg{neg}"
                try:
                    res[key] = probe(N, K, am, bm, pad, neg, swap)
                except Exception as e:
                    res[key] = "EXC " + repr(e)
                print(key, res[key], flush=True)

if "--probe-only" in sys.argv:
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1) if os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True) is None else None
    sys.exit(0)
# ---- logmel parity
def logmel_err(name, y):
    ref = R.waveform_to_log_mel(y)
    out = P.waveform_to_log_mel(torch.from_numpy(y).float().cuda()).cpu().numpy()
    err = np.abs(out - ref)
    return float(err.max()), int(np.argmax(err.max(axis=1))), float(ref.min()), float(ref.max())

n = 480000
for name, fn in signals.ALL.items():
    y = fn(n, 0)
    try:
        r = logmel_err(name, y)
    except Exception as e:
        r = "EXC " + repr(e)
    res["logmel_" + name] = r
    print("logmel", name, r, flush=True)
for nn in (31680, 2880001):
    y = signals.hdr(nn, 1)
    try:
        r = logmel_err("hdr", y)
    except Exception as e:
        r = "EXC " + repr(e)
    res[f"logmel_hdr_{nn}"] = r
    print("logmel hdr", nn, r, flush=True)

# stft parity
y = signals.hdr(100000, 2)
try:
    S = P.multichannel_stft(y[:, None]); Sr = R.multichannel_stft(y[:, None])
    res["stft_rel"] = float(np.abs(S - Sr).max() / np.abs(Sr).max())
    lm = P.multichannel_complex_to_log_mel(Sr); lmr = R.multichannel_complex_to_log_mel(Sr)
    res["c2lm"] = float(np.abs(lm - lmr).max())
except Exception as e:
    res["stft_rel"] = "EXC " + repr(e)
print("stft", res.get("stft_rel"), res.get("c2lm"), flush=True)

# ---- timing
try:
    for B in (16, 64, 256):
        w = (torch.randn(B, 2880000, device="cuda") * 0.1).clamp_(-1, 1)
        for _ in range(2): P.waveform_to_log_mel(w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): P.waveform_to_log_mel(w)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[f"time_B{B}_ms"] = ms
        print(f"B={B}: {ms:.3f} ms  -> {B*60/3600/(ms*1e-3):.1f} audio-h/s, {B*182*75.5e6*4/3/ms/1e9:.1f} TFLOP/s(mma)", flush=True)
        del w
except Exception as e:
    res["time"] = "EXC " + repr(e)
    print(res["time"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "first_gpu.json"), "w"), indent=1)
