"""Where does the batch-independent ~1.1 ms of a Cnn_AvgPooling forward go?

Times sedb_cnn_forward at B in {1, 16, 128, 256} three ways: (a) the raw C-ABI call in a loop (CUDA events = device
time, perf_counter = host time to enqueue), (b) through the drop-in nn.Module, (c) M5 likewise at 128 / 1024 frames.
Writes gpurun_out/fixed_cost.json.
"""
import ctypes, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200  # noqa
from sed_b200 import _ext
from sed_b200.models._native import aligned_ptr
import refmodels

lib = _ext.load()
res = {}
m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
m = m.cuda().eval()
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
N = 50


def timed(fn, n=N):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    w0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    w_enq = (time.perf_counter() - w0) * 1e3 / n
    torch.cuda.synchronize()
    w_all = (time.perf_counter() - w0) * 1e3 / n
    return {"event_ms": a.elapsed_time(b) / n, "host_enqueue_ms": w_enq, "wall_ms": w_all}


for B in (1, 16, 128, 256):
    x = torch.randn(B, 1, 182, 64, device="cuda")
    with torch.no_grad():
        m.logits(x)
    h = m._native.get(x.device, m._native_tensors())
    need = lib.sedb_cnn_workspace_bytes(h, B, 182)
    ws = torch.zeros(need + 256, dtype=torch.uint8, device="cuda")
    wp, wb = aligned_ptr(ws)
    out = torch.empty(B, 176, 1, device="cuda")
    st = _ext.stream_ptr()
    xp, op = ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr())

    def raw():
        _ext.check(lib.sedb_cnn_forward(h, xp, B, 182, None, op, wp, wb, st))

    def mod():
        with torch.no_grad():
            m.logits(x)

    res[f"cnn_B{B}_cabi"] = timed(raw)
    res[f"cnn_B{B}_module"] = timed(mod)
    print(B, res[f"cnn_B{B}_cabi"], res[f"cnn_B{B}_module"], flush=True)

m5, _ = refmodels.seeded_m5()
m5 = m5.cuda().eval()
for B in (16, 128, 1024):
    x = torch.randn(B, 1, 31680, device="cuda") * 0.1

    def mod5():
        with torch.no_grad():
            m5(x)

    res[f"m5_B{B}_module"] = timed(mod5, 20)
    print("m5", B, res[f"m5_B{B}_module"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "fixed_cost.json"), "w"), indent=1)
