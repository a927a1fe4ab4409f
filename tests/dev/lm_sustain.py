import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sed_b200
from sed_b200.dataset.spectogram import preprocess as P
C = 256
g = torch.Generator(device="cuda").manual_seed(1)
wave = torch.empty(C, 2880000, device="cuda")
for i in range(C): wave[i] = (torch.randn(2880000, device="cuda", generator=g) * 0.1).clamp_(-1, 1)
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_slowdown", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for _ in range(3): P.waveform_to_log_mel(wave)
torch.cuda.synchronize()
print("idle", smi())
for rep in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): P.waveform_to_log_mel(wave)
    b.record()
    s = smi()
    torch.cuda.synchronize()
    print(rep, round(a.elapsed_time(b) / 10, 4), s)
    if rep == 7: time.sleep(3)
