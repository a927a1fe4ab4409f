import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200, refmodels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m, _ = refmodels.seeded_m5(); m = m.cuda()
x = torch.randn(B, 1, 31680, device="cuda") * 0.1
for _ in range(3): m(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): m(x)
e1.record(); torch.cuda.synchronize()
print(f"m5 B={B}: {e0.elapsed_time(e1)/5:.3f} ms")
