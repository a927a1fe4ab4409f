import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sed_b200, refmodels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG); m = m.cuda().eval()
x = torch.randn(B, 1, 182, 64, device="cuda")
with torch.no_grad():
    for _ in range(3): m.logits(x)
torch.cuda.synchronize()
