import copy, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200.train import DataParallelTrainer
from sed_b200.utils.common import WeightedBCE
import refmodels
torch.manual_seed(0)
a, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
b = copy.deepcopy(a)
a, b = a.cuda(), b.cuda()
crit = WeightedBCE(5, True)
opt = torch.optim.Adam(a.parameters(), lr=1e-3, amsgrad=True)
tr = DataParallelTrainer(b, crit, lr=1e-3)
for it in range(3):
    x = refmodels.cnn_inputs(30, 300 + it, batch=4).cuda()
    y = (torch.rand(4, 30, 1, generator=torch.Generator().manual_seed(it)) > 0.8).float().cuda()
    a.train(); la = crit(a(x), y); opt.zero_grad(); la.backward(); opt.step()
    lb = tr.step(x, y)
    ga = torch.cat([p.grad.reshape(-1) for p in a.parameters()])
    print(it, float(la), float(lb), "grad diff", float((ga - tr.flat.grad).abs().max()), "grad max", float(ga.abs().max()))
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        d = float((p1 - p2).abs().max())
        if d > 1e-5: print("   ", n1, d, float(p1.abs().max()))
