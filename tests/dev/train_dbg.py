"""Localise native-training errors: small model configs / shapes against float64 autograd."""
import copy, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200.models.spectogram_models import Cnn_AvgPooling
from sed_b200.utils.common import WeightedBCE
from oracle import cnn_ref
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
crit = WeightedBCE(recall_factor=5, multi_frame=True)
cases = [([(32, 1)], 4, 8), ([(32, 1)], 4, 30), ([(32, 2)], 4, 30), ([(32, 1), (64, 1)], 4, 16), ([(32, 2), (64, 2)], 4, 30),
         ([(32, 2), (64, 2), (128, 2), (128, 1)], 4, 30), ([(32, 2), (64, 2), (128, 2), (128, 1)], 64, 30), ([(128, 1)], 8, 8)]
for cfg, B, T in cases:
    torch.manual_seed(0)
    m = Cnn_AvgPooling(1, model_config=cfg)
    m.load_state_dict(cnn_ref.randomize_bn_({k: v.clone() for k, v in m.state_dict().items()}, seed=7))
    m = m.cuda().train()
    ref = copy.deepcopy(m).double(); ref.native_training = False; ref.train()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, 1, T, 64, device="cuda", generator=g) * 1.5
    y = (torch.rand(B, T, 1, device="cuda", generator=g) > 0.8).float()
    out = m(x); loss = crit(out, y); loss.backward()
    out_r = ref(x.double()); loss_r = crit(out_r, y.double()); loss_r.backward()
    print(f"cfg {cfg} B {B} T {T}: logits err {(out.double()-out_r).abs().max().item():.2e} (max {out_r.abs().max().item():.2f}) "
          f"loss rel {abs(loss.item()-loss_r.item())/abs(loss_r.item()):.2e}")
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        rel = ((p.grad.double() - q.grad).norm() / q.grad.norm()).item()
        print(f"    {n:34s} rel {rel:.2e}  norm {q.grad.norm().item():.3e}")
