"""Compare the native training step's intermediates (Z, A, dZ per layer) with torch autograd in float64."""
import copy, ctypes, os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200 import _ext
from sed_b200.models.spectogram_models import Cnn_AvgPooling
from sed_b200.models._native import aligned_ptr
from sed_b200.utils.common import WeightedBCE
from oracle import cnn_ref
lib = _ext.load()
crit = WeightedBCE(recall_factor=5, multi_frame=True)
cfg = eval(sys.argv[1]) if len(sys.argv) > 1 else [(32, 1)]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = int(sys.argv[3]) if len(sys.argv) > 3 else 30
torch.manual_seed(0)
m = Cnn_AvgPooling(1, model_config=cfg)
m.load_state_dict(cnn_ref.randomize_bn_({k: v.clone() for k, v in m.state_dict().items()}, seed=7))
m = m.cuda().train()
ref = copy.deepcopy(m).double(); ref.native_training = False; ref.train()
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(B, 1, T, 64, device="cuda", generator=g) * 1.5
y = (torch.rand(B, T, 1, device="cuda", generator=g) > 0.8).float()
out = m(x); loss = crit(out, y); loss.backward()
# reference intermediates
zs, acts = [], []
h = x.double()
for blk in ref.conv_blocks:
    z1 = blk.conv1(h); z1.retain_grad(); a1 = F.relu(blk.bn1(z1)); a1.retain_grad()
    z2 = blk.conv2(a1); z2.retain_grad(); a2 = F.avg_pool2d(F.relu(blk.bn2(z2)), blk.pool_size); a2.retain_grad()
    zs += [z1, z2]; acts += [a1, a2]; h = a2
o = torch.mean(h, dim=3).transpose(1, 2)
from sed_b200.models.spectogram_models import interpolate
o = interpolate(ref.event_fc(o), 2 ** ref.num_pools)
crit(o, y.double()).backward()
hd = m._train_handle(x.device)
lay = (ctypes.c_longlong * 256)()
_ext.check(lib.sedb_debug_train_layout(hd, B, T, lay, 256))
n = lay[0]
ws = m._native.workspace(x.device, ("train", B, T), 0)
ptr, _ = aligned_ptr(ws)
off0 = ptr.value - ws.data_ptr()
raw = ws[off0:]
def planes_f32(off, C, H, W, S):
    t = raw[off:off + B * (C // 8) * S * 32].view(torch.float32).view(B, C // 8, S, 8)[:, :, 8:8 + (H + 2) * (W + 2)]
    t = t.reshape(B, C // 8, H + 2, W + 2, 8)[:, :, 1:H + 1, 1:W + 1]
    return t.permute(0, 1, 4, 2, 3).reshape(B, C, H, W).double()
def planes_bf(off, C, H, W, S):
    t = raw[off:off + B * 2 * (C // 8) * S * 16].view(torch.bfloat16).view(B, 2, C // 8, S, 8)[:, :, :, 8:8 + (H + 2) * (W + 2)]
    full = t.reshape(B, 2, C // 8, H + 2, W + 2, 8).double()
    full = full[:, 0] + full[:, 1]
    inner = full[:, :, 1:H + 1, 1:W + 1].permute(0, 1, 4, 2, 3).reshape(B, C, H, W)
    border = full.abs().sum() - full[:, :, 1:H + 1, 1:W + 1].abs().sum()
    return inner, float(border)
for l in range(n):
    C, H, W, pool, zo, zS, ao, aS, dzo, dzS, so, wgi = [lay[4 + 12 * l + i] for i in range(12)]
    Z = planes_f32(zo, C, H, W, zS)
    A, ab = planes_bf(ao, C, H // pool, W // pool, aS)
    ez = ((Z - zs[l]).norm() / zs[l].norm()).item()
    ea = ((A - acts[l]).norm() / acts[l].norm()).item()
    msg = f"layer {l}: C {C} {H}x{W} pool {pool} wg(n_pc*1000+Pb) {wgi}: Z rel {ez:.2e}  A rel {ea:.2e} (padding abs sum {ab:.1e})"
    if l >= 1:
        dZ, db = planes_bf(dzo, C, H, W, dzS)
        ed = ((dZ - zs[l].grad).norm() / zs[l].grad.norm()).item()
        # row-wise error profile
        rows = ((dZ - zs[l].grad) ** 2).sum((0, 1, 3)).sqrt() / (zs[l].grad ** 2).sum((0, 1, 3)).sqrt().clamp_min(1e-30)
        msg += f"  dZ rel {ed:.2e} (padding {db:.1e}) worst row {int(rows.argmax())} {rows.max().item():.2e}"
    print(msg)
for (nme, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
    d = (p.grad.double() - q.grad)
    print(f"    {nme:34s} rel {(d.norm() / q.grad.norm()).item():.2e}")
# ---- BN backward sums of the last layer and the final content of G (= dL/dA_0 for a one-block model)
st = raw[:8 * 4096].view(torch.float64)
l = n - 1
C, H, W, pool, zo, zS, ao, aS, dzo, dzS, so, wgi = [lay[4 + 12 * l + i] for i in range(12)]
s1 = st[so + 2 * C: so + 3 * C]; s2 = st[so + 3 * C: so + 4 * C]
zl = zs[l]; bn = ref.conv_blocks[l // 2].bn2 if l % 2 else ref.conv_blocks[l // 2].bn1
yl = F.batch_norm(zl, None, None, bn.weight, bn.bias, True, 0.0, 1e-5)
xh = (zl - zl.mean((0, 2, 3), keepdim=True)) / torch.sqrt(zl.var((0, 2, 3), unbiased=False, keepdim=True) + 1e-5)
# gradient wrt the (un-pooled) relu output
ga = torch.autograd.grad(crit(interpolate(ref.event_fc(torch.mean(F.avg_pool2d(F.relu(yl), pool), dim=3).transpose(1, 2)), 2 ** ref.num_pools), y.double()), yl)[0]
print("s1 rel", ((s1 - ga.sum((0, 2, 3))).norm() / ga.sum((0, 2, 3)).norm()).item(), "s2 rel", ((s2 - (ga * xh).sum((0, 2, 3))).norm() / (ga * xh).sum((0, 2, 3)).norm()).item())
print("s1 ours", s1[:4].tolist(), "ref", ga.sum((0, 2, 3))[:4].tolist())
if n == 2:
    C0, H0, W0 = lay[4], lay[5], lay[6]
    S_g = ((8 + (H0 + 2) * (W0 + 2) + 8 + 7) // 8) * 8
    G = planes_f32(lay[1], C0, H0, W0, S_g)
    print("G(dL/dA_0) rel", ((G - acts[0].grad).norm() / acts[0].grad.norm()).item())
r1 = ga.sum((0, 2, 3))
print("per-channel s1 rel err", ((s1 - r1).abs() / r1.abs()).tolist()[:32])
dZ, _ = planes_bf(dzo, C, H, W, dzS)
err = (dZ - zs[l].grad)
print("dZ err by row", (err ** 2).sum((0, 1, 3)).sqrt().tolist())
print("dZ ref by row", (zs[l].grad ** 2).sum((0, 1, 3)).sqrt().tolist())
print("dZ err by image", (err ** 2).sum((1, 2, 3)).sqrt().tolist())
print("dZ err by col", (err ** 2).sum((0, 1, 2)).sqrt().tolist()[:12])
print("dZ err by channel", (err ** 2).sum((0, 2, 3)).sqrt().tolist())
