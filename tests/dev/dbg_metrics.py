import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sed_b200, signals, refmodels
from sed_b200.dataset.spectogram import preprocess as P
from oracle import logmel_ref, cnn_ref
m, sd = refmodels.seeded_cnn(refmodels.MAIN_CFG); m = m.cuda()
with torch.no_grad(): m.event_fc.weight.mul_(25.0)
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
ys = np.stack([signals.hdr(2880000, 30 + i) for i in range(2)])
lm = logmel_ref.waveform_to_log_mel(ys)
mean, std = lm.mean((0, 1)), lm.std((0, 1))
x_ref = torch.from_numpy(((lm - mean) / std)[:, None].astype(np.float32))
with torch.no_grad():
    p_ref = torch.sigmoid(cnn_ref.cnn_avgpooling_forward(sd, x_ref, [2, 2, 2, 1])).numpy()
x_gpu = P.waveform_to_log_mel(torch.from_numpy(ys).float().cuda(), mean=mean, std=std)[:, None]
print("logmel (normalised) max diff", float((x_gpu.cpu() - x_ref).abs().max()))
p_gg = m.logits(x_gpu).cpu().numpy()
p_og = m.logits(x_ref.cuda()).cpu().numpy()
with torch.no_grad():
    p_gc = torch.sigmoid(cnn_ref.cnn_avgpooling_forward(sd, x_gpu.cpu(), [2, 2, 2, 1])).numpy()
print(os.environ.get("SEDB_LIB_PATH", "new")[-14:], "gpu->gpu", np.abs(p_gg - p_ref).max(), "oracle->gpuCNN", np.abs(p_og - p_ref).max(), "gpuLM->cpuCNN", np.abs(p_gc - p_ref).max())
