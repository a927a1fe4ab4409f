"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sed_b200
from sed_b200.dataset.spectogram import preprocess as P
import refmodels, signals
ys = np.stack([signals.hdr(100000, i) for i in range(3)])
x = P.waveform_to_log_mel(torch.from_numpy(ys).float().cuda())
pcm = torch.from_numpy(np.clip(np.round(ys[:, :, None].repeat(2, axis=2) * 20000), -32768, 32767).astype(np.int16))
x16 = P.pcm16_to_log_mel(pcm.cuda())                                   # 2-channel PCM, vector path
x16b = P.pcm16_to_log_mel(pcm[:, :, :1].repeat(1, 1, 3).contiguous().cuda())   # 3 channels, scalar path
m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG); m = m.cuda()
p = m.logits(torch.randn(2, 1, 37, 64, device="cuda"))
w, _ = refmodels.seeded_m5(); w = w.cuda()
q = w(refmodels.m5_inputs(2).cuda())
torch.cuda.synchronize()
print("ok", x.shape, x16.shape, x16b.shape, p.shape, q.shape)
