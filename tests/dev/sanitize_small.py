"""compute-sanitizer target: small invocations of every kernel family (log-mel, CNN inference, M5, training step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import sed_b200, refmodels, signals
from sed_b200.dataset.spectogram import preprocess as P
from sed_b200.train import DataParallelTrainer
from sed_b200.utils.common import WeightedBCE
y = np.stack([signals.hdr(100000, i) for i in range(2)])
lm = P.waveform_to_log_mel(torch.from_numpy(y).float().cuda())
# more frames than SMs (consecutive runs: block scale taken from the previous frame) with level steps (rejected attempts)
rng = np.random.default_rng(0)
y2 = rng.standard_normal((64, 15840 * 7 + 100)).astype(np.float32) * 0.1
for i in range(0, y2.shape[1], 15840):
    y2[:, i:i + 15840] *= (1e-3, 0.5, 0.05, 1.0, 0.0, 0.2)[(i // 15840) % 6]
lm2 = P.waveform_to_log_mel(torch.from_numpy(np.clip(y2, -1, 1)).cuda())
lm3 = P.pcm16_to_log_mel(torch.from_numpy((np.clip(y2[:40], -1, 1) * 32767).astype(np.int16)).cuda())
# sample-rate conversion: tensor-core path (160 phases, three tiles per clip) and the CUDA-core FIR (3 phases)
from sed_b200.dataset import dataset_utils as DU
rs1 = DU.resample(torch.from_numpy(y2[:3, :44100]).cuda(), 44100, 48000)
rs2 = DU.resample(torch.from_numpy(y2[:3, :16000]).cuda(), 16000, 48000)
m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG); m = m.cuda()
p = m.logits(torch.randn(3, 1, 61, 64, device="cuda"))
m5, _ = refmodels.seeded_m5(); m5 = m5.cuda()
q = m5(torch.randn(2, 1, 31680, device="cuda") * 0.1)
tr = DataParallelTrainer(m, WeightedBCE(recall_factor=5, multi_frame=True), lr=1e-4)
for _ in range(2):
    loss = tr.step(torch.randn(5, 1, 30, 64, device="cuda"), (torch.rand(5, 30, 1, device="cuda") > 0.8).float())
torch.cuda.synchronize()
print("ok", lm.shape, p.shape, q.shape, float(loss))
