import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sed_b200, refmodels
from sed_b200.models.spectogram_models import Cnn_AvgPooling
from sed_b200.train import DataParallelTrainer
from sed_b200.utils.common import WeightedBCE
torch.manual_seed(0)
m = Cnn_AvgPooling(1, model_config=refmodels.MAIN_CFG).cuda()
tr = DataParallelTrainer(m, WeightedBCE(recall_factor=5, multi_frame=True), lr=1e-6)
x = torch.randn(64, 1, 30, 64, device="cuda"); y = (torch.rand(64, 30, 1, device="cuda") > 0.8).float()
for _ in range(3): tr.step(x, y)
torch.cuda.synchronize()
