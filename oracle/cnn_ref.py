"""CPU restatement of the reference's CNN forward passes (TEST INFRASTRUCTURE ONLY).

Functional float32 PyTorch-CPU restatements, driven by a ``state_dict`` with the reference's own key names:

* :func:`cnn_avgpooling_forward`  <- ``Cnn_AvgPooling.forward`` / ``ConvBlock.forward`` / ``interpolate``
  (models/spectogram_models.py:9-22, 153-160, 185-202)
* :func:`m5_forward`              <- ``M5.forward`` (models/waveform_models.py:59-71)
* :func:`weighted_bce`            <- ``WeightedBCE.__call__`` (utils/common.py:16-30)

Pinned against the verbatim reference modules imported from /root/reference by
``tests/golden/make_golden.py`` (fixtures in ``tests/golden/``) and, when the reference tree is present, directly by
``tests/test_oracle_cnn.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=BN_EPS)


def cnn_num_blocks(sd):
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("conv_blocks."))


def cnn_avgpooling_forward(sd, x, pools):
    """x: (B, 1, T, 64) float32 -> logits (B, T', classes); ``pools`` = pool size per block (model_config[i][1])."""
    n_blocks = cnn_num_blocks(sd)
    assert len(pools) == n_blocks
    x = x.to(torch.float32)
    for b in range(n_blocks):
        pre = f"conv_blocks.{b}"
        x = F.relu(_bn(F.conv2d(x, sd[pre + ".conv1.weight"], None, stride=1, padding=1), sd, pre + ".bn1"))
        x = F.relu(_bn(F.conv2d(x, sd[pre + ".conv2.weight"], None, stride=1, padding=1), sd, pre + ".bn2"))
        x = F.avg_pool2d(x, kernel_size=pools[b])
    x = torch.mean(x, dim=3).transpose(1, 2)
    y = F.linear(x, sd["event_fc.weight"], sd["event_fc.bias"])
    num_pools = 1 + sum(1 for p in pools[1:] if p == 2)          # spectogram_models.py:167-172
    ratio = 2 ** num_pools
    B, T, C = y.shape
    return y[:, :, None, :].repeat(1, 1, ratio, 1).reshape(B, T * ratio, C)


_M5_LAYOUT = [  # (block, conv index, bn index, stride, padding); MaxPool(4) follows blocks 1-4
    (1, 0, 1, 4, 39),
    (2, 0, 1, 1, 1), (2, 3, 4, 1, 1),
    (3, 0, 1, 1, 1), (3, 3, 4, 1, 1),
    (4, 0, 1, 1, 1), (4, 3, 4, 1, 1),
    (5, 0, 1, 1, 1), (5, 3, 4, 1, 1),
]


def m5_forward(sd, x):
    """x: (B, 1, 31680) float32 -> logits (B, classes)."""
    x = x.to(torch.float32)
    for i, (blk, ci, bi, stride, pad) in enumerate(_M5_LAYOUT):
        pre = f"conv_block{blk}"
        x = F.conv1d(x, sd[f"{pre}.{ci}.weight"], sd[f"{pre}.{ci}.bias"], stride=stride, padding=pad)
        x = F.relu(_bn(x, sd, f"{pre}.{bi}"))
        last_of_block = (i + 1 == len(_M5_LAYOUT)) or (_M5_LAYOUT[i + 1][0] != blk)
        if last_of_block and blk <= 4:
            x = F.max_pool1d(x, 4, 4)
    return F.linear(torch.mean(x, dim=2), sd["fc.weight"], sd["fc.bias"])


def weighted_bce(output, target, recall_factor, multi_frame=True):
    """utils/common.py:16-30."""
    if multi_frame:
        n = min(output.shape[1], target.shape[1])
        o, t = output[:, :n], target[:, :n]
    else:
        o, t = output.reshape(-1), target
    return F.binary_cross_entropy_with_logits(o, t, pos_weight=torch.tensor([float(recall_factor)]))


def randomize_bn_(sd, seed=0):
    """Randomise BatchNorm affine/running stats in place (defaults mu=0, var=1 would hide BN-fold bugs)."""
    g = torch.Generator().manual_seed(seed)
    for k in list(sd.keys()):
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.3
        elif k.endswith("running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 1.5 + 0.5
        elif (".bn" in k or _is_bn1d_key(k, sd)) and k.endswith("weight") and sd[k].dim() == 1:
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
        elif (".bn" in k or _is_bn1d_key(k, sd)) and k.endswith("bias") and sd[k].dim() == 1 and \
                k.replace("bias", "running_mean") in sd:
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.2
    return sd


def _is_bn1d_key(k, sd):
    return k.rsplit(".", 1)[0] + ".running_mean" in sd
