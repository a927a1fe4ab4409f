"""The two helpers of the reference's ``utils/common.py`` that its model/config modules import
(``human_format`` :102-113, ``count_parameters`` :116-117), without the matplotlib dependency."""

_SUFFIXES = ('', 'K', 'M', 'G', 'T', 'P')


def human_format(num):
    """48000 -> '48.0K': one decimal and a thousands suffix, as the reference prints it."""
    idx = 0
    while abs(num) >= 1000:
        idx += 1
        num = num / 1000.0
    return f"{num:.1f}{_SUFFIXES[idx]}"


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


class WeightedBCE:
    """Binary cross-entropy on logits with a positive-class weight (the reference's loss, utils/common.py:11-30).

    ``multi_frame=True``: output/target are (batch, frames, classes) and are cropped to their common number of frames
    (pooling can shorten the output); otherwise the output is flattened to (batch,).
    """

    def __init__(self, recall_factor, multi_frame):
        import torch
        self.recall_factor = torch.tensor([float(recall_factor)])
        self.multi_frame = multi_frame

    def __call__(self, output, target):
        from torch.nn.functional import binary_cross_entropy_with_logits
        if self.multi_frame:
            n = min(output.shape[1], target.shape[1])
            output, target = output[:, :n], target[:, :n]
        else:
            output = output.reshape(-1)
        return binary_cross_entropy_with_logits(output, target, pos_weight=self.recall_factor.to(output.device))
