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
