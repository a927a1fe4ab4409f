"""Generates ``train_reference.npz``: one training iteration of the VERBATIM reference (train.py:96-103) on CPU.

    python tests/golden/make_train_golden.py          (authoring container: needs /root/reference)

For seeded weights (reference init under torch.manual_seed(0), BatchNorm affine parameters and running statistics
randomised as in the inference goldens) and seeded inputs/targets, it runs ``model.train(); out = model(x);
loss = WeightedBCE(recall_factor=5, multi_frame=True)(out, y); loss.backward()`` with the reference's own
``models/spectogram_models.Cnn_AvgPooling`` and ``utils/common.WeightedBCE`` in float32 AND float64 and stores, per case:
logits, loss, the BatchNorm running statistics after the forward pass, and for every parameter gradient its L2 norm and
a strided sample of its elements (float64 run = the value; the float32 run's deviation from it is stored as the
reference's own noise floor).

ReLU ties.  A gradient is discontinuous where a BatchNorm output crosses zero: an element with |y| below the forward
rounding error (~1e-6 in float32, for the reference's own float32 run as much as for this repo's kernels) may be masked
either way, and ONE such flip moves the relative L2 error of a gradient tensor to ~1e-2 -- far above the 1e-3 parity bar,
without either side being wrong.  The strict cases are therefore small and seeded so that every BatchNorm output of the
float64 run keeps |y| >= MARGIN (the seed search is part of this script; the achieved margin is stored); the larger
``loose`` case is kept with its flip-sized tolerance, and the full-size check lives in
tests/test_gpu_train_native.py (float64 autograd with this repo's masks forced).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from make_golden import import_reference, MAIN_CFG  # noqa: E402
from oracle import cnn_ref  # noqa: E402

STRICT = {"B2_T16": (2, 16, 16), "B4_T8": (4, 8, 8), "B3_T13": (3, 13, 16)}     # batch, frames, target frames: tie-free seeds
LOOSE = {"B4_T30": (4, 30, 30), "B2_T182": (2, 182, 182)}
SAMPLE = 384
MARGIN = 2e-5


def case_inputs(B, T, Tt, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, T, 64, generator=g) * 1.5
    y = (torch.rand(B, Tt, 1, generator=g) > 0.8).float()
    return x, y


def sample_idx(n):
    return np.unique(np.linspace(0, n - 1, min(n, SAMPLE)).astype(np.int64))


def relu_margin(RefCnn, x):
    """min |BatchNorm output| over the network in a float64 train-mode forward pass"""
    torch.manual_seed(0)
    m = RefCnn(1, model_config=MAIN_CFG)
    m.load_state_dict(cnn_ref.randomize_bn_({k: v.clone() for k, v in m.state_dict().items()}, seed=7))
    m = m.double().train()
    margins = []
    hooks = [mod.register_forward_hook(lambda _m, _i, o: margins.append(float(o.abs().min())))
             for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        m(x.double())
    for h in hooks:
        h.remove()
    return min(margins)


def main():
    RefCnn, _, _, _ = import_reference()
    from utils.common import WeightedBCE            # the reference's loss (utils/common.py:11-30)
    torch.set_num_threads(8)
    out = {}
    cases = {}
    for name, (B, T, Tt) in STRICT.items():
        for seed in range(2000, 2400):
            x, y = case_inputs(B, T, Tt, seed)
            mg = relu_margin(RefCnn, x)
            if mg >= MARGIN:
                cases[name] = (B, T, Tt, seed)
                out[f"{name}_relu_margin"] = np.float64(mg)
                print(name, "seed", seed, "min |BN output|", mg)
                break
        else:
            raise SystemExit(f"no tie-free seed for {name}")
    for name, (B, T, Tt) in LOOSE.items():
        cases[name] = (B, T, Tt, 1000 + T)
    for name, (B, T, Tt, seed) in cases.items():
        out[f"{name}_shape"] = np.array([B, T, Tt, seed])
        x, y = case_inputs(B, T, Tt, seed)
        res = {}
        for dt in (torch.float64, torch.float32):
            torch.manual_seed(0)
            m = RefCnn(1, model_config=MAIN_CFG)
            sd = cnn_ref.randomize_bn_({k: v.clone() for k, v in m.state_dict().items()}, seed=7)
            m.load_state_dict(sd)
            m = m.to(dt)
            m.train()
            logits = m(x.to(dt))
            loss = WeightedBCE(recall_factor=5, multi_frame=True)(logits, y.to(dt))
            loss.backward()
            res[dt] = (logits.detach(), loss.detach(), [p.grad.detach().reshape(-1) for p in m.parameters()],
                       {k: v.detach().clone() for k, v in m.state_dict().items() if "running" in k})
        l64, loss64, g64, rs64 = res[torch.float64]
        l32, loss32, g32, _ = res[torch.float32]
        out[f"{name}_logits"] = l64.numpy().astype(np.float32)
        out[f"{name}_loss"] = np.float64(loss64)
        names = [n for n, _ in m.named_parameters()]
        for i, n in enumerate(names):
            idx = sample_idx(g64[i].numel())
            out[f"{name}_grad_{i}_idx"] = idx
            out[f"{name}_grad_{i}_val"] = g64[i].numpy()[idx]
            out[f"{name}_grad_{i}_norm"] = np.float64(g64[i].norm())
            out[f"{name}_grad_{i}_ref32_relerr"] = np.float64((g32[i].double() - g64[i]).norm() / g64[i].norm())
        for k, v in rs64.items():
            out[f"{name}_{k}"] = v.numpy().astype(np.float32)
        worst = max(float(out[f"{name}_grad_{i}_ref32_relerr"]) for i in range(len(names)))
        print(name, "loss", float(loss64), "logits", tuple(l64.shape), "params", len(names),
              "reference fp32-vs-fp64 worst rel grad err", worst)
    out["param_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "train_reference.npz"), **out)
    print("saved", os.path.getsize(os.path.join(HERE, "train_reference.npz")), "bytes")


if __name__ == "__main__":
    main()
