"""GPU tests of the training-step support: the fused Adam-amsgrad update and the flat-bucket
trainer reproduce torch.optim.Adam(amsgrad=True) as constructed at the reference's train.py:85."""
import copy
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import sed_b200
from sed_b200 import _ext
from sed_b200.train import DataParallelTrainer
from sed_b200.utils.common import WeightedBCE
import refmodels


def test_fused_adam_amsgrad_matches_torch():
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(10007, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True)
    p = p0.clone().cuda()
    m, v, vm = (torch.zeros_like(p) for _ in range(3))
    lib = _ext.load()
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())      # noqa: E731
    for step in range(1, 6):
        grad = torch.randn(10007, generator=g) * (0.1 if step % 2 else 3.0)      # exercises the running max
        ref.grad = grad.clone()
        opt.step()
        gd = (grad * 4.0).cuda()                                                 # as if summed over 4 ranks
        _ext.check(lib.sedb_adam_amsgrad_step(ptr(p), ptr(gd), ptr(m), ptr(v), ptr(vm), p.numel(), 1e-3, 0.9, 0.999,
                                              1e-8, 0.0, step, 0.25, None))
        assert torch.allclose(p.cpu(), ref.data, atol=2e-7, rtol=1e-5)


def test_trainer_step_matches_reference_loop():
    """Three iterations of train.py:96-103 on the drop-in module: the flat-bucket trainer against torch's own
    Adam(amsgrad=True) fed with the same gradients.  (Adam's first steps are sign-like, lr * g/|g|, so two independent
    cuDNN backward passes that differ in the last bit of a near-zero gradient legitimately diverge by 2 lr; the update
    rule itself is what must agree.)"""
    torch.manual_seed(0)
    a, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
    b = copy.deepcopy(a)
    a, b = a.cuda(), b.cuda()
    crit = WeightedBCE(recall_factor=5, multi_frame=True)
    opt = torch.optim.Adam(a.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True)
    trainer = DataParallelTrainer(b, crit, lr=1e-3)
    for it in range(3):
        x = refmodels.cnn_inputs(30, 300 + it, batch=4).cuda()
        y = (torch.rand(4, 30, 1, generator=torch.Generator().manual_seed(it)) > 0.8).float().cuda()
        a.train()
        with torch.no_grad():
            loss_a = crit(a(x), y)                       # same weights -> same loss (also keeps BN stats in step)
        loss_b = trainer.step(x, y)
        assert abs(float(loss_a) - float(loss_b)) < 3e-4    # torch fp32 forward vs the native split-bf16 forward
        for pa, pb in zip(a.parameters(), b.parameters()):
            pa.grad = pb.grad.detach().clone()
        opt.step()
        for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
            assert n1 == n2 and torch.allclose(p1, p2, atol=1e-6, rtol=1e-5), (it, n1)
    assert trainer.step_count == 3 and float(trainer.max_exp_avg_sq.max()) > 0
    # the updated weights are what the native inference path now sees (the handle repacks after in-place updates)
    a.eval(); b.eval()
    x = refmodels.cnn_inputs(30, 999, batch=2).cuda()
    assert torch.allclose(a(x), b(x), atol=1e-3)


def test_weighted_bce_matches_formula():
    crit = WeightedBCE(recall_factor=5, multi_frame=True)
    out = torch.randn(2, 24, 1)
    tgt = (torch.rand(2, 30, 1) > 0.7).float()
    ref = torch.nn.functional.binary_cross_entropy_with_logits(out, tgt[:, :24], pos_weight=torch.tensor([5.0]))
    assert torch.allclose(crit(out, tgt), ref)
    flat = WeightedBCE(recall_factor=2, multi_frame=False)
    o2, t2 = torch.randn(6, 1), (torch.rand(6) > 0.5).float()
    assert torch.allclose(flat(o2, t2), torch.nn.functional.binary_cross_entropy_with_logits(
        o2.reshape(-1), t2, pos_weight=torch.tensor([2.0])))


def test_native_handle_sees_fused_updates_without_a_train_forward():
    """The fused update writes the parameters behind torch's back; the trainer must invalidate the packed weights of
    the native inference handle (version bump), even when no train-mode forward touched the BatchNorm buffers."""
    from oracle import cnn_ref
    torch.manual_seed(3)
    m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
    m = m.cuda().eval()
    trainer = DataParallelTrainer(m, WeightedBCE(recall_factor=5, multi_frame=True), lr=5e-2)
    x = refmodels.cnn_inputs(30, 77, batch=2).cuda()
    with torch.no_grad():
        y0 = m.logits(x).clone()                        # packs the weights
    g = torch.Generator(device="cuda").manual_seed(5)
    trainer.flat.grad.copy_(torch.randn(trainer.flat.numel, device="cuda", generator=g))
    trainer.apply_update()                              # first Adam step: every weight moves by ~lr
    with torch.no_grad():
        y1 = m.logits(x)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = torch.sigmoid(cnn_ref.cnn_avgpooling_forward(sd, x.cpu(), [p for _, p in refmodels.MAIN_CFG]))
    assert (y1 - y0).abs().max() > 1e-3                 # the update is visible ...
    assert (y1.cpu() - ref).abs().max() < 1e-3          # ... and it is exactly the updated model


def test_trainer_graph_replay_equals_eager_and_checkpoint_roundtrip(tmp_path):
    """The captured CUDA graph of the whole iteration (native forward / loss / backward / update) replays to the same
    parameters as eager execution; the optimizer state loads into torch.optim.Adam(amsgrad=True) (train.py:85,123-128)."""
    crit = WeightedBCE(recall_factor=5, multi_frame=True)
    models, trainers = [], []
    for graph in (False, True):
        m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
        m = m.cuda()
        models.append(m)
        trainers.append(DataParallelTrainer(m, crit, lr=1e-3, graph=graph))
    losses = [[], []]
    for it in range(4):
        x = refmodels.cnn_inputs(30, 700 + it, batch=8).cuda()
        y = (torch.rand(8, 30, 1, generator=torch.Generator().manual_seed(it)) > 0.8).float().cuda()
        for k, tr in enumerate(trainers):
            losses[k].append(float(tr.step(x, y)))
        if it == 0:
            # identical state in, identical arithmetic: the gradients agree up to the order of the few float-atomic sums.
            # (Later iterations may legitimately drift apart: Adam turns a gradient of the size of its eps = 1e-8 into an
            # update of order lr whatever its relative accuracy, and a ReLU tie flipped by such a last-bit difference
            # moves a gradient tensor by ~1e-2 -- see tests/test_gpu_train_native.py.)
            rel_g = float((trainers[0].flat.grad - trainers[1].flat.grad).norm() / trainers[0].flat.grad.norm())
            assert rel_g < 1e-5, rel_g
            for (n, p), (_, q) in zip(models[0].named_buffers(), models[1].named_buffers()):
                assert torch.allclose(p.float(), q.float(), atol=1e-6, rtol=1e-5), n
        if it == 1:
            for tr in trainers:
                tr.decay_lr(0.5)                           # the device-resident learning rate follows
    assert np.allclose(losses[0], losses[1], rtol=1e-3)
    for (n, p), (_, q) in zip(models[0].named_parameters(), models[1].named_parameters()):
        assert float((p - q).norm() / p.norm()) < 1e-3, n          # 4 steps of lr = 1e-3: an un-applied update would show
    assert trainers[1].step_count == 4 and abs(trainers[1].lr - 5e-4) < 1e-12
    # checkpoint in the reference's format; the optimizer part loads into torch's Adam
    ck = {"iterations": 4, "model": models[1].state_dict(), "optimizer": trainers[1].state_dict()}
    path = os.path.join(tmp_path, "iteration_4.pth")
    torch.save(ck, path)
    ck2 = torch.load(path, weights_only=False)
    ref_model, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
    ref_model = ref_model.cuda()
    ref_model.load_state_dict(ck2["model"])
    opt = torch.optim.Adam(ref_model.parameters(), lr=1.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True)
    opt.load_state_dict(ck2["optimizer"])
    assert abs(opt.param_groups[0]["lr"] - 5e-4) < 1e-12
    st = opt.state[next(iter(ref_model.parameters()))]
    assert float(st["step"]) == 4.0 and float(st["max_exp_avg_sq"].abs().max()) > 0
    tr3 = DataParallelTrainer(ref_model, crit, lr=1.0)
    tr3.load_state_dict(ck2["optimizer"])
    assert tr3.step_count == 4 and torch.equal(tr3.exp_avg, trainers[1].exp_avg)


def test_train_dropin_runs_the_reference_loop(tmp_path):
    """sed_b200.train.train: the reference's train() signature on a tiny in-memory loader; writes its checkpoints."""
    from sed_b200.train import train

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 32

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(i)
            return torch.randn(1, 30, 64, generator=g), (torch.rand(30, 1, generator=g) > 0.8).float()

        def get_validation_sampler(self, max_validate_num=None):
            for i in range(2):
                g = torch.Generator().manual_seed(100 + i)
                yield torch.randn(1, 1, 182, 64, generator=g), (torch.rand(1, 182, 1, generator=g) > 0.8).float(), f"clip{i}"

    loader = torch.utils.data.DataLoader(DS(), batch_size=8, shuffle=False, drop_last=True)
    m, _ = refmodels.seeded_cnn(refmodels.MAIN_CFG)
    logs = []
    trainer, hist = train(m, loader, WeightedBCE(recall_factor=5, multi_frame=True), num_steps=6, lr=1e-4, log_freq=3,
                          outputs_dir=str(tmp_path), device=torch.device("cuda"), log=logs.append)
    assert trainer.step_count == 6 and len(hist["train_loss"]) == 6 and all(np.isfinite(hist["train_loss"]))
    assert len(hist["val"]) == 2 and len(hist["val"][0][1]) == 2            # two evaluations of two clips
    assert os.path.exists(os.path.join(tmp_path, "checkpoints", "iteration_3.pth"))
    ck = torch.load(os.path.join(tmp_path, "checkpoints", "iteration_6.pth"), weights_only=False)
    assert set(ck) == {"iterations", "model", "optimizer"} and ck["iterations"] == 6
    assert len(logs) == 2 and "lr:" in logs[0]
