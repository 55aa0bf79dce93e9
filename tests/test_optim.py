"""Fused Novograd + LR schedule (lightning_asr_b200.optim, csrc/optim.cu) against the oracle and against the golden
fixture produced by the reference's own optimizer (tests/golden/make_golden_optim.py)."""
import os
import sys

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)

from golden_common import optim_case  # noqa: E402


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def test_host_schedule_mirror_matches_reference_lrs():
    """optim.CosineAnnealingWarmupRestarts (host mirror) reproduces the reference scheduler's LR sequence."""
    from lightning_asr_b200.optim import CosineAnnealingWarmupRestarts

    class FakeOpt:
        param_groups = [{"lr": 123.0}]

    hyper, sched, steps = optim_case.CASES["train_py"]
    opt = FakeOpt()
    sch = CosineAnnealingWarmupRestarts(opt, **sched)
    fx = _load("optim.pt")["train_py"]
    for k in range(steps):
        assert abs(opt.param_groups[0]["lr"] - fx["lrs"][k]) <= 1e-15, k
        sch.step()
    # the epoch-given branch (cosine_annearing_with_warmup.py:73-85): inside the first cycle it just sets the position
    sch2 = CosineAnnealingWarmupRestarts(FakeOpt(), **sched)
    sch2.step(4)
    assert (sch2.cycle, sch2.step_in_cycle, sch2.cur_cycle_steps, sch2.last_epoch) == (0, 4, 6, 4)
    sch2.step(17)  # cycle_mult = 2: n = int(log2(17/6 + 1)) = 1 -> cycle 1, 11 steps into a 12-step cycle
    assert (sch2.cycle, sch2.step_in_cycle, sch2.cur_cycle_steps) == (1, 11, 12)


def test_novograd_rejects_bad_arguments_like_the_reference():
    from lightning_asr_b200 import _lib
    from lightning_asr_b200.optim import Novograd

    p = [torch.nn.Parameter(torch.zeros(4))]
    with pytest.raises(ValueError):
        Novograd(p, lr=-1.0)
    with pytest.raises(ValueError):
        Novograd(p, betas=(1.0, 0.5))
    with pytest.raises(ValueError):
        Novograd(p, eps=-1e-3)
    with pytest.raises(_lib.LasrError):  # no CPU path
        Novograd(p, lr=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(optim_case.CASES))
def test_fused_novograd_matches_reference_fixture(case):
    from lightning_asr_b200.optim import CosineAnnealingWarmupRestarts, Novograd

    hyper, sched, steps = optim_case.CASES[case]
    fx = _load("optim.pt")[case]
    params = [torch.nn.Parameter(p.cuda()) for p in optim_case.params()]
    opt = Novograd(params, **hyper)
    if sched is not None:
        opt.attach_schedule(CosineAnnealingWarmupRestarts(opt, **sched))
    snaps = dict(fx["snaps"])
    for k in range(steps):
        for p, g in zip(params, optim_case.grads(k)):
            p.grad = g.cuda()
        opt.step()
        assert abs(opt.last_lr() - fx["lrs"][k]) <= 1e-7 * max(fx["lrs"][k], 1e-4), (k, opt.last_lr(), fx["lrs"][k])
        if k in snaps:
            for a, b in zip(params, snaps[k]):
                assert torch.allclose(a.detach().cpu(), b, rtol=2e-6, atol=1e-8), (case, k)
    for p, v, m in zip(params, fx["exp_avg_sq"], fx["exp_avg"]):
        st = opt.state[p]
        assert abs(float(st["exp_avg_sq"]) - v) <= 2e-6 * abs(v)
        # 24 steps of m = b1*m + g' in fp32: FMA contraction on the device vs separate mul/add in torch
        assert torch.allclose(st["exp_avg"].cpu(), m, rtol=2e-5, atol=2e-7)
    # bf16 shadows are refreshed by the update pass
    for p in params:
        o = opt.bank.offsets[id(p)]
        sh = opt.bank.shadow[o:o + p.numel()].view_as(p)
        assert torch.equal(sh, p.detach().bfloat16())


@pytest.mark.gpu
def test_fused_novograd_large_tensor_and_odd_sizes_vs_oracle():
    """Tensors spanning several chunks and sizes that are not multiples of 4 (scalar tails)."""
    from lightning_asr_b200.optim import Novograd
    from oracle import optim_oracle

    g = torch.Generator().manual_seed(5)
    shapes = [(1024, 87), (10001,), (3,), (4097,)]
    ref = [torch.randn(s, generator=g) * 0.2 for s in shapes]
    params = [torch.nn.Parameter(p.clone().cuda()) for p in ref]
    opt = Novograd(params, lr=3e-3, betas=(0.8, 0.5), weight_decay=1e-4)
    state = [{} for _ in ref]
    for k in range(4):
        grads = [torch.randn(s, generator=g) * 0.05 for s in shapes]
        for p, gr in zip(params, grads):
            p.grad = gr.cuda()
        opt.step()
        optim_oracle.novograd_step(ref, grads, state, 3e-3, betas=(0.8, 0.5), weight_decay=1e-4)
    for a, b in zip(params, ref):
        assert torch.allclose(a.detach().cpu(), b, rtol=5e-6, atol=1e-8)


@pytest.mark.gpu
def test_novograd_checkpoint_resume_matches_uninterrupted_run():
    """save -> load -> continue must equal the uninterrupted run: the moments live in flat device buffers and
    the LR schedule in a device struct, neither of which the inherited load_state_dict touches."""
    import copy

    from lightning_asr_b200.optim import CosineAnnealingWarmupRestarts, Novograd

    hyper, sched, steps = optim_case.CASES["train_py"]

    def make():
        params = [torch.nn.Parameter(p.cuda()) for p in optim_case.params()]
        opt = Novograd(params, **hyper)
        opt.attach_schedule(CosineAnnealingWarmupRestarts(opt, **sched))
        return params, opt

    def run(params, opt, lo, hi):
        for k in range(lo, hi):
            for p, g in zip(params, optim_case.grads(k)):
                p.grad = g.cuda()
            opt.step()

    pa, oa = make()
    run(pa, oa, 0, 12)
    pb, ob = make()
    run(pb, ob, 0, 5)
    ckpt = copy.deepcopy({"opt": ob.state_dict(), "params": [p.detach().clone() for p in pb]})
    assert ckpt["opt"]["state"][0]["exp_avg"].abs().sum() > 0 and ckpt["opt"]["lasr"]["steps"] == 5
    pc, oc = make()  # a fresh process: new parameters, new optimizer, then resume
    with torch.no_grad():
        for p, q in zip(pc, ckpt["params"]):
            p.copy_(q)
    oc.load_state_dict(ckpt["opt"])
    assert oc.state[pc[0]]["exp_avg"].data_ptr() == oc._exp_avg.data_ptr() + 4 * oc.bank.offsets[id(pc[0])]
    assert oc.schedule_state() == ob.schedule_state()
    run(pc, oc, 5, 12)
    assert oc.last_lr() == oa.last_lr()
    for a, c in zip(pa, pc):
        assert torch.equal(a.detach(), c.detach())
    for a, c in zip(pa, pc):
        assert torch.equal(oa.state[a]["exp_avg"], oc.state[c]["exp_avg"])
        assert torch.equal(oa.state[a]["exp_avg_sq"], oc.state[c]["exp_avg_sq"])
        assert oc.state[c]["step"] == 12
