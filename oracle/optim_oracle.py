"""CPU restatement of the reference optimizer path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  novograd_step        scheduler/novograd.py:75-145  (Novograd.step; amsgrad=False, luc=False as train.py:46 uses it)
  CosineWarmupOracle   scheduler/cosine_annearing_with_warmup.py:21-89 (CosineAnnealingWarmupRestarts, epoch=None path)

Pinned by tests/test_oracle_golden.py::test_optimizer_oracle_matches_reference against tests/golden/optim.pt, which
tests/golden/make_golden_optim.py produced by running the reference's OWN classes (imported from /root/reference).
"""
import math

import torch


def novograd_step(params, grads, state, lr, betas=(0.95, 0.98), eps=1e-8, weight_decay=0.0, grad_averaging=False):
    """One Novograd step over lists of fp32 tensors, in place.  `state` is a list of dicts (empty on first use)."""
    beta1, beta2 = betas
    for p, g, st in zip(params, grads, state):
        if not st:  # :103-112
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p)
            st["exp_avg_sq"] = torch.zeros([])
        st["step"] += 1
        norm = g.norm().pow(2)  # :113 -- ONE second moment per tensor ("layer-wise")
        if st["exp_avg_sq"] == 0:  # :115-118
            st["exp_avg_sq"] = norm.clone()
        else:
            st["exp_avg_sq"] = st["exp_avg_sq"] * beta2 + (1.0 - beta2) * norm
        denom = st["exp_avg_sq"].sqrt() + eps  # :126
        g = g / denom  # :128
        if weight_decay != 0:
            g = g + weight_decay * p  # :129-130  (decay added AFTER the normalisation)
        if grad_averaging:
            g = g * (1 - beta1)  # :131-132
        st["exp_avg"].mul_(beta1).add_(g)  # :133
        p.add_(st["exp_avg"], alpha=-lr)  # :143


class CosineWarmupOracle:
    """The learning-rate sequence of CosineAnnealingWarmupRestarts driven with step() (epoch=None), including the
    construction-time step of torch's _LRScheduler base class.  `lr` is what the optimizer uses NEXT."""

    def __init__(self, first_cycle_steps, cycle_mult=1.0, max_lr=0.1, min_lr=0.001, warmup_steps=0, gamma=1.0):
        assert warmup_steps < first_cycle_steps  # :30
        self.first_cycle_steps = first_cycle_steps
        self.cycle_mult = cycle_mult
        self.base_max_lr = max_lr
        self.max_lr = max_lr
        self.min_lr = min_lr
        self.warmup_steps = warmup_steps
        self.gamma = gamma
        self.cur_cycle_steps = first_cycle_steps
        self.cycle = 0
        self.step_in_cycle = -1
        self.last_epoch = -1
        self.lr = min_lr  # init_lr :47-51
        self.step()  # _LRScheduler.__init__ performs one step

    def get_lr(self):  # :53-62
        if self.step_in_cycle == -1:
            return self.min_lr
        if self.step_in_cycle < self.warmup_steps:
            return (self.max_lr - self.min_lr) * self.step_in_cycle / self.warmup_steps + self.min_lr
        return self.min_lr + (self.max_lr - self.min_lr) * (
            1 + math.cos(math.pi * (self.step_in_cycle - self.warmup_steps) / (self.cur_cycle_steps - self.warmup_steps))
        ) / 2

    def step(self):  # :64-72, 86-89
        self.last_epoch += 1
        self.step_in_cycle += 1
        if self.step_in_cycle >= self.cur_cycle_steps:
            self.cycle += 1
            self.step_in_cycle -= self.cur_cycle_steps
            self.cur_cycle_steps = int((self.cur_cycle_steps - self.warmup_steps) * self.cycle_mult) + self.warmup_steps
        self.max_lr = self.base_max_lr * (self.gamma ** self.cycle)
        self.lr = self.get_lr()
        return self.lr
