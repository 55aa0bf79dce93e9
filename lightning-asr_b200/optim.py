"""Drop-in optimizer / LR schedule of the reference's training script, fused on the device.

  Novograd                        mirrors scheduler/novograd.py:30-145 (same constructor, same `state_dict` layout:
                                  state[p] = {step, exp_avg, exp_avg_sq}); train.py:46 builds it with
                                  lr=learning_rate, weight_decay=weight_decay, betas=(0.8, 0.5)
  CosineAnnealingWarmupRestarts   mirrors scheduler/cosine_annearing_with_warmup.py:6-89; train.py:53-55 steps it once
                                  per optimizer step ('interval': 'step')

The reference's step() walks ~100 parameter tensors with a grad.norm(), a host-syncing `if exp_avg_sq == 0` and ~6
elementwise launches each.  Here all parameters live in ONE flat fp32 buffer (runtime.ParamBank) and a step is three
launches of `lasr_novograd_step` (csrc/optim.cu) with no host round trip, so TrainEngine captures the optimizer -- and
the schedule, whose state is a small device struct -- inside the step's CUDA graph.  The bf16 weight shadows consumed
by the tensor-core kernels are refreshed by the update pass itself.

No CPU path: parameters must be CUDA tensors (LasrError otherwise).
"""
import ctypes
import math

import torch
from torch.optim.optimizer import Optimizer

from . import _lib, runtime

_CHUNK = 4096  # elements per CTA of the norm / update passes


class _SchedStruct(ctypes.Structure):
    """include/lasr.h: lasr_lr_sched_t"""
    _fields_ = [("base_max_lr", ctypes.c_double), ("max_lr", ctypes.c_double), ("min_lr", ctypes.c_double),
                ("cycle_mult", ctypes.c_double), ("gamma", ctypes.c_double), ("lr", ctypes.c_double),
                ("first_cycle_steps", ctypes.c_int32), ("cur_cycle_steps", ctypes.c_int32),
                ("warmup_steps", ctypes.c_int32), ("cycle", ctypes.c_int32), ("step_in_cycle", ctypes.c_int32),
                ("last_epoch", ctypes.c_int32)]


def _check_valid_opt_params(lr, eps, betas):  # scheduler/novograd.py:21-27
    if lr < 0:
        raise ValueError(f"Invalid learning rate: {lr}")
    if eps < 0:
        raise ValueError(f"Invalid epsilon value: {eps}")
    if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
        raise ValueError(f"Betas have to be between 0 and 1: {betas}")


class Novograd(Optimizer):
    """Novograd (https://arxiv.org/abs/1905.11286) with the reference's exact update rule; see module docstring.

    `bank`: the runtime.ParamBank that owns the parameters (TrainEngine passes its own).  Without one the optimizer
    re-homes the parameters it is given into a private bank (their values are preserved; `.data` become views)."""

    def __init__(self, params, lr=1e-3, betas=(0.95, 0.98), eps=1e-8, weight_decay=0, grad_averaging=False,
                 amsgrad=False, luc=False, luc_trust=1e-3, luc_eps=1e-8, bank=None):
        _check_valid_opt_params(lr, eps, betas)
        if amsgrad or luc:
            raise NotImplementedError("amsgrad / luc are not on the reference's training path (train.py:46) and are "
                                      "not implemented by the fused kernel")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_averaging=grad_averaging,
                        amsgrad=amsgrad)
        self.luc, self.luc_trust, self.luc_eps = luc, luc_trust, luc_eps
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError("the fused Novograd step handles one parameter group (as train.py:46 builds)")
        plist = [p for p in self.param_groups[0]["params"]]
        if not plist or not plist[0].is_cuda:
            raise _lib.LasrError("Novograd: parameters must live on a CUDA device (there is no CPU path)")
        _lib.require_device()
        if bank is None:
            bank = runtime.current()
        if bank is None or any(id(p) not in bank.offsets for p in plist):
            bank = runtime.ParamBank(plist)
        self.bank = bank
        dev = bank.device
        self._plist = plist
        # chunk table: pieces of <= _CHUNK elements, each inside one parameter tensor
        off, ln, par = [], [], []
        for i, p in enumerate(plist):
            o, n = bank.offsets[id(p)], p.numel()
            for s in range(0, n, _CHUNK):
                off.append(o + s)
                ln.append(min(_CHUNK, n - s))
                par.append(i)
        self._chunk_off = torch.tensor(off, dtype=torch.int32, device=dev)
        self._chunk_len = torch.tensor(ln, dtype=torch.int32, device=dev)
        self._chunk_par = torch.tensor(par, dtype=torch.int32, device=dev)
        self._exp_avg = torch.zeros(bank.numel, dtype=torch.float32, device=dev)
        self._exp_avg_sq = torch.zeros(len(plist), dtype=torch.float32, device=dev)
        self._denom = torch.zeros(len(plist), dtype=torch.float32, device=dev)
        self._lr_use = torch.zeros(1, dtype=torch.float32, device=dev)
        self._sched_dev = None  # uint8 tensor holding a lasr_lr_sched_t once a scheduler is attached
        self._steps = 0
        for i, p in enumerate(plist):  # reference state layout (scheduler/novograd.py:103-112), as views
            o = bank.offsets[id(p)]
            self.state[p] = {"step": 0, "exp_avg": self._exp_avg[o:o + p.numel()].view_as(p),
                             "exp_avg_sq": self._exp_avg_sq[i]}

    # -- schedule on the device ------------------------------------------------------------------------------
    def attach_schedule(self, sched):
        """Move a CosineAnnealingWarmupRestarts' state into device memory; from now on every step() uses the
        schedule's current learning rate and then advances it (what train.py's per-step scheduler.step() does)."""
        st = _SchedStruct(sched.base_max_lr, sched.max_lr, sched.min_lr, float(sched.cycle_mult), sched.gamma,
                          float(self.param_groups[0]["lr"]), sched.first_cycle_steps, int(sched.cur_cycle_steps),
                          sched.warmup_steps, sched.cycle, sched.step_in_cycle, sched.last_epoch)
        raw = torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8).clone()
        self._sched_dev = raw.to(self.bank.device)

    def schedule_state(self):
        """Device schedule state as a dict (synchronises)."""
        if self._sched_dev is None:
            return None
        st = _SchedStruct.from_buffer_copy(bytes(self._sched_dev.cpu().numpy().tobytes()))
        return {name: getattr(st, name) for name, _ in _SchedStruct._fields_}

    def last_lr(self):
        """The learning rate the most recent step() applied (synchronises)."""
        return float(self._lr_use.item())

    # -- checkpoint / resume -----------------------------------------------------------------------------------
    def state_dict(self):
        """torch.optim layout (state[i] = {step, exp_avg, exp_avg_sq} like scheduler/novograd.py) plus the device-side
        pieces the fused step really reads: the LR-schedule struct and the step count."""
        sd = super().state_dict()
        sd["lasr"] = {"steps": self._steps,
                      "sched": None if self._sched_dev is None else self._sched_dev.detach().cpu().clone(),
                      "lr_use": self._lr_use.detach().cpu().clone()}
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """Resume: the inherited loader replaces self.state with fresh tensors, but the fused kernel only reads the flat
        buffers -- copy the loaded moments into them, re-create the views, restore the schedule struct and step count."""
        state_dict = dict(state_dict)
        extra = state_dict.pop("lasr", None)
        super().load_state_dict(state_dict)
        bank = self.bank
        steps = 0
        for i, p in enumerate(self._plist):
            st = self.state.get(p)
            o = bank.offsets[id(p)]
            if st is not None and "exp_avg" in st:
                self._exp_avg[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1).to(self._exp_avg.dtype))
                self._exp_avg_sq[i].copy_(torch.as_tensor(st["exp_avg_sq"]).reshape(()).to(self._exp_avg_sq.dtype))
                steps = max(steps, int(st.get("step", 0)))
            else:
                self._exp_avg[o:o + p.numel()].zero_()
                self._exp_avg_sq[i].zero_()
            self.state[p] = {"step": int(st.get("step", 0)) if st else 0,
                             "exp_avg": self._exp_avg[o:o + p.numel()].view_as(p), "exp_avg_sq": self._exp_avg_sq[i]}
        self._steps = steps
        if extra is not None:
            self._steps = int(extra.get("steps", steps))
            if extra.get("sched") is not None:
                if self._sched_dev is not None and self._sched_dev.numel() == extra["sched"].numel():
                    self._sched_dev.copy_(extra["sched"].to(bank.device))  # in place: a captured graph keeps the pointer
                else:
                    self._sched_dev = extra["sched"].to(bank.device).clone()
            if extra.get("lr_use") is not None:
                self._lr_use.copy_(extra["lr_use"].to(bank.device))
        bank.invalidate_shadow()  # the caller usually reloads the model too: force a fresh bf16 cast

    # -- the step ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        bank, group = self.bank, self.param_groups[0]
        # gradients must be the bank's flat views (they are inside TrainEngine); copy stray ones in
        for p in self._plist:
            o = bank.offsets[id(p)]
            view_ptr = bank.grads.data_ptr() + 4 * o
            if p.grad is None:
                raise _lib.LasrError("Novograd.step: a parameter has no gradient; the fused step updates every "
                                     "parameter it was built with")
            if p.grad.data_ptr() != view_ptr:
                bank.grads[o:o + p.numel()].copy_(p.grad.reshape(-1))
        norms = runtime.zeros((len(self._plist),), torch.float64, bank.device)
        beta1, beta2 = group["betas"]
        _lib.call("lasr_novograd_step", bank.master, bank.grads, self._exp_avg, bank.shadow, self._chunk_off,
                  self._chunk_len, self._chunk_par, self._chunk_off.numel(), norms, self._exp_avg_sq, self._denom,
                  len(self._plist), self._sched_dev, self._lr_use, group["lr"], beta1, beta2, group["eps"],
                  group["weight_decay"], 1 if group["grad_averaging"] else 0)
        bank.shadow_synced = True  # the update pass rewrote the bf16 shadows: the next step needs no cast launch
        self._steps += 1
        for st in self.state.values():
            st["step"] = self._steps
        return loss


class CosineAnnealingWarmupRestarts:
    """Host mirror of scheduler/cosine_annearing_with_warmup.py (constructor arguments, attributes, get_lr / step).
    It keeps `optimizer.param_groups[i]['lr']` up to date exactly like the reference; when the optimizer is the fused
    Novograd, `optimizer.attach_schedule(self)` additionally moves the state to the device so the per-step update
    needs no host involvement (step() here then only advances the host copy, e.g. for logging)."""

    def __init__(self, optimizer, first_cycle_steps, cycle_mult=1., max_lr=0.1, min_lr=0.001, warmup_steps=0, gamma=1.,
                 last_epoch=-1):
        assert warmup_steps < first_cycle_steps
        self.optimizer = optimizer
        self.first_cycle_steps = first_cycle_steps
        self.cycle_mult = cycle_mult
        self.base_max_lr = max_lr
        self.max_lr = max_lr
        self.min_lr = min_lr
        self.warmup_steps = warmup_steps
        self.gamma = gamma
        self.cur_cycle_steps = first_cycle_steps
        self.cycle = 0
        self.step_in_cycle = last_epoch
        self.last_epoch = last_epoch
        self.base_lrs = []
        for g in optimizer.param_groups:  # init_lr
            g["lr"] = min_lr
            self.base_lrs.append(min_lr)
        self.step()  # torch's _LRScheduler base class performs one step at construction

    def get_lr(self):
        if self.step_in_cycle == -1:
            return list(self.base_lrs)
        if self.step_in_cycle < self.warmup_steps:
            return [(self.max_lr - b) * self.step_in_cycle / self.warmup_steps + b for b in self.base_lrs]
        c = math.cos(math.pi * (self.step_in_cycle - self.warmup_steps) / (self.cur_cycle_steps - self.warmup_steps))
        return [b + (self.max_lr - b) * (1 + c) / 2 for b in self.base_lrs]

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
            self.step_in_cycle += 1
            if self.step_in_cycle >= self.cur_cycle_steps:
                self.cycle += 1
                self.step_in_cycle -= self.cur_cycle_steps
                self.cur_cycle_steps = int((self.cur_cycle_steps - self.warmup_steps) * self.cycle_mult) + self.warmup_steps
        elif epoch >= self.first_cycle_steps:
            if self.cycle_mult == 1.:
                self.step_in_cycle = epoch % self.first_cycle_steps
                self.cycle = epoch // self.first_cycle_steps
            else:
                n = int(math.log(epoch / self.first_cycle_steps * (self.cycle_mult - 1) + 1, self.cycle_mult))
                self.cycle = n
                self.step_in_cycle = epoch - int(self.first_cycle_steps * (self.cycle_mult ** n - 1) / (self.cycle_mult - 1))
                self.cur_cycle_steps = self.first_cycle_steps * self.cycle_mult ** n
        else:
            self.cur_cycle_steps = self.first_cycle_steps
            self.step_in_cycle = epoch
        self.max_lr = self.base_max_lr * (self.gamma ** self.cycle)
        self.last_epoch = math.floor(epoch)
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g["lr"] = lr

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)
