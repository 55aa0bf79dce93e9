"""Step runtime: flat parameter / gradient / shadow buffers and a zero-initialised bump arena.

Why: one training step of asr13x1 touches ~100 parameter tensors and ~60 small accumulators (BatchNorm statistics,
BN-backward totals, weight gradients accumulated with RED).  Allocating, zero-filling and casting each of them with
its own launch costs more than the kernels take, so the step engine (trainer.TrainEngine) installs a `ParamBank`:

  master  fp32 flat buffer; every parameter is a view into it (state_dict / optimizers see ordinary Parameters)
  shadow  bf16 flat buffer, refreshed by ONE cast launch per step; pointwise weights are consumed from it in place
  grads   fp32 flat buffer; p.grad are views; wgrad kernels accumulate straight into them (no autograd add, no copy);
          the buffer doubles as the NCCL all-reduce buckets (ddp.GradSync)
  arena   fp64-aligned scratch that follows `grads` in the same allocation, so ONE memset per step zeroes both

Without an installed bank (plain `module(x)` / `loss.backward()` usage, unit tests) every helper falls back to
per-call torch allocations and per-call casts: same kernels, same results, more launches.
"""
import os

import torch

from . import _lib

_ALIGN = 64  # elements: keeps bf16 views 128-byte and fp32 views 256-byte aligned (TMA needs 16)


class ParamBank:
    def __init__(self, module, arena_bytes=16 << 20):
        # a module, or (optim.Novograd without an engine) a plain list of parameters
        params = [p for p in (module.parameters() if hasattr(module, "parameters") else module)]
        if not params or not params[0].is_cuda:
            raise _lib.LasrError("ParamBank needs a module that already lives on the GPU")
        cur = _BANK["bank"]
        if cur is not None and any(id(p) in cur.offsets for p in params):
            raise _lib.LasrError(
                "these parameters already live in the installed ParamBank: a second bank would re-home them and "
                "silently detach the first one's master / shadow / gradient buffers (captured CUDA graphs and the "
                "fused optimizer keep using those).  Share runtime.current(), or runtime.uninstall() first.")
        self.device = params[0].device
        self.params = params
        self.offsets = {}
        off = 0
        for p in params:
            if p.dtype != torch.float32:
                raise _lib.LasrError("parameters must be fp32 masters")
            self.offsets[id(p)] = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.master = torch.zeros(off, device=self.device, dtype=torch.float32)
        self.shadow = torch.zeros(off, device=self.device, dtype=torch.bfloat16)
        self._zeroed = torch.zeros(off * 4 + arena_bytes, device=self.device, dtype=torch.uint8)
        self.grads = self._zeroed[: off * 4].view(torch.float32)
        self.arena = self._zeroed[off * 4:]
        self._arena_off = 0
        self.arena_high = 0
        self.armed = False
        self.shadow_fresh = False
        self.shadow_synced = False  # set by optim.Novograd: its update pass already rewrote the bf16 shadows
        self.on_grad_ready = None  # callable(param) installed by ddp.GradSync
        # device-resident step counter: added to the dropout seeds inside the kernels, so that a replayed CUDA graph
        # (frozen kernel arguments) still draws fresh masks; only advanced when some module actually drops
        self.step_counter = torch.zeros((), device=self.device, dtype=torch.int64)
        self.count_steps = any(getattr(m, "drop_rate", 0.0) > 0.0 for m in module.modules()) \
            if hasattr(module, "modules") else False
        # weight-gradient kernels are off the critical path of backward (nothing reads them before the step ends):
        # defer() launches them on a lower-priority side stream so they fill the SMs that the main chain leaves idle
        # at kernel boundaries and tails.  Their inputs are kept alive until join_side().  OPT-IN (LASR_SIDE_WGRAD=1):
        # measured on B200 it is a loss (asr13x1 step 3.88 -> 4.01 ms) -- the persistent kernels are bound by L2 / HBM
        # throughput, not by idle SMs, so co-running them only disturbs their static work partition.
        self.side = torch.cuda.Stream(priority=0) if os.environ.get("LASR_SIDE_WGRAD", "0") == "1" else None
        self.keepalive = []
        with torch.no_grad():
            for p in params:
                o = self.offsets[id(p)]
                view = self.master[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grads[o:o + p.numel()].view_as(p)

    # -- per-step protocol -----------------------------------------------------------------------------------
    def begin_step(self):
        """One memset (gradients + arena) and one cast (bf16 weight shadows; skipped when the fused optimizer of the
        previous step already refreshed them)."""
        self._zeroed.zero_()
        self._arena_off = 0
        if self.count_steps:
            self.step_counter.add_(1)
        if not self.shadow_synced:
            _lib.call("lasr_cast_weight", self.master, self.shadow, 1, self.numel, 0, _lib.LASR_BF16)
        self.shadow_synced = False
        self.shadow_fresh = True
        self.armed = True
        for p in self.params:  # an optimizer or user code may have replaced .grad
            o = self.offsets[id(p)]
            if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * o:
                p.grad = self.grads[o:o + p.numel()].view_as(p)

    def end_step(self):
        self.armed = False
        self.shadow_fresh = False  # the optimizer is about to change the masters

    def join_side(self):
        """The current stream waits for every deferred weight-gradient kernel; their inputs may be freed again."""
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self.keepalive.clear()

    def invalidate_shadow(self):
        """Call after changing parameters behind the runtime's back (load_state_dict, manual edits)."""
        self.shadow_synced = False
        self.shadow_fresh = False

    # -- services --------------------------------------------------------------------------------------------
    def zeros(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * (8 if dtype == torch.float64 else 4)
        nbytes_al = (nbytes + 255) // 256 * 256
        self.arena_high = max(self.arena_high, self._arena_off + nbytes_al)
        if not self.armed or self._arena_off + nbytes_al > self.arena.numel():
            return None
        v = self.arena[self._arena_off:self._arena_off + nbytes].view(dtype).view(shape)
        self._arena_off += nbytes_al
        return v

    def shadow_of(self, p):
        o = self.offsets.get(id(p))
        if o is None or not self.shadow_fresh:
            return None
        return self.shadow[o:o + p.numel()].view_as(p)

    def grad_of(self, p):
        o = self.offsets.get(id(p))
        if o is None or not self.armed:
            return None
        return self.grads[o:o + p.numel()].view_as(p)


_BANK = {"bank": None}


def install(bank):
    """Make `bank` the step runtime.  From here on parameters and their bf16 shadows are only written by the step
    prologue (memset + cast launches, full stream-order barriers), so the kernels may prefetch them under programmatic
    dependent launch (include/lasr.h: lasr_set_early_param_loads)."""
    _BANK["bank"] = bank
    _lib.load().lasr_set_early_param_loads(1)


def uninstall():
    _BANK["bank"] = None
    _lib.load().lasr_set_early_param_loads(0)


def current():
    return _BANK["bank"]


def zeros(shape, dtype, device):
    """Zero-initialised accumulator: an arena slice inside an armed step, else a fresh torch.zeros."""
    b = _BANK["bank"]
    if b is not None:
        v = b.zeros(tuple(shape), dtype)
        if v is not None:
            return v
    return torch.zeros(tuple(shape), device=device, dtype=dtype)


def weight(p, dtype):
    """The parameter in the compute dtype: fp32 -> itself; bf16 -> the per-step shadow view, else a per-call cast."""
    if dtype == torch.float32:
        return p.detach()
    b = _BANK["bank"]
    if b is not None:
        v = b.shadow_of(p)
        if v is not None:
            return v
    out = torch.empty(p.shape, device=p.device, dtype=dtype)
    _lib.call("lasr_cast_weight", p.detach().contiguous(), out, 1, p.numel(), 0, _lib.dtype_code(dtype))
    return out


_DROP = {"base": 0x5DEECE66D, "calls": 0}


def set_dropout_seed(seed):
    """Seed of the fused dropout masks (the reference uses torch's global Philox stream; ours is independent of it)."""
    _DROP["base"] = int(seed)
    _DROP["calls"] = 0


def next_dropout_stream():
    """-> (seed, seed_dev) for one dropout call site.  Every call gets its own seed; inside an armed step the bank's
    device step counter is added in the kernel, so a CUDA-graph replay draws new masks although `seed` is frozen."""
    _DROP["calls"] += 1
    seed = (_DROP["base"] + 0x9E3779B97F4A7C15 * _DROP["calls"]) & 0xFFFFFFFFFFFFFFFF
    b = _BANK["bank"]
    if b is not None and b.armed and b.count_steps:
        return seed, b.step_counter
    return seed, None


def grad_sink(p):
    """-> (buffer to ACCUMULATE the gradient of parameter p into, returned_to_autograd).
    Inside an armed step the buffer is p.grad's flat-bucket view and autograd gets None (nothing to add or copy);
    otherwise a fresh zeroed tensor that the Function returns as the gradient."""
    b = _BANK["bank"]
    if b is not None:
        v = b.grad_of(p)
        if v is not None:
            return v, None
    g = torch.zeros_like(p, dtype=torch.float32)
    return g, g


def defer(fn, *keep):
    """Run fn() -- kernels whose results nothing in the rest of backward reads (weight gradients) -- on the bank's side
    stream, ordered after everything enqueued so far on the current stream.  Without an armed bank: just fn()."""
    b = _BANK["bank"]
    if b is None or not b.armed or b.side is None:
        fn()
        return
    b.side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(b.side):
        fn()
    b.keepalive.extend(keep)


def grad_ready(*params):
    """Tell the gradient exchange that these parameters' gradients are final for this step (launch order = stream
    order, so the all-reduce of a completed bucket can be enqueued right away)."""
    b = _BANK["bank"]
    if b is not None and b.armed and b.on_grad_ready is not None:
        for p in params:
            if p is not None:
                b.on_grad_ready(p)
