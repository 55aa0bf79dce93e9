"""The caller of the hot path: a PyTorch-Lightning-free mirror of train.py's LightingModule (train.py:23-198) and a
step engine that runs one training step (H2D -> forward -> CTC -> backward [-> gradient all-reduce] [-> optimizer])
on one GPU, optionally replayed from a CUDA graph.

  LightingModule.training_step / validation_step / forward keep the reference's batch contract
      batch = (inputs [N,1,64,T] fp32, targets [N,S] long, percents [N] fp32, target_sizes [N] int32, paths)
  and its arithmetic: t_lengths = torch.mul(T', percents).int() (:76), loss = mean(CTCLoss(...)) (:77-78).

pytorch_lightning / hydra / comet are the reference's control plane and are out of scope;
`self.log` becomes a plain dict (`self.logged`).
"""
import os

import torch
import torch.nn as nn

from . import _lib, runtime
from .ctc import CTCLoss
from .metrics import WER
from .quartznet import build_model

SAMPLE_RATE = 16000


def num_frames(num_samples):
    """T = 1 + (S + 2*pad) // hop with pad=32, hop=160 (data_module.py:68-70)."""
    return 1 + (num_samples + 64) // 160


def synthetic_batch(n, seconds, n_labels, seed=1234, ragged=False, features=True):
    """SURVEY.md section 8d synthetic inputs.  features=True draws the normalised log-mel tensor directly
    (randn [N,1,64,T], zero after each utterance's length like _collate_fn's padding); features=False returns raw
    waveforms 0.05*randn clamped to [-1, 1] ([N, S] fp32) plus per-utterance sample counts for the GPU frontend."""
    g = torch.Generator().manual_seed(seed)
    S = int(round(seconds * SAMPLE_RATE))
    T = num_frames(S)
    frac = torch.linspace(0.6, 1.0, n) if ragged else torch.ones(n)
    if features:
        lens_t = torch.clamp((frac * T).round().long(), 1, T)
        lens_t[-1] = T
        x = torch.randn(n, 1, 64, T, generator=g)
        t = torch.arange(T)
        x = x * (t[None, :] < lens_t[:, None])[:, None, None, :]
        percents = lens_t.float() / float(T)  # data_module.py:244
        first = x
    else:
        lens_s = torch.clamp((frac * S).round().long(), 400, S)
        lens_s[-1] = S
        w = (0.05 * torch.randn(n, S, generator=g)).clamp_(-1.0, 1.0)
        s = torch.arange(S)
        first = (w * (s[None, :] < lens_s[:, None]), lens_s.int())
        lens_t = 1 + (lens_s + 64) // 160
        percents = lens_t.float() / float(T)
    Tp = (T - 1) // 2 + 1
    t_len = torch.mul(Tp, percents).int()
    tgt_len = torch.clamp(t_len // 4, min=1).int()
    S_max = int(tgt_len.max())
    targets = torch.randint(0, n_labels, (n, S_max), generator=g)
    targets = targets * (torch.arange(S_max)[None, :] < tgt_len[:, None])
    return first, targets.long(), percents, tgt_len, [f"synthetic_{i}" for i in range(n)]


def profile_step_graph(run, replays=5, only=None):
    """Per-kernel device time of a step AS IT RUNS INSIDE A CUDA GRAPH: `run()` is captured once more with a timing
    event recorded before and after every C-ABI call as external event-record nodes, the graph is replayed `replays`
    times and the event pairs are read after each replay.  -> (calls, step_ms) with calls = [(name, int/float args,
    pointer flags, median ms)] in launch order and step_ms = median device time of the instrumented replay.
    The instrumented graph differs from the timed one: every event node costs about as much as an empty kernel and
    removes the programmatic-dependent-launch overlap between its neighbours, so a fully instrumented step runs ~30 %
    longer than the plain one.  `only` (callable(name) -> bool) restricts the events to some entry points -- bench.py
    times ONE kernel family per captured graph so that the rest of the step runs undisturbed; calls without events
    come back with ms = None."""
    import statistics

    s = torch.cuda.Stream(priority=-1)
    s.wait_stream(torch.cuda.current_stream())
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    _lib.PROFILE, _lib.PROFILE_EXTERNAL, _lib.PROFILE_ONLY = [], True, only
    try:
        with torch.cuda.graph(graph, stream=s):
            run()
    finally:
        prof, _lib.PROFILE, _lib.PROFILE_EXTERNAL, _lib.PROFILE_ONLY = _lib.PROFILE, None, False, None
    torch.cuda.synchronize()
    per_call = [[] for _ in prof]
    steps = []
    for _ in range(replays + 1):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        graph.replay()
        t1.record()
        torch.cuda.synchronize()
        steps.append(t0.elapsed_time(t1))
        for i, (_, _, _, e0, e1) in enumerate(prof):
            if e0 is not None:
                per_call[i].append(e0.elapsed_time(e1))
    calls = [(name, args, has, statistics.median(ts[1:]) if ts else None)
             for (name, args, has, _, _), ts in zip(prof, per_call)]
    return calls, statistics.median(steps[1:])


def _flat_views(tensors, device, pin=False):
    """One flat byte buffer on `device` + a view per tensor of `tensors` (same dtype / shape, 256-byte aligned);
    non-tensor entries pass through."""
    offs, total = [], 0
    for t in tensors:
        if torch.is_tensor(t):
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        else:
            offs.append(None)
    flat = torch.zeros(max(total, 256), dtype=torch.uint8, device=device)
    if pin:
        flat = flat.pin_memory()
    views = []
    for t, o in zip(tensors, offs):
        if o is None:
            views.append(t)
        else:
            nb = t.numel() * t.element_size()
            views.append(flat[o:o + nb].view(t.dtype).view(t.shape))
    return flat, views


class LightingModule(nn.Module):
    """train.py:23-198 without pytorch_lightning.  `encoder` is the drop-in MyModel2 of `model_name`."""

    def __init__(self, learning_rate=5e-3, weight_decay=1e-4, labels=None, total_epoch=50, drop_rate=0., mask=False,
                 use_cer=False, model_name="asr13x1", precision=None, in_c=64):
        super().__init__()
        self.learning_rate = learning_rate
        self.weight_decay = weight_decay
        self.total_epoch = total_epoch
        self.labels = labels
        self.wer = WER(vocabulary=labels, use_cer=use_cer)  # :195
        self.loss = CTCLoss(blank=len(labels), reduction="none")  # :196
        self.encoder = build_model(model_name, labels, drop_rate=drop_rate, mask=mask, in_c=in_c,
                                   precision=precision)  # :197
        self.logged = {}

    def log(self, name, value, **_):
        self.logged[name] = value

    def configure_optimizers(self, steps_per_epoch=1000, bank=None):
        """train.py:36-62: Novograd(lr, weight_decay, betas=(0.8, 0.5)) + CosineAnnealingWarmupRestarts stepped once
        per optimizer step.  Returns the same ([optimizer], [{'scheduler', 'interval', 'monitor'}]) pair; the schedule
        is also attached to the fused optimizer so the per-step LR update happens on the device."""
        from .optim import CosineAnnealingWarmupRestarts, Novograd
        novo_optim = Novograd(self.parameters(), lr=self.learning_rate, weight_decay=self.weight_decay,
                              betas=(0.8, 0.5), bank=bank)
        lr_scheduler = CosineAnnealingWarmupRestarts(novo_optim, first_cycle_steps=self.total_epoch * steps_per_epoch,
                                                     cycle_mult=2, max_lr=self.learning_rate, min_lr=1e-4,
                                                     warmup_steps=1000, gamma=0.5)
        novo_optim.attach_schedule(lr_scheduler)
        return [novo_optim], [{"scheduler": lr_scheduler, "interval": "step", "monitor": "val_loss"}]

    def forward(self, inputs, percentage):
        return self.encoder(inputs, percentage)  # :34

    # -- reference-shaped steps ---------------------------------------------------------------------------------
    def training_step_tensors(self, batch):
        """train.py:71-78 -> (loss, out [N,T',V'], t_lengths)."""
        input, trans, percentage, trans_lengths = batch[0], batch[1], batch[2], batch[3]
        out = self.encoder(input, percentage)
        t_lengths = torch.mul(out.size(1), percentage).int()
        loss = torch.mean(self.loss(out.transpose(0, 1), trans, t_lengths, trans_lengths))
        return loss, out, t_lengths

    def training_step(self, batch, batch_idx=0, log_wer=True):
        loss, out, t_lengths = self.training_step_tensors(batch)
        self.log("train_loss", loss)
        if log_wer:  # train.py:80 does this every step (a D2H sync); callers may make it periodic
            self.log("train_wer", self.wer(out.argmax(dim=-1, keepdim=False), batch[1], batch[3], t_lengths))
        return loss

    def validation_step(self, batch, batch_idx=0):
        """train.py:88-116 (model must be in eval mode, as PL puts it)."""
        with torch.no_grad():
            loss, out, t_lengths = self.training_step_tensors(batch)
            pred = out.argmax(dim=-1, keepdim=False)
            wer = self.wer(pred, batch[1], batch[3], t_lengths)
            self.log("val_wer", wer)
            self.log("val_loss", loss)
            return {"val_loss": loss, "input": batch[0], "val_wer": wer,
                    "pred": self.wer.ctc_decoder_predictions_tensor(pred, t_lengths),
                    "true": self.wer.decode_reference(batch[1], batch[3]), "path": batch[-1]}

    def test_step(self, batch, batch_idx=0):
        r = self.validation_step(batch, batch_idx)
        return {"test_loss": r["val_loss"], "input": r["input"], "test_wer": r["val_wer"], "pred": r["pred"],
                "true": r["true"], "path": r["path"]}

    # -- fast path: same arithmetic, log-probs never materialised ------------------------------------------------
    def training_step_fused(self, batch):
        """loss = mean_n CTC(log_softmax(decoder(encoder(x)))) with the decoder GEMM, the row log-sum-exp, the CTC
        lattices and the (softmax - occupancy) gradient fused (functions.FusedDecoderCTCFn).  -> (loss, logits, t_len)"""
        nll, logits, t_lengths = self.encoder.forward_fused_ctc(batch[0], batch[2], batch[1], batch[3])
        return torch.mean(nll), logits, t_lengths


class TrainEngine:
    """One process per GPU.  step_host(batch) is the end-to-end call (pinned host batch -> H2D -> step -> loss on the
    host); step_device() re-runs the step on the batch already resident in HBM.

    The engine owns a runtime.ParamBank (flat fp32 masters / bf16 shadows / fp32 gradients + zeroed arena: one memset
    and one cast launch per step) and, with graph=True, captures the whole step -- memset, cast, forward, CTC,
    backward, gradient all-reduce, optimizer -- into ONE CUDA graph that is replayed every step."""

    def __init__(self, module, example_batch, graph=True, optimizer=None, grad_sync=None, fused=True, world_sync=None,
                 augment=False, dither=False, wave_dtype=torch.int16):
        """example_batch: the reference's training tuple (inputs [N,1,64,T] fp32 features, targets, percents,
        target_sizes, ...) -- or, for the step that starts from WAVEFORMS (the log-mel frontend on the training path,
        data_module.py:150-174 moved onto the device), ((waves [N, S_max], num_samples [N]), targets, _,
        target_sizes, ...): the waveforms travel as 16-bit PCM (wave_dtype=torch.float32 keeps fp32 samples, used by the
        parity tests), the frontend runs inside the step graph, `augment` / `dither` switch the train-time crop +
        SpecAugment and the dither on (drawn on the device, re-keyed every step by the bank's step counter)."""
        _lib.require_device()
        self.module = module
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.bank = runtime.ParamBank(module)
        runtime.install(self.bank)
        if optimizer == "novograd":  # the reference's configure_optimizers (train.py:36-62) over this engine's bank
            optimizer = module.configure_optimizers(bank=self.bank)[0][0]
        elif callable(optimizer) and not isinstance(optimizer, torch.optim.Optimizer):
            optimizer = optimizer(self.bank)
        self.optimizer = optimizer
        self.grad_sync = grad_sync
        if world_sync is not None:  # (group, bucket_mb): build the gradient exchange over the bank's flat buffer
            if torch.distributed.is_initialized() and torch.distributed.get_world_size(world_sync[0]) > 1:
                # leave NCCL's CTAs their SMs: the persistent kernels size their grids for the rest (include/lasr.h)
                nccl_ctas = int(os.environ.get("NCCL_MAX_CTAS", "8"))
                _lib.load().lasr_set_sm_budget(max(64, 148 - int(os.environ.get("LASR_SM_RESERVE", str(nccl_ctas)))))
            from . import ddp
            tail = world_sync[2] if len(world_sync) > 2 else (0.75, 2.5, 4.0)
            self.grad_sync = ddp.GradSync(module, group=world_sync[0], bucket_mb=world_sync[1], overlap=True,
                                          bank=self.bank, tail_mb=tail)
        self.fused = fused
        self.waveform = isinstance(example_batch[0], (tuple, list))
        self.frontend = None
        if self.waveform:
            from . import frontend as fe
            waves, ns = example_batch[0]
            self._wave_dtype = wave_dtype
            w = fe.pcm16(waves) if (wave_dtype == torch.int16 and waves.dtype != torch.int16) else waves.to(wave_dtype)
            example_batch = (w.contiguous(), example_batch[1], torch.as_tensor(ns).int(), example_batch[3])
            from .quartznet import resolve_dtype
            self.bank.count_steps = self.bank.count_steps or augment or dither
            self.frontend = fe.DeviceFrontend(w.shape[0], w.shape[1], self.dev, resolve_dtype(module.encoder.precision),
                                              augment=augment, dither=dither, seed_dev=self.bank.step_counter)
            self.uniforms = None  # optional [N, 6] fp64 device tensor: the augmentation parity hook
        # TWO pinned host staging sets: the H2D copies are asynchronous, so the host may only overwrite a set once the
        # copy that last read it has completed (event per set); with two sets the host can stage batch i+1 while the
        # H2D of batch i is still in flight (the two-deep pipeline of step_host(prefetch_next=..., defer_loss=True))
        # every set (pinned host x 2, device staging, device static) is ONE flat buffer with the batch's tensors as
        # 256-byte-aligned views: a batch crosses PCIe as one copy and is taken over with one device-to-device copy
        # (four small tensors as four copies each cost ~20 us of serial launches per step between graph replays)
        proto = [t.contiguous() if torch.is_tensor(t) else t for t in example_batch[:4]]
        self._host_flats, self._host_sets = [], []
        for _ in range(2):
            flat, views = _flat_views(proto, "cpu", pin=True)
            for v, t in zip(views, proto):
                if torch.is_tensor(t):
                    v.copy_(t)
            self._host_flats.append(flat)
            self._host_sets.append(views)
        self._host_idx = 0
        self._h2d_evt = [None, None]
        self.static_flat, self.static = _flat_views(proto, self.dev)
        self.static_flat.copy_(self._host_flats[0], non_blocking=True)
        self.h2d_bytes = self._host_flats[0].numel()  # what one step copies (tensor bytes + < 1 KB of alignment)
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        self._loss_ring = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._loss_evt = [None, None]
        self._host_steps = 0
        self.loss_dev = torch.zeros((), device=self.dev, dtype=torch.float32)
        self._gout = None  # constant seed of the backward: d mean(nll) / d nll
        self.graph = None
        self.use_graph = graph
        # double-buffered input pipeline: the NEXT batch's H2D runs on a copy stream while this step computes
        self.copy_stream = torch.cuda.Stream()
        self.staging_flat, self.staging = _flat_views(proto, self.dev)
        self._staged_evt = None
        self._consumed_evt = None

    @property
    def host(self):
        """The pinned staging set holding the most recently staged batch."""
        return self._host_sets[self._host_idx]

    def _stage_host(self, batch):
        """Copy a new batch into the OTHER pinned set (waiting until the H2D that last read it has finished)."""
        i = self._host_idx ^ 1
        if self._h2d_evt[i] is not None:
            self._h2d_evt[i].synchronize()
        if self.waveform and isinstance(batch[0], (tuple, list)):  # ((waves, num_samples), targets, _, target_sizes)
            from . import frontend as fe
            w = batch[0][0]
            w = fe.pcm16(w) if (self._wave_dtype == torch.int16 and w.dtype != torch.int16) else w.to(self._wave_dtype)
            batch = (w, batch[1], torch.as_tensor(batch[0][1]).int(), batch[3])
        for h, src in zip(self._host_sets[i], batch[:4]):
            if torch.is_tensor(src):
                if src.shape != h.shape:
                    raise ValueError(f"TrainEngine batches must keep their shape: {tuple(src.shape)} vs {tuple(h.shape)}")
                h.copy_(src)
        self._host_idx = i

    def _h2d(self, dst_flat, stream):
        """Queue the H2D of the current pinned set into `dst_flat` on `stream` and remember when it completes."""
        dst_flat.copy_(self._host_flats[self._host_idx], non_blocking=True)
        evt = torch.cuda.Event()
        evt.record(stream)
        self._h2d_evt[self._host_idx] = evt

    def _step_eager(self):
        m = self.module
        self.bank.begin_step()
        if self.grad_sync is not None:
            self.grad_sync.zero_and_attach()
        if self.waveform:
            # static = [waves, targets, num_samples, target_sizes]: frontend -> encoder -> fused decoder + CTC
            feats, percents = self.frontend(self.static[0], self.static[2], self.uniforms)
            x = m.encoder.encoder.forward_ntc(feats, percents)
            nll, _, _ = m.encoder.fused_ctc_from_encoded(x, percents, self.static[1], self.static[3])
        elif self.fused:
            nll, _, _ = m.encoder.forward_fused_ctc(self.static[0], self.static[2], self.static[1], self.static[3])
        else:
            nll = None
            loss, _, _ = m.training_step_tensors(self.static)
        if nll is not None:
            # loss = mean(nll) (train.py:64-86): the backward is seeded with the constant d loss / d nll = 1/N and the mean
            # is written straight into loss_dev -- autograd's own route (ones -> mul -> expand -> contiguous copy, then a
            # copy of the loss) is four more dependent launches between the CTC lattices and the CTC gradient pass
            if self._gout is None or self._gout.shape != nll.shape:
                self._gout = torch.full(nll.shape, 1.0 / nll.numel(), device=nll.device, dtype=torch.float32)
            nll.backward(gradient=self._gout)
            loss = None
        else:
            loss.backward()
        self.bank.join_side()  # deferred weight-gradient kernels (runtime.defer) are part of this step
        if self.grad_sync is not None:
            self.grad_sync(m)
        if self.optimizer is not None:
            self.optimizer.step()  # optim.Novograd: three launches over the flat buffers, arena still armed
        self.bank.end_step()
        if loss is None:
            torch.mean(nll.detach(), dim=0, out=self.loss_dev)
        else:
            self.loss_dev.copy_(loss.detach())

    def _capture(self):
        # the step's main chain is captured from a HIGH-priority stream; the deferred weight-gradient kernels
        # (runtime.defer) sit on the bank's default-priority side stream and only fill SMs the chain leaves idle
        s = torch.cuda.Stream(priority=-1)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._step_eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            self._step_eager()
        torch.cuda.synchronize()
        self.graph = graph

    def step_device(self):
        """One step on the resident batch; returns the device loss scalar (no sync)."""
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._step_eager()
        return self.loss_dev

    def load_batch(self, batch):
        """Pinned-host staging + async H2D of a new batch of the SAME shapes."""
        self._stage_host(batch)
        self._h2d(self.static_flat, torch.cuda.current_stream())

    def prefetch(self, batch=None):
        """Start the H2D of the NEXT batch (pinned host -> device staging buffers) on the copy stream; it overlaps
        whatever the compute stream is doing.  `batch=None` re-sends the pinned example batch (bench.py)."""
        if batch is not None:
            self._stage_host(batch)
        if self._consumed_evt is not None:
            self.copy_stream.wait_event(self._consumed_evt)  # the previous staging contents have been taken over
        with torch.cuda.stream(self.copy_stream):
            self._h2d(self.staging_flat, self.copy_stream)
            self._staged_evt = self._h2d_evt[self._host_idx]

    def step_host(self, batch=None, prefetch_next=False, defer_loss=False):
        """End-to-end step: H2D of the batch from pinned memory, the step, D2H of the loss.  Returns a float.
        If prefetch() staged this batch earlier, the inputs are taken over with a device-to-device copy instead of
        waiting for PCIe; `prefetch_next` (True: the pinned example batch, or a batch tuple) starts the next batch's
        H2D before this step's loss is awaited, so the copy hides under the step.
        defer_loss=True keeps the host one step ahead of the device: the call enqueues this step (and its loss D2H)
        and returns the PREVIOUS step's loss (None on the first call; flush_loss() returns the last one), so the next
        graph launch is already queued when the device finishes -- what a training loop that logs the loss does."""
        cur = torch.cuda.current_stream()
        if batch is None and self._staged_evt is not None:
            cur.wait_event(self._staged_evt)
            self.static_flat.copy_(self.staging_flat, non_blocking=True)
            self._consumed_evt = torch.cuda.Event()
            self._consumed_evt.record(cur)
            self._staged_evt = None
        elif batch is not None:
            self.load_batch(batch)
        else:
            self._h2d(self.static_flat, cur)
        self.step_device()
        if prefetch_next is not False and prefetch_next is not None:
            self.prefetch(None if prefetch_next is True else prefetch_next)
        if not defer_loss:
            self.loss_host.copy_(self.loss_dev, non_blocking=True)
            cur.synchronize()
            return float(self.loss_host)
        slot = self._host_steps & 1
        self._loss_ring[slot].copy_(self.loss_dev, non_blocking=True)
        self._loss_evt[slot] = torch.cuda.Event()
        self._loss_evt[slot].record(cur)
        self._host_steps += 1
        prev = slot ^ 1
        if self._loss_evt[prev] is None:
            return None
        self._loss_evt[prev].synchronize()
        return float(self._loss_ring[prev])

    def close(self):
        """Release the step runtime (the parameters stay valid views of this engine's flat master buffer)."""
        if runtime.current() is self.bank:
            runtime.uninstall()

    def flush_loss(self):
        """Loss of the most recent deferred step (waits for it)."""
        slot = (self._host_steps - 1) & 1
        if self._host_steps == 0 or self._loss_evt[slot] is None:
            return None
        self._loss_evt[slot].synchronize()
        return float(self._loss_ring[slot])

class InferEngine:
    """Validation / inference path on one GPU (BASELINE config 5; predict.py:43-62, train.py:88-116 without the loss):
    raw waveforms -> log-mel frontend (csrc/logmel.cu) -> encoder in eval mode -> decoder logits -> greedy CTC decode
    (argmax + collapse, utils/asr_metrics.py:153-171).  Everything between the H2D of the waveforms and the D2H of the
    token ids runs on the device with static buffers, so the whole pass is ONE CUDA graph.

      step_device()             re-runs the pass on the resident waveforms -> (tokens [N, T'] int32, counts [N] int32)
      step_host(prefetch_next)  pinned waveforms -> H2D (double buffered, like TrainEngine) -> pass -> tokens on host
      transcripts()             strings through WER.decode_tokens (host side, like the reference)
    """

    def __init__(self, module, waves, num_samples, graph=True, wave_dtype=None):
        """waves [N, S_max]: 16-bit PCM (torch.int16: the samples as they sit in the file; x / 32768 on the device like
        torchaudio.load(normalize=True), predict.py:46) or fp32 in [-1, 1].  wave_dtype=None keeps the dtype given,
        torch.int16 converts float input once (frontend.pcm16): a quarter of the H2D bytes of fp32 samples -- at
        256 x 30 s that is 245 MB instead of 491 MB per pass, the difference between a copy that hides under the pass
        and one that does not."""
        from . import frontend, ops
        _lib.require_device()
        self.module = module.eval()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.frontend, self.ops = frontend, ops
        if wave_dtype is None:
            wave_dtype = torch.int16 if waves.dtype == torch.int16 else torch.float32
        if wave_dtype not in (torch.int16, torch.float32):
            raise _lib.LasrError("InferEngine: wave_dtype must be torch.int16 or torch.float32")
        self._wave_dtype = wave_dtype
        waves = self._wire(waves)
        self.N, self.S = waves.shape
        ns = torch.as_tensor(num_samples, dtype=torch.int64)
        if int(ns.max()) > self.S or int(ns.min()) <= frontend.N_FFT // 2:
            raise _lib.LasrError("num_samples must be in (256, S_max]")
        self.T = frontend.num_frames(int(ns.max()))
        self.ns = ns.to(self.dev, dtype=torch.int32)
        frames = frontend.num_frames(ns)
        self.percents = (frames.float() / float(self.T)).to(self.dev)  # data_module.py:244
        self.host = waves.pin_memory()
        self.static = self.host.to(self.dev, non_blocking=True)
        self.staging = torch.empty_like(self.static)
        self.h2d_bytes = self.host.numel() * self.host.element_size()
        from .quartznet import resolve_dtype
        self.dtype = resolve_dtype(module.encoder.precision)
        self.basis, self.mel_idx, self.mel_w = frontend.constants(self.dev)
        Lp = _lib.load().lasr_logmel_padded_len(self.T)
        self.parts = torch.empty((3, self.N, Lp), device=self.dev, dtype=torch.bfloat16)
        self.db = torch.empty((self.N, self.T, frontend.N_MELS), device=self.dev, dtype=torch.float32)
        self.stats = torch.zeros((self.N, 2), device=self.dev, dtype=torch.float64)
        self.feats = torch.empty((self.N, self.T, frontend.N_MELS), device=self.dev, dtype=self.dtype)
        self.tokens = self.counts = None
        self.tokens_host = self.counts_host = None
        self.copy_stream = torch.cuda.Stream()
        self._staged_evt = self._consumed_evt = None
        self.use_graph, self.graph = graph, None
        # weights do not change during inference: ONE bf16 shadow cast here instead of one per layer and pass.
        # If the module is already owned by the installed bank (validation during training: the TrainEngine's graph and
        # the fused optimizer keep writing THAT bank's masters / shadows) the engine shares it -- re-homing the
        # parameters into a second bank would silently freeze both state_dict() and this engine's weights.
        cur = runtime.current()
        params = list(module.parameters())
        if cur is not None and all(id(p) in cur.offsets for p in params):
            self.bank, self._owns_bank = cur, False
        else:
            self.bank, self._owns_bank = runtime.ParamBank(module), True  # raises if another installed bank owns them
            runtime.install(self.bank)
        self.refresh_weights()

    def _wire(self, waves):
        """The batch in its wire format (contiguous, CPU or CUDA)."""
        if self._wave_dtype == torch.int16:
            return (waves if waves.dtype == torch.int16 else self.frontend.pcm16(waves)).contiguous()
        if waves.dtype == torch.int16:
            return (waves.float() / 32768.0).contiguous()
        return waves.float().contiguous()

    def refresh_weights(self):
        """Re-cast the bf16 weight shadows (call after load_state_dict / any parameter update; with a bank shared with a
        TrainEngine, call it before validating so the shadows reflect the latest optimizer step)."""
        b = self.bank
        _lib.call("lasr_cast_weight", b.master, b.shadow, 1, b.numel, 0, _lib.LASR_BF16)
        b.shadow_fresh = True

    def close(self):
        """Release the step runtime if this engine installed it."""
        if self._owns_bank and runtime.current() is self.bank:
            runtime.uninstall()

    def _run(self):
        fe, ops = self.frontend, self.ops
        model = self.module.encoder  # MyModel2
        with torch.no_grad():
            self.stats.zero_()
            _lib.call("lasr_logmel_prepare_wave", self.static, 1 if self._wave_dtype == torch.int16 else 0, None, 0, None,
                      None, self.ns, self.parts, self.N, self.S, self.T)
            _lib.call("lasr_logmel_fwd", self.parts, self.basis, self.mel_idx, self.mel_w, self.ns, self.db, self.stats,
                      self.N, self.T, 6)
            _lib.call("lasr_logmel_normalize", self.db, self.stats, self.ns, None, self.feats, self.N, self.T,
                      _lib.dtype_code(self.dtype))
            x = model.encoder.forward_ntc(self.feats, self.percents)  # [N, T', 1024]
            V = len(model.labels) + 1
            w = runtime.weight(model.decoder.weight, x.dtype).view(V, -1)
            logits = ops.pwconv_fwd(x, w, bias=model.decoder.bias.detach(), ldy=(V + 7) // 8 * 8)
            t_len = ops.out_lengths(x.shape[1], self.percents)  # train.py:76
            _, tokens, counts = ops.greedy_decode(logits, t_len, V, V - 1)
        if self.tokens is None:
            self.tokens, self.counts = torch.empty_like(tokens), torch.empty_like(counts)
            self.tokens_host = torch.empty(tokens.shape, dtype=tokens.dtype).pin_memory()
            self.counts_host = torch.empty(counts.shape, dtype=counts.dtype).pin_memory()
        self.tokens.copy_(tokens)
        self.counts.copy_(counts)

    _step_eager = _run  # bench.py's per-kernel breakdown hook

    def step_device(self):
        if self.use_graph:
            if self.graph is None:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(2):
                        self._run()
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run()
                torch.cuda.synchronize()
                self.graph = g
            self.graph.replay()
        else:
            self._run()
        return self.tokens, self.counts

    def prefetch(self, waves=None):
        if waves is not None:
            self.host.copy_(self._wire(waves))
        if self._consumed_evt is not None:
            self.copy_stream.wait_event(self._consumed_evt)
        with torch.cuda.stream(self.copy_stream):
            self.staging.copy_(self.host, non_blocking=True)
            self._staged_evt = torch.cuda.Event()
            self._staged_evt.record(self.copy_stream)

    def step_host(self, prefetch_next=False):
        cur = torch.cuda.current_stream()
        if self._staged_evt is not None:
            cur.wait_event(self._staged_evt)
            self.static.copy_(self.staging, non_blocking=True)
            self._consumed_evt = torch.cuda.Event()
            self._consumed_evt.record(cur)
            self._staged_evt = None
        else:
            self.static.copy_(self.host, non_blocking=True)
        self.step_device()
        if prefetch_next is not False and prefetch_next is not None:
            self.prefetch(None if prefetch_next is True else prefetch_next)
        self.tokens_host.copy_(self.tokens, non_blocking=True)
        self.counts_host.copy_(self.counts, non_blocking=True)
        cur.synchronize()
        return self.tokens_host, self.counts_host

    def transcripts(self):
        """Collapsed token ids -> strings (labels_map join, utils/asr_metrics.py:168)."""
        labels = self.module.labels
        toks, cnts = self.tokens.cpu(), self.counts.cpu()
        return ["".join(labels[int(c)] for c in toks[i, :int(cnts[i])]) for i in range(toks.shape[0])]
