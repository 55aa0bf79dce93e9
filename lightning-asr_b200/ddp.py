"""Data-parallel gradient exchange: the one collective on the path.

The reference gets it implicitly from pytorch_lightning's `accelerator: ddp` (conf/conf.yaml:30, train.py:239), i.e.
torch DistributedDataParallel: SUM all-reduce of every gradient, averaged over ranks, BatchNorm statistics stay
per-rank (no SyncBN, train.py:233-251).  Here: one process per GPU, parameters' gradients live in a few flat fp32
buckets ordered by backward completion (decoder first, first_cnn last); each bucket is all-reduced with NCCL on a
side stream as soon as it is complete, while the backward of earlier layers continues on the compute stream; the
optimizer (or the caller) waits on the side stream.  `world_size == 1` is a no-op.

The 1/world_size averaging is folded into the all-reduce (ReduceOp.AVG on NCCL; SUM + scale on gloo for the CPU tests).
"""
import torch
import torch.distributed as dist


def broadcast_parameters(module, src=0, group=None):
    """DDP's construction-time broadcast: every rank starts from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)


class GradSync:
    """Flat-bucket gradient all-reduce.

    With a runtime.ParamBank (the step engine's mode) the buckets are contiguous slices of the bank's flat gradient
    buffer -- the wgrad kernels have already accumulated into them, nothing is copied -- and a bucket is launched the
    moment the last of its parameters is reported by runtime.grad_ready().  Without a bank (generic autograd usage,
    the CPU tests) gradients are attached as views of private flat buckets and post-accumulate hooks do the
    reporting.  Call the object after backward to drain: it reduces whatever has not fired and joins the side stream.
    """

    def __init__(self, module, group=None, bucket_mb=8.0, overlap=True, bank=None, tail_mb=(0.75, 2.5, 4.0)):
        """bucket_mb: size of the buckets that complete early in backward (decoder side).  tail_mb: sizes of the LAST
        buckets to complete, innermost first -- the final bucket (first_cnn + the first block) finishes with the very last
        weight-gradient kernel and nothing but the optimizer is left to hide its all-reduce behind, so it is kept small
        (latency-bound, ~25 us) and the ones before it grow geometrically; () restores uniform buckets."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bank = bank
        self.params = [p for p in module.parameters() if p.requires_grad]
        # walk the parameters in FORWARD order (= reverse completion order): the tail buckets first, then uniform ones
        cap_rest = int(bucket_mb * 1024 * 1024 / 4)
        caps = [min(int(mb * 1024 * 1024 / 4), cap_rest) for mb in tail_mb]
        groups, cur, cur_n = [], [], 0
        for p in self.params:
            cap = caps[len(groups)] if len(groups) < len(caps) else cap_rest
            if cur and cur_n + p.numel() > cap:
                groups.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            groups.append(cur)
        # buckets in backward completion order (decoder / last layers first), parameters inside likewise
        groups = [list(reversed(g)) for g in reversed(groups)]
        self.buckets = []  # (flat fp32 tensor, [(param, offset, numel)])
        for g in groups:
            if bank is not None:
                lo = min(bank.offsets[id(p)] for p in g)
                hi = max(bank.offsets[id(p)] + p.numel() for p in g)
                flat = bank.grads[lo:hi]
                items = [(p, bank.offsets[id(p)] - lo, p.numel()) for p in g]
            else:
                flat = torch.zeros(sum(p.numel() for p in g), device=g[0].device, dtype=torch.float32)
                items, off = [], 0
                for p in g:
                    items.append((p, off, p.numel()))
                    off += p.numel()
            self.buckets.append((flat, items))
        self.enabled = True  # False: skip the collective itself (bench.py measures how much of it is exposed)
        self.overlap = overlap and self.world > 1
        self.comm_stream = None
        if self.overlap and self.params and self.params[0].is_cuda:
            self.comm_stream = torch.cuda.Stream()
        self._bucket_of = {}
        for bi, (_, items) in enumerate(self.buckets):
            for p, _, _ in items:
                self._bucket_of[id(p)] = bi
        self._left = [len(items) for _, items in self.buckets]
        self._fired = [False] * len(self.buckets)
        if bank is not None:
            bank.on_grad_ready = self._ready if self.overlap else None
        elif self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._ready)

    # no-bank mode: gradients are views of the private flat buckets, so backward's accumulation writes into them
    def attach_grad_views(self):
        if self.bank is not None:
            return
        for flat, items in self.buckets:
            for p, off, n in items:
                p.grad = flat[off:off + n].view_as(p)

    def zero_and_attach(self):
        """Start of a step (no-bank mode zeroes its own buckets; the bank zeroes itself in begin_step)."""
        if self.bank is None:
            for flat, _ in self.buckets:
                flat.zero_()
            self.attach_grad_views()
        self._left = [len(items) for _, items in self.buckets]
        self._fired = [False] * len(self.buckets)

    def _reduce_bucket(self, bi):
        flat, items = self.buckets[bi]
        views = self.bank is not None or all(
            p.grad is not None and p.grad.data_ptr() == flat[off:off + n].data_ptr() for p, off, n in items)
        if not views:
            for p, off, n in items:
                if p.grad is None:
                    flat[off:off + n].zero_()
                else:
                    flat[off:off + n].copy_(p.grad.reshape(-1))
        if self.world > 1 and self.enabled:
            if flat.is_cuda:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(self.world)
        if not views:
            for p, off, n in items:
                if p.grad is not None:
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))

    def _launch(self, bi):
        self._fired[bi] = True
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            if self.bank is not None and getattr(self.bank, "side", None) is not None:
                self.comm_stream.wait_stream(self.bank.side)  # weight gradients are produced on the side stream
            with torch.cuda.stream(self.comm_stream):
                self._reduce_bucket(bi)
        else:
            self._reduce_bucket(bi)

    def _ready(self, p):
        bi = self._bucket_of.get(id(p))
        if bi is None or self._fired[bi]:
            return
        self._left[bi] -= 1
        if self._left[bi] == 0:
            self._launch(bi)

    def __call__(self, module=None):
        """Finish the exchange: reduce every bucket that has not fired, then join the side stream."""
        if self.world == 1:
            return
        for bi in range(len(self.buckets)):
            if not self._fired[bi]:
                self._launch(bi)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
