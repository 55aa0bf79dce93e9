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

    grad_sync = GradSync(module, bucket_mb=8); after `loss.backward()` call `grad_sync(module)` (synchronous variant),
    or install hooks with `overlap=True` so each bucket's all-reduce starts when its last gradient is produced.
    """

    def __init__(self, module, group=None, bucket_mb=8.0, overlap=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in module.parameters() if p.requires_grad]
        # reverse registration order ~ backward completion order (decoder / last layers first)
        order = list(reversed(self.params))
        self.buckets = []  # list of (flat fp32 tensor, [(param, offset, numel)])
        cap = int(bucket_mb * 1024 * 1024 / 4)
        cur, cur_n = [], 0
        for p in order:
            if cur and cur_n + p.numel() > cap:
                self._close(cur, cur_n)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._close(cur, cur_n)
        self.overlap = overlap and self.world > 1
        self.comm_stream = None
        self._pending = {}
        self._events = []
        if self.overlap and self.params and self.params[0].is_cuda:
            self.comm_stream = torch.cuda.Stream()
        self._bucket_of = {}
        for bi, (_, items) in enumerate(self.buckets):
            for p, _, _ in items:
                self._bucket_of[id(p)] = bi
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)

    def _close(self, params, n):
        dev = params[0].device
        flat = torch.zeros(n, device=dev, dtype=torch.float32)
        items, off = [], 0
        for p in params:
            items.append((p, off, p.numel()))
            off += p.numel()
        self.buckets.append((flat, items))

    # gradients are views of the flat buckets, so backward's accumulation writes straight into them (no copy)
    def attach_grad_views(self):
        for flat, items in self.buckets:
            for p, off, n in items:
                p.grad = flat[off:off + n].view_as(p)

    def zero_and_attach(self):
        for flat, _ in self.buckets:
            flat.zero_()
        self.attach_grad_views()
        self._pending = {}

    def _reduce_bucket(self, bi):
        flat, items = self.buckets[bi]
        views = all(p.grad is not None and p.grad.data_ptr() == flat[off:off + n].data_ptr() for p, off, n in items)
        if not views:
            for p, off, n in items:
                if p.grad is None:
                    flat[off:off + n].zero_()
                else:
                    flat[off:off + n].copy_(p.grad.reshape(-1))
        if self.world > 1:
            if flat.is_cuda:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(self.world)
        if not views:
            for p, off, n in items:
                if p.grad is not None:
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))

    def _hook(self, p):
        bi = self._bucket_of[id(p)]
        left = self._pending.get(bi)
        if left is None:
            left = len(self.buckets[bi][1])
        left -= 1
        self._pending[bi] = left
        if left == 0:
            self._pending[bi] = None
            if self.comm_stream is not None:
                self.comm_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.comm_stream):
                    self._reduce_bucket(bi)
            else:
                self._reduce_bucket(bi)

    def __call__(self, module=None):
        """Finish the exchange: in overlap mode wait for the side stream; otherwise reduce every bucket now."""
        if self.world == 1:
            return
        if self.overlap:
            # buckets whose parameters received no gradient this step never fired: reduce them now
            for bi, left in list(self._pending.items()):
                if left is not None:
                    self._pending[bi] = None
                    self._reduce_bucket(bi)
            self._pending = {}
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            for bi in range(len(self.buckets)):
                self._reduce_bucket(bi)
