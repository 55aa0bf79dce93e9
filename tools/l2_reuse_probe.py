"""Does the L2 carry a BatchNorm backward's operands from the reduce pass to the apply pass?  Times (reduce; apply)
back to back against the passes alone (an L2-sized flush before every timed region)."""
import statistics
import sys
import torch
sys.path.insert(0, ".")
from lightning_asr_b200 import _lib, ops
_lib.require_device()
dev = "cuda"
N, T = 32, 801
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, iters=9):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts[2:])


for c in (256, 512):
    y = torch.randn(N, T, c, device=dev).bfloat16()
    r = torch.randn(N, T, c, device=dev).bfloat16()
    g = torch.ones(c, device=dev)
    b = torch.zeros(c, device=dev)
    rm, rv, nbt = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.zeros((), device=dev, dtype=torch.long)
    sums = lambda t: torch.stack([t.double().reshape(-1, c).sum(0), (t.double().reshape(-1, c) ** 2).sum(0)])
    bn1 = ops.BNForward(g, b, rm, rv, nbt, sums(y))
    bn2 = ops.BNForward(g, b, rm.clone(), rv.clone(), nbt.clone(), sums(r))
    bits = ops.relu_bits_alloc(N, T, c, dev)
    out = ops.bn_apply_act(y, bn1, r, bn2, relu_bits=bits)
    dout = torch.randn(N, T, c, device=dev).bfloat16()
    totals = torch.zeros(3, c, device=dev, dtype=torch.float64)
    lengths = torch.full((N,), T, device=dev, dtype=torch.int32)
    dg = torch.zeros(4, c, device=dev)
    red = lambda: ops.bn_act_bwd_reduce(dout, None, y, r, ops.ACT_RELU, totals, relu_bits=bits)
    app = lambda: ops.bn_act_bwd_apply(dout, None, y, r, None, None, totals, None, (g, bn1.save, dg[0], dg[1]),
                                       (g, bn2.save, dg[2], dg[3]), lengths, ops.ACT_RELU, relu_bits=bits)
    t_r, t_a = timeit(red), timeit(app)
    t_ra = timeit(lambda: (red(), app()))
    t_rr = timeit(lambda: (red(), red()))
    print(f"C={c}: reduce {t_r:.1f} us, apply {t_a:.1f} us, reduce;apply {t_ra:.1f} us (apply after reduce: {t_ra - t_r:.1f}), "
          f"reduce;reduce {t_rr:.1f} us (second reduce: {t_rr - t_r:.1f})")
