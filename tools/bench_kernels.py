"""Per-kernel timings at the asr13x1 config-2 shapes (N=32, T'=801, bf16): CUDA events on the launching stream, an
L2-sized scratch write between iterations (cold L2), median of `iters`.  Prints us, achieved GB/s (algorithmic bytes)
and TFLOP/s per kernel.   python tools/bench_kernels.py [gemm|dw|dwcm|bn|ctc|all]"""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from lightning_asr_b200 import _lib, ops  # noqa: E402

_lib.require_device()
dev = "cuda"
N, T = 32, 801
M = N * T
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, iters=7):
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


def report(name, us, nbytes, flops):
    print(f"{name:44s} {us:8.1f} us  {nbytes / us / 1e3:8.1f} GB/s  {flops / us / 1e6:8.1f} TFLOP/s", flush=True)


def bench_calib():
    one = torch.zeros(64, device=dev)
    out = torch.zeros(64, device=dev, dtype=torch.bfloat16)
    report("calibration: 64-element cast kernel", timeit(lambda: _lib.call("lasr_cast_weight", one, out, 1, 64, 0, 1)), 0, 0)
    for cin, cout in [(256, 256), (512, 512)]:
        x = torch.randn(M, cin, device=dev).bfloat16()
        w = (torch.randn(cout, cin, device=dev) / cin ** 0.5).bfloat16()
        y = torch.empty(M, cout, device=dev, dtype=torch.bfloat16)
        report(f"torch.matmul (cuBLAS) {cin}->{cout}", timeit(lambda: torch.matmul(x, w.t(), out=y)), 2 * M * (cin + cout), 2.0 * M * cin * cout)
    x = torch.randn(M, 512, device=dev).bfloat16()
    y = torch.empty_like(x)
    report("torch copy_ [M,512] bf16", timeit(lambda: y.copy_(x)), 2 * M * 512 * 2, 0)


def bench_gemm():
    for cin, cout in [(64, 256), (256, 256), (256, 512), (512, 512), (512, 1024), (1024, 29)]:
        x = torch.randn(M, cin, device=dev).bfloat16()
        w = (torch.randn(cout, cin, device=dev) / cin ** 0.5).bfloat16()
        lengths = torch.full((N,), T, device=dev, dtype=torch.int32)
        stats = torch.zeros(2, cout, device=dev, dtype=torch.float64)
        ld = (cout + 7) // 8 * 8
        y = torch.empty(M, ld, device=dev, dtype=torch.bfloat16)
        by, fl = 2 * M * (cin + cout), 2.0 * M * cin * cout
        report(f"pwconv_fwd+mask+stats {cin}->{cout}", timeit(lambda: ops.pwconv_fwd(x, w, lengths=lengths, T=T, stats=stats, out=y, ldy=ld)), by, fl)
        report(f"pwconv_fwd plain      {cin}->{cout}", timeit(lambda: ops.pwconv_fwd(x, w, out=y, ldy=ld)), by, fl)
        dy = torch.randn(M, ld, device=dev).bfloat16()
        report(f"pwconv_dgrad          {cin}->{cout}", timeit(lambda: ops.pwconv_dgrad(dy, w, lddy=ld)), by, fl)
        dw = torch.zeros(cout, cin, device=dev)
        report(f"pwconv_wgrad          {cin}->{cout}", timeit(lambda: ops.pwconv_wgrad(dy, x, out=dw, Cout=cout)), by + 4 * cin * cout, fl)


def bench_dw():
    for c, k in [(256, 33), (256, 39), (256, 51), (512, 51), (512, 63), (512, 75)]:
        x = torch.randn(N, T, c, device=dev).bfloat16()
        w = torch.randn(c, 1, k, device=dev) / k ** 0.5
        by, fl = 2 * M * c * 2, 2.0 * M * c * k
        report(f"dwconv_fwd   C={c} k={k}", timeit(lambda: ops.dwconv_fwd(x, w)), by, fl)
        add = torch.randn(N, T, c, device=dev).bfloat16()
        report(f"dwconv_dgrad+addend C={c} k={k}", timeit(lambda: ops.dwconv_fwd(x, w, flip=True, addend=add)), by + 2 * M * c, fl)
        dw = torch.zeros(c, 1, k, device=dev)
        report(f"dwconv_wgrad C={c} k={k}", timeit(lambda: ops.dwconv_wgrad(x, add, k, out=dw)), by, fl)


def bench_dwcm():
    """TMA-fed depthwise kernels reading channel-major series, and the BatchNorm pass that writes them"""
    for c, k in [(256, 33), (256, 39), (256, 51), (512, 51), (512, 63), (512, 75)]:
        x = torch.randn(N, T, c, device=dev).bfloat16()
        w = torch.randn(c, 1, k, device=dev) / k ** 0.5
        xs = ops.series_from_ntc(x, k)
        by, fl = 2 * M * c * 2, 2.0 * M * c * k
        report(f"dwconv_fwd_cm   C={c} k={k}", timeit(lambda: ops.dwconv_fwd_cm(xs, w)), by, fl)
        add = torch.randn(N, T, c, device=dev).bfloat16()
        report(f"dwconv_dgrad_cm+addend C={c} k={k}", timeit(lambda: ops.dwconv_fwd_cm(xs, w, flip=True, addend=add)), by + 2 * M * c, fl)
        adds = ops.series_from_ntc(add, k)
        report(f"dwconv_dgrad_cm+series addend C={c} k={k}", timeit(lambda: ops.dwconv_fwd_cm(xs, w, flip=True, addend=adds)), by + 2 * M * c, fl)
        dw = torch.zeros(c, 1, k, device=dev)
        report(f"dwconv_wgrad_cm C={c} k={k}", timeit(lambda: ops.dwconv_wgrad_cm(xs, adds, k, out=dw)), by, fl)
        report(f"dwconv_bwd_cm (dgrad+series addend, wgrad) C={c} k={k}", timeit(lambda: ops.dwconv_bwd_cm(xs, adds, w, addend=adds, out_dw=dw)), 2 * M * c * 4, 2 * fl)
    for cin, cout, k in [(256, 256, 33), (256, 512, 51), (512, 512, 63)]:
        dy = torch.randn(N, T, cout, device=dev).bfloat16()
        dr = torch.randn(N, T, cout, device=dev).bfloat16()
        w = (torch.randn(cout, cin, device=dev) / cin ** 0.5).bfloat16()
        by, fl = 2 * M * (cin + cout), 2.0 * M * cin * cout
        report(f"pwconv_dgrad2 (grouped) {cin}<-{cout}", timeit(lambda: ops.pwconv_dgrad2(dy, w, dr, w)), 2 * by, 2 * fl)
        report(f"pwconv_dgrad_cm (grouped, series out) {cin}<-{cout}", timeit(lambda: ops.pwconv_dgrad_cm(dy, w, k, dr, w)), 2 * by, 2 * fl)
    for c in (256, 512):
        y = torch.randn(N, T, c, device=dev).bfloat16()
        r = torch.randn(N, T, c, device=dev).bfloat16()
        g = torch.ones(c, device=dev)
        b = torch.zeros(c, device=dev)
        rm, rv, nbt = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.zeros((), device=dev, dtype=torch.long)

        def sums(t):
            t2 = t.double().reshape(-1, c)
            return torch.stack([t2.sum(0), (t2 * t2).sum(0)])
        bn1 = ops.BNForward(g, b, rm, rv, nbt, sums(y))
        bn2 = ops.BNForward(g, b, rm.clone(), rv.clone(), nbt.clone(), sums(r))
        report(f"bn_apply_act_fwd (res) C={c}", timeit(lambda: ops.bn_apply_act(y, bn1, r, bn2)), 2 * M * c * 3, 0)
        report(f"bn_apply_act_fwd_cm (res) C={c}", timeit(lambda: ops.bn_apply_act(y, bn1, r, bn2, cm_k=51)), 2 * M * c * 4, 0)
    for c in (256, 512):  # inference shape (config 5): everything streams through DRAM
        Nb, Tb = 256, 1501
        y = torch.randn(Nb, Tb, c, device=dev).bfloat16()
        r = torch.randn(Nb, Tb, c, device=dev).bfloat16()
        g = torch.ones(c, device=dev)
        b = torch.zeros(c, device=dev)
        rm, rv, nbt = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.zeros((), device=dev, dtype=torch.long)
        bn1 = ops.BNForward(g, b, rm, rv, nbt, None)
        bn2 = ops.BNForward(g, b, rm.clone(), rv.clone(), nbt.clone(), None)
        report(f"bn_apply_act_fwd eval b256x30s C={c}", timeit(lambda: ops.bn_apply_act(y, bn1, r, bn2)), 2 * Nb * Tb * c * 3, 0)
        report(f"bn_apply_act_fwd_cm eval b256x30s C={c}", timeit(lambda: ops.bn_apply_act(y, bn1, r, bn2, cm_k=51)), 2 * Nb * Tb * c * 4, 0)


def bench_bn():
    for c in (256, 512):
        y = torch.randn(N, T, c, device=dev).bfloat16()
        r = torch.randn(N, T, c, device=dev).bfloat16()
        g = torch.ones(c, device=dev)
        b = torch.zeros(c, device=dev)
        rm, rv, nbt = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.zeros((), device=dev, dtype=torch.long)

        def sums(t):
            t2 = t.double().reshape(-1, c)
            return torch.stack([t2.sum(0), (t2 * t2).sum(0)])
        s1, s2 = sums(y), sums(r)
        bn1 = ops.BNForward(g, b, rm, rv, nbt, s1)
        bn2 = ops.BNForward(g, b, rm.clone(), rv.clone(), nbt.clone(), s2)
        out = ops.bn_apply_act(y, bn1, r, bn2)
        report(f"bn_apply_act_fwd (res) C={c}", timeit(lambda: ops.bn_apply_act(y, bn1, r, bn2)), 2 * M * c * 3, 0)
        dout = torch.randn(N, T, c, device=dev).bfloat16()
        totals = torch.zeros(3, c, device=dev, dtype=torch.float64)
        report(f"bn_act_bwd_reduce (res) C={c}", timeit(lambda: ops.bn_act_bwd_reduce(dout, out, y, r, ops.ACT_RELU, totals)), 2 * M * c * 4, 0)
        bits = ops.relu_bits_alloc(N, T, c, dev)
        ops.bn_apply_act(y, bn1, r, bn2, relu_bits=bits)
        report(f"bn_act_bwd_reduce (res, relu bits) C={c}", timeit(lambda: ops.bn_act_bwd_reduce(dout, None, y, r, ops.ACT_RELU, totals, relu_bits=bits)), 2 * M * c * 3, 0)
        lengths = torch.full((N,), T, device=dev, dtype=torch.int32)
        dg = torch.zeros(4, c, device=dev)
        report(f"bn_act_bwd_apply (res) C={c}",
               timeit(lambda: ops.bn_act_bwd_apply(dout, out, y, r, None, None, totals, None, (g, bn1.save, dg[0], dg[1]),
                                                   (g, bn2.save, dg[2], dg[3]), lengths, ops.ACT_RELU)), 2 * M * c * 6, 0)
        report(f"bn_act_bwd_apply (res, relu bits) C={c}",
               timeit(lambda: ops.bn_act_bwd_apply(dout, None, y, r, None, None, totals, None, (g, bn1.save, dg[0], dg[1]),
                                                   (g, bn2.save, dg[2], dg[3]), lengths, ops.ACT_RELU, relu_bits=bits)), 2 * M * c * 5, 0)


def bench_ctc():
    V, ld = 29, 32
    S = T // 4
    logits = torch.randn(N, T, ld, device=dev).bfloat16()
    targets = torch.randint(0, 28, (N, S), device=dev)
    il = torch.full((N,), T, device=dev, dtype=torch.int32)
    tl = torch.full((N,), S, device=dev, dtype=torch.int32)
    lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
    report("log_softmax_fwd (lse only) V=29", timeit(lambda: ops.log_softmax_fwd(logits, V, want_lp=False)), 2 * M * ld, 0)
    nll, alpha, beta, scales = ops.ctc_fwd(logits, lse, targets, il, tl, V, 28, want_beta=True)
    report("ctc_fwd (alpha+beta lattices)", timeit(lambda: ops.ctc_fwd(logits, lse, targets, il, tl, V, 28, want_beta=True)),
           8 * M * (2 * S + 1), 0)
    import os
    for mode, what in (("0", "round-1 barrier kernel"), ("1", "warp pipeline, state pair per lane")):
        os.environ["LASR_CTC_WARP"] = mode
        report(f"ctc_fwd LASR_CTC_WARP={mode} {what}",
               timeit(lambda: ops.ctc_fwd(logits, lse, targets, il, tl, V, 28, want_beta=True)), 8 * M * (2 * S + 1), 0)
    os.environ.pop("LASR_CTC_WARP")
    go = torch.full((N,), 1.0 / N, device=dev)
    for mode in ("0", "1"):
        os.environ["LASR_CTC_GRAD_SMALL"] = mode
        report(f"ctc_bwd LASR_CTC_GRAD_SMALL={mode}", timeit(lambda: ops.ctc_bwd(logits, lse, targets, il, tl, alpha, beta, nll, go, V, 28, ld, torch.bfloat16, scales=scales)),
               8 * M * (2 * S + 1) + 4 * M * ld, 0)
    os.environ.pop("LASR_CTC_GRAD_SMALL")
    report("ctc_bwd (fused softmax grad)", timeit(lambda: ops.ctc_bwd(logits, lse, targets, il, tl, alpha, beta, nll, go, V, 28, ld, torch.bfloat16, scales=scales)),
           8 * M * (2 * S + 1) + 4 * M * ld, 0)


def bench_ctc_aishell():
    """config 4: the 4334-class decoder -- CTC gradient pass and the decoder-bias column sum"""
    import os
    V, ld = 4334, 4336
    S = T // 4
    logits = torch.randn(N, T, ld, device=dev).bfloat16()
    targets = torch.randint(0, V - 1, (N, S), device=dev)
    il = torch.full((N,), T, device=dev, dtype=torch.int32)
    tl = torch.full((N,), S, device=dev, dtype=torch.int32)
    lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
    nll, alpha, beta, scales = ops.ctc_fwd(logits, lse, targets, il, tl, V, V - 1, want_beta=True)
    go = torch.full((N,), 1.0 / N, device=dev)
    for mode in ("0", "1"):
        os.environ["LASR_CTC_GRAD_LARGE"] = mode
        report(f"ctc_bwd V=4334 LASR_CTC_GRAD_LARGE={mode}",
               timeit(lambda: ops.ctc_bwd(logits, lse, targets, il, tl, alpha, beta, nll, go, V, V - 1, ld, torch.bfloat16, scales=scales)),
               8 * M * (2 * S + 1) + 4 * M * ld, 0)
    os.environ.pop("LASR_CTC_GRAD_LARGE")
    out = torch.zeros(V, device=dev)
    report("colsum [M, 4336] bf16 (decoder bias gradient)", timeit(lambda: ops.colsum(logits.view(M, ld), V, out=out)), 2 * M * ld, 0)


def bench_lstm():
    """the Context variants' BiLSTM recurrence (config 4: N = 32, T' = 801; config 3: N = 64, T' = 1001)"""
    import os
    H = 40
    for nn, tt in ((32, 801), (64, 1001)):
        pre = (torch.randn(nn, tt, 8 * H, device=dev) * 0.7).bfloat16()
        whh = torch.randn(2, 4 * H, H, device=dev) * 0.2
        lens = torch.full((nn,), tt, device=dev, dtype=torch.int32)
        dout = torch.randn(nn, tt, 2 * H, device=dev).bfloat16()
        for mode in ("1", "2", "3"):
            os.environ["LASR_LSTM_V1"] = mode
            out, gates, cells = ops.bilstm_fwd(pre, whh, lens, H)
            dwhh = torch.zeros_like(whh)
            report(f"bilstm_fwd N={nn} T={tt} LASR_LSTM_V1={mode}", timeit(lambda: ops.bilstm_fwd(pre, whh, lens, H)), 0, 0)
            report(f"bilstm_bwd N={nn} T={tt} LASR_LSTM_V1={mode}", timeit(lambda: ops.bilstm_bwd(dout, out, gates, cells, whh, lens, dwhh, H)), 0, 0)
        os.environ.pop("LASR_LSTM_V1")


def bench_frontend():
    from lightning_asr_b200 import frontend

    S = 16000 * 16
    w = (0.05 * torch.randn(N, S, device=dev)).clamp(-1, 1)
    ns = [S] * N
    frontend.logmel_batch(w, ns, out_dtype=torch.bfloat16)
    Tm = frontend.num_frames(S)
    report("logmel_batch (3 launches, 32 x 16 s)", timeit(lambda: frontend.logmel_batch(w, ns, out_dtype=torch.bfloat16)),
           4 * N * S + 4 * N * 64 * Tm, 2.0 * N * Tm * 320 * 512 * 6)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
for name, fn in [("calib", bench_calib), ("gemm", bench_gemm), ("dw", bench_dw), ("dwcm", bench_dwcm), ("bn", bench_bn), ("ctc", bench_ctc), ("aishell", bench_ctc_aishell), ("lstm", bench_lstm), ("frontend", bench_frontend)]:
    if which in ("all", name):
        fn()
