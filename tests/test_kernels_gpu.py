"""Per-kernel parity (through the C ABI) against plain torch fp32/fp64 restatements of the same op.
Tolerances: fp32 path rel 1e-4 (north_star), bf16 path rel 1e-2."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

DTYPES = [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)]


@pytest.fixture(scope="module")
def ops():
    from lightning_asr_b200 import _lib, ops
    _lib.require_device()
    return ops


def ntc(x):  # [N, C, T] -> [N, T, C]
    return x.transpose(1, 2).contiguous()


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("C,K,stride,T", [(64, 33, 2, 401), (256, 33, 1, 501), (256, 39, 1, 130), (336, 51, 1, 257),
                                           (512, 63, 1, 300), (512, 75, 1, 129), (512, 87, 1, 64), (256, 33, 1, 7),
                                           (256, 33, 1, 1501), (512, 63, 1, 2100), (64, 87, 1, 1024)])
def test_dwconv_fwd_wgrad_dgrad(ops, dtype, tol, C, K, stride, T):
    torch.manual_seed(C + K)
    N = 3
    x = torch.randn(N, C, T, device="cuda").to(dtype)
    w = torch.randn(C, 1, K, device="cuda") / K ** 0.5
    y = ops.dwconv_fwd(ntc(x), w, stride=stride)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    yr = F.conv1d(xr, wr, stride=stride, padding=K // 2, groups=C)
    assert y.shape == (N, yr.shape[2], C)
    assert rel_err(y.float(), ntc(yr)) < tol
    dy = torch.randn_like(yr).to(dtype)
    yr.backward(dy.double())
    dw = ops.dwconv_wgrad(ntc(x), ntc(dy), K, stride=stride)
    assert dw.shape == (C, 1, K)
    assert rel_err(dw, wr.grad) < tol
    if stride == 1:
        addend = torch.randn(N, T, C, device="cuda").to(dtype)
        dx = ops.dwconv_fwd(ntc(dy), w, stride=1, flip=True, addend=addend)
        assert rel_err(dx.float(), ntc(xr.grad) + addend.double()) < tol


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("M,Cin,Cout,T", [(1002, 64, 256, 501), (900, 256, 256, 300), (1000, 336, 512, 250),
                                           (777, 512, 1024, 777), (640, 1024, 29, 320), (300, 1024, 4334, 100),
                                           (25632, 256, 256, 801), (12800, 512, 512, 400)])
def test_pwconv_fwd_mask_stats_dgrad_wgrad(ops, dtype, tol, M, Cin, Cout, T):
    torch.manual_seed(M)
    x = torch.randn(M, Cin, device="cuda").to(dtype)
    w = (torch.randn(Cout, Cin, device="cuda") / Cin ** 0.5).to(dtype)
    nb = M // T
    lengths = torch.randint(T // 2, T + 1, (nb,), device="cuda", dtype=torch.int32)
    ld = (Cout + 7) // 8 * 8  # odd widths (decoder: V' = 29 / 4334) use a padded row pitch
    stats = torch.zeros(2, Cout, device="cuda", dtype=torch.float64)
    y = ops.pwconv_fwd(x, w, lengths=lengths, T=T, stats=stats, ldy=ld)
    assert y.shape == (M, ld)
    ref = x.double() @ w.double().t()
    t = torch.arange(M, device="cuda")
    ref = ref * ((t % T) < lengths[(t // T).clamp_max(nb - 1)]).unsqueeze(1)
    assert rel_err(y[:, :Cout].float(), ref) < tol
    assert rel_err(stats[0], ref.sum(0)) < 10 * tol or ref.sum(0).norm() < 1e-3
    assert rel_err(stats[1], (ref * ref).sum(0)) < tol
    bias = torch.randn(Cout, device="cuda")
    y2 = ops.pwconv_fwd(x, w, bias=bias, ldy=ld)
    assert rel_err(y2[:, :Cout].float(), x.double() @ w.double().t() + bias.double()) < tol
    dy = torch.zeros(M, ld, device="cuda", dtype=dtype)
    dy[:, :Cout] = torch.randn(M, Cout, device="cuda").to(dtype)
    dw = ops.pwconv_wgrad(dy, x, Cout=Cout)
    assert dw.shape == (Cout, Cin)
    assert rel_err(dw, dy[:, :Cout].double().t() @ x.double()) < tol
    dx = ops.pwconv_dgrad(dy, w, lddy=ld)
    assert rel_err(dx.float(), dy[:, :Cout].double() @ w.double()) < tol
    db = ops.colsum(dy, Cout)
    assert rel_err(db, dy[:, :Cout].double().sum(0)) < tol


@pytest.mark.parametrize("dtype,tol", DTYPES)
@pytest.mark.parametrize("C,with_res", [(256, True), (512, False), (336, True), (1024, False)])
def test_bn_fwd_bwd(ops, dtype, tol, C, with_res):
    """BN(train) + residual + ReLU forward and backward vs autograd on the same (rounded) inputs."""
    torch.manual_seed(C)
    N, T = 4, 203
    y = (torch.randn(N, T, C, device="cuda") * 2 + 0.5).to(dtype)
    r = (torch.randn(N, T, C, device="cuda") - 0.3).to(dtype) if with_res else None
    lengths = torch.tensor([T, T - 20, T // 2, 1], device="cuda", dtype=torch.int32)
    g1, b1 = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    g2, b2 = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")

    def stats_of(t):  # the statistics epilogue of the pointwise GEMM, through an identity weight
        st = torch.zeros(2, C, device="cuda", dtype=torch.float64)
        ops.pwconv_fwd(t, torch.eye(C, device="cuda").to(dtype), stats=st)
        return st

    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    nbt = torch.zeros((), device="cuda", dtype=torch.long)
    bn1 = ops.BNForward(g1, b1, rm, rv, nbt, stats_of(y))
    bn2 = ops.BNForward(g2, b2, None, None, None, stats_of(r)) if with_res else None
    out = ops.bn_apply_act(y, bn1, r, bn2, None, ops.ACT_RELU)
    assert int(nbt) == 1

    yd = y.double().requires_grad_(True)
    rd = r.double().requires_grad_(True) if with_res else None
    g1d, b1d = g1.double().requires_grad_(True), b1.double().requires_grad_(True)
    g2d, b2d = g2.double().requires_grad_(True), b2.double().requires_grad_(True)
    rmd, rvd = torch.zeros(C, device="cuda", dtype=torch.double), torch.ones(C, device="cuda", dtype=torch.double)
    z = F.batch_norm(yd.reshape(-1, C), rmd, rvd, g1d, b1d, True, 0.1, 1e-3).reshape(N, T, C)
    if with_res:
        z = z + F.batch_norm(rd.reshape(-1, C), None, None, g2d, b2d, True, 0.1, 1e-3).reshape(N, T, C)
    ref = torch.relu(z)
    assert rel_err(out.float(), ref) < tol
    assert rel_err(rm, rmd) < 1e-4 and rel_err(rv, rvd) < 1e-4
    # eval mode reads the running statistics
    bn_eval = ops.BNForward(g1, b1, rm, rv, nbt)
    out_eval = ops.bn_apply_act(y, bn_eval, None, None, None, ops.ACT_NONE)
    ref_eval = F.batch_norm(y.double().reshape(-1, C), rmd, rvd, g1.double(), b1.double(), False, 0.1, 1e-3)
    assert rel_err(out_eval.float().reshape(-1, C), ref_eval) < tol and int(nbt) == 1

    dout = torch.randn(N, T, C, device="cuda").to(dtype)
    # the reference masks the conv output BEFORE BN, so d(conv out) is zeroed at masked frames
    keep = (torch.arange(T, device="cuda")[None, :] < lengths[:, None]).unsqueeze(-1)
    ref.backward(dout.double())
    totals = torch.zeros(3, C, device="cuda", dtype=torch.float64)
    ops.bn_act_bwd_reduce(dout, out, y, r, ops.ACT_RELU, totals)
    dg1, db1 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dg2, db2 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dy, dr = ops.bn_act_bwd_apply(dout, out, y, r, None, None, totals, None, (g1, bn1.save, dg1, db1),
                                  (g2, bn2.save, dg2, db2) if with_res else None, lengths, ops.ACT_RELU)
    gtol = tol * 3
    assert rel_err(dy.float(), yd.grad * keep) < gtol
    assert rel_err(dg1, g1d.grad) < gtol and rel_err(db1, b1d.grad) < gtol
    if with_res:
        assert rel_err(dr.float(), rd.grad) < gtol
        assert rel_err(dg2, g2d.grad) < gtol and rel_err(db2, b2d.grad) < gtol
    # standalone coefficient kernel agrees with the fused prologue
    coef = ops.bn_bwd_coef(totals, N * T, 1, g1, bn1.save, None, None)
    g = dout.float() * (out.float() > 0)
    dy2 = (coef[0] * g + coef[1] * y.float() + coef[2]) * keep
    assert rel_err(dy2, yd.grad * keep) < gtol


@pytest.mark.parametrize("V,dtype", [(29, torch.float32), (29, torch.bfloat16), (4334, torch.float32),
                                      (4334, torch.bfloat16)])
def test_ctc_matches_torch(ops, V, dtype):
    torch.manual_seed(V)
    N, T = 5, 120
    ld = (V + 7) // 8 * 8
    logits = torch.zeros(N, T, ld, device="cuda")
    logits[:, :, :V] = torch.randn(N, T, V, device="cuda") * 2
    logits = logits.to(dtype)
    in_len = torch.tensor([T, T - 7, 60, 33, 2], device="cuda", dtype=torch.int32)
    tgt_len = torch.tensor([30, 25, 29, 1, 5], device="cuda", dtype=torch.int32)  # last one infeasible (5 > 2)
    S = 30
    targets = torch.randint(0, V - 1, (N, S), device="cuda")
    targets[2, :10] = 3  # repeated labels need the blank between them
    blank = V - 1
    lse, lp = ops.log_softmax_fwd(logits, V, want_lp=True)
    lp_ref = F.log_softmax(logits[..., :V].double(), dim=-1)
    assert rel_err(lp, lp_ref) < 1e-5
    lpd = lp_ref.clone().requires_grad_(True)
    nll_ref = F.ctc_loss(lpd.transpose(0, 1), targets, in_len.long(), tgt_len.long(), blank=blank, reduction="none")
    gout = torch.rand(N, device="cuda") + 0.5
    (nll_ref[:4] * gout[:4].double()).sum().backward()
    for x, l in ((lp, None), (logits, lse)):
        nll, alpha, beta, scales = ops.ctc_fwd(x, l, targets, in_len, tgt_len, V, blank, want_beta=True)
        assert torch.isinf(nll[4]) and nll[4] > 0
        assert rel_err(nll[:4], nll_ref[:4]) < 1e-4
        grad = ops.ctc_bwd(x, l, targets, in_len, tgt_len, alpha, beta, nll, gout, V, blank, x.shape[-1],
                           torch.float32, scales=scales)
        # torch's own fp32 CTC gradient is only ~2e-4 accurate; compare against the fp64 run
        assert rel_err(grad[:4, :, :V], lpd.grad[:4]) < 5e-4
        assert grad[:4, :, V:].abs().max().item() == 0 if x.shape[-1] > V else True
        assert grad[1, T - 7:].abs().max().item() == 0


def test_greedy_decode_bit_exact(ops, labels28):
    from oracle import ctc_oracle
    torch.manual_seed(3)
    N, T, V = 6, 257, 29
    x = torch.randn(N, T, V, device="cuda")
    x[:, :, 28] += 1.5  # plenty of blanks
    x[0, 10:20] = x[0, 10:11]  # repeated frames -> repeats collapse
    x[1, 5, 3] = x[1, 5, 7] = 9.0  # exact tie -> lowest index
    lens = torch.tensor([T, 200, 1, 0, 131, 64], device="cuda", dtype=torch.int32)
    for ln in (lens, None):
        amax, tokens, counts = ops.greedy_decode(x, ln, V, 28)
        assert torch.equal(amax, x.argmax(-1))
        toks_ref, _ = ctc_oracle.ctc_decoder_predictions(x.argmax(-1).cpu().tolist(), labels28,
                                                         None if ln is None else ln.cpu().tolist())
        for i in range(N):
            assert tokens[i, : counts[i]].cpu().tolist() == toks_ref[i]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_bilstm_matches_torch_packed_lstm(dtype, tol):
    """functions.BiLstmFn vs the reference's BatchLSTM arithmetic (models/QuartNetContext.py:186-199):
    pack_padded_sequence -> nn.LSTM(256, 40, bidirectional) -> pad_packed_sequence on the CPU in fp32,
    forward and every gradient, ragged lengths including a length-1 and a full-length utterance."""
    from lightning_asr_b200.functions import BiLstmFn

    torch.manual_seed(3)
    N, T, Cin, H = 5, 37, 256, 40
    lens = torch.tensor([37, 1, 20, 36, 9], dtype=torch.int32)
    ref = torch.nn.LSTM(Cin, H, num_layers=1, batch_first=True, bidirectional=True)
    x = torch.randn(N, T, Cin) * 0.5
    gout = torch.randn(N, T, 2 * H)
    xr = x.clone().requires_grad_(True)
    packed = torch.nn.utils.rnn.pack_padded_sequence(xr, lens.long(), batch_first=True, enforce_sorted=False)
    yr, _ = ref(packed)
    yr, _ = torch.nn.utils.rnn.pad_packed_sequence(yr, batch_first=True, total_length=T)
    (yr * gout).sum().backward()

    params = [torch.nn.Parameter(p.detach().clone().cuda()) for p in
              (ref.weight_ih_l0, ref.weight_hh_l0, ref.bias_ih_l0, ref.bias_hh_l0, ref.weight_ih_l0_reverse,
               ref.weight_hh_l0_reverse, ref.bias_ih_l0_reverse, ref.bias_hh_l0_reverse)]
    xg = x.cuda().to(dtype).requires_grad_(True)
    y = BiLstmFn.apply(xg, lens.cuda(), *params)
    (y.float() * gout.cuda()).sum().backward()
    assert rel_err(y.float().cpu(), yr.detach()) < tol
    assert torch.count_nonzero(y[1, 1:]) == 0 and torch.count_nonzero(y[4, 9:]) == 0  # pad_packed_sequence zeros
    assert rel_err(xg.grad.float().cpu(), xr.grad) < tol
    refs = (ref.weight_ih_l0, ref.weight_hh_l0, ref.bias_ih_l0, ref.bias_hh_l0, ref.weight_ih_l0_reverse,
            ref.weight_hh_l0_reverse, ref.bias_ih_l0_reverse, ref.bias_hh_l0_reverse)
    for p, r in zip(params, refs):
        assert rel_err(p.grad.cpu(), r.grad) < tol


@pytest.mark.parametrize("N,T,S,scale,gtol", [(8, 801, 200, 1.0, 3e-3), (8, 801, 100, 3.0, 3e-3),
                                              (4, 1501, 60, 2.0, 1e-2)])
def test_ctc_long_lattices_keep_the_mass_near_the_diagonal(ops, N, T, S, scale, gtol):
    """long lattices with short / ragged targets and peaky emissions: the forward mass near the alignment diagonal sits
    hundreds of octaves below the column maximum (this is what ruled out linear-domain lattices with a shared exponent per
    frame as the default); loss and gradient against torch's fp64 CTC"""
    torch.manual_seed(T + S)
    V, ld = 29, 32
    logits = (torch.randn(N, T, ld, device="cuda") * scale).bfloat16()
    targets = torch.randint(0, V - 1, (N, S), device="cuda")
    il = torch.full((N,), T, device="cuda", dtype=torch.int32)
    il[1] = T - 37
    tl = torch.full((N,), S, device="cuda", dtype=torch.int32)
    tl[2] = S // 2
    tl[3] = 1
    lpd = torch.log_softmax(logits[..., :V].double(), -1).requires_grad_(True)
    ref = F.ctc_loss(lpd.transpose(0, 1), targets, il.long(), tl.long(), blank=V - 1, reduction="none")
    ref.sum().backward()
    lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
    nll, alpha, beta, scales = ops.ctc_fwd(logits, lse, targets, il, tl, V, V - 1, want_beta=True)
    assert torch.isfinite(nll).all()
    assert rel_err(nll, ref) < 1e-4
    grad = ops.ctc_bwd(logits, lse, targets, il, tl, alpha, beta, nll, torch.ones(N, device="cuda"), V, V - 1, ld,
                       torch.float32, scales=scales)
    # (softmax - occupancy): what torch returns as the gradient w.r.t. the log-probs.  The occupancy is
    # exp(alpha + beta - log P - lp) with fp32 lattices whose entries reach |log alpha| ~ 2000 .. 8000 here: one ulp there
    # is 2.4e-4 .. 1e-3, which bounds the relative accuracy of any fp32 log-space CTC gradient (torch's own included)
    assert rel_err(grad[..., :V], lpd.grad) < gtol


@pytest.mark.parametrize("N,T,S,dtype", [(5, 120, 30, torch.float32), (32, 801, 200, torch.bfloat16),
                                         (16, 401, 100, torch.bfloat16), (3, 257, 300, torch.float32),
                                         (2, 64, 1, torch.bfloat16), (4, 40, 15, torch.float32),
                                         (3, 500, 511, torch.bfloat16), (2, 1300, 600, torch.bfloat16)])
@pytest.mark.parametrize("use_lse", [True, False])
def test_ctc_warp_pipelined_lattices_are_bit_identical(ops, N, T, S, dtype, use_lse):
    """The default lattice kernel (csrc/ctc.cu ctc_lattice_warp2_kernel: a state pair per lane in registers, neighbours by
    warp shuffle, warp-boundary states through a shared-memory mailbox, no CTA barrier per frame) runs the arithmetic of
    the round-1 barrier kernel: alpha, beta, nll and the gradient must not differ in a single bit -- ragged lengths,
    lattices of 1 to 19 warps (both launch-bound variants), empty / one-label / infeasible targets, log-prob and
    logit + lse inputs."""
    import os
    torch.manual_seed(N * T + S)
    V, ld = 29, 32
    logits = (torch.randn(N, T, ld, device="cuda") * 2).to(dtype)
    targets = torch.randint(0, V - 1, (N, S), device="cuda")
    targets[0, : min(S, 12)] = 5  # repeats: no skip transitions there
    il = torch.full((N,), T, device="cuda", dtype=torch.int32)
    tl = torch.full((N,), min(S, T // 2), device="cuda", dtype=torch.int32)
    il[1] = T - 37 if T > 40 else T
    tl[1] = max(int(tl[1]) // 2, 1)
    if N > 2:
        tl[2] = 1
    if N > 3:
        tl[3] = 0
    if N > 4:
        il[4] = 2  # infeasible unless S <= 2
    lse, lp = ops.log_softmax_fwd(logits, V, want_lp=not use_lse)
    x, l = (logits, lse) if use_lse else (lp, None)
    gout = torch.rand(N, device="cuda") + 0.5
    old = os.environ.get("LASR_CTC_WARP")
    results = {}
    try:
        for mode in ("0", "1"):
            os.environ["LASR_CTC_WARP"] = mode
            nll, alpha, beta, scales = ops.ctc_fwd(x, l, targets, il, tl, V, V - 1, want_beta=True)
            grad = ops.ctc_bwd(x, l, targets, il, tl, alpha, beta, nll, gout, V, V - 1, x.shape[-1], torch.float32,
                               scales=scales)
            torch.cuda.synchronize()
            results[mode] = (nll.clone(), alpha.clone(), beta.clone(), grad.clone())
    finally:
        if old is None:
            os.environ.pop("LASR_CTC_WARP", None)
        else:
            os.environ["LASR_CTC_WARP"] = old
    nll0, a0, b0, g0 = results["0"]
    feasible = torch.isfinite(nll0)
    assert feasible.sum() >= min(N, 2)
    for mode in ("1",):  # the default: a (blank, label) state pair per lane
        nll, a, b, g = results[mode]
        assert torch.equal(torch.isfinite(nll), feasible), mode
        assert torch.equal(nll[feasible], nll0[feasible]), mode
        assert bool((nll[~feasible] > 0).all()), mode
        for n in range(N):
            if not bool(feasible[n]):
                continue
            Tn, Lp = int(il[n]), 2 * int(tl[n]) + 1
            assert torch.equal(a[n, :Tn, :Lp], a0[n, :Tn, :Lp]), (mode, n)
            assert torch.equal(b[n, :Tn, :Lp], b0[n, :Tn, :Lp]), (mode, n)
            assert torch.equal(g[n], g0[n]), (mode, n)


@pytest.mark.parametrize("N,T,S,dtype,gdtype", [(5, 120, 30, torch.float32, torch.float32),
                                                (8, 801, 200, torch.bfloat16, torch.bfloat16),
                                                (3, 257, 100, torch.bfloat16, torch.float32)])
def test_ctc_gradient_pass_for_small_vocabularies_is_bit_identical(ops, N, T, S, dtype, gdtype):
    """ctc_grad_small_kernel (8 frames of one utterance per CTA, labels and the row's log-probs in shared memory) against
    the general gradient pass on the same lattices: same bits, incl. zero rows past an utterance's end and for an
    infeasible utterance."""
    import os
    torch.manual_seed(N + T + S)
    V, ld = 29, 32
    logits = (torch.randn(N, T, ld, device="cuda") * 2).to(dtype)
    targets = torch.randint(0, V - 1, (N, S), device="cuda")
    il = torch.full((N,), T, device="cuda", dtype=torch.int32)
    tl = torch.full((N,), min(S, T // 2), device="cuda", dtype=torch.int32)
    il[1] = T - 37
    tl[1] = 3
    tl[2] = 0
    if N > 4:
        il[4] = 2
    lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
    gout = torch.rand(N, device="cuda") + 0.5
    nll, alpha, beta, scales = ops.ctc_fwd(logits, lse, targets, il, tl, V, V - 1, want_beta=True)
    old = os.environ.get("LASR_CTC_GRAD_SMALL")
    grads = {}
    try:
        for mode in ("0", "1"):
            os.environ["LASR_CTC_GRAD_SMALL"] = mode
            grads[mode] = ops.ctc_bwd(logits, lse, targets, il, tl, alpha, beta, nll, gout, V, V - 1, ld, gdtype,
                                      scales=scales).clone()
            torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("LASR_CTC_GRAD_SMALL", None)
        else:
            os.environ["LASR_CTC_GRAD_SMALL"] = old
    feasible = torch.isfinite(nll)
    assert torch.equal(grads["0"][feasible], grads["1"][feasible])
    assert grads["1"][1, T - 37:].abs().max().item() == 0


@pytest.mark.parametrize("M,C,ld", [(25632, 4334, 4336), (1000, 512, 512), (777, 257, 264), (64, 1024, 1024)])
def test_colsum_vectorised_bf16(ops, M, C, ld):
    """the decoder-bias gradient of a large vocabulary: column sums of a padded bf16 [M, ld] matrix (16-byte vectors, 8 row
    lanes per block) against an fp64 sum; the pad columns stay out of it and `out` is accumulated into."""
    torch.manual_seed(C)
    x = torch.randn(M, ld, device="cuda").bfloat16()
    x[:, C:] = 7.0  # pad columns must not leak into the sums
    out = torch.full((C,), 0.5, device="cuda")
    got = ops.colsum(x, C, out=out)  # x is the padded [M, ld] matrix, C its valid columns
    ref = x[:, :C].double().sum(0) + 0.5
    assert rel_err(got, ref) < 1e-5


@pytest.mark.parametrize("N,T,S,V,dtype,gdtype", [(4, 120, 30, 300, torch.bfloat16, torch.bfloat16),
                                                  (6, 401, 60, 4334, torch.bfloat16, torch.bfloat16),
                                                  (3, 257, 40, 4334, torch.bfloat16, torch.float32),
                                                  (3, 100, 20, 1000, torch.float32, torch.float32)])
def test_ctc_gradient_pass_for_large_vocabularies_is_bit_identical(ops, N, T, S, V, dtype, gdtype):
    """ctc_grad_large_kernel (occupancy per target POSITION + a class -> first-position table shared by the 8 frame rows
    of a CTA, 16-byte class vectors) against the general pass with its dense per-warp occ[V]: same bits wherever a class
    sits at one position, 1 ulp where colliding atomics of one class may be ordered differently -- repeated labels,
    ragged / empty / infeasible utterances, the vector and the scalar class loop."""
    import os
    torch.manual_seed(N + T + V)
    ld = (V + 7) // 8 * 8
    logits = torch.zeros(N, T, ld, device="cuda")
    logits[..., :V] = torch.randn(N, T, V, device="cuda") * 2
    logits = logits.to(dtype)
    targets = torch.randint(0, V - 1, (N, S), device="cuda")
    targets[0, ::3] = 17          # one class at many positions
    targets[1, 5:9] = 4           # adjacent repeats (no skip transitions)
    il = torch.full((N,), T, device="cuda", dtype=torch.int32)
    tl = torch.full((N,), S, device="cuda", dtype=torch.int32)
    il[1] = T - 37
    tl[1] = 9
    tl[2] = 0
    if N > 3:
        il[3] = 2  # infeasible
    lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
    gout = torch.rand(N, device="cuda") + 0.5
    nll, alpha, beta, scales = ops.ctc_fwd(logits, lse, targets, il, tl, V, V - 1, want_beta=True)
    old = os.environ.get("LASR_CTC_GRAD_LARGE")
    grads = {}
    try:
        for mode in ("0", "1"):
            os.environ["LASR_CTC_GRAD_LARGE"] = mode
            grads[mode] = ops.ctc_bwd(logits, lse, targets, il, tl, alpha, beta, nll, gout, V, V - 1, ld, gdtype,
                                      scales=scales).clone()
            torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("LASR_CTC_GRAD_LARGE", None)
        else:
            os.environ["LASR_CTC_GRAD_LARGE"] = old
    feasible = torch.isfinite(nll)
    assert int(feasible.sum()) >= 2
    # same additions per class; when several positions of ONE class collide in a shared-memory atomic the hardware may
    # serialise them in another order than in the dense layout: 1 ulp of fp32 on such a class (invisible in bf16)
    if gdtype == torch.bfloat16:
        # bf16 -> bf16: the large-vocabulary pass takes exp(x - lse) from ex2.approx (2^-22 relative), the general pass
        # from libm: the fp32 values agree to ~1e-6, so a bf16 rounding flips on a few elements in a thousand, by one ulp
        g0, g1 = grads["0"][feasible].float(), grads["1"][feasible].float()
        d = (g1 - g0).abs()
        assert (d > 0).float().mean().item() < 5e-3
        assert bool((d <= 2.0 ** -7 * torch.maximum(g0.abs(), g1.abs()) + 1e-30).all())
        assert rel_err(g1, g0) < 1e-3
    else:
        assert rel_err(grads["1"][feasible], grads["0"][feasible]) < 1e-6
        once = torch.ones(V, dtype=torch.bool, device="cuda")
        once[17] = once[4] = False  # the classes planted at several positions
        assert torch.equal(grads["0"][feasible][..., :V][..., once], grads["1"][feasible][..., :V][..., once])
    assert grads["1"][1, T - 37:].abs().max().item() == 0
    # and against torch's fp64 CTC on the same log-probs
    lpd = torch.log_softmax(logits[..., :V].double(), -1).requires_grad_(True)
    ref = F.ctc_loss(lpd.transpose(0, 1), targets, il.long(), tl.long(), blank=V - 1, reduction="none")
    (ref[feasible] * gout[feasible].double()).sum().backward()
    # (fp32 log-space lattices of |log alpha| ~ 1e3 bound the accuracy of the occupancy at ~1e-3 -- see the long-lattice
    # test above; the kernel under test is pinned bit for bit to the general pass, which the torch comparisons pin)
    tol = 1.5e-3 if gdtype == torch.float32 else 1e-2
    assert rel_err(grads["1"][feasible][..., :V].float(), lpd.grad[feasible]) < tol


def _bilstm_generations(ops, N, T, dtype, modes):
    """forward + backward of the BiLSTM recurrence under LASR_LSTM_V1 = each of `modes` on the same seeded inputs"""
    import os
    torch.manual_seed(N * T)
    H = 40
    pre = (torch.randn(N, T, 8 * H, device="cuda") * 0.7).to(dtype)
    whh = torch.randn(2, 4 * H, H, device="cuda") * 0.2
    lens = torch.full((N,), T, device="cuda", dtype=torch.int32)
    lens[0] = max(T // 2, 1)
    if N > 2:
        lens[2] = 0
    dout = torch.randn(N, T, 2 * H, device="cuda").to(dtype)
    res = {}
    old = os.environ.get("LASR_LSTM_V1")
    try:
        for mode in modes:
            if mode == "default":
                os.environ.pop("LASR_LSTM_V1", None)
            else:
                os.environ["LASR_LSTM_V1"] = mode
            out, gates, cells = ops.bilstm_fwd(pre, whh, lens, H)
            dwhh = torch.zeros_like(whh)
            dpre = ops.bilstm_bwd(dout, out, gates, cells, whh, lens, dwhh, H)
            torch.cuda.synchronize()
            res[mode] = (out.clone(), gates.clone(), cells.clone(), dpre.clone(), dwhh.clone())
    finally:
        if old is None:
            os.environ.pop("LASR_LSTM_V1", None)
        else:
            os.environ["LASR_LSTM_V1"] = old
    return res, lens


@pytest.mark.parametrize("N,T,dtype", [(3, 57, torch.float32), (4, 401, torch.bfloat16), (2, 9, torch.float32)])
def test_bilstm_one_barrier_kernels_are_bit_identical(ops, N, T, dtype):
    """csrc/lstm.cu second generation (the four gates of a unit in four adjacent lanes: shuffles instead of a shared-memory
    round trip, ONE barrier per frame, the backward's gate-gradient chain on all 160 threads) against the two-barrier
    kernels: outputs, saved gates / cells, dpre and dW_hh must not differ in a bit -- ragged, zero and full lengths."""
    res, lens = _bilstm_generations(ops, N, T, dtype, ("1", "2"))
    o1, g1, c1, p1, w1 = res["1"]
    o2, g2, c2, p2, w2 = res["2"]
    assert torch.equal(o1, o2) and torch.equal(p1, p2)
    for n in range(N):  # gates / cells are only defined inside an utterance
        ln = int(lens[n])
        assert torch.equal(g1[n, :ln], g2[n, :ln]) and torch.equal(c1[n, :ln], c2[n, :ln])
    # dW_hh: per-CTA register sums added with atomics from 2N CTAs -- the order of those few additions is not fixed
    assert rel_err(w2, w1) < 1e-6


@pytest.mark.parametrize("N,T", [(3, 57), (4, 401), (2, 9), (1, 1), (2, 4), (2, 5)])
def test_bilstm_third_generation_fp32_structure(ops, N, T):
    """third generation with libm activations (LASR_LSTM_V1=3 on fp32): packed FFMA2 dot products, 32-bit row indices, the
    unrolled rings running virtual frames past the end, per-lane gate gradients.  The forward must reproduce generation 1
    bit for bit (FFMA2 = two IEEE FMAs); the backward re-associates a few products: 1e-6.  Lengths around the ring sizes
    (4 and 8) included."""
    res, lens = _bilstm_generations(ops, N, T, torch.float32, ("1", "3"))
    o1, g1, c1, p1, w1 = res["1"]
    o3, g3, c3, p3, w3 = res["3"]
    assert torch.equal(o1, o3)
    for n in range(N):
        ln = int(lens[n])
        assert torch.equal(g1[n, :ln], g3[n, :ln]) and torch.equal(c1[n, :ln], c3[n, :ln])
        assert p3[n, ln:].abs().max().item() == 0 if ln < T else True
    assert rel_err(p3, p1) < 1e-6
    assert rel_err(w3, w1) < 1e-6


@pytest.mark.parametrize("N,T", [(4, 401), (3, 57), (2, 8)])
def test_bilstm_third_generation_bf16_fast_activations(ops, N, T):
    """the bf16 default (third generation, sigmoid / tanh through ex2.approx + rcp.approx, abs. error ~2e-7) against the
    libm kernels: saved fp32 gates and cells within 2e-6 absolute, outputs and gradients within bf16 rounding noise (a
    1e-7 difference flips a bf16 rounding now and then: norm-relative 2e-3, 2^-9 = 2e-3 per flipped element)."""
    res, lens = _bilstm_generations(ops, N, T, torch.bfloat16, ("1", "default"))
    o1, g1, c1, p1, w1 = res["1"]
    o3, g3, c3, p3, w3 = res["default"]
    for n in range(N):
        ln = int(lens[n])
        if ln == 0:
            continue
        assert (g1[n, :ln] - g3[n, :ln]).abs().max().item() < 5e-6
        assert (c1[n, :ln] - c3[n, :ln]).abs().max().item() < 2e-5 * max(1.0, c1[n, :ln].abs().max().item())
        assert o3[n, ln:].abs().max().item() == 0 if ln < T else True
    assert rel_err(o3, o1) < 2e-3
    assert rel_err(p3, p1) < 4e-3
    assert rel_err(w3, w1) < 2e-3
