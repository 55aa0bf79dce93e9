"""Channel-major series path (include/lasr.h "channel-major series"): the BatchNorm apply pass that writes the series
companion, and the TMA-fed depthwise kernels that read it.  Yardsticks: fp64 torch ops on the same bf16 inputs
(models/QuartNet.py:14-21,24,30,35-37)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from lightning_asr_b200 import _lib, ops
    _lib.require_device()
    return ops


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _dw_ref(x, w, flip=False):
    """x [N, T, C] -> fp64 depthwise conv, channels-last"""
    K = w.shape[-1]
    ww = w.double().flip(-1) if flip else w.double()
    return F.conv1d(x.double().transpose(1, 2), ww, padding=K // 2, groups=x.shape[-1]).transpose(1, 2)


@pytest.mark.parametrize("T", [801, 157, 896, 897, 2000])
@pytest.mark.parametrize("K", [33, 39, 75])
def test_series_layout_round_trip(ops, T, K):
    x = torch.randn(3, T, 64, device="cuda").bfloat16()
    s = ops.series_from_ntc(x, K)
    assert s.S % 128 == 0 and s.off % 8 == 0 and s.off >= K // 2 and s.S >= s.off + T + 48
    back, buf = ops.series_to_ntc(s)
    assert torch.equal(back, x)
    assert buf[:, :, :s.off].abs().max().item() == 0 and buf[:, :, s.off + T:].abs().max().item() == 0


@pytest.mark.parametrize("N,T,C,K", [(4, 157, 256, 33), (32, 801, 256, 33), (32, 801, 256, 39), (32, 801, 512, 51),
                                     (32, 801, 512, 63), (32, 801, 512, 75), (8, 801, 512, 87), (3, 801, 64, 33),
                                     (5, 2001, 256, 51), (2, 3000, 128, 75), (1, 40, 64, 33), (7, 896, 64, 39),
                                     (6, 897, 64, 39)])
@pytest.mark.parametrize("flip", [False, True])
def test_dwconv_fwd_cm_matches_fp64(ops, N, T, C, K, flip):
    torch.manual_seed(N * 1000 + T + C + K)
    x = torch.randn(N, T, C, device="cuda").bfloat16()
    w = (torch.randn(C, 1, K, device="cuda") / K ** 0.5)
    xs = ops.series_from_ntc(x, K)
    y = ops.dwconv_fwd_cm(xs, w, flip=flip)
    # the kernel multiplies bf16-rounded taps
    ref = _dw_ref(x, w.bfloat16().float(), flip)
    assert rel_err(y, ref) < 6e-3  # bf16 output rounding
    # and agrees with the gather kernel on the same inputs
    y_old = ops.dwconv_fwd(x, w, flip=flip)
    assert rel_err(y, y_old.float()) < 6e-3
    addend = torch.randn_like(x)
    y2 = ops.dwconv_fwd_cm(xs, w, flip=flip, addend=addend)
    assert rel_err(y2, ref + addend.double()) < 6e-3


@pytest.mark.parametrize("N,T,C,K,has_r,has_gate", [(4, 157, 256, 33, True, False), (32, 801, 512, 75, True, False),
                                                    (3, 801, 64, 39, False, False), (5, 301, 128, 51, True, True),
                                                    (2, 897, 64, 33, False, True)])
def test_bn_apply_cm_matches_plain_pass_and_layout(ops, N, T, C, K, has_r, has_gate):
    torch.manual_seed(T + C)
    dev = "cuda"
    y = torch.randn(N, T, C, device=dev).bfloat16()
    r = torch.randn(N, T, C, device=dev).bfloat16() if has_r else None
    gate = torch.rand(N, C, device=dev) if has_gate else None

    def bn(src):
        sums = torch.stack((src.double().sum((0, 1)), (src.double() ** 2).sum((0, 1)))).contiguous()
        return ops.BNForward(torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev), torch.zeros(C, device=dev),
                             torch.ones(C, device=dev), torch.zeros((), device=dev, dtype=torch.int64), sums)

    torch.manual_seed(1)
    b1, b2 = bn(y), (bn(r) if has_r else None)
    out_ref = ops.bn_apply_act(y, b1, r, b2, gate)
    rm_ref, rv_ref = b1.running_mean.clone(), b1.running_var.clone()
    torch.manual_seed(1)
    b1c, b2c = bn(y), (bn(r) if has_r else None)
    out, xs = ops.bn_apply_act(y, b1c, r, b2c, gate, cm_k=K)
    assert torch.equal(out, out_ref)
    assert torch.equal(b1c.running_mean, rm_ref) and torch.equal(b1c.running_var, rv_ref)
    assert torch.equal(b1c.save, b1.save)
    back, buf = ops.series_to_ntc(xs)
    assert torch.equal(back, out)
    assert buf[:, :, :xs.off].abs().max().item() == 0 and buf[:, :, xs.off + T:].abs().max().item() == 0
    # and the depthwise conv reads it
    w = torch.randn(C, 1, K, device=dev) / K ** 0.5
    d = ops.dwconv_fwd_cm(xs, w)
    assert rel_err(d, _dw_ref(out, w.bfloat16().float())) < 6e-3


@pytest.mark.parametrize("N,T,Cin,Cout,K", [(32, 801, 256, 256, 33), (32, 801, 256, 512, 51), (32, 801, 512, 512, 63),
                                            (5, 157, 512, 512, 75), (3, 2001, 256, 256, 39), (1, 40, 128, 256, 33),
                                            (7, 897, 512, 512, 87)])
@pytest.mark.parametrize("grouped", [False, True])
def test_pwconv_dgrad_cm_writes_the_series_of_the_plain_dgrad(ops, N, T, Cin, Cout, K, grouped):
    torch.manual_seed(T + Cin + Cout)
    dy = torch.randn(N, T, Cout, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, device="cuda") / Cout ** 0.5).bfloat16()
    ref = ops.pwconv_dgrad(dy, w)
    if grouped:
        dy2 = torch.randn(N, T, Cout, device="cuda").bfloat16()
        w2 = (torch.randn(Cout, Cin, device="cuda") / Cout ** 0.5).bfloat16()
        s1, s2 = ops.pwconv_dgrad_cm(dy, w, K, dy2, w2)
        back2, buf2 = ops.series_to_ntc(s2)
        assert rel_err(back2, ops.pwconv_dgrad(dy2, w2)) < 1e-2
        assert buf2[:, :, :s2.off].abs().max().item() == 0 and buf2[:, :, s2.off + T:].abs().max().item() == 0
    else:
        s1 = ops.pwconv_dgrad_cm(dy, w, K)
    back, buf = ops.series_to_ntc(s1)
    assert rel_err(back, ref) < 1e-2
    assert rel_err(back, (dy.double().reshape(-1, Cout) @ w.double()).reshape(N, T, Cin)) < 1e-2
    assert buf[:, :, :s1.off].abs().max().item() == 0 and buf[:, :, s1.off + T:].abs().max().item() == 0
    assert torch.isfinite(buf.float()).all()


@pytest.mark.parametrize("N,T,C,K", [(32, 801, 256, 33), (32, 801, 256, 39), (32, 801, 512, 51), (32, 801, 512, 75),
                                     (8, 801, 512, 87), (3, 157, 64, 33), (5, 2001, 256, 51), (1, 40, 64, 33),
                                     (6, 897, 64, 39), (2, 3000, 128, 63)])
@pytest.mark.parametrize("addend", ["none", "ntc", "series"])
def test_dwconv_bwd_cm_matches_fp64(ops, N, T, C, K, addend):
    torch.manual_seed(N + T + C + K)
    x = torch.randn(N, T, C, device="cuda").bfloat16()
    dy = torch.randn(N, T, C, device="cuda").bfloat16()
    w = torch.randn(C, 1, K, device="cuda") / K ** 0.5
    xs, dys = ops.series_from_ntc(x, K), ops.series_from_ntc(dy, K)
    add = torch.randn(N, T, C, device="cuda").bfloat16() if addend != "none" else None
    add_arg = ops.series_from_ntc(add, K) if addend == "series" else add
    # yardstick: fp64 autograd of the conv on the same bf16 operands (taps rounded to bf16 in the data gradient)
    xr = x.double().transpose(1, 2).requires_grad_(True)
    wr = w.double().requires_grad_(True)
    F.conv1d(xr, wr, padding=K // 2, groups=C).backward(dy.double().transpose(1, 2))
    dx_ref = _dw_ref(dy, w.bfloat16().float(), flip=True)
    if add is not None:
        dx_ref = dx_ref + add.double()
    dx, dw = ops.dwconv_bwd_cm(xs, dys, w, addend=add_arg)
    assert rel_err(dx, dx_ref) < 6e-3
    assert rel_err(dw, wr.grad) < 2e-3
    # separate launches agree
    dw2 = ops.dwconv_wgrad_cm(xs, dys, K)
    assert rel_err(dw2, wr.grad) < 2e-3
    dx2 = ops.dwconv_fwd_cm(dys, w, flip=True, addend=add_arg)
    assert torch.equal(dx2, dx)


@pytest.mark.parametrize("N,T,C,cm", [(4, 157, 256, True), (4, 157, 256, False), (32, 801, 512, True), (3, 40, 64, False),
                                      (2, 808, 1024, False)])
def test_relu_bits_replace_the_output_tensor_in_the_backward_passes(ops, N, T, C, cm):
    """forward passes write one byte per (frame, 8 channels); the backward passes driven by the bits must produce the
    same totals / gradients, bit for bit, as when they read the output tensor"""
    torch.manual_seed(C + T)
    dev = "cuda"
    y = torch.randn(N, T, C, device=dev).bfloat16()
    r = torch.randn(N, T, C, device=dev).bfloat16()

    def bn(src):
        sums = torch.stack((src.double().sum((0, 1)), (src.double() ** 2).sum((0, 1)))).contiguous()
        return ops.BNForward(torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev), torch.zeros(C, device=dev),
                             torch.ones(C, device=dev), torch.zeros((), device=dev, dtype=torch.int64), sums)

    torch.manual_seed(2)
    b1, b2 = bn(y), bn(r)
    bits = ops.relu_bits_alloc(N, T, C, dev)
    bits.fill_(0xAA)
    if cm:
        out, _ = ops.bn_apply_act(y, b1, r, b2, cm_k=33, relu_bits=bits)
    else:
        out = ops.bn_apply_act(y, b1, r, b2, relu_bits=bits)
    # the bits are the sign of the stored output
    Tb = (T + 7) // 8
    ref = torch.zeros(N, Tb * 8, C // 8, device=dev, dtype=torch.int32)
    pos = (out.float() > 0).view(N, T, C // 8, 8).int()
    ref[:, :T] = (pos << torch.arange(8, device=dev, dtype=torch.int32)).sum(-1)
    got = bits.view(N, Tb, C // 8, 8).permute(0, 1, 3, 2).reshape(N, Tb * 8, C // 8).int()
    assert torch.equal(got[:, :T], ref[:, :T])
    dout = torch.randn(N, T, C, device=dev).bfloat16()
    lengths = torch.randint(T // 2, T + 1, (N,), device=dev, dtype=torch.int32)
    res = []
    for use_bits in (False, True):
        totals = torch.zeros(3, C, device=dev, dtype=torch.float64)
        ops.bn_act_bwd_reduce(dout, out, y, r, ops.ACT_RELU, totals, relu_bits=bits if use_bits else None)
        dg = torch.zeros(4, C, device=dev)
        dy, dr = ops.bn_act_bwd_apply(dout, out, y, r, None, None, totals, None, (b1.gamma, b1.save, dg[0], dg[1]),
                                      (b2.gamma, b2.save, dg[2], dg[3]), lengths, ops.ACT_RELU,
                                      relu_bits=bits if use_bits else None)
        res.append((totals, dy, dr, dg))
    assert rel_err(res[1][0], res[0][0]) < 1e-6  # fp32 partial sums, different CTA order
    assert rel_err(res[1][1], res[0][1]) < 1e-2 and rel_err(res[1][2], res[0][2]) < 1e-2
    assert rel_err(res[1][3], res[0][3]) < 1e-5


@pytest.mark.parametrize("N,T,C,K", [(32, 801, 256, 33), (8, 801, 512, 75), (3, 157, 64, 39), (4, 2001, 128, 87)])
def test_prebuilt_toeplitz_factors_give_the_same_results(ops, N, T, C, K):
    """the BatchNorm pass builds the Toeplitz factors of the conv that reads its series; forward and backward launches
    that fetch them must reproduce the launches that build them in their own prologue bit for bit"""
    torch.manual_seed(K + C)
    dev = "cuda"
    y = torch.randn(N, T, C, device=dev).bfloat16()
    sums = torch.stack((y.double().sum((0, 1)), (y.double() ** 2).sum((0, 1)))).contiguous()
    bn = ops.BNForward(torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev), torch.zeros(C, device=dev),
                       torch.ones(C, device=dev), torch.zeros((), device=dev, dtype=torch.int64), sums)
    w = torch.randn(C, 1, K, device=dev) / K ** 0.5
    out, xs = ops.bn_apply_act(y, bn, cm_k=K, cm_w=w)
    assert xs.toep is not None and xs.toep_flip is not None
    d = ops.dwconv_fwd_cm(xs, w)
    toep, toep_flip = xs.toep, xs.toep_flip
    xs.toep = xs.toep_flip = None
    assert torch.equal(d, ops.dwconv_fwd_cm(xs, w))
    dy = ops.series_from_ntc(torch.randn(N, T, C, device=dev).bfloat16(), K)
    dx1, dw1 = ops.dwconv_bwd_cm(xs, dy, w, toep_flip=toep_flip)
    dx2, dw2 = ops.dwconv_bwd_cm(xs, dy, w)
    assert torch.equal(dx1, dx2)
    assert rel_err(dw1, dw2) < 1e-5  # fp32 atomics in a different CTA split order
    assert rel_err(d, _dw_ref(out, w.bfloat16().float())) < 6e-3
