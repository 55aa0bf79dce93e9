"""Thin Python wrappers around the C ABI: allocate outputs with torch, pass raw device pointers.

Every function here launches hand-written sm_100a kernels on the current torch CUDA stream.  There is no
eager / CPU fallback: inputs must be CUDA tensors and the shared library must be loadable.

Activations are channels-last: [N, T, C] contiguous (C fastest).
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, call, dtype_code  # noqa: F401


def _chk(t, name):
    if not t.is_cuda:
        raise _lib.LasrError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.LasrError(f"{name} must be contiguous")
    return t


def out_lengths(T, percents):
    """lengths = torch.mul(T, percents).int()  -- models/QuartNet.py:311 / train.py:76, fp32 multiply then truncate."""
    return torch.mul(T, percents.float()).int()


def nct_to_ntc(x, dtype):
    """[N, C, T] fp32 -> [N, T, C] dtype."""
    _chk(x, "x")
    N, C, T = x.shape
    y = torch.empty((N, T, C), device=x.device, dtype=dtype)
    call("lasr_nct_to_ntc", x, y, N, C, T, dtype_code(dtype))
    return y


def ntc_to_nct(y):
    _chk(y, "y")
    N, T, C = y.shape
    x = torch.empty((N, C, T), device=y.device, dtype=torch.float32)
    call("lasr_ntc_to_nct", y, x, N, C, T, dtype_code(y.dtype))
    return x


def cast_weight(w2d, dtype, transpose=False):
    """fp32 [R, C] master weight -> dtype copy ([R, C] or [C, R])."""
    _chk(w2d, "w")
    R, C = w2d.shape
    out = torch.empty((C, R) if transpose else (R, C), device=w2d.device, dtype=dtype)
    call("lasr_cast_weight", w2d, out, R, C, 1 if transpose else 0, dtype_code(dtype))
    return out


def dw_out_len(T_in, K, stride):
    return (T_in + 2 * (K // 2) - K) // stride + 1


def dwconv_fwd(x, w, stride=1, flip=False, addend=None):
    """x [N, T_in, C], w [C, 1, K] fp32 (nn.Conv1d(groups=C).weight) -> y [N, T_out, C]."""
    _chk(x, "x"), _chk(w, "w")
    N, T_in, C = x.shape
    K = w.shape[-1]
    T_out = dw_out_len(T_in, K, stride)
    y = torch.empty((N, T_out, C), device=x.device, dtype=x.dtype)
    call("lasr_dwconv1d_fwd", x, w, y, addend, N, T_in, T_out, C, K, stride, 1 if flip else 0, dtype_code(x.dtype))
    return y


def dwconv_wgrad(x, dy, K, stride=1, out=None):
    """dw [C, 1, K] fp32 += sum dy * shifted x  (out: a zeroed / accumulating buffer, e.g. the flat-bucket view)."""
    _chk(x, "x"), _chk(dy, "dy")
    N, T_in, C = x.shape
    T_out = dy.shape[1]
    dw = out if out is not None else torch.zeros((C, 1, K), device=x.device, dtype=torch.float32)
    call("lasr_dwconv1d_wgrad", x, dy, dw, N, T_in, T_out, C, K, stride, dtype_code(x.dtype))
    return dw


def dwconv_bwd(x, dy, w, addend=None, out_dw=None):
    """Stride-1 backward in one launch: -> (dx [N, T, C], dw [C, 1, K] fp32 accumulated into out_dw if given)."""
    _chk(x, "x"), _chk(dy, "dy"), _chk(w, "w")
    N, T, C = x.shape
    K = w.shape[-1]
    if dy.shape != x.shape:
        raise _lib.LasrError("dwconv_bwd: stride-1 layers only (dy and x must have the same shape)")
    dx = torch.empty_like(dy)
    dw = out_dw if out_dw is not None else torch.zeros((C, 1, K), device=x.device, dtype=torch.float32)
    call("lasr_dwconv1d_bwd", x, dy, w, addend, dx, dw, N, T, C, K, dtype_code(x.dtype))
    return dx, dw


class Series:
    """Channel-major companion of a channels-last activation (include/lasr.h, "channel-major series"): the operand of
    the TMA-fed depthwise kernels.  t: bf16 [C, N, S]; built for a depthwise conv of kernel size K over T frames."""
    __slots__ = ("t", "N", "T", "C", "K", "S", "off", "toep", "toep_flip")

    def __init__(self, t, N, T, C, K):
        self.t, self.N, self.T, self.C, self.K = t, N, T, C, K
        self.S, self.off = cm_pitch(T, K), cm_offset(K)
        # Toeplitz factors of the conv that reads the series (forward / reversed taps), built by the pass that wrote it
        self.toep = self.toep_flip = None


def cm_offset(K):
    return _lib.load().lasr_cm_offset(K)


def cm_pitch(T, K):
    return _lib.load().lasr_cm_pitch(T, K)


def cm_supported(C, K, dtype, stride=1):
    """a depthwise layer (C channels, kernel K) can read its input as a Series"""
    return dtype == torch.bfloat16 and stride == 1 and C % 64 == 0 and K % 2 == 1 and 3 <= K <= 89


def series_from_ntc(x, K):
    """Reference construction of a Series from a channels-last tensor with torch ops (tests / the layers whose input is
    not produced by a BatchNorm pass)."""
    N, T, C = x.shape
    S, off = cm_pitch(T, K), cm_offset(K)
    buf = torch.zeros((C, N, S), device=x.device, dtype=x.dtype)
    buf[:, :, off:off + T] = x.permute(2, 0, 1)
    g = torch.arange(S // 8, device=x.device)
    perm = g ^ ((g >> 3) & 1)
    t = buf.view(C, N, S // 8, 8)[:, :, perm, :].reshape(C, N, S).contiguous()
    return Series(t, N, T, C, K)


def series_to_ntc(s):
    """inverse of series_from_ntc (tests)"""
    g = torch.arange(s.S // 8, device=s.t.device)
    perm = g ^ ((g >> 3) & 1)
    buf = s.t.view(s.C, s.N, s.S // 8, 8)[:, :, perm, :].reshape(s.C, s.N, s.S)
    return buf[:, :, s.off:s.off + s.T].permute(1, 2, 0).contiguous(), buf


def _addends(addend):
    """addend: None | channels-last tensor | Series -> (channels-last ptr arg, series ptr arg)"""
    if addend is None:
        return None, None
    if isinstance(addend, Series):
        return None, addend.t
    return _chk(addend, "addend"), None


def dwconv_fwd_cm(xs, w, flip=False, addend=None):
    """depthwise conv (stride 1, bf16) reading its input from a Series -> y [N, T, C] channels-last.
    addend: channels-last [N, T, C] tensor or a Series laid out like xs."""
    _chk(w, "w")
    K = w.shape[-1]
    if K != xs.K:
        raise _lib.LasrError(f"dwconv_fwd_cm: the series was laid out for K={xs.K}, the conv has K={K}")
    y = torch.empty((xs.N, xs.T, xs.C), device=xs.t.device, dtype=torch.bfloat16)
    a_cl, a_cm = _addends(addend)
    call("lasr_dwconv1d_fwd_cm", xs.t, w, y, a_cl, a_cm, xs.toep_flip if flip else xs.toep, xs.N, xs.T, xs.C, K, xs.S,
         1 if flip else 0)
    return y


def dwconv_wgrad_cm(xs, dys, K, out=None):
    """dw [C, 1, K] fp32 += sum dy * shifted x from the Series of x and dy."""
    dw = out if out is not None else torch.zeros((xs.C, 1, K), device=xs.t.device, dtype=torch.float32)
    call("lasr_dwconv1d_wgrad_cm", xs.t, dys.t, dw, xs.N, xs.T, xs.C, K, xs.S)
    return dw


def dwconv_bwd_cm(xs, dys, w, addend=None, out_dw=None, toep_flip=None):
    """Stride-1 backward in one launch from Series operands -> (dx [N, T, C] channels-last, dw [C, 1, K] fp32).
    toep_flip: the prebuilt Toeplitz factors of the reversed taps (Series.toep_flip of the forward's input)."""
    _chk(w, "w")
    K = w.shape[-1]
    if K != xs.K or K != dys.K or (xs.N, xs.T, xs.C) != (dys.N, dys.T, dys.C):
        raise _lib.LasrError("dwconv_bwd_cm: the series of x and dy must share the conv's layout")
    dx = torch.empty((xs.N, xs.T, xs.C), device=xs.t.device, dtype=torch.bfloat16)
    dw = out_dw if out_dw is not None else torch.zeros((xs.C, 1, K), device=xs.t.device, dtype=torch.float32)
    a_cl, a_cm = _addends(addend)
    call("lasr_dwconv1d_bwd_cm", xs.t, dys.t, w, a_cl, a_cm, toep_flip, dx, dw, xs.N, xs.T, xs.C, K, xs.S)
    return dx, dw


def pwconv_dgrad_cm(dy, w, K, dy2=None, w2=None):
    """Data gradient of a 1x1 conv written as the channel-major Series a depthwise conv of kernel size K consumes:
    dy [N, T, Cout], w [Cout, Cin] -> Series of dx [N, T, Cin]; (dy2, w2): the block's residual conv in the same launch
    -> (series1, series2).  Raises LasrError(UNSUPPORTED) for shapes the kernels do not take."""
    _chk(dy, "dy"), _chk(w, "w")
    N, T, Cout = dy.shape
    Cin = w.shape[1]
    s1 = Series(None, N, T, Cin, K)
    s1.t = torch.empty((Cin, N, s1.S), device=dy.device, dtype=dy.dtype)
    s2 = None
    if dy2 is not None:
        _chk(dy2, "dy2"), _chk(w2, "w2")
        s2 = Series(None, N, T, Cin, K)
        s2.t = torch.empty((Cin, N, s2.S), device=dy.device, dtype=dy.dtype)
    call("lasr_pwconv_dgrad_cm", dy, w, s1.t, dy2, w2, s2.t if s2 is not None else None, N, T, Cin, Cout, s1.S, s1.off)
    return (s1, s2) if dy2 is not None else s1


def pwconv_fwd(x2d, w, bias=None, lengths=None, T=0, stats=None, out=None, ldy=None):
    """y[M, Cout] = x[M, Cin] w[Cout, Cin]^T (+bias), MaskCNN row mask; `stats` (double [2, Cout], zeroed) receives the
    BatchNorm batch sums.  x2d may be [N, T, Cin]; w [Cout, Cin(, 1)] in x's dtype.  Output row pitch ldy >= Cout."""
    _chk(x2d, "x"), _chk(w, "w")
    Cin = x2d.shape[-1]
    M = x2d.numel() // Cin
    Cout = w.shape[0]
    if w.shape[1] != Cin or w.dtype != x2d.dtype:
        raise _lib.LasrError(f"pwconv weight {tuple(w.shape)}/{w.dtype} does not match input Cin={Cin}/{x2d.dtype}")
    ldy = Cout if ldy is None else ldy
    y = out if out is not None else torch.empty(x2d.shape[:-1] + (ldy,), device=x2d.device, dtype=x2d.dtype)
    call("lasr_pwconv_fwd", x2d, w, y, bias, lengths, T, stats, M, Cin, Cout, Cin, Cin, ldy, dtype_code(x2d.dtype))
    return y


def pwconv_fwd_fused(x2d, w, bias, residual=None, lengths=None, T=0, relu=True):
    """y = act(mask(x w^T) + bias [+ residual]) -- the eval-mode conv1x1 -> MaskCNN -> BN(folded) -> [+res] -> ReLU chain
    as one GEMM (bf16 only)."""
    _chk(x2d, "x"), _chk(w, "w")
    Cin = x2d.shape[-1]
    M = x2d.numel() // Cin
    Cout = w.shape[0]
    if w.shape[1] != Cin or w.dtype != x2d.dtype:
        raise _lib.LasrError(f"pwconv weight {tuple(w.shape)}/{w.dtype} does not match input Cin={Cin}/{x2d.dtype}")
    y = torch.empty(x2d.shape[:-1] + (Cout,), device=x2d.device, dtype=x2d.dtype)
    if residual is not None:
        _chk(residual, "residual")
    call("lasr_pwconv_fwd_fused", x2d, w, y, bias, residual, lengths, T, 1 if relu else 0, M, Cin, Cout, Cin, Cin, Cout,
         Cout, dtype_code(x2d.dtype))
    return y


def pwconv_dgrad(dy, w, lddy=None):
    """dx[M, Cin] = dy[M, Cout] w[Cout, Cin]   (w in its native layout; dy row pitch lddy >= Cout)."""
    _chk(dy, "dy"), _chk(w, "w")
    Cout, Cin = w.shape[0], w.shape[1]
    lddy = dy.shape[-1] if lddy is None else lddy
    M = dy.numel() // lddy
    dx = torch.empty(dy.shape[:-1] + (Cin,), device=dy.device, dtype=dy.dtype)
    call("lasr_pwconv_dgrad", dy, w, dx, M, Cin, Cout, lddy, Cin, Cin, dtype_code(dy.dtype))
    return dx


def pwconv_wgrad(dy, x, out=None, Cout=None):
    """dw[Cout, Cin] fp32 += dy[M, :Cout]^T x[M, Cin]   (out zeroed / accumulating)."""
    _chk(dy, "dy"), _chk(x, "x")
    lddy, Cin = dy.shape[-1], x.shape[-1]
    Cout = lddy if Cout is None else Cout
    M = x.numel() // Cin
    dw = out if out is not None else torch.zeros((Cout, Cin), device=x.device, dtype=torch.float32)
    call("lasr_pwconv_wgrad", dy, x, dw, M, Cin, Cout, lddy, Cin, Cin, dtype_code(x.dtype))
    return dw


def pwconv_fwd2(x1, w1, lengths1, stats1, x2, w2, lengths2, stats2, T):
    """Two 1x1 convs of identical shape in one launch: y1 = mask1(x1 w1^T), y2 = mask2(x2 w2^T) (+ BN statistics)."""
    for t in (x1, w1, x2, w2):
        _chk(t, "operand")
    Cin, Cout = x1.shape[-1], w1.shape[0]
    if x2.shape != x1.shape or w2.shape != w1.shape or w1.shape[1] != Cin or w1.dtype != x1.dtype or x2.dtype != x1.dtype:
        raise _lib.LasrError("pwconv_fwd2: the two problems must have identical shapes and dtypes")
    M = x1.numel() // Cin
    y1 = torch.empty(x1.shape[:-1] + (Cout,), device=x1.device, dtype=x1.dtype)
    y2 = torch.empty_like(y1)
    call("lasr_pwconv_fwd2", x1, w1, y1, lengths1, stats1, x2, w2, y2, lengths2, stats2, T, M, Cin, Cout,
         dtype_code(x1.dtype))
    return y1, y2


def pwconv_dgrad2(dy1, w1, dy2, w2):
    """dx1 = dy1 w1 and dx2 = dy2 w2 (identical shapes, dense operands) in one launch."""
    for t in (dy1, w1, dy2, w2):
        _chk(t, "operand")
    Cout, Cin = w1.shape[0], w1.shape[1]
    if dy2.shape != dy1.shape or w2.shape != w1.shape or dy1.shape[-1] != Cout:
        raise _lib.LasrError("pwconv_dgrad2: the two problems must have identical dense shapes")
    M = dy1.numel() // Cout
    dx1 = torch.empty(dy1.shape[:-1] + (Cin,), device=dy1.device, dtype=dy1.dtype)
    dx2 = torch.empty_like(dx1)
    call("lasr_pwconv_dgrad2", dy1, w1, dx1, dy2, w2, dx2, M, Cin, Cout, dtype_code(dy1.dtype))
    return dx1, dx2


def pwconv_wgrad2(dy1, x1, out1, dy2, x2, out2):
    """out1 += dy1^T x1 and out2 += dy2^T x2 (same shapes) in one launch."""
    for t in (dy1, x1, dy2, x2):
        _chk(t, "operand")
    lddy, Cin = dy1.shape[-1], x1.shape[-1]
    if dy2.shape != dy1.shape or x2.shape != x1.shape or out1.shape != out2.shape:
        raise _lib.LasrError("pwconv_wgrad2: the two problems must have identical shapes")
    M = x1.numel() // Cin
    call("lasr_pwconv_wgrad2", dy1, x1, out1, dy2, x2, out2, M, Cin, lddy, lddy, Cin, Cin, dtype_code(x1.dtype))


def colsum(x, C, out=None):
    """out[c] += sum over rows of x[..., c], c < C."""
    ld = x.shape[-1]
    M = x.numel() // ld
    o = out if out is not None else torch.zeros((C,), device=x.device, dtype=torch.float32)
    call("lasr_colsum", x, o, M, C, ld, dtype_code(x.dtype))
    return o


class _BN(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("sums", "gamma", "beta", "running_mean", "running_var",
                                               "num_batches_tracked", "save_mean", "save_invstd")]


class _BNBwd(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("gamma", "mean", "invstd", "dgamma", "dbeta")]


def _p(t):
    return None if t is None else t.data_ptr()


class _Drop(ctypes.Structure):
    _fields_ = [("mask", ctypes.c_void_p), ("mode", ctypes.c_int), ("p", ctypes.c_float), ("seed", ctypes.c_uint64),
                ("seed_dev", ctypes.c_void_p)]


class Dropout:
    """lasr_dropout_t: nn.Dropout(p) fused into the BatchNorm passes (models/QuartNet.py:27,38,149).
    mask: uint8 [N, T, C] keep mask (1 = keep).  mode 'read' applies a caller-supplied mask (parity hook); mode
    'generate' draws it on the device (Philox4x32-10 keyed by `seed`) in the forward pass and stores it in `mask`."""

    def __init__(self, mask, p, mode="read", seed=0, seed_dev=None):
        if mask.dtype != torch.uint8 or not mask.is_cuda or not mask.is_contiguous():
            raise _lib.LasrError("dropout mask must be a contiguous CUDA uint8 tensor")
        if not (0.0 < p < 1.0):
            raise _lib.LasrError("dropout p must be in (0, 1)")
        self.mask, self.p = mask, float(p)
        self.seed_dev = seed_dev  # int64 device scalar (kept alive here)
        self.c = _Drop(mask.data_ptr(), {"read": 1, "generate": 2}[mode], float(p), int(seed) & 0xFFFFFFFFFFFFFFFF,
                       _p(seed_dev))

    @property
    def ptr(self):
        return ctypes.addressof(self.c)

    def backward_view(self):
        """The same mask, read-only mode (what the backward passes use)."""
        return Dropout(self.mask, self.p, "read")


class BNForward:
    """lasr_bn_t for one BatchNorm call.  training: `sums` = double [2, C] batch statistics (from the GEMM epilogue),
    save [2, C] receives mean / invstd.  eval: sums None, running statistics are read."""

    def __init__(self, gamma, beta, running_mean, running_var, nbt, sums=None):
        self.gamma, self.beta = gamma, beta
        self.running_mean, self.running_var, self.nbt = running_mean, running_var, nbt
        self.sums = sums
        self.training = sums is not None
        self.save = torch.empty((2, gamma.numel()), device=gamma.device, dtype=torch.float32) if self.training else None
        self.c = _BN(_p(sums), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                     _p(nbt) if self.training else None, _p(self.save[0]) if self.training else None,
                     _p(self.save[1]) if self.training else None)

    @property
    def ptr(self):
        return ctypes.addressof(self.c)


def bn_coeffs(bn, count, eps, momentum, side_effects):
    C = bn.gamma.numel()
    coef = torch.empty((2, C), device=bn.gamma.device, dtype=torch.float32)
    call("lasr_bn_coeffs", bn.ptr, C, count, eps, momentum, coef[0], coef[1], 1 if side_effects else 0)
    return coef[0], coef[1]


def sum_over_time(y):
    N, T, C = y.shape
    sums = torch.empty((N, C), device=y.device, dtype=torch.float32)
    call("lasr_sum_over_time", y, sums, N, T, C, dtype_code(y.dtype))
    return sums


def relu_bits_alloc(N, T, C, device):
    """storage of the ReLU gate bits of an [N, T, C] activation (include/lasr.h: one byte per frame and 8 channels)"""
    return torch.empty((N, (T + 7) // 8, C // 8, 8), device=device, dtype=torch.uint8)


def bn_apply_act(y, bn1, r=None, bn2=None, gate=None, act=ACT_RELU, eps=1e-3, momentum=0.1, side_effects=True,
                 drop=None, cm_k=None, relu_bits=None, cm_w=None):
    """out = act(BN1(y) [* gate] [* dropout] [+ BN2(r)]) in one pass; performs the training side effects of both BNs.
    cm_k = kernel size of the depthwise conv that consumes the result: also writes the channel-major Series companion
    and returns (out, series); cm_w = that conv's taps [C, 1, cm_k]: the pass also builds its Toeplitz factors
    (series.toep / .toep_flip).  relu_bits (relu_bits_alloc): receives the sign bits the backward passes read instead of
    `out`."""
    N, T, C = y.shape
    out = torch.empty_like(y)
    if drop is not None and drop.mask.numel() != y.numel():
        raise _lib.LasrError("dropout mask must have one byte per activation element")
    if cm_k is not None:
        if drop is not None or not cm_supported(C, cm_k, y.dtype):
            raise _lib.LasrError("bn_apply_act: no channel-major companion for this configuration")
        xs = Series(None, N, T, C, cm_k)
        xs.t = torch.empty((C, N, xs.S), device=y.device, dtype=y.dtype)
        if cm_w is not None:
            ks16 = _lib.load().lasr_cm_ks(cm_k) * 16
            xs.toep = torch.empty((C, ks16), device=y.device, dtype=y.dtype)
            xs.toep_flip = torch.empty((C, ks16), device=y.device, dtype=y.dtype)
        call("lasr_bn_apply_act_fwd_cm", y, bn1.ptr, r, bn2.ptr if bn2 is not None else None, gate, out, xs.t, N, T, C,
             xs.S, xs.off, eps, momentum, act, 1 if (side_effects and bn1.training) else 0, relu_bits, cm_w, xs.toep,
             xs.toep_flip, cm_k)
        return out, xs
    call("lasr_bn_apply_act_fwd", y, bn1.ptr, r, bn2.ptr if bn2 is not None else None, gate, out, N * T, C, T, N * T,
         eps, momentum, act, 1 if (side_effects and bn1.training) else 0, drop.ptr if drop is not None else None,
         dtype_code(y.dtype), relu_bits)
    return out


def bn_bwd_chunks(N, T):
    return _lib.load().lasr_bn_bwd_chunks(N, T)


def bn_act_bwd_reduce(dout, out, y, r, act, totals, per_n=None, drop=None, relu_bits=None):
    """totals: double [3, C] (zeroed); [4, C] with dropout.  relu_bits: the forward's sign bits (then `out` is not read
    and may be None)."""
    N, T, C = y.shape
    if totals.shape[0] < (4 if drop is not None else 3):
        raise _lib.LasrError("bn_act_bwd_reduce: totals needs 4 slots with dropout, 3 without")
    call("lasr_bn_act_bwd_reduce", dout, out if relu_bits is None else None, y, r, totals, per_n, N, T, C, act,
         drop.ptr if drop is not None else None, dtype_code(y.dtype), relu_bits)
    return totals


def bn_bwd_coef(totals, count, slot_gx, gamma, save, dgamma, dbeta):
    C = gamma.numel()
    coef = torch.empty((3, C), device=gamma.device, dtype=torch.float32)
    call("lasr_bn_bwd_coef", totals, C, count, slot_gx, gamma, save[0], save[1], dgamma, dbeta, coef)
    return coef


def bn_act_bwd_apply(dout, out, y, r, gate, extra, totals, coef1, bn1, bn2, lengths, act, drop=None, relu_bits=None):
    """bn1 / bn2: (gamma, save [2,C], dgamma, dbeta) tuples or None."""
    N, T, C = y.shape
    dy = torch.empty_like(y)
    dr = torch.empty_like(y) if r is not None else None
    s1 = _BNBwd(_p(bn1[0]), _p(bn1[1][0]), _p(bn1[1][1]), _p(bn1[2]), _p(bn1[3])) if bn1 is not None else None
    s2 = _BNBwd(_p(bn2[0]), _p(bn2[1][0]), _p(bn2[1][1]), _p(bn2[2]), _p(bn2[3])) if bn2 is not None else None
    call("lasr_bn_act_bwd_apply", dout, out if relu_bits is None else None, y, r, gate, extra, totals, coef1,
         ctypes.addressof(s1) if s1 is not None else None, ctypes.addressof(s2) if s2 is not None else None, N * T,
         lengths, T, dy, dr, N * T, C, act, drop.ptr if drop is not None else None, dtype_code(y.dtype), relu_bits)
    return dy, dr


def se_excite_fwd(sums, scale, shift, T, w1, w2):
    N, C = sums.shape
    Cr = w1.shape[0]
    s = torch.empty((N, C), device=sums.device, dtype=torch.float32)
    hidden = torch.empty((N, Cr), device=sums.device, dtype=torch.float32)
    gate = torch.empty((N, C), device=sums.device, dtype=torch.float32)
    call("lasr_se_excite_fwd", sums, scale, shift, T, w1, w2, s, hidden, gate, N, C, Cr)
    return s, hidden, gate


def se_excite_bwd(per_n, scale, shift, T, w1, w2, s, hidden, gate, dw1, dw2):
    N, C = gate.shape
    Cr = w1.shape[0]
    extra = torch.empty((N, C), device=gate.device, dtype=torch.float32)
    ws = torch.empty((N * (C + Cr),), device=gate.device, dtype=torch.float32)
    call("lasr_se_excite_bwd", per_n, 1, scale, shift, T, w1, w2, s, hidden, gate, extra, dw1, dw2, ws, N, C, Cr)
    return extra


def se_bn_bwd_finalize(per_n, N, T, gate, extra, sums_y, gamma, save, dgamma, dbeta):
    C = gate.shape[1]
    coef = torch.empty((3, C), device=gate.device, dtype=torch.float32)
    call("lasr_se_bn_bwd_finalize", per_n, N, 1, C, T, gate, extra, sums_y, gamma, save[0], save[1], dgamma, dbeta,
         coef)
    return coef


def bilstm_fwd(pre, whh, lengths, hidden):
    """pre [N, T, 8H] input projections (+ biases) -> (out [N, T, 2H], gates [N, T, 2, H, 4] f32, cells [N, T, 2, H] f32)."""
    _chk(pre, "pre"), _chk(whh, "whh")
    N, T = pre.shape[0], pre.shape[1]
    if pre.shape[2] != 8 * hidden or tuple(whh.shape) != (2, 4 * hidden, hidden) or whh.dtype != torch.float32:
        raise _lib.LasrError(f"bilstm: pre {tuple(pre.shape)} / whh {tuple(whh.shape)} do not match hidden={hidden}")
    out = torch.empty((N, T, 2 * hidden), device=pre.device, dtype=pre.dtype)
    gates = torch.empty((N, T, 2, hidden, 4), device=pre.device, dtype=torch.float32)
    cells = torch.empty((N, T, 2, hidden), device=pre.device, dtype=torch.float32)
    call("lasr_bilstm_fwd", pre, whh, lengths, out, gates, cells, N, T, hidden, dtype_code(pre.dtype))
    return out, gates, cells


def bilstm_bwd(dout, out, gates, cells, whh, lengths, dwhh, hidden):
    """-> dpre [N, T, 8H] in dout's dtype; dwhh [2, 4H, H] f32 is accumulated into."""
    _chk(dout, "dout"), _chk(out, "out")
    N, T = out.shape[0], out.shape[1]
    dpre = torch.empty((N, T, 8 * hidden), device=out.device, dtype=out.dtype)
    call("lasr_bilstm_bwd", dout, out, gates, cells, whh, lengths, dpre, dwhh, N, T, hidden, dtype_code(out.dtype))
    return dpre


def log_softmax_fwd(logits, V, want_lp=True):
    """logits [..., ld] -> (lse [...], lp [..., V] fp32 or None)."""
    ld = logits.shape[-1]
    M = logits.numel() // ld
    lse = torch.empty(logits.shape[:-1], device=logits.device, dtype=torch.float32)
    lp = torch.empty(logits.shape[:-1] + (V,), device=logits.device, dtype=torch.float32) if want_lp else None
    call("lasr_log_softmax_fwd", logits, lse, lp, M, V, ld, dtype_code(logits.dtype))
    return lse, lp


def log_softmax_bwd(dlp, lp, ld, dtype):
    V = lp.shape[-1]
    M = lp.numel() // V
    dlogits = torch.empty(lp.shape[:-1] + (ld,), device=lp.device, dtype=dtype)
    call("lasr_log_softmax_bwd", dlp, lp, dlogits, M, V, ld, dtype_code(dtype))
    return dlogits


# Linear-domain ("scaled") lattice kernels with one shared power-of-two exponent per frame: 266 -> 198 us at the config-2
# shape, but OFF by default -- on long lattices the forward mass near the alignment diagonal can sit more than 2^126 below
# the column maximum (measured: T' = 801, S = 100, logits * 3: nll off by 5 %; short targets in a long batch: inf), which
# a shared exponent flushes to zero.  The log-space kernels have no such limit.  LASR_CTC_SCALED=1 opts in.
CTC_SCALED = os.environ.get("LASR_CTC_SCALED", "0") == "1"


def ctc_fwd(x, lse, targets, input_lengths, target_lengths, V, blank, want_beta):
    """x [N, T, ld] log-probs (lse None) or logits (+lse [N,T]).  -> nll [N], alpha, beta|None, scales|None.
    scales (int32 workspace) is not None when the scaled single-warp lattices ran (include/lasr.h): pass it on to
    ctc_bwd together with alpha / beta."""
    N, T, ld = x.shape
    S_max = max(int(targets.shape[1]), 1) if targets.dim() == 2 else 1
    Lp = 2 * S_max + 1
    alpha = torch.empty((N, T, Lp), device=x.device, dtype=torch.float32)
    beta = torch.empty((N, T, Lp), device=x.device, dtype=torch.float32) if want_beta else None
    nll = torch.empty((N,), device=x.device, dtype=torch.float32)
    scales = emis = None
    if CTC_SCALED and Lp <= 32 * 33:
        lib = _lib.load()
        scales = torch.empty((lib.lasr_ctc_scales_bytes(N, T) // 4,), device=x.device, dtype=torch.int32)
        emis = torch.empty((lib.lasr_ctc_emis_bytes(N, T, S_max) // 4,), device=x.device, dtype=torch.float32)
    call("lasr_ctc_fwd", x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, scales, emis, N, T, V, ld,
         S_max, blank, dtype_code(x.dtype))
    return nll, alpha, beta, scales


def ctc_bwd(x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, grad_out, V, blank, ldg, grad_dtype,
            scales=None):
    N, T, ld = x.shape
    S_max = max(int(targets.shape[1]), 1)
    grad = torch.empty((N, T, ldg), device=x.device, dtype=grad_dtype)
    call("lasr_ctc_bwd", x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, scales, grad_out, grad, N, T, V,
         ld, ldg, S_max, blank, dtype_code(x.dtype), dtype_code(grad_dtype))
    return grad


def greedy_decode(x, lengths, V, blank, collapse=True):
    """x [N, T, ld] scores -> (argmax [N,T] int64, tokens [N,T] int32, counts [N] int32)."""
    N, T, ld = x.shape
    amax = torch.empty((N, T), device=x.device, dtype=torch.int64)
    tokens = torch.empty((N, T), device=x.device, dtype=torch.int32) if collapse else None
    counts = torch.empty((N,), device=x.device, dtype=torch.int32) if collapse else None
    call("lasr_greedy_decode", x, lengths, amax, tokens, counts, N, T, V, ld, blank, dtype_code(x.dtype))
    return amax, tokens, counts


def ctc_collapse(predictions, lengths, blank):
    """predictions [N, T] int64 -> (tokens [N, T] int32, counts [N] int32): utils/asr_metrics.py:159-167."""
    _chk(predictions, "predictions")
    N, T = predictions.shape
    tokens = torch.empty((N, T), device=predictions.device, dtype=torch.int32)
    counts = torch.empty((N,), device=predictions.device, dtype=torch.int32)
    call("lasr_ctc_collapse", predictions, lengths, tokens, counts, N, T, blank)
    return tokens, counts
