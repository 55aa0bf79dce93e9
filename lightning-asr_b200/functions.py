"""autograd.Function wrappers: one Function per fused unit of the hot path.

  SepConvBNFn      SeprationConv (+ optionally the whole QuartNetBlock with its residual branch):
                   dw conv -> 1x1 conv (+MaskCNN, BN statistics) -> [residual 1x1 conv (+BN statistics)] -> BN finalize
                   -> one apply pass (BN, SE gate, residual add, ReLU).       models/QuartNet.py:8-78
  Conv1x1BNReLUFn  last_cnn2: 1x1 conv -> BN -> ReLU.                          models/QuartNet.py:145-149
  DecoderLogSoftmaxFn  decoder 1x1 conv (+bias) -> log_softmax.                models/QuartNet.py:275,282-290
  CTCLossFn        torch.nn.CTCLoss(blank, reduction='none').                   train.py:196
  FusedDecoderCTCFn  decoder + log-softmax + CTC with the gradient written straight as d(logits).

All activations inside are channels-last [N, T, C]; parameters keep the reference's shapes (state_dict schema).
Backward passes are hand-scheduled kernel sequences, not autograd graphs.
"""
import torch

from . import ops
from .ops import ACT_NONE, ACT_RELU

BN_EPS = 1e-3  # nn.BatchNorm1d(out_ch, eps=1e-3)  models/QuartNet.py:24,64,147
BN_MOMENTUM = 0.1


def _bump(nbt):
    if nbt is not None:
        nbt.add_(1)


class SepConvBNFn(torch.autograd.Function):
    """y = act( [gate *] BN1(mask(pw(dw(x)))) [+ BN2(res(x))] ).

    forward args:
      x [N, T_in, Cin] channels-last; lengths int32 [N] or None (mask off)
      res_x: input of the residual branch when it differs from x (QuartNetBlock with repeat > 1), else None
      dw_w [Cin, 1, K], pw_w [Cout, Cin, 1], bn_w, bn_b [Cout]
      res_w [Cout, Cin, 1] or None, rbn_w, rbn_b (residual branch, QuartNetBlock.reside)
      se_w1 [Cout/8, Cout], se_w2 [Cout, Cout/8] or None (SELayer.fc.0 / fc.2)
      bn_buffers = (running_mean, running_var, num_batches_tracked) for bn; rbn_buffers likewise
      stride, relu (bool), training (bool)
    """

    @staticmethod
    def forward(ctx, x, res_x, lengths, dw_w, pw_w, bn_w, bn_b, res_w, rbn_w, rbn_b, se_w1, se_w2, bn_buffers,
                rbn_buffers, stride, relu, training):
        dt = x.dtype
        N, T_in, Cin = x.shape
        K = dw_w.shape[-1]
        Cout = pw_w.shape[0]
        act = ACT_RELU if relu else ACT_NONE
        wt = ops.cast_weight(dw_w.view(Cin, K), torch.float32, transpose=True)  # [K, Cin] tap-major
        d = ops.dwconv_fwd(x, wt, stride=stride)
        T = d.shape[1]
        pw_s = ops.cast_weight(pw_w.view(Cout, Cin), dt)
        y, stats1 = ops.pwconv_fwd(d, pw_s, lengths=lengths, T=T, want_stats=training)
        has_res = res_w is not None
        r = stats2 = res_s = None
        if has_res:
            res_s = ops.cast_weight(res_w.view(Cout, res_w.shape[1]), dt)
            r, stats2 = ops.pwconv_fwd(x if res_x is None else res_x, res_s, want_stats=training)
        count = N * T
        if training:
            st1 = ops.bn_finalize(stats1, count, bn_w, bn_b, bn_buffers[0], bn_buffers[1], BN_EPS, BN_MOMENTUM)
            _bump(bn_buffers[2])
            st2 = None
            if has_res:
                st2 = ops.bn_finalize(stats2, count, rbn_w, rbn_b, rbn_buffers[0], rbn_buffers[1], BN_EPS, BN_MOMENTUM)
                _bump(rbn_buffers[2])
        else:
            st1 = ops.bn_eval_coeffs(bn_w, bn_b, bn_buffers[0], bn_buffers[1], BN_EPS)
            st2 = ops.bn_eval_coeffs(rbn_w, rbn_b, rbn_buffers[0], rbn_buffers[1], BN_EPS) if has_res else None
        gate = s = hidden = sums_y = None
        if se_w1 is not None:
            sums_y = ops.sum_over_time(y)
            s, hidden, gate = ops.se_excite_fwd(sums_y, st1, T, se_w1, se_w2)
        out = ops.bn_apply_act(y, st1, r, st2, gate, act)

        ctx.saved = (x, res_x, lengths, wt, d, y, r, out, st1, st2, gate, s, hidden, sums_y)
        ctx.params = (dw_w, pw_w, bn_w, res_w, rbn_w, se_w1, se_w2)
        ctx.cfg = (stride, act, training, K, Cin, Cout)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout):
        if dout is None:
            return (None,) * 17
        x, res_x, lengths, wt, d, y, r, out, st1, st2, gate, s, hidden, sums_y = ctx.saved
        dw_w, pw_w, bn_w, res_w, rbn_w, se_w1, se_w2 = ctx.params
        stride, act, training, K, Cin, Cout = ctx.cfg
        if not training:
            raise RuntimeError("lightning_asr_b200: backward through eval-mode BatchNorm is not implemented")
        dt = x.dtype
        dout = dout.contiguous()
        N, T, _ = y.shape
        count = N * T
        has_res = r is not None
        partials, chunks = ops.bn_act_bwd_reduce(dout, out, y, r, act)
        d_bn_w = torch.zeros_like(bn_w)
        d_bn_b = torch.zeros_like(bn_w)
        d_rbn_w = d_rbn_b = None
        extra = d_se1 = d_se2 = None
        if gate is not None:
            extra, d_se1, d_se2 = ops.se_excite_bwd(partials, chunks, st1, T, se_w1, se_w2, s, hidden, gate)
            coef1 = ops.se_bn_bwd_finalize(partials, N, chunks, T, gate, extra, sums_y, bn_w, st1, d_bn_w, d_bn_b)
        else:
            coef1 = ops.bn_bwd_finalize(partials, count, 0, 1, bn_w, st1, d_bn_w, d_bn_b)
        coef2 = None
        if has_res:
            d_rbn_w = torch.zeros_like(rbn_w)
            d_rbn_b = torch.zeros_like(rbn_w)
            coef2 = ops.bn_bwd_finalize(partials, count, 0, 2, rbn_w, st2, d_rbn_w, d_rbn_b)
        dy, dr = ops.bn_act_bwd_apply(dout, out, y, r, gate, extra, coef1, coef2, lengths, act)

        # pointwise conv: weight grads (split-K tcgen05 MN-major GEMM) and data grads (same NT kernel, W^T)
        d_pw = ops.pwconv_wgrad(dy, d).view(Cout, Cin, 1)
        pw_t = ops.cast_weight(pw_w.view(Cout, Cin), dt, transpose=True)  # [Cin, Cout]
        dd, _ = ops.pwconv_fwd(dy, pw_t)
        d_res = None
        dxr = None
        d_res_x = None
        need_dx = ctx.needs_input_grad[0]
        if has_res:
            rin = x if res_x is None else res_x
            Cres = rin.shape[-1]
            d_res = ops.pwconv_wgrad(dr, rin).view(Cout, Cres, 1)
            if (res_x is None and need_dx) or (res_x is not None and ctx.needs_input_grad[1]):
                res_t = ops.cast_weight(res_w.view(Cout, Cres), dt, transpose=True)
                dxr, _ = ops.pwconv_fwd(dr, res_t)
                if res_x is not None:
                    d_res_x, dxr = dxr, None
        # depthwise conv
        dwt = ops.dwconv_wgrad(x, dd, K, stride=stride)  # [K, Cin]
        d_dw = ops.cast_weight(dwt, torch.float32, transpose=True).view(Cin, 1, K)
        dx = None
        if need_dx:
            if stride != 1:
                raise RuntimeError("lightning_asr_b200: data gradient of the stride-2 first conv is not needed/implemented")
            dx = ops.dwconv_fwd(dd, wt, stride=1, flip=True, addend=dxr)
        return (dx, d_res_x, None, d_dw, d_pw, d_bn_w, d_bn_b, d_res, d_rbn_w, d_rbn_b, d_se1, d_se2, None, None, None,
                None, None)


class Conv1x1BNReLUFn(torch.autograd.Function):
    """last_cnn2: relu(BN(conv1x1(x)))  (no mask)   models/QuartNet.py:145-149; relu=False gives the bare
    residual branch conv1x1 -> BN (models/QuartNet.py:62-65) used by the unfused dropout path."""

    @staticmethod
    def forward(ctx, x, w, bn_w, bn_b, bn_buffers, training, relu=True):
        dt = x.dtype
        N, T, Cin = x.shape
        Cout = w.shape[0]
        w_s = ops.cast_weight(w.view(Cout, Cin), dt)
        y, stats = ops.pwconv_fwd(x, w_s, want_stats=training)
        if training:
            st = ops.bn_finalize(stats, N * T, bn_w, bn_b, bn_buffers[0], bn_buffers[1], BN_EPS, BN_MOMENTUM)
            _bump(bn_buffers[2])
        else:
            st = ops.bn_eval_coeffs(bn_w, bn_b, bn_buffers[0], bn_buffers[1], BN_EPS)
        act = ACT_RELU if relu else ACT_NONE
        out = ops.bn_apply_act(y, st, act=act)
        ctx.saved = (x, y, out, st)
        ctx.params = (w, bn_w)
        ctx.training = training
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, out, st = ctx.saved
        w, bn_w = ctx.params
        if not ctx.training:
            raise RuntimeError("lightning_asr_b200: backward through eval-mode BatchNorm is not implemented")
        dt = x.dtype
        dout = dout.contiguous()
        N, T, Cin = x.shape
        Cout = w.shape[0]
        act = ctx.act
        partials, _ = ops.bn_act_bwd_reduce(dout, out, y, None, act)
        d_bn_w = torch.zeros_like(bn_w)
        d_bn_b = torch.zeros_like(bn_w)
        coef = ops.bn_bwd_finalize(partials, N * T, 0, 1, bn_w, st, d_bn_w, d_bn_b)
        dy, _ = ops.bn_act_bwd_apply(dout, out, y, None, None, None, coef, None, None, act)
        d_w = ops.pwconv_wgrad(dy, x).view(Cout, Cin, 1)
        dx = None
        if ctx.needs_input_grad[0]:
            w_t = ops.cast_weight(w.view(Cout, Cin), dt, transpose=True)
            dx, _ = ops.pwconv_fwd(dy, w_t)
        return dx, d_w, d_bn_w, d_bn_b, None, None, None


def _pad8(v):
    return (v + 7) // 8 * 8


def _decoder_logits(x, w, b):
    """logits [N, T, ld] with ld = V rounded up to 8 (16-byte rows for TMA); padded columns are exactly 0."""
    dt = x.dtype
    V, Cin = w.shape[0], w.shape[1]
    ld = _pad8(V)
    w_s = ops.cast_weight(w.view(V, Cin), dt, pad_rows=ld)
    bias = torch.zeros((ld,), device=x.device, dtype=torch.float32)
    bias[:V] = b
    logits, _ = ops.pwconv_fwd(x, w_s, bias=bias)
    return logits


def _decoder_backward(x, w, dlogits, need_dx):
    """dlogits [N, T, ld] (padded columns zero) -> dx, dw [V, Cin, 1], db [V]."""
    dt = x.dtype
    V, Cin = w.shape[0], w.shape[1]
    ld = dlogits.shape[-1]
    dw_pad = ops.pwconv_wgrad(dlogits, x)  # [ld, Cin]
    d_w = dw_pad[:V].reshape(V, Cin, 1)
    # bias gradient = column sums of dlogits: reuse the pointwise wgrad against a ones column
    ones = torch.ones(x.shape[:-1] + (8,), device=x.device, dtype=dt)
    d_b = ops.pwconv_wgrad(dlogits, ones)[:V, 0].contiguous()
    dx = None
    if need_dx:
        w_t = torch.zeros((Cin, ld), device=x.device, dtype=dt)
        w_t[:, :V] = w.view(V, Cin).t().to(dt)
        dx, _ = ops.pwconv_fwd(dlogits, w_t)
    return dx, d_w, d_b


class DecoderLogSoftmaxFn(torch.autograd.Function):
    """log_softmax(conv1x1(x) + b) -> [N, T, V] fp32   models/QuartNet.py:282-290"""

    @staticmethod
    def forward(ctx, x, w, b):
        V = w.shape[0]
        logits = _decoder_logits(x, w, b)
        _, lp = ops.log_softmax_fwd(logits, V, want_lp=True)
        ctx.saved = (x, lp, logits.shape[-1])
        ctx.params = (w,)
        return lp

    @staticmethod
    def backward(ctx, dlp):
        x, lp, ld = ctx.saved
        (w,) = ctx.params
        dlogits = ops.log_softmax_bwd(dlp.contiguous().float(), lp, ld, x.dtype)
        dx, d_w, d_b = _decoder_backward(x, w, dlogits, ctx.needs_input_grad[0])
        return dx, d_w, d_b


def _as_ntv(log_probs):
    """Accept torch.nn.CTCLoss's [T, N, V] argument (usually `out.transpose(0, 1)`, a strided view of a contiguous
    [N, T, V] tensor) and return a contiguous [N, T, V] tensor without copying when possible."""
    if log_probs.dim() != 3:
        raise ValueError("log_probs must be [T, N, C]")
    ntv = log_probs.transpose(0, 1)
    return ntv if ntv.is_contiguous() else ntv.contiguous()


class CTCLossFn(torch.autograd.Function):
    """torch.nn.CTCLoss(blank, reduction='none', zero_infinity=False) on [T, N, V] log-probs -> nll [N]."""

    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank):
        x = _as_ntv(log_probs)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        V = x.shape[-1]
        need_grad = log_probs.requires_grad
        nll, alpha, beta = ops.ctc_fwd(x, None, targets, input_lengths, target_lengths, V, blank, want_beta=need_grad)
        ctx.saved = (x, targets, input_lengths, target_lengths, alpha, beta, nll)
        ctx.blank = blank
        ctx.in_dtype = log_probs.dtype
        return nll

    @staticmethod
    def backward(ctx, gout):
        x, targets, il, tl, alpha, beta, nll = ctx.saved
        V = x.shape[-1]
        grad = ops.ctc_bwd(x, None, targets, il, tl, alpha, beta, nll, gout.contiguous().float(), V, ctx.blank, V,
                           torch.float32)
        return grad.transpose(0, 1).to(ctx.in_dtype), None, None, None, None


class FusedDecoderCTCFn(torch.autograd.Function):
    """nll [N] = CTC(log_softmax(conv1x1(x) + b)) without materialising log-probs; backward emits d(logits) =
    (softmax - occupancy) * g directly."""

    @staticmethod
    def forward(ctx, x, w, b, targets, input_lengths, target_lengths, blank):
        V = w.shape[0]
        logits = _decoder_logits(x, w, b)
        lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
        nll, alpha, beta = ops.ctc_fwd(logits, lse, targets, input_lengths, target_lengths, V, blank, want_beta=True)
        ctx.saved = (x, logits, lse, targets, input_lengths, target_lengths, alpha, beta, nll)
        ctx.params = (w,)
        ctx.blank = blank
        ctx.mark_non_differentiable(logits)
        return nll, logits

    @staticmethod
    def backward(ctx, gout, _glogits):
        x, logits, lse, targets, il, tl, alpha, beta, nll = ctx.saved
        (w,) = ctx.params
        V = w.shape[0]
        ld = logits.shape[-1]
        dlogits = ops.ctc_bwd(logits, lse, targets, il, tl, alpha, beta, nll, gout.contiguous().float(), V, ctx.blank,
                              ld, x.dtype)
        dx, d_w, d_b = _decoder_backward(x, w, dlogits, ctx.needs_input_grad[0])
        return dx, d_w, d_b, None, None, None, None
