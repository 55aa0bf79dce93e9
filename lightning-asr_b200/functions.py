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
import os

import torch

from . import ops, runtime
from .ops import ACT_NONE, ACT_RELU

BN_EPS = 1e-3  # nn.BatchNorm1d(out_ch, eps=1e-3)  models/QuartNet.py:24,64,147
BN_MOMENTUM = 0.1
RELU_BITS = os.environ.get("LASR_RELU_BITS", "1") != "0"  # A/B switch: backward passes read `out` again


def _stats(C, device):
    return runtime.zeros((2, C), torch.float64, device)


def _make_drop(drop, shape, device, training):
    """drop = None | (p, mask_or_None): -> ops.Dropout for the forward pass (None when inactive).
    A supplied uint8 keep-mask [N, T, C] is applied as is (parity hook); otherwise the forward kernel draws one."""
    if drop is None or not training or drop[0] <= 0.0:
        return None
    p, mask = drop[0], drop[1]
    if mask is not None:
        return ops.Dropout(mask.contiguous(), p, "read")
    seed, seed_dev = runtime.next_dropout_stream()
    return ops.Dropout(torch.empty(shape, device=device, dtype=torch.uint8), p, "generate", seed, seed_dev)


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
      drop = None | (p, keep_mask uint8 [N, T, Cout] or None): nn.Dropout(p) on the normalised branch, fused into the
             apply pass (before the residual add, models/QuartNet.py:38,76); a supplied mask is the parity hook
    Kernel sequence (training): dwconv -> pw GEMM (+mask, +BN sums) [-> residual GEMM (+BN sums)] [-> SE squeeze /
    excite] -> ONE apply pass.  Backward: reduce -> apply -> pw wgrad, pw dgrad [-> res wgrad, res dgrad] -> dw wgrad,
    dw dgrad (+ residual dgrad fused as addend).
    """

    @staticmethod
    def forward(ctx, x, res_x, lengths, dw_w, pw_w, bn_w, bn_b, res_w, rbn_w, rbn_b, se_w1, se_w2, bn_buffers,
                rbn_buffers, stride, relu, training, drop=None, xs=None, cm_out=None):
        """xs: ops.Series companion of x (channel-major, written by the pass that produced x) or None;
        cm_out: None | [K_next, taps_next]: ask the apply pass for the Series companion of the OUTPUT, laid out for a
        depthwise conv of kernel size K_next (taps given: its Toeplitz factors are built too); the Series replaces the
        list's first element (autograd Functions return tensors only)."""
        dt = x.dtype
        dev = x.device
        N, T_in, Cin = x.shape
        K = dw_w.shape[-1]
        Cout = pw_w.shape[0]
        act = ACT_RELU if relu else ACT_NONE
        if xs is not None and (stride != 1 or xs.K != K or (xs.N, xs.T, xs.C) != (N, T_in, Cin)):
            xs = None
        if xs is not None:
            d = ops.dwconv_fwd_cm(xs, dw_w.detach())  # TMA-fed: the series already exist in memory
        else:
            d = ops.dwconv_fwd(x, dw_w.detach(), stride=stride)
        T = d.shape[1]
        pw_s = runtime.weight(pw_w, dt).view(Cout, Cin)
        sums1 = _stats(Cout, dev) if training else None
        has_res = res_w is not None
        r = res_s = bn2 = None
        if has_res:
            rin = x if res_x is None else res_x
            res_s = runtime.weight(res_w, dt).view(Cout, rin.shape[-1])
            sums2 = _stats(Cout, dev) if training else None
            if rin.shape == d.shape:  # the block's two 1x1 convs have the same shape: one grouped launch
                y, r = ops.pwconv_fwd2(d, pw_s, lengths, sums1, rin, res_s, None, sums2, T)
            else:
                y = ops.pwconv_fwd(d, pw_s, lengths=lengths, T=T, stats=sums1)
                r = ops.pwconv_fwd(rin, res_s, stats=sums2)
            bn2 = ops.BNForward(rbn_w.detach(), rbn_b.detach(), rbn_buffers[0], rbn_buffers[1], rbn_buffers[2], sums2)
        else:
            y = ops.pwconv_fwd(d, pw_s, lengths=lengths, T=T, stats=sums1)
        bn1 = ops.BNForward(bn_w.detach(), bn_b.detach(), bn_buffers[0], bn_buffers[1], bn_buffers[2], sums1)
        gate = s = hidden = sums_y = scale1 = shift1 = None
        se_side = True
        if se_w1 is not None:
            # the excitation needs BN1's coefficients before the apply pass: s = scale*mean_t(y) + shift
            scale1, shift1 = ops.bn_coeffs(bn1, N * T, BN_EPS, BN_MOMENTUM, side_effects=False)
            sums_y = ops.sum_over_time(y)
            s, hidden, gate = ops.se_excite_fwd(sums_y, scale1, shift1, T, se_w1.detach(), se_w2.detach())
        dr_ = _make_drop(drop, (N, T, Cout), dev, training)
        cm_k = cm_out[0] if cm_out else None
        cm_w = cm_out[1].detach() if (cm_out and len(cm_out) > 1 and cm_out[1] is not None) else None
        # the backward passes read the ReLU gate as one byte per 8 channels instead of the whole output tensor
        bits = ops.relu_bits_alloc(N, T, Cout, dev) if (training and relu and dr_ is None and RELU_BITS) else None
        if cm_k is not None and dr_ is None and ops.cm_supported(Cout, cm_k, dt):
            out, cm_out[0] = ops.bn_apply_act(y, bn1, r, bn2, gate, act, BN_EPS, BN_MOMENTUM, side_effects=se_side,
                                              cm_k=cm_k, relu_bits=bits, cm_w=cm_w)
        else:
            if cm_out:
                cm_out[0] = None
            out = ops.bn_apply_act(y, bn1, r, bn2, gate, act, BN_EPS, BN_MOMENTUM, side_effects=se_side, drop=dr_,
                                   relu_bits=bits)

        # save_for_backward (not attributes): holding `out` on ctx directly would create an uncollectable
        # node <-> tensor cycle and leak every step's activations
        ctx.save_for_backward(x, res_x, lengths, d, y, r, out, bn1.save, bn2.save if bn2 is not None else None, gate, s,
                              hidden, sums_y, scale1, shift1, pw_s, res_s, dw_w, pw_w, bn_w, bn_b, res_w, rbn_w, rbn_b,
                              se_w1, se_w2, dr_.mask if dr_ is not None else None, xs.t if xs is not None else None,
                              bits, xs.toep_flip if xs is not None else None)
        ctx.cfg = (stride, act, training, K, Cin, Cout, dr_.p if dr_ is not None else 0.0)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout):
        if dout is None:
            return (None,) * 20
        (x, res_x, lengths, d, y, r, out, save1, save2, gate, s, hidden, sums_y, scale1, shift1, pw_s, res_s, dw_w,
         pw_w, bn_w, bn_b, res_w, rbn_w, rbn_b, se_w1, se_w2, drop_mask, xs_t, bits, toep_flip) = ctx.saved_tensors
        stride, act, training, K, Cin, Cout, drop_p = ctx.cfg
        drop = ops.Dropout(drop_mask, drop_p, "read") if drop_mask is not None else None
        if not training:
            raise RuntimeError("lightning_asr_b200: backward through eval-mode BatchNorm is not implemented")
        dev = x.device
        dout = dout.contiguous()
        N, T, _ = y.shape
        has_res = r is not None
        g_bn_w, ret_bn_w = runtime.grad_sink(bn_w)
        g_bn_b, ret_bn_b = runtime.grad_sink(bn_b)
        g_rbn_w = ret_rbn_w = g_rbn_b = ret_rbn_b = None
        if has_res:
            g_rbn_w, ret_rbn_w = runtime.grad_sink(rbn_w)
            g_rbn_b, ret_rbn_b = runtime.grad_sink(rbn_b)
        totals = runtime.zeros((4 if drop is not None else 3, Cout), torch.float64, dev)
        per_n = runtime.zeros((N, 3, Cout), torch.float32, dev) if gate is not None else None
        ops.bn_act_bwd_reduce(dout, out, y, r, act, totals, per_n, drop=drop, relu_bits=bits)
        extra = coef1 = None
        ret_se1 = ret_se2 = None
        bn1_side = (bn_w.detach(), save1, g_bn_w, g_bn_b)
        if gate is not None:
            g_se1, ret_se1 = runtime.grad_sink(se_w1)
            g_se2, ret_se2 = runtime.grad_sink(se_w2)
            extra = ops.se_excite_bwd(per_n, scale1, shift1, T, se_w1.detach(), se_w2.detach(), s, hidden, gate,
                                      g_se1, g_se2)
            coef1 = ops.se_bn_bwd_finalize(per_n, N, T, gate, extra, sums_y, bn_w.detach(), save1, g_bn_w, g_bn_b)
            bn1_side = None
        bn2_side = (rbn_w.detach(), save2, g_rbn_w, g_rbn_b) if has_res else None
        dy, dr = ops.bn_act_bwd_apply(dout, out, y, r, gate, extra, totals, coef1, bn1_side, bn2_side, lengths, act,
                                      drop=drop, relu_bits=bits)
        runtime.grad_ready(bn_w, bn_b, rbn_w, rbn_b, se_w1, se_w2)

        # pointwise conv: weight gradient (split-K tcgen05 MN-major GEMM) and data gradient (W as MN-major B operand)
        g_pw, ret_pw = runtime.grad_sink(pw_w)
        ret_res = None
        dxr = None
        d_res_x = None
        need_dx = ctx.needs_input_grad[0]
        rin = g_res = None
        if has_res:
            rin = x if res_x is None else res_x
            g_res, ret_res = runtime.grad_sink(res_w)
        if has_res and rin.shape == d.shape:
            # the block's two weight gradients have the same shape: one grouped launch
            def _wgrads():
                ops.pwconv_wgrad2(dy, d, g_pw.view(Cout, Cin), dr, rin, g_res.view(Cout, Cin))
                runtime.grad_ready(pw_w, res_w)
            runtime.defer(_wgrads, dy, d, dr, rin)
        else:
            def _pw_wgrad():
                ops.pwconv_wgrad(dy, d, out=g_pw)
                runtime.grad_ready(pw_w)
            runtime.defer(_pw_wgrad, dy, d)
            if has_res:
                def _res_wgrad():
                    ops.pwconv_wgrad(dr, rin, out=g_res)
                    runtime.grad_ready(res_w)
                runtime.defer(_res_wgrad, dr, rin)
        want_dxr = has_res and ((res_x is None and need_dx) or (res_x is not None and ctx.needs_input_grad[1]))
        # series path (csrc/dwconv_cm.cu): the forward read x as a channel-major series; the data-gradient GEMMs write the
        # depthwise conv's upstream gradient (and the residual branch's gradient, its addend) in the same format
        series = (xs_t is not None and need_dx and stride == 1 and Cin % 128 == 0 and Cout <= 512 and Cout % 64 == 0
                  and (not want_dxr or (res_x is None and res_s.shape == pw_s.shape)))
        if series:
            xs = ops.Series(xs_t, x.shape[0], x.shape[1], Cin, K)
            if want_dxr:
                dd_s, dxr_s = ops.pwconv_dgrad_cm(dy, pw_s, K, dr, res_s)
            else:
                dd_s, dxr_s = ops.pwconv_dgrad_cm(dy, pw_s, K), None
            g_dw, ret_dw = runtime.grad_sink(dw_w)
            dx, _ = ops.dwconv_bwd_cm(xs, dd_s, dw_w.detach(), addend=dxr_s, out_dw=g_dw, toep_flip=toep_flip)
            runtime.grad_ready(dw_w)
            return (dx, d_res_x, None, ret_dw, ret_pw, ret_bn_w, ret_bn_b, ret_res, ret_rbn_w, ret_rbn_b, ret_se1,
                    ret_se2, None, None, None, None, None, None, None, None)
        if want_dxr and res_s.shape == pw_s.shape:
            dd, dxr = ops.pwconv_dgrad2(dy, pw_s, dr, res_s)  # both data gradients in one grouped launch
        else:
            dd = ops.pwconv_dgrad(dy, pw_s)
            if want_dxr:
                dxr = ops.pwconv_dgrad(dr, res_s)
        if want_dxr and res_x is not None:
            d_res_x, dxr = dxr, None
        # depthwise conv
        g_dw, ret_dw = runtime.grad_sink(dw_w)
        dx = None
        if need_dx and stride == 1:
            # data gradient (+ the residual branch's gradient as addend) and weight gradient in one launch
            dx, _ = ops.dwconv_bwd(x, dd, dw_w.detach(), addend=dxr, out_dw=g_dw)
            runtime.grad_ready(dw_w)
        else:
            def _dw_wgrad():
                ops.dwconv_wgrad(x, dd, K, stride=stride, out=g_dw)
                runtime.grad_ready(dw_w)
            runtime.defer(_dw_wgrad, x, dd)
            if need_dx:
                raise RuntimeError("lightning_asr_b200: data gradient of the stride-2 first conv is not needed/implemented")
        return (dx, d_res_x, None, ret_dw, ret_pw, ret_bn_w, ret_bn_b, ret_res, ret_rbn_w, ret_rbn_b, ret_se1, ret_se2,
                None, None, None, None, None, None, None, None)


class Conv1x1BNReLUFn(torch.autograd.Function):
    """last_cnn2: relu(BN(conv1x1(x)))  (no mask)   models/QuartNet.py:145-149; relu=False gives the bare
    residual branch conv1x1 -> BN (models/QuartNet.py:62-65) used by the unfused dropout path."""

    @staticmethod
    def forward(ctx, x, w, bn_w, bn_b, bn_buffers, training, relu=True, drop=None):
        dt = x.dtype
        N, T, Cin = x.shape
        Cout = w.shape[0]
        w_s = runtime.weight(w, dt).view(Cout, Cin)
        sums = _stats(Cout, x.device) if training else None
        y = ops.pwconv_fwd(x, w_s, stats=sums)
        bn = ops.BNForward(bn_w.detach(), bn_b.detach(), bn_buffers[0], bn_buffers[1], bn_buffers[2], sums)
        act = ACT_RELU if relu else ACT_NONE
        dr_ = _make_drop(drop, (N, T, Cout), x.device, training)  # nn.Dropout after the ReLU (models/QuartNet.py:149)
        bits = ops.relu_bits_alloc(N, T, Cout, x.device) if (training and relu and dr_ is None and RELU_BITS) else None
        out = ops.bn_apply_act(y, bn, act=act, eps=BN_EPS, momentum=BN_MOMENTUM, drop=dr_, relu_bits=bits)
        ctx.save_for_backward(x, y, out, bn.save, w_s, w, bn_w, bn_b, dr_.mask if dr_ is not None else None, bits)
        ctx.training = training
        ctx.act = act
        ctx.drop_p = dr_.p if dr_ is not None else 0.0
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, out, save, w_s, w, bn_w, bn_b, drop_mask, bits = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("lightning_asr_b200: backward through eval-mode BatchNorm is not implemented")
        dout = dout.contiguous()
        N, T, Cin = x.shape
        Cout = w.shape[0]
        act = ctx.act
        drop = ops.Dropout(drop_mask, ctx.drop_p, "read") if drop_mask is not None else None
        totals = runtime.zeros((4 if drop is not None else 3, Cout), torch.float64, x.device)
        ops.bn_act_bwd_reduce(dout, out, y, None, act, totals, drop=drop, relu_bits=bits)
        g_w, ret_w = runtime.grad_sink(bn_w)
        g_b, ret_b = runtime.grad_sink(bn_b)
        dy, _ = ops.bn_act_bwd_apply(dout, out, y, None, None, None, totals, None, (bn_w.detach(), save, g_w, g_b),
                                     None, None, act, drop=drop, relu_bits=bits)
        runtime.grad_ready(bn_w, bn_b)
        g_cw, ret_cw = runtime.grad_sink(w)

        def _wgrad():
            ops.pwconv_wgrad(dy, x, out=g_cw)
            runtime.grad_ready(w)
        runtime.defer(_wgrad, dy, x)
        dx = ops.pwconv_dgrad(dy, w_s) if ctx.needs_input_grad[0] else None
        return dx, ret_cw, ret_w, ret_b, None, None, None, None


def _pad8(v):
    return (v + 7) // 8 * 8


def _decoder_logits(x, w, b):
    """logits [N, T, ld] with ld = V rounded up to 8 (16-byte rows for TMA).  Only the first V columns are written /
    meaningful: the weight's missing rows are zero-filled by TMA, every consumer reads columns < V only."""
    dt = x.dtype
    V, Cin = w.shape[0], w.shape[1]
    w_s = runtime.weight(w, dt).view(V, Cin)
    logits = ops.pwconv_fwd(x, w_s, bias=b.detach(), ldy=_pad8(V))
    return logits, w_s


def _decoder_backward(x, w, b, w_s, dlogits, need_dx):
    """dlogits [N, T, ld] (columns >= V are zero) -> dx; accumulates dw [V, Cin, 1] and db [V]."""
    V = w.shape[0]
    g_w, ret_w = runtime.grad_sink(w)
    g_b, ret_b = runtime.grad_sink(b)

    def _wgrad():
        ops.pwconv_wgrad(dlogits, x, out=g_w, Cout=V)
        ops.colsum(dlogits, V, out=g_b)
        runtime.grad_ready(w, b)
    runtime.defer(_wgrad, dlogits, x)
    dx = ops.pwconv_dgrad(dlogits, w_s, lddy=dlogits.shape[-1]) if need_dx else None
    return dx, ret_w, ret_b


class DecoderLogSoftmaxFn(torch.autograd.Function):
    """log_softmax(conv1x1(x) + b) -> [N, T, V] fp32   models/QuartNet.py:282-290"""

    @staticmethod
    def forward(ctx, x, w, b):
        V = w.shape[0]
        logits, w_s = _decoder_logits(x, w, b)
        _, lp = ops.log_softmax_fwd(logits, V, want_lp=True)
        ctx.save_for_backward(x, lp, w_s, w, b)
        ctx.ld = logits.shape[-1]
        return lp

    @staticmethod
    def backward(ctx, dlp):
        x, lp, w_s, w, b = ctx.saved_tensors
        ld = ctx.ld
        dlogits = ops.log_softmax_bwd(dlp.contiguous().float(), lp, ld, x.dtype)
        return _decoder_backward(x, w, b, w_s, dlogits, ctx.needs_input_grad[0])


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
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank, zero_infinity=False):
        x = _as_ntv(log_probs)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        V = x.shape[-1]
        need_grad = log_probs.requires_grad
        nll, alpha, beta, scales = ops.ctc_fwd(x, None, targets, input_lengths, target_lengths, V, blank,
                                               want_beta=need_grad)
        ctx.save_for_backward(x, targets, input_lengths, target_lengths, alpha, beta, nll, scales)
        ctx.blank = blank
        ctx.in_dtype = log_probs.dtype
        ctx.zero_infinity = zero_infinity
        return nll

    @staticmethod
    def backward(ctx, gout):
        x, targets, il, tl, alpha, beta, nll, scales = ctx.saved_tensors
        V = x.shape[-1]
        grad = ops.ctc_bwd(x, None, targets, il, tl, alpha, beta, nll, gout.contiguous().float(), V, ctx.blank, V,
                           torch.float32, scales=scales)
        if ctx.zero_infinity:
            # torch.nn.CTCLoss(zero_infinity=True) zeroes the gradient of infeasible utterances (their lattice terms
            # are inf / NaN); the reference's own setting is zero_infinity=False (train.py:196)
            grad = torch.where(torch.isfinite(nll)[:, None, None], grad, torch.zeros_like(grad))
        return grad.transpose(0, 1).to(ctx.in_dtype), None, None, None, None, None


class FusedDecoderCTCFn(torch.autograd.Function):
    """nll [N] = CTC(log_softmax(conv1x1(x) + b)) without materialising log-probs; backward emits d(logits) =
    (softmax - occupancy) * g directly."""

    @staticmethod
    def forward(ctx, x, w, b, targets, input_lengths, target_lengths, blank):
        V = w.shape[0]
        logits, w_s = _decoder_logits(x, w, b)
        lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
        nll, alpha, beta, scales = ops.ctc_fwd(logits, lse, targets, input_lengths, target_lengths, V, blank,
                                               want_beta=True)
        ctx.save_for_backward(x, logits, lse, targets, input_lengths, target_lengths, alpha, beta, nll, w_s, w, b, scales)
        ctx.blank = blank
        ctx.mark_non_differentiable(logits)
        # without this autograd hands backward a materialised zero gradient for `logits`: a fill of the whole
        # [N, T', V'] tensor per step (58 us at the 4334-class vocabulary)
        ctx.set_materialize_grads(False)
        return nll, logits

    @staticmethod
    def backward(ctx, gout, _glogits):
        x, logits, lse, targets, il, tl, alpha, beta, nll, w_s, w, b, scales = ctx.saved_tensors
        if gout is None:  # nll unused downstream (grads are not materialised)
            gout = torch.zeros_like(nll)
        V = w.shape[0]
        ld = logits.shape[-1]
        dlogits = ops.ctc_bwd(logits, lse, targets, il, tl, alpha, beta, nll, gout.contiguous().float(), V, ctx.blank,
                              ld, x.dtype, scales=scales)
        dx, d_w, d_b = _decoder_backward(x, w, b, w_s, dlogits, ctx.needs_input_grad[0])
        return dx, d_w, d_b, None, None, None, None


class BiLstmFn(torch.autograd.Function):
    """Bidirectional single-layer LSTM over variable-length utterances, channels-last, no host sync:
        c[n, t, :] = [h_fwd(t) | h_reverse(t)] for t < lengths[n], zeros after          models/QuartNetContext.py:171-173,186-199
    = pack_padded_sequence(enforce_sorted=False) -> nn.LSTM(Cin, 40, bidirectional=True) -> pad_packed_sequence.
    Kernel sequence: ONE pointwise GEMM (Cin -> 320, bias b_ih + b_hh) for the input projections of all frames and both
    directions -> lasr_bilstm_fwd (recurrence).  Backward: lasr_bilstm_bwd (recurrence, dW_hh) -> weight-gradient GEMM
    (dW_ih), column sums (biases), data-gradient GEMM (dx).  Parameters are nn.LSTM's own (state_dict schema)."""

    @staticmethod
    def forward(ctx, x, lengths, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        dt, dev = x.dtype, x.device
        N, T, Cin = x.shape
        H = w_hh.shape[1]
        w_cat = torch.cat((runtime.weight(w_ih, dt), runtime.weight(w_ih_r, dt)), dim=0).contiguous()  # [8H, Cin]
        bias = torch.cat((b_ih.detach() + b_hh.detach(), b_ih_r.detach() + b_hh_r.detach()), dim=0).float().contiguous()
        whh = torch.stack((w_hh.detach(), w_hh_r.detach()), dim=0).float().contiguous()  # [2, 4H, H]
        pre = ops.pwconv_fwd(x, w_cat, bias=bias)  # [N, T, 8H]
        out, gates, cells = ops.bilstm_fwd(pre, whh, lengths, H)
        ctx.save_for_backward(x, lengths, w_cat, whh, out, gates, cells, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r,
                              b_hh_r)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x, lengths, w_cat, whh, out, gates, cells, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r,
         b_hh_r) = ctx.saved_tensors
        H = w_hh.shape[1]
        G = 4 * H
        dwhh = runtime.zeros((2, G, H), torch.float32, x.device)
        dpre = ops.bilstm_bwd(dout.contiguous(), out, gates, cells, whh, lengths, dwhh, H)  # [N, T, 2G]
        dwcat = runtime.zeros((2 * G, x.shape[-1]), torch.float32, x.device)
        ops.pwconv_wgrad(dpre, x, out=dwcat, Cout=2 * G)
        db = ops.colsum(dpre, 2 * G, out=runtime.zeros((2 * G,), torch.float32, x.device))
        rets = []
        for p, g in ((w_ih, dwcat[:G]), (w_hh, dwhh[0]), (b_ih, db[:G]), (b_hh, db[:G]),
                     (w_ih_r, dwcat[G:]), (w_hh_r, dwhh[1]), (b_ih_r, db[G:]), (b_hh_r, db[G:])):
            sink, ret = runtime.grad_sink(p)
            if ret is None:
                sink.add_(g.view_as(sink))  # flat-bucket view: accumulate in place, autograd gets nothing
                rets.append(None)
            else:
                rets.append(g.reshape(p.shape).clone())
        runtime.grad_ready(w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r)
        dx = ops.pwconv_dgrad(dpre, w_cat) if ctx.needs_input_grad[0] else None
        return (dx, None) + tuple(rets)
