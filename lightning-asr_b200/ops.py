"""Thin Python wrappers around the C ABI: allocate outputs with torch, pass raw device pointers.

Every function here launches hand-written sm_100a kernels on the current torch CUDA stream.  There is no
eager / CPU fallback: inputs must be CUDA tensors and the shared library must be loadable.

Activations are channels-last: [N, T, C] contiguous (C fastest).
"""
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


def cast_weight(w2d, dtype, transpose=False, pad_rows=0):
    """fp32 [R, C] master weight -> dtype shadow ([R, C] or [C, R]); optional zero rows appended (non-transposed)."""
    _chk(w2d, "w")
    R, C = w2d.shape
    if transpose:
        out = torch.empty((C, R), device=w2d.device, dtype=dtype)
    elif pad_rows > R:
        out = torch.zeros((pad_rows, C), device=w2d.device, dtype=dtype)
    else:
        out = torch.empty((R, C), device=w2d.device, dtype=dtype)
    call("lasr_cast_weight", w2d, out, R, C, 1 if transpose else 0, dtype_code(dtype))
    return out


def dw_out_len(T_in, K, stride):
    return (T_in + 2 * (K // 2) - K) // stride + 1


def dwconv_fwd(x, wt, stride=1, flip=False, addend=None):
    """x [N, T_in, C], wt [K, C] fp32 tap-major -> y [N, T_out, C]."""
    _chk(x, "x"), _chk(wt, "wt")
    N, T_in, C = x.shape
    K = wt.shape[0]
    T_out = dw_out_len(T_in, K, stride)
    y = torch.empty((N, T_out, C), device=x.device, dtype=x.dtype)
    call("lasr_dwconv1d_fwd", x, wt, y, addend, N, T_in, T_out, C, K, stride, 1 if flip else 0, dtype_code(x.dtype))
    return y


def dwconv_wgrad(x, dy, K, stride=1):
    """-> dwt [K, C] fp32 tap-major."""
    _chk(x, "x"), _chk(dy, "dy")
    N, T_in, C = x.shape
    T_out = dy.shape[1]
    dwt = torch.zeros((K, C), device=x.device, dtype=torch.float32)
    call("lasr_dwconv1d_wgrad", x, dy, dwt, N, T_in, T_out, C, K, stride, dtype_code(x.dtype))
    return dwt


def pwconv_fwd(x2d, w, bias=None, lengths=None, T=0, want_stats=False, out=None):
    """y[M, Cout] = x[M, Cin] w[Cout, Cin]^T (+bias), MaskCNN row mask, BN partial statistics.

    x2d may be a [N, T, Cin] tensor (flattened).  Returns (y, stats or None)."""
    _chk(x2d, "x"), _chk(w, "w")
    Cin = x2d.shape[-1]
    M = x2d.numel() // Cin
    Cout = w.shape[0]
    if w.shape[1] != Cin or w.dtype != x2d.dtype:
        raise _lib.LasrError(f"pwconv weight {tuple(w.shape)}/{w.dtype} does not match input Cin={Cin}/{x2d.dtype}")
    y = out if out is not None else torch.empty(x2d.shape[:-1] + (Cout,), device=x2d.device, dtype=x2d.dtype)
    stats = None
    if want_stats:
        groups = ((M + 127) // 128) * 4
        stats = torch.empty((groups, 2, Cout), device=x2d.device, dtype=torch.float32)
    call("lasr_pwconv_fwd", x2d, w, y, bias, lengths, T, stats, M, Cin, Cout, Cin, Cin, Cout, dtype_code(x2d.dtype))
    return y, stats


def pwconv_wgrad(dy, x, out=None):
    """dw[Cout, Cin] fp32 = dy[M, Cout]^T x[M, Cin]."""
    _chk(dy, "dy"), _chk(x, "x")
    Cout, Cin = dy.shape[-1], x.shape[-1]
    M = x.numel() // Cin
    dw = out if out is not None else torch.zeros((Cout, Cin), device=x.device, dtype=torch.float32)
    call("lasr_pwconv_wgrad", dy, x, dw, M, Cin, Cout, Cout, Cin, Cin, dtype_code(x.dtype))
    return dw


class BNState:
    """Per-call BatchNorm coefficients (all fp32 [C])."""

    __slots__ = ("mean", "invstd", "scale", "shift")

    def __init__(self, C, device):
        buf = torch.empty((4, C), device=device, dtype=torch.float32)
        self.mean, self.invstd, self.scale, self.shift = buf[0], buf[1], buf[2], buf[3]


def bn_finalize(stats, count, gamma, beta, running_mean, running_var, eps, momentum):
    groups, _, C = stats.shape
    st = BNState(C, stats.device)
    call("lasr_bn_finalize", stats, groups, C, count, eps, momentum, gamma, beta, st.mean, st.invstd, st.scale,
         st.shift, running_mean, running_var)
    return st


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps):
    C = gamma.numel()
    st = BNState(C, gamma.device)
    call("lasr_bn_eval_coeffs", gamma, beta, running_mean, running_var, eps, st.scale, st.shift, C)
    return st


def sum_over_time(y):
    N, T, C = y.shape
    sums = torch.empty((N, C), device=y.device, dtype=torch.float32)
    call("lasr_sum_over_time", y, sums, N, T, C, dtype_code(y.dtype))
    return sums


def bn_apply_act(y, st1, r=None, st2=None, gate=None, act=ACT_RELU):
    N, T, C = y.shape
    out = torch.empty_like(y)
    call("lasr_bn_apply_act_fwd", y, st1.scale, st1.shift, r, st2.scale if st2 else None,
         st2.shift if st2 else None, gate, out, N * T, C, T, act, dtype_code(y.dtype))
    return out


def bn_bwd_chunks(N, T):
    return _lib.load().lasr_bn_bwd_chunks(N, T)


def bn_act_bwd_reduce(dout, out, y, r, act):
    N, T, C = y.shape
    chunks = bn_bwd_chunks(N, T)
    partials = torch.empty((N * chunks, 3, C), device=y.device, dtype=torch.float32)
    call("lasr_bn_act_bwd_reduce", dout, out, y, r, partials, N, T, C, chunks, act, dtype_code(y.dtype))
    return partials, chunks


def bn_bwd_finalize(partials, count, idx_g, idx_gx, gamma, st, dgamma, dbeta):
    groups, nslots, C = partials.shape
    coef = torch.empty((3, C), device=partials.device, dtype=torch.float32)
    call("lasr_bn_bwd_finalize", partials, groups, nslots, C, count, idx_g, idx_gx, gamma, st.mean, st.invstd, dgamma,
         dbeta, coef)
    return coef


def bn_act_bwd_apply(dout, out, y, r, gate, extra, coef1, coef2, lengths, act):
    N, T, C = y.shape
    dy = torch.empty_like(y)
    dr = torch.empty_like(y) if r is not None else None
    call("lasr_bn_act_bwd_apply", dout, out, y, r, gate, extra, coef1, coef2, lengths, T, dy, dr, N * T, C, act,
         dtype_code(y.dtype))
    return dy, dr


def se_excite_fwd(sums, st, T, w1, w2):
    N, C = sums.shape
    Cr = w1.shape[0]
    s = torch.empty((N, C), device=sums.device, dtype=torch.float32)
    hidden = torch.empty((N, Cr), device=sums.device, dtype=torch.float32)
    gate = torch.empty((N, C), device=sums.device, dtype=torch.float32)
    call("lasr_se_excite_fwd", sums, st.scale, st.shift, T, w1, w2, s, hidden, gate, N, C, Cr)
    return s, hidden, gate


def se_excite_bwd(partials, chunks, st, T, w1, w2, s, hidden, gate):
    N, C = gate.shape
    Cr = w1.shape[0]
    extra = torch.empty((N, C), device=gate.device, dtype=torch.float32)
    dw1 = torch.zeros_like(w1)
    dw2 = torch.zeros_like(w2)
    call("lasr_se_excite_bwd", partials, chunks, st.scale, st.shift, T, w1, w2, s, hidden, gate, extra, dw1, dw2, N, C,
         Cr)
    return extra, dw1, dw2


def se_bn_bwd_finalize(partials, N, chunks, T, gate, extra, sums_y, gamma, st, dgamma, dbeta):
    C = gate.shape[1]
    coef = torch.empty((3, C), device=gate.device, dtype=torch.float32)
    call("lasr_se_bn_bwd_finalize", partials, N, chunks, C, T, gate, extra, sums_y, gamma, st.mean, st.invstd, dgamma,
         dbeta, coef)
    return coef


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


def ctc_fwd(x, lse, targets, input_lengths, target_lengths, V, blank, want_beta):
    """x [N, T, ld] log-probs (lse None) or logits (+lse [N,T]).  -> nll [N], alpha, beta|None."""
    N, T, ld = x.shape
    S_max = max(int(targets.shape[1]), 1) if targets.dim() == 2 else 1
    Lp = 2 * S_max + 1
    alpha = torch.empty((N, T, Lp), device=x.device, dtype=torch.float32)
    beta = torch.empty((N, T, Lp), device=x.device, dtype=torch.float32) if want_beta else None
    nll = torch.empty((N,), device=x.device, dtype=torch.float32)
    call("lasr_ctc_fwd", x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, N, T, V, ld, S_max, blank,
         dtype_code(x.dtype))
    return nll, alpha, beta


def ctc_bwd(x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, grad_out, V, blank, ldg, grad_dtype):
    N, T, ld = x.shape
    S_max = max(int(targets.shape[1]), 1)
    grad = torch.empty((N, T, ldg), device=x.device, dtype=grad_dtype)
    call("lasr_ctc_bwd", x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, grad_out, grad, N, T, V, ld,
         ldg, S_max, blank, dtype_code(x.dtype), dtype_code(grad_dtype))
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
