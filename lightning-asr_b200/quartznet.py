"""Drop-in modules for the reference's three encoder variants.

Mirrors (same class names, constructor signatures, forward contract and state_dict key schema):
    models/QuartNet.py           SeprationConv :8-52, QuartNetBlock :55-78, QuartNet12 :120-173, MyModel2 :264-291
    models/QuartNetContext.py    QuartNet12 :125-184 (BiLSTM(256->40) splice :171-173, block6), BatchLSTM :186-199
    models/QuartNetContextSE.py  SELayer :8-23, SeprationConv :25-72 (SE between BN and ReLU, :55)

The nn.Conv1d / nn.BatchNorm1d / nn.Linear sub-modules are PARAMETER CONTAINERS only: they are created in the
reference's order so that (a) reference checkpoints load with strict=True and vice versa, (b) the default
initialisation under a given torch seed is identical.  Their forward() is never called: compute goes through the
hand-written sm_100a kernels in functions.py, on channels-last activations.  Only MyModel2 is a layout boundary:
it takes the reference's [N, 1, F, T] fp32 features and returns [N, T', V+1] fp32 log-probs.
"""
import os

import weakref

import torch
import torch.nn as nn

from . import _lib, ops
from .functions import (BiLstmFn, Conv1x1BNReLUFn, DecoderLogSoftmaxFn, FusedDecoderCTCFn, SepConvBNFn)

_PRECISION = {"default": "auto"}


def set_default_precision(p):
    """'bf16' (tcgen05 path), 'fp32' (exact-parity FFMA path) or 'auto' (bf16 under torch.autocast(bf16/fp16))."""
    if p not in ("auto", "bf16", "fp32"):
        raise ValueError(p)
    _PRECISION["default"] = p


def resolve_dtype(precision=None):
    p = precision or _PRECISION["default"]
    if p == "auto":
        if torch.is_autocast_enabled():
            return torch.bfloat16  # the reference's fp16 AMP (conf/conf.yaml:27-29) maps to bf16 on B200
        return torch.float32
    return torch.bfloat16 if p == "bf16" else torch.float32


def _bn_buffers(bn):
    return (bn.running_mean, bn.running_var, bn.num_batches_tracked)


_FOLD = {"enabled": os.environ.get("LASR_EVAL_FOLD", "0") == "1"}
# channel-major series companions between blocks (csrc/dwconv_cm.cu); LASR_CM=0 is the A/B switch back to the gather kernels
_CM = {"enabled": os.environ.get("LASR_CM", "1") != "0", "max_bytes": int(os.environ.get("LASR_CM_MAX_MB", "96")) << 20}


def set_eval_folding(on):
    """Inference fast path: fold eval-mode BatchNorm into the 1x1 conv and run conv -> mask -> BN ->
    (+ residual) -> ReLU as one GEMM epilogue (bf16, no-grad, non-SE blocks).  OPT-IN: validated against the unfused
    kernels and the oracle (tests/test_model_gpu.py), but measured slower on B200 (config 5: 11.6 -> 12.7 ms): the
    per-thread residual row reads of the GEMM epilogue cost more than the BatchNorm pass they replace."""
    _FOLD["enabled"] = bool(on)


def _fold_ok(module, x):
    return (_FOLD["enabled"] and not module.training and not torch.is_grad_enabled() and x.dtype == torch.bfloat16)


def _folded(conv, bn, cache_owner, tag):
    """(w' [Cout, Cin] bf16, bias [Cout] fp32) with eval-mode BatchNorm folded into the 1x1 conv; cached on the module
    and rebuilt when any of the tensors involved changed (in-place updates bump torch's version counters)."""
    key = (conv.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, conv.weight.data_ptr())
    cache = cache_owner.__dict__.setdefault("_fold_cache", {})
    hit = cache.get(tag)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    with torch.no_grad():
        scale = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
        shift = bn.bias.double() - bn.running_mean.double() * scale
        w = (conv.weight.double().view(conv.weight.shape[0], -1) * scale[:, None]).to(torch.bfloat16).contiguous()
        b = shift.float().contiguous()
    cache[tag] = (key, w, b)
    return w, b


class SELayer(nn.Module):
    """Parameter container for models/QuartNetContextSE.py:8-23 (fc.0: C->C/r, fc.2: C/r->C, no biases)."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel, bias=False),
            nn.Sigmoid(),
        )


class SeprationConv(nn.Module):
    """dw conv -> 1x1 conv -> (identity shuffle) -> MaskCNN -> BN -> [SE] -> ReLU(if not last) -> dropout."""

    def __init__(self, in_ch, out_ch, k=33, last=False, mask=True, dilation=1, stride=1, drop_rate=0.1, se=False,
                 _with_se=False):
        super().__init__()
        if dilation != 1:
            raise NotImplementedError("dilation > 1 is never used by the reference's in-scope models")
        self.last = last
        self.mask = mask
        self.k = k
        self.stride = stride
        self.depthwise_conv = nn.Conv1d(in_ch, in_ch, kernel_size=(k,), stride=(stride,), padding=(k // 2,),
                                        groups=in_ch, dilation=(dilation,), bias=False)
        self.pointwise_conv = nn.Conv1d(in_ch, out_ch, kernel_size=(1,), stride=(1,), bias=False)
        self.bn = nn.BatchNorm1d(out_ch, eps=1e-3)
        self.se = SELayer(out_ch, reduction=8) if _with_se else None  # ContextSE: `se=` kwarg is ignored (:46)
        self.drop_rate = drop_rate

    def _se_weights(self):
        if self.se is None:
            return None, None
        return self.se.fc[0].weight, self.se.fc[2].weight

    def forward(self, x, lengths, residual=None, res_x=None, drop_mask=None, next_k=None, next_w=None):
        """x [N, T, Cin] channels-last.  `residual` = (conv1x1, bn) of the enclosing block to fuse (then ReLU is
        applied after the add, models/QuartNet.py:75-77).  nn.Dropout(drop_rate) (:27,38) is fused into the apply pass;
        `drop_mask` (uint8 keep mask [N, T, Cout]) replaces the device-drawn mask (parity hook).
        next_k, next_w: kernel size and taps of the depthwise conv that consumes the result (None: unknown / not a
        depthwise conv).
        When given (bf16), the apply pass also writes the channel-major series companion of the output and attaches it
        as `out._lasr_series`; a depthwise conv whose input carries a matching companion reads it through TMA
        (csrc/dwconv_cm.cu) instead of gathering the series from the channels-last tensor."""
        lens = lengths if self.mask else None
        drop = (self.drop_rate, drop_mask) if (self.drop_rate > 0.0 and self.training) else None
        se1, se2 = self._se_weights()
        xs = getattr(x, "_lasr_series", None) if _CM["enabled"] else None
        # the companion pays (one more write stream in the apply pass) while the depthwise conv can read it back from
        # L2; at inference batch sizes (hundreds of MB per activation) everything streams through HBM and it loses
        small = x.shape[0] * x.shape[1] * self.pointwise_conv.out_channels * 2 <= _CM["max_bytes"]
        cm_out = [next_k, next_w] if (next_k is not None and _CM["enabled"] and small) else None
        if self.se is None and _fold_ok(self, x):
            # eval fast path: dw conv -> [residual GEMM + bias] -> ONE GEMM with the whole block epilogue
            d = ops.dwconv_fwd(x, self.depthwise_conv.weight.detach(), stride=self.stride)
            w1, b1 = _folded(self.pointwise_conv, self.bn, self, "pw")
            r = None
            if residual is not None:
                rconv, rbn = residual
                w2, b2 = _folded(rconv, rbn, self, "res")
                r = ops.pwconv_fwd(x if res_x is None else res_x, w2, bias=b2)
            relu = True if residual is not None else (not self.last)
            return ops.pwconv_fwd_fused(d, w1, b1, residual=r, lengths=lens, T=d.shape[1], relu=relu)
        if residual is not None:
            rconv, rbn = residual
            out = SepConvBNFn.apply(x, res_x, lens, self.depthwise_conv.weight, self.pointwise_conv.weight,
                                    self.bn.weight, self.bn.bias, rconv.weight, rbn.weight, rbn.bias, se1, se2,
                                    _bn_buffers(self.bn), _bn_buffers(rbn), self.stride, True, self.training, drop, xs,
                                    cm_out)
        else:
            out = SepConvBNFn.apply(x, None, lens, self.depthwise_conv.weight, self.pointwise_conv.weight,
                                    self.bn.weight, self.bn.bias, None, None, None, se1, se2, _bn_buffers(self.bn), None,
                                    self.stride, not self.last, self.training, drop, xs, cm_out)
        if cm_out and cm_out[0] is not None:
            out._lasr_series = cm_out[0]
        return out


class QuartNetBlock(nn.Module):
    def __init__(self, repeat=3, in_ch=1, out_ch=32, k=33, mask=True, drop_rate=0., _with_se=False):
        super().__init__()
        seq = []
        for _ in range(0, repeat - 1):
            # reference quirk (models/QuartNet.py:60): `mask` lands in the positional `last` slot, so inner seps get
            # last=mask and the constructor default mask=True
            seq.append(SeprationConv(in_ch, in_ch, k, mask, drop_rate=drop_rate, _with_se=_with_se))
        self.reside = nn.Sequential(
            nn.Conv1d(in_ch, out_ch, kernel_size=(1,), bias=False),
            nn.BatchNorm1d(out_ch, eps=1e-3),
        )
        seq.append(SeprationConv(in_ch, out_ch, k=k, last=True, mask=mask, drop_rate=drop_rate, _with_se=_with_se))
        self.seq = nn.ModuleList(seq)
        self.drop_rate = drop_rate

    def forward(self, x, lengths, drop_masks=None, next_k=None, next_w=None):
        """drop_masks: optional list of uint8 keep masks, one per SeprationConv of `seq` (parity hook).
        next_k: kernel size of the depthwise conv that follows the block (see SeprationConv.forward)."""
        start = x
        for i, m in enumerate(self.seq[:-1]):
            x = m(x, lengths, drop_mask=None if drop_masks is None else drop_masks[i], next_k=self.seq[i + 1].k,
                  next_w=self.seq[i + 1].depthwise_conv.weight)
        last = self.seq[-1]
        # dropout sits between BN and the residual add (models/QuartNet.py:38,76): the apply pass does both
        return last(x, lengths, residual=(self.reside[0], self.reside[1]), res_x=None if x is start else start,
                    drop_mask=None if drop_masks is None else drop_masks[-1], next_k=next_k, next_w=next_w)


class BatchLSTM(nn.Module):
    """models/QuartNetContext.py:186-199.  `self.rnn` is a plain nn.LSTM and stays the owner of the parameters, so the
    checkpoint keys (`context_rnn.rnn.weight_ih_l0`, ..., `..._reverse`) are the reference's; the arithmetic runs in
    functions.BiLstmFn (one input-projection GEMM + the recurrence kernels of csrc/lstm.cu), channels-last, with the
    utterance lengths read on the device: no `.cpu()` sync, no pack / pad copies, CUDA-graph capturable.

    forward(x [N, T, in_ch], length int32 [N] on the device) -> (c [N, T, 2*out_ch], None); frames >= length are zero
    exactly as pad_packed_sequence leaves them."""

    def __init__(self, in_ch=128, out_ch=128, batch_first=True, bidirection=True, num_layers=1, dropout=0.):
        super().__init__()
        if not (batch_first and bidirection and num_layers == 1 and dropout == 0. and out_ch == 40):
            raise NotImplementedError("BatchLSTM: the B200 kernel implements the reference's only configuration "
                                      "(batch_first, bidirectional, 1 layer, hidden 40; QuartNetContext.py:157)")
        self.batch_first = batch_first
        self.rnn = nn.LSTM(in_ch, out_ch, num_layers=num_layers, batch_first=batch_first, bidirectional=bidirection,
                           dropout=dropout)

    def forward(self, x, length, total_length=None):
        r = self.rnn
        c = BiLstmFn.apply(x.contiguous(), length, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0,
                           r.weight_ih_l0_reverse, r.weight_hh_l0_reverse, r.bias_ih_l0_reverse, r.bias_hh_l0_reverse)
        return c, None


_ASR13X1 = [  # (name, in, out, k)   models/QuartNet.py:130-144
    ("block1", 256, 256, 33), ("block12", 256, 256, 33), ("block13", 256, 256, 33),
    ("block2", 256, 256, 39), ("block22", 256, 256, 39), ("block23", 256, 256, 39),
    ("block3", 256, 512, 51), ("block32", 512, 512, 51), ("block33", 512, 512, 51),
    ("block4", 512, 512, 63), ("block42", 512, 512, 63), ("block43", 512, 512, 63),
    ("block5", 512, 512, 75),
]


class QuartNet12(nn.Module):
    """variant: 'base' (models/QuartNet.py), 'context' (QuartNetContext.py), 'contextse' (QuartNetContextSE.py)."""

    def __init__(self, drop_rate=0., mask=False, in_c=64, variant="base"):
        super().__init__()
        if variant not in ("base", "context", "contextse"):
            raise ValueError(variant)
        self.variant = variant
        se = variant == "contextse"
        ctx = variant != "base"
        self.first_cnn = SeprationConv(in_ch=in_c, out_ch=256, k=33, last=False, mask=mask, stride=2,
                                       drop_rate=drop_rate, _with_se=se)
        self.block_names = []
        for name, cin, cout, k in _ASR13X1:
            if ctx and name == "block3":
                cin = 336  # 256 + 2*40 context channels (models/QuartNetContext.py:143)
            setattr(self, name, QuartNetBlock(repeat=1, in_ch=cin, out_ch=cout, k=k, mask=mask, drop_rate=drop_rate,
                                              _with_se=se))
            self.block_names.append(name)
        if ctx:
            self.block6 = QuartNetBlock(repeat=1, in_ch=512, out_ch=512, k=87, mask=mask, drop_rate=drop_rate,
                                        _with_se=se)
            self.block_names.append("block6")
        self.last_cnn2 = nn.Sequential(
            nn.Conv1d(512, 1024, kernel_size=(1,), stride=(1,), bias=False),
            nn.BatchNorm1d(1024, eps=1e-3),
            nn.ReLU(inplace=True),
            nn.Dropout(drop_rate),
        )
        if ctx:
            self.context_rnn = BatchLSTM(in_ch=256, out_ch=40, batch_first=True, bidirection=True)
        self.drop_rate = drop_rate

    def forward_ntc(self, x, percents, drop_masks=None):
        """x [N, T, F] channels-last features -> [N, T', 1024] channels-last.
        drop_masks: optional dict module name -> uint8 keep mask [N, T', C] ('first_cnn', block names, 'last_cnn2'):
        the parity hook for drop_rate > 0 (torch's Philox stream cannot be reproduced)."""
        T_out = (x.shape[1] - 1) // 2 + 1
        lengths = ops.out_lengths(T_out, percents.to(x.device))
        dm = drop_masks or {}
        blocks = [getattr(self, name) for name in self.block_names]
        x = self.first_cnn(x, lengths, drop_mask=dm.get("first_cnn"), next_k=blocks[0].seq[0].k,
                           next_w=blocks[0].seq[0].depthwise_conv.weight)
        for i, name in enumerate(self.block_names):
            # the next depthwise conv reads this block's output, unless the BiLSTM splice comes in between
            spliced = name == "block23" and self.variant != "base"
            nxt = blocks[i + 1].seq[0] if (i + 1 < len(blocks) and not spliced) else None
            x = blocks[i](x, lengths, drop_masks=[dm[name]] if name in dm else None,
                          next_k=None if nxt is None else nxt.k,
                          next_w=None if nxt is None else nxt.depthwise_conv.weight)
            if name == "block23" and self.variant != "base":
                # models/QuartNetContext.py:171-173: length = (T' * percents).int() is the same `lengths` tensor; it
                # stays on the device (the reference's `.cpu()` sync is gone)
                c, _ = self.context_rnn(x, lengths)
                x = torch.cat((x, c), dim=2).contiguous()
        if _fold_ok(self, x):
            w, b = _folded(self.last_cnn2[0], self.last_cnn2[1], self, "last_cnn2")
            x = ops.pwconv_fwd_fused(x, w, b, relu=True)
        else:
            drop = (self.drop_rate, dm.get("last_cnn2")) if (self.drop_rate > 0.0 and self.training) else None
            x = Conv1x1BNReLUFn.apply(x, self.last_cnn2[0].weight, self.last_cnn2[1].weight, self.last_cnn2[1].bias,
                                      _bn_buffers(self.last_cnn2[1]), self.training, True, drop)
        # the decoder / CTC side needs the same (T' * percents).int(): hand it over with the output it belongs to
        self._lengths_of = (weakref.ref(x), lengths)
        return x

    def forward(self, input, percents, precision=None, drop_masks=None):
        """input [N, 1, F, T] fp32 (reference layout) -> [N, T', 1024] channels-last (internal layout)."""
        _lib.require_device()
        dt = resolve_dtype(precision)
        feats = input.squeeze(dim=1).contiguous().float()
        return self.forward_ntc(ops.nct_to_ntc(feats, dt), percents, drop_masks=drop_masks)


class MyModel2(nn.Module):
    """MyModel2(labels, drop_rate=0., mask=False[, in_c=64]) -- models/QuartNet.py:264-291 and siblings.

    forward(input [N,1,F,T] float, percents [N] float) -> log_probs [N, T', len(labels)+1] fp32.
    """

    variant = "base"

    def __init__(self, labels, drop_rate=0., mask=False, in_c=64, precision=None):
        super().__init__()
        self.labels = labels
        self.precision = precision
        self.encoder = QuartNet12(drop_rate=drop_rate, mask=mask, in_c=in_c, variant=self.variant)
        self.decoder = nn.Conv1d(1024, len(self.labels) + 1, kernel_size=(1,))

    def encode(self, input, percents, drop_masks=None):
        return self.encoder(input, percents, precision=self.precision, drop_masks=drop_masks)

    def forward(self, input, percents, drop_masks=None):
        x = self.encode(input, percents, drop_masks=drop_masks)
        return DecoderLogSoftmaxFn.apply(x, self.decoder.weight, self.decoder.bias)

    def forward_fused_ctc(self, input, percents, targets, target_lengths):
        """Training fast path: encoder -> decoder -> log-softmax -> CTC in one chain, log-probs never materialised.
        Returns (nll [N], logits [N, T', ld] for decoding, t_lengths)."""
        x = self.encode(input, percents)
        return self.fused_ctc_from_encoded(x, percents, targets, target_lengths)

    def fused_ctc_from_encoded(self, x, percents, targets, target_lengths):
        held = getattr(self.encoder, "_lengths_of", None)
        if held is not None and held[0]() is x:  # computed by the encoder pass that produced x: two launches fewer
            t_lengths = held[1]
        else:
            t_lengths = ops.out_lengths(x.shape[1], percents.to(x.device))
        nll, logits = FusedDecoderCTCFn.apply(x, self.decoder.weight, self.decoder.bias,
                                              targets.to(x.device).long().contiguous(), t_lengths,
                                              target_lengths.to(x.device).int().contiguous(), len(self.labels))
        return nll, logits, t_lengths


class MyModel2Context(MyModel2):
    variant = "context"


class MyModel2ContextSE(MyModel2):
    variant = "contextse"


# model_name -> class.  In the reference `model_name` (conf/conf.yaml:1) is only a label and the model is chosen by a
# hard-coded import (train.py:13-14); this registry is the switch SURVEY.md 8b asks for.
MODEL_REGISTRY = {
    "asr12x1": MyModel2, "asr13x1": MyModel2, "quartnet": MyModel2,
    "asr13x1context": MyModel2Context, "quartnetcontext": MyModel2Context,
    "asr13x1contextse": MyModel2ContextSE, "quartnetcontextse": MyModel2ContextSE,
}


def build_model(model_name, labels, drop_rate=0., mask=False, in_c=64, precision=None):
    key = model_name.lower().replace("_", "").replace("-", "")
    if key not in MODEL_REGISTRY:
        raise KeyError(f"unknown model_name {model_name!r}; known: {sorted(MODEL_REGISTRY)}")
    return MODEL_REGISTRY[key](labels, drop_rate=drop_rate, mask=mask, in_c=in_c, precision=precision)
