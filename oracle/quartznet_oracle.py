"""CPU oracle (test infrastructure, see oracle/__init__.py): the QuartzNet-style encoder of the reference,
restated as pure functions over a state_dict.

Follows, line by line:
  models/QuartNet.py           SeprationConv.forward :29-39, channel_shuffle(groups=1) :41-52 (identity),
                               QuartNetBlock.forward :71-78, QuartNet12.forward :152-173, MyModel2.forward :280-291,
                               MaskCNN.forward :309-321
  models/QuartNetContext.py    BiLSTM splice :171-173, BatchLSTM.forward :194-199, block3 in=336 :143, block6 :150,:182
  models/QuartNetContextSE.py  SELayer.forward :19-23, SE between BN and ReLU :55

It never instantiates the reference's classes: parameters come in as a plain dict with the reference's
state_dict keys (the checkpoint schema), so it also documents that schema.  Works in fp32 and
fp64 (pass a .double() state dict and inputs) -- SURVEY.md section 10 uses the fp64 run as the yardstick.

Pinned by tests/test_oracle_golden.py against fixtures generated from the reference's own modules.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # nn.BatchNorm1d(out_ch, eps=1e-3): models/QuartNet.py:24,64,147
BN_MOMENTUM = 0.1

ASR13X1_BLOCKS = [  # models/QuartNet.py:130-144
    ("block1", 33), ("block12", 33), ("block13", 33), ("block2", 39), ("block22", 39), ("block23", 39),
    ("block3", 51), ("block32", 51), ("block33", 51), ("block4", 63), ("block42", 63), ("block43", 63),
    ("block5", 75),
]


def mask_lengths(T, percents):
    """models/QuartNet.py:311 -- torch.mul(x.size(2), percents).int(): fp32 product, truncation toward zero."""
    return torch.mul(T, percents.float()).int()


def mask_cnn(x, percents):
    """models/QuartNet.py:309-321: zero x[i, :, len_i:] with len_i = int(T * p_i)."""
    lengths = mask_lengths(x.size(2), percents)
    t = torch.arange(x.size(2), device=x.device)
    keep = t[None, :] < lengths[:, None].to(x.device)
    return x * keep[:, None, :].to(x.dtype)


def batch_norm(x, sd, prefix, training, update_buffers):
    """nn.BatchNorm1d(eps=1e-3): batch statistics (biased var) in training; running statistics in eval.
    Running stats (momentum 0.1, unbiased var) are updated in `sd` when update_buffers is set."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        mean = x.mean(dim=(0, 2))
        var = x.var(dim=(0, 2), unbiased=False)
        if update_buffers:
            n = x.size(0) * x.size(2)
            with torch.no_grad():
                rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach().to(rm.dtype))
                rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * (var.detach() * n / max(n - 1, 1)).to(rv.dtype))
                key = prefix + ".num_batches_tracked"
                if key in sd:
                    sd[key] += 1
    else:
        mean, var = rm.to(x.dtype), rv.to(x.dtype)
    xhat = (x - mean[None, :, None]) / torch.sqrt(var[None, :, None] + BN_EPS)
    return xhat * w[None, :, None] + b[None, :, None]


def se_layer(x, sd, prefix):
    """models/QuartNetContextSE.py:19-23: mean over ALL T -> Linear -> ReLU -> Linear -> Sigmoid -> scale."""
    y = x.mean(dim=2)
    y = torch.relu(y @ sd[prefix + ".fc.0.weight"].t())
    y = torch.sigmoid(y @ sd[prefix + ".fc.2.weight"].t())
    return x * y[:, :, None]


def sep_conv(x, percents, sd, prefix, last, mask, stride, training, update_buffers, drop_mask=None, relu_mask=None):
    """models/QuartNet.py:29-39 (+ models/QuartNetContextSE.py:55).  relu_mask: parity hook, see block()."""
    dw = sd[prefix + ".depthwise_conv.weight"]
    k = dw.shape[-1]
    x = F.conv1d(x, dw, stride=stride, padding=k // 2, groups=dw.shape[0])
    x = F.conv1d(x, sd[prefix + ".pointwise_conv.weight"])
    # channel_shuffle(groups=1) is the identity permutation (:41-52)
    if mask:
        x = mask_cnn(x, percents)
    x = batch_norm(x, sd, prefix + ".bn", training, update_buffers)
    if (prefix + ".se.fc.0.weight") in sd:
        x = se_layer(x, sd, prefix + ".se")
    if not last:
        x = torch.relu(x) if relu_mask is None else x * relu_mask
    if drop_mask is not None:  # dropout with an externally supplied keep-mask / (1-p) scale (parity hook)
        x = x * drop_mask
    return x


def block(x, percents, sd, prefix, mask, training, update_buffers, relu_mask=None, drop_masks=None,
          inner_relu_masks=None):
    """models/QuartNet.py:71-78, any `repeat` (the :60 quirk -- `mask` lands in the inner seps' `last` slot -- is
    pinned by tests/test_oracle_golden.py::test_block_repeat_quirk_matches_reference and exercised on the GPU by
    tests/test_parity_gpu.py).
    relu_mask (parity hook, like drop_mask): replaces the final ReLU's own gate by a supplied 0/1 tensor so a
    reduced-precision implementation can be checked on an identical gating pattern.
    inner_relu_masks: the same hook for the ReLUs of the inner SeprationConvs (repeat > 1, present iff not mask).
    drop_masks: optional list (one per SeprationConv of the block) of dropout factors [N, C, T] (0 or 1/(1-p))."""
    start = x
    i = 0
    while (prefix + f".seq.{i + 1}.depthwise_conv.weight") in sd:
        # inner seps: constructed as SeprationConv(in, in, k, mask, ...) -> last=mask, mask=True (:60)
        x = sep_conv(x, percents, sd, prefix + f".seq.{i}", last=bool(mask), mask=True, stride=1, training=training,
                     update_buffers=update_buffers, drop_mask=None if drop_masks is None else drop_masks[i],
                     relu_mask=None if inner_relu_masks is None else inner_relu_masks[i])
        i += 1
    x = sep_conv(x, percents, sd, prefix + f".seq.{i}", last=True, mask=mask, stride=1, training=training,
                 update_buffers=update_buffers, drop_mask=None if drop_masks is None else drop_masks[i])
    r = F.conv1d(start, sd[prefix + ".reside.0.weight"])
    r = batch_norm(r, sd, prefix + ".reside.1", training, update_buffers)
    if relu_mask is not None:
        return (x + r) * relu_mask
    return torch.relu(x + r)


def context_lstm(x, percents, sd, prefix):
    """models/QuartNetContext.py:171-173,194-199: pack -> BiLSTM(256->40) -> pad -> concat on channels."""
    length = (x.size(2) * percents).int().cpu()
    hidden = sd[prefix + ".rnn.weight_hh_l0"].shape[1]
    lstm = torch.nn.LSTM(x.size(1), hidden, num_layers=1, batch_first=True, bidirectional=True).to(x.dtype)
    with torch.no_grad():
        for name, p in lstm.named_parameters():
            p.copy_(sd[prefix + ".rnn." + name])
    # route gradients to the caller's tensors: functional call with the dict's tensors
    params = {name: sd[prefix + ".rnn." + name] for name, _ in lstm.named_parameters()}
    packed = torch.nn.utils.rnn.pack_padded_sequence(x.transpose(1, 2), enforce_sorted=False, lengths=length,
                                                     batch_first=True)
    out, _ = torch.func.functional_call(lstm, params, (packed,))
    c, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
    return torch.cat((x, c.transpose(1, 2)), dim=1)


def encoder(x, percents, sd, prefix="encoder", mask=False, training=True, update_buffers=False, taps=None,
            drop_masks=None):
    """QuartNet12.forward (models/QuartNet.py:152-173; Context variants models/QuartNetContext.py:163-184).
    x [N, 1, F, T] -> [N, 1024, T'].  `taps` (dict) collects intermediate activations by module name.
    drop_masks: dict module name ('first_cnn', block names, 'last_cnn2') -> dropout factors [N, C, T'] standing in
    for nn.Dropout(drop_rate) at models/QuartNet.py:38 and :149 (torch's random stream is not reproducible)."""
    dm = drop_masks or {}
    x = x.squeeze(dim=1)
    x = sep_conv(x, percents, sd, prefix + ".first_cnn", last=False, mask=mask, stride=2, training=training,
                 update_buffers=update_buffers, drop_mask=dm.get("first_cnn"))
    if taps is not None:
        taps["first_cnn"] = x
    names = [n for n, _ in ASR13X1_BLOCKS]
    has_ctx = (prefix + ".context_rnn.rnn.weight_ih_l0") in sd
    if has_ctx:
        names.append("block6")
    for name in names:
        x = block(x, percents, sd, prefix + "." + name, mask, training, update_buffers,
                  drop_masks=[dm[name]] if name in dm else None)
        if taps is not None:
            taps[name] = x
        if name == "block23" and has_ctx:
            x = context_lstm(x, percents, sd, prefix + ".context_rnn")
    x = F.conv1d(x, sd[prefix + ".last_cnn2.0.weight"])
    x = batch_norm(x, sd, prefix + ".last_cnn2.1", training, update_buffers)
    x = torch.relu(x)
    if "last_cnn2" in dm:
        x = x * dm["last_cnn2"]
    if taps is not None:
        taps["last_cnn2"] = x
    return x


def model(x, percents, sd, mask=False, training=True, update_buffers=False, taps=None, drop_masks=None):
    """MyModel2.forward (models/QuartNet.py:280-291): encoder -> decoder conv (bias) -> [N,T',V'] -> log_softmax."""
    h = encoder(x, percents, sd, "encoder", mask, training, update_buffers, taps, drop_masks)
    logits = F.conv1d(h, sd["decoder.weight"], sd["decoder.bias"])
    return F.log_softmax(logits.transpose(1, 2), dim=-1)
