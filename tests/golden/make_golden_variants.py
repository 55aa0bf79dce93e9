"""Golden fixtures for the configurations the round-1 fixtures did not pin (run in the build container only):

    python tests/golden/make_golden_variants.py        -> tests/golden/variants.pt

Every entry is an output of the REFERENCE's own modules (imported from /root/reference) on seeded inputs:

  mask_false   models/QuartNet.py :: MyModel2(mask=False) -- the constructor default (:265): no MaskCNN anywhere.
  repeat2_*    models/QuartNet.py :: QuartNetBlock(repeat=2) with mask=True / mask=False: the inner SeprationConv is
               built as SeprationConv(in, in, k, mask) so `mask` lands in the positional `last` slot (:60).
  dropout_*    MyModel2(drop_rate=0.2) in train mode for the base and ContextSE variants.  torch's dropout stream is
               not reproducible across implementations, so nn.Dropout.forward is patched to multiply by a keep mask
               drawn from a seeded generator keyed by the CALL ORDER ("drop/<i>"); the placement of every dropout is
               still decided by the reference's own forward code.  Tests regenerate the same masks (golden_common.drop_factor).
  aishell      models/QuartNetContext.py :: MyModel2 with the 4333-symbol AISHELL vocabulary (BASELINE config 4):
               losses, gradient norms, a column subset of the log-probs.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from golden_common import LABELS28, aishell_labels, block_inputs, drop_factor, golden_weights, model_inputs  # noqa: E402
from make_golden import import_reference  # noqa: E402

DROP_P = 0.2


def _schema(m):
    return [(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()]


def _train_step(model, labels, x, percents, targets, tgt_len):
    """One training step of the reference module in fp64 (the fp64 run of the same reference code is the
    noise-free yardstick; whole-network fp32 gradients are only 3e-3..1e-2 accurate); results stored as fp32."""
    model.double()
    model.zero_grad()
    out = model(x.double(), percents)
    t_len = torch.mul(out.size(1), percents).int()
    nll = torch.nn.CTCLoss(blank=len(labels), reduction="none")(out.transpose(0, 1), targets, t_len, tgt_len)
    torch.mean(nll).backward()
    return out.detach().float(), nll.detach().float()


def make_mask_false(mods):
    model = mods["base"].MyModel2(LABELS28, drop_rate=0.0)  # mask defaults to False (models/QuartNet.py:265)
    schema = _schema(model)
    model.load_state_dict(golden_weights(schema), strict=True)
    model.train()
    out, nll = _train_step(model, LABELS28, *model_inputs())
    keep = ["encoder.first_cnn.pointwise_conv.weight", "encoder.block3.seq.0.bn.weight", "decoder.bias",
            "encoder.block5.seq.0.depthwise_conv.weight"]
    return {"schema": schema, "train_out": out, "train_nll": nll,
            "grad_norm": {k: float(p.grad.double().norm()) for k, p in model.named_parameters()},
            "grad": {k: p.grad.detach().float() for k, p in model.named_parameters() if k in keep}}


def make_repeat2(mods, mask):
    blk = mods["base"].QuartNetBlock(repeat=2, in_ch=64, out_ch=96, k=33, mask=mask, drop_rate=0.0)
    schema = _schema(blk)
    blk.load_state_dict(golden_weights(schema), strict=True)
    blk.train()
    x, percents, dout = block_inputs(64, 96)
    x = x.clone().requires_grad_(True)
    out = blk(x, percents)
    out.backward(dout)
    return {"schema": schema, "out": out.detach().clone(), "dx": x.grad.detach().clone(),
            "inner_last": bool(blk.seq[0].last), "inner_mask": bool(blk.seq[0].mask),
            "grad": {k: p.grad.detach().clone() for k, p in blk.named_parameters()}}


def make_dropout(mods, variant):
    model = mods[variant].MyModel2(LABELS28, drop_rate=DROP_P, mask=True)
    schema = _schema(model)
    model.load_state_dict(golden_weights(schema), strict=True)
    model.train()
    calls = []
    real = torch.nn.Dropout.forward

    def patched(self, inp):
        if not self.training or self.p == 0.0:
            return inp
        calls.append(tuple(inp.shape))
        return inp * drop_factor(len(calls) - 1, inp.shape, self.p).to(inp.dtype)

    torch.nn.Dropout.forward = patched
    try:
        out, nll = _train_step(model, LABELS28, *model_inputs())
    finally:
        torch.nn.Dropout.forward = real
    keep = ["encoder.first_cnn.bn.weight", "encoder.block4.seq.0.depthwise_conv.weight", "encoder.last_cnn2.1.bias",
            "encoder.block1.reside.0.weight"]
    return {"schema": schema, "p": DROP_P, "calls": calls, "train_out": out, "train_nll": nll,
            "grad_norm": {k: float(p.grad.double().norm()) for k, p in model.named_parameters()},
            "grad": {k: p.grad.detach().float() for k, p in model.named_parameters() if k in keep}}


def make_aishell(mods):
    labels = aishell_labels()
    model = mods["context"].MyModel2(labels, drop_rate=0.0, mask=True)
    schema = _schema(model)
    model.load_state_dict(golden_weights(schema), strict=True)
    model.train()
    x, percents, targets, tgt_len = model_inputs(n_labels=len(labels))
    out, nll = _train_step(model, labels, x, percents, targets, tgt_len)
    cols = torch.unique(torch.cat([targets.reshape(-1), torch.tensor([len(labels), 0, 1, 2, 4000])]))
    keep = ["decoder.bias", "encoder.context_rnn.rnn.weight_hh_l0", "encoder.block6.seq.0.depthwise_conv.weight"]
    return {"schema": schema, "train_nll": nll, "cols": cols, "train_out_cols": out[:, :, cols].clone(),
            "train_out_argmax": out.argmax(dim=-1),
            "grad_norm": {k: float(p.grad.double().norm()) for k, p in model.named_parameters()},
            "grad": {k: p.grad.detach().float() for k, p in model.named_parameters() if k in keep}}


if __name__ == "__main__":
    torch.set_num_threads(8)
    _, _, mods = import_reference()
    fx = {"mask_false": make_mask_false(mods), "repeat2_mask": make_repeat2(mods, True),
          "repeat2_nomask": make_repeat2(mods, False), "dropout_base": make_dropout(mods, "base"),
          "dropout_contextse": make_dropout(mods, "contextse"), "aishell": make_aishell(mods)}
    torch.save(fx, os.path.join(HERE, "variants.pt"))
    for k, v in fx.items():
        print(k, {kk: (tuple(vv.shape) if torch.is_tensor(vv) else type(vv).__name__) for kk, vv in v.items()})
    print("size", os.path.getsize(os.path.join(HERE, "variants.pt")))
