"""CPU oracle (test infrastructure): the training / validation step around the model (train.py:64-116), restated
without pytorch_lightning.  Used as the checker in tests, and timed as the reported CPU baseline (`kind: "port"`)."""
import torch

from . import ctc_oracle, quartznet_oracle


def training_step(sd, batch, labels, mask=True, training=True, update_buffers=False):
    """train.py:64-86.  batch = (inputs [N,1,64,T], targets [N,S] long, percents [N], target_sizes [N] int)
    Returns (loss scalar, log_probs [N,T',V'], t_lengths)."""
    inputs, targets, percents, target_sizes = batch[:4]
    out = quartznet_oracle.model(inputs, percents, sd, mask=mask, training=training, update_buffers=update_buffers)
    t_lengths = torch.mul(out.size(1), percents).int()  # :76
    nll = torch.nn.functional.ctc_loss(out.transpose(0, 1), targets, t_lengths, target_sizes, blank=len(labels),
                                       reduction="none", zero_infinity=False)  # :196, :77-78
    return torch.mean(nll), out, t_lengths


def validation_metrics(out, t_lengths, targets, target_sizes, labels, use_cer=False):
    """train.py:97-106: greedy argmax decode + WER against the reference strings."""
    pred = out.argmax(dim=-1).cpu().tolist()
    toks, hyps = ctc_oracle.ctc_decoder_predictions(pred, labels, t_lengths.cpu().tolist())
    refs = ctc_oracle.decode_reference(targets.cpu().tolist(), target_sizes.cpu().tolist(), labels)
    return toks, hyps, refs, ctc_oracle.word_error_rate(hyps, refs, use_cer)
