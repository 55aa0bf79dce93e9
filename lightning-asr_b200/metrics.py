"""Greedy CTC decode + WER/CER metric with the reference's call contract (utils/asr_metrics.py:62-228).

The argmax and the collapse rule (keep p iff (p != previous or previous == blank) and p != blank, :163-167) run in
the greedy-decode kernels; only the compact token ids ([N] counts + [N, T] int32) cross to the host, where the
strings and the Levenshtein distance (editdistance.eval in the reference, :54,:220 -- O(S^2) on tiny strings, not a
GPU workload) are produced.
"""
import torch

from . import ops


def _levenshtein(a, b):
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def word_error_rate(hypotheses, references, use_cer=False):
    """utils/asr_metrics.py:26-59."""
    scores = 0
    words = 0
    if len(hypotheses) != len(references):
        raise ValueError(
            "In word error rate calculation, hypotheses and reference"
            " lists must have the same number of elements. But I got:"
            "{0} and {1} correspondingly".format(len(hypotheses), len(references))
        )
    for h, r in zip(hypotheses, references):
        h_list, r_list = (list(h), list(r)) if use_cer else (h.split(), r.split())
        words += len(r_list)
        scores += _levenshtein(h_list, r_list)
    return 1.0 * scores / words if words != 0 else float("inf")


class WER(torch.nn.Module):
    """WER(vocabulary, batch_dim_index=0, use_cer=False, ctc_decode=True, log_prediction=True) -- same state
    semantics as the reference: `update` OVERWRITES scores / words (:222-223), `compute` = scores / words."""

    def __init__(self, vocabulary, batch_dim_index=0, use_cer=False, ctc_decode=True, log_prediction=True,
                 dist_sync_on_step=False):
        super().__init__()
        if batch_dim_index != 0:
            raise NotImplementedError("batch_dim_index must be 0 (the only value the reference uses)")
        self.batch_dim_index = batch_dim_index
        self.blank_id = len(vocabulary)
        self.labels_map = dict([(i, vocabulary[i]) for i in range(len(vocabulary))])
        self.use_cer = use_cer
        self.ctc_decode = ctc_decode
        self.log_prediction = log_prediction
        self.dist_sync_on_step = dist_sync_on_step
        self.scores = torch.tensor(0.0)
        self.words = torch.tensor(0.0)

    # -- device side --------------------------------------------------------------------------------------------
    def decode_tokens(self, predictions, predictions_len=None):
        """predictions [N, T] integer argmax ids (CUDA) -> list of token-id lists (collapse kernel)."""
        if not predictions.is_cuda:
            raise RuntimeError("lightning_asr_b200.WER decodes on the GPU; predictions must be a CUDA tensor")
        pred = predictions.long().contiguous()
        lens = None
        if predictions_len is not None:
            lens = torch.as_tensor(predictions_len).to(pred.device).int().contiguous()
        tokens, counts = ops.ctc_collapse(pred, lens, self.blank_id)
        tokens, counts = tokens.cpu(), counts.cpu().tolist()
        return [tokens[i, : counts[i]].tolist() for i in range(len(counts))]

    def decode_scores(self, scores, lengths=None):
        """scores [N, T, V'] (log-probs or logits, V' = blank + 1 valid classes) -> token-id lists; argmax fused."""
        _, tokens, counts = ops.greedy_decode(scores, lengths, self.blank_id + 1, self.blank_id)
        tokens, counts = tokens.cpu(), counts.cpu().tolist()
        return [tokens[i, : counts[i]].tolist() for i in range(len(counts))]

    # -- reference API ------------------------------------------------------------------------------------------
    def ctc_decoder_predictions_tensor(self, predictions, predictions_len=None):
        return ["".join(self.labels_map[c] for c in toks) for toks in self.decode_tokens(predictions, predictions_len)]

    def decode_reference(self, targets, target_lengths):
        tg = targets.long().cpu()
        tl = target_lengths.long().cpu().tolist()
        return ["".join(self.labels_map[c] for c in tg[i][: tl[i]].tolist()) for i in range(tg.shape[0])]

    def update(self, predictions, targets, target_lengths, t_lengths=None):
        with torch.no_grad():
            references = self.decode_reference(targets, target_lengths)
            if not self.ctc_decode:
                raise NotImplementedError("Implement me if you need non-CTC decode on predictions")
            hypotheses = self.ctc_decoder_predictions_tensor(predictions, t_lengths)
        words = 0.0
        scores = 0.0
        for h, r in zip(hypotheses, references):
            h_list, r_list = (list(h), list(r)) if self.use_cer else (h.split(), r.split())
            words += len(r_list)
            scores += _levenshtein(h_list, r_list)
        self.scores = torch.tensor(scores)
        self.words = torch.tensor(words)

    def compute(self):
        return self.scores.detach().float() / self.words.detach().float()

    def forward(self, predictions, targets, target_lengths, t_lengths=None):
        """torchmetrics' compute_on_step=True behaviour: update, then return this batch's value."""
        self.update(predictions, targets, target_lengths, t_lengths)
        return self.compute()
