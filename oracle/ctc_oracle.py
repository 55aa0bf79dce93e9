"""CPU oracle (test infrastructure): CTC loss, greedy CTC decode, reference strings, WER/CER.

CTC: the reference calls torch.nn.CTCLoss(blank=len(labels), reduction='none') (train.py:196) on
`out.transpose(0, 1)` with t_lengths = torch.mul(T', percentage).int() (train.py:76-78).  The arithmetic lives in
the third-party dependency torch (pinned torch==1.8.1 in requirements.txt:1; installed 2.11.0), ATen LossCTC.cpp,
which implements Graves et al. 2006 in log space.  `ctc_nll_and_grad` restates that published algorithm in numpy
float64 (alpha/beta recursions, eq. 16 gradient as ATen defines it: softmax - occupancy, zero past the input length);
tests pin it against torch.nn.functional.ctc_loss run in the build container (golden fixtures) and at run time.

Decode / WER: restated from utils/asr_metrics.py:153-171 (greedy collapse), :173-185 (references),
:46-59 and :211-220 (word / char error with Levenshtein distance; editdistance==0.5.3 is absent, so the classic
dynamic programme is restated here).
"""
import numpy as np


def _logsumexp(*xs):
    m = max(xs)
    if m == -np.inf:
        return -np.inf
    return m + np.log(sum(np.exp(x - m) for x in xs))


def ctc_nll_and_grad(log_probs, targets, input_length, target_length, blank):
    """log_probs [T, V] float64 (one utterance), targets [S] ints.  Returns (nll, grad [T, V]) where grad is the
    gradient ATen reports for log_probs: exp(lp) - exp(log_occupancy + nll - lp) for t < input_length, else 0."""
    lp = np.asarray(log_probs, dtype=np.float64)
    T, V = lp.shape
    Tn, Sn = int(input_length), int(target_length)
    ext = [blank]
    for s in range(Sn):
        ext += [int(targets[s]), blank]
    L = len(ext)
    alpha = np.full((T, L), -np.inf)
    beta = np.full((T, L), -np.inf)
    grad = np.zeros_like(lp)
    if Tn == 0:
        return (0.0 if Sn == 0 else np.inf), grad
    alpha[0, 0] = lp[0, blank]
    if L > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, Tn):
        for s in range(L):
            a = alpha[t - 1, s]
            b = alpha[t - 1, s - 1] if s >= 1 else -np.inf
            c = alpha[t - 1, s - 2] if (s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]) else -np.inf
            alpha[t, s] = _logsumexp(a, b, c) + lp[t, ext[s]]
    tail = [alpha[Tn - 1, L - 1]] + ([alpha[Tn - 1, L - 2]] if L > 1 else [])
    ll = _logsumexp(*tail)
    nll = -ll
    beta[Tn - 1, L - 1] = lp[Tn - 1, blank]
    if L > 1:
        beta[Tn - 1, L - 2] = lp[Tn - 1, ext[L - 2]]
    for t in range(Tn - 2, -1, -1):
        for s in range(L):
            a = beta[t + 1, s]
            b = beta[t + 1, s + 1] if s + 1 < L else -np.inf
            c = beta[t + 1, s + 2] if (s + 2 < L and ext[s] != blank and ext[s] != ext[s + 2]) else -np.inf
            beta[t, s] = _logsumexp(a, b, c) + lp[t, ext[s]]
    with np.errstate(invalid="ignore", over="ignore"):
        for t in range(Tn):
            occ = np.full(V, -np.inf)
            for s in range(L):
                occ[ext[s]] = _logsumexp(occ[ext[s]], alpha[t, s] + beta[t, s])
            grad[t] = np.exp(lp[t]) - np.exp(occ + nll - lp[t])
    return nll, grad


def greedy_collapse(prediction, blank_id):
    """utils/asr_metrics.py:161-167: keep p iff (p != previous or previous == blank) and p != blank."""
    decoded = []
    previous = blank_id
    for p in prediction:
        if (p != previous or previous == blank_id) and p != blank_id:
            decoded.append(p)
        previous = p
    return decoded


def ctc_decoder_predictions(predictions, labels, predictions_len=None):
    """utils/asr_metrics.py:153-171.  predictions: [N][T] ints.  Returns (token id lists, strings)."""
    blank_id = len(labels)
    toks, hyps = [], []
    for i, pred in enumerate(predictions):
        pred = [int(p) for p in pred]
        if predictions_len is not None:
            pred = pred[: int(predictions_len[i])]
        d = greedy_collapse(pred, blank_id)
        toks.append(d)
        hyps.append("".join(labels[c] for c in d))
    return toks, hyps


def decode_reference(targets, target_lengths, labels):
    """utils/asr_metrics.py:173-185."""
    return ["".join(labels[int(c)] for c in tgt[: int(n)]) for tgt, n in zip(targets, target_lengths)]


def levenshtein(a, b):
    """Plain edit distance (what editdistance.eval computes, utils/asr_metrics.py:54,220)."""
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
    """utils/asr_metrics.py:26-59 / the per-batch score of WER.update+compute :211-228."""
    if len(hypotheses) != len(references):
        raise ValueError("hypotheses and references must have the same number of elements")
    scores = words = 0
    for h, r in zip(hypotheses, references):
        h_list, r_list = (list(h), list(r)) if use_cer else (h.split(), r.split())
        words += len(r_list)
        scores += levenshtein(h_list, r_list)
    return 1.0 * scores / words if words != 0 else float("inf")
