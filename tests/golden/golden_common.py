"""Seeded inputs shared by make_golden.py (which runs the reference in the build container) and the tests (which
run the oracle / the CUDA path anywhere): everything is regenerated from fixed seeds so only OUTPUTS are stored."""
import zlib

import torch

LABELS28 = [" ", "'"] + [chr(ord("a") + i) for i in range(26)]  # conf/conf.yaml:12-13


def _gen(key):
    return torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)


def golden_weights(schema):
    """Deterministic, non-trivial state dict for a (key, shape, dtype) schema: conv / linear / LSTM weights
    ~ N(0, 1/fan_in), BatchNorm gamma ~ U(0.5, 1.5), beta / running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5)."""
    sd = {}
    for key, shape, dtype in schema:
        g = _gen(key)
        shape = tuple(shape)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros(shape, dtype=torch.long)
        elif key.endswith("running_var") or (key.endswith(".weight") and len(shape) == 1):
            sd[key] = torch.rand(shape, generator=g) + 0.5
        elif key.endswith("running_mean") or key.endswith(".bias") or "bias_" in key:
            sd[key] = torch.randn(shape, generator=g) * 0.1
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[key] = torch.randn(shape, generator=g) / max(fan_in, 1) ** 0.5
    return sd


def model_inputs(n=3, frames=131, n_labels=28):
    """features [n,1,64,frames] (zero beyond each utterance, like _collate_fn), percents with max == 1 (needed by the
    Context variants), targets / target lengths with S_i = T'_i // 4."""
    g = torch.Generator().manual_seed(77)
    x = torch.randn(n, 1, 64, frames, generator=g)
    lens = torch.tensor([frames, int(frames * 0.8), int(frames * 0.61)][:n])
    x = x * (torch.arange(frames)[None, :] < lens[:, None])[:, None, None, :]
    percents = lens.float() / float(frames)
    Tp = (frames - 1) // 2 + 1
    t_len = torch.mul(Tp, percents).int()
    tgt_len = (t_len // 4).int()
    targets = torch.randint(0, n_labels, (n, int(tgt_len.max())), generator=g)
    targets = targets * (torch.arange(targets.shape[1])[None, :] < tgt_len[:, None])
    return x, percents, targets.long(), tgt_len


def seeded_wave(num_samples, seed):
    """SURVEY.md 8d waveform: 0.05 * randn clamped to [-1, 1], with a slow sine so the spectrum is not flat."""
    g = torch.Generator().manual_seed(1000 + seed)
    t = torch.arange(num_samples, dtype=torch.float32)
    w = 0.05 * torch.randn(num_samples, generator=g) + 0.2 * torch.sin(2 * torch.pi * (220.0 + 30 * seed) * t / 16000)
    return w.clamp_(-1.0, 1.0)


CTC_CASES = {"v29": (29, 40, 4, 9), "v4334": (4334, 12, 2, 4)}  # V, T, N, S


def ctc_inputs(name):
    """-> (log_probs [N,T,V] fp64, targets [N,S], in_len [N], tg_len [N], blank)."""
    V, T, N, S = CTC_CASES[name]
    g = _gen("ctc/" + name)
    lp = torch.log_softmax(torch.randn(N, T, V, generator=g, dtype=torch.float64) * 2, dim=-1)
    targets = torch.randint(0, V - 1, (N, S), generator=g)
    targets[0, :3] = 2  # repeated labels need a blank between them
    in_len = torch.tensor([T, T - 3, T // 2, 3][:N])
    tg_len = torch.tensor([S, S - 2, 1, 5][:N])  # v29's last utterance is infeasible (5 labels in 3 frames) -> inf
    return lp, targets, in_len, tg_len, V - 1


class optim_case:
    """Seeded parameters / gradients for the optimizer fixtures (tests/golden/make_golden_optim.py)."""
    SHAPES = [(64, 32, 1), (32, 1, 33), (64,), (29, 130, 1), (29,), (7,)]
    SNAP_STEPS = (0, 1, 4, 9)
    # name: (Novograd kwargs, scheduler kwargs or None, steps)
    CASES = {
        "train_py": (dict(lr=5e-3, weight_decay=1e-4, betas=(0.8, 0.5)),  # train.py:46, conf/conf.yaml:21-22
                     dict(first_cycle_steps=6, cycle_mult=2, max_lr=5e-3, min_lr=1e-4, warmup_steps=2, gamma=0.5), 24),
        "fixed_lr_avg": (dict(lr=1e-2, weight_decay=0.0, betas=(0.95, 0.98), grad_averaging=True), None, 5),
    }

    @staticmethod
    def params():
        return [torch.randn(s, generator=_gen(f"optim.p{i}")) * 0.1 for i, s in enumerate(optim_case.SHAPES)]

    @staticmethod
    def grads(step):
        return [torch.randn(s, generator=_gen(f"optim.g{i}.{step}")) * (0.01 + 0.003 * i)
                for i, s in enumerate(optim_case.SHAPES)]


def aishell_labels():
    """A 4333-symbol vocabulary the size of data/aishell1-vocab.txt (only the SIZE matters: V' = 4334)."""
    return [chr(0x4E00 + i) for i in range(4333)]


def drop_factor(call_index, shape, p):
    """Dropout factor tensor (0 or 1/(1-p)) of the call_index-th nn.Dropout call of a forward pass, reference layout
    [N, C, T]: seeded per call, so make_golden_variants.py (which patches nn.Dropout.forward with it) and the tests
    (which feed the same masks to the oracle / the CUDA kernels) agree without storing the masks."""
    keep = torch.rand(tuple(shape), generator=_gen(f"drop/{call_index}")) >= p
    return keep.float() / (1.0 - p)


def block_inputs(cin, cout, n=3, frames=75):
    """Seeded input / percents / upstream gradient for a single QuartNetBlock, reference layout [N, C, T]."""
    g = _gen(f"block/{cin}/{cout}")
    x = torch.randn(n, cin, frames, generator=g)
    percents = torch.tensor([1.0, 0.83, 0.52][:n])
    dout = torch.randn(n, cout, frames, generator=g)
    return x, percents, dout
