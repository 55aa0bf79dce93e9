"""Pin the CPU oracle (oracle/) to the golden fixtures under tests/golden/, which are outputs of the REFERENCE's own
code run in the build container (tests/golden/make_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)

from golden_common import LABELS28, ctc_inputs, golden_weights, model_inputs, seeded_wave  # noqa: E402

from oracle import ctc_oracle, frontend_oracle, quartznet_oracle, train_oracle  # noqa: E402


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.mark.parametrize("variant", ["base", "context", "contextse"])
def test_model_oracle_matches_reference(variant):
    fx = _load(f"model_{variant}.pt")
    x, percents, targets, tgt_len = model_inputs()
    for mode in ("train", "eval"):
        sd = golden_weights(fx["schema"])
        out = quartznet_oracle.model(x, percents, sd, mask=True, training=(mode == "train"))
        assert rel_err(out, fx[mode + "_out"]) < 2e-6
        t_len = torch.mul(out.size(1), percents).int()
        nll = torch.nn.functional.ctc_loss(out.transpose(0, 1), targets, t_len, tgt_len, blank=28, reduction="none")
        assert rel_err(nll, fx[mode + "_nll"]) < 2e-6
    # gradients + running statistics of one training step
    sd = golden_weights(fx["schema"])
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    loss, out, _ = train_oracle.training_step(sd, (x, targets, percents, tgt_len), LABELS28, mask=True, training=True,
                                              update_buffers=True)
    loss.backward()
    for k, ref in fx["grad_norm"].items():
        got = float(sd[k].grad.double().norm())
        assert abs(got - ref) <= 5e-3 * max(ref, 1e-12), (k, got, ref)  # whole-network fp32 gradient noise floor
    for k, ref in fx["grad"].items():
        assert rel_err(sd[k].grad, ref) < 1e-2, k
    for k, ref in fx["running"].items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(ref)
        else:
            assert rel_err(sd[k].float(), ref.float()) < 1e-5, k


def test_frontend_oracle_matches_reference():
    fx = _load("frontend.pt")
    waves = {"w8000": seeded_wave(8000, 1), "w12345": seeded_wave(12345, 2), "w400": seeded_wave(400, 3)}
    for name, w in waves.items():
        got = frontend_oracle.logmel(w)
        ref = fx["features"][name]
        assert got.shape == ref.shape
        assert got.shape[2] == frontend_oracle.num_frames(w.numel())
        assert (got - ref).abs().max().item() < 2e-5, name
    c = fx["collate"]
    batch = [(fx["features"]["w8000"], [3, 4, 5, 6], "a"), (fx["features"]["w12345"], [7, 8], "b"),
             (fx["features"]["w400"], [9, 10, 11], "c")]
    inputs, targets, percents, sizes, paths = frontend_oracle.collate(batch)
    assert torch.equal(inputs, c["inputs"]) and torch.equal(targets, c["targets"])
    assert torch.equal(percents, c["percents"]) and torch.equal(sizes, c["sizes"]) and paths == c["paths"]


def test_decode_oracle_matches_reference():
    fx = _load("decode.pt")
    pred, lens = fx["pred"].tolist(), fx["lens"].tolist()
    _, hyp_len = ctc_oracle.ctc_decoder_predictions(pred, LABELS28, lens)
    _, hyp_all = ctc_oracle.ctc_decoder_predictions(pred, LABELS28, None)
    assert hyp_len == fx["hyp_len"] and hyp_all == fx["hyp_all"]
    refs = ctc_oracle.decode_reference(fx["targets"].tolist(), fx["target_lens"].tolist(), LABELS28)
    assert refs == fx["refs"]
    assert ctc_oracle.word_error_rate(hyp_len, refs) == pytest.approx(fx["wer"])
    assert ctc_oracle.word_error_rate(hyp_len, refs, use_cer=True) == pytest.approx(fx["cer"])
    assert ctc_oracle.word_error_rate(hyp_len, refs) == pytest.approx(fx["wer_update"], rel=1e-6)
    with pytest.raises(ValueError):
        ctc_oracle.word_error_rate(["a"], ["a", "b"])  # utils/asr_metrics.py:40-45


@pytest.mark.parametrize("name", ["v29", "v4334"])
def test_ctc_oracle_matches_torch(name):
    fx = _load("ctc.pt")[name]
    lp, targets, in_len, tg_len, blank = ctc_inputs(name)
    assert float(lp.sum()) == pytest.approx(fx["lp_checksum"], rel=1e-12)  # same seeded inputs as the fixture
    for n in range(lp.shape[0]):
        nll, grad = ctc_oracle.ctc_nll_and_grad(lp[n].numpy(), targets[n].numpy(), int(in_len[n]), int(tg_len[n]),
                                                blank)
        ref = float(fx["nll"][n])
        if np.isinf(ref):
            assert np.isinf(nll) and nll > 0  # infeasible alignment -> +inf (zero_infinity=False)
            continue
        assert nll == pytest.approx(ref, rel=1e-10)
        g = fx["grad"]
        if isinstance(g, dict):
            cols = g["cols"].numpy()
            assert np.abs(grad[:, cols] - g["values"][n].numpy()).max() < 1e-9
        else:
            assert np.abs(grad - g[n].numpy()).max() < 1e-9
        assert np.abs(grad[int(in_len[n]):]).max(initial=0.0) == 0.0


def test_mask_lengths_truncation():
    # models/QuartNet.py:311: fp32 product then .int() truncation
    p = torch.tensor([1.0, 0.999, 0.5, 0.0013])
    assert quartznet_oracle.mask_lengths(801, p).tolist() == [801, 800, 400, 1]
    x = torch.ones(4, 2, 801)
    y = quartznet_oracle.mask_cnn(x, p)
    assert y.sum(dim=(1, 2)).tolist() == [1602.0, 1600.0, 800.0, 2.0]


def test_optimizer_oracle_matches_reference():
    """oracle/optim_oracle.py vs tests/golden/optim.pt (the reference's own Novograd + CosineAnnealingWarmupRestarts)."""
    from golden_common import optim_case
    from oracle import optim_oracle

    fx = _load("optim.pt")
    for name, (hyper, sched, steps) in optim_case.CASES.items():
        params = optim_case.params()
        state = [{} for _ in params]
        sch = optim_oracle.CosineWarmupOracle(**sched) if sched else None
        h = dict(hyper)
        lr0 = h.pop("lr")
        snaps = dict(fx[name]["snaps"])
        for k in range(steps):
            lr = sch.lr if sch else lr0
            assert abs(lr - fx[name]["lrs"][k]) <= 1e-15
            optim_oracle.novograd_step(params, optim_case.grads(k), state, lr, **h)
            if sch:
                sch.step()
            if k in snaps:
                for a, b in zip(params, snaps[k]):
                    assert torch.allclose(a, b, rtol=1e-6, atol=1e-8), (name, k)
        for st, v, m in zip(state, fx[name]["exp_avg_sq"], fx[name]["exp_avg"]):
            assert abs(float(st["exp_avg_sq"]) - v) <= 1e-6 * abs(v)
            assert torch.allclose(st["exp_avg"], m, rtol=1e-6, atol=1e-8)


def test_augment_oracle_matches_reference():
    """oracle.frontend_oracle.draw_augment + logmel(crop=, bands=) vs tests/golden/augment.pt = the reference's own
    AudioParser.parse_audio(mask=True) with its random generators seeded (tests/golden/make_golden_augment.py)."""
    import random

    fx = _load("augment.pt")
    for name, d in fx.items():
        w = seeded_wave(d["samples"], d["seed"])
        start, kept, bands = frontend_oracle.draw_augment(d["samples"], random.Random(d["rng_seed"]))
        got = frontend_oracle.logmel(w, crop=(start, kept), bands=bands)
        assert got.shape == d["features"].shape, name
        assert rel_err(got, d["features"]) < 1e-5, name
        # the product's host-side draw (no oracle import there) follows the same arithmetic
        from lightning_asr_b200 import frontend
        assert frontend.draw_augment(d["samples"], random.Random(d["rng_seed"])) == (start, kept, bands)


# ---------------------------------------------------------------------------------------------------------------
# variants.pt (tests/golden/make_golden_variants.py): mask=False, QuartNetBlock(repeat=2), drop_rate > 0, AISHELL vocab
# ---------------------------------------------------------------------------------------------------------------
from golden_common import aishell_labels, block_inputs, drop_factor  # noqa: E402

DROP_NAMES = ["first_cnn", "block1", "block12", "block13", "block2", "block22", "block23", "block3", "block32",
              "block33", "block4", "block42", "block43", "block5"]


def drop_masks_for(fx, with_block6):
    """The keep factors make_golden_variants.py fed the reference's nn.Dropout calls, keyed by module name (the call
    order of the reference's forward: first_cnn, the blocks in order, last_cnn2)."""
    names = DROP_NAMES + (["block6"] if with_block6 else []) + ["last_cnn2"]
    assert len(names) == len(fx["calls"])
    return {n: drop_factor(i, fx["calls"][i], fx["p"]) for i, n in enumerate(names)}


def _train_grads(sd, x, percents, targets, tgt_len, labels, mask, drop_masks=None):
    """fp64 oracle step (the variants fixture is the reference's own fp64 run, stored as fp32)."""
    for k in list(sd):
        if sd[k].is_floating_point():
            sd[k] = sd[k].double()
            if "running" not in k:
                sd[k].requires_grad_(True)
    if drop_masks is not None:
        drop_masks = {k: v.double() for k, v in drop_masks.items()}
    out = quartznet_oracle.model(x.double(), percents, sd, mask=mask, training=True, drop_masks=drop_masks)
    t_len = torch.mul(out.size(1), percents).int()
    nll = torch.nn.functional.ctc_loss(out.transpose(0, 1), targets, t_len, tgt_len, blank=len(labels), reduction="none")
    nll.mean().backward()
    return out.detach(), nll.detach()


def _check_grads(sd, fx, norm_tol=1e-5, tol=1e-5):
    for k, ref in fx["grad_norm"].items():
        got = float(sd[k].grad.double().norm())
        assert abs(got - ref) <= norm_tol * max(ref, 1e-12), (k, got, ref)
    for k, ref in fx["grad"].items():
        assert rel_err(sd[k].grad, ref) < tol, k


def test_oracle_mask_false_matches_reference():
    fx = _load("variants.pt")["mask_false"]
    sd = golden_weights(fx["schema"])
    out, nll = _train_grads(sd, *model_inputs(), LABELS28, mask=False)
    assert rel_err(out, fx["train_out"]) < 2e-6
    assert rel_err(nll, fx["train_nll"]) < 2e-6
    _check_grads(sd, fx)


@pytest.mark.parametrize("mask", [True, False])
def test_block_repeat_quirk_matches_reference(mask):
    """QuartNetBlock(repeat=2): `mask` lands in the inner SeprationConv's `last` slot (models/QuartNet.py:60)."""
    fx = _load("variants.pt")["repeat2_mask" if mask else "repeat2_nomask"]
    assert fx["inner_last"] == mask and fx["inner_mask"] is True
    sd = {"b." + k: v.requires_grad_(v.is_floating_point() and "running" not in k)
          for k, v in golden_weights(fx["schema"]).items()}
    x, percents, dout = block_inputs(64, 96)
    x = x.clone().requires_grad_(True)
    out = quartznet_oracle.block(x, percents, sd, "b", mask=mask, training=True, update_buffers=False)
    out.backward(dout)
    assert rel_err(out, fx["out"]) < 2e-6
    assert rel_err(x.grad, fx["dx"]) < 1e-5
    for k, ref in fx["grad"].items():
        assert rel_err(sd["b." + k].grad, ref) < 1e-5, k


@pytest.mark.parametrize("variant", ["base", "contextse"])
def test_oracle_dropout_placement_matches_reference(variant):
    fx = _load("variants.pt")["dropout_" + variant]
    sd = golden_weights(fx["schema"])
    dm = drop_masks_for(fx, with_block6=(variant != "base"))
    out, nll = _train_grads(sd, *model_inputs(), LABELS28, mask=True, drop_masks=dm)
    assert rel_err(out, fx["train_out"]) < 2e-6
    assert rel_err(nll, fx["train_nll"]) < 2e-6
    _check_grads(sd, fx)


def test_oracle_aishell_vocab_matches_reference():
    fx = _load("variants.pt")["aishell"]
    labels = aishell_labels()
    sd = golden_weights(fx["schema"])
    x, percents, targets, tgt_len = model_inputs(n_labels=len(labels))
    out, nll = _train_grads(sd, x, percents, targets, tgt_len, labels, mask=True)
    assert out.shape[-1] == 4334
    assert rel_err(out[:, :, fx["cols"]], fx["train_out_cols"]) < 2e-6
    assert torch.equal(out.argmax(dim=-1), fx["train_out_argmax"])
    assert rel_err(nll, fx["train_nll"]) < 2e-6
    _check_grads(sd, fx)
