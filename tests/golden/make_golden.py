"""Generate the golden fixtures that pin the CPU oracle to the REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The reference ships no tests / golden vectors of its own, so every fixture here is an output of
the reference's own code executed under the installed torch / torchaudio on seeded inputs:

  model_<variant>.pt   models/QuartNet.py | QuartNetContext.py | QuartNetContextSE.py :: MyModel2, imported as-is,
                       weights = golden_weights() (deterministic per key, so the 20 MB state dict is not stored),
                       train-mode and eval-mode log-probs, torch.nn.CTCLoss (train.py:196) losses, gradient norms,
                       BatchNorm running statistics after the step, and the state_dict schema (keys / shapes).
  frontend.pt          data_module.py :: AudioParser.parse_audio, imported with the absent control-plane packages
                       (pytorch_lightning, hydra, omegaconf) stubbed, torchaudio.load patched to return the seeded
                       waveform and the dither (torch.randn_like, :155) patched to zero; plus _collate_fn (:222-248).
  decode.pt            utils/asr_metrics.py :: WER.ctc_decoder_predictions_tensor / decode_reference / word_error_rate
                       imported with torchmetrics / editdistance stubbed (editdistance.eval -> plain Levenshtein).
  ctc.pt               torch.nn.CTCLoss(blank=V, reduction='none') in fp64 on seeded log-probs: nll and gradient
                       (the arithmetic lives in torch, third-party; this pins the numpy restatement).
Fixtures are small (a few hundred KB in total) and committed together with this script.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from golden_common import LABELS28, golden_weights, model_inputs, seeded_wave  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    sys.path.insert(0, REF)
    _stub("pytorch_lightning", LightningDataModule=object, LightningModule=torch.nn.Module)
    _stub("hydra", main=lambda **kw: (lambda f: f))
    _stub("omegaconf", DictConfig=dict, OmegaConf=object, ListConfig=list)

    class Metric(torch.nn.Module):
        def __init__(self, **kw):
            super().__init__()

        def add_state(self, name, default, **kw):
            setattr(self, name, default)

    _stub("torchmetrics", Metric=Metric)

    def lev(a, b):
        prev = list(range(len(b) + 1))
        for i, x in enumerate(a, 1):
            cur = [i]
            for j, y in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
            prev = cur
        return prev[-1]

    _stub("editdistance", eval=lev)
    import torchaudio

    if not hasattr(torchaudio, "set_audio_backend"):
        torchaudio.set_audio_backend = lambda *_: None
    import data_module
    from models import QuartNet, QuartNetContext, QuartNetContextSE
    from utils import asr_metrics

    return data_module, asr_metrics, {"base": QuartNet, "context": QuartNetContext, "contextse": QuartNetContextSE}


def make_model_fixture(variant, mod):
    torch.manual_seed(0)
    model = mod.MyModel2(LABELS28, drop_rate=0.0, mask=True)
    schema = [(k, tuple(v.shape), str(v.dtype)) for k, v in model.state_dict().items()]
    model.load_state_dict(golden_weights(schema), strict=True)
    x, percents, targets, tgt_len = model_inputs()
    fx = {"schema": schema}
    for mode in ("train", "eval"):
        model.load_state_dict(golden_weights(schema), strict=True)
        model.train(mode == "train")
        model.zero_grad()
        out = model(x, percents)  # [N, T', 29]
        t_len = torch.mul(out.size(1), percents).int()  # train.py:76
        nll = torch.nn.CTCLoss(blank=len(LABELS28), reduction="none")(out.transpose(0, 1), targets, t_len, tgt_len)
        fx[mode + "_out"] = out.detach().clone()
        fx[mode + "_nll"] = nll.detach().clone()
        if mode == "train":
            torch.mean(nll).backward()  # train.py:77
            fx["grad_norm"] = {k: float(p.grad.double().norm()) for k, p in model.named_parameters()}
            keep = ["encoder.first_cnn.depthwise_conv.weight", "encoder.block1.reside.1.weight", "decoder.bias",
                    "encoder.last_cnn2.1.bias"]
            fx["grad"] = {k: p.grad.detach().clone() for k, p in model.named_parameters() if k in keep}
            sd = model.state_dict()
            fx["running"] = {k: sd[k].detach().clone() for k in sd
                             if k.endswith("running_mean") or k.endswith("running_var")
                             or k.endswith("num_batches_tracked")}
            fx["running"] = {k: v for k, v in fx["running"].items()
                             if k.startswith("encoder.first_cnn") or k.startswith("encoder.block5")
                             or k.startswith("encoder.last_cnn2")}
    torch.save(fx, os.path.join(HERE, f"model_{variant}.pt"))
    print(variant, "log-probs", tuple(fx["train_out"].shape), "nll", fx["train_nll"].tolist())


def make_frontend_fixture(data_module):
    import torchaudio

    fx = {}
    waves = {"w8000": seeded_wave(8000, 1), "w12345": seeded_wave(12345, 2), "w400": seeded_wave(400, 3)}
    parser = data_module.AudioParser()
    real_load, real_randn_like = torchaudio.load, torch.randn_like
    try:
        torch.randn_like = lambda t, *a, **k: torch.zeros_like(t)  # dither off (:155 is unseeded randomness)
        feats = {}
        for name, w in waves.items():
            torchaudio.load = lambda *_a, _w=w, **_k: (_w.clone().reshape(1, -1), 16000)
            feats[name] = parser.parse_audio(object(), mask=False).clone()  # not a str -> no os.path check (:151)
    finally:
        torchaudio.load, torch.randn_like = real_load, real_randn_like
    fx["features"] = feats
    # _collate_fn (:222-248) on those three utterances
    dm = data_module.LibriDataModule.__new__(data_module.LibriDataModule)
    batch = [(feats["w8000"], [3, 4, 5, 6], "a"), (feats["w12345"], [7, 8], "b"), (feats["w400"], [9, 10, 11], "c")]
    inputs, targets, percents, sizes, paths = dm._collate_fn(batch)
    fx["collate"] = {"inputs": inputs, "targets": targets, "percents": percents, "sizes": sizes, "paths": paths}
    torch.save(fx, os.path.join(HERE, "frontend.pt"))
    print("frontend", {k: tuple(v.shape) for k, v in feats.items()})


def make_decode_fixture(asr_metrics):
    g = torch.Generator().manual_seed(11)
    V = len(LABELS28)
    pred = torch.randint(0, V + 1, (6, 80), generator=g)
    pred[:, ::3] = V  # plenty of blanks
    pred[0, 10:20] = 5  # repeats
    pred[1, :] = V  # all blank
    lens = torch.tensor([80, 64, 0, 1, 33, 79])
    wer = asr_metrics.WER(vocabulary=LABELS28)
    hyp_len = wer.ctc_decoder_predictions_tensor(pred, lens)
    hyp_all = wer.ctc_decoder_predictions_tensor(pred)
    targets = torch.randint(0, V, (6, 20), generator=g)
    tl = torch.tensor([20, 5, 1, 0, 13, 7])
    refs = wer.decode_reference(targets, tl)
    fx = {"pred": pred, "lens": lens, "hyp_len": hyp_len, "hyp_all": hyp_all, "targets": targets, "target_lens": tl,
          "refs": refs, "wer": asr_metrics.word_error_rate(hyp_len, refs), "cer": asr_metrics.word_error_rate(
              hyp_len, refs, use_cer=True)}
    wer.update(pred, targets, tl, lens)
    fx["wer_update"] = float(wer.compute())
    torch.save(fx, os.path.join(HERE, "decode.pt"))
    print("decode", hyp_len[:2], fx["wer"], fx["cer"], fx["wer_update"])


def make_ctc_fixture():
    from golden_common import CTC_CASES, ctc_inputs

    cases = {}
    for name in CTC_CASES:
        lp, targets, in_len, tg_len, blank = ctc_inputs(name)
        lp = lp.clone().requires_grad_(True)
        nll = torch.nn.CTCLoss(blank=blank, reduction="none")(lp.transpose(0, 1), targets, in_len, tg_len)
        finite = torch.isfinite(nll)
        nll[finite].sum().backward()
        grad = lp.grad.clone()
        if grad.shape[-1] > 1000:  # keep the fixture small: a column subset of the gradient
            cols = torch.unique(torch.cat([targets.reshape(-1), torch.tensor([blank, 0, 1, 17])]))
            grad = {"cols": cols, "values": grad[:, :, cols].clone(), "abs_sum": float(grad.abs().sum())}
        cases[name] = {"lp_checksum": float(lp.detach().sum()), "nll": nll.detach().clone(), "grad": grad}
    torch.save(cases, os.path.join(HERE, "ctc.pt"))
    print("ctc", {k: v["nll"].tolist() for k, v in cases.items()})


if __name__ == "__main__":
    torch.set_num_threads(8)
    data_module, asr_metrics, model_mods = import_reference()
    for variant, mod in model_mods.items():
        make_model_fixture(variant, mod)
    make_frontend_fixture(data_module)
    make_decode_fixture(asr_metrics)
    make_ctc_fixture()
