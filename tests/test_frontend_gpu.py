"""Log-mel frontend (csrc/logmel.cu) against the reference's torchaudio pipeline: the committed golden features
(tests/golden/frontend.pt, generated from torchaudio.transforms in the build container) and the CPU oracle."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from golden_common import seeded_wave  # noqa: E402

from conftest import rel_err  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def frontend():
    from lightning_asr_b200 import _lib, frontend

    _lib.require_device()
    return frontend


def _batch(waves):
    S_max = max(w.numel() for w in waves)
    x = torch.zeros(len(waves), S_max)
    for i, w in enumerate(waves):
        x[i, : w.numel()] = w
    return x.cuda(), [w.numel() for w in waves]


def test_logmel_matches_golden_torchaudio_features(frontend):
    """fp32 tolerance of the north star: rel 1e-4 (features are ~N(0,1) per utterance); also the collate contract."""
    fx = torch.load(os.path.join(GOLDEN, "frontend.pt"))
    names = ["w8000", "w12345", "w400"]
    waves = [seeded_wave(8000, 1), seeded_wave(12345, 2), seeded_wave(400, 3)]
    x, ns = _batch(waves)
    out = frontend.logmel_batch(x, ns, out_dtype=torch.float32)
    T_max = out["inputs"].shape[-1]
    for i, name in enumerate(names):
        ref = fx["features"][name]
        T = ref.shape[-1]
        got = out["inputs"][i, :, :, :T].cpu()
        assert rel_err(got, ref) < 1e-4, name
        assert (got - ref).abs().max().item() < 2e-3, name
        assert float(out["inputs"][i, :, :, T:].abs().max()) == 0.0 if T < T_max else True
        # channels-last twin is the same data
        assert torch.equal(out["ntc"][i].t().cpu(), out["inputs"][i, 0].cpu())
    # batched output == the reference's collate of per-utterance features (data_module.py:222-248)
    c = fx["collate"]
    order = [0, 1, 2]
    assert rel_err(out["inputs"][order].cpu(), c["inputs"]) < 1e-4
    assert torch.allclose(out["percents"], c["percents"])


@pytest.mark.parametrize("products,tol", [(6, 1e-4), (3, 1e-3)])
def test_logmel_matches_oracle_ragged_with_dither(frontend, products, tol):
    from oracle import frontend_oracle

    g = torch.Generator().manual_seed(5)
    lens = [16000 * 3, 16000 * 2 + 123, 9000, 16000 * 3 - 1]
    # speech-like dynamic range: a decaying spectrum, not white noise (exercises the cancellation-heavy low bins)
    waves = []
    for n in lens:
        w = torch.randn(n, generator=g)
        w = torch.cumsum(w, 0)
        w = w - w.mean()
        waves.append((0.3 * w / w.abs().max()).float())
    dither = [torch.randn(n, generator=g) for n in lens]
    x, ns = _batch(waves)
    d, _ = _batch(dither)
    out = frontend.logmel_batch(x, ns, dither=d, products=products)
    for i, n in enumerate(lens):
        ref = frontend_oracle.logmel(waves[i], dither=dither[i])
        T = ref.shape[-1]
        assert T == int(out["frames"][i])
        assert rel_err(out["inputs"][i, :, :, :T].cpu(), ref) < tol, (i, products)


def test_audio_parser_contract(frontend, tmp_path):
    """AudioParser.parse_audio(path) -> [1, 64, T] (data_module.py:150-174); missing file raises like :151-152."""
    import wave

    import numpy as np

    w = seeded_wave(8000, 1)
    pcm = (w.clamp(-1, 1) * 32767).round().to(torch.int16).numpy()
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(16000)
        wf.writeframes(pcm.astype("<i2").tobytes())
    parser = frontend.AudioParser(dither=False)
    feat = parser.parse_audio(path)
    from oracle import frontend_oracle

    ref = frontend_oracle.logmel(torch.from_numpy(pcm.astype(np.float32) / 32768.0))
    assert feat.shape == ref.shape
    assert rel_err(feat.cpu(), ref) < 1e-4
    with pytest.raises(FileExistsError):
        parser.parse_audio(str(tmp_path / "missing.wav"))


def test_device_augmentation_matches_reference_fixture(frontend):
    """logmel_batch(starts=, bands=) -- crop after pre-emphasis + SpecAugment in the dB domain + statistics correction,
    all on the device -- against the reference's own parse_audio(mask=True) output (tests/golden/augment.pt)."""
    import os
    import random

    fx = torch.load(os.path.join(GOLDEN, "augment.pt"), weights_only=False)
    names = sorted(fx)
    waves = [seeded_wave(fx[k]["samples"], fx[k]["seed"]) for k in names]
    draws = [frontend.draw_augment(fx[k]["samples"], random.Random(fx[k]["rng_seed"])) for k in names]
    S = max(len(w) for w in waves)
    x = torch.zeros(len(waves), S)
    for i, w in enumerate(waves):
        x[i, :len(w)] = w
    out = frontend.logmel_batch(x.cuda(), [d[1] for d in draws], starts=[d[0] for d in draws],
                                bands=[list(d[2]) for d in draws])
    for i, k in enumerate(names):
        ref = fx[k]["features"]
        T = ref.shape[-1]
        assert T == int(out["frames"][i])
        got = out["inputs"][i, :, :, :T].cpu()
        assert rel_err(got, ref) < 1e-4, k
        f0, fw, t0, tw = draws[i][2]
        # masked cells hold exactly the normalised zero: -mean/std, one constant
        if fw > 0:
            band = got[0, f0:f0 + fw, :]
            assert float(band.max() - band.min()) == 0.0


def test_augment_draw_on_device_matches_reference_arithmetic(frontend):
    """lasr_augment_draw fed the six random() values of a seeded random.Random gives exactly the (start, kept, bands) of
    the host restatement frontend.draw_augment (pinned to the reference's parse_audio(mask=True) by the fixture test
    above) -- i.e. the device kernel restates data_module.py:138-148,97-122 in the reference's draw order."""
    import random

    from lightning_asr_b200 import _lib
    sizes = [20000, 48000, 256000, 16000 * 7 + 13, 4000, 300000, 65537]
    seeds = [3, 904, 905, 17, 99, 1234, 5]
    want, unif = [], []
    for S, sd in zip(sizes, seeds):
        want.append(frontend.draw_augment(S, random.Random(sd)))
        r = random.Random(sd)
        unif.append([r.random() for _ in range(6)])
    N = len(sizes)
    T_max = frontend.num_frames(max(sizes))
    ns = torch.tensor(sizes, device="cuda", dtype=torch.int32)
    u = torch.tensor(unif, device="cuda", dtype=torch.float64)
    starts = torch.zeros(N, device="cuda", dtype=torch.int32)
    kept = torch.zeros_like(starts)
    bands = torch.zeros(N, 4, device="cuda", dtype=torch.int32)
    perc = torch.zeros(N, device="cuda")
    _lib.call("lasr_augment_draw", ns, u, 0, None, starts, kept, bands, perc, N, T_max, 1, 1)
    for i, (st, kp, bd) in enumerate(want):
        assert (int(starts[i]), int(kept[i]), tuple(bands[i].tolist())) == (st, kp, tuple(bd)), i
        assert float(perc[i]) == pytest.approx(frontend.num_frames(kp) / T_max, rel=1e-7)
    # augmentation switched off: the whole utterance, empty bands
    _lib.call("lasr_augment_draw", ns, u, 0, None, starts, kept, bands, perc, N, T_max, 0, 0)
    assert torch.equal(kept, ns) and int(starts.abs().sum()) == 0 and int(bands.abs().sum()) == 0
    # Philox draws: inside the reference's ranges, different per utterance, per seed and per device step counter
    ns2 = torch.full((64,), 160000, device="cuda", dtype=torch.int32)
    outs = []
    for seed, ctr in ((7, None), (8, None), (7, torch.tensor(3, device="cuda", dtype=torch.int64))):
        st2, kp2 = torch.zeros(64, device="cuda", dtype=torch.int32), torch.zeros(64, device="cuda", dtype=torch.int32)
        bd2 = torch.zeros(64, 4, device="cuda", dtype=torch.int32)
        _lib.call("lasr_augment_draw", ns2, None, seed, ctr, st2, kp2, bd2, None, 64, 1001, 1, 1)
        assert int(kp2.min()) > 0.96 * 160000 - 3200 and int((st2 + kp2).max()) <= 160000
        assert int(st2.max()) <= 3200 and int(bd2[:, 1].max()) < 27 and int((bd2[:, 0] + bd2[:, 1]).max()) <= 64
        assert int((bd2[:, 2] + bd2[:, 3]).max()) <= 1001 and len(set(kp2.tolist())) > 32
        outs.append(kp2.clone())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("wire", ["fp32", "int16"])
def test_device_frontend_matches_reference_fixture(frontend, wire):
    """DeviceFrontend (static buffers, draws on the device, what TrainEngine captures in its step graph) against the
    reference's own parse_audio(mask=True) output; the int16 wire format against the same utterances quantised to
    16-bit PCM first (the reference decodes 16-bit wav files, data_module.py:153)."""
    import os
    import random

    from oracle import frontend_oracle
    fx = torch.load(os.path.join(GOLDEN, "augment.pt"), weights_only=False)
    names = sorted(fx)
    waves = [seeded_wave(fx[k]["samples"], fx[k]["seed"]) for k in names]
    unif = []
    for k in names:
        r = random.Random(fx[k]["rng_seed"])
        unif.append([r.random() for _ in range(6)])
    S = max(len(w) for w in waves)
    x = torch.zeros(len(waves), S)
    for i, w in enumerate(waves):
        x[i, :len(w)] = w
    ns = torch.tensor([len(w) for w in waves], device="cuda", dtype=torch.int32)
    xd = x.cuda() if wire == "fp32" else frontend.pcm16(x).cuda()
    fe = frontend.DeviceFrontend(len(waves), S, "cuda", torch.float32, augment=True)
    feats, perc = fe(xd.contiguous(), ns, torch.tensor(unif, device="cuda", dtype=torch.float64))
    for i, k in enumerate(names):
        start, kept, bands = frontend.draw_augment(fx[k]["samples"], random.Random(fx[k]["rng_seed"]))
        if wire == "fp32":
            ref = fx[k]["features"]
        else:  # the oracle (pinned to the same fixture) on the quantised samples
            ref = frontend_oracle.logmel(frontend.pcm16(waves[i]).float() / 32768.0, crop=(start, kept), bands=bands)
        T = ref.shape[-1]
        assert float(perc[i]) == pytest.approx(T / fe.T_max, rel=1e-7)
        got = feats[i, :T].t().cpu()
        assert rel_err(got, ref[0]) < 1e-4, k
        assert float(feats[i, T:].abs().max()) == 0.0 if T < fe.T_max else True  # collate's zero padding
