"""CPU oracle (test infrastructure): the log-mel frontend and the collate contract.

Restated from data_module.py (which cannot be imported: needs pytorch_lightning / hydra / torchaudio.load):
  AudioParser.__init__ :59-73   MelSpectrogram(sr=16000, n_fft=512, pad=32, win_length=320, hop_length=160,
                                n_mels=64) + AmplitudeToDB(stype="power")
  AudioParser.parse_audio :150-174 from the waveform tensor onward (dither :155 is random -> injected or off,
                                crop :158-159 and SpecAugment :163-165 are train-time randomness -> off)
  LibriDataModule._collate_fn :222-248
The STFT / mel arithmetic is the third-party torchaudio (pinned 0.8.1, installed 2.11.0); its defaults that matter
(SURVEY.md a1): periodic Hann window, center=True, reflect padding, power=2, onesided, HTK mel scale, norm=None,
f_min 0, f_max sr/2; dB: 10*log10(clamp(x, 1e-10)), no top_db.  Restated with torch.stft + an explicit HTK
filterbank; pinned against torchaudio.transforms in the build container (tests/golden) -- max abs diff 0.0 there.
"""
import math

import torch

SR = 16000
N_FFT = 512
WIN = 320
HOP = 160
PAD = 32
N_MELS = 64


def hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(n_freqs=N_FFT // 2 + 1, n_mels=N_MELS, f_min=0.0, f_max=SR / 2.0, sr=SR, dtype=torch.float32):
    """torchaudio.functional.melscale_fbanks(..., norm=None, mel_scale='htk'): triangular filters [n_freqs, n_mels]."""
    all_freqs = torch.linspace(0, sr // 2, n_freqs, dtype=torch.float64)
    m_min, m_max = hz_to_mel_htk(f_min), hz_to_mel_htk(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.min(down, up), min=0.0)
    return fb.to(dtype)


def preemphasis(y):
    """data_module.py:157: y[0], y[t] - 0.97*y[t-1]   (y [1, S] or [S])."""
    y2 = y.reshape(1, -1)
    return torch.cat((y2[:, 0].unsqueeze(1), y2[:, 1:] - 0.97 * y2[:, :-1]), dim=1)


def num_frames(num_samples):
    return 1 + (num_samples + 2 * PAD) // HOP


def draw_augment(num_samples, rng):
    """The random draws of parse_audio(mask=True) for one utterance, in the reference's order and arithmetic:
    sub_secquence(weight=0.98) :138-148 -> (location, target_length), the slice is x[:, location:target_length] (sic);
    spec_augment(freq_mask=27, time_mask=0.07) :97-122 -> (w_x, w_y, rect_x, rect_y) on the cropped spectrogram.
    `rng` is anything with .uniform(a, b) (the reference uses np.random for the crop and an unseeded random.Random()
    for the bands).  -> (start, kept_samples, (f0, fw, t0, tw))"""
    length = num_samples
    target_length = int(length * rng.uniform(0.98, 1))
    location = int(rng.uniform(0, length - target_length))
    kept = max(target_length - location, 0)
    T = num_frames(kept)
    freq_mask, time_mask = 27, int(T * 0.07)
    w_x = int(rng.uniform(0, freq_mask))
    w_y = int(rng.uniform(0, time_mask))
    rect_x = int(rng.uniform(0, N_MELS - w_x))
    rect_y = int(rng.uniform(0, T - w_y))
    return location, kept, (rect_x, w_x, rect_y, w_y)


def logmel(wave, dither=None, dtype=torch.float32, crop=None, bands=None):
    """wave [S] (one utterance) -> normalised log-mel [1, 64, T]; data_module.py:155-172.  mask=False by default;
    crop=(start, kept_samples) and bands=(f0, fw, t0, tw) replay the train-time augmentation (:158-159, :163-165) with
    explicit draws."""
    y = wave.reshape(1, -1).to(dtype)
    if dither is not None:
        y = y + 1e-5 * dither.reshape(1, -1).to(dtype)  # :155 (torch.randn_like in the reference)
    y = preemphasis(y)
    if crop is not None:
        y = y[:, crop[0]:crop[0] + crop[1]]  # :158-159
    y = torch.nn.functional.pad(y, (PAD, PAD), "constant")  # MelSpectrogram(pad=32)
    window = torch.hann_window(WIN, periodic=True, dtype=dtype)
    spec = torch.stft(y, N_FFT, hop_length=HOP, win_length=WIN, window=window, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    power = spec.real ** 2 + spec.imag ** 2  # [1, 257, T]
    fb = mel_filterbank(dtype=dtype)  # [257, 64]
    mel = torch.matmul(power.transpose(1, 2), fb).transpose(1, 2)  # [1, 64, T]
    db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))  # AmplitudeToDB(stype="power"), ref 1.0
    if bands is not None:  # :110-121: both bands set to 0 in the dB domain, before the statistics
        f0, fw, t0, tw = bands
        db = db.clone()
        db[0, f0:f0 + fw, :] = 0
        db[0, :, t0:t0 + tw] = 0
    std, mean = torch.std_mean(db)  # :171 global, unbiased
    return (db - mean) / std  # :172


def collate(batch):
    """data_module.py:222-248.  batch: list of (feats [1,64,T_i], token id list, path)."""
    longest = max(batch, key=lambda s: s[0].size(2))[0]
    freq, max_t = longest.size(1), longest.size(2)
    max_s = len(max(batch, key=lambda s: len(s[1]))[1])
    n = len(batch)
    inputs = torch.zeros(n, 1, freq, max_t)
    percents = torch.zeros(n, dtype=torch.float32)
    target_sizes = torch.zeros(n, dtype=torch.int32)
    targets = torch.zeros(n, max_s)
    paths = []
    for i, (feat, txt, path) in enumerate(batch):
        t = feat.size(2)
        inputs[i, 0, :, :t] = feat.squeeze(0)
        percents[i] = t / float(max_t)
        target_sizes[i] = len(txt)
        targets[i, : len(txt)] = torch.tensor(txt, dtype=torch.float32)
        paths.append(path)
    return inputs, targets.long(), percents, target_sizes, paths
