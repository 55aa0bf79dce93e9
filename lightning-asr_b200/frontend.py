"""Log-mel frontend + collate: the host-side mirror of data_module.py's AudioParser / _collate_fn.

  AudioParser(win_len=0.02, sr=16000).parse_audio(audio, mask=False) -> [1, 64, T]      data_module.py:58-174
  logmel_batch(waves [N, S_max], num_samples [N])  -> (inputs [N,1,64,T_max], percents [N])   the batched GPU entry point
  collate(batch)                                                                        data_module.py:222-248

All arithmetic runs in csrc/logmel.cu (tcgen05 windowed-DFT GEMM, error-compensated); this file only builds the
constant tables (window x DFT basis split in three bf16 terms, sparse HTK mel filterbank) once per device.
The reference runs this per utterance on CPU DataLoader workers; here it is batched on the GPU and can emit the
encoder's channels-last input directly (out_ntc).  Random train-time augmentation (crop :158-159, SpecAugment
:163-165) runs on the device too: the draws restate the reference's arithmetic (draw_augment on the host with any
random.Random, lasr_augment_draw on the device with Philox or caller-supplied uniforms) and are pinned to the
reference's own parse_audio(mask=True) by tests/golden/augment.pt.  DeviceFrontend is the static-buffer, CUDA-graph
capturable pipeline TrainEngine runs inside its step: int16 PCM on the wire, features never leave the device.
"""
import math

import torch

from . import _lib

SR = 16000
N_FFT = 512
WIN = 320
HOP = 160
PAD = 32
N_MELS = 64
N_BINS = N_FFT // 2 + 1

_CONST = {}


def num_frames(num_samples):
    """T = 1 + (S + 2*pad) // hop  (MelSpectrogram(pad=32, hop_length=160, center=True), data_module.py:68-70)."""
    return 1 + (num_samples + 2 * PAD) // HOP


def _hz_to_mel(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(sr=SR):
    """torchaudio.functional.melscale_fbanks(257, 0, sr/2, 64, sr, norm=None, mel_scale='htk') -> [257, 64] fp32."""
    all_freqs = torch.linspace(0, sr // 2, N_BINS, dtype=torch.float64)
    m_pts = torch.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), N_MELS + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0).to(torch.float32)


def _split3(x64):
    """fp32 value -> three bf16 terms h + m + l (24 mantissa bits)."""
    x = x64.to(torch.float32)
    h = x.to(torch.bfloat16)
    r1 = x - h.float()
    m = r1.to(torch.bfloat16)
    r2 = r1 - m.float()
    return torch.stack([h, m, r2.to(torch.bfloat16)])


def constants(device, sr=SR):
    """(basis [3, 512, 320] bf16, mel_idx [257, 2] int32, mel_w [257, 2] fp32) on `device`, built once."""
    key = (str(device), sr)
    if key in _CONST:
        return _CONST[key]
    i = torch.arange(WIN, dtype=torch.float64)
    win = torch.hann_window(WIN, periodic=True, dtype=torch.float64)
    f = torch.arange(N_FFT // 2, dtype=torch.float64)  # bins 0..255
    ang = 2.0 * math.pi * f[:, None] * i[None, :] / N_FFT
    basis = torch.empty(N_FFT, WIN, dtype=torch.float64)
    basis[0::2] = torch.cos(ang) * win
    basis[1::2] = torch.sin(ang) * win
    basis[1] = torch.cos(math.pi * i) * win  # the purely real Nyquist bin takes the slot of im(0) == 0
    fb = mel_filterbank(sr)
    idx = torch.zeros(N_BINS, 2, dtype=torch.int32)
    w = torch.zeros(N_BINS, 2, dtype=torch.float32)
    for b in range(N_BINS):
        nz = torch.nonzero(fb[b]).flatten().tolist()
        if len(nz) > 2:
            raise _lib.LasrError("mel filterbank is not 2-sparse per bin; unsupported sample rate")
        for j, m in enumerate(nz):
            idx[b, j] = m
            w[b, j] = fb[b, m]
    out = (_split3(basis).contiguous().to(device), idx.contiguous().to(device), w.contiguous().to(device))
    _CONST[key] = out
    return out


def draw_augment(num_samples, rng):
    """The random draws of parse_audio(mask=True) (data_module.py:158-165) for one utterance of `num_samples` samples,
    in the reference's order and arithmetic: sub_secquence(weight=0.98) (:138-148; the slice is x[:, loc:L], so
    L - loc samples survive) then spec_augment(freq_mask=27, time_mask=0.07) (:97-122) on the cropped spectrogram.
    `rng`: anything with .uniform(a, b), e.g. a seeded random.Random -- the reference's own generators are unseeded.
    -> (start, kept_samples, (f0, fw, t0, tw)) for logmel_batch(starts=..., num_samples=kept, bands=...)."""
    target_length = int(num_samples * rng.uniform(0.98, 1))
    location = int(rng.uniform(0, num_samples - target_length))
    kept = max(target_length - location, 0)
    T = num_frames(kept)
    w_x = int(rng.uniform(0, 27))
    w_y = int(rng.uniform(0, int(T * 0.07)))
    rect_x = int(rng.uniform(0, N_MELS - w_x))
    rect_y = int(rng.uniform(0, T - w_y))
    return location, kept, (rect_x, w_x, rect_y, w_y)


def logmel_batch(waves, num_samples, dither=None, out_dtype=None, want_nct=True, products=6, sr=SR, starts=None,
                 bands=None):
    """waves [N, S_max] fp32 CUDA (zero padded), num_samples [N] int -> dict with
         'inputs'  [N, 1, 64, T_max] fp32 (reference layout, if want_nct)
         'ntc'     [N, T_max, 64] out_dtype (channels-last encoder input, if out_dtype is not None)
         'percents'[N] fp32 = T_n / T_max  (data_module.py:244), 'frames' [N] int32
       Train-time augmentation on the device (draw_augment): starts [N] int = first kept sample of each utterance's
       pre-emphasised waveform (num_samples then counts the KEPT samples), bands [N, 4] int = SpecAugment
       (f0, fw, t0, tw) applied in the dB domain before the normalisation.
    """
    if not waves.is_cuda:
        raise _lib.LasrError("logmel_batch needs CUDA tensors (no CPU fallback)")
    waves = waves.contiguous().float()
    N, S_max = waves.shape
    ns_host = num_samples.cpu() if isinstance(num_samples, torch.Tensor) else torch.tensor(num_samples)
    st_host = None
    if starts is not None:
        st_host = (starts.cpu() if isinstance(starts, torch.Tensor) else torch.tensor(starts)).long()
        if int(st_host.min()) < 0 or int((st_host + ns_host.long()).max()) > S_max:
            raise _lib.LasrError("starts + num_samples must stay inside the waveform")
    if int(ns_host.max()) > S_max or int(ns_host.min()) <= N_FFT // 2:
        raise _lib.LasrError("num_samples must be in (256, S_max] (reflect padding needs more than n_fft/2 samples)")
    T_max = num_frames(int(ns_host.max()))
    ns = ns_host.to(device=waves.device, dtype=torch.int32)
    dev = waves.device
    basis, mel_idx, mel_w = constants(dev, sr)
    Lp = _lib.load().lasr_logmel_padded_len(T_max)
    parts = torch.empty((3, N, Lp), device=dev, dtype=torch.bfloat16)
    st = st_host.to(device=dev, dtype=torch.int32) if st_host is not None else None
    _lib.call("lasr_logmel_prepare_crop", waves, dither, st, ns, parts, N, S_max, T_max)
    db = torch.empty((N, T_max, N_MELS), device=dev, dtype=torch.float32)
    stats = torch.zeros((N, 2), device=dev, dtype=torch.float64)
    _lib.call("lasr_logmel_fwd", parts, basis, mel_idx, mel_w, ns, db, stats, N, T_max, products)
    if bands is not None:
        b = (bands if isinstance(bands, torch.Tensor) else torch.tensor(bands)).to(device=dev, dtype=torch.int32)
        if tuple(b.shape) != (N, 4):
            raise _lib.LasrError("bands must be [N, 4] = (f0, fw, t0, tw) per utterance")
        _lib.call("lasr_spec_augment", db, stats, ns, b.contiguous(), N, T_max)
    out_nct = torch.empty((N, 1, N_MELS, T_max), device=dev, dtype=torch.float32) if want_nct else None
    out_ntc = torch.empty((N, T_max, N_MELS), device=dev, dtype=out_dtype) if out_dtype is not None else None
    _lib.call("lasr_logmel_normalize", db, stats, ns, out_nct, out_ntc, N, T_max,
              _lib.dtype_code(out_dtype) if out_dtype is not None else _lib.LASR_F32)
    frames = num_frames(ns_host.long())
    return {"inputs": out_nct, "ntc": out_ntc, "percents": frames.float() / float(T_max), "frames": frames.int(),
            "db": db}


class DeviceFrontend:
    """waveforms resident on the device -> channels-last log-mel features, all buffers static (CUDA-graph capturable):

        feats [N, T_max, 64] out_dtype, percents [N] fp32 = DeviceFrontend(...)(waves [N, S_max], num_samples [N] int32)

    waves: int16 PCM (x / 32768, what torchaudio.load(normalize=True) yields, data_module.py:153) or fp32 in [-1, 1].
    T_max = frames of S_max samples: the batch is padded to the bucket length, percents[n] = T_n / T_max (:244).
    augment=True: train-time crop + SpecAugment drawn on the device (lasr_augment_draw: the reference's arithmetic, Philox
    keyed by seed + the bank's step counter, or `uniforms` [N, 6] fp64 supplied per call = the parity hook);
    dither=True: y += 1e-5 * N(0, 1) drawn in the prepare kernel (:155)."""

    def __init__(self, N, S_max, device, out_dtype, augment=False, dither=False, seed=0x6C617372, seed_dev=None):
        self.N, self.S_max, self.dev, self.out_dtype = N, S_max, torch.device(device), out_dtype
        self.augment, self.dither, self.seed, self.seed_dev = augment, dither, int(seed), seed_dev
        self.T_max = num_frames(S_max)
        self.basis, self.mel_idx, self.mel_w = constants(self.dev)
        Lp = _lib.load().lasr_logmel_padded_len(self.T_max)
        dev = self.dev
        self.parts = torch.empty((3, N, Lp), device=dev, dtype=torch.bfloat16)
        self.db = torch.empty((N, self.T_max, N_MELS), device=dev, dtype=torch.float32)
        self.stats = torch.zeros((N, 2), device=dev, dtype=torch.float64)
        self.feats = torch.empty((N, self.T_max, N_MELS), device=dev, dtype=out_dtype)
        self.starts = torch.zeros((N,), device=dev, dtype=torch.int32)
        self.kept = torch.zeros((N,), device=dev, dtype=torch.int32)
        self.bands = torch.zeros((N, 4), device=dev, dtype=torch.int32)
        self.percents = torch.zeros((N,), device=dev, dtype=torch.float32)

    def __call__(self, waves, num_samples, uniforms=None):
        if waves.dtype not in (torch.int16, torch.float32) or tuple(waves.shape) != (self.N, self.S_max):
            raise _lib.LasrError("DeviceFrontend: waves must be int16 or float32 [N, S_max] on the device")
        if not waves.is_cuda or not waves.is_contiguous() or num_samples.dtype != torch.int32:
            raise _lib.LasrError("DeviceFrontend: contiguous CUDA waves and int32 num_samples expected")
        N, T = self.N, self.T_max
        aug = 1 if self.augment else 0
        _lib.call("lasr_augment_draw", num_samples, uniforms, self.seed, self.seed_dev if uniforms is None else None,
                  self.starts, self.kept, self.bands, self.percents, N, T, aug, aug)
        self.stats.zero_()
        wave_code = 1 if waves.dtype == torch.int16 else 0
        _lib.call("lasr_logmel_prepare_wave", waves, wave_code, None, (self.seed ^ 0x5DEECE66D) if self.dither else 0,
                  self.seed_dev, self.starts if self.augment else None, self.kept, self.parts, N, self.S_max, T)
        _lib.call("lasr_logmel_fwd", self.parts, self.basis, self.mel_idx, self.mel_w, self.kept, self.db, self.stats, N, T,
                  6)
        if self.augment:
            _lib.call("lasr_spec_augment", self.db, self.stats, self.kept, self.bands, N, T)
        _lib.call("lasr_logmel_normalize", self.db, self.stats, self.kept, None, self.feats, N, T,
                  _lib.dtype_code(self.out_dtype))
        return self.feats, self.percents


def pcm16(waves):
    """float waveform in [-1, 1] -> 16-bit PCM (the wire format: x * 32768 rounded, saturated)."""
    return (waves.float() * 32768.0).round().clamp_(-32768, 32767).to(torch.int16)


def _load_wav(path_or_file):
    """PCM wav -> float32 [1, S] in [-1, 1) and its sample rate (torchaudio.load(..., normalize=True) contract,
    data_module.py:153).  Decoding is I/O, outside the hot path: stdlib `wave` covers the 16-bit PCM files that
    scripts/get_libri.py / get_aishell.py produce."""
    import wave

    import numpy as np

    with wave.open(path_or_file, "rb") as wf:
        sr, ch, sw, n = wf.getframerate(), wf.getnchannels(), wf.getsampwidth(), wf.getnframes()
        raw = wf.readframes(n)
    if sw != 2:
        raise _lib.LasrError("only 16-bit PCM wav files are supported")
    a = np.frombuffer(raw, dtype="<i2").reshape(-1, ch).T.astype(np.float32) / 32768.0
    return torch.from_numpy(a.copy()), sr


class AudioParser:
    """data_module.py:58-174.  parse_audio(audio_path, mask=False) -> [1, 64, T] fp32 (on the GPU).
    `audio_path` may be a path / file-like (as in the reference) or an already decoded waveform tensor [S] / [1, S].
    mask=True applies the train-time crop (:158-159) and SpecAugment (:163-165) with draws from `self.random`, an
    unseeded random.Random() exactly like the reference's (:65); assign a seeded one for reproducible runs."""

    def __init__(self, win_len=0.02, sr=16000, device="cuda", dither=True):
        import random

        self.sr = sr
        self.win_len = win_len  # kept for signature parity; the reference hard-codes 320/160 too (:68-70)
        self.device = torch.device(device)
        self.dither = dither
        self.random = random.Random()  # :65

    def parse_audio(self, audio_path, mask=False):
        if isinstance(audio_path, torch.Tensor):
            y = audio_path.reshape(1, -1)
        else:
            import os

            if isinstance(audio_path, str) and not os.path.exists(audio_path):
                raise FileExistsError("audio_path not exits")  # data_module.py:151-152 (sic)
            y, _ = _load_wav(audio_path)
            y = y[:1]
        y = y.to(self.device, torch.float32)
        d = torch.randn_like(y) if self.dither else None  # :155
        if mask:
            start, kept, bands = draw_augment(y.shape[1], self.random)
            out = logmel_batch(y, [kept], dither=d, sr=self.sr, starts=[start], bands=[list(bands)])
        else:
            out = logmel_batch(y, [y.shape[1]], dither=d, sr=self.sr)
        return out["inputs"][0]


def collate(batch):
    """data_module.py:222-248.  batch: list of (feats [1,64,T_i], token ids, path) -> the training-step tuple."""
    longest = max(batch, key=lambda s: s[0].size(2))[0]
    freq, max_t = longest.size(1), longest.size(2)
    max_s = len(max(batch, key=lambda s: len(s[1]))[1])
    n = len(batch)
    inputs = torch.zeros(n, 1, freq, max_t, device=longest.device)
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
