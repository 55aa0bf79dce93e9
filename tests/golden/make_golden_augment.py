"""Golden fixture for the train-time augmentation path: the REFERENCE's own
AudioParser.parse_audio(mask=True) (data_module.py:150-174 -> sub_secquence :138-148, spec_augment :97-122), imported
from /root/reference with the control-plane packages stubbed (see make_golden.import_reference), run with its random
draws made reproducible: np.random.uniform and AudioParser.rand are replaced by one seeded random.Random, dither off.

    python tests/golden/make_golden_augment.py      (build container only)
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from golden_common import seeded_wave  # noqa: E402
from make_golden import import_reference  # noqa: E402


def main():
    import torchaudio

    data_module, _, _ = import_reference()
    parser = data_module.AudioParser()
    fx = {}
    real_load, real_randn_like, real_uniform = torchaudio.load, torch.randn_like, np.random.uniform
    try:
        torch.randn_like = lambda t, *a, **k: torch.zeros_like(t)
        for name, (n, seed) in {"a20000": (20000, 4), "a48000": (48000, 5)}.items():
            w = seeded_wave(n, seed)
            rng = random.Random(900 + seed)
            np.random.uniform = lambda a, b, _r=rng: _r.uniform(a, b)
            parser.rand = rng
            torchaudio.load = lambda *_a, _w=w, **_k: (_w.clone().reshape(1, -1), 16000)
            fx[name] = {"samples": n, "seed": seed, "rng_seed": 900 + seed,
                        "features": parser.parse_audio(object(), mask=True).clone()}
            print(name, tuple(fx[name]["features"].shape))
    finally:
        torchaudio.load, torch.randn_like, np.random.uniform = real_load, real_randn_like, real_uniform
    torch.save(fx, os.path.join(HERE, "augment.pt"))


if __name__ == "__main__":
    main()
