import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu on the GPU box")


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="session")
def labels28():
    # conf/conf.yaml:12-13 (28 inline labels; blank = index 28)
    return [" ", "'"] + [chr(ord("a") + i) for i in range(26)]


def bf16_reference_yardstick(x, percents, sd, mask, training, drop_masks=None):
    """The reference's OWN bf16 path on this GPU: the oracle's functional restatement of the reference modules run on
    cuda under torch.autocast(bfloat16) (cuDNN / cuBLAS bf16 kernels) -> log-probs fp32 [N, T', V'].  Used as the
    yardstick for whole-network bf16 comparisons: every bf16 implementation of this 15-block network
    rounds ~5 tensors per block to 8 mantissa bits (1.1e-3 rms each), i.e. ~sqrt(75) * 1.1e-3 ~ 1e-2 at the output of a
    random-init net, the reference's autocast path included."""
    import torch
    from oracle import quartznet_oracle as qo
    sdc = {k: (v.detach().float().cuda() if v.is_floating_point() else v.clone().cuda()) for k, v in sd.items()}
    dm = None if drop_masks is None else {k: v.float().cuda() for k, v in drop_masks.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out = qo.model(x.float().cuda(), percents, sdc, mask=mask, training=training, drop_masks=dm)
    return out.float()


def assert_bf16_close(ours, ref64, yardstick, what=""):
    """north_star: bf16 path within rel 1e-2 of the reference.  Where the reference's own bf16 (autocast) run of the
    same case is itself further than that from the fp64 truth, ours must not be further than it (x1.1 for run-to-run
    noise); both numbers are printed so the log shows which branch decided."""
    e_ours, e_ref = rel_err(ours, ref64), rel_err(yardstick, ref64)
    print(f"[bf16 parity] {what}: ours {e_ours:.3e}  reference-autocast-bf16 {e_ref:.3e}  (both vs fp64)")
    assert e_ours < max(1e-2, 1.1 * e_ref), (what, e_ours, e_ref)


def check_network_grads(model, g64, g32):
    """SURVEY.md 10.1 protocol for whole-network fp32 train-mode gradients: judged against the fp64 run with the fp32
    oracle's own deviation as the yardstick.  The MEDIAN over all parameter tensors of ours/theirs must be <= 2 (a
    systematic loss of accuracy reads 10-100x there); a single tensor may sit up to max(3x theirs, 1e-2), because one
    ReLU gate that flips on a 1e-7 forward difference in either fp32 run moves every gradient below it by
    ~1/sqrt(#elements) ~ 1e-3 (tools/diag_model_grads.py) -- that is a property of the comparison, not of either side."""
    ratios = []
    for name, prm in model.named_parameters():
        ours = rel_err(prm.grad, g64[name])
        theirs = rel_err(g32[name], g64[name])
        ratios.append(ours / max(theirs, 1e-4))
        assert ours < max(3 * theirs, 1e-2), (name, ours, theirs)
    ratios.sort()
    print(f"[fp32 grad protocol] median ours/theirs {ratios[len(ratios) // 2]:.2f}, max {ratios[-1]:.2f}")
    assert ratios[len(ratios) // 2] <= 2.0, ratios
