"""Edge cases of the hot path on the B200 kernels: empty / degenerate / ragged inputs and error behaviour
(the reference's own behaviour for the same inputs is the yardstick: torch ops in fp64 or the oracle)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from lightning_asr_b200 import _lib, ops as o
    _lib.require_device()
    return o


def test_ctc_empty_targets_and_single_frame(ops):
    """Empty transcript (loss = -sum log p(blank)), one-frame utterances, a one-label target on one frame."""
    torch.manual_seed(1)
    N, T, V = 4, 9, 29
    ld = 32
    logits = torch.zeros(N, T, ld, device="cuda")
    logits[:, :, :V] = torch.randn(N, T, V, device="cuda")
    in_len = torch.tensor([T, 1, 1, 5], device="cuda", dtype=torch.int32)
    tgt_len = torch.tensor([0, 0, 1, 2], device="cuda", dtype=torch.int32)
    targets = torch.randint(0, V - 1, (N, 2), device="cuda")
    lp_ref = F.log_softmax(logits[..., :V].double(), dim=-1).requires_grad_(True)
    nll_ref = F.ctc_loss(lp_ref.transpose(0, 1), targets, in_len.long(), tgt_len.long(), blank=V - 1, reduction="none")
    nll_ref.sum().backward()
    lse, _ = ops.log_softmax_fwd(logits, V, want_lp=False)
    nll, alpha, beta, scales = ops.ctc_fwd(logits, lse, targets, in_len, tgt_len, V, V - 1, want_beta=True)
    assert rel_err(nll, nll_ref) < 1e-5
    assert rel_err(nll[0], -lp_ref[0, :, V - 1].sum()) < 1e-5  # empty target = all blanks
    grad = ops.ctc_bwd(logits, lse, targets, in_len, tgt_len, alpha, beta, nll, torch.ones(N, device="cuda"), V, V - 1,
                       ld, torch.float32, scales=scales)
    assert rel_err(grad[..., :V], lp_ref.grad) < 5e-4
    assert torch.isfinite(grad).all()


def test_depthwise_shorter_than_kernel_and_single_frame(ops):
    """T < K (every tap window hangs over both edges) and T == 1, fp32 and bf16, forward + both gradients."""
    torch.manual_seed(2)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1e-2)):
        for T, K, C in ((20, 33, 256), (1, 51, 512), (7, 87, 512)):
            x = torch.randn(3, T, C, device="cuda").to(dtype)
            w = torch.randn(C, 1, K, device="cuda") / K ** 0.5
            xr = x.float().transpose(1, 2).contiguous().requires_grad_(True)
            wr = w.clone().requires_grad_(True)
            if dtype == torch.bfloat16:  # the tensor-core path rounds the taps to bf16
                wr = w.bfloat16().float().requires_grad_(True)
            yr = F.conv1d(xr, wr, padding=K // 2, groups=C)
            y = ops.dwconv_fwd(x, w)
            assert y.shape == (3, T, C)
            assert rel_err(y.float(), yr.transpose(1, 2)) < tol, (T, K, dtype)
            dy = torch.randn(3, T, C, device="cuda").to(dtype)
            yr.backward(dy.float().transpose(1, 2))
            dx = ops.dwconv_fwd(dy, w, flip=True)
            assert rel_err(dx.float(), xr.grad.transpose(1, 2)) < tol, (T, K, dtype)
            dw = ops.dwconv_wgrad(x, dy, K)
            assert rel_err(dw, wr.grad) < 3 * tol, (T, K, dtype)


def test_pointwise_tiny_and_fully_masked(ops):
    """M far below one 128-row tile, and a batch whose every frame is masked (MaskCNN with length 0)."""
    torch.manual_seed(3)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1e-2)):
        x = torch.randn(1, 5, 256, device="cuda").to(dtype)
        w = (torch.randn(512, 256, device="cuda") / 16).to(dtype)
        y = ops.pwconv_fwd(x, w)
        assert rel_err(y.float(), x.float() @ w.float().t()) < tol
        lengths = torch.zeros(1, device="cuda", dtype=torch.int32)
        stats = torch.zeros(2, 512, device="cuda", dtype=torch.float64)
        y0 = ops.pwconv_fwd(x, w, lengths=lengths, T=5, stats=stats)
        assert torch.count_nonzero(y0) == 0 and torch.count_nonzero(stats) == 0
        lengths = torch.tensor([3], device="cuda", dtype=torch.int32)
        y3 = ops.pwconv_fwd(x, w, lengths=lengths, T=5)
        assert torch.count_nonzero(y3[0, 3:]) == 0 and rel_err(y3[0, :3].float(), y[0, :3].float()) < 1e-6


def test_bilstm_zero_and_full_lengths():
    from lightning_asr_b200.functions import BiLstmFn

    torch.manual_seed(4)
    ref = torch.nn.LSTM(256, 40, batch_first=True, bidirectional=True).cuda()
    params = [getattr(ref, n) for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0",
                                         "weight_ih_l0_reverse", "weight_hh_l0_reverse", "bias_ih_l0_reverse",
                                         "bias_hh_l0_reverse")]
    x = torch.randn(3, 11, 256, device="cuda", requires_grad=True)
    lens = torch.tensor([0, 11, 4], device="cuda", dtype=torch.int32)
    y = BiLstmFn.apply(x, lens, *params)
    assert torch.count_nonzero(y[0]) == 0 and torch.count_nonzero(y[2, 4:]) == 0
    y.sum().backward()
    assert torch.isfinite(x.grad).all() and torch.count_nonzero(x.grad[0]) == 0 and torch.count_nonzero(x.grad[2, 4:]) == 0
    with torch.no_grad():  # fp32 reference on the CPU (cuDNN's LSTM would run its GEMMs in TF32)
        cpu = torch.nn.LSTM(256, 40, batch_first=True, bidirectional=True)
        cpu.load_state_dict({k: v.cpu() for k, v in ref.state_dict().items()})
        full, _ = cpu(x[1:2].detach().cpu())
    assert rel_err(y[1:2], full) < 1e-5


def test_logmel_shortest_utterance_and_rejections():
    from lightning_asr_b200 import _lib, frontend
    from oracle import frontend_oracle

    g = torch.Generator().manual_seed(5)
    w = 0.1 * torch.randn(257, generator=g)  # reflect padding needs > n_fft/2 samples: 257 is the minimum
    out = frontend.logmel_batch(w.reshape(1, -1).cuda(), [257])
    ref = frontend_oracle.logmel(w)
    assert out["inputs"].shape[-1] == ref.shape[-1] == 3
    assert rel_err(out["inputs"][0].cpu(), ref) < 1e-4
    with pytest.raises(_lib.LasrError):
        frontend.logmel_batch(w[:256].reshape(1, -1).cuda(), [256])
    with pytest.raises(_lib.LasrError):
        frontend.logmel_batch(w.reshape(1, -1), [257])  # CPU tensor: no fallback


def test_model_batch_with_a_very_short_utterance():
    """An utterance that covers 2 % of the padded batch (MaskCNN zeroes almost everything, CTC sees 3 frames)."""
    import lightning_asr_b200.quartznet as q
    from lightning_asr_b200.ctc import CTCLoss
    from oracle import quartznet_oracle as qo

    labels = [" ", "'"] + [chr(ord("a") + i) for i in range(26)]
    torch.manual_seed(0)
    model = q.MyModel2(labels, mask=True, precision="fp32")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().train()
    N, T = 3, 301
    x = torch.randn(N, 1, 64, T)
    p = torch.tensor([1.0, 0.02, 0.5])
    out = model(x.cuda(), p.cuda())
    ref = qo.model(x, p, sd, mask=True, training=True, update_buffers=False)
    assert rel_err(out, ref) < 1e-4
    Tp = out.shape[1]
    t_len = torch.mul(Tp, p).int()
    assert int(t_len[1]) == 3
    tgt_len = torch.tensor([20, 1, 10], dtype=torch.int32)
    targets = torch.randint(0, 28, (N, 20))
    nll = CTCLoss(blank=28, reduction="none")(out.transpose(0, 1), targets.cuda(), t_len.cuda(), tgt_len.cuda())
    nll_ref = F.ctc_loss(ref.transpose(0, 1), targets, t_len.long(), tgt_len.long(), blank=28, reduction="none")
    assert rel_err(nll, nll_ref) < 1e-4
    nll.mean().backward()
    assert all(torch.isfinite(p_.grad).all() for p_ in model.parameters())
