"""Isolate the head (last_cnn2 -> decoder -> log_softmax -> CTC) on the real activation distribution:
oracle fp64 / fp32 vs our fp32 kernels, fed with the SAME encoder activation (block5 output of the fp64 run)."""
import copy, sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import lightning_asr_b200.quartznet as q
from lightning_asr_b200.functions import Conv1x1BNReLUFn, DecoderLogSoftmaxFn
from lightning_asr_b200.ctc import CTCLoss
from oracle import quartznet_oracle as qo


def rel_err(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


labels = [" ", "'"] + [chr(ord("a") + i) for i in range(26)]
torch.manual_seed(0)
model = q.MyModel2(labels, mask=True, precision="fp32")
sd0 = copy.deepcopy(model.state_dict())
N, T = 4, 301
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 1, 64, T, generator=g); p = torch.linspace(0.6, 1.0, N)
Tp = (T - 1) // 2 + 1
t_len = torch.mul(Tp, p).int(); tgt_len = (t_len // 4).int()
targets = torch.randint(0, 28, (N, int(tgt_len.max())))
sd64 = {k: (v.detach().clone().double() if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
taps = {}
with torch.no_grad():
    qo.model(x.double(), p, sd64, mask=True, training=True, taps=taps)
h64 = taps["block5"].detach()  # [N, 512, T']


def head(dtype):
    sd = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    h = h64.clone().to(dtype).requires_grad_(True)
    y = F.conv1d(h, sd["encoder.last_cnn2.0.weight"])
    y = torch.relu(qo.batch_norm(y, sd, "encoder.last_cnn2.1", True, False))
    y.retain_grad()
    logits = F.conv1d(y, sd["decoder.weight"], sd["decoder.bias"])
    lp = F.log_softmax(logits.transpose(1, 2), dim=-1)
    lp.retain_grad()
    nll = F.ctc_loss(lp.transpose(0, 1), targets, t_len, tgt_len, blank=28, reduction="none")
    nll.mean().backward()
    return dict(h=h.grad, y=y.grad, lp=lp.grad, w=sd["encoder.last_cnn2.0.weight"].grad,
                g=sd["encoder.last_cnn2.1.weight"].grad, b=sd["encoder.last_cnn2.1.bias"].grad,
                dw=sd["decoder.weight"].grad, db=sd["decoder.bias"].grad)


r64, r32 = head(torch.float64), head(torch.float32)
model = model.cuda().train()
enc = model.encoder
hg = h64.detach().clone().float().cuda().transpose(1, 2).contiguous().requires_grad_(True)
y = Conv1x1BNReLUFn.apply(hg, enc.last_cnn2[0].weight, enc.last_cnn2[1].weight, enc.last_cnn2[1].bias,
                          q._bn_buffers(enc.last_cnn2[1]), True, True)
y.retain_grad()
lp = DecoderLogSoftmaxFn.apply(y, model.decoder.weight, model.decoder.bias)
lp.retain_grad()
nll = CTCLoss(blank=28, reduction="none")(lp.transpose(0, 1), targets.cuda(), t_len.cuda(), tgt_len.cuda())
nll.mean().backward()
ours = dict(h=hg.grad.transpose(1, 2), y=y.grad.transpose(1, 2), lp=lp.grad, w=enc.last_cnn2[0].weight.grad,
            g=enc.last_cnn2[1].weight.grad, b=enc.last_cnn2[1].bias.grad, dw=model.decoder.weight.grad,
            db=model.decoder.bias.grad)
for k in ours:
    print(f"{k:3s} ours {rel_err(ours[k], r64[k]):.2e}  torch32 {rel_err(r32[k], r64[k]):.2e}")
# structure of the BN-backward cancellation at this layer
yg = r64["y"]
print("d(last_cnn2 out): |mean_t| / rms  =", (yg.mean(dim=(0, 2)).abs().mean() / yg.pow(2).mean().sqrt()).item())
