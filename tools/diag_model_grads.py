import copy, sys, torch
sys.path.insert(0, ".")
def rel_err(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
import lightning_asr_b200.quartznet as q
from lightning_asr_b200.ctc import CTCLoss
from oracle import quartznet_oracle as qo
labels = [" ", "'"] + [chr(ord("a") + i) for i in range(26)]
torch.manual_seed(0)
model = getattr(q, sys.argv[1] if len(sys.argv) > 1 else "MyModel2")(labels, mask=True, precision="fp32")
sd0 = copy.deepcopy(model.state_dict())
model = model.cuda().train()
N, T = 4, 301
g = torch.Generator().manual_seed(0)
x = torch.randn(N, 1, 64, T, generator=g); p = torch.linspace(0.6, 1.0, N)
Tp = (T - 1) // 2 + 1
t_len = torch.mul(Tp, p).int(); tgt_len = (t_len // 4).int()
targets = torch.randint(0, 28, (N, int(tgt_len.max())))
def run(dtype):
    sd = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
    for v in sd.values():
        if v.is_floating_point(): v.requires_grad_(True)
    taps = {}
    out = qo.model(x.to(dtype), p, sd, mask=True, training=True, taps=taps)
    out.retain_grad()
    nll = torch.nn.functional.ctc_loss(out.transpose(0, 1), targets, t_len, tgt_len, blank=28, reduction="none")
    nll.mean().backward()
    return out, nll.detach(), sd
out64, nll64, sd64 = run(torch.float64)
out32, nll32, sd32 = run(torch.float32)
out = model(x.cuda(), p.cuda()); out.retain_grad()
nll = CTCLoss(blank=28, reduction="none")(out.transpose(0, 1), targets.cuda(), t_len.cuda(), tgt_len.cuda())
nll.mean().backward()
print("out", rel_err(out, out64), rel_err(out32, out64), "nll", rel_err(nll, nll64), rel_err(nll32, nll64))
print("dlp ours vs 64", rel_err(out.grad, out64.grad), " torch32 vs 64", rel_err(out32.grad, out64.grad))
for name, prm in model.named_parameters():
    print(f"{name:50s} ours {rel_err(prm.grad, sd64[name].grad):.2e} torch32 {rel_err(sd32[name].grad, sd64[name].grad):.2e}")
