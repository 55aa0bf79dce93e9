"""Accuracy of the BiLSTM kernels vs an fp64 run of torch's LSTM, next to torch's own fp32 run (same inputs)."""
import sys
import torch
sys.path.insert(0, ".")
from lightning_asr_b200.functions import BiLstmFn

def rel(a, b):
    return float((a.detach().double() - b.detach().double()).norm() / b.detach().double().norm())

torch.manual_seed(3)
N, T, Cin, H = 8, 501, 256, 40
lens = torch.tensor([501, 1, 200, 360, 90, 480, 333, 501], dtype=torch.int32)
ref = torch.nn.LSTM(Cin, H, num_layers=1, batch_first=True, bidirectional=True)
x = torch.randn(N, T, Cin) * 0.5
gout = torch.randn(N, T, 2 * H)
names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0", "weight_ih_l0_reverse", "weight_hh_l0_reverse",
         "bias_ih_l0_reverse", "bias_hh_l0_reverse"]

def run_torch(dtype):
    m = torch.nn.LSTM(Cin, H, num_layers=1, batch_first=True, bidirectional=True).to(dtype)
    m.load_state_dict({k: v.to(dtype) for k, v in ref.state_dict().items()})
    xr = x.to(dtype).requires_grad_(True)
    pk = torch.nn.utils.rnn.pack_padded_sequence(xr, lens.long(), batch_first=True, enforce_sorted=False)
    y, _ = m(pk)
    y, _ = torch.nn.utils.rnn.pad_packed_sequence(y, batch_first=True, total_length=T)
    (y * gout.to(dtype)).sum().backward()
    return y.detach(), xr.grad, [getattr(m, n).grad for n in names]

y64, dx64, g64 = run_torch(torch.float64)
y32, dx32, g32 = run_torch(torch.float32)
params = [torch.nn.Parameter(getattr(ref, n).detach().clone().cuda()) for n in names]
xg = x.cuda().detach().requires_grad_(True)
y = BiLstmFn.apply(xg, lens.cuda(), *params)
(y * gout.cuda()).sum().backward()
print(f"{'tensor':26s} {'ours vs f64':>12s} {'torch f32 vs f64':>18s}")
print(f"{'out':26s} {rel(y.cpu(), y64):12.3e} {rel(y32, y64):18.3e}")
print(f"{'dx':26s} {rel(xg.grad.cpu(), dx64):12.3e} {rel(dx32, dx64):18.3e}")
for n, p, a, b in zip(names, params, g32, g64):
    print(f"{n:26s} {rel(p.grad.cpu(), b):12.3e} {rel(a, b):18.3e}")
