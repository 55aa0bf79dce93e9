import sys, torch
sys.path.insert(0, ".")
import lightning_asr_b200.quartznet as q
from oracle import quartznet_oracle as qo
def rel_err(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
for (cin, cout, k) in [(256, 256, 33), (256, 256, 39), (256, 512, 51), (512, 512, 51), (512, 512, 63), (512, 512, 75)]:
    torch.manual_seed(k)
    blk = q.QuartNetBlock(repeat=1, in_ch=cin, out_ch=cout, k=k, mask=True).cuda().train()
    N, T = 4, 151
    x = torch.relu(torch.randn(N, cin, T))
    p = torch.tensor([1.0, 0.9, 0.77, 0.6])
    dout = torch.randn(N, cout, T)
    res = {}
    for dtype in (torch.float64, torch.float32):
        sd = {"b." + k_: (v.detach().to(dtype).cpu() if v.is_floating_point() else v.cpu()) for k_, v in blk.state_dict().items()}
        for v in sd.values():
            if v.is_floating_point(): v.requires_grad_(True)
        xr = x.to(dtype).requires_grad_(True)
        ref = qo.block(xr, p, sd, "b", mask=True, training=True, update_buffers=False)
        ref.backward(dout.to(dtype))
        res[dtype] = (ref.detach(), xr.grad, sd)
    xg = x.cuda().transpose(1, 2).contiguous().detach().requires_grad_(True)
    out = blk(xg, torch.mul(T, p).int().cuda())
    out.backward(dout.cuda().transpose(1, 2).contiguous())
    r64, r32 = res[torch.float64], res[torch.float32]
    print(f"block {cin}->{cout} k{k}: out ours {rel_err(out.transpose(1,2), r64[0]):.1e} t32 {rel_err(r32[0], r64[0]):.1e} | dx ours {rel_err(xg.grad.transpose(1,2), r64[1]):.1e} t32 {rel_err(r32[1], r64[1]):.1e}")
    for name, prm in blk.named_parameters():
        print(f"    {name:32s} ours {rel_err(prm.grad, r64[2]['b.'+name].grad):.1e} t32 {rel_err(r32[2]['b.'+name].grad, r64[2]['b.'+name].grad):.1e}")
