"""ncu driver: a few pointwise-conv GEMM launches at config-2 shapes."""
import sys
import torch
sys.path.insert(0, ".")
from lightning_asr_b200 import _lib, ops
_lib.require_device()
M, T, N = 32 * 801, 801, 32
for cin, cout in [(256, 256), (512, 512)]:
    x = torch.randn(M, cin, device="cuda").bfloat16()
    w = (torch.randn(cout, cin, device="cuda") / cin ** 0.5).bfloat16()
    lengths = torch.full((N,), T, device="cuda", dtype=torch.int32)
    stats = torch.zeros(2, cout, device="cuda", dtype=torch.float64)
    for _ in range(2):
        y = ops.pwconv_fwd(x, w, lengths=lengths, T=T, stats=stats)
    dx = ops.pwconv_dgrad(y, w)
torch.cuda.synchronize()
print("ok")
