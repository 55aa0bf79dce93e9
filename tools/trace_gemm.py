"""Timeline of the weight-stationary GEMM: %globaltimer stamps per CTA role and tile (debug hook in gemm_tc.cu)."""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from lightning_asr_b200 import _lib, ops
_lib.require_device()
lib = _lib.load()
lib.lasr_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
lib.lasr_debug_set_gemm_trace.restype = None
M, T, N = 32 * 801, 801, 32
names = ["k:start/setup/Wready/end", "mma:acc_free", "mma:first_A", "mma:commit", "epi:tmem_full", "epi:stg_free", "epi:staged"]
for cin, cout in [(256, 256), (512, 512)]:
    x = torch.randn(M, cin, device="cuda").bfloat16()
    w = (torch.randn(cout, cin, device="cuda") / cin ** 0.5).bfloat16()
    lengths = torch.full((N,), T, device="cuda", dtype=torch.int32)
    stats = torch.zeros(2, cout, device="cuda", dtype=torch.float64)
    y = ops.pwconv_fwd(x, w, lengths=lengths, T=T, stats=stats)
    trace = torch.zeros(148 * 8 * 16, device="cuda", dtype=torch.int64)
    torch.cuda.synchronize()
    lib.lasr_debug_set_gemm_trace(trace.data_ptr())
    y = ops.pwconv_fwd(x, w, lengths=lengths, T=T, stats=stats)
    torch.cuda.synchronize()
    lib.lasr_debug_set_gemm_trace(None)
    tr = trace.cpu().view(148, 8, 16)
    t0 = int(tr[:, 0, 0][tr[:, 0, 0] > 0].min())
    print(f"=== {cin}->{cout}: kernel start skew over CTAs: {int(tr[:,0,0][tr[:,0,0]>0].max()) - t0} ns; last CTA end {int(tr[:,0,3].max()) - t0} ns")
    for cta in (0, 1, 100):
        print(f"-- CTA {cta}")
        for slot, nm in enumerate(names):
            vals = [int(v) - t0 for v in tr[cta, slot] if int(v) > 0]
            print(f"   {nm:26s} {vals}")
