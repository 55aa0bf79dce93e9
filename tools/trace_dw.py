"""Timeline of the tensor-core depthwise conv (debug hook in dwconv_tc.cu): %globaltimer stamps per role and item."""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from lightning_asr_b200 import _lib, ops
_lib.require_device()
lib = _lib.load()
lib.lasr_debug_set_dw_trace.argtypes = [ctypes.c_void_p]
lib.lasr_debug_set_dw_trace.restype = None
N, T = 32, 801
names = ["prod:start", "prod:stored", "mma:acc_free", "mma:series_full", "mma:issued", "epi:tmem_full", "epi:done",
         "cta:entry/prologue_done/pdl_wait_done/roles_done"]
for c, k in [(512, 63), (256, 33)]:
    x = torch.randn(N, T, c, device="cuda").bfloat16()
    w = torch.randn(c, 1, k, device="cuda") / k ** 0.5
    y = ops.dwconv_fwd(x, w)
    trace = torch.zeros(148 * 8 * 16, device="cuda", dtype=torch.int64)
    torch.cuda.synchronize()
    lib.lasr_debug_set_dw_trace(trace.data_ptr())
    y = ops.dwconv_fwd(x, w)
    torch.cuda.synchronize()
    lib.lasr_debug_set_dw_trace(None)
    tr = trace.cpu().view(148, 8, 16)
    t0 = int(tr[:, 7, 0][tr[:, 7, 0] > 0].min()) if int(tr[:, 7, 0].max()) > 0 else int(tr[:, 0, 0][tr[:, 0, 0] > 0].min())
    print(f"=== C={c} k={k}: last epilogue end {int(tr[:, 6].max()) - t0} ns")
    print("   CTA entry spread (ns):", int(tr[:, 7, 0][tr[:, 7, 0] > 0].max()) - t0 if int(tr[:, 7, 0].max()) > 0 else None)
    for cta in (0, 77):
        print(f"-- CTA {cta}")
        for slot, nm in enumerate(names):
            vals = [int(v) - t0 for v in tr[cta, slot] if int(v) > 0]
            print(f"   {nm:18s} {vals}")
